// decode_host.cu -- TEST INFRASTRUCTURE, not the product.
//
// Instantiates the device decoder of flow-pipeline_b200/csrc/decode.cuh (the template every kernel parses
// records with) as plain host code, so the CPU test suite can run the very same source over the golden
// vectors and fuzz sets and compare it with the oracle.  The byte source is the global-memory one (ByteSrc);
// the shared-memory source differs only in how a word is fetched.  Nothing in libflowagg.so links this.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../flow-pipeline_b200/csrc/decode.cuh"

extern "C" {

struct dh_flow {
    uint64_t time_received, sampling_rate, time_flow_start, bytes, packets;
    uint32_t type, sequence_num, src_as, dst_as, etype, proto, src_port, dst_port;
    uint32_t src_len, dst_len, sampler_len, pad;
    uint8_t src[16], dst[16], sampler[16];
};

}  // extern "C"

namespace {

void put_addr(const uint32_t w[4], uint8_t out[16])
{
    for (int k = 0; k < 4; k++) {
        out[4 * k + 0] = (uint8_t)(w[k] >> 24);
        out[4 * k + 1] = (uint8_t)(w[k] >> 16);
        out[4 * k + 2] = (uint8_t)(w[k] >> 8);
        out[4 * k + 3] = (uint8_t)w[k];
    }
}

template <uint32_t NEED>
void run(const fa::ByteSrc &src, const uint32_t *offsets, size_t n, bool framed, dh_flow *out, uint8_t *valid)
{
    for (size_t i = 0; i < n; i++) {
        fa::Flow f;
        fa::flow_reset(f);
        const uint32_t o0 = offsets[i], o1 = offsets[i + 1];
        const bool ok = o1 >= o0 && fa::decode_record<NEED>(src, o0, o1, framed, f);
        valid[i] = ok ? 1 : 0;
        dh_flow &d = out[i];
        memset(&d, 0, sizeof d);
        if (!ok) continue;
        d.time_received = f.time_received;
        d.sampling_rate = f.sampling_rate;
        d.time_flow_start = f.time_flow_start;
        d.bytes = f.bytes;
        d.packets = f.packets;
        d.type = f.type;
        d.sequence_num = f.sequence_num;
        d.src_as = f.src_as;
        d.dst_as = f.dst_as;
        d.etype = f.etype;
        d.proto = f.proto;
        d.src_port = f.src_port;
        d.dst_port = f.dst_port;
        d.src_len = f.src_len;
        d.dst_len = f.dst_len;
        d.sampler_len = f.sampler_len;
        put_addr(f.src, d.src);
        put_addr(f.dst, d.dst);
        put_addr(f.sampler, d.sampler);
    }
}

}  // namespace

// the NEED masks the kernels instantiate (kernels.cuh: KeyTraits<MODE>::NEED | values | weight)
static const uint32_t kNeed[] = {
    fa::F_ALL,
    fa::F_SRC_AS | fa::F_DST_AS | fa::F_BYTES | fa::F_PACKETS,                                            // aspair
    fa::F_TIME_RECEIVED | fa::F_SRC_AS | fa::F_DST_AS | fa::F_ETYPE | fa::F_BYTES | fa::F_PACKETS,         // flows5m
    fa::F_SRC_ADDR | fa::F_BYTES | fa::F_PACKETS | fa::F_SAMPLING_RATE,                                   // srcaddr, weighted
    fa::F_SRC_ADDR | fa::F_DST_ADDR | fa::F_SRC_PORT | fa::F_DST_PORT | fa::F_PROTO | fa::F_BYTES | fa::F_PACKETS,  // 5tuple
    fa::F_DST_PORT | fa::F_BYTES | fa::F_PACKETS,                                                         // dstport
};

extern "C" int dh_need_count(void) { return (int)(sizeof kNeed / sizeof kNeed[0]); }
extern "C" uint32_t dh_need_mask(int sel) { return kNeed[sel]; }

// buf[0,len): the records; offsets[n+1] relative to buf.  Returns 0, or -1 for a bad selector.
extern "C" int dh_decode(const uint8_t *buf, size_t len, const uint32_t *offsets, size_t n, int framed, int sel, dh_flow *out,
                         uint8_t *valid)
{
    std::vector<uint32_t> words(len / 4 + 8, 0u);
    if (len) memcpy(words.data(), buf, len);
    fa::ByteSrc src;
    src.words = words.data();
    src.limit_word = (uint32_t)(words.size() - 1);
    switch (sel) {
    case 0: run<kNeed[0]>(src, offsets, n, framed != 0, out, valid); break;
    case 1: run<kNeed[1]>(src, offsets, n, framed != 0, out, valid); break;
    case 2: run<kNeed[2]>(src, offsets, n, framed != 0, out, valid); break;
    case 3: run<kNeed[3]>(src, offsets, n, framed != 0, out, valid); break;
    case 4: run<kNeed[4]>(src, offsets, n, framed != 0, out, valid); break;
    case 5: run<kNeed[5]>(src, offsets, n, framed != 0, out, valid); break;
    default: return -1;
    }
    return 0;
}
