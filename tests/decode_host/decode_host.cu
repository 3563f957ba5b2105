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

// tags: nullptr = the order-agnostic decoder alone; otherwise the shape fast path over that table first
// (the kernels' arrangement), *n_fast counting the records it took.
template <uint32_t NEED>
void run(const fa::ByteSrc &src, const uint32_t *offsets, size_t n, bool framed, dh_flow *out, uint8_t *valid,
         const uint16_t *tags = nullptr, uint32_t n_tags = 0, uint64_t *n_fast = nullptr)
{
    fa::ShapeTable sh;
    fa::shape_build(tags, tags ? n_tags : 0, NEED, sh);
    for (size_t i = 0; i < n; i++) {
        fa::Flow f;
        fa::flow_reset(f);
        const uint32_t o0 = offsets[i], o1 = offsets[i + 1];
        bool ok = false;
        if (o1 >= o0) {
            if (tags && i + 1 < n && offsets[i + 2] >= o1) {
                // the kernels' two-records-side-by-side walk: record i together with its successor, keep i's result
                uint32_t p2[2] = {o0, o1};
                const uint32_t e2[2] = {o1, offsets[i + 2]};
                fa::Flow f2[2];
                fa::flow_reset(f2[0]);
                fa::flow_reset(f2[1]);
                bool t2[2];
                fa::decode_records_shape<NEED, 2>(sh, src, p2, e2, framed, f2, t2);
                ok = t2[0];
                if (ok) f = f2[0];
            } else {
                ok = tags && fa::decode_record_shape<NEED>(sh, src, o0, o1, framed, f);
            }
            if (ok && n_fast) ++*n_fast;
            if (!ok) {
                fa::flow_reset(f);
                ok = fa::decode_record<NEED>(src, o0, o1, framed, f);
            }
        }
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

struct MarkSeen {
    uint8_t *seen;  // per tag value: 1 seen | 2 a 1..4-byte varint | 4 a 5-byte varint
    __host__ __device__ void operator()(uint32_t tagval, uint32_t vb) const
    {
        seen[tagval & 0x3fffu] |= 1 | (vb >= 1 && vb <= 4 ? 2 : 0) | (vb == 5 ? 4 : 0);
    }
};

static int dispatch(const uint8_t *buf, size_t len, const uint32_t *offsets, size_t n, int framed, int sel, dh_flow *out, uint8_t *valid,
                    const uint16_t *tags, uint32_t n_tags, uint64_t *n_fast)
{
    std::vector<uint32_t> words(len / 4 + 8, 0u);
    if (len) memcpy(words.data(), buf, len);
    fa::ByteSrc src;
    src.words = words.data();
    src.limit_word = (uint32_t)(words.size() - 1);
    switch (sel) {
    case 0: run<kNeed[0]>(src, offsets, n, framed != 0, out, valid, tags, n_tags, n_fast); break;
    case 1: run<kNeed[1]>(src, offsets, n, framed != 0, out, valid, tags, n_tags, n_fast); break;
    case 2: run<kNeed[2]>(src, offsets, n, framed != 0, out, valid, tags, n_tags, n_fast); break;
    case 3: run<kNeed[3]>(src, offsets, n, framed != 0, out, valid, tags, n_tags, n_fast); break;
    case 4: run<kNeed[4]>(src, offsets, n, framed != 0, out, valid, tags, n_tags, n_fast); break;
    case 5: run<kNeed[5]>(src, offsets, n, framed != 0, out, valid, tags, n_tags, n_fast); break;
    default: return -1;
    }
    return 0;
}

// buf[0,len): the records; offsets[n+1] relative to buf.  Returns 0, or -1 for a bad selector.
extern "C" int dh_decode(const uint8_t *buf, size_t len, const uint32_t *offsets, size_t n, int framed, int sel, dh_flow *out,
                         uint8_t *valid)
{
    return dispatch(buf, len, offsets, n, framed, sel, out, valid, nullptr, 0, nullptr);
}

// What k_learn_shape does, on the host: the ascending, de-duplicated tag values of the first n_sample records
// (every `stride`-th one).  Returns the number of tags written (<= cap).
extern "C" uint32_t dh_learn_shape(const uint8_t *buf, size_t len, const uint32_t *offsets, size_t n, int framed, size_t stride, size_t n_sample,
                                   uint16_t *tags, uint32_t cap)
{
    std::vector<uint32_t> words(len / 4 + 8, 0u);
    if (len) memcpy(words.data(), buf, len);
    fa::ByteSrc src;
    src.words = words.data();
    src.limit_word = (uint32_t)(words.size() - 1);
    std::vector<uint8_t> seen(1u << 14, 0);
    if (!stride) stride = 1;
    for (size_t k = 0, i = 0; k < n_sample && i < n; k++, i += stride)
        if (offsets[i + 1] >= offsets[i])
            fa::shape_collect(src, offsets[i], offsets[i + 1], framed != 0, MarkSeen{seen.data()});
    uint32_t m = 0;
    for (uint32_t t = 0; t < (1u << 14); t++)
        if (seen[t] && m < cap) tags[m++] = (uint16_t)(t | ((seen[t] & 2) ? fa::kTagSaw4 : 0) | ((seen[t] & 4) ? fa::kTagSaw5 : 0));
    return m;
}

// The kernels' arrangement: shape fast path over `tags` (ascending tag values) first, the order-agnostic decoder
// for every record it does not decide.  *n_fast = records the fast path took.
extern "C" int dh_decode_shaped(const uint8_t *buf, size_t len, const uint32_t *offsets, size_t n, int framed, int sel, const uint16_t *tags,
                                uint32_t n_tags, dh_flow *out, uint8_t *valid, uint64_t *n_fast)
{
    static const uint16_t none = 0;
    uint64_t dummy = 0;
    if (n_fast) *n_fast = 0;
    return dispatch(buf, len, offsets, n, framed, sel, out, valid, tags ? tags : &none, n_tags, n_fast ? n_fast : &dummy);
}
