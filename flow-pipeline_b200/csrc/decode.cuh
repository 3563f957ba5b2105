// decode.cuh -- device-side proto3 decoder for flowprotob.FlowMessage.
//
// One thread decodes one record from a byte tile staged in shared memory (or,
// for tiles that do not fit, straight from global memory).  The rules are those
// of proto.Unmarshal as called at inserter/inserter.go:124 (golang/protobuf
// v1.4.3 -> google.golang.org/protobuf, go.mod:7) for the field table of
// pb-ext/flow.pb.go:58-143:
//   any field order; last value wins; bytes replaced; unknown numbers and known
//   numbers with a foreign wire type skipped by wire type; groups skipped with
//   matching end tags; uint32/enum keep the low 32 bits; varints <= 10 bytes with
//   the 10th byte <= 1; field number 1..2^29-1 (1..2^31-1 inside a skipped
//   group); wire types 6/7, stray end-group, truncation -> the record is bad
//   (skipped and counted, inserter.go:125-126); proto3 strings 100/101 must be
//   valid UTF-8.
//
// Hot path: an unaligned 64-bit window is assembled from three aligned 32-bit
// shared-memory loads; a 1-2 byte tag and a <=5-byte varint (or a 1-byte length)
// are decoded from that window without further loads or loops.  Everything else
// (long varints, long tags, groups, UTF-8) takes a byte-wise slow path kept out
// of line.
#pragma once
#include <stdint.h>

namespace fa {

// which fields a kernel needs; everything else is skipped and its code removed
enum : uint32_t {
    F_TYPE = 1u << 0,
    F_TIME_RECEIVED = 1u << 1,
    F_SAMPLING_RATE = 1u << 2,
    F_SEQUENCE_NUM = 1u << 3,
    F_SRC_ADDR = 1u << 4,
    F_DST_ADDR = 1u << 5,
    F_BYTES = 1u << 6,
    F_PACKETS = 1u << 7,
    F_SAMPLER_ADDR = 1u << 8,
    F_SRC_AS = 1u << 9,
    F_DST_AS = 1u << 10,
    F_PROTO = 1u << 11,
    F_SRC_PORT = 1u << 12,
    F_DST_PORT = 1u << 13,
    F_ETYPE = 1u << 14,
    F_TIME_FLOW_START = 1u << 15,
    F_ALL = 0xFFFFu
};

struct Flow {
    unsigned long long time_received, sampling_rate, time_flow_start, bytes, packets;
    uint32_t type, sequence_num, src_as, dst_as, etype, proto, src_port, dst_port;
    uint32_t src[4], dst[4], sampler[4];  // big-endian words, zero right-padded (FixedString(16))
    uint32_t src_len, dst_len, sampler_len;
};

#define FA_MAX_GROUP_DEPTH 32

// Byte source over aligned 32-bit words.  `words` may point to shared or global
// memory (the compiler keeps the address space after inlining).  limit_word is
// the last word index that may be touched (over-reads are clamped to it).
struct ByteSrc {
    const uint32_t *words;
    uint32_t limit_word;
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return words[i < limit_word ? i : limit_word]; }
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return (word(pos >> 2) >> ((pos & 3u) * 8u)) & 0xffu; }
    // bytes pos..pos+7 as little-endian (lo, hi)
    __device__ __forceinline__ void window(uint32_t pos, uint32_t &lo, uint32_t &hi) const
    {
        const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
        const uint32_t a = word(i), b = word(i + 1), c = word(i + 2);
        lo = __funnelshift_r(a, b, sh);
        hi = __funnelshift_r(b, c, sh);
    }
};

// byte source without clamping, for tiles resident in (padded) shared memory
struct SmemSrc {
    const uint32_t *words;
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return words[i]; }
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return (words[pos >> 2] >> ((pos & 3u) * 8u)) & 0xffu; }
    __device__ __forceinline__ void window(uint32_t pos, uint32_t &lo, uint32_t &hi) const
    {
        const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
        const uint32_t a = words[i], b = words[i + 1], c = words[i + 2];
        lo = __funnelshift_r(a, b, sh);
        hi = __funnelshift_r(b, c, sh);
    }
};

// ---- slow paths (out of line) -------------------------------------------------

// protowire.ConsumeVarint, byte at a time.  Returns bytes consumed or 0 on error.
template <class Src>
__device__ __noinline__ uint32_t varint_slow(const Src &s, uint32_t pos, uint32_t end, unsigned long long &v)
{
    unsigned long long x = 0;
    for (uint32_t i = 0; i < 10; i++) {
        if (pos + i >= end) return 0;  // truncated
        const unsigned long long y = s.byte(pos + i);
        if (i == 9) {
            if (y >= 2) return 0;  // overflow
            x |= y << 63;
            v = x;
            return 10;
        }
        x |= (y & 0x7f) << (7 * i);
        if (y < 0x80) {
            v = x;
            return i + 1;
        }
    }
    return 0;
}

template <class Src>
__device__ __noinline__ bool utf8_valid(const Src &s, uint32_t pos, uint32_t n)
{
    uint32_t i = 0;
    while (i < n) {
        const uint32_t c = s.byte(pos + i);
        if (c < 0x80) {
            i++;
            continue;
        }
        uint32_t need, lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) need = 1;
        else if (c == 0xE0) { need = 2; lo = 0xA0; }
        else if (c >= 0xE1 && c <= 0xEC) need = 2;
        else if (c == 0xED) { need = 2; hi = 0x9F; }
        else if (c >= 0xEE && c <= 0xEF) need = 2;
        else if (c == 0xF0) { need = 3; lo = 0x90; }
        else if (c >= 0xF1 && c <= 0xF3) need = 3;
        else if (c == 0xF4) { need = 3; hi = 0x8F; }
        else return false;
        if (i + need >= n) return false;
        const uint32_t c1 = s.byte(pos + i + 1);
        if (c1 < lo || c1 > hi) return false;
        for (uint32_t k = 2; k <= need; k++) {
            const uint32_t ck = s.byte(pos + i + k);
            if (ck < 0x80 || ck > 0xBF) return false;
        }
        i += need + 1;
    }
    return true;
}

// protowire.ConsumeFieldValue(StartGroupType): returns the position after the
// matching end-group tag, or 0xFFFFFFFF on error.
template <class Src>
__device__ __noinline__ uint32_t skip_group(const Src &s, uint32_t pos, uint32_t end, uint32_t start_num)
{
    uint32_t stack[FA_MAX_GROUP_DEPTH];
    int depth = 0;
    stack[depth++] = start_num;
    while (depth > 0) {
        unsigned long long tag, v;
        uint32_t n = varint_slow(s, pos, end, tag);
        if (!n) return 0xFFFFFFFFu;
        pos += n;
        const unsigned long long num = tag >> 3;
        const uint32_t wt = (uint32_t)tag & 7u;
        if (num < 1 || num > 0x7fffffffull) return 0xFFFFFFFFu;
        switch (wt) {
        case 0:
            n = varint_slow(s, pos, end, v);
            if (!n) return 0xFFFFFFFFu;
            pos += n;
            break;
        case 1:
            if (end - pos < 8) return 0xFFFFFFFFu;
            pos += 8;
            break;
        case 2:
            n = varint_slow(s, pos, end, v);
            if (!n) return 0xFFFFFFFFu;
            pos += n;
            if (v > (unsigned long long)(end - pos)) return 0xFFFFFFFFu;
            pos += (uint32_t)v;
            break;
        case 3:
            if (depth >= FA_MAX_GROUP_DEPTH) return 0xFFFFFFFFu;
            stack[depth++] = (uint32_t)num;
            break;
        case 4:
            if (stack[depth - 1] != (uint32_t)num) return 0xFFFFFFFFu;
            depth--;
            break;
        case 5:
            if (end - pos < 4) return 0xFFFFFFFFu;
            pos += 4;
            break;
        default:
            return 0xFFFFFFFFu;
        }
    }
    return pos;
}

// ---- field stores ---------------------------------------------------------------

template <uint32_t NEED>
__device__ __forceinline__ void store_varint_field(Flow &f, uint32_t num, unsigned long long v)
{
    // consumeUint64 / consumeUint32 / consumeEnum: last value wins, u32 = low 32 bits
    switch (num) {
    case 1: if (NEED & F_TYPE) f.type = (uint32_t)v; break;
    case 2: if (NEED & F_TIME_RECEIVED) f.time_received = v; break;
    case 3: if (NEED & F_SAMPLING_RATE) f.sampling_rate = v; break;
    case 4: if (NEED & F_SEQUENCE_NUM) f.sequence_num = (uint32_t)v; break;
    case 9: if (NEED & F_BYTES) f.bytes = v; break;
    case 10: if (NEED & F_PACKETS) f.packets = v; break;
    case 14: if (NEED & F_SRC_AS) f.src_as = (uint32_t)v; break;
    case 15: if (NEED & F_DST_AS) f.dst_as = (uint32_t)v; break;
    case 20: if (NEED & F_PROTO) f.proto = (uint32_t)v; break;
    case 21: if (NEED & F_SRC_PORT) f.src_port = (uint32_t)v; break;
    case 22: if (NEED & F_DST_PORT) f.dst_port = (uint32_t)v; break;
    case 30: if (NEED & F_ETYPE) f.etype = (uint32_t)v; break;
    case 38: if (NEED & F_TIME_FLOW_START) f.time_flow_start = v; break;
    default: break;
    }
}

// first min(len,16) payload bytes -> 4 big-endian words, zero right-padded
template <class Src>
__device__ __forceinline__ void load_addr(const Src &s, uint32_t pos, uint32_t len, uint32_t out[4])
{
    const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
    uint32_t w[5];
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = s.word(i + k);
    const uint32_t n = len < 16u ? len : 16u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t le = __funnelshift_r(w[k], w[k + 1], sh);
        const uint32_t be = __byte_perm(le, 0, 0x0123);
        const int nb = (int)n - 4 * k;  // valid bytes in this word
        const uint32_t mask = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : (0xFFFFFFFFu << (32 - 8 * nb)));
        out[k] = be & mask;
    }
}

// ---- the decoder -------------------------------------------------------------------

// Decode the message occupying [pos,end) of the source.  Returns true when
// proto.Unmarshal would return nil.  `f` must be zero-initialised (m.Reset()).
template <uint32_t NEED, class Src>
__device__ __forceinline__ bool decode_message(const Src &s, uint32_t pos, const uint32_t end, Flow &f)
{
    while (pos < end) {
        uint32_t lo, hi;
        s.window(pos, lo, hi);
        // ---- tag ----
        uint32_t tag, tn;
        if (!(lo & 0x80u)) {
            tag = lo & 0x7fu;
            tn = 1;
        } else if (!(lo & 0x8000u)) {
            tag = (lo & 0x7fu) | ((lo >> 1) & 0x3f80u);
            tn = 2;
        } else {
            unsigned long long t64;
            tn = varint_slow(s, pos, end, t64);
            if (!tn) return false;
            if ((t64 >> 3) > 0x1fffffffull) return false;
            tag = (uint32_t)t64;
            s.window(pos + tn, lo, hi);  // re-centre the window on the value
            pos += tn;
            tn = 0;
        }
        const uint32_t num = tag >> 3, wt = tag & 7u;
        if (num == 0) return false;
        pos += tn;
        if (pos > end) return false;  // tag ran past the end (tn==2 with one byte left)
        // value window: bytes pos.. (at least 6 valid)
        const uint32_t sh = tn * 8u;
        const uint32_t xlo = __funnelshift_r(lo, hi, sh);
        const uint32_t xhi = hi >> sh;
        if (wt == 0) {
            // ---- varint ----
            unsigned long long v;
            uint32_t vn;
            const uint32_t stop_lo = ~xlo & 0x80808080u;
            if (stop_lo) {
                // 1..4 bytes
                const uint32_t t = __ffs(stop_lo);  // 8,16,24,32
                vn = t >> 3;
                const uint32_t x = xlo & (0xFFFFFFFFu >> (32u - t));
                v = (x & 0x7fu) | ((x >> 1) & 0x3f80u) | ((x >> 2) & 0x1fc000u) | ((x >> 3) & 0xfe00000u);
            } else if (!(xhi & 0x80u)) {
                // 5 bytes (every Unix timestamp since 1978)
                vn = 5;
                const uint32_t x = xlo;
                const uint32_t low28 = (x & 0x7fu) | ((x >> 1) & 0x3f80u) | ((x >> 2) & 0x1fc000u) | ((x >> 3) & 0xfe00000u);
                v = (unsigned long long)low28 | ((unsigned long long)(xhi & 0x7fu) << 28);
            } else {
                vn = varint_slow(s, pos, end, v);
                if (!vn) return false;
            }
            pos += vn;
            if (pos > end) return false;
            store_varint_field<NEED>(f, num, v);
        } else if (wt == 2) {
            // ---- length-delimited ----
            unsigned long long ln;
            uint32_t vn;
            if (!(xlo & 0x80u)) {
                ln = xlo & 0x7fu;
                vn = 1;
            } else {
                vn = varint_slow(s, pos, end, ln);
                if (!vn) return false;
            }
            pos += vn;
            if (pos > end || ln > (unsigned long long)(end - pos)) return false;
            const uint32_t n = (uint32_t)ln;
            if (num == 6) {
                if (NEED & F_SRC_ADDR) { load_addr(s, pos, n, f.src); f.src_len = n; }
            } else if (num == 7) {
                if (NEED & F_DST_ADDR) { load_addr(s, pos, n, f.dst); f.dst_len = n; }
            } else if (num == 11) {
                if (NEED & F_SAMPLER_ADDR) { load_addr(s, pos, n, f.sampler); f.sampler_len = n; }
            } else if (num == 100 || num == 101) {
                if (!utf8_valid(s, pos, n)) return false;
            }
            pos += n;
        } else if (wt == 5) {
            if (end - pos < 4) return false;
            pos += 4;
        } else if (wt == 1) {
            if (end - pos < 8) return false;
            pos += 8;
        } else if (wt == 3) {
            pos = skip_group(s, pos, end, num);
            if (pos == 0xFFFFFFFFu) return false;
        } else {
            return false;  // end-group at top level, or wire type 6/7
        }
    }
    return true;
}

// Decode the record occupying [pos,end): bare message, or varint(len) || message
// whose length must fill the span exactly.
template <uint32_t NEED, class Src>
__device__ __forceinline__ bool decode_record(const Src &s, uint32_t pos, uint32_t end, bool framed, Flow &f)
{
    if (framed) {
        if (pos >= end) return false;
        uint32_t lo, hi;
        s.window(pos, lo, hi);
        unsigned long long mlen;
        uint32_t n;
        if (!(lo & 0x80u)) {
            mlen = lo & 0x7fu;
            n = 1;
        } else if (!(lo & 0x8000u)) {
            mlen = (lo & 0x7fu) | ((lo >> 1) & 0x3f80u);
            n = 2;
        } else {
            n = varint_slow(s, pos, end, mlen);
            if (!n) return false;
        }
        if (pos + n > end || mlen != (unsigned long long)(end - pos - n)) return false;
        pos += n;
    }
    return decode_message<NEED>(s, pos, end, f);
}

__device__ __forceinline__ void flow_reset(Flow &f)
{
    f.time_received = f.sampling_rate = f.time_flow_start = f.bytes = f.packets = 0;
    f.type = f.sequence_num = f.src_as = f.dst_as = f.etype = f.proto = f.src_port = f.dst_port = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) f.src[i] = f.dst[i] = f.sampler[i] = 0;
    f.src_len = f.dst_len = f.sampler_len = 0;
}

}  // namespace fa
