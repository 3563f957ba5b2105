// decode.cuh -- device-side proto3 decoder for flowprotob.FlowMessage.
//
// One thread decodes one record from a byte tile staged in shared memory (or,
// for records that do not fit the tile, straight from global memory).  The rules
// are those of proto.Unmarshal as called at inserter/inserter.go:124
// (golang/protobuf v1.4.3 -> google.golang.org/protobuf, go.mod:7) for the field
// table of pb-ext/flow.pb.go:58-143:
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
// are decoded from that window branch-free; the value lands in the wanted
// register through predicated selects (no switch, no local memory).  Long
// varints, long tags, groups and UTF-8 take byte-wise slow paths kept out of
// line; they return by value so nothing on the hot path has its address taken.
#pragma once
#include <stdint.h>

// The decoder also compiles as plain host C++ (nvcc's host pass): tests/decode_host builds it that way so the
// CPU test suite can drive the very same template against the oracle and the golden vectors.  That build is test
// infrastructure; nothing in the product calls the host instantiation.
#define FA_DEV __host__ __device__
#ifdef __CUDA_ARCH__
#define FA_UNROLL _Pragma("unroll")
#else
#define FA_UNROLL
#endif

namespace fa {

FA_DEV __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
FA_DEV __forceinline__ uint32_t byte_swap(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}
FA_DEV __forceinline__ uint32_t find_first_set(uint32_t x)  // 1-based, 0 for x == 0
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)x);
#else
    return (uint32_t)__builtin_ffs((int)x);
#endif
}
FA_DEV __forceinline__ uint32_t shl_clamp(uint32_t n)  // 1 << n, 0 for n >= 32 (shl.b32 clamps)
{
#ifdef __CUDA_ARCH__
    uint32_t bit;
    asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(n));
    return bit;
#else
    return n < 32u ? 1u << n : 0u;
#endif
}
FA_DEV __forceinline__ uint32_t load_word(const uint32_t *p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

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

// Byte source over aligned 32-bit words in GLOBAL memory; over-reads are clamped
// to limit_word (the last word of the buffer).
struct ByteSrc {
    const uint32_t *words;
    uint32_t limit_word;
    FA_DEV __forceinline__ uint32_t word(uint32_t i) const { return load_word(words + (i < limit_word ? i : limit_word)); }
    FA_DEV __forceinline__ uint32_t byte(uint32_t pos) const { return (word(pos >> 2) >> ((pos & 3u) * 8u)) & 0xffu; }
    FA_DEV __forceinline__ void window(uint32_t pos, uint32_t &lo, uint32_t &hi) const
    {
        const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
        const uint32_t a = word(i), b = word(i + 1), c = word(i + 2);
        lo = funnel_r(a, b, sh);
        hi = funnel_r(b, c, sh);
    }
};

#ifdef __CUDACC__
// Byte source over a tile resident in (padded) SHARED memory, addressed by its
// 32-bit shared-window address so the loads are LDS, never generic.
struct SmemSrc {
    uint32_t base;  // shared address of tile byte 0
    __device__ __forceinline__ uint32_t word(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + i * 4u));
        return v;
    }
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + pos));
        return v;
    }
    // bytes pos..pos+7 as little-endian (lo, hi)
    __device__ __forceinline__ void window(uint32_t pos, uint32_t &lo, uint32_t &hi) const
    {
        const uint32_t a = base + (pos & ~3u), sh = pos * 8u;  // the funnel shift uses sh mod 32
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(a));
#ifdef FA_W2_COND
        // experiment: the third word is only needed for an unaligned cursor
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, 0;\n\t@p ld.shared.u32 %0, [%1+8];\n\t}" : "=r"(w2) : "r"(a), "r"(pos & 3u));
#else
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(a));
#endif
        lo = funnel_r(w0, w1, sh);
        hi = funnel_r(w1, w2, sh);
    }
};
#endif  // __CUDACC__

// ---- slow paths (out of line, everything by value) -----------------------------

struct VarintRes {
    unsigned long long v;
    uint32_t n;  // bytes consumed, 0 = error
};

// protowire.ConsumeVarint, byte at a time
template <class Src>
FA_DEV __noinline__ VarintRes varint_slow(const Src s, uint32_t pos, uint32_t end)
{
    VarintRes r;
    r.v = 0;
    r.n = 0;
    unsigned long long x = 0;
    for (uint32_t i = 0; i < 10; i++) {
        if (pos + i >= end) return r;  // truncated
        const unsigned long long y = s.byte(pos + i);
        if (i == 9) {
            if (y >= 2) return r;  // overflow
            r.v = x | (y << 63);
            r.n = 10;
            return r;
        }
        x |= (y & 0x7f) << (7 * i);
        if (y < 0x80) {
            r.v = x;
            r.n = i + 1;
            return r;
        }
    }
    return r;
}

template <class Src>
FA_DEV __noinline__ bool utf8_valid(const Src s, uint32_t pos, uint32_t n)
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
FA_DEV __noinline__ uint32_t skip_group(const Src s, uint32_t pos, uint32_t end, uint32_t start_num)
{
    uint32_t stack[FA_MAX_GROUP_DEPTH];
    int depth = 0;
    stack[depth++] = start_num;
    while (depth > 0) {
        VarintRes t = varint_slow(s, pos, end);
        if (!t.n) return 0xFFFFFFFFu;
        pos += t.n;
        const unsigned long long num = t.v >> 3;
        const uint32_t wt = (uint32_t)t.v & 7u;
        if (num < 1 || num > 0x7fffffffull) return 0xFFFFFFFFu;
        switch (wt) {
        case 0: {
            VarintRes v = varint_slow(s, pos, end);
            if (!v.n) return 0xFFFFFFFFu;
            pos += v.n;
            break;
        }
        case 1:
            if (end - pos < 8) return 0xFFFFFFFFu;
            pos += 8;
            break;
        case 2: {
            VarintRes v = varint_slow(s, pos, end);
            if (!v.n) return 0xFFFFFFFFu;
            pos += v.n;
            if (v.v > (unsigned long long)(end - pos)) return 0xFFFFFFFFu;
            pos += (uint32_t)v.v;
            break;
        }
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

// ---- field stores: predicated selects, no branches ----------------------------------

template <uint32_t NEED>
FA_DEV __forceinline__ void store_varint_field(Flow &f, const uint32_t num, const unsigned long long v)
{
    // consumeUint64 / consumeUint32 / consumeEnum: last value wins, u32 = low 32 bits
    const uint32_t v32 = (uint32_t)v;
    if (NEED & F_TYPE) f.type = num == 1 ? v32 : f.type;
    if (NEED & F_TIME_RECEIVED) f.time_received = num == 2 ? v : f.time_received;
    if (NEED & F_SAMPLING_RATE) f.sampling_rate = num == 3 ? v : f.sampling_rate;
    if (NEED & F_SEQUENCE_NUM) f.sequence_num = num == 4 ? v32 : f.sequence_num;
    if (NEED & F_BYTES) f.bytes = num == 9 ? v : f.bytes;
    if (NEED & F_PACKETS) f.packets = num == 10 ? v : f.packets;
    if (NEED & F_SRC_AS) f.src_as = num == 14 ? v32 : f.src_as;
    if (NEED & F_DST_AS) f.dst_as = num == 15 ? v32 : f.dst_as;
    if (NEED & F_PROTO) f.proto = num == 20 ? v32 : f.proto;
    if (NEED & F_SRC_PORT) f.src_port = num == 21 ? v32 : f.src_port;
    if (NEED & F_DST_PORT) f.dst_port = num == 22 ? v32 : f.dst_port;
    if (NEED & F_ETYPE) f.etype = num == 30 ? v32 : f.etype;
    if (NEED & F_TIME_FLOW_START) f.time_flow_start = num == 38 ? v : f.time_flow_start;
}

// first min(len,16) payload bytes -> 4 big-endian words, zero right-padded
template <class Src>
FA_DEV __forceinline__ void load_addr(const Src s, uint32_t pos, uint32_t len, uint32_t out[4])
{
    const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
    uint32_t w[5];
FA_UNROLL
    for (int k = 0; k < 5; k++) w[k] = s.word(i + k);
    const uint32_t n = len < 16u ? len : 16u;
FA_UNROLL
    for (int k = 0; k < 4; k++) {
        const uint32_t le = funnel_r(w[k], w[k + 1], sh);
        const uint32_t be = byte_swap(le);
        const int nb = (int)n - 4 * k;  // valid bytes in this word
        const uint32_t mask = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : (0xFFFFFFFFu << (32 - 8 * nb)));
        out[k] = be & mask;
    }
}

// which varint field numbers a kernel reads: bit n of (m1:m0) = field n
__host__ __device__ constexpr uint32_t need_mask_lo(uint32_t need)
{
    return ((need & F_TYPE) ? 1u << 1 : 0u) | ((need & F_TIME_RECEIVED) ? 1u << 2 : 0u) | ((need & F_SAMPLING_RATE) ? 1u << 3 : 0u) |
           ((need & F_SEQUENCE_NUM) ? 1u << 4 : 0u) | ((need & F_BYTES) ? 1u << 9 : 0u) | ((need & F_PACKETS) ? 1u << 10 : 0u) |
           ((need & F_SRC_AS) ? 1u << 14 : 0u) | ((need & F_DST_AS) ? 1u << 15 : 0u) | ((need & F_PROTO) ? 1u << 20 : 0u) |
           ((need & F_SRC_PORT) ? 1u << 21 : 0u) | ((need & F_DST_PORT) ? 1u << 22 : 0u) | ((need & F_ETYPE) ? 1u << 30 : 0u);
}
__host__ __device__ constexpr uint32_t need_mask_hi(uint32_t need) { return (need & F_TIME_FLOW_START) ? 1u << (38 - 32) : 0u; }

template <uint32_t NEED>
FA_DEV __forceinline__ bool varint_field_needed(uint32_t num)
{
    constexpr uint32_t M0 = need_mask_lo(NEED), M1 = need_mask_hi(NEED);
    // shl.b32 clamps: a shift of 32 or more yields 0, so out-of-range numbers are "not needed"
    uint32_t bit;
    bit = shl_clamp(num);
    bool need = (bit & M0) != 0u;
    if (M1 != 0u) {
        uint32_t bit1;
        bit1 = shl_clamp(num - 32u);  // num < 32 wraps to a huge shift: 0
        need = need || (bit1 & M1) != 0u;
    }
    return need;
}

// everything that is neither a varint nor length-delimited: fixed32/fixed64 skips, group skips,
// and the error cases.  Returns the new position, 0xFFFFFFFF on error.
template <class Src>
FA_DEV __noinline__ uint32_t skip_other(const Src s, uint32_t pos, uint32_t end, uint32_t wt, uint32_t num)
{
    if (pos > end) return 0xFFFFFFFFu;
    if (wt == 5) return end - pos < 4 ? 0xFFFFFFFFu : pos + 4;
    if (wt == 1) return end - pos < 8 ? 0xFFFFFFFFu : pos + 8;
    if (wt == 3) return skip_group(s, pos, end, num);
    return 0xFFFFFFFFu;  // end-group at top level, or wire type 6/7
}

// ---- the decoder -------------------------------------------------------------------

#define FA_POS_ERR 0xFFFFFFFFu

// Decode the message occupying [pos,end) of the source.  Returns true when
// proto.Unmarshal would return nil.  `f` must be zero-initialised (m.Reset()).
// Errors park pos at FA_POS_ERR, which ends the loop; the record is good iff the
// loop ends exactly on `end`.  Values written by a record that is then rejected are
// never looked at.
template <uint32_t NEED, class Src>
FA_DEV __forceinline__ bool decode_message(const Src s, uint32_t pos, const uint32_t end, Flow &f)
{
    uint32_t min_num = 1u;  // smallest field number seen: 0 is illegal; checked once, after the loop
    while (pos < end) {
        uint32_t lo, hi;
        s.window(pos, lo, hi);
        uint32_t num, wt, xlo, xhi;
        if (!(lo & 0x80u)) {
            // 1-byte tag: fields 1..15 (9 of the 13 fields every producer sends)
            wt = lo & 7u;
            num = (lo >> 3) & 0xfu;
            pos += 1u;
            xlo = funnel_r(lo, hi, 8);
            xhi = hi >> 8;
        } else if (!(lo & 0x8000u)) {
            // 2-byte tag: fields 16..2047
            const uint32_t tag = (lo & 0x7fu) | ((lo >> 1) & 0x3f80u);
            wt = tag & 7u;
            num = tag >> 3;
            pos += 2u;
            xlo = funnel_r(lo, hi, 16);
            xhi = hi >> 16;
        } else {
            // three bytes or more: field numbers >= 2048 or an over-long encoding (rare)
            const VarintRes t = varint_slow(s, pos, end);
            if (!t.n || (t.v >> 3) > 0x1fffffffull) {
                pos = FA_POS_ERR;
                continue;
            }
            wt = (uint32_t)t.v & 7u;
            num = (uint32_t)(t.v >> 3);
            pos += t.n;
            s.window(pos, xlo, xhi);  // the window restarts on the value
        }
        min_num = min(min_num, num);
        // xlo/xhi: bytes pos.. (at least 6 valid)
        if (__builtin_expect(wt == 0, 1)) {
            // ---- varint ----
            const uint32_t stop = ~xlo & 0x80808080u;  // terminator bytes among the first four
            if (__builtin_expect(stop != 0u, 1)) {
                // 1..4 bytes
                const uint32_t t = find_first_set(stop);  // 8,16,24,32
                pos += t >> 3;
                if (varint_field_needed<NEED>(num)) {
                    const uint32_t x = xlo & (0xFFFFFFFFu >> ((32u - t) & 31u));
                    const uint32_t v = (x & 0x7fu) | ((x >> 1) & 0x3f80u) | ((x >> 2) & 0x1fc000u) | ((x >> 3) & 0xfe00000u);
                    store_varint_field<NEED>(f, num, (unsigned long long)v);
                }
            } else if (!(xhi & 0x80u)) {
                // 5 bytes: every Unix timestamp since 1978
                pos += 5u;
                if (varint_field_needed<NEED>(num)) {
                    const uint32_t x = xlo;
                    const uint32_t low28 = (x & 0x7fu) | ((x >> 1) & 0x3f80u) | ((x >> 2) & 0x1fc000u) | ((x >> 3) & 0xfe00000u);
                    const uint32_t b4 = xhi & 0x7fu;  // bits 28..34
                    store_varint_field<NEED>(f, num, ((unsigned long long)(b4 >> 4) << 32) | (unsigned long long)(low28 | (b4 << 28)));
                }
            } else {
                const VarintRes r = varint_slow(s, pos, end);  // six bytes or more
                pos = r.n ? pos + r.n : FA_POS_ERR;
                store_varint_field<NEED>(f, num, r.v);
            }
        } else if (wt == 2) {
            // ---- length-delimited ----
            uint32_t n;
            if (__builtin_expect((xlo & 0x80u) != 0u, 0)) {
                const VarintRes r = varint_slow(s, pos, end);
                n = (uint32_t)r.v;
                pos = (r.n && r.v <= 0xFFFFFFFFull) ? pos + r.n : FA_POS_ERR;
            } else {
                n = xlo & 0x7fu;
                pos += 1;
            }
            if (pos > end || n > end - pos) {
                pos = FA_POS_ERR;
                continue;
            }
            if ((NEED & F_SRC_ADDR) && num == 6) {
                load_addr(s, pos, n, f.src);
                f.src_len = n;
            } else if ((NEED & F_DST_ADDR) && num == 7) {
                load_addr(s, pos, n, f.dst);
                f.dst_len = n;
            } else if ((NEED & F_SAMPLER_ADDR) && num == 11) {
                load_addr(s, pos, n, f.sampler);
                f.sampler_len = n;
            } else if (__builtin_expect(num == 100 || num == 101, 0)) {
                if (!utf8_valid(s, pos, n)) {
                    pos = FA_POS_ERR;
                    continue;
                }
            }
            pos += n;
        } else {
            pos = skip_other(s, pos, end, wt, num);  // FA_POS_ERR on error
        }
    }
    // good iff the loop ended exactly on `end` (anything that ran past the record is an error)
    // and no field carried the illegal number 0
    return pos == end && min_num != 0u;
}

// Decode the record occupying [pos,end): bare message, or varint(len) || message
// whose length must fill the span exactly.
template <uint32_t NEED, class Src>
FA_DEV __forceinline__ bool decode_record(const Src s, uint32_t pos, uint32_t end, bool framed, Flow &f)
{
    if (framed) {
        if (pos >= end) return false;
        uint32_t lo, hi;
        s.window(pos, lo, hi);
        unsigned long long mlen;
        uint32_t n;
        if ((lo & 0x8080u) == 0x8080u) {
            const VarintRes r = varint_slow(s, pos, end);
            if (!r.n) return false;
            mlen = r.v;
            n = r.n;
        } else {
            const bool one = !(lo & 0x80u);
            mlen = one ? (lo & 0x7fu) : ((lo & 0x7fu) | ((lo >> 1) & 0x3f80u));
            n = one ? 1u : 2u;
        }
        if (pos + n > end || mlen != (unsigned long long)(end - pos - n)) return false;
        pos += n;
    }
    return decode_message<NEED>(s, pos, end, f);
}

FA_DEV __forceinline__ void flow_reset(Flow &f)
{
    f.time_received = f.sampling_rate = f.time_flow_start = f.bytes = f.packets = 0;
    f.type = f.sequence_num = f.src_as = f.dst_as = f.etype = f.proto = f.src_port = f.dst_port = 0;
FA_UNROLL
    for (int i = 0; i < 4; i++) f.src[i] = f.dst[i] = f.sampler[i] = 0;
    f.src_len = f.dst_len = f.sampler_len = 0;
}

}  // namespace fa
