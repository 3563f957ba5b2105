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

FA_DEV __forceinline__ void flow_reset(Flow &f)
{
    f.time_received = f.sampling_rate = f.time_flow_start = f.bytes = f.packets = 0;
    f.type = f.sequence_num = f.src_as = f.dst_as = f.etype = f.proto = f.src_port = f.dst_port = 0;
FA_UNROLL
    for (int i = 0; i < 4; i++) f.src[i] = f.dst[i] = f.sampler[i] = 0;
    f.src_len = f.dst_len = f.sampler_len = 0;
}

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

// ---- the shape fast path -----------------------------------------------------------------------
//
// Every protobuf serializer (Go's proto.Marshal at mocker/mocker.go:97 included) writes each field at most once,
// in ascending field-number order, and proto3 leaves zero values out.  The records of one producer therefore
// all look like a SUBSEQUENCE of one short list of tags.  The fast path walks that list in lock step -- "does my
// cursor sit on this tag?  then consume the field" -- with no per-lane dispatch on field number or wire type:
//
//   * the fields a kernel KEEPS are known at compile time (its NEED mask): each is a piece of straight-line code
//     with its tag, wire type and destination register as immediates;
//   * between two kept fields sits a SEGMENT of the ShapeTable: the other tags the producer sends, learned from a
//     sample of the batch (shape_collect / shape_build), validated and skipped by one small loop.
//
// The fast path ACCEPTS a record only when its bytes are exactly a sequence of such fields in ascending order, each
// with a 1..5-byte varint, a 1-byte length or a fixed-width value, ending on the record's last byte; the values it
// kept are then the values proto.Unmarshal keeps (no duplicates, so last-wins is moot).  Anything else -- unknown
// or repeated tags, another order, long varints, groups, strings -- is NOT decided here: the caller re-parses the
// record with decode_record (the order-agnostic decoder above).  Results never depend on the table.
constexpr uint32_t kShapeMax = 48;
constexpr uint32_t kShapeNoSlot = 31;
constexpr uint32_t kKeptFields = 16;
// flags a learned tag value carries above its 14 bits
constexpr uint16_t kTagSaw4 = 1u << 14;  // a varint of 1..4 bytes was seen for this tag
constexpr uint16_t kTagSaw5 = 1u << 15;  // a 5-byte varint was seen

// skip-step word: bits 0-15 the tag bytes as they appear in the stream (little endian), then one-hot kind bits.  No
// kind bit = "varint of 1..4 bytes": the commonest step is the one with no flag set.  The three varint kinds differ
// in the lengths they take (V5: exactly 5 bytes -- every Unix timestamp since 1978 --, V45: 1..5); which one a tag
// gets is learned from the lengths seen in the sample.
constexpr uint32_t kStepV5 = 1u << 24, kStepV45 = 1u << 25, kStepLen = 1u << 26, kStepFixed32 = 1u << 27, kStepFixed64 = 1u << 28;
constexpr uint32_t kStepNotPlain = 0x1Fu << 24;

struct ShapeTable {
    // seg[i]: skip steps in front of the i-th kept field of the kernel (i = its rank among the kernel's kept fields in
    // tag order; the last used entry is the tail behind the last kept field): count with one-byte tags | count with
    // two-byte tags << 8.  Their words follow each other in step[], one-byte tags first within a segment.
    uint32_t seg[kKeptFields + 1];
    uint32_t step[kShapeMax + 1];  // + one spare word (shape_skip reads one step ahead)
};

// schema of pb-ext/flow.pb.go:58-143 for the fields the kernels keep: F_* bit index -> tag value (num << 3 | wt).
// Ascending, like the bits themselves.
FA_DEV constexpr uint32_t shape_tag_of_bit(int bit)
{
    constexpr uint32_t t[kKeptFields] = {(1u << 3) | 0u,  (2u << 3) | 0u,  (3u << 3) | 0u,  (4u << 3) | 0u,  (6u << 3) | 2u,  (7u << 3) | 2u,
                                        (9u << 3) | 0u,  (10u << 3) | 0u, (11u << 3) | 2u, (14u << 3) | 0u, (15u << 3) | 0u, (20u << 3) | 0u,
                                        (21u << 3) | 0u, (22u << 3) | 0u, (30u << 3) | 0u, (38u << 3) | 0u};
    return t[bit];
}
FA_DEV __forceinline__ uint32_t shape_slot_of(uint32_t tagval)
{
    for (int b = 0; b < (int)kKeptFields; b++)
        if (shape_tag_of_bit(b) == tagval) return (uint32_t)b;
    return kShapeNoSlot;
}
// tag value -> its bytes in the stream, little endian (one byte below 128, two below 2^14)
FA_DEV constexpr uint32_t shape_raw_tag(uint32_t tagval) { return tagval >= 128u ? ((tagval & 0x7fu) | 0x80u | ((tagval >> 7) << 8)) : tagval; }

// Can a field with this tag value be a step at all?  Field numbers 1..2047 (1-2 tag bytes), wire types
// 0/1/2/5, and not the two proto3 strings (100, 101: UTF-8 is checked by the generic decoder only).
FA_DEV __forceinline__ bool shape_tag_ok(uint32_t tagval)
{
    const uint32_t num = tagval >> 3, wt = tagval & 7u;
    if (num < 1u || num > 2047u) return false;
    if (wt != 0u && wt != 1u && wt != 2u && wt != 5u) return false;
    if ((num == 100u || num == 101u) && wt == 2u) return false;
    return true;
}

// Learned tag values (bits 14/15 = kTagSaw4/kTagSaw5, neither = lengths unknown) -> table for a kernel that keeps the
// fields in `need`.  More than kShapeMax tags: no learned steps (the kept fields alone are still walked).
FA_DEV inline void shape_build(const uint16_t *tags, uint32_t n, uint32_t need, ShapeTable &t)
{
    for (uint32_t i = 0; i <= kKeptFields; i++) t.seg[i] = 0;
    if (n > kShapeMax) n = 0;
    uint32_t n_steps = 0, segment = 0, lo = 0;  // lo: first tag value of the current segment
    // segment boundaries: the kept fields' tag values, ascending; the segment after the last one has no upper bound
    for (int bit = 0; bit <= (int)kKeptFields; bit++) {
        if (bit < (int)kKeptFields && !((need >> bit) & 1u)) continue;
        const uint32_t hi = bit < (int)kKeptFields ? shape_tag_of_bit(bit) : 0x10000u;
        for (int two = 0; two < 2; two++)  // one-byte tags first (ascending order has them first anyway)
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t tagval = tags[i] & 0x3fffu, wt = tagval & 7u;
                if (!shape_tag_ok(tagval) || (tagval >= 128u) != (two != 0)) continue;
                if (tagval < lo || tagval >= hi) continue;
                const uint32_t slot = shape_slot_of(tagval);
                if (slot != kShapeNoSlot && ((need >> slot) & 1u)) continue;  // a kept field: compiled code, not a step
                const bool saw4 = (tags[i] & kTagSaw4) != 0, saw5 = (tags[i] & kTagSaw5) != 0;
                const uint32_t kind = wt == 0u ? (saw4 == saw5 ? kStepV45 : (saw5 ? kStepV5 : 0u))
                                               : (wt == 2u ? kStepLen : (wt == 5u ? kStepFixed32 : kStepFixed64));
                t.step[n_steps++] = shape_raw_tag(tagval) | kind;
                t.seg[segment] += two ? 1u << 8 : 1u;
            }
        segment++;
        lo = hi + 1u;  // the next segment starts behind this kept field's tag value
    }
    t.step[n_steps] = 0u;  // the look-ahead word
}

// Walk one sampled record and report every field a step could stand for: calls seen(tagval, value_bytes) per field
// (value_bytes: the varint's length, 0 for the other wire types).  Stops at the first thing the fast path would not
// take (the record then teaches what it showed so far).
template <class Src, class Seen>
FA_DEV __forceinline__ void shape_collect(const Src s, uint32_t pos, const uint32_t end, bool framed, Seen seen)
{
    if (framed) {  // skip the length prefix (1-2 bytes; longer ones are not worth learning from)
        if (pos >= end) return;
        const uint32_t b0 = s.byte(pos);
        pos += (b0 & 0x80u) ? 2u : 1u;
    }
    for (uint32_t guard = 0; guard < 2u * kShapeMax && pos < end; guard++) {
        const uint32_t b0 = s.byte(pos);
        uint32_t tagval = b0 & 0x7fu, tl = 1u;
        if (b0 & 0x80u) {
            if (pos + 1u >= end) return;
            const uint32_t b1 = s.byte(pos + 1u);
            if ((b1 & 0x80u) || b1 == 0u) return;  // 3+ bytes, or an over-long encoding
            tagval |= b1 << 7;
            tl = 2u;
        }
        if (!shape_tag_ok(tagval)) return;
        pos += tl;
        const uint32_t wt = tagval & 7u;
        uint32_t vb = 0;
        if (wt == 0u) {
            uint32_t k = 0;
            while (k < 5u && pos + k < end && (s.byte(pos + k) & 0x80u)) k++;
            if (k >= 5u || pos + k >= end) return;
            vb = k + 1u;
            pos += vb;
        } else if (wt == 2u) {
            if (pos >= end) return;
            const uint32_t ln = s.byte(pos);
            if (ln & 0x80u) return;
            pos += 1u + ln;
        } else {
            pos += wt == 5u ? 4u : 8u;
        }
        if (pos > end) return;
        seen(tagval, vb);
    }
}

// 4 x 7 payload bits of the varint bytes in x -> 28-bit value
FA_DEV __forceinline__ uint32_t shape_low28(uint32_t x)
{
    return (x & 0x7fu) | ((x >> 1) & 0x3f80u) | ((x >> 2) & 0x1fc000u) | ((x >> 3) & 0xfe00000u);
}

// NR records' share of `count` skip steps starting at byte offset jo of sh.step, tags TWO ? two bytes : one byte long.
// The table walk (step word, loop, dispatch on the step's kind) is paid once per step for all NR records.
template <bool TWO, int NR, class Src>
FA_DEV __forceinline__ void shape_skip(const ShapeTable &sh, const Src s, uint32_t &jo, const uint32_t count, uint32_t (&pos)[NR],
                                       const uint32_t (&end)[NR])
{
    constexpr uint32_t TMASK = TWO ? 0xffffu : 0xffu, TSH = TWO ? 16u : 8u, TL = TWO ? 2u : 1u;
    const uint32_t je = jo + count * 4u;
    // the next step's word is fetched one step ahead (an indexed constant load sits on the loop's critical path otherwise);
    // step[] has a spare word behind the last step, so the look-ahead never leaves the table
    uint32_t st_next = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(sh.step) + jo);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (; jo != je; jo += 4u) {
        const uint32_t st = st_next;
        st_next = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(sh.step) + jo + 4u);
        uint32_t x[NR], hi[NR];
        bool m[NR];
FA_UNROLL
        for (int q = 0; q < NR; q++) {
            uint32_t lo;
            s.window(pos[q], lo, hi[q]);
            m[q] = (((lo ^ st) & TMASK) == 0u) && pos[q] < end[q];
            x[q] = funnel_r(lo, hi[q], TSH);  // value bytes 0-3
        }
        if (!(st & kStepNotPlain)) {
            // a varint of 1..4 bytes: the commonest step
FA_UNROLL
            for (int q = 0; q < NR; q++) {
                const uint32_t t = find_first_set(~x[q] & 0x80808080u);  // 8, 16, 24, 32; 0: no terminator in four bytes
                pos[q] += (m[q] && t != 0u) ? TL + (t >> 3) : 0u;
            }
        } else if (st & kStepV5) {
FA_UNROLL
            for (int q = 0; q < NR; q++) {
                const bool five = (x[q] & 0x80808080u) == 0x80808080u && !((hi[q] >> TSH) & 0x80u);
                pos[q] += (m[q] && five) ? TL + 5u : 0u;
            }
        } else if (st & kStepLen) {
FA_UNROLL
            for (int q = 0; q < NR; q++) pos[q] += (m[q] && !(x[q] & 0x80u)) ? TL + 1u + (x[q] & 0x7fu) : 0u;  // one-byte lengths only
        } else if (st & kStepV45) {
FA_UNROLL
            for (int q = 0; q < NR; q++) {
                const uint32_t stop = ~x[q] & 0x80808080u;
                uint32_t t = find_first_set(stop);
                t = (stop == 0u && !((hi[q] >> TSH) & 0x80u)) ? 40u : t;
                pos[q] += (m[q] && t != 0u) ? TL + (t >> 3) : 0u;
            }
        } else {
FA_UNROLL
            for (int q = 0; q < NR; q++) pos[q] += m[q] ? TL + ((st & kStepFixed32) ? 4u : 8u) : 0u;
        }
    }
}

// The kept field BIT (an F_* bit index) of NR records: tag, wire type and destination are immediates.
template <int BIT, int NR, class Src>
FA_DEV __forceinline__ void shape_keep(const Src s, uint32_t (&pos)[NR], const uint32_t (&end)[NR], Flow (&f)[NR])
{
    constexpr uint32_t TAG = shape_tag_of_bit(BIT), RAW = shape_raw_tag(TAG);
    constexpr bool TWO = TAG >= 128u;
    constexpr uint32_t TMASK = TWO ? 0xffffu : 0xffu, TSH = TWO ? 16u : 8u, TL = TWO ? 2u : 1u;
FA_UNROLL
    for (int q = 0; q < NR; q++) {
        uint32_t lo, hi;
        s.window(pos[q], lo, hi);
        bool m = ((lo & TMASK) == RAW) && pos[q] < end[q];
        const uint32_t x = funnel_r(lo, hi, TSH), y = hi >> TSH;  // value bytes 0-3, 4..
        if ((TAG & 7u) == 2u) {
            // bytes: SrcAddr / DstAddr / SamplerAddress
            const uint32_t ln = x & 0x7fu;
            m = m && !(x & 0x80u);  // two-byte lengths: not decided here
            uint32_t a[4];
            load_addr(s, pos[q] + TL + 1u, ln, a);
            uint32_t *dst = BIT == 4 ? f[q].src : (BIT == 5 ? f[q].dst : f[q].sampler);
            uint32_t &dlen = BIT == 4 ? f[q].src_len : (BIT == 5 ? f[q].dst_len : f[q].sampler_len);
FA_UNROLL
            for (int k = 0; k < 4; k++) dst[k] = m ? a[k] : dst[k];
            dlen = m ? ln : dlen;
            pos[q] += m ? TL + 1u + ln : 0u;
        } else {
            const uint32_t stop = ~x & 0x80808080u;  // terminators among the first four bytes
            uint32_t t = find_first_set(stop);       // 8, 16, 24, 32; 0: none
            t = (stop == 0u && !(y & 0x80u)) ? 40u : t;
            m = m && t != 0u;  // six bytes or more: not decided here
            const uint32_t keep_bits = t >= 32u ? 0xFFFFFFFFu : (0xFFFFFFFFu >> ((32u - t) & 31u));
            const uint32_t b4 = t == 40u ? (y & 0x7fu) : 0u;  // bits 28..34
            const uint32_t vlo = shape_low28(x & keep_bits) | (b4 << 28), vhi = b4 >> 4;
            const unsigned long long v = ((unsigned long long)vhi << 32) | vlo;
            // consumeUint64 / consumeUint32 / consumeEnum: u32 = low 32 bits
            if (BIT == 0) f[q].type = m ? vlo : f[q].type;
            if (BIT == 1) f[q].time_received = m ? v : f[q].time_received;
            if (BIT == 2) f[q].sampling_rate = m ? v : f[q].sampling_rate;
            if (BIT == 3) f[q].sequence_num = m ? vlo : f[q].sequence_num;
            if (BIT == 6) f[q].bytes = m ? v : f[q].bytes;
            if (BIT == 7) f[q].packets = m ? v : f[q].packets;
            if (BIT == 9) f[q].src_as = m ? vlo : f[q].src_as;
            if (BIT == 10) f[q].dst_as = m ? vlo : f[q].dst_as;
            if (BIT == 11) f[q].proto = m ? vlo : f[q].proto;
            if (BIT == 12) f[q].src_port = m ? vlo : f[q].src_port;
            if (BIT == 13) f[q].dst_port = m ? vlo : f[q].dst_port;
            if (BIT == 14) f[q].etype = m ? vlo : f[q].etype;
            if (BIT == 15) f[q].time_flow_start = m ? v : f[q].time_flow_start;
            pos[q] += m ? TL + (t >> 3) : 0u;
        }
    }
}

// segment SEG of the table (both tag lengths), then -- for BIT < 16 -- the kept field BIT, then on to the next bit
template <uint32_t NEED, int BIT, int SEG, int NR, class Src>
struct ShapeWalk {
    static FA_DEV __forceinline__ void run(const ShapeTable &sh, const Src s, uint32_t &jo, uint32_t (&pos)[NR], const uint32_t (&end)[NR],
                                           Flow (&f)[NR])
    {
        constexpr bool KEPT = BIT < (int)kKeptFields && ((NEED >> (BIT < (int)kKeptFields ? BIT : 0)) & 1u);
        if (KEPT || BIT == (int)kKeptFields) {
            const uint32_t sg = sh.seg[SEG];
            shape_skip<false, NR>(sh, s, jo, sg & 0xffu, pos, end);
            shape_skip<true, NR>(sh, s, jo, (sg >> 8) & 0xffu, pos, end);
        }
        if (KEPT) shape_keep<(BIT < (int)kKeptFields ? BIT : 0), NR>(s, pos, end, f);
        ShapeWalk<NEED, BIT + 1, SEG + (KEPT ? 1 : 0), NR, Src>::run(sh, s, jo, pos, end, f);
    }
};
template <uint32_t NEED, int SEG, int NR, class Src>
struct ShapeWalk<NEED, (int)kKeptFields + 1, SEG, NR, Src> {
    static FA_DEV __forceinline__ void run(const ShapeTable &, const Src, uint32_t &, uint32_t (&)[NR], const uint32_t (&)[NR], Flow (&)[NR]) {}
};

// Framed or bare records [pos,end) through the shape fast path, NR at a time.  taken[q] true: f[q] holds record q's
// fields.  false: undecided -- the caller resets f[q] and runs decode_record (the order-agnostic decoder).  Records
// with pos == end are idle lanes' placeholders: never taken.  f must be zero-initialised.
template <uint32_t NEED, int NR, class Src>
FA_DEV __forceinline__ void decode_records_shape(const ShapeTable &sh, const Src s, uint32_t (&pos)[NR], const uint32_t (&end)[NR], bool framed,
                                                 Flow (&f)[NR], bool (&taken)[NR])
{
FA_UNROLL
    for (int q = 0; q < NR; q++) {
        taken[q] = pos[q] < end[q];  // empty spans are the order-agnostic decoder's business
        if (framed) {  // varint(len) || message, len must fill the span: mocker/mocker.go:98-101
            uint32_t lo, hi;
            s.window(pos[q], lo, hi);
            const bool one = !(lo & 0x80u);
            const uint32_t mlen = one ? (lo & 0x7fu) : ((lo & 0x7fu) | ((lo >> 1) & 0x3f80u));
            const uint32_t hn = one ? 1u : 2u;
            taken[q] = taken[q] && (one || !(lo & 0x8000u)) && pos[q] + hn <= end[q] && mlen == end[q] - pos[q] - hn;
            pos[q] = taken[q] ? pos[q] + hn : end[q];
        }
    }
    uint32_t jo = 0;
    ShapeWalk<NEED, 0, 0, NR, Src>::run(sh, s, jo, pos, end, f);
FA_UNROLL
    for (int q = 0; q < NR; q++) taken[q] = taken[q] && pos[q] == end[q];
}

// one record (the CPU harness and single-record callers)
template <uint32_t NEED, class Src>
FA_DEV __forceinline__ bool decode_record_shape(const ShapeTable &sh, const Src s, uint32_t pos, uint32_t end, bool framed, Flow &f)
{
    uint32_t p[1] = {pos};
    const uint32_t e[1] = {end};
    Flow g[1] = {f};
    bool taken[1];
    decode_records_shape<NEED, 1>(sh, s, p, e, framed, g, taken);
    f = g[0];
    return taken[0];
}

}  // namespace fa
