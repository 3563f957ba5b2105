// kernels.cuh -- sm_100a kernels of the flow-aggregation stage.
//
//   k_decode_aggregate<MODE,W>  fused kernel 1 -> kernel 2: length-delimited
//                               FlowMessage bytes -> group table (+ sketch).  The
//                               columnar intermediate never touches HBM.
//   k_decode_columns            kernel 1 alone: bytes -> 20 decoded columns in HBM
//                               (inserter.go:142-157 row + create.sh:36-59 columns).
//   k_aggregate_columns<MODE>   kernel 2 alone: columns -> group table (+ sketch).
//   k_compact_rows<KW>          flush: occupied slots -> dense fa_row array.
//   k_estimate<KW>              sketch estimate of every group (top-K candidates).
//
// The path is integer / memory bound: no tensor cores anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flowagg.h"
#include "decode.cuh"

namespace fa {

constexpr int kThreads = 256;            // one record per thread per tile
constexpr int kTileRecords = kThreads;   // records per CTA tile
constexpr int kTileBytes = 40 * 1024;    // staged bytes per tile (mean mocker tile: 21.6 KB)
constexpr int kTilePad = 64;             // over-read slack behind the tile

struct Counters {
    unsigned long long n_bad, n_nokey, n_dropped, n_groups, flush_rows, pad[3];
};

struct SubmitParams {
    const uint8_t *buf;       // device bytes; buf[0] is host/stream byte `base`
    unsigned long long base;  // offsets[] are relative to the stream, buf to base (multiple of 16)
    unsigned long long len;   // bytes readable behind buf (rounded up to 16 by the owner)
    const uint32_t *offsets;  // n_records + 1
    uint32_t n_records;
    uint32_t framed;
    // group table
    uint8_t *slots;
    uint32_t slot_mask;
    uint32_t scale;  // FA_CFG_SCALE_SAMPLING
    // sketch (nullptr = off)
    unsigned long long *cms;
    uint32_t cms_depth, cms_wlog2;
    Counters *counters;
};

// ---- key modes ----------------------------------------------------------------------

template <int MODE> struct KeyTraits;
template <> struct KeyTraits<FA_KEY_FLOWS5M> { static constexpr int KW = 4; static constexpr uint32_t NEED = F_TIME_RECEIVED | F_SRC_AS | F_DST_AS | F_ETYPE; };
template <> struct KeyTraits<FA_KEY_ASPAIR> { static constexpr int KW = 2; static constexpr uint32_t NEED = F_SRC_AS | F_DST_AS; };
template <> struct KeyTraits<FA_KEY_SRCADDR> { static constexpr int KW = 4; static constexpr uint32_t NEED = F_SRC_ADDR; };
template <> struct KeyTraits<FA_KEY_DSTADDR> { static constexpr int KW = 4; static constexpr uint32_t NEED = F_DST_ADDR; };
template <> struct KeyTraits<FA_KEY_5TUPLE> { static constexpr int KW = 11; static constexpr uint32_t NEED = F_SRC_ADDR | F_DST_ADDR | F_SRC_PORT | F_DST_PORT | F_PROTO; };
template <> struct KeyTraits<FA_KEY_SRCPORT> { static constexpr int KW = 1; static constexpr uint32_t NEED = F_SRC_PORT; };
template <> struct KeyTraits<FA_KEY_DSTPORT> { static constexpr int KW = 1; static constexpr uint32_t NEED = F_DST_PORT; };

// GROUP BY key of one flow.  false = the flow cannot form the key (an address
// longer than FixedString(16), create.sh:15-16).
template <int MODE>
__device__ __forceinline__ bool make_key(const Flow &f, uint32_t *key)
{
    if (MODE == FA_KEY_FLOWS5M) {
        // toStartOfFiveMinute on the DateTime (UInt32) column: create.sh:39,96
        const uint32_t t = (uint32_t)f.time_received;
        key[0] = t - t % 300u;
        key[1] = f.src_as;
        key[2] = f.dst_as;
        key[3] = f.etype;
        return true;
    } else if (MODE == FA_KEY_ASPAIR) {
        key[0] = f.src_as;
        key[1] = f.dst_as;
        return true;
    } else if (MODE == FA_KEY_SRCADDR) {
#pragma unroll
        for (int i = 0; i < 4; i++) key[i] = f.src[i];
        return f.src_len <= 16;
    } else if (MODE == FA_KEY_DSTADDR) {
#pragma unroll
        for (int i = 0; i < 4; i++) key[i] = f.dst[i];
        return f.dst_len <= 16;
    } else if (MODE == FA_KEY_5TUPLE) {
#pragma unroll
        for (int i = 0; i < 4; i++) key[i] = f.src[i];
#pragma unroll
        for (int i = 0; i < 4; i++) key[4 + i] = f.dst[i];
        key[8] = f.src_port;
        key[9] = f.dst_port;
        key[10] = f.proto;
        return f.src_len <= 16 && f.dst_len <= 16;
    } else if (MODE == FA_KEY_SRCPORT) {
        key[0] = f.src_port;
        return true;
    } else {
        key[0] = f.dst_port;
        return true;
    }
}

// table / sketch hash; same arithmetic as the oracle's fo_hash64
template <int KW>
__device__ __forceinline__ unsigned long long hash64(const uint32_t *key)
{
    unsigned long long h = 0x243F6A8885A308D3ull;
#pragma unroll
    for (int i = 0; i < KW; i += 2) {
        unsigned long long w = key[i];
        if (i + 1 < KW) w |= (unsigned long long)key[i + 1] << 32;
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    h ^= h >> 30;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 27;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 31;
    return h;
}

// ---- group table: open addressing, linear probing, in HBM (L2-resident when small) ---
//
// slot = { u32 state; u32 key[KW]; pad to 8; u64 bytes, packets, count }
template <int KW> struct SlotLayout {
    static constexpr uint32_t VAL_OFF = (4u + 4u * KW + 7u) & ~7u;
    static constexpr uint32_t BYTES = VAL_OFF + 24u;
};
enum : uint32_t { SLOT_EMPTY = 0, SLOT_BUSY = 1, SLOT_READY = 2 };

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// sum(Bytes), sum(Packets), count() for one flow (create.sh:105-107).  UInt64
// wrap-around is native to the 64-bit reduction.
template <int KW>
__device__ __forceinline__ void table_add(const SubmitParams &p, const uint32_t *key, unsigned long long h,
                                          unsigned long long bytes, unsigned long long packets,
                                          unsigned long long count)
{
    uint32_t slot = (uint32_t)(h >> 32) & p.slot_mask;
    for (uint32_t probe = 0; probe <= p.slot_mask; probe++) {
        uint8_t *s = p.slots + (size_t)slot * SlotLayout<KW>::BYTES;
        uint32_t *state = reinterpret_cast<uint32_t *>(s);
        uint32_t *skey = state + 1;
        uint32_t st = ld_acquire_u32(state);
        if (st == SLOT_EMPTY) {
            const uint32_t old = atomicCAS(state, (uint32_t)SLOT_EMPTY, (uint32_t)SLOT_BUSY);
            if (old == SLOT_EMPTY) {
#pragma unroll
                for (int i = 0; i < KW; i++) skey[i] = key[i];
                st_release_u32(state, SLOT_READY);
                atomicAdd(&p.counters->n_groups, 1ull);
                st = SLOT_READY;
            } else {
                st = old;
            }
        }
        while (st == SLOT_BUSY) {  // another thread is publishing this slot's key
            __nanosleep(32);
            st = ld_acquire_u32(state);
        }
        bool same = true;
#pragma unroll
        for (int i = 0; i < KW; i++) same &= (skey[i] == key[i]);
        if (same) {
            unsigned long long *val = reinterpret_cast<unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
            red_add_u64(val + 0, bytes);
            red_add_u64(val + 1, packets);
            red_add_u64(val + 2, count);
            return;
        }
        slot = (slot + 1) & p.slot_mask;
    }
    atomicAdd(&p.counters->n_dropped, count);
}

// count-min sketch update: idx_j = (lo32(h) + j*(hi32(h)|1)) mod w
__device__ __forceinline__ void cms_add(const SubmitParams &p, unsigned long long h, unsigned long long weight)
{
    const uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32) | 1u;
    const uint32_t mask = (1u << p.cms_wlog2) - 1u;
    for (uint32_t j = 0; j < p.cms_depth; j++) {
        const uint32_t idx = (a + j * b) & mask;
        red_add_u64(p.cms + ((size_t)j << p.cms_wlog2) + idx, weight);
    }
}

template <int MODE>
__device__ __forceinline__ void aggregate_flow(const SubmitParams &p, const Flow &f, uint32_t &nokey)
{
    constexpr int KW = KeyTraits<MODE>::KW;
    uint32_t key[KW];
    if (!make_key<MODE>(f, key)) {
        nokey++;
        return;
    }
    const unsigned long long h = hash64<KW>(key);
    unsigned long long b = f.bytes, pk = f.packets;
    if (p.scale) {  // sum(Bytes*SamplingRate): viz-ch.json:74
        b *= f.sampling_rate;
        pk *= f.sampling_rate;
    }
    if (p.slots) table_add<KW>(p, key, h, b, pk, 1ull);
    if (p.cms) cms_add(p, h, f.bytes * f.sampling_rate);  // viz-ch.json:233 weight
}

// ---- tile staging -------------------------------------------------------------------------

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

struct TileInfo {
    uint32_t r0, n;      // first record, record count
    uint32_t b0, b1;     // stream byte span of the tile
    uint32_t a0;         // b0 rounded down to 16
    bool staged;         // bytes [a0, b1) are in shared memory
};

// Stage the byte span of records [r0, r0+n) into shared memory with coalesced
// 16-byte streaming loads.  Falls back (staged=false) when the span is not sane
// or does not fit.
__device__ __forceinline__ TileInfo stage_tile(const SubmitParams &p, uint32_t tile, uint32_t *smem_words)
{
    TileInfo t;
    t.r0 = tile * kTileRecords;
    t.n = min((uint32_t)kTileRecords, p.n_records - t.r0);
    t.b0 = __ldg(p.offsets + t.r0);
    t.b1 = __ldg(p.offsets + t.r0 + t.n);
    t.a0 = t.b0 & ~15u;
    const unsigned long long end = p.base + p.len;
    t.staged = t.b0 <= t.b1 && t.b0 >= p.base && (unsigned long long)t.b1 <= end && (t.b1 - t.a0) <= (uint32_t)kTileBytes;
    if (t.staged) {
        const uint32_t n16 = (t.b1 - t.a0 + 15u) >> 4;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.buf + ((unsigned long long)t.a0 - p.base));
        uint4 *dst = reinterpret_cast<uint4 *>(smem_words);
        for (uint32_t i = threadIdx.x; i < n16; i += kThreads) dst[i] = ldg_stream(src + i);
    }
    __syncthreads();
    return t;
}

// one record from global memory (tile too large for shared memory, or offsets
// out of order): correct, slow, rare
template <uint32_t NEED>
__device__ __noinline__ bool decode_record_global(const SubmitParams &p, uint32_t o0, uint32_t o1, Flow &f)
{
    const unsigned long long end = p.base + p.len;
    if (o0 > o1 || o0 < p.base || (unsigned long long)o1 > end) return false;
    if (o0 == o1) return p.framed == 0;  // empty bare message decodes to all-zero; empty framed span is bad
    ByteSrc s;
    s.words = reinterpret_cast<const uint32_t *>(p.buf);
    s.limit_word = (uint32_t)(((p.len + 15ull) & ~15ull) / 4ull) - 1u;
    return decode_record<NEED>(s, (uint32_t)(o0 - p.base), (uint32_t)(o1 - p.base), p.framed != 0, f);
}

template <uint32_t NEED>
__device__ __forceinline__ bool decode_tile_record(const SubmitParams &p, const TileInfo &t, const uint32_t *smem_words,
                                                   uint32_t r, Flow &f)
{
    const uint32_t o0 = __ldg(p.offsets + r), o1 = __ldg(p.offsets + r + 1);
    flow_reset(f);
    if (t.staged && o0 >= t.b0 && o1 <= t.b1 && o0 <= o1) {
        SmemSrc s;
        s.words = smem_words;
        return decode_record<NEED>(s, o0 - t.a0, o1 - t.a0, p.framed != 0, f);
    }
    return decode_record_global<NEED>(p, o0, o1, f);
}

__device__ __forceinline__ void flush_counts(const SubmitParams &p, uint32_t bad, uint32_t nokey)
{
    bad = __reduce_add_sync(0xFFFFFFFFu, bad);
    nokey = __reduce_add_sync(0xFFFFFFFFu, nokey);
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&p.counters->n_bad, (unsigned long long)bad);
        if (nokey) atomicAdd(&p.counters->n_nokey, (unsigned long long)nokey);
    }
}

// ---- fused decode + aggregate --------------------------------------------------------------

template <int MODE, bool WEIGHTED>
__global__ void __launch_bounds__(kThreads) k_decode_aggregate(const SubmitParams p, const uint32_t n_tiles)
{
    extern __shared__ __align__(16) uint32_t smem_words[];
    constexpr uint32_t NEED = KeyTraits<MODE>::NEED | F_BYTES | F_PACKETS | (WEIGHTED ? F_SAMPLING_RATE : 0u);
    uint32_t bad = 0, nokey = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileInfo t = stage_tile(p, tile, smem_words);
        if (threadIdx.x < t.n) {
            Flow f;
            if (decode_tile_record<NEED>(p, t, smem_words, t.r0 + threadIdx.x, f)) {
                if (!WEIGHTED) f.sampling_rate = 1;  // unused unless scale/cms, which imply WEIGHTED
                aggregate_flow<MODE>(p, f, nokey);
            } else {
                bad++;  // inserter.go:125-126: log, skip the row
            }
        }
        __syncthreads();  // tile buffer is reused
    }
    flush_counts(p, bad, nokey);
}

// ---- kernel 1 alone: decode to columns -------------------------------------------------------

struct Columns {
    uint8_t *valid;
    unsigned long long *time_received, *time_flow_start, *sampling_rate, *bytes, *packets;
    uint32_t *type, *sequence_num, *src_as, *dst_as, *etype, *proto, *src_port, *dst_port;
    uint4 *src_addr, *dst_addr, *sampler_addr;
    uint8_t *src_addr_len, *dst_addr_len, *sampler_addr_len;
};

__device__ __forceinline__ uint4 addr_bytes(const uint32_t be[4])
{
    // big-endian words back to memory byte order
    return make_uint4(__byte_perm(be[0], 0, 0x0123), __byte_perm(be[1], 0, 0x0123), __byte_perm(be[2], 0, 0x0123),
                      __byte_perm(be[3], 0, 0x0123));
}

__global__ void __launch_bounds__(kThreads) k_decode_columns(const SubmitParams p, const uint32_t n_tiles, const Columns c)
{
    extern __shared__ __align__(16) uint32_t smem_words[];
    uint32_t bad = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileInfo t = stage_tile(p, tile, smem_words);
        if (threadIdx.x < t.n) {
            const uint32_t r = t.r0 + threadIdx.x;
            Flow f;
            const bool ok = decode_tile_record<F_ALL>(p, t, smem_words, r, f);
            if (!ok) {
                flow_reset(f);
                bad++;
            }
            c.valid[r] = ok ? 1 : 0;
            c.time_received[r] = f.time_received;
            c.time_flow_start[r] = f.time_flow_start;
            c.sampling_rate[r] = f.sampling_rate;
            c.bytes[r] = f.bytes;
            c.packets[r] = f.packets;
            c.type[r] = f.type;
            c.sequence_num[r] = f.sequence_num;
            c.src_as[r] = f.src_as;
            c.dst_as[r] = f.dst_as;
            c.etype[r] = f.etype;
            c.proto[r] = f.proto;
            c.src_port[r] = f.src_port;
            c.dst_port[r] = f.dst_port;
            c.src_addr[r] = addr_bytes(f.src);
            c.dst_addr[r] = addr_bytes(f.dst);
            c.sampler_addr[r] = addr_bytes(f.sampler);
            c.src_addr_len[r] = (uint8_t)min(f.src_len, 255u);
            c.dst_addr_len[r] = (uint8_t)min(f.dst_len, 255u);
            c.sampler_addr_len[r] = (uint8_t)min(f.sampler_len, 255u);
        }
        __syncthreads();
    }
    flush_counts(p, bad, 0);
}

// ---- kernel 2 alone: columns -> table / sketch -------------------------------------------------

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_aggregate_columns(const SubmitParams p, const Columns c)
{
    uint32_t nokey = 0;
    for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < p.n_records; r += gridDim.x * kThreads) {
        if (!c.valid[r]) continue;
        Flow f;
        flow_reset(f);
        constexpr uint32_t NEED = KeyTraits<MODE>::NEED;
        f.bytes = c.bytes[r];
        f.packets = c.packets[r];
        f.sampling_rate = c.sampling_rate[r];
        if (NEED & F_TIME_RECEIVED) f.time_received = c.time_received[r];
        if (NEED & F_SRC_AS) f.src_as = c.src_as[r];
        if (NEED & F_DST_AS) f.dst_as = c.dst_as[r];
        if (NEED & F_ETYPE) f.etype = c.etype[r];
        if (NEED & F_PROTO) f.proto = c.proto[r];
        if (NEED & F_SRC_PORT) f.src_port = c.src_port[r];
        if (NEED & F_DST_PORT) f.dst_port = c.dst_port[r];
        if (NEED & F_SRC_ADDR) {
            const uint4 a = c.src_addr[r];
            f.src[0] = __byte_perm(a.x, 0, 0x0123); f.src[1] = __byte_perm(a.y, 0, 0x0123);
            f.src[2] = __byte_perm(a.z, 0, 0x0123); f.src[3] = __byte_perm(a.w, 0, 0x0123);
            f.src_len = c.src_addr_len[r] == 255 ? 17u : c.src_addr_len[r];
        }
        if (NEED & F_DST_ADDR) {
            const uint4 a = c.dst_addr[r];
            f.dst[0] = __byte_perm(a.x, 0, 0x0123); f.dst[1] = __byte_perm(a.y, 0, 0x0123);
            f.dst[2] = __byte_perm(a.z, 0, 0x0123); f.dst[3] = __byte_perm(a.w, 0, 0x0123);
            f.dst_len = c.dst_addr_len[r] == 255 ? 17u : c.dst_addr_len[r];
        }
        aggregate_flow<MODE>(p, f, nokey);
    }
    flush_counts(p, 0, nokey);
}

// ---- flush: occupied slots -> dense rows -------------------------------------------------------

template <int KW>
__global__ void __launch_bounds__(256) k_compact_rows(const uint8_t *slots, uint32_t n_slots, fa_row *rows,
                                                      unsigned long long cap, Counters *counters)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        const uint8_t *s = slots + (size_t)i * SlotLayout<KW>::BYTES;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
        if (w[0] != SLOT_READY) continue;
        const unsigned long long at = atomicAdd(&counters->flush_rows, 1ull);
        if (at >= cap) continue;
        fa_row r;
#pragma unroll
        for (int k = 0; k < FA_MAX_KEY_WORDS; k++) r.key[k] = k < KW ? w[1 + k] : 0u;
        const unsigned long long *v = reinterpret_cast<const unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
        r.bytes = v[0];
        r.packets = v[1];
        r.count = v[2];
        rows[at] = r;
    }
}

// ---- sketch estimate of every group (top-K candidates) -------------------------------------------

template <int KW>
__global__ void __launch_bounds__(256) k_estimate(const uint8_t *slots, uint32_t n_slots, const unsigned long long *cms,
                                                  uint32_t depth, uint32_t wlog2, fa_hh *out, unsigned long long cap,
                                                  Counters *counters)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        const uint8_t *s = slots + (size_t)i * SlotLayout<KW>::BYTES;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
        if (w[0] != SLOT_READY) continue;
        uint32_t key[KW];
#pragma unroll
        for (int k = 0; k < KW; k++) key[k] = w[1 + k];
        const unsigned long long h = hash64<KW>(key);
        const uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32) | 1u, mask = (1u << wlog2) - 1u;
        unsigned long long est = ~0ull;
        for (uint32_t j = 0; j < depth; j++) {
            const unsigned long long cnt = cms[((size_t)j << wlog2) + ((a + j * b) & mask)];
            est = cnt < est ? cnt : est;
        }
        const unsigned long long at = atomicAdd(&counters->flush_rows, 1ull);
        if (at >= cap) continue;
        fa_hh hh;
#pragma unroll
        for (int k = 0; k < FA_MAX_KEY_WORDS; k++) hh.key[k] = k < KW ? key[k] : 0u;
        hh.estimate = est;
        out[at] = hh;
    }
}

}  // namespace fa
