// kernels.cuh -- sm_100a kernels of the flow-aggregation stage.
//
//   k_stream<AggConsumer<MODE,W>>  fused kernel 1 -> kernel 2: length-delimited
//                                  FlowMessage bytes -> group table (+ sketch).  The
//                                  columnar intermediate never touches HBM.
//   k_stream<ColConsumer>          kernel 1 alone: bytes -> 20 decoded columns in HBM
//                                  (inserter.go:142-157 row + create.sh:36-59 columns).
//   k_learn_shape                  the batch's field list for the decoder's lock-step fast path (decode.cuh).
//   k_aggregate_columns<MODE>      kernel 2 alone: columns -> group table (+ sketch).
//   k_table_init / k_merge_hot / k_compact_rows / k_estimate   table reset, hot-replica fold, flush, top-K candidates.
//
// Stream kernel: persistent CTAs (4 per SM, 256 threads), each walking tiles of <= 256 records.  A tile's byte
// span AND its slice of the offsets array are brought into one of the CTA's TWO shared-memory buffers by bulk-async
// copies (cp.async.bulk, the 1-D TMA path: SASS UBLKCP) signalled on an mbarrier, with an L2 evict-first policy so
// the stream does not push the group table out of L2.  While the 8 warps parse tile k out of one buffer (one thread
// = one record), the copy of tile k+1 lands in the other; the warp that finishes reading a buffer LAST re-arms it
// with tile k+2 (its byte bounds were fetched by warp 0 a whole parse earlier and handed over through shared
// memory), so no warp ever waits at a block-wide barrier and no global load sits between two tiles.
//
// The path is integer / memory bound: no tensor cores anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flowagg.h"
#include "decode.cuh"

namespace fa {

constexpr int kThreads = 256;            // one record per thread per tile
constexpr int kTileRecords = kThreads;
constexpr int kTilePad = 64;             // over-read slack behind each tile buffer
constexpr int kTileBytesMax = 112 * 1024; // staged bytes per tile buffer, upper bound (mocker tile of 256: 21.6 KB -> 23 KB)
constexpr int kStreamBlocksPerSM = 4;    // 2 x 23 KB buffers per CTA: 4 CTAs = 32 warps per SM, 64 registers per thread
constexpr int kStreamOffBytes = (kThreads + 8) * 4;  // a tile's slice of the offsets array (n+1 words), per buffer
constexpr int kStreamCtlBytes = 64;      // mbarriers, reader counters, tile bounds behind the buffers

constexpr uint32_t kHotReplicas = 64;  // CTA b uses replica b mod 64
constexpr uint32_t kHotSlots = 1024;   // slots per replica (power of two)
constexpr uint32_t kHotProbes = 8;     // bounded probe sequence; a miss falls through to the main table

struct Counters {
    unsigned long long n_bad, n_nokey, n_dropped, n_groups, flush_rows;
    unsigned long long n_slow;  // records the shape fast path did not decide (parsed by the order-agnostic decoder)
    unsigned int side_state, pad0;
    // key-repetition statistics of the two most recent submits: {lanes whose key repeats inside their warp,
    // lanes looked at}, sampled from every 64th tile.  Submit i decides from what submit i-1 saw.
    unsigned int hint[2][2];
};

struct SubmitParams {
    const uint8_t *buf;       // device bytes; buf[0] is stream byte `base`
    unsigned long long base;  // offsets[] are relative to the stream, buf to base (multiple of 16)
    unsigned long long len;   // bytes readable behind buf (the owner pads the allocation to 16)
    const uint32_t *offsets;  // n_records + 1
    uint32_t n_records;
    uint32_t framed;
    uint32_t lane_shift;    // log2 of the records between neighbouring lanes of a warp (host-chosen, see pick_lane_stride)
    uint32_t tile_records;    // records per tile (multiple of 32, <= kTileRecords)
    uint32_t tile_bytes;      // shared-memory bytes per tile buffer (multiple of 16)
    uint32_t n_tiles;
    uint32_t offsets_aligned;  // offsets is 16-byte aligned: a tile's slice of it can be staged by a bulk copy
    // group table
    uint8_t *slots;
    uint32_t slot_mask;
    uint32_t scale;  // FA_CFG_SCALE_SAMPLING
    // sketch (nullptr = off)
    unsigned long long *cms;
    uint32_t cms_depth, cms_wlog2;
    Counters *counters;
    uint32_t hint_set;  // this submit writes counters->hint[hint_set], reads hint[hint_set ^ 1]
    // hot-key replicas: kHotReplicas small tables of kHotSlots slots (same slot layout as the main table)
    uint8_t *hot_slots;
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

// table / sketch hash; same arithmetic as the checker's restatement
template <int KW>
__host__ __device__ __forceinline__ unsigned long long hash64(const uint32_t *key)
{
    unsigned long long h = 0x243F6A8885A308D3ull;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
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
// Three slot layouts, all with values {u64 bytes, packets, count} behind the key:
//   KW <= 2   32 B  { u64 key | 3 x u64 }            claimed by a 64-bit CAS on the key
//   KW == 4   48 B  { u128 key | 3 x u64 | pad }     claimed by a 128-bit CAS on the key
//   KW == 11  72 B  { u32 state, u32 key[11] | 3 x u64 }   claimed through a state word
// For the CAS layouts the all-ones key marks an empty slot; the one real key that
// is all ones lives in a reserved side slot behind the table (index = capacity).
// One 32-byte sector per record for the AS-pair roll-up.  Lookups are relaxed
// GPU-scope loads served by L2 (no L1 invalidation on the hot path).
template <int KW> struct SlotLayout;
template <> struct SlotLayout<1> { static constexpr uint32_t BYTES = 32, VAL_OFF = 8; };
template <> struct SlotLayout<2> { static constexpr uint32_t BYTES = 32, VAL_OFF = 8; };
template <> struct SlotLayout<4> { static constexpr uint32_t BYTES = 48, VAL_OFF = 16; };
template <> struct SlotLayout<11> { static constexpr uint32_t BYTES = 72, VAL_OFF = 48; };
enum : uint32_t { SLOT_EMPTY = 0, SLOT_BUSY = 1, SLOT_READY = 2 };

__device__ __forceinline__ uint32_t ld_relaxed_u32(const void *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const void *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_relaxed_u128(const void *p, unsigned long long &lo, unsigned long long &hi)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const void *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(void *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void cas_u128(void *p, unsigned long long cmp_lo, unsigned long long cmp_hi, unsigned long long new_lo,
                                         unsigned long long new_hi, unsigned long long &old_lo, unsigned long long &old_hi)
{
    asm volatile(
        "{\n\t.reg .b128 c, s, d;\n\t"
        "mov.b128 c, {%3, %4};\n\t"
        "mov.b128 s, {%5, %6};\n\t"
        "atom.global.relaxed.gpu.cas.b128 d, [%2], c, s;\n\t"
        "mov.b128 {%0, %1}, d;\n\t}"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(p), "l"(cmp_lo), "l"(cmp_hi), "l"(new_lo), "l"(new_hi)
        : "memory");
}

__device__ __forceinline__ void slot_add(uint8_t *vals, unsigned long long bytes, unsigned long long packets, unsigned long long count)
{
    // sum(Bytes), sum(Packets), count(): create.sh:105-107; UInt64 wrap-around is native
    unsigned long long *v = reinterpret_cast<unsigned long long *>(vals);
    red_add_u64(v + 0, bytes);
    red_add_u64(v + 1, packets);
    red_add_u64(v + 2, count);
}

template <int KW>
__device__ __forceinline__ void side_slot_add(const SubmitParams &p, unsigned long long bytes, unsigned long long packets,
                                              unsigned long long count)
{
    uint8_t *s = p.slots + ((size_t)p.slot_mask + 1) * SlotLayout<KW>::BYTES;
    if (atomicCAS(&p.counters->side_state, 0u, 1u) == 0u) atomicAdd(&p.counters->n_groups, 1ull);
    slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
}

template <int KW>
__device__ __forceinline__ void table_add(const SubmitParams &p, const uint32_t *key, unsigned long long h,
                                          unsigned long long bytes, unsigned long long packets, unsigned long long count)
{
    uint32_t slot = (uint32_t)(h >> 32) & p.slot_mask;
    if (KW <= 2) {
        const unsigned long long k = (unsigned long long)key[0] | (KW == 2 ? (unsigned long long)key[1] << 32 : 0ull);
        if (k == ~0ull) return side_slot_add<KW>(p, bytes, packets, count);
#pragma unroll 1
        for (uint32_t probe = 0; probe <= p.slot_mask; probe++) {
            uint8_t *s = p.slots + (size_t)slot * SlotLayout<KW>::BYTES;
            unsigned long long cur = ld_relaxed_u64(s);
            if (cur == ~0ull) {
                cur = atomicCAS(reinterpret_cast<unsigned long long *>(s), ~0ull, k);
                if (cur == ~0ull) {
                    atomicAdd(&p.counters->n_groups, 1ull);
                    cur = k;
                }
            }
            if (cur == k) return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
            slot = (slot + 1) & p.slot_mask;
        }
    } else if (KW == 4) {
        const unsigned long long klo = (unsigned long long)key[0] | ((unsigned long long)key[1] << 32);
        const unsigned long long khi = (unsigned long long)key[2] | ((unsigned long long)key[3] << 32);
        if ((klo & khi) == ~0ull) return side_slot_add<KW>(p, bytes, packets, count);
#pragma unroll 1
        for (uint32_t probe = 0; probe <= p.slot_mask; probe++) {
            uint8_t *s = p.slots + (size_t)slot * SlotLayout<KW>::BYTES;
            unsigned long long clo, chi;
            ld_relaxed_u128(s, clo, chi);  // one 16-byte access: a consistent snapshot of the key
            if ((clo & chi) == ~0ull) {
                cas_u128(s, ~0ull, ~0ull, klo, khi, clo, chi);
                if ((clo & chi) == ~0ull) {
                    atomicAdd(&p.counters->n_groups, 1ull);
                    clo = klo;
                    chi = khi;
                }
            }
            if (clo == klo && chi == khi) return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
            slot = (slot + 1) & p.slot_mask;
        }
    } else {
        // wide keys (5-tuple): EMPTY -> BUSY (CAS) -> key written -> READY (release).  A reader
        // that sees READY through a relaxed load re-reads with acquire only when it has to
        // decide a MISmatch on a key it may have read before it was published.
#pragma unroll 1
        for (uint32_t probe = 0; probe <= p.slot_mask; probe++) {
            uint8_t *s = p.slots + (size_t)slot * SlotLayout<KW>::BYTES;
            uint32_t *state = reinterpret_cast<uint32_t *>(s);
            uint32_t *skey = state + 1;
            uint32_t st = ld_relaxed_u32(state);
            if (st == SLOT_EMPTY) {
                const uint32_t old = atomicCAS(state, (uint32_t)SLOT_EMPTY, (uint32_t)SLOT_BUSY);
                if (old == SLOT_EMPTY) {
#pragma unroll
                    for (int i = 0; i < KW; i++) skey[i] = key[i];
                    st_release_u32(state, SLOT_READY);
                    atomicAdd(&p.counters->n_groups, 1ull);
                    return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
                }
                st = old;
            }
            while (st != SLOT_READY) {  // another thread is publishing this slot's key
                __nanosleep(32);
                st = ld_acquire_u32(state);
            }
            // state READY was observed before these loads are issued (dependent branch above)
            bool same = true;
#pragma unroll
            for (int i = 0; i < KW; i++) same &= (ld_relaxed_u32(skey + i) == key[i]);
            if (same) return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
            slot = (slot + 1) & p.slot_mask;
        }
    }
    atomicAdd(&p.counters->n_dropped, count);
}

// Bounded insert into one hot-key replica (KW <= 4 layouts only).  Returns false when the probe
// sequence is exhausted (replica full of other keys): the caller then uses the main table.
// Does not count groups: replicas are folded into the main table before anything reads it.
template <int KW>
__device__ __forceinline__ bool hot_add(uint8_t *replica, const uint32_t *key, unsigned long long h, unsigned long long bytes,
                                        unsigned long long packets)
{
    uint32_t slot = (uint32_t)(h >> 20) & (kHotSlots - 1u);
    if (KW <= 2) {
        const unsigned long long k = (unsigned long long)key[0] | (KW == 2 ? (unsigned long long)key[1] << 32 : 0ull);
        if (k == ~0ull) return false;
#pragma unroll 1
        for (uint32_t probe = 0; probe < kHotProbes; probe++) {
            uint8_t *s = replica + (size_t)slot * SlotLayout<KW>::BYTES;
            unsigned long long cur = ld_relaxed_u64(s);
            if (cur == ~0ull) {
                cur = atomicCAS(reinterpret_cast<unsigned long long *>(s), ~0ull, k);
                if (cur == ~0ull) cur = k;
            }
            if (cur == k) {
                slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, 1ull);
                return true;
            }
            slot = (slot + 1) & (kHotSlots - 1u);
        }
    } else {
        const unsigned long long klo = (unsigned long long)key[0] | ((unsigned long long)key[1] << 32);
        const unsigned long long khi = (unsigned long long)key[2] | ((unsigned long long)key[3] << 32);
        if ((klo & khi) == ~0ull) return false;
#pragma unroll 1
        for (uint32_t probe = 0; probe < kHotProbes; probe++) {
            uint8_t *s = replica + (size_t)slot * SlotLayout<KW>::BYTES;
            unsigned long long clo, chi;
            ld_relaxed_u128(s, clo, chi);
            if ((clo & chi) == ~0ull) {
                cas_u128(s, ~0ull, ~0ull, klo, khi, clo, chi);
                if ((clo & chi) == ~0ull) {
                    clo = klo;
                    chi = khi;
                }
            }
            if (clo == klo && chi == khi) {
                slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, 1ull);
                return true;
            }
            slot = (slot + 1) & (kHotSlots - 1u);
        }
    }
    return false;
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

// ---- one decoded flow into the group table (+ sketch), in two halves -----------------------------------
//
// begin: key, hash, sketch update, and the LOAD of the first slot of the probe sequence (hot replica or main table);
// finish: compare, then the three reductions -- or, when the first slot is empty or holds another key, the ordinary
// probe loop from the start (CAS claim included).  The stream kernel runs `finish` one tile after `begin`, so the
// L2 round trip of the probe hides behind the parsing of the next record.
template <int KW>
struct Probe {
    static constexpr int KEY64 = KW <= 2 ? 1 : 2;
    uint8_t *slot;                   // first slot of the probe sequence; nullptr = nothing in flight
    unsigned long long k[KEY64];     // the flow's key
    unsigned long long c[KEY64];     // what the slot held
    unsigned long long h, b, pk;
    bool hot;
};

template <int KW>
__device__ __forceinline__ uint8_t *hot_replica(const SubmitParams &p)
{
    return p.hot_slots + (size_t)(blockIdx.x & (kHotReplicas - 1u)) * kHotSlots * SlotLayout<(KW <= 4 ? KW : 1)>::BYTES;
}

// Returns the low hash bits of the key, or 0 with have=false when the flow has no key.
template <int MODE>
__device__ __forceinline__ uint32_t aggregate_begin(const SubmitParams &p, const Flow &f, uint32_t &nokey, bool hot, bool &have,
                                                    Probe<(KeyTraits<MODE>::KW <= 4 ? KeyTraits<MODE>::KW : 1)> &pr)
{
    constexpr int KW = KeyTraits<MODE>::KW;
    uint32_t key[KW];
    have = make_key<MODE>(f, key);
    if (!have) {
        nokey++;
        return 0u;
    }
    const unsigned long long h = hash64<KW>(key);
    unsigned long long b = f.bytes, pk = f.packets;
    if (p.scale) {  // sum(Bytes*SamplingRate): viz-ch.json:74
        b *= f.sampling_rate;
        pk *= f.sampling_rate;
    }
    if (p.cms) cms_add(p, h, f.bytes * f.sampling_rate);  // viz-ch.json:233 weight
    if (p.slots) {
        if (KW > 4) {
            table_add<KW>(p, key, h, b, pk, 1ull);  // wide keys: probed in place (state-word protocol)
        } else {
            constexpr int K4 = KW <= 4 ? KW : 1;
            constexpr int HI = Probe<K4>::KEY64 - 1;
            pr.k[0] = (unsigned long long)key[0] | (K4 >= 2 ? (unsigned long long)key[K4 >= 2 ? 1 : 0] << 32 : 0ull);
            if (K4 == 4) pr.k[HI] = (unsigned long long)key[K4 == 4 ? 2 : 0] | ((unsigned long long)key[K4 == 4 ? 3 : 0] << 32);
            const bool all_ones = K4 == 4 ? (pr.k[0] & pr.k[HI]) == ~0ull : pr.k[0] == ~0ull;
            if (all_ones) {
                side_slot_add<K4>(p, b, pk, 1ull);
            } else {
                uint8_t *s = hot ? hot_replica<K4>(p) + (size_t)((uint32_t)(h >> 20) & (kHotSlots - 1u)) * SlotLayout<K4>::BYTES
                                 : p.slots + (size_t)((uint32_t)(h >> 32) & p.slot_mask) * SlotLayout<K4>::BYTES;
                if (K4 == 4) ld_relaxed_u128(s, pr.c[0], pr.c[HI]);
                else pr.c[0] = ld_relaxed_u64(s);
                pr.slot = s;
                pr.h = h;
                pr.b = b;
                pr.pk = pk;
                pr.hot = hot;
            }
        }
    }
    return (uint32_t)h;
}

// The slot holds the key: the three reductions.  Empty, or another key: the ordinary probe loop from the start (CAS
// claim included).
template <int KW>
__device__ __forceinline__ void aggregate_finish(const SubmitParams &p, Probe<KW> &pr)
{
    if (!pr.slot) return;
    constexpr int HI = Probe<KW>::KEY64 - 1;
    bool same = pr.c[0] == pr.k[0];
    if (KW == 4) same = same && pr.c[HI] == pr.k[HI];
    if (same) {
        slot_add(pr.slot + SlotLayout<KW>::VAL_OFF, pr.b, pr.pk, 1ull);
    } else {
        uint32_t key[KW];
        key[0] = (uint32_t)pr.k[0];
        if (KW >= 2) key[KW >= 2 ? 1 : 0] = (uint32_t)(pr.k[0] >> 32);
        if (KW == 4) {
            key[KW == 4 ? 2 : 0] = (uint32_t)pr.k[HI];
            key[KW == 4 ? 3 : 0] = (uint32_t)(pr.k[HI] >> 32);
        }
        bool done = false;
        if (pr.hot) done = hot_add<KW>(hot_replica<KW>(p), key, pr.h, pr.b, pr.pk);
        if (!done) table_add<KW>(p, key, pr.h, pr.b, pr.pk, 1ull);
    }
    pr.slot = nullptr;
}

// ---- tile staging: one bulk-async copy per tile ----------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (1-D TMA), completion on the mbarrier, L2 evict-first
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}

// Staged span of a tile whose records occupy stream bytes [b0, b1): as much as fits the buffer, starting at b0
// rounded down to 16 (the bulk copy's alignment).  Returns the stream byte one past the staged range (= b0 & ~15
// when nothing can be staged: insane bounds).  Records that end beyond it (oversized tiles, out-of-order offsets)
// are parsed from global memory instead.  Stream positions fit 32 bits (offsets are u32).
__device__ __forceinline__ uint32_t tile_staged_end(const SubmitParams &p, uint32_t b0, uint32_t b1)
{
    const uint32_t a0 = b0 & ~15u, base = (uint32_t)p.base, end = base + (uint32_t)p.len;
    const bool sane = b0 <= b1 && b0 >= base && b1 <= end;
    return a0 + (sane ? min((b1 - a0 + 15u) & ~15u, p.tile_bytes) : 0u);
}

// One thread: stage tile t (records' bytes [b0, b1)) into a buffer -- the bytes AND the tile's slice of the offsets
// array (n+1 words), two bulk copies completing on the buffer's barrier.  The copies move whole 16-byte units; the odd
// tail of the offsets slice (one word for a full tile of 256) is stored by this thread before it arms the barrier.
__device__ __forceinline__ void tile_issue(const SubmitParams &p, uint32_t t, uint32_t b0, uint32_t b1, uint32_t buf, uint32_t *soff,
                                           uint2 *bounds, uint32_t bar)
{
    const uint32_t r0 = t * p.tile_records, n = min(p.tile_records, p.n_records - r0);
    const uint32_t *src = p.offsets + r0;
    const uint32_t s_end = tile_staged_end(p, b0, b1), a0 = b0 & ~15u, nbytes = s_end - a0;
    *bounds = make_uint2(b0, s_end);
    const uint32_t bulk_words = p.offsets_aligned ? ((n + 1u) & ~3u) : 0u;
    for (uint32_t i = bulk_words; i < n; i++) soff[i] = __ldg(src + i);
    soff[n] = b1;
    // the buffer's last readers used ordinary loads; the copies write through the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (nbytes + bulk_words) {
        mbar_expect_tx(bar, nbytes + bulk_words * 4u);
        if (nbytes) bulk_load(buf, p.buf + (a0 - (uint32_t)p.base), nbytes, bar);
        if (bulk_words) bulk_load(smem_u32(soff), src, bulk_words * 4u, bar);
    } else {
        mbar_arrive(bar);
    }
}

// ---- consumers ----------------------------------------------------------------------------------

struct Columns {
    uint8_t *valid;
    unsigned long long *time_received, *time_flow_start, *sampling_rate, *bytes, *packets;
    uint32_t *type, *sequence_num, *src_as, *dst_as, *etype, *proto, *src_port, *dst_port;
    uint4 *src_addr, *dst_addr, *sampler_addr;
    uint8_t *src_addr_len, *dst_addr_len, *sampler_addr_len;
};

struct TileParams {
    SubmitParams p;
    Columns c;         // only read by ColConsumer
    ShapeTable shape;  // the batch's field list, specialised to the consumer's NEED mask (decode.cuh)
};

// ---- hot keys --------------------------------------------------------------------------------------
//
// Real flow data is skewed (the reference's own mocker draws from 9 AS pairs, mocker.go:61-62): with one
// shared slot per key the hottest slots serialise in L2 (measured: 3.5 G flows/s on 18 groups against 23 G
// on 65 536).  Every submit samples how often keys repeat inside a warp (every 64th tile); when the previous
// submit of the context saw >= 1/16 of its lanes repeating, CTAs send their updates to one of 64 small replica
// tables (same fire-and-forget reductions, 64x less contention per address); replicas are folded into the main
// table before anything reads it.  The very first submit of a context assumes hot keys.
template <int MODE, bool WEIGHTED>
struct AggConsumer {
    static constexpr int KW = KeyTraits<MODE>::KW;
    static constexpr int PKW = KW <= 4 ? KW : 1;  // layout of the in-flight probe (unused for wide keys)
    static constexpr uint32_t NEED = KeyTraits<MODE>::NEED | F_BYTES | F_PACKETS | (WEIGHTED ? F_SAMPLING_RATE : 0u);
    static constexpr bool HOT = KW <= 4;                 // 5-tuples are high-cardinality by nature
    static constexpr bool PERMUTE = true;                // nothing is stored per record: lanes may take any record
    struct Item {
        uint32_t h32;
        bool have;
    };
    typedef Probe<PKW> State;
    static __device__ __forceinline__ void state_clear(State &st) { st.slot = nullptr; }
    static __device__ __forceinline__ void item_clear(Item &it)
    {
        it.h32 = 0;
        it.have = false;
    }
    static __device__ __forceinline__ bool want_hot(const SubmitParams &p)
    {
        if (!HOT || !p.hot_slots) return false;
        const unsigned int d = __ldg(&p.counters->hint[p.hint_set ^ 1u][0]), n = __ldg(&p.counters->hint[p.hint_set ^ 1u][1]);
        return n != 0u && d * 16u >= n;  // >= 1/16 of the sampled lanes repeat
    }
    // first half: everything up to the load of the first probe
    static __device__ __forceinline__ void begin(const TileParams &tp, uint32_t, bool ok, Flow &f, uint32_t &bad, uint32_t &nokey, bool hot,
                                                 Item &it, State &st)
    {
        if (ok) {
            if (!WEIGHTED) f.sampling_rate = 1;  // unused unless scale/cms, which imply WEIGHTED
            it.h32 = aggregate_begin<MODE>(tp.p, f, nokey, hot, it.have, st);
        } else {
            bad++;  // inserter.go:125-126: log, skip the row
        }
    }
    static __device__ __forceinline__ void finish(const TileParams &tp, State &st) { aggregate_finish<PKW>(tp.p, st); }
    // every 64th tile measures how often keys repeat inside a warp (feeds the next submit's decision)
    static __device__ __forceinline__ void sample_repeats(const SubmitParams &p, uint32_t tile, const Item &it)
    {
        if (!HOT || (tile & 63u) != 0u) return;
        const uint32_t h32 = it.have ? it.h32 : (0x9E3779B9u * (threadIdx.x + 1u));
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h32);
        const uint32_t dups = __popc(__ballot_sync(0xFFFFFFFFu, it.have && __popc(peers) > 1));
        const uint32_t lanes = __popc(__ballot_sync(0xFFFFFFFFu, it.have));
        if ((threadIdx.x & 31) == 0 && lanes) {
            atomicAdd(&p.counters->hint[p.hint_set][0], dups);
            atomicAdd(&p.counters->hint[p.hint_set][1], lanes);
        }
    }
};

__device__ __forceinline__ uint4 addr_bytes(const uint32_t be[4])
{
    // big-endian words back to memory byte order
    return make_uint4(__byte_perm(be[0], 0, 0x0123), __byte_perm(be[1], 0, 0x0123), __byte_perm(be[2], 0, 0x0123),
                      __byte_perm(be[3], 0, 0x0123));
}

struct ColConsumer {
    static constexpr uint32_t NEED = F_ALL;
    static constexpr bool PERMUTE = false;  // column stores stay coalesced: lane i writes row r0+i
    struct Item {};
    struct State {};
    static __device__ __forceinline__ void state_clear(State &) {}
    static __device__ __forceinline__ void item_clear(Item &) {}
    static __device__ __forceinline__ bool want_hot(const SubmitParams &) { return false; }
    static __device__ __forceinline__ void sample_repeats(const SubmitParams &, uint32_t, const Item &) {}
    static __device__ __forceinline__ void finish(const TileParams &, State &) {}
    static __device__ __forceinline__ void begin(const TileParams &tp, uint32_t r, bool ok, Flow &f, uint32_t &bad, uint32_t &, bool, Item &, State &)
    {
        const Columns &c = tp.c;
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
};

// A record the shape fast path did not decide, parsed by the order-agnostic decoder from the staged tile and consumed
// on the spot.  Out of line with its own Flow, so the fast path keeps its Flow in registers.  Returns bad | nokey << 1.
template <class Consumer>
__device__ __noinline__ uint32_t record_from_tile(const TileParams &tp, uint32_t r, uint32_t smem_base, uint32_t p0, uint32_t p1, bool hot)
{
    uint32_t bad = 0, nokey = 0;
    Flow f;
    flow_reset(f);
    SmemSrc s;
    s.base = smem_base;
    const bool ok = decode_record<Consumer::NEED>(s, p0, p1, tp.p.framed != 0, f);
    typename Consumer::Item it;
    typename Consumer::State st;
    Consumer::item_clear(it);
    Consumer::state_clear(st);
    Consumer::begin(tp, r, ok, f, bad, nokey, hot, it, st);
    Consumer::finish(tp, st);
    return bad | (nokey << 1);
}

// one record straight from global memory (it did not fit the staged tile, or its
// offsets are out of order): correct, slow, rare.  Returns bad | nokey << 1.
template <class Consumer>
__device__ __noinline__ uint32_t record_from_global(const TileParams &tp, uint32_t r, uint32_t o0, uint32_t o1)
{
    uint32_t bad = 0, nokey = 0;
    const SubmitParams &p = tp.p;
    const unsigned long long end = p.base + p.len;
    Flow f;
    flow_reset(f);
    bool ok;
    if (o0 > o1 || o0 < p.base || (unsigned long long)o1 > end) {
        ok = false;
    } else if (o0 == o1) {
        ok = p.framed == 0;  // empty bare message decodes to all-zero; an empty framed span is bad
    } else {
        ByteSrc s;
        s.words = reinterpret_cast<const uint32_t *>(p.buf);
        s.limit_word = (uint32_t)(((p.len + 15ull) & ~15ull) / 4ull) - 1u;
        ok = decode_record<Consumer::NEED>(s, (uint32_t)(o0 - p.base), (uint32_t)(o1 - p.base), p.framed != 0, f);
    }
    typename Consumer::Item it;
    typename Consumer::State st;
    Consumer::item_clear(it);
    Consumer::state_clear(st);
    Consumer::begin(tp, r, ok, f, bad, nokey, false, it, st);
    Consumer::finish(tp, st);
    return bad | (nokey << 1);
}

__device__ __forceinline__ void flush_counts(const SubmitParams &p, uint32_t bad, uint32_t nokey, uint32_t slow)
{
    bad = __reduce_add_sync(0xFFFFFFFFu, bad);
    nokey = __reduce_add_sync(0xFFFFFFFFu, nokey);
    slow = __reduce_add_sync(0xFFFFFFFFu, slow);
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&p.counters->n_bad, (unsigned long long)bad);
        if (nokey) atomicAdd(&p.counters->n_nokey, (unsigned long long)nokey);
        if (slow) atomicAdd(&p.counters->n_slow, (unsigned long long)slow);
    }
}

// Which record of the tile does this thread parse?  Lanes of a warp read their records from shared memory in
// lock step, so the bank pattern is set by the byte distance between the records of neighbouring lanes.  With
// consecutive records and near-constant record sizes that distance can resonate with the 32 x 4-byte banks
// (measured: 86-byte records = 21.5 words, 3 lanes apart = 64.5 words -> every third lane on one bank).  Each
// group of d = 2^shift warps therefore shares a run of 32*d records, neighbouring lanes taking records d apart;
// d is chosen per batch from its mean record size (pick_lane_stride, host side).  A bijection on [0, kThreads).
__device__ __forceinline__ uint32_t record_of_thread(uint32_t shift)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    return ((warp >> shift) << (5u + shift)) + (lane << shift) + (warp & ((1u << shift) - 1u));
}

// ---- the stream kernel: decode (+ consume) tiles blockIdx.x, blockIdx.x + gridDim.x, ... ---------------------
//
// Shared memory: two tile buffers of tile_bytes + kTilePad each, two offsets slices of kStreamOffBytes, then
// {full[2] mbarriers, readers_done[2], bounds[2], next_bounds[2][2]}.

template <class Consumer>
__global__ void __launch_bounds__(kThreads, kStreamBlocksPerSM) k_stream(const __grid_constant__ TileParams tp)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const SubmitParams &p = tp.p;
    const uint32_t buf_stride = p.tile_bytes + kTilePad;
    uint8_t *offs_area = smem + 2u * buf_stride;                             // two offsets slices of kStreamOffBytes
    uint8_t *ctl = offs_area + 2u * kStreamOffBytes;
    const uint32_t bar0 = smem_u32(ctl);                                     // full[b] at bar0 + 8 b
    unsigned int *readers_done = reinterpret_cast<unsigned int *>(ctl + 16); // [2]
    uint2 *bounds = reinterpret_cast<uint2 *>(ctl + 32);                     // [2]: {b0, s_end} of the tile in buffer b
    volatile uint32_t *next_bounds = reinterpret_cast<volatile uint32_t *>(ctl + 48);  // [2][2]: byte bounds of the tile to stage next in buffer b
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_tiles = p.n_tiles, step = gridDim.x;
    const uint32_t n_warps = kThreads / 32;

    // byte bounds of tile t's records: which = 0 the first record's start, 1 the last record's end
    auto tile_bound = [&](uint32_t t, uint32_t which) {
        const uint32_t r0 = t * p.tile_records, n = min(p.tile_records, p.n_records - r0);
        return __ldg(p.offsets + r0 + (which ? n : 0u));
    };
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        readers_done[0] = readers_done[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint32_t k = 0; k < 2u; k++) {  // the first two tiles of this CTA
            const uint32_t t = blockIdx.x + k * step;
            if (t < n_tiles)
                tile_issue(p, t, tile_bound(t, 0), tile_bound(t, 1), smem0 + k * buf_stride,
                           reinterpret_cast<uint32_t *>(offs_area + k * kStreamOffBytes), bounds + k, bar0 + 8u * k);
        }
    }
    __syncthreads();

    const bool hot = Consumer::want_hot(p);  // uniform over the grid
    const bool framed = p.framed != 0;
    uint32_t bad = 0, nokey = 0, slow = 0;
    uint32_t in_tile = threadIdx.x;
    if (Consumer::PERMUTE && p.lane_shift) in_tile = record_of_thread(p.lane_shift);
    typename Consumer::State state;  // the previous tile's record: its first probe is in flight during this tile's parse
    Consumer::state_clear(state);

    uint32_t k = 0;
#pragma unroll 1
    for (uint32_t t = blockIdx.x; t < n_tiles; t += step, k++) {
        const uint32_t b = k & 1u;
        const uint32_t r0 = t * p.tile_records, n = min(p.tile_records, p.n_records - r0);
        // short tails keep the identity mapping (the permutation is a bijection on full tiles only)
        const uint32_t mine = n == (uint32_t)kThreads ? in_tile : threadIdx.x;
        const bool active = mine < n;
        const uint32_t r = r0 + mine;
        const uint32_t *soff = reinterpret_cast<const uint32_t *>(offs_area + b * kStreamOffBytes);
        // warp 0 fetches the byte bounds of the tile that will follow this one in buffer b (lane 0: begin, lane 1: end);
        // they travel through shared memory to whichever warp re-arms the buffer, a whole parse later
        const bool refill = t + 2u * step < n_tiles;
        uint32_t nbound = 0;
        if (refill && threadIdx.x < 2u) nbound = tile_bound(t + 2u * step, threadIdx.x);

        mbar_wait(bar0 + 8u * b, (k >> 1) & 1u);
        const uint2 bb = bounds[b];  // {b0, s_end}
        const uint32_t o0 = active ? soff[mine] : 0u, o1 = active ? soff[mine + 1u] : 0u;
        const uint32_t buf = smem0 + b * buf_stride, a0 = bb.x & ~15u;
        typename Consumer::Item item;
        Consumer::item_clear(item);
        // every lane walks the field list (lanes without a staged record walk an empty span); cursors are absolute
        // shared-window addresses (the buffers are 16-byte aligned, so alignment arithmetic is unchanged)
        const bool from_tile = active && o0 >= bb.x && o0 <= o1 && o1 <= bb.y;
        const uint32_t p0 = buf + (from_tile ? o0 - a0 : 0u), p1 = buf + (from_tile ? o1 - a0 : 0u);
        Flow f;
        flow_reset(f);
        SmemSrc s;
        s.base = 0u;
        const bool fast = decode_record_shape<Consumer::NEED>(tp.shape, s, p0, p1, framed, f) && from_tile;
        if (!fast && active) {  // parsed by the order-agnostic decoder and consumed on the spot
            uint32_t res;
            if (from_tile) {
                slow++;
                res = record_from_tile<Consumer>(tp, r, 0u, p0, p1, hot);
            } else {
                res = record_from_global<Consumer>(tp, r, o0, o1);
            }
            bad += res & 1u;
            nokey += res >> 1;
        }
        // this warp is done with buffer b: count it (the answer -- am I the last? -- is looked at after the table work)
        if (refill && threadIdx.x < 2u) next_bounds[2u * b + threadIdx.x] = nbound;
        __syncwarp();
        unsigned int seen = 0;
        if (lane == 0) {
            __threadfence_block();
            seen = atomicAdd(&readers_done[b], 1u);
        }
        Consumer::finish(tp, state);  // the previous tile's record: its probe has had a whole parse to come back
        if (fast) Consumer::begin(tp, r, true, f, bad, nokey, hot, item, state);
        Consumer::sample_repeats(p, t, item);
        // buffer b is free once all warps are past it; the last one re-arms it with tile k+2 of this CTA
        if (lane == 0 && seen == n_warps - 1u) {
            readers_done[b] = 0u;
            if (refill)
                tile_issue(p, t + 2u * step, next_bounds[2u * b], next_bounds[2u * b + 1u], buf,
                           reinterpret_cast<uint32_t *>(offs_area + b * kStreamOffBytes), bounds + b, bar0 + 8u * b);
        }
    }
    Consumer::finish(tp, state);
    flush_counts(p, bad, nokey, slow);
}

// ---- the batch's field list ------------------------------------------------------------------------------
//
// 256 records spread over the batch are walked (shape_collect, decode.cuh) and every tag a fast-path step could
// stand for is marked; the ascending list of marked tag values goes to (pinned) host memory, where the next launches
// turn it into their ShapeTable.  Purely a speed matter: results never depend on the table.
struct ShapeLearned {
    uint32_t n;      // tag values written (0 when more than kShapeMax distinct tags were seen: no fast path)
    uint32_t total;  // distinct tags seen
    uint16_t tagval[kShapeMax];
};

struct MarkTag {
    unsigned int *bits;  // three bitmaps of 2^14 bits: tag seen | with a 1..4-byte varint | with a 5-byte varint
    __device__ __forceinline__ void operator()(uint32_t tagval, uint32_t vb) const
    {
        const uint32_t w = (tagval >> 5) & 511u, bit = 1u << (tagval & 31u);
        atomicOr(&bits[w], bit);
        if (vb >= 1u && vb <= 4u) atomicOr(&bits[512u + w], bit);
        if (vb == 5u) atomicOr(&bits[1024u + w], bit);
    }
};

__global__ void __launch_bounds__(256) k_learn_shape(const SubmitParams p, ShapeLearned *out)
{
    __shared__ unsigned int bits[3 * 512];  // one bit per tag value < 2^14, three times (MarkTag)
    for (uint32_t i = threadIdx.x; i < 3u * 512u; i += blockDim.x) bits[i] = 0u;
    __syncthreads();
    const uint32_t n = p.n_records;
    const uint32_t r = n >= 256u ? (uint32_t)(((unsigned long long)threadIdx.x * n) >> 8) : threadIdx.x;
    if (r < n) {
        const uint32_t o0 = __ldg(p.offsets + r), o1 = __ldg(p.offsets + r + 1);
        if (o0 < o1 && o0 >= p.base && (unsigned long long)o1 <= p.base + p.len) {
            ByteSrc s;
            s.words = reinterpret_cast<const uint32_t *>(p.buf);
            s.limit_word = (uint32_t)(((p.len + 15ull) & ~15ull) / 4ull) - 1u;
            shape_collect(s, (uint32_t)(o0 - p.base), (uint32_t)(o1 - p.base), p.framed != 0, MarkTag{bits});
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (uint32_t w = 0; w < 512u; w++) {
            unsigned int v = bits[w];
            while (v) {
                const uint32_t bit = __ffs((int)v) - 1u;
                v &= v - 1u;
                if (m < kShapeMax)
                    out->tagval[m] = (uint16_t)((w * 32u + bit) | (((bits[512u + w] >> bit) & 1u) ? kTagSaw4 : 0u) |
                                                (((bits[1024u + w] >> bit) & 1u) ? kTagSaw5 : 0u));
                m++;
            }
        }
        out->total = m;
        out->n = m <= kShapeMax ? m : 0u;
        __threadfence_system();
    }
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
        bool have;
        Probe<(KeyTraits<MODE>::KW <= 4 ? KeyTraits<MODE>::KW : 1)> pr;
        pr.slot = nullptr;
        aggregate_begin<MODE>(p, f, nokey, false, have, pr);
        aggregate_finish(p, pr);
    }
    flush_counts(p, 0, nokey, 0);
}

// ---- table reset / flush / top-K candidates ----------------------------------------------------------

// empty table: all-ones keys (CAS layouts) or state 0 (wide keys), zero values; n_slots includes the side slot
template <int KW>
__global__ void __launch_bounds__(256) k_table_init(uint8_t *slots, unsigned long long n_slots)
{
    constexpr uint32_t WORDS = SlotLayout<KW>::BYTES / 8;
    constexpr uint32_t KEY_WORDS64 = KW <= 2 ? 1 : (KW == 4 ? 2 : 0);
    unsigned long long *w = reinterpret_cast<unsigned long long *>(slots);
    const unsigned long long total = n_slots * WORDS;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x)
        w[i] = (i % WORDS) < KEY_WORDS64 ? ~0ull : 0ull;
}

// Fold the hot-key replicas into the main table and empty them (KW <= 4 layouts).
template <int KW>
__global__ void __launch_bounds__(256) k_merge_hot(const SubmitParams p, uint32_t n_hot_slots)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hot_slots; i += gridDim.x * blockDim.x) {
        uint8_t *s = p.hot_slots + (size_t)i * SlotLayout<KW>::BYTES;
        unsigned long long *w = reinterpret_cast<unsigned long long *>(s);
        uint32_t key[KW];
        constexpr int KEY64 = KW <= 2 ? 1 : 2;
        bool empty = true;
#pragma unroll
        for (int k = 0; k < KEY64; k++) empty &= (w[k] == ~0ull);
        if (empty) continue;
        key[0] = (uint32_t)w[0];
        if (KW >= 2) key[KW >= 2 ? 1 : 0] = (uint32_t)(w[0] >> 32);
        if (KW == 4) {
            key[KW == 4 ? 2 : 0] = (uint32_t)w[KEY64 - 1];
            key[KW == 4 ? 3 : 0] = (uint32_t)(w[KEY64 - 1] >> 32);
        }
        const unsigned long long *v = reinterpret_cast<const unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
        table_add<KW>(p, key, hash64<KW>(key), v[0], v[1], v[2]);
#pragma unroll
        for (int k = 0; k < KEY64; k++) w[k] = ~0ull;
        unsigned long long *vv = reinterpret_cast<unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
        vv[0] = vv[1] = vv[2] = 0ull;
    }
}

// is_side: the reserved slot behind the table (the all-ones key); it is occupied iff side_claimed (counters->side_state,
// which is what n_groups counted) -- its count alone would miss a key whose merged count is 0.
template <int KW>
__device__ __forceinline__ bool slot_read(const uint8_t *s, bool is_side, bool side_claimed, uint32_t *key)
{
    if (KW <= 2) {
        const unsigned long long k = *reinterpret_cast<const unsigned long long *>(s);
        key[0] = (uint32_t)k;
        if (KW == 2) key[1] = (uint32_t)(k >> 32);
        const unsigned long long cnt = reinterpret_cast<const unsigned long long *>(s + SlotLayout<KW>::VAL_OFF)[2];
        (void)cnt;
        return is_side ? side_claimed : k != ~0ull;
    } else if (KW == 4) {
        const unsigned long long lo = reinterpret_cast<const unsigned long long *>(s)[0], hi = reinterpret_cast<const unsigned long long *>(s)[1];
        key[0] = (uint32_t)lo; key[1] = (uint32_t)(lo >> 32); key[2] = (uint32_t)hi; key[3] = (uint32_t)(hi >> 32);
        const unsigned long long cnt = reinterpret_cast<const unsigned long long *>(s + SlotLayout<KW>::VAL_OFF)[2];
        (void)cnt;
        return is_side ? side_claimed : (lo & hi) != ~0ull;
    } else {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
#pragma unroll
        for (int k = 0; k < KW; k++) key[k] = w[1 + k];
        return !is_side && w[0] == SLOT_READY;
    }
}

template <int KW>
__global__ void __launch_bounds__(256) k_compact_rows(const uint8_t *slots, unsigned long long n_slots, fa_row *rows,
                                                      unsigned long long cap, Counters *counters)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint8_t *s = slots + i * SlotLayout<KW>::BYTES;
        uint32_t key[KW];
        if (!slot_read<KW>(s, i == n_slots - 1, counters->side_state != 0u, key)) continue;
        if (i == n_slots - 1) {  // the side slot holds the all-ones key
#pragma unroll
            for (int k = 0; k < KW; k++) key[k] = 0xFFFFFFFFu;
        }
        const unsigned long long at = atomicAdd(&counters->flush_rows, 1ull);
        if (at >= cap) continue;
        fa_row r;
#pragma unroll
        for (int k = 0; k < FA_MAX_KEY_WORDS; k++) r.key[k] = k < KW ? key[k] : 0u;
        const unsigned long long *v = reinterpret_cast<const unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
        r.bytes = v[0];
        r.packets = v[1];
        r.count = v[2];
        rows[at] = r;
    }
}

// Which of n_owners contexts merges a key in a box-wide roll-up.  The table index uses hi32(h), this uses lo32(h).
__host__ __device__ __forceinline__ uint32_t key_owner(unsigned long long h, uint32_t n_owners) { return (uint32_t)h % n_owners; }

// SummingMergeTree's merge step (create.sh:88-90) for rows that are already aggregates: another context's
// flush output, possibly read straight from a peer GPU's memory over NVLink.  n_owners > 1 keeps only the
// rows this context owns (the hash-partitioned exchange of fa_flush_box): the key is read first and the values
// only for owned rows.
template <int KW>
__global__ void __launch_bounds__(256) k_add_rows(const SubmitParams p, const fa_row *rows, unsigned long long n, uint32_t owner,
                                                  uint32_t n_owners)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t key[KW];
#pragma unroll
        for (int k = 0; k < KW; k++) key[k] = rows[i].key[k];
        const unsigned long long h = hash64<KW>(key);
        if (n_owners > 1u && key_owner(h, n_owners) != owner) continue;
        table_add<KW>(p, key, h, rows[i].bytes, rows[i].packets, rows[i].count);
    }
}

// sketch estimate of every group (top-K candidates)
template <int KW>
__global__ void __launch_bounds__(256) k_estimate(const uint8_t *slots, unsigned long long n_slots, const unsigned long long *cms,
                                                  uint32_t depth, uint32_t wlog2, fa_hh *out, unsigned long long cap,
                                                  Counters *counters)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint8_t *s = slots + i * SlotLayout<KW>::BYTES;
        uint32_t key[KW];
        if (!slot_read<KW>(s, i == n_slots - 1, counters->side_state != 0u, key)) continue;
        if (i == n_slots - 1) {
#pragma unroll
            for (int k = 0; k < KW; k++) key[k] = 0xFFFFFFFFu;
        }
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
