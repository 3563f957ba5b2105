// kernels.cuh -- sm_100a kernels of the flow-aggregation stage.
//
//   k_tile<AggConsumer<MODE,W>>  fused kernel 1 -> kernel 2: length-delimited
//                                FlowMessage bytes -> group table (+ sketch).  The
//                                columnar intermediate never touches HBM.
//   k_tile<ColConsumer>          kernel 1 alone: bytes -> 20 decoded columns in HBM
//                                (inserter.go:142-157 row + create.sh:36-59 columns).
//   k_aggregate_columns<MODE>    kernel 2 alone: columns -> group table (+ sketch).
//   k_table_init / k_merge_hot / k_compact_rows / k_estimate   table reset, hot-replica fold, flush, top-K candidates.
//
// Tile kernel: one CTA = one tile of <= 256 records.  The tile's byte span is
// brought into shared memory by ONE bulk-async copy (cp.async.bulk, the 1-D TMA
// path: SASS UBLKCP) signalled on an mbarrier, with an L2 evict-first policy so the
// stream does not push the group table out of L2; then one thread parses one
// record from shared memory.  Up to 8 CTAs are resident per SM, so copies of some
// tiles overlap the parsing of others without any intra-CTA pipeline.
//
// The path is integer / memory bound: no tensor cores anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flowagg.h"
#include "decode.cuh"

namespace fa {

constexpr int kThreads = 256;            // one record per thread per tile
constexpr int kTilePad = 64;             // over-read slack behind the tile (then the mbarrier)
constexpr int kTileBytesMax = 112 * 1024; // staged bytes per tile, upper bound (mocker tile of 256: 21.6 KB -> 24 KB)

constexpr uint32_t kHotReplicas = 64;  // CTA b uses replica b mod 64
constexpr uint32_t kHotSlots = 1024;   // slots per replica (power of two)
constexpr uint32_t kHotProbes = 8;     // bounded probe sequence; a miss falls through to the main table
constexpr uint32_t kCandHotSlots = 8192;  // FA_CFG_TOPK_ONLY: slots per replica (twice the candidates a top-1000 context keeps)

// per group table (a context has two: the one being filled and the one being flushed)
struct TableState {
    unsigned long long n_groups, n_dropped, flush_rows;
    unsigned int side_state, pad0;
};

// Context-wide counters, laid out by who touches them WHILE A LAUNCH RUNS: a load from an L2 line that atomics are queueing
// on waits behind them (ncu, configs[2]: every CTA's prologue load of total_weight sat ~13 k cycles behind the per-warp
// atomics on total_weight_acc next to it -- 24 % of the kernel's stall samples), so what a launch reads and what it hammers
// never share a 128-byte line.
struct Counters {
    // line 0: rare atomics (bad records; every 64th tile's statistics), read by every CTA's prologue (hint)
    unsigned long long n_bad, n_nokey;
    // key-repetition statistics of the two most recent submits: {lanes whose key repeats inside their warp,
    // lanes looked at}, sampled from every 64th tile.  Submit i decides from what submit i-1 saw.
    unsigned int hint[2][2];
    unsigned long long pad0[12];
    // line 1: FA_CFG_TOPK_ONLY: sum of the sketched weights the admission threshold of the RUNNING launch scales with --
    // read by every CTA, written only between launches (k_publish_weight)
    unsigned long long total_weight;
    unsigned long long pad1[15];
    // line 2: the running launch adds its own weights here (one atomic per warp); nothing reads it until the launch is over
    unsigned long long total_weight_acc;
    unsigned long long pad2[15];
};
static_assert(sizeof(Counters) == 384, "three 128-byte lines");

struct SubmitParams {
    const uint8_t *buf;       // device bytes; buf[0] is stream byte `base`
    unsigned long long base;  // offsets[] are relative to the stream, buf to base (multiple of 16)
    unsigned long long len;   // bytes readable behind buf (the owner pads the allocation to 16)
    const uint32_t *offsets;  // n_records + 1
    uint32_t n_records;
    uint32_t framed;
    uint32_t lane_shift;    // log2 of the records between neighbouring lanes of a warp (host-chosen, see pick_lane_stride)
    uint32_t tile_records;    // records per CTA tile (<= blockDim.x)
    uint32_t tile_bytes;      // shared-memory bytes for the tile (multiple of 16); barrier sits behind it
    // group table
    uint8_t *slots;
    uint32_t slot_mask;
    uint32_t scale;  // FA_CFG_SCALE_SAMPLING
    // sketch (nullptr = off)
    unsigned long long *cms;
    uint32_t cms_depth, cms_wlog2;
    uint32_t hot_mask;     // slots per hot replica - 1 (kHotSlots - 1; FA_CFG_TOPK_ONLY uses kCandHotSlots - 1)
    uint32_t admit_shift;  // FA_CFG_TOPK_ONLY: a key enters the (bounded) candidate table only once its sketch estimate
                           // reaches total_weight >> admit_shift; 0 = every key is a candidate (the exact group table)
    Counters *counters;
    TableState *tstate;  // of the table behind `slots`
    uint32_t hint_set;  // this submit writes counters->hint[hint_set], reads hint[hint_set ^ 1]
    uint32_t sample_seed;  // FA_CFG_TOPK_ONLY: distinguishes the submits in the admission draw (a function of the submit's number)
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

// Where a key lives in its own group table (and hot replica).  Nothing outside the library sees this value -- the sketch rows
// and the box-wide owner of a key come from hash64 itself, which the checker restates -- so it could be cheaper than hash64 for
// narrow keys; a 6-instruction multiplicative mix measured 3 % SLOWER than the 40-instruction hash64 on configs[1]
// (profiles/r02/experiments), so it is hash64.
template <int KW>
__device__ __forceinline__ unsigned long long slot_hash(const uint32_t *key)
{
    return hash64<KW>(key);
}

// ---- group table: open addressing, linear probing, in HBM (L2-resident when small) ---
//
// Three slot layouts, all with values {u64 bytes, packets, count} behind the key:
//   KW <= 2   32 B  { u64 key | 3 x u64 }            claimed by a 64-bit CAS on the key
//   KW == 4   48 B  { u128 key | 3 x u64 | pad }     claimed by a 128-bit CAS on the key
//   KW == 11  split: a 32-byte HEAD per slot { u64 head = fingerprint << 32 | state | 3 x u64 } in one array and,
//             behind all heads, a 64-byte KEY record per slot { u32 key[11], zero pad } in a second, 64-byte-aligned array
// For the CAS layouts the all-ones key marks an empty slot; the one real key that
// is all ones lives in a reserved side slot behind the table (index = capacity).
// One 32-byte sector per record for the AS-pair roll-up.  Lookups are relaxed
// GPU-scope loads served by L2 (no L1 invalidation on the hot path).
//
// The wide-key table is sized for HBM, not L2 (BASELINE configs[4]: 100 M distinct 5-tuples), so it is laid out by DRAM
// sector: a probe reads ONE sector (the head); the fingerprint (the hash bits the slot index does not use) settles a
// mismatch without touching the key; a claim writes the key record as two 256-bit stores, one whole sector each
// (no read-for-ownership: measured, four 16-byte stores still made L2 fetch both sectors), and the three sums land in the head sector the probe already brought into L2.
// Insert = 32 B read + 96 B written back; a repeated key = head + key record read.  (Round 1's 72-byte slot straddled
// three sectors and cost ~6 sector reads + 3 write-backs per insert, ncu: profiles/r02/experiments.)
template <int KW> struct SlotLayout;
template <> struct SlotLayout<1> { static constexpr uint32_t BYTES = 32, VAL_OFF = 8; };
template <> struct SlotLayout<2> { static constexpr uint32_t BYTES = 32, VAL_OFF = 8; };
template <> struct SlotLayout<4> { static constexpr uint32_t BYTES = 48, VAL_OFF = 16; };
template <> struct SlotLayout<11> { static constexpr uint32_t BYTES = 32, VAL_OFF = 8; };  // the head array; keys: wide_key_*
constexpr uint32_t kWideKeyBytes = 64;
// byte offset of the key array behind n_slots heads (n_slots counts the side slot, which wide keys never use)
__host__ __device__ __forceinline__ unsigned long long wide_key_offset(unsigned long long n_slots) { return (n_slots * 32ull + 127ull) & ~127ull; }
// allocation size of a table of n_slots slots
__host__ __device__ __forceinline__ unsigned long long table_alloc_bytes(int kw, unsigned long long n_slots)
{
    if (kw <= 2) return n_slots * 32ull;
    if (kw == 4) return n_slots * 48ull;
    return wide_key_offset(n_slots) + n_slots * kWideKeyBytes;
}
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
__device__ __forceinline__ unsigned long long ld_acquire_u64(const void *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(void *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// one 256-bit store (sm_100: SASS STG.E.ENL2.256): a whole, aligned 32-byte sector in a single request, so L2 has
// nothing to merge and nothing to fetch before it can write the sector back
__device__ __forceinline__ void st_sector(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h)
{
    const unsigned long long q0 = (unsigned long long)a | ((unsigned long long)b << 32), q1 = (unsigned long long)c | ((unsigned long long)d << 32);
    const unsigned long long q2 = (unsigned long long)e | ((unsigned long long)f << 32), q3 = (unsigned long long)g | ((unsigned long long)h << 32);
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(q0), "l"(q1), "l"(q2), "l"(q3) : "memory");
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
    if (atomicCAS(&p.tstate->side_state, 0u, 1u) == 0u) atomicAdd(&p.tstate->n_groups, 1ull);
    slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
}

// OUTLINE: the call sits inside a __noinline__ device function.  ptxas 12.9 narrows a 256-bit vector store to its first
// element there (SASS: STG.E.64 instead of STG.E.ENL2.256 -- found by a failing parity test, pinned by
// tests/test_host_logic.py::test_key_record_stores_are_whole_sectors), so that one, rare, path keeps four 16-byte stores.
template <int KW, bool OUTLINE = false>
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
                    atomicAdd(&p.tstate->n_groups, 1ull);
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
                    atomicAdd(&p.tstate->n_groups, 1ull);
                    clo = klo;
                    chi = khi;
                }
            }
            if (clo == klo && chi == khi) return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
            // PTX does not promise that a 16-byte vector load is one atomic access: a snapshot that LOOKS half written (one
            // half ours, or one half still all ones) is re-read through the CAS unit, whose answer is atomic, before the probe
            // moves on -- a torn read must never send a key to a second slot
            if (clo == klo || chi == khi || clo == ~0ull || chi == ~0ull) {
                cas_u128(s, ~0ull, ~0ull, klo, khi, clo, chi);
                if ((clo & chi) == ~0ull) {  // it was empty after all and is ours now
                    atomicAdd(&p.tstate->n_groups, 1ull);
                    clo = klo;
                    chi = khi;
                }
                if (clo == klo && chi == khi) return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
            }
            slot = (slot + 1) & p.slot_mask;
        }
    } else {
        // wide keys (5-tuple): head EMPTY(0) -> fingerprint|BUSY (CAS) -> key record written -> fingerprint|READY (release)
        const unsigned long long fp = h << 32;  // lo32(h): the slot index comes from hi32(h)
        uint8_t *keys = p.slots + wide_key_offset((unsigned long long)p.slot_mask + 2ull);
        uint32_t kk[11];
#pragma unroll
        for (int i = 0; i < 11; i++) kk[i] = i < KW ? key[i < KW ? i : 0] : 0u;
#pragma unroll 1
        for (uint32_t probe = 0; probe <= p.slot_mask; probe++) {
            uint8_t *s = p.slots + (size_t)slot * SlotLayout<KW>::BYTES;
            uint8_t *kr = keys + (size_t)slot * kWideKeyBytes;
            unsigned long long hd = ld_relaxed_u64(s);
            if (hd == 0ull) {
                hd = atomicCAS(reinterpret_cast<unsigned long long *>(s), 0ull, fp | SLOT_BUSY);
                if (hd == 0ull) {
                    // ours: two whole-sector stores write the key record, so L2 never has to fetch it
                    if (OUTLINE) {
                        st_v4(kr, kk[0], kk[1], kk[2], kk[3]);
                        st_v4(kr + 16, kk[4], kk[5], kk[6], kk[7]);
                        st_v4(kr + 32, kk[8], kk[9], kk[10], 0u);
                        st_v4(kr + 48, 0u, 0u, 0u, 0u);
                    } else {
                        st_sector(kr, kk[0], kk[1], kk[2], kk[3], kk[4], kk[5], kk[6], kk[7]);
                        st_sector(kr + 32, kk[8], kk[9], kk[10], 0u, 0u, 0u, 0u, 0u);
                    }
                    st_release_u64(s, fp | SLOT_READY);
                    atomicAdd(&p.tstate->n_groups, 1ull);
                    return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
                }
            }
            // the fingerprint is in place from the claim on: a different one settles the mismatch without reading the key
            // (and without waiting for a slot that is still being published)
            if ((hd & 0xFFFFFFFF00000000ull) == fp) {
                // the key record is ordered behind the publisher's stores only through an acquire load of READY
                do {
                    hd = ld_acquire_u64(s);
                    if ((uint32_t)hd != SLOT_READY) __nanosleep(32);
                } while ((uint32_t)hd != SLOT_READY);
                const uint4 k0 = ld_relaxed_v4(kr), k1 = ld_relaxed_v4(kr + 16), k2 = ld_relaxed_v4(kr + 32);
                const bool same = k0.x == kk[0] && k0.y == kk[1] && k0.z == kk[2] && k0.w == kk[3] && k1.x == kk[4] && k1.y == kk[5] &&
                                  k1.z == kk[6] && k1.w == kk[7] && k2.x == kk[8] && k2.y == kk[9] && k2.z == kk[10];
                if (same) return slot_add(s + SlotLayout<KW>::VAL_OFF, bytes, packets, count);
            }
            slot = (slot + 1) & p.slot_mask;
        }
    }
    atomicAdd(&p.tstate->n_dropped, count);
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

// Bounded insert into the candidate table of FA_CFG_TOPK_ONLY: at most kCandProbes slots are looked at; a key that finds
// neither itself nor a free slot is simply not a candidate (yet: its next flow tries again, and the table is pruned after
// every submit).  KW == 4 layout (addresses).
constexpr uint32_t kCandProbes = 16;
__device__ __forceinline__ void candidate_add(const SubmitParams &p, const uint32_t *key, unsigned long long h, unsigned long long bytes,
                                              unsigned long long packets, unsigned long long count)
{
    const unsigned long long klo = (unsigned long long)key[0] | ((unsigned long long)key[1] << 32);
    const unsigned long long khi = (unsigned long long)key[2] | ((unsigned long long)key[3] << 32);
    if ((klo & khi) == ~0ull) return side_slot_add<4>(p, bytes, packets, count);
    uint32_t slot = (uint32_t)(h >> 32) & p.slot_mask;
#pragma unroll 1
    for (uint32_t probe = 0; probe < kCandProbes; probe++) {
        uint8_t *s = p.slots + (size_t)slot * SlotLayout<4>::BYTES;
        unsigned long long clo, chi;
        ld_relaxed_u128(s, clo, chi);
        if ((clo & chi) == ~0ull) {
            cas_u128(s, ~0ull, ~0ull, klo, khi, clo, chi);
            if ((clo & chi) == ~0ull) {
                atomicAdd(&p.tstate->n_groups, 1ull);
                clo = klo;
                chi = khi;
            }
        }
        if (clo == klo && chi == khi) return slot_add(s + SlotLayout<4>::VAL_OFF, bytes, packets, count);
        slot = (slot + 1) & p.slot_mask;
    }
}

// the same search continued from a first slot that has already been loaded (clo:chi = its key words)
__device__ __forceinline__ uint8_t *candidate_find_from(const SubmitParams &p, unsigned long long klo, unsigned long long khi, uint32_t slot,
                                                        unsigned long long clo, unsigned long long chi)
{
#pragma unroll 1
    for (uint32_t probe = 0; probe < kCandProbes; probe++) {
        uint8_t *s = p.slots + (size_t)slot * SlotLayout<4>::BYTES;
        if (probe) ld_relaxed_u128(s, clo, chi);
        if (clo == klo && chi == khi) return s;
        if ((clo & chi) == ~0ull) return nullptr;
        slot = (slot + 1) & p.slot_mask;
    }
    return nullptr;
}

// the key's slot in the candidate table, or nullptr (bounded probe; an empty slot ends the search)
__device__ __forceinline__ uint8_t *candidate_find(const SubmitParams &p, unsigned long long klo, unsigned long long khi, unsigned long long h)
{
    uint32_t slot = (uint32_t)(h >> 32) & p.slot_mask;
#pragma unroll 1
    for (uint32_t probe = 0; probe < kCandProbes; probe++) {
        uint8_t *s = p.slots + (size_t)slot * SlotLayout<4>::BYTES;
        unsigned long long clo, chi;
        ld_relaxed_u128(s, clo, chi);
        if (clo == klo && chi == khi) return s;
        if ((clo & chi) == ~0ull) return nullptr;
        slot = (slot + 1) & p.slot_mask;
    }
    return nullptr;
}

// A candidate's flow: its sums AND its sketch weight go to the slot (the 48-byte slot's spare word holds the weight not yet
// in the sketch); k_apply_pending moves the weight into the sketch after the launch.  One hot address per heavy key would
// serialise in L2 exactly like the hot groups of a roll-up, so the CTA's replica takes the update when it can.
__device__ __forceinline__ void slot_add_pending(uint8_t *s, unsigned long long bytes, unsigned long long packets, unsigned long long count,
                                                 unsigned long long weight)
{
    unsigned long long *v = reinterpret_cast<unsigned long long *>(s + SlotLayout<4>::VAL_OFF);
    red_add_u64(v + 0, bytes);
    red_add_u64(v + 1, packets);
    red_add_u64(v + 2, count);
    red_add_u64(v + 3, weight);
}

__device__ __forceinline__ bool hot_add_pending(uint8_t *replica, uint32_t hot_mask, unsigned long long klo, unsigned long long khi,
                                                unsigned long long h, unsigned long long bytes, unsigned long long packets, unsigned long long count,
                                                unsigned long long weight)
{
    uint32_t slot = (uint32_t)(h >> 20) & hot_mask;
#pragma unroll 1
    for (uint32_t probe = 0; probe < 4u; probe++) {
        uint8_t *s = replica + (size_t)slot * SlotLayout<4>::BYTES;
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
            slot_add_pending(s, bytes, packets, count, weight);
            return true;
        }
        slot = (slot + 1) & hot_mask;
    }
    return false;
}

// One decoded flow into the group table (+ sketch).  hot: this submit sends updates through the CTA's
// replica first (keys repeat a lot: one shared slot per key would serialise in L2).
// Returns the low hash bits of the key, or 0 with have=false when the flow has no key.
template <int MODE, bool OUTLINE = false>
__device__ __forceinline__ uint32_t aggregate_flow(const SubmitParams &p, const Flow &f, uint32_t &nokey, bool hot, bool &have,
                                                   unsigned long long &total_w, const unsigned long long admit_bar = 0, const uint32_t rec = 0)
{
    constexpr int KW = KeyTraits<MODE>::KW;
    uint32_t key[KW];
    have = make_key<MODE>(f, key);
    if (!have) {
        nokey++;
        return 0u;
    }
    const unsigned long long h = hash64<KW>(key);
    const unsigned long long hs = h;  // slot_hash<KW>(key)
    unsigned long long b = f.bytes, pk = f.packets;
    if (p.scale) {  // sum(Bytes*SamplingRate): viz-ch.json:74
        b *= f.sampling_rate;
        pk *= f.sampling_rate;
    }
    if (KW == 4 && p.cms && p.admit_shift) {
        // heavy hitters only (viz-ch.json:233's top-N): the sketch sees every flow, the candidate table only keys whose estimate
        // has reached the admission threshold -- memory is bounded by 2^admit_shift candidates, whatever the number of keys
        const unsigned long long w = f.bytes * f.sampling_rate;
        const unsigned long long klo = (unsigned long long)key[0] | ((unsigned long long)key[KW == 4 ? 1 : 0] << 32);
        const unsigned long long khi = (unsigned long long)key[KW == 4 ? 2 : 0] | ((unsigned long long)key[KW == 4 ? 3 : 0] << 32);
        // Two lookups decide a flow's path -- the CTA's replica (a key found there was a candidate when it got there, and
        // candidates only leave between launches) and the candidate table -- and their first probes are independent loads:
        // both are issued before either is looked at, so the common cases cost ONE L2 round trip, not two in a row.
        // (lanes of a warp holding the same key combining before they go to memory measured 1.5 % slower: profiles/r02/experiments)
        const unsigned long long cnt = 1ull, wsum = w;
        const bool keyed = p.slots && (klo & khi) != ~0ull;
        uint8_t *rep = p.hot_slots ? p.hot_slots + (size_t)(blockIdx.x & (kHotReplicas - 1u)) * ((size_t)p.hot_mask + 1u) * SlotLayout<4>::BYTES : nullptr;
        uint8_t *rs = nullptr, *cs = nullptr;
        unsigned long long rlo = 0ull, rhi = 0ull, clo = ~0ull, chi = ~0ull;
        const uint32_t cslot = (uint32_t)(h >> 32) & p.slot_mask;
        if (keyed) {
            if (rep) {
                rs = rep + (size_t)((uint32_t)(h >> 20) & p.hot_mask) * SlotLayout<4>::BYTES;
                ld_relaxed_u128(rs, rlo, rhi);
            }
            ld_relaxed_u128(p.slots + (size_t)cslot * SlotLayout<4>::BYTES, clo, chi);
        }
        const bool in_replica = rs && rlo == klo && rhi == khi;
        if (in_replica) {
            slot_add_pending(rs, b, pk, cnt, wsum);
        } else if (keyed && (cs = candidate_find_from(p, klo, khi, cslot, clo, chi)) != nullptr) {
            // a candidate (a heavy key) this replica has not met yet, or met behind a collision: slot and sketch weight through
            // the replica when it has room; the sketch is settled after the launch
            bool done = false;
            if (rep) done = hot_add_pending(rep, p.hot_mask, klo, khi, h, b, pk, cnt, wsum);
            if (!done) slot_add_pending(cs, b, pk, cnt, wsum);
        } else {
            // sample and hold (Estan & Varghese): the sketch takes the flow through fire-and-forget reductions, and the flow
            // makes its key a candidate with probability min(1, w / bar), bar = total weight / (64 K) -- at most 64 K
            // admissions are expected per launch whatever the number of keys, and a key that weighs several bars (every
            // top-K key does) is admitted with probability 1 - exp(-weight / bar).  The draw is a hash of (key, record
            // number, submit number): the same stream admits the same keys on every run.
            cms_add(p, h, wsum);
            uint32_t u = (uint32_t)(h >> 11) ^ (rec * 0x9E3779B9u) ^ p.sample_seed;
            u ^= u >> 16; u *= 0x7FEB352Du; u ^= u >> 15; u *= 0x846CA68Bu; u ^= u >> 16;
            const unsigned long long lo = (unsigned long long)u * admit_bar, hi = __umul64hi((unsigned long long)u, admit_bar);
            const bool admit = wsum >= admit_bar || hi < (wsum >> 32) || (hi == (wsum >> 32) && lo < (wsum << 32));   // u / 2^32 < w / bar
            if (p.slots && admit) candidate_add(p, key, h, b, pk, cnt);
        }
        total_w += w;
        return (uint32_t)h;
    }
    if (p.slots) {
        bool done = false;
        if (KW <= 4 && hot)
            done = hot_add<(KW <= 4 ? KW : 1)>(p.hot_slots + (size_t)(blockIdx.x & (kHotReplicas - 1u)) * kHotSlots * SlotLayout<(KW <= 4 ? KW : 1)>::BYTES, key, hs, b, pk);
        if (!done) table_add<KW, OUTLINE>(p, key, hs, b, pk, 1ull);
    }
    if (p.cms) cms_add(p, h, f.bytes * f.sampling_rate);  // viz-ch.json:233 weight
    return (uint32_t)(hs >> 32);
}

// ---- tile staging: one bulk-async copy per tile ----------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

struct TileInfo {
    uint32_t r0, n;   // first record, record count
    uint32_t b0;      // stream byte of the tile's first record
    uint32_t a0;      // b0 rounded down to 16: stream byte of shared-memory byte 0
    uint32_t s_end;   // stream byte one past the staged range (a0 if nothing is staged)
};

// Stage as much of the byte span of records [r0, r0+n) as fits into shared memory.
// Records that end beyond s_end (oversized tiles, out-of-order offsets) are parsed
// from global memory instead.
__device__ __forceinline__ TileInfo stage_tile(const SubmitParams &p, uint32_t tile, uint8_t *smem)
{
    TileInfo t;
    t.r0 = tile * p.tile_records;
    t.n = min(p.tile_records, p.n_records - t.r0);
    t.b0 = __ldg(p.offsets + t.r0);
    const uint32_t b1 = __ldg(p.offsets + t.r0 + t.n);
    t.a0 = t.b0 & ~15u;
    const unsigned long long end = p.base + p.len;
    const bool sane = t.b0 <= b1 && t.b0 >= p.base && (unsigned long long)b1 <= end;
    uint32_t nbytes = 0;
    if (sane) nbytes = min((b1 - t.a0 + 15u) & ~15u, p.tile_bytes);
    t.s_end = t.a0 + nbytes;
    const uint32_t bar = smem_u32(smem + p.tile_bytes + kTilePad);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (nbytes) {
            mbar_expect_tx(bar, nbytes);
            bulk_load(smem_u32(smem), p.buf + ((unsigned long long)t.a0 - p.base), nbytes, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    return t;
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
    Columns c;  // only read by ColConsumer
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
    static constexpr uint32_t NEED = KeyTraits<MODE>::NEED | F_BYTES | F_PACKETS | (WEIGHTED ? F_SAMPLING_RATE : 0u);
#ifndef FA_AGG_MIN_BLOCKS
#define FA_AGG_MIN_BLOCKS 8
#endif
#ifndef FA_AGG_MIN_BLOCKS_W4
#define FA_AGG_MIN_BLOCKS_W4 5
#endif
    // 32 registers per thread (8 CTAs per SM) for the roll-ups; 48 (5 CTAs) for weighted address keys -- sketch + candidate paths
    // are bound by L2 latency, not by issue slots, and measured 1.92 / 1.76 / 1.69 ms per slab at 32 / 40 / 48 registers -- and for
    // the 5-tuple
    static constexpr int MIN_BLOCKS = KW <= 2 ? FA_AGG_MIN_BLOCKS : (KW == 4 ? (WEIGHTED ? FA_AGG_MIN_BLOCKS_W4 : FA_AGG_MIN_BLOCKS) : 5);
    static constexpr bool HOT = KW <= 4;                 // 5-tuples are high-cardinality by nature
    static constexpr bool PERMUTE = true;                // nothing is stored per record: lanes may take any record
    struct Item {
        uint32_t h32;
        bool have;
        unsigned long long weight;  // sketched weight of the record (FA_CFG_TOPK_ONLY keeps the running total)
        unsigned long long bar;     // FA_CFG_TOPK_ONLY: estimate a key needs to become a candidate in this launch
    };
    // total weight seen before this launch >> admit_shift (nothing writes total_weight while the launch runs); loaded once,
    // early, so that its latency hides behind the tile copy
    static __device__ __forceinline__ unsigned long long admit_bar(const SubmitParams &p)
    {
        return (KW == 4 && WEIGHTED && p.admit_shift) ? (__ldg(&p.counters->total_weight) >> p.admit_shift) : 0ull;
    }
    static __device__ __forceinline__ void item_clear(Item &it)
    {
        it.h32 = 0;
        it.have = false;
        it.weight = 0;
        it.bar = 0;
    }
    static __device__ __forceinline__ void add_weight_one(const SubmitParams &p, const Item &it)  // a record parsed out of line
    {
        if (KW == 4 && WEIGHTED && p.admit_shift && it.weight) atomicAdd(&p.counters->total_weight_acc, it.weight);
    }
    static __device__ __forceinline__ void set_bar(Item &it, unsigned long long bar) { it.bar = bar; }
    // the tile's sketched weight into the context's running total (one atomic per warp)
    static __device__ __forceinline__ void add_weight(const SubmitParams &p, const Item &it)
    {
        if (KW != 4 || !WEIGHTED || !p.admit_shift) return;
        unsigned long long w = it.weight;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, d);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&p.counters->total_weight_acc, w);
    }
    static __device__ __forceinline__ bool want_hot(const SubmitParams &p)
    {
        if (!HOT || !p.hot_slots) return false;
        const unsigned int d = __ldg(&p.counters->hint[p.hint_set ^ 1u][0]), n = __ldg(&p.counters->hint[p.hint_set ^ 1u][1]);
        return n != 0u && d * 16u >= n;  // >= 1/16 of the sampled lanes repeat
    }
    template <bool OUTLINE = false>
    static __device__ __forceinline__ void consume(const TileParams &tp, uint32_t r, bool ok, Flow &f, uint32_t &bad, uint32_t &nokey, bool hot,
                                                   Item &it)
    {
        if (ok) {
            if (!WEIGHTED) f.sampling_rate = 1;  // unused unless scale/cms, which imply WEIGHTED
            it.h32 = aggregate_flow<MODE, OUTLINE>(tp.p, f, nokey, hot, it.have, it.weight, it.bar, r);
        } else {
            bad++;  // inserter.go:125-126: log, skip the row
        }
    }
    // every 64th tile measures how often keys repeat inside a warp (feeds the next submit's decision)
    static __device__ __forceinline__ void sample_repeats(const SubmitParams &p, const Item &it)
    {
        if (!HOT || (blockIdx.x & 63u) != 0u) return;
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
    static constexpr int MIN_BLOCKS = 4;  // all 16 fields live: 64 registers per thread
    static constexpr bool PERMUTE = false;  // column stores stay coalesced: lane i writes row r0+i
    struct Item {};
    static __device__ __forceinline__ void item_clear(Item &) {}
    static __device__ __forceinline__ bool want_hot(const SubmitParams &) { return false; }
    static __device__ __forceinline__ void sample_repeats(const SubmitParams &, const Item &) {}
    static __device__ __forceinline__ void add_weight(const SubmitParams &, const Item &) {}
    static __device__ __forceinline__ void add_weight_one(const SubmitParams &, const Item &) {}
    static __device__ __forceinline__ unsigned long long admit_bar(const SubmitParams &) { return 0ull; }
    static __device__ __forceinline__ void set_bar(Item &, unsigned long long) {}
    template <bool OUTLINE = false>
    static __device__ __forceinline__ void consume(const TileParams &tp, uint32_t r, bool ok, Flow &f, uint32_t &bad, uint32_t &nokey, bool, Item &)
    {
        consume(tp, r, ok, f, bad, nokey);
    }
    static __device__ __forceinline__ void consume(const TileParams &tp, uint32_t r, bool ok, Flow &f, uint32_t &bad, uint32_t &)
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

// one record straight from global memory (it did not fit the staged tile, or its
// offsets are out of order): correct, slow, rare.  Out of line with its own Flow so
// the hot path keeps its Flow in registers.
// Returns bad | nokey << 1.
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
    Consumer::item_clear(it);
    Consumer::set_bar(it, Consumer::admit_bar(tp.p));
    Consumer::template consume<true>(tp, r, ok, f, bad, nokey, false, it);
    Consumer::add_weight_one(tp.p, it);
    return bad | (nokey << 1);
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

// Which record of the tile does this thread parse?  Lanes of a warp read their records from shared memory in
// lock step, so the bank pattern is set by the byte distance between the records of neighbouring lanes.  With
// consecutive records and near-constant record sizes that distance can resonate with the 32 x 4-byte banks
// (measured: 86-byte records = 21.5 words, 3 lanes apart = 64.5 words -> 3-4-way conflicts on every load, fused
// kernel 0.87 ms instead of 0.56 ms).  Each group of d = 2^shift warps therefore shares a run of 32*d records,
// neighbouring lanes taking records d apart; d in {1,2,4,8} is chosen per batch from its mean record size so that the
// predicted bank multiplicity is smallest (pick_lane_stride, host side; the host passes log2 d, so no division
// here).  A bijection on [0, THREADS).
__device__ __forceinline__ uint32_t record_of_thread(uint32_t shift)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    return ((warp >> shift) << (5u + shift)) + (lane << shift) + (warp & ((1u << shift) - 1u));
}

// ---- the tile kernel: decode (+ consume) one tile per CTA ------------------------------------

template <class Consumer, int THREADS>
__global__ void __launch_bounds__(THREADS, (Consumer::MIN_BLOCKS * kThreads) / THREADS) k_tile(const __grid_constant__ TileParams tp)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const SubmitParams &p = tp.p;
    const TileInfo t = stage_tile(p, blockIdx.x, smem);
    const bool hot = Consumer::want_hot(p);  // uniform over the grid; the load overlaps the tile copy
    const unsigned long long bar = Consumer::admit_bar(p);
    uint32_t bad = 0, nokey = 0;
    uint32_t in_tile = threadIdx.x;
    // full tiles only (the mapping is a bijection on [0, THREADS)); short tails keep the identity
    if (Consumer::PERMUTE && p.lane_shift && t.n == (uint32_t)THREADS && (THREADS / 32) >= (1 << p.lane_shift)) in_tile = record_of_thread(p.lane_shift);
    const bool active = in_tile < t.n;
    const uint32_t r = t.r0 + in_tile;
    uint32_t o0 = 0, o1 = 0;
    if (active) {  // in flight while the bulk copy lands
        o0 = __ldg(p.offsets + r);
        o1 = __ldg(p.offsets + r + 1);
    }
    mbar_wait(smem_u32(smem + p.tile_bytes + kTilePad), 0);
    typename Consumer::Item item;
    Consumer::item_clear(item);
    Consumer::set_bar(item, bar);
    if (active) {
        if (o0 >= t.b0 && o0 <= o1 && o1 <= t.s_end) {
            Flow f;
            flow_reset(f);
            SmemSrc s;
            s.base = smem_u32(smem);
            const bool ok = decode_record<Consumer::NEED>(s, o0 - t.a0, o1 - t.a0, p.framed != 0, f);
            Consumer::consume(tp, r, ok, f, bad, nokey, hot, item);
        } else {
            const uint32_t res = record_from_global<Consumer>(tp, r, o0, o1);  // updates the table itself
            bad += res & 1u;
            nokey += res >> 1;
        }
    }
    Consumer::sample_repeats(p, item);
    Consumer::add_weight(p, item);
    flush_counts(p, bad, nokey);
}

// ---- kernel 2 alone: columns -> table / sketch -------------------------------------------------

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_aggregate_columns(const SubmitParams p, const Columns c)
{
    uint32_t nokey = 0;
    unsigned long long wsum = 0;
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
        aggregate_flow<MODE>(p, f, nokey, false, have, wsum, p.admit_shift ? (__ldg(&p.counters->total_weight) >> p.admit_shift) : 0ull, r);
    }
    if (wsum) atomicAdd(&p.counters->total_weight_acc, wsum);
    flush_counts(p, 0, nokey);
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
        table_add<KW>(p, key, slot_hash<KW>(key), v[0], v[1], v[2]);
#pragma unroll
        for (int k = 0; k < KEY64; k++) w[k] = ~0ull;
        unsigned long long *vv = reinterpret_cast<unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
        vv[0] = vv[1] = vv[2] = 0ull;
    }
}

// Slot i of a table of n_slots (the last one is the reserved side slot of the all-ones key; it is occupied iff
// side_claimed -- counters->side_state, which is what n_groups counted: its count alone would miss a key whose merged
// count is 0).  Returns whether the slot holds a group; `vals` = its three sums.
template <int KW>
__device__ __forceinline__ bool slot_read(const uint8_t *slots, unsigned long long n_slots, unsigned long long i, bool side_claimed, uint32_t *key,
                                          const unsigned long long *&vals)
{
    const uint8_t *s = slots + i * SlotLayout<KW>::BYTES;
    const bool is_side = i == n_slots - 1;
    vals = reinterpret_cast<const unsigned long long *>(s + SlotLayout<KW>::VAL_OFF);
    if (KW <= 2) {
        const unsigned long long k = *reinterpret_cast<const unsigned long long *>(s);
        key[0] = (uint32_t)k;
        if (KW == 2) key[KW == 2 ? 1 : 0] = (uint32_t)(k >> 32);
        return is_side ? side_claimed : k != ~0ull;
    } else if (KW == 4) {
        const unsigned long long lo = reinterpret_cast<const unsigned long long *>(s)[0], hi = reinterpret_cast<const unsigned long long *>(s)[1];
        key[0] = (uint32_t)lo; key[KW == 4 ? 1 : 0] = (uint32_t)(lo >> 32); key[KW == 4 ? 2 : 0] = (uint32_t)hi; key[KW == 4 ? 3 : 0] = (uint32_t)(hi >> 32);
        return is_side ? side_claimed : (lo & hi) != ~0ull;
    } else {
        const unsigned long long hd = *reinterpret_cast<const unsigned long long *>(s);
        if (is_side || (uint32_t)hd != SLOT_READY) return false;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(slots + wide_key_offset(n_slots) + i * kWideKeyBytes);
#pragma unroll
        for (int k = 0; k < KW; k++) key[k] = w[k];
        return true;
    }
}

template <int KW>
__global__ void __launch_bounds__(256) k_compact_rows(const uint8_t *slots, unsigned long long n_slots, fa_row *rows,
                                                      unsigned long long cap, TableState *counters)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t key[KW];
        const unsigned long long *v;
        if (!slot_read<KW>(slots, n_slots, i, counters->side_state != 0u, key, v)) continue;
        if (i == n_slots - 1) {  // the side slot holds the all-ones key
#pragma unroll
            for (int k = 0; k < KW; k++) key[k] = 0xFFFFFFFFu;
        }
        const unsigned long long at = atomicAdd(&counters->flush_rows, 1ull);
        if (at >= cap) continue;
        fa_row r;
#pragma unroll
        for (int k = 0; k < FA_MAX_KEY_WORDS; k++) r.key[k] = k < KW ? key[k] : 0u;
        r.bytes = v[0];
        r.packets = v[1];
        r.count = v[2];
        rows[at] = r;
    }
}

// ---- FA_CFG_TOPK_ONLY: keep the candidate table small --------------------------------------------------------------
//
// After every submit: publish the running total of the sketched weights (the next launch's admission threshold scales
// with it), then rebuild the candidate table from the candidates whose sketch estimate is still at or above the
// threshold -- keys admitted early, when little had been seen and the bar was low, leave again.  Three launches over
// <= table_capacity slots: collect survivors, empty the table, re-insert.
// fold the replicas' candidate updates (sums + pending sketch weight) into the main candidate table, empty the replicas
__global__ void __launch_bounds__(256) k_merge_hot_candidates(const SubmitParams p, unsigned long long n_hot_slots)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_hot_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint8_t *s = p.hot_slots + (size_t)i * SlotLayout<4>::BYTES;
        unsigned long long *w = reinterpret_cast<unsigned long long *>(s);
        const unsigned long long klo = w[0], khi = w[1];
        if ((klo & khi) == ~0ull) continue;
        unsigned long long *v = reinterpret_cast<unsigned long long *>(s + SlotLayout<4>::VAL_OFF);
        uint32_t key[4] = {(uint32_t)klo, (uint32_t)(klo >> 32), (uint32_t)khi, (uint32_t)(khi >> 32)};
        const unsigned long long h = hash64<4>(key);
        uint8_t *cs = candidate_find(p, klo, khi, h);
        if (cs) {
            slot_add_pending(cs, v[0], v[1], v[2], v[3]);
        } else {  // left the table meanwhile (cannot happen within one launch, but never lose sketch weight)
            cms_add(p, h, v[3]);
            candidate_add(p, key, h, v[0], v[1], v[2]);
        }
        w[0] = w[1] = ~0ull;
        v[0] = v[1] = v[2] = v[3] = 0ull;
    }
}

// the candidates' weight that has not reached the sketch yet: add it now (the sketch is linear: when a weight lands does not matter)
__global__ void __launch_bounds__(256) k_apply_pending(const SubmitParams p, unsigned long long n_slots)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint8_t *s = p.slots + i * SlotLayout<4>::BYTES;
        const unsigned long long *k = reinterpret_cast<const unsigned long long *>(s);
        if ((k[0] & k[1]) == ~0ull) continue;
        unsigned long long *pend = reinterpret_cast<unsigned long long *>(s + SlotLayout<4>::VAL_OFF) + 3;
        const unsigned long long w = *pend;
        if (!w) continue;
        *pend = 0ull;
        uint32_t key[4] = {(uint32_t)k[0], (uint32_t)(k[0] >> 32), (uint32_t)k[1], (uint32_t)(k[1] >> 32)};
        cms_add(p, hash64<4>(key), w);
    }
}

__global__ void k_publish_weight(Counters *counters, TableState *ts)
{
    counters->total_weight = counters->total_weight_acc;
    ts->flush_rows = 0;
}

__global__ void __launch_bounds__(256) k_prune_collect(const uint8_t *slots, unsigned long long n_slots, const unsigned long long *cms, uint32_t depth,
                                                       uint32_t wlog2, uint32_t admit_shift, const Counters *counters, fa_row *rows, TableState *ts)
{
    const unsigned long long bar = counters->total_weight >> admit_shift;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t key[4];
        const unsigned long long *v;
        if (!slot_read<4>(slots, n_slots, i, ts->side_state != 0u, key, v)) continue;
        if (i == n_slots - 1) key[0] = key[1] = key[2] = key[3] = 0xFFFFFFFFu;
        const unsigned long long h = hash64<4>(key);
        const uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32) | 1u, mask = (1u << wlog2) - 1u;
        unsigned long long est = ~0ull;
        for (uint32_t j = 0; j < depth; j++) {
            const unsigned long long cnt = cms[((size_t)j << wlog2) + ((a + j * b) & mask)];
            est = cnt < est ? cnt : est;
        }
        if (est < bar) continue;
        const unsigned long long at = atomicAdd(&ts->flush_rows, 1ull);
        fa_row r;
#pragma unroll
        for (int k = 0; k < FA_MAX_KEY_WORDS; k++) r.key[k] = k < 4 ? key[k] : 0u;
        r.bytes = v[0];
        r.packets = v[1];
        r.count = v[2];
        rows[at] = r;
    }
}

// re-insert the ts->flush_rows survivors (the count lives on the device: no host round trip between the launches)
__global__ void __launch_bounds__(256) k_prune_reinsert(const SubmitParams p, const fa_row *rows)
{
    const unsigned long long n = p.tstate->flush_rows;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t key[4];
#pragma unroll
        for (int k = 0; k < 4; k++) key[k] = rows[i].key[k];
        table_add<4>(p, key, hash64<4>(key), rows[i].bytes, rows[i].packets, rows[i].count);
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
        table_add<KW>(p, key, slot_hash<KW>(key), rows[i].bytes, rows[i].packets, rows[i].count);
    }
}

// sketch estimate of every group (top-K candidates)
template <int KW>
__global__ void __launch_bounds__(256) k_estimate(const uint8_t *slots, unsigned long long n_slots, const unsigned long long *cms,
                                                  uint32_t depth, uint32_t wlog2, fa_hh *out, unsigned long long cap,
                                                  TableState *counters)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t key[KW];
        const unsigned long long *v;
        if (!slot_read<KW>(slots, n_slots, i, counters->side_state != 0u, key, v)) continue;
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
