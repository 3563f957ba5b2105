// flowagg.cu -- implementation of the C ABI in include/flowagg.h.
//
// Host side of the B200-native replacement for (*state).buffer / (*state).flush
// (inserter/inserter.go:90-165) and the flows_5m roll-up
// (compose/clickhouse/create.sh:92-110).  All compute is in the sm_100a kernels
// of kernels.cuh; there is no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#define FA_STR2(x) #x
#define FA_STR(x) FA_STR2(x)
#define FA_STR_CUDART FA_STR(CUDART_VERSION)

#include "../../include/flowagg.h"
#include "kernels.cuh"
#include "mocker_gen.h"

using namespace fa;

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------

struct fa_ctx {
    fa_config cfg;
    int kw = 0;
    uint32_t slot_bytes = 0;
    uint64_t capacity = 0;
    bool weighted = false;
    uint32_t admit_shift = 0;  // FA_CFG_TOPK_ONLY: log2(64 * topk_k), 0 otherwise
    // offsets-free host submits run one call behind: the batch staged by the latest fa_submit(offsets = NULL) is indexed
    // and launched by the NEXT call (or by whatever reads the context), so that the one host look the index needs waits
    // while the next batch's host-to-device copy is already running
    bool frame_pending = false;
    int frame_stage = 0;
    size_t frame_len = 0;
    uint32_t frame_flags = 0;
    uint32_t hot_slots_per_replica = kHotSlots;

    cudaStream_t stream = nullptr;  // compute
    cudaStream_t copy_stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;

    uint8_t *d_slots = nullptr;
    uint8_t *d_hot = nullptr;  // hot-key replicas (key modes with KW <= 4)
    bool hot_dirty = false;    // replicas may hold sums not yet folded into d_slots
    unsigned long long *d_cms = nullptr, *d_cms_global = nullptr;
    size_t cms_words = 0;
    Counters *d_counters = nullptr;
    Counters *h_counters = nullptr;  // pinned
    TableState *d_ts = nullptr;      // state of the table behind d_slots
    TableState *h_ts = nullptr;      // pinned: last snapshot of the current table's state
    TableState *h_ts_drain = nullptr; // pinned: state of the table an asynchronous flush is draining

    // asynchronous flush (fa_flush_begin / fa_flush_end): the filled table is swapped for a spare, empty one and drained
    // on a side stream while the next submits already fill the spare
    uint8_t *d_slots_spare = nullptr;
    TableState *d_ts_spare = nullptr;
    cudaStream_t flush_stream = nullptr;
    cudaEvent_t ev_swap = nullptr, ev_flush_done = nullptr, ev_spare_ready = nullptr;
    void *d_fscratch = nullptr;  // the drain's own scratch (d_scratch belongs to the submit / query paths)
    size_t fscratch_bytes = 0;
    bool flush_pending = false, flush_deferred = false;
    uint32_t flush_flags = 0, flush_m = 0;          // flags of the pending flush; rows the drain sorted and copied
    fa_row *flush_rows_in = nullptr;                // every compacted row of the drained table (device, in d_fscratch)

    // host-submit staging (double buffered)
    uint8_t *d_stage[2] = {nullptr, nullptr};
    uint32_t *d_stage_off[2] = {nullptr, nullptr};
    cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    int next_stage = 0;
    uint8_t *h_slab[2] = {nullptr, nullptr};
    uint32_t *h_slab_off[2] = {nullptr, nullptr};
    cudaEvent_t ev_slab[2] = {nullptr, nullptr};
    uint32_t *d_frame_off = nullptr;  // offsets found on the GPU (offsets == NULL submits)
    size_t frame_off_cap = 0;

    // kernel-1 columns of the last submit
    Columns cols{};
    uint64_t cols_n = 0;
    void *cols_block = nullptr;

    // flush / top-K scratch
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *h_bounce = nullptr;  // pinned landing area for flushed rows
    size_t bounce_bytes = 0;

    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;

    // device time of the decode/aggregate kernels: an event pair around every launch, read back lazily
    static constexpr int kBusyRing = 32;
    cudaEvent_t ev_busy[kBusyRing][2] = {};
    uint64_t busy_head = 0, busy_tail = 0;  // pairs recorded / pairs added to busy_us
    double busy_us = 0;

    uint64_t n_records = 0, n_submits = 0, bytes_in = 0;
    uint64_t last_groups = 0;  // rows of the previous flush: sizes the speculative (single-sync) flush
    uint64_t n_kernels = 0;  // launches of this library's own kernels (cub's are not counted)
    std::string last_error;
};

#define FA_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char b__[512];                                                                         \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            (ctx)->last_error = b__;                                                               \
            return FA_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

static const int k_key_words[FA_KEY_MODES] = {4, 2, 4, 4, 11, 1, 1};

static uint32_t slot_bytes_for(int kw) { return kw <= 2 ? 32u : (kw == 4 ? 48u : 32u); }  // wide keys: the head; the key records sit behind the heads (table_alloc_bytes)

extern "C" const char *fa_strerror(int s)
{
    switch (s) {
    case FA_OK: return "ok";
    case FA_ERR_INVALID: return "invalid argument";
    case FA_ERR_CUDA: return "CUDA error";
    case FA_ERR_NOMEM: return "out of memory";
    case FA_ERR_CAPACITY: return "output capacity too small";
    case FA_ERR_TABLE_FULL: return "group table full";
    case FA_ERR_NCCL: return "NCCL error";
    case FA_ERR_FRAMING: return "stream ends inside a record";
    default: return "unknown status";
    }
}

extern "C" const char *fa_last_error(const fa_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

extern "C" const char *fa_build_info(void) { return "libflowagg abi=" FA_STR(FA_ABI_VERSION) " arch=sm_100a cuda=" FA_STR_CUDART; }

static int ensure_scratch(fa_ctx *c, size_t bytes)
{
    if (bytes <= c->scratch_bytes) return FA_OK;
    if (c->d_scratch) cudaFree(c->d_scratch);
    c->d_scratch = nullptr;
    c->scratch_bytes = 0;
    FA_CUDA(c, cudaMalloc(&c->d_scratch, bytes));
    c->scratch_bytes = bytes;
    return FA_OK;
}

// empty `slots` (capacity + 1 of them) on `stream`; with_hot: the hot-key replicas too
template <int KW>
static cudaError_t launch_table_init(fa_ctx *c, uint8_t *slots, cudaStream_t stream, bool with_hot)
{
    const unsigned long long n_slots = c->capacity + 1;
    const unsigned long long words = n_slots * (SlotLayout<KW>::BYTES / 8);
    const int grid = (int)std::min<unsigned long long>((words + 255) / 256, (unsigned long long)c->num_sms * 32);
    k_table_init<KW><<<grid, 256, 0, stream>>>(slots, n_slots);
    c->n_kernels++;
    if (with_hot && c->d_hot) {
        k_table_init<KW><<<c->num_sms, 256, 0, stream>>>(c->d_hot, (unsigned long long)kHotReplicas * c->hot_slots_per_replica);
        c->n_kernels++;
        c->hot_dirty = false;
    }
    return cudaGetLastError();
}

static void fill_table_params(fa_ctx *c, SubmitParams &p)
{
    p.slots = c->d_slots;
    p.slot_mask = (uint32_t)(c->capacity - 1);
    p.counters = c->d_counters;
    p.tstate = c->d_ts;
    p.hot_slots = c->d_hot;
    p.hot_mask = c->hot_slots_per_replica - 1u;
}

template <int KW>
static cudaError_t launch_merge_hot(fa_ctx *c)
{
    SubmitParams p{};
    fill_table_params(c, p);
    k_merge_hot<KW><<<c->num_sms, 256, 0, c->stream>>>(p, kHotReplicas * kHotSlots);
    c->n_kernels++;
    return cudaGetLastError();
}

// fold the hot-key replicas into the main table (before anything reads it)
static int merge_hot(fa_ctx *c)
{
    if (!c->d_hot || !c->hot_dirty) return FA_OK;
    cudaError_t e;
    switch (c->kw) {
    case 1: e = launch_merge_hot<1>(c); break;
    case 2: e = launch_merge_hot<2>(c); break;
    default: e = launch_merge_hot<4>(c); break;
    }
    FA_CUDA(c, e);
    c->hot_dirty = false;
    return FA_OK;
}

// empty table: all-ones keys (CAS layouts) / state 0 (wide keys), zero sums
static int table_init_at(fa_ctx *c, uint8_t *slots, cudaStream_t stream, bool with_hot)
{
    if (!slots) return FA_OK;
    cudaError_t e;
    switch (c->kw) {
    case 1: e = launch_table_init<1>(c, slots, stream, with_hot); break;
    case 2: e = launch_table_init<2>(c, slots, stream, with_hot); break;
    case 4: e = launch_table_init<4>(c, slots, stream, with_hot); break;
    default: e = launch_table_init<11>(c, slots, stream, with_hot); break;
    }
    FA_CUDA(c, e);
    return FA_OK;
}
static int table_init(fa_ctx *c) { return table_init_at(c, c->d_slots, c->stream, true); }

extern "C" void fa_destroy(fa_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_slots);
    cudaFree(c->d_hot);
    cudaFree(c->d_cms);
    cudaFree(c->d_cms_global);
    cudaFree(c->d_counters);
    cudaFreeHost(c->h_counters);
    if (c->flush_stream) cudaStreamSynchronize(c->flush_stream);
    cudaFree(c->d_ts);
    cudaFree(c->d_ts_spare);
    cudaFreeHost(c->h_ts);
    cudaFree(c->d_slots_spare);
    cudaFree(c->d_fscratch);
    if (c->ev_swap) cudaEventDestroy(c->ev_swap);
    if (c->ev_flush_done) cudaEventDestroy(c->ev_flush_done);
    if (c->ev_spare_ready) cudaEventDestroy(c->ev_spare_ready);
    if (c->flush_stream) cudaStreamDestroy(c->flush_stream);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_stage[i]);
        cudaFree(c->d_stage_off[i]);
        cudaFreeHost(c->h_slab[i]);
        cudaFreeHost(c->h_slab_off[i]);
        if (c->ev_staged[i]) cudaEventDestroy(c->ev_staged[i]);
        if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
        if (c->ev_slab[i]) cudaEventDestroy(c->ev_slab[i]);
    }
    cudaFree(c->d_frame_off);
    cudaFree(c->cols_block);
    cudaFree(c->d_scratch);
    cudaFreeHost(c->h_bounce);
    for (auto &pair : c->ev_busy)
        for (cudaEvent_t e : pair)
            if (e) cudaEventDestroy(e);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int alloc_columns(fa_ctx *c)
{
    const size_t n = c->cfg.max_batch_records;
    // one block: 5 u64, 8 u32, 3 x 16 B, 4 u8 columns
    const size_t bytes = n * (5 * 8 + 8 * 4 + 3 * 16 + 4) + 20 * 256 + 1024;
    FA_CUDA(c, cudaMalloc(&c->cols_block, bytes));
    uint8_t *p = (uint8_t *)c->cols_block;
    auto take = [&](size_t b) { uint8_t *r = p; p += (b + 255) & ~(size_t)255; return r; };
    Columns &k = c->cols;
    k.src_addr = (uint4 *)take(n * 16);
    k.dst_addr = (uint4 *)take(n * 16);
    k.sampler_addr = (uint4 *)take(n * 16);
    k.time_received = (unsigned long long *)take(n * 8);
    k.time_flow_start = (unsigned long long *)take(n * 8);
    k.sampling_rate = (unsigned long long *)take(n * 8);
    k.bytes = (unsigned long long *)take(n * 8);
    k.packets = (unsigned long long *)take(n * 8);
    k.type = (uint32_t *)take(n * 4);
    k.sequence_num = (uint32_t *)take(n * 4);
    k.src_as = (uint32_t *)take(n * 4);
    k.dst_as = (uint32_t *)take(n * 4);
    k.etype = (uint32_t *)take(n * 4);
    k.proto = (uint32_t *)take(n * 4);
    k.src_port = (uint32_t *)take(n * 4);
    k.dst_port = (uint32_t *)take(n * 4);
    k.valid = take(n);
    k.src_addr_len = take(n);
    k.dst_addr_len = take(n);
    k.sampler_addr_len = take(n);
    return FA_OK;
}

extern "C" int fa_create(const fa_config *cfg, fa_ctx **out)
{
    if (!cfg || !out || cfg->abi_version != FA_ABI_VERSION) return FA_ERR_INVALID;
    if (cfg->key_mode >= FA_KEY_MODES) return FA_ERR_INVALID;
    if ((cfg->flags & FA_CFG_NO_AGGREGATE) && !(cfg->flags & FA_CFG_COLUMNS)) return FA_ERR_INVALID;
    fa_ctx *c = new (std::nothrow) fa_ctx();
    if (!c) return FA_ERR_NOMEM;
    *out = c;  // returned even on failure so the caller can read fa_last_error, then fa_destroy
    c->cfg = *cfg;
    if (!c->cfg.table_capacity) c->cfg.table_capacity = 1ull << 17;
    if (!c->cfg.cms_depth) c->cfg.cms_depth = 4;
    if (!c->cfg.cms_width_log2) c->cfg.cms_width_log2 = 20;
    if (!c->cfg.max_batch_bytes) c->cfg.max_batch_bytes = 256ull << 20;
    if (!c->cfg.max_batch_records) c->cfg.max_batch_records = 4u << 20;
    if (c->cfg.cms_depth > 16 || c->cfg.cms_width_log2 > 30 || c->cfg.cms_width_log2 < 4) return FA_ERR_INVALID;
    if (c->cfg.max_batch_bytes > 0xFFFFFFF0ull) return FA_ERR_INVALID;  // offsets are u32
    uint64_t cap = 16;
    while (cap < c->cfg.table_capacity) cap <<= 1;
    if (cap > (1ull << 32)) return FA_ERR_INVALID;
    c->capacity = cap;
    c->kw = k_key_words[c->cfg.key_mode];
    c->slot_bytes = slot_bytes_for(c->kw);
    c->weighted = (c->cfg.flags & (FA_CFG_CMS | FA_CFG_SCALE_SAMPLING)) != 0;
    if (c->cfg.flags & FA_CFG_TOPK_ONLY) {
        // heavy hitters only: a sketch plus a bounded candidate table, address keys (KW == 4)
        if (!(c->cfg.flags & FA_CFG_CMS) || c->kw != 4 || (c->cfg.flags & (FA_CFG_NO_AGGREGATE | FA_CFG_SCALE_SAMPLING))) return FA_ERR_INVALID;
        const uint64_t k = c->cfg.topk_k ? c->cfg.topk_k : 1000;
        uint32_t shift = 6;  // 64 x K candidates can truly weigh more than total / (64 K)
        while ((1ull << (shift - 6)) < k) shift++;
        c->admit_shift = shift;
        const uint64_t want = 4ull << shift;
        if (cfg->table_capacity == 0) {
            c->capacity = want;
        } else if (c->capacity < want) {
            c->last_error = "FA_CFG_TOPK_ONLY needs table_capacity >= 256 * topk_k";
            return FA_ERR_INVALID;
        }
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        c->last_error = "no CUDA device: libflowagg has no CPU fallback";
        return FA_ERR_CUDA;
    }
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    cudaDeviceProp prop;
    FA_CUDA(c, cudaGetDeviceProperties(&prop, c->cfg.device));
    if (prop.major != 10) {
        char b[160];
        snprintf(b, sizeof b, "device %d is sm_%d%d; libflowagg carries sm_100a code only", c->cfg.device, prop.major, prop.minor);
        c->last_error = b;
        return FA_ERR_CUDA;
    }
    c->num_sms = prop.multiProcessorCount;
    if (c->cfg.stream || (c->cfg.flags & FA_CFG_CALLER_STREAM)) {
        c->stream = (cudaStream_t)c->cfg.stream;  // NULL here = the legacy default stream
    } else {
        FA_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    FA_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    FA_CUDA(c, cudaEventCreate(&c->ev_t0));
    FA_CUDA(c, cudaEventCreate(&c->ev_t1));
    for (int i = 0; i < 2; i++) {
        FA_CUDA(c, cudaEventCreateWithFlags(&c->ev_staged[i], cudaEventDisableTiming));
        FA_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
        FA_CUDA(c, cudaEventCreateWithFlags(&c->ev_slab[i], cudaEventDisableTiming));
    }
    FA_CUDA(c, cudaMalloc(&c->d_counters, sizeof(Counters)));
    FA_CUDA(c, cudaMemsetAsync(c->d_counters, 0, sizeof(Counters), c->stream));
    // nothing is known about the keys yet: the first submit combines per tile (at worst ~25 % slower than
    // it had to be; not combining hot keys would be several times slower)
    FA_CUDA(c, cudaMemsetAsync(&c->d_counters->hint[1][0], 1, 8, c->stream));
    FA_CUDA(c, cudaHostAlloc(&c->h_counters, sizeof(Counters), cudaHostAllocDefault));
    FA_CUDA(c, cudaMalloc(&c->d_ts, sizeof(TableState)));
    FA_CUDA(c, cudaMemsetAsync(c->d_ts, 0, sizeof(TableState), c->stream));
    FA_CUDA(c, cudaHostAlloc(&c->h_ts, 2 * sizeof(TableState), cudaHostAllocDefault));
    memset(c->h_ts, 0, 2 * sizeof(TableState));
    c->h_ts_drain = c->h_ts + 1;
    if (!(c->cfg.flags & FA_CFG_NO_AGGREGATE)) {
        FA_CUDA(c, cudaMalloc(&c->d_slots, table_alloc_bytes(c->kw, c->capacity + 1)));  // + the side slot
        c->hot_slots_per_replica = c->admit_shift ? kCandHotSlots : kHotSlots;
        if (c->kw <= 4) FA_CUDA(c, cudaMalloc(&c->d_hot, (size_t)kHotReplicas * c->hot_slots_per_replica * c->slot_bytes));
        int rc = table_init(c);
        if (rc) return rc;
    }
    if (c->cfg.flags & FA_CFG_CMS) {
        c->cms_words = (size_t)c->cfg.cms_depth << c->cfg.cms_width_log2;
        FA_CUDA(c, cudaMalloc(&c->d_cms, c->cms_words * 8));
        FA_CUDA(c, cudaMemsetAsync(c->d_cms, 0, c->cms_words * 8, c->stream));
    }
    if (c->cfg.flags & FA_CFG_COLUMNS) {
        int rc = alloc_columns(c);
        if (rc) return rc;
    }
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    return FA_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel launch
// ---------------------------------------------------------------------------------------------

template <class Consumer, int THREADS>
static cudaError_t launch_tile_t(fa_ctx *c, const TileParams &tp, uint32_t n_tiles)
{
    const size_t smem = (size_t)tp.p.tile_bytes + kTilePad + 16;  // tile buffer + over-read pad + mbarrier
    static thread_local unsigned long long configured = 0;  // bit d: done for device d (per kernel instantiation, per thread)
    const unsigned long long dev_bit = 1ull << (c->cfg.device & 63);
    if (smem > 48 * 1024 && !(configured & dev_bit)) {  // the attribute is per device
        cudaError_t e = cudaFuncSetAttribute(k_tile<Consumer, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTileBytesMax + kTilePad + 16));
        if (e != cudaSuccess) return e;
        configured |= dev_bit;
    }
    k_tile<Consumer, THREADS><<<n_tiles, THREADS, smem, c->stream>>>(tp);
    c->n_kernels++;
    return cudaGetLastError();
}

template <class Consumer>
static cudaError_t launch_tile(fa_ctx *c, const TileParams &tp, uint32_t n_tiles)
{
    if (tp.p.tile_records <= 128) return launch_tile_t<Consumer, 128>(c, tp, n_tiles);
    if (tp.p.tile_records <= 256) return launch_tile_t<Consumer, 256>(c, tp, n_tiles);
    if (tp.p.tile_records <= 512) return launch_tile_t<Consumer, 512>(c, tp, n_tiles);
    return launch_tile_t<Consumer, 1024>(c, tp, n_tiles);
}

template <int MODE>
static cudaError_t launch_fused_mode(fa_ctx *c, const TileParams &tp, uint32_t n_tiles)
{
    return c->weighted ? launch_tile<AggConsumer<MODE, true>>(c, tp, n_tiles) : launch_tile<AggConsumer<MODE, false>>(c, tp, n_tiles);
}

template <int MODE>
static cudaError_t launch_agg_columns_mode(fa_ctx *c, const SubmitParams &p, int grid)
{
    k_aggregate_columns<MODE><<<grid, kThreads, 0, c->stream>>>(p, c->cols);
    c->n_kernels++;
    return cudaGetLastError();
}

#define FA_DISPATCH_MODE(mode, CALL)                       \
    switch (mode) {                                        \
    case FA_KEY_FLOWS5M: e = CALL(FA_KEY_FLOWS5M); break;  \
    case FA_KEY_ASPAIR: e = CALL(FA_KEY_ASPAIR); break;    \
    case FA_KEY_SRCADDR: e = CALL(FA_KEY_SRCADDR); break;  \
    case FA_KEY_DSTADDR: e = CALL(FA_KEY_DSTADDR); break;  \
    case FA_KEY_5TUPLE: e = CALL(FA_KEY_5TUPLE); break;    \
    case FA_KEY_SRCPORT: e = CALL(FA_KEY_SRCPORT); break;  \
    default: e = CALL(FA_KEY_DSTPORT); break;              \
    }

// Shared-memory bank conflicts of the record cursors: lane j of a warp starts at byte ~ j*d*avg of the tile, so
// near-constant record sizes with avg/4 close to a multiple of 32/k words put every k-th lane on one bank (mocker
// records of 86 B: 11 lanes per bank at d = 1).  Choose the record stride d between neighbouring lanes
// that minimises the worst start-bank multiplicity; ties keep the smaller d (coalesced offset loads).
static uint32_t pick_lane_stride(double avg)
{
    static const int forced = getenv("FA_LANE_STRIDE") ? atoi(getenv("FA_LANE_STRIDE")) : 0;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return (uint32_t)forced;
    uint32_t best_d = 1, best_conf = 33;
    for (uint32_t d = 1; d <= 8; d *= 2) {
        uint32_t cnt[32] = {0}, conf = 0;
        for (uint32_t j = 0; j < 32; j++) conf = std::max(conf, ++cnt[(uint32_t)((double)(j * d) * avg / 4.0) & 31u]);
        if (conf < best_conf) {
            best_conf = conf;
            best_d = d;
        }
    }
    return best_d;
}

// Add the device time of launches whose event pairs have completed (all of them when wait is set).
static void busy_collect(fa_ctx *c, bool wait)
{
    while (c->busy_tail < c->busy_head) {
        cudaEvent_t *pair = c->ev_busy[c->busy_tail % fa_ctx::kBusyRing];
        if (wait) cudaEventSynchronize(pair[1]);
        else if (cudaEventQuery(pair[1]) != cudaSuccess) break;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pair[0], pair[1]) == cudaSuccess) c->busy_us += (double)ms * 1e3;
        c->busy_tail++;
    }
    cudaGetLastError();  // cudaErrorNotReady is not an error
}

// Launch decode(+aggregate) over records already in device memory.
static int launch_batch(fa_ctx *c, const uint8_t *d_buf, uint64_t base, uint64_t len, const uint32_t *d_offsets,
                        uint32_t n_records, uint32_t flags)
{
    if (n_records == 0) return FA_OK;
    TileParams tp{};
    SubmitParams &p = tp.p;
    p.buf = d_buf;
    p.base = base;
    p.len = len;
    p.offsets = d_offsets;
    p.n_records = n_records;
    p.framed = (flags & FA_FRAMED) ? 1u : 0u;
    fill_table_params(c, p);
    p.scale = (c->cfg.flags & FA_CFG_SCALE_SAMPLING) ? 1u : 0u;
    p.cms = c->d_cms;
    p.cms_depth = c->cfg.cms_depth;
    p.cms_wlog2 = c->cfg.cms_width_log2;
    p.admit_shift = c->admit_shift;
    p.hint_set = (uint32_t)(c->n_submits & 1u);
    p.sample_seed = (uint32_t)c->n_submits * 0x85EBCA6Bu;
    if (c->d_hot) c->hot_dirty = true;
    FA_CUDA(c, cudaMemsetAsync(&c->d_counters->hint[p.hint_set][0], 0, 8, c->stream));  // this submit's statistics start at zero
    // tile shape from the batch's mean record size: 256 records per CTA when their bytes fit
    // the shared-memory budget, fewer for fat records
    const double avg = (double)len / (double)n_records;
    static const uint32_t max_tile = getenv("FA_TILE_RECORDS") ? (uint32_t)atoi(getenv("FA_TILE_RECORDS")) : (uint32_t)kThreads;
    uint32_t tr = std::min<uint32_t>(std::max<uint32_t>(max_tile & ~31u, 32u), 1024u);
    while (tr > 32 && (double)tr * avg * 1.06 + 512.0 > (double)kTileBytesMax) tr -= 32;
    uint32_t tb = (uint32_t)((double)tr * avg * 1.06 + 512.0);
    tb = (tb + 1023u) & ~1023u;
    if (tb > (uint32_t)kTileBytesMax) tb = kTileBytesMax;
    if (tb < 4096u) tb = 4096u;
    p.tile_records = tr;
    p.tile_bytes = tb;
    const uint32_t d = pick_lane_stride(avg);
    p.lane_shift = d == 8 ? 3u : (d == 4 ? 2u : (d == 2 ? 1u : 0u));
    tp.c = c->cols;
    const uint32_t n_tiles = (n_records + tr - 1) / tr;
    cudaError_t e = cudaSuccess;
    busy_collect(c, c->busy_head - c->busy_tail >= (uint64_t)fa_ctx::kBusyRing);  // a full ring waits for its oldest pair
    cudaEvent_t *busy = c->ev_busy[c->busy_head % fa_ctx::kBusyRing];
    if (!busy[0]) {
        FA_CUDA(c, cudaEventCreate(&busy[0]));
        FA_CUDA(c, cudaEventCreate(&busy[1]));
    }
    FA_CUDA(c, cudaEventRecord(busy[0], c->stream));
    if (c->cfg.flags & FA_CFG_COLUMNS) {
        if (n_records > c->cfg.max_batch_records) return FA_ERR_INVALID;
        e = launch_tile<ColConsumer>(c, tp, n_tiles);
        c->cols_n = n_records;
        if (e == cudaSuccess && !(c->cfg.flags & FA_CFG_NO_AGGREGATE)) {
            const int g2 = (int)std::min<uint32_t>((n_records + kThreads - 1) / kThreads, (uint32_t)(c->num_sms * 8));
#define CALL_AGG(M) launch_agg_columns_mode<M>(c, p, g2)
            FA_DISPATCH_MODE(c->cfg.key_mode, CALL_AGG)
#undef CALL_AGG
        }
    } else {
#define CALL_FUSED(M) launch_fused_mode<M>(c, tp, n_tiles)
        FA_DISPATCH_MODE(c->cfg.key_mode, CALL_FUSED)
#undef CALL_FUSED
    }
    FA_CUDA(c, e);
    if (c->admit_shift && c->d_slots) {
        // FA_CFG_TOPK_ONLY: publish the total weight, keep only the candidates still above the bar (kernels.cuh: k_prune_*)
        const unsigned long long n_slots = c->capacity + 1;
        int rc = ensure_scratch(c, n_slots * sizeof(fa_row));
        if (rc) return rc;
        const int g = (int)std::min<uint64_t>((n_slots + 255) / 256, (uint64_t)c->num_sms * 8);
        // the heavy keys' updates sit in the replicas and their sketch weight in the slots: settle both first
        k_merge_hot_candidates<<<c->num_sms * 4, 256, 0, c->stream>>>(p, (unsigned long long)kHotReplicas * c->hot_slots_per_replica);
        k_apply_pending<<<g, 256, 0, c->stream>>>(p, c->capacity);
        c->n_kernels += 2;
        c->hot_dirty = false;
        k_publish_weight<<<1, 1, 0, c->stream>>>(c->d_counters, c->d_ts);
        k_prune_collect<<<g, 256, 0, c->stream>>>(c->d_slots, n_slots, c->d_cms, c->cfg.cms_depth, c->cfg.cms_width_log2, c->admit_shift,
                                                   c->d_counters, (fa_row *)c->d_scratch, c->d_ts);
        c->n_kernels += 2;
        FA_CUDA(c, cudaGetLastError());
        rc = table_init_at(c, c->d_slots, c->stream, false);
        if (rc) return rc;
        FA_CUDA(c, cudaMemsetAsync(&c->d_ts->n_groups, 0, 8, c->stream));
        FA_CUDA(c, cudaMemsetAsync(&c->d_ts->side_state, 0, 4, c->stream));
        SubmitParams pp{};
        fill_table_params(c, pp);
        k_prune_reinsert<<<g, 256, 0, c->stream>>>(pp, (const fa_row *)c->d_scratch);
        c->n_kernels++;
        FA_CUDA(c, cudaGetLastError());
    }
    FA_CUDA(c, cudaEventRecord(busy[1], c->stream));
    c->busy_head++;
    c->n_submits++;
    c->n_records += n_records;
    return FA_OK;
}

// ---------------------------------------------------------------------------------------------
// framing on the GPU (offsets == NULL): see frame.cuh
// ---------------------------------------------------------------------------------------------
#include "frame.cuh"

// ---------------------------------------------------------------------------------------------
// ingest
// ---------------------------------------------------------------------------------------------

static int ensure_staging(fa_ctx *c)
{
    if (c->d_stage[0]) return FA_OK;
    for (int i = 0; i < 2; i++) {
        FA_CUDA(c, cudaMalloc(&c->d_stage[i], c->cfg.max_batch_bytes + 256));
        FA_CUDA(c, cudaMalloc(&c->d_stage_off[i], ((size_t)c->cfg.max_batch_records + 1) * 4));
    }
    return FA_OK;
}

extern "C" int fa_host_buffer(fa_ctx *c, int slot, uint8_t **buf, size_t *cap_bytes, uint32_t **offsets, size_t *cap_records)
{
    if (!c || slot < 0 || slot > 1) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    if (!c->h_slab[slot]) {
        FA_CUDA(c, cudaHostAlloc(&c->h_slab[slot], c->cfg.max_batch_bytes, cudaHostAllocDefault));
        FA_CUDA(c, cudaHostAlloc(&c->h_slab_off[slot], ((size_t)c->cfg.max_batch_records + 1) * 4, cudaHostAllocDefault));
    } else {
        FA_CUDA(c, cudaEventSynchronize(c->ev_slab[slot]));
    }
    if (buf) *buf = c->h_slab[slot];
    if (cap_bytes) *cap_bytes = c->cfg.max_batch_bytes;
    if (offsets) *offsets = c->h_slab_off[slot];
    if (cap_records) *cap_records = c->cfg.max_batch_records;
    return FA_OK;
}

// index and launch the offsets-free batch the previous fa_submit staged (no-op when there is none)
static int finish_frame(fa_ctx *c)
{
    if (!c->frame_pending) return FA_OK;
    c->frame_pending = false;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    const int i = c->frame_stage;
    FA_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_staged[i], 0));
    uint32_t n_found = 0;
    int rc = frame_index_device(c, c->d_stage[i], c->frame_len, &n_found);
    if (rc == FA_OK) rc = launch_batch(c, c->d_stage[i], 0, c->frame_len, c->d_frame_off, n_found, c->frame_flags);
    cudaEventRecord(c->ev_consumed[i], c->stream);  // the stage is free again whatever happened to its batch
    return rc;
}
#define FA_DRAIN(c)                      \
    do {                                 \
        const int rc_ = finish_frame(c); \
        if (rc_) return rc_;             \
    } while (0)

extern "C" int fa_submit_device(fa_ctx *c, const uint8_t *d_buf, size_t len, const uint32_t *d_offsets, uint32_t n_records,
                                uint32_t flags)
{
    if (!c || (!d_buf && len)) return FA_ERR_INVALID;
    if (((uintptr_t)d_buf & 15u) || len > 0xFFFFFFF0ull) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    if (!d_offsets) {
        if (!(flags & FA_FRAMED)) return FA_ERR_INVALID;
        uint32_t n_found = 0;
        int rc = frame_index_device(c, d_buf, len, &n_found);
        if (rc) return rc;
        d_offsets = c->d_frame_off;
        n_records = n_found;
    }
    c->bytes_in += len;
    return launch_batch(c, d_buf, 0, len, d_offsets, n_records, flags);
}

extern "C" int fa_submit(fa_ctx *c, const uint8_t *buf, size_t len, const uint32_t *offsets, uint32_t n_records, uint32_t flags)
{
    if (!c || (!buf && len)) return FA_ERR_INVALID;
    if (len > 0xFFFFFFF0ull) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    int rc = ensure_staging(c);
    if (rc) return rc;
    if (!offsets) {
        // no boundaries from the host: stage the whole stream, index it on the GPU
        if (!(flags & FA_FRAMED)) return FA_ERR_INVALID;
        if (len > c->cfg.max_batch_bytes) return FA_ERR_INVALID;
        const int i = c->next_stage;
        c->next_stage ^= 1;
        FA_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[i], 0));
        FA_CUDA(c, cudaMemcpyAsync(c->d_stage[i], buf, len, cudaMemcpyHostToDevice, c->copy_stream));
        FA_CUDA(c, cudaEventRecord(c->ev_staged[i], c->copy_stream));
        for (int s = 0; s < 2; s++)
            if (buf >= c->h_slab[s] && c->h_slab[s] && buf < c->h_slab[s] + c->cfg.max_batch_bytes)
                FA_CUDA(c, cudaEventRecord(c->ev_slab[s], c->copy_stream));
        // the previous offsets-free batch (the other stage): its index's host look overlaps the copy just enqueued
        rc = finish_frame(c);
        c->frame_pending = true;
        c->frame_stage = i;
        c->frame_len = len;
        c->frame_flags = flags;
        c->bytes_in += len;
        return rc;
    }
    FA_DRAIN(c);
    if (n_records == 0) return FA_OK;
    // cut into batches at record boundaries; copy of batch i+1 overlaps kernel of batch i
    uint32_t r0 = 0;
    while (r0 < n_records) {
        const uint64_t a0 = offsets[r0] & ~15u;
        const uint64_t lim = a0 + c->cfg.max_batch_bytes;
        const uint32_t rmax = (uint32_t)std::min<uint64_t>((uint64_t)n_records, (uint64_t)r0 + c->cfg.max_batch_records);
        // largest r1 in (r0, rmax] with offsets[r1] <= lim  (offsets assumed non-decreasing; the
        // kernel re-validates every span)
        const uint32_t *lo = offsets + r0 + 1, *hi = offsets + rmax + 1;
        const uint32_t *it = std::upper_bound(lo, hi, (uint32_t)std::min<uint64_t>(lim, 0xFFFFFFFFull));
        uint32_t r1 = (uint32_t)(it - offsets) - 1;
        if (r1 <= r0) r1 = r0 + 1;
        // the batch must be a monotone run inside the buffer; stop at the first boundary that is not
        bool lone_bad = offsets[r0] > len;
        if (!lone_bad) {
            uint32_t prev = offsets[r0], v = r0 + 1;
            for (; v <= r1; v++) {
                const uint32_t o = offsets[v];
                if (o < prev || o > len) break;
                prev = o;
            }
            if (v <= r1) {
                if (v == r0 + 1) lone_bad = true;  // record r0 itself has an insane span
                else r1 = v - 1;
            }
        }
        if (lone_bad) r1 = r0 + 1;
        uint64_t b1 = lone_bad ? a0 : offsets[r1];
        if (!lone_bad && b1 - a0 > c->cfg.max_batch_bytes) {
            c->last_error = "a single record exceeds max_batch_bytes";
            return FA_ERR_INVALID;
        }
        const int i = c->next_stage;
        c->next_stage ^= 1;
        FA_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[i], 0));
        if (b1 > a0) FA_CUDA(c, cudaMemcpyAsync(c->d_stage[i], buf + a0, b1 - a0, cudaMemcpyHostToDevice, c->copy_stream));
        FA_CUDA(c, cudaMemcpyAsync(c->d_stage_off[i], offsets + r0, ((size_t)(r1 - r0) + 1) * 4, cudaMemcpyHostToDevice,
                                   c->copy_stream));
        FA_CUDA(c, cudaEventRecord(c->ev_staged[i], c->copy_stream));
        FA_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_staged[i], 0));
        // a lone record with an insane span is launched over an empty buffer: the kernel rejects and counts it
        rc = launch_batch(c, c->d_stage[i], lone_bad ? 0 : a0, b1 - a0, c->d_stage_off[i], r1 - r0, flags);
        if (rc) return rc;
        FA_CUDA(c, cudaEventRecord(c->ev_consumed[i], c->stream));
        r0 = r1;
    }
    for (int s = 0; s < 2; s++)
        if (c->h_slab[s] && buf >= c->h_slab[s] && buf < c->h_slab[s] + c->cfg.max_batch_bytes)
            FA_CUDA(c, cudaEventRecord(c->ev_slab[s], c->copy_stream));
    c->bytes_in += len;
    return FA_OK;
}

extern "C" int fa_sync(fa_ctx *c)
{
    if (!c) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    FA_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    return FA_OK;
}

// context counters and the CURRENT table's state -> h_counters / h_ts
static int read_counters(fa_ctx *c)
{
    FA_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaMemcpyAsync(c->h_ts, c->d_ts, sizeof(TableState), cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    return FA_OK;
}

extern "C" int fa_stats_get(fa_ctx *c, fa_stats *out)
{
    if (!c || !out) return FA_ERR_INVALID;
    int rc = fa_sync(c);
    if (rc) return rc;
    rc = merge_hot(c);
    if (rc) return rc;
    rc = read_counters(c);
    if (rc) return rc;
    out->n_records = c->n_records;
    out->n_bad = c->h_counters->n_bad;
    out->n_nokey = c->h_counters->n_nokey;
    out->n_dropped = c->h_ts->n_dropped;
    out->n_groups = c->h_ts->n_groups;
    out->n_submits = c->n_submits;
    out->bytes_in = c->bytes_in;
    out->n_kernels = c->n_kernels;
    busy_collect(c, true);
    out->gpu_busy_us = (uint64_t)c->busy_us;
    return FA_OK;
}

// ---------------------------------------------------------------------------------------------
// emit
// ---------------------------------------------------------------------------------------------

template <int KW>
static cudaError_t launch_compact(fa_ctx *c, fa_row *d_rows, unsigned long long cap)
{
    const unsigned long long n_slots = c->capacity + 1;  // + the side slot
    const int grid = (int)std::min<uint64_t>((c->capacity + 255) / 256, (uint64_t)c->num_sms * 16);
    k_compact_rows<KW><<<grid, 256, 0, c->stream>>>(c->d_slots, n_slots, d_rows, cap, c->d_ts);
    c->n_kernels++;
    return cudaGetLastError();
}

template <int KW>
static cudaError_t launch_estimate(fa_ctx *c, const unsigned long long *cms, fa_hh *d_out, unsigned long long cap)
{
    const unsigned long long n_slots = c->capacity + 1;
    const int grid = (int)std::min<uint64_t>((c->capacity + 255) / 256, (uint64_t)c->num_sms * 16);
    k_estimate<KW><<<grid, 256, 0, c->stream>>>(c->d_slots, n_slots, cms, c->cfg.cms_depth, c->cfg.cms_width_log2, d_out, cap,
                                                c->d_ts);
    c->n_kernels++;
    return cudaGetLastError();
}

#define FA_DISPATCH_KW(kw, CALL)     \
    switch (kw) {                    \
    case 1: e = CALL(1); break;      \
    case 2: e = CALL(2); break;      \
    case 4: e = CALL(4); break;      \
    default: e = CALL(11); break;    \
    }

static int reset_table(fa_ctx *c)
{
    int rc = table_init(c);
    if (rc) return rc;
    FA_CUDA(c, cudaMemsetAsync(&c->d_ts->n_groups, 0, 8, c->stream));
    FA_CUDA(c, cudaMemsetAsync(&c->d_ts->n_dropped, 0, 8, c->stream));
    FA_CUDA(c, cudaMemsetAsync(&c->d_ts->side_state, 0, 4, c->stream));
    return FA_OK;
}

// ---- ORDER BY on the device: stable LSD radix sort (cub) over the key words, last word first ----
__global__ void k_iota(uint32_t *perm, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = i;
}
__global__ void k_sort_key(const fa_row *rows, const uint32_t *perm, uint32_t n, int word, uint32_t *keys)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = rows[perm[i]].key[word];
}
// speculative variant: only the first counters->flush_rows rows exist; the rest of [0,n) sorts to the end
__global__ void k_sort_key_guarded(const fa_row *rows, const uint32_t *perm, uint32_t n, int word, uint32_t *keys, const TableState *counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t src = perm[i];
        keys[i] = (unsigned long long)src < counters->flush_rows ? rows[src].key[word] : 0xFFFFFFFFu;
    }
}
__global__ void k_gather_rows(const fa_row *in, const uint32_t *perm, uint32_t n, fa_row *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

// rows (unsorted, device) -> sorted (device); returns the pointer holding the sorted rows
static int sort_rows_device(fa_ctx *c, size_t groups, fa_row **sorted)
{
    const uint32_t n = (uint32_t)groups;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (int)n, 0, 32, c->stream);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t rows_b = al(groups * sizeof(fa_row)), u32_b = al(groups * 4);
    const size_t need = 2 * rows_b + 4 * u32_b + al(cub_bytes);
    // the unsorted rows already sit at the start of the scratch block: grow it without losing them
    if (need > c->scratch_bytes) {
        void *bigger = nullptr;
        FA_CUDA(c, cudaMalloc(&bigger, need));
        FA_CUDA(c, cudaMemcpyAsync(bigger, c->d_scratch, groups * sizeof(fa_row), cudaMemcpyDeviceToDevice, c->stream));
        FA_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_scratch);
        c->d_scratch = bigger;
        c->scratch_bytes = need;
    }
    uint8_t *base = (uint8_t *)c->d_scratch;
    fa_row *rows_in = (fa_row *)base;
    fa_row *rows_out = (fa_row *)(base + rows_b);
    uint32_t *keys_a = (uint32_t *)(base + 2 * rows_b), *keys_b = (uint32_t *)(base + 2 * rows_b + u32_b);
    uint32_t *perm_a = (uint32_t *)(base + 2 * rows_b + 2 * u32_b), *perm_b = (uint32_t *)(base + 2 * rows_b + 3 * u32_b);
    void *cub_tmp = base + 2 * rows_b + 4 * u32_b;
    const int g = (int)((n + 255) / 256);
    k_iota<<<g, 256, 0, c->stream>>>(perm_a, n);
    c->n_kernels++;
    for (int w = c->kw - 1; w >= 0; w--) {
        k_sort_key<<<g, 256, 0, c->stream>>>(rows_in, perm_a, n, w, keys_a);
        c->n_kernels++;
        FA_CUDA(c, cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, perm_a, perm_b, (int)n, 0, 32, c->stream));
        std::swap(perm_a, perm_b);
    }
    k_gather_rows<<<g, 256, 0, c->stream>>>(rows_in, perm_a, n, rows_out);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    *sorted = rows_out;
    return FA_OK;
}

// Steady-state flush with ONE synchronisation.  The row count is guessed from the previous flush (a 5-minute
// roll-up has about as many groups as the last one), so compaction, ORDER BY and the copies are all enqueued at
// once -- behind the kernels that are still running -- and the host waits a single time.  Compaction keeps
// every row in scratch, so a wrong guess loses nothing: *done stays false and the exact path re-reads them.
static int flush_speculative(fa_ctx *c, fa_row *rows, size_t cap, size_t *n, uint32_t flags, bool *done)
{
    *done = false;
    const uint64_t all_rows = c->capacity + 1;                     // compaction bound
    const uint64_t m64 = std::min<uint64_t>({(uint64_t)cap, all_rows, c->last_groups + c->last_groups / 8 + 256});
    if (all_rows > (1ull << 22) || m64 < c->last_groups || m64 >= (1ull << 31)) return FA_OK;  // big tables / small caller array: exact path
    const uint32_t m = (uint32_t)m64;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)m,
                                    0, 32, c->stream);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t in_b = al(all_rows * sizeof(fa_row)), out_b = al((size_t)m * sizeof(fa_row)), u32_b = al((size_t)m * 4);
    int rc = ensure_scratch(c, in_b + out_b + 4 * u32_b + al(cub_bytes));
    if (rc) return rc;
    uint8_t *base = (uint8_t *)c->d_scratch;
    fa_row *rows_in = (fa_row *)base, *rows_out = (fa_row *)(base + in_b);
    uint32_t *keys_a = (uint32_t *)(base + in_b + out_b), *keys_b = (uint32_t *)(base + in_b + out_b + u32_b);
    uint32_t *perm_a = (uint32_t *)(base + in_b + out_b + 2 * u32_b), *perm_b = (uint32_t *)(base + in_b + out_b + 3 * u32_b);
    void *cub_tmp = base + in_b + out_b + 4 * u32_b;
    FA_CUDA(c, cudaMemsetAsync(&c->d_ts->flush_rows, 0, 8, c->stream));
    cudaError_t e;
#define CALL_COMPACT(K) launch_compact<K>(c, rows_in, all_rows)
    FA_DISPATCH_KW(c->kw, CALL_COMPACT)
#undef CALL_COMPACT
    FA_CUDA(c, e);
    const int g = (int)((m + 255) / 256);
    k_iota<<<g, 256, 0, c->stream>>>(perm_a, m);
    c->n_kernels++;
    for (int w = c->kw - 1; w >= 0; w--) {
        k_sort_key_guarded<<<g, 256, 0, c->stream>>>(rows_in, perm_a, m, w, keys_a, c->d_ts);
        c->n_kernels++;
        FA_CUDA(c, cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, perm_a, perm_b, (int)m, 0, 32, c->stream));
        std::swap(perm_a, perm_b);
    }
    k_gather_rows<<<g, 256, 0, c->stream>>>(rows_in, perm_a, m, rows_out);  // entries past the true count are never read
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    const size_t bytes = (size_t)m * sizeof(fa_row);
    cudaPointerAttributes attr{};
    const bool pinned_dst = cudaPointerGetAttributes(&attr, rows) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    void *dst = rows;
    if (!pinned_dst) {
        if (bytes > c->bounce_bytes) {
            if (c->h_bounce) cudaFreeHost(c->h_bounce);
            c->h_bounce = nullptr;
            c->bounce_bytes = 0;
            const size_t want = std::max<size_t>(bytes * 2, 1u << 20);
            FA_CUDA(c, cudaHostAlloc(&c->h_bounce, want, cudaHostAllocDefault));
            c->bounce_bytes = want;
        }
        dst = c->h_bounce;
    }
    FA_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaMemcpyAsync(c->h_ts, c->d_ts, sizeof(TableState), cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaMemcpyAsync(dst, rows_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaStreamSynchronize(c->stream));  // the one wait
    const uint64_t groups = c->h_ts->n_groups, dropped = c->h_ts->n_dropped;
    if (groups > m) {  // the roll-up grew by more than 1/8: every row is still in scratch, the exact path takes over
        c->last_groups = groups;
        return FA_OK;
    }
    c->last_groups = groups;
    *n = (size_t)groups;
    if (!pinned_dst) memcpy(rows, c->h_bounce, (size_t)groups * sizeof(fa_row));
    if (!(flags & FA_FLUSH_KEEP)) {
        rc = reset_table(c);  // enqueued only now that the rows are safely on the host; the next submit queues behind it
        if (rc) return rc;
    }
    *done = true;
    return dropped ? FA_ERR_TABLE_FULL : FA_OK;
}

// ---------------------------------------------------------------------------------------------
// asynchronous flush: swap tables, drain the filled one on a side stream
// ---------------------------------------------------------------------------------------------
//
// (*state).flush holds the inserter's global mutex while it talks to the database (inserter.go:90-111), so buffer()
// stalls behind every flush.  Here the flush must not stall the stream either: fa_flush_begin folds the hot-key replicas,
// swaps the filled table for a spare, empty one -- the very next fa_submit already aggregates into the spare -- and
// enqueues compaction, ORDER BY, the copies to pinned host memory and the emptying of the old table on the context's
// flush stream; fa_flush_end waits for that (a blocking, non-spinning event wait) and hands the rows over.

template <int KW>
static cudaError_t launch_compact_at(fa_ctx *c, const uint8_t *slots, TableState *ts, fa_row *d_rows, unsigned long long cap, cudaStream_t st)
{
    const unsigned long long n_slots = c->capacity + 1;  // + the side slot
    const int grid = (int)std::min<uint64_t>((c->capacity + 255) / 256, (uint64_t)c->num_sms * 16);
    k_compact_rows<KW><<<grid, 256, 0, st>>>(slots, n_slots, d_rows, cap, ts);
    c->n_kernels++;
    return cudaGetLastError();
}

static bool flush_async_ok(const fa_ctx *c) { return c->d_slots && c->capacity + 1 <= (1ull << 22); }

// ORDER BY of the first min(m, ts->flush_rows) rows of rows_in into rows_out, on `st` (stable LSD radix sort over the
// key words, last word first; entries past the true row count sort to the end and are never read)
static int sort_rows_guarded(fa_ctx *c, cudaStream_t st, const fa_row *rows_in, fa_row *rows_out, uint32_t m, const TableState *ts, uint8_t *work,
                             size_t cub_bytes)
{
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t u32_b = al((size_t)m * 4);
    uint32_t *keys_a = (uint32_t *)work, *keys_b = (uint32_t *)(work + u32_b);
    uint32_t *perm_a = (uint32_t *)(work + 2 * u32_b), *perm_b = (uint32_t *)(work + 3 * u32_b);
    void *cub_tmp = work + 4 * u32_b;
    const int g = (int)((m + 255) / 256);
    k_iota<<<g, 256, 0, st>>>(perm_a, m);
    c->n_kernels++;
    for (int w = c->kw - 1; w >= 0; w--) {
        k_sort_key_guarded<<<g, 256, 0, st>>>(rows_in, perm_a, m, w, keys_a, ts);
        c->n_kernels++;
        FA_CUDA(c, cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, perm_a, perm_b, (int)m, 0, 32, st));
        std::swap(perm_a, perm_b);
    }
    k_gather_rows<<<g, 256, 0, st>>>(rows_in, perm_a, m, rows_out);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    return FA_OK;
}

extern "C" int fa_flush_begin(fa_ctx *c, uint32_t flags)
{
    if (!c || !c->d_slots || c->flush_pending) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    c->flush_flags = flags;
    if ((flags & (FA_FLUSH_KEEP | FA_FLUSH_UNSORTED)) || !flush_async_ok(c)) {
        // peeks, unsorted dumps and huge tables are drained by fa_flush_end itself, in place
        c->flush_pending = c->flush_deferred = true;
        return FA_OK;
    }
    if (!c->flush_stream) {
        // highest priority: the drain's small kernels must slip in between the CTAs of the fused kernels that are filling
        // the spare table (those retire every few microseconds), not wait for a whole kernel to end
        int prio_lo = 0, prio_hi = 0;
        FA_CUDA(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        FA_CUDA(c, cudaStreamCreateWithPriority(&c->flush_stream, cudaStreamNonBlocking, prio_hi));
        FA_CUDA(c, cudaEventCreateWithFlags(&c->ev_swap, cudaEventDisableTiming));
        FA_CUDA(c, cudaEventCreateWithFlags(&c->ev_spare_ready, cudaEventDisableTiming));
        FA_CUDA(c, cudaEventCreateWithFlags(&c->ev_flush_done, cudaEventDisableTiming | cudaEventBlockingSync));
        FA_CUDA(c, cudaMalloc(&c->d_slots_spare, table_alloc_bytes(c->kw, c->capacity + 1)));
        FA_CUDA(c, cudaMalloc(&c->d_ts_spare, sizeof(TableState)));
        FA_CUDA(c, cudaMemsetAsync(c->d_ts_spare, 0, sizeof(TableState), c->flush_stream));
        int rc0 = table_init_at(c, c->d_slots_spare, c->flush_stream, false);
        if (rc0) return rc0;
        FA_CUDA(c, cudaEventRecord(c->ev_spare_ready, c->flush_stream));
    }
    int rc = merge_hot(c);  // on the main stream, into the table about to be drained
    if (rc) return rc;
    // rows to sort and copy: a 5-minute roll-up has about as many groups as the last one; compaction keeps EVERY row in
    // scratch, so a wrong guess loses nothing (fa_flush_end re-sorts the exact count)
    const uint64_t all_rows = c->capacity + 1;
    const uint32_t m = (uint32_t)std::min<uint64_t>(all_rows, c->last_groups ? c->last_groups + c->last_groups / 8 + 256 : all_rows);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (int)all_rows, 0, 32, c->flush_stream);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t in_b = al(all_rows * sizeof(fa_row)), work_b = 4 * al(all_rows * 4) + al(cub_bytes);
    const size_t need = 2 * in_b + work_b;
    if (need > c->fscratch_bytes) {
        FA_CUDA(c, cudaStreamSynchronize(c->flush_stream));
        if (c->d_fscratch) cudaFree(c->d_fscratch);
        c->d_fscratch = nullptr;
        c->fscratch_bytes = 0;
        FA_CUDA(c, cudaMalloc(&c->d_fscratch, need));
        c->fscratch_bytes = need;
    }
    const size_t host_b = (size_t)all_rows * sizeof(fa_row);
    if (host_b > c->bounce_bytes) {
        if (c->h_bounce) cudaFreeHost(c->h_bounce);
        c->h_bounce = nullptr;
        c->bounce_bytes = 0;
        FA_CUDA(c, cudaHostAlloc(&c->h_bounce, host_b, cudaHostAllocDefault));
        c->bounce_bytes = host_b;
    }
    uint8_t *base = (uint8_t *)c->d_fscratch;
    fa_row *rows_in = (fa_row *)base, *rows_out = (fa_row *)(base + in_b);
    // swap: the spare (emptied by the previous drain) takes over
    FA_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_spare_ready, 0));
    FA_CUDA(c, cudaEventRecord(c->ev_swap, c->stream));
    uint8_t *old_slots = c->d_slots;
    TableState *old_ts = c->d_ts;
    c->d_slots = c->d_slots_spare;
    c->d_ts = c->d_ts_spare;
    c->d_slots_spare = old_slots;
    c->d_ts_spare = old_ts;
    // the drain
    cudaStream_t fs = c->flush_stream;
    FA_CUDA(c, cudaStreamWaitEvent(fs, c->ev_swap, 0));
    FA_CUDA(c, cudaMemsetAsync(&old_ts->flush_rows, 0, 8, fs));
    cudaError_t e;
#define CALL_COMPACT(K) launch_compact_at<K>(c, old_slots, old_ts, rows_in, all_rows, fs)
    FA_DISPATCH_KW(c->kw, CALL_COMPACT)
#undef CALL_COMPACT
    FA_CUDA(c, e);
    rc = sort_rows_guarded(c, fs, rows_in, rows_out, m, old_ts, base + 2 * in_b, cub_bytes);
    if (rc) return rc;
    FA_CUDA(c, cudaMemcpyAsync(c->h_ts_drain, old_ts, sizeof(TableState), cudaMemcpyDeviceToHost, fs));
    FA_CUDA(c, cudaMemcpyAsync(c->h_bounce, rows_out, (size_t)m * sizeof(fa_row), cudaMemcpyDeviceToHost, fs));
    rc = table_init_at(c, old_slots, fs, false);
    if (rc) return rc;
    FA_CUDA(c, cudaMemsetAsync(old_ts, 0, sizeof(TableState), fs));
    FA_CUDA(c, cudaEventRecord(c->ev_spare_ready, fs));
    FA_CUDA(c, cudaEventRecord(c->ev_flush_done, fs));
    c->flush_m = m;
    c->flush_rows_in = rows_in;
    c->flush_pending = true;
    c->flush_deferred = false;
    return FA_OK;
}

static int flush_sync(fa_ctx *c, fa_row *rows, size_t cap, size_t *n, uint32_t flags);

extern "C" int fa_flush_end(fa_ctx *c, fa_row *rows, size_t cap, size_t *n)
{
    if (!c || !n || !c->flush_pending) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    if (c->flush_deferred) {
        FA_DRAIN(c);
        const int rc = flush_sync(c, rows, cap, n, c->flush_flags);
        if (rc != FA_ERR_CAPACITY) c->flush_pending = c->flush_deferred = false;
        return rc;
    }
    FA_CUDA(c, cudaEventSynchronize(c->ev_flush_done));  // blocking wait: the host thread sleeps, it does not spin
    const uint64_t groups = c->h_ts_drain->n_groups, dropped = c->h_ts_drain->n_dropped;
    *n = (size_t)groups;
    if (groups > cap || (groups && !rows)) return FA_ERR_CAPACITY;  // the rows stay in scratch: call again with a larger array
    if (groups > c->flush_m) {
        // the roll-up grew by more than 1/8 since the last flush: every row is still in scratch, order the exact count
        const uint32_t m = (uint32_t)groups;
        size_t cub_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (int)(c->capacity + 1), 0, 32, c->flush_stream);
        auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t in_b = al((c->capacity + 1) * sizeof(fa_row));
        uint8_t *base = (uint8_t *)c->d_fscratch;
        fa_row *rows_out = (fa_row *)(base + in_b);
        // the drained table's state was zeroed; the guard needs the row count again
        TableState tmp{};
        tmp.flush_rows = groups;
        FA_CUDA(c, cudaMemcpyAsync(c->d_ts_spare, &tmp, sizeof tmp, cudaMemcpyHostToDevice, c->flush_stream));
        int rc = sort_rows_guarded(c, c->flush_stream, c->flush_rows_in, rows_out, m, c->d_ts_spare, base + 2 * in_b, cub_bytes);
        if (rc) return rc;
        FA_CUDA(c, cudaMemcpyAsync(c->h_bounce, rows_out, (size_t)m * sizeof(fa_row), cudaMemcpyDeviceToHost, c->flush_stream));
        FA_CUDA(c, cudaMemsetAsync(c->d_ts_spare, 0, sizeof(TableState), c->flush_stream));
        FA_CUDA(c, cudaEventRecord(c->ev_spare_ready, c->flush_stream));
        FA_CUDA(c, cudaStreamSynchronize(c->flush_stream));
    }
    if (groups) memcpy(rows, c->h_bounce, (size_t)groups * sizeof(fa_row));
    c->last_groups = groups;
    c->flush_pending = false;
    return dropped ? FA_ERR_TABLE_FULL : FA_OK;
}

extern "C" int fa_flush(fa_ctx *c, fa_row *rows, size_t cap, size_t *n, uint32_t flags)
{
    if (!c || !n || !c->d_slots || c->flush_pending) return FA_ERR_INVALID;
    FA_DRAIN(c);
    return flush_sync(c, rows, cap, n, flags);
}

// the synchronous flush: everything on the context's own stream, in place
static int flush_sync(fa_ctx *c, fa_row *rows, size_t cap, size_t *n, uint32_t flags)
{
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    static const bool dbg = getenv("FA_DEBUG_FLUSH") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
    };
    const auto t_in = now();
    if (dbg) cudaStreamSynchronize(c->stream);  // separates "waiting for the submitted kernels" from the flush itself
    const auto t_start = now();
    // everything submitted so far is ordered before this on the context's stream (copies hand over through
    // events), so no host-side wait is needed before enqueueing
    int rc = merge_hot(c);
    if (rc) return rc;
    if (!(flags & FA_FLUSH_UNSORTED) && c->last_groups && rows) {
        bool done = false;
        rc = flush_speculative(c, rows, cap, n, flags, &done);
        if (rc || done) return rc;
    }
    // exact path: two synchronisations, one to learn the row count, one for the rows
    rc = read_counters(c);
    if (rc) return rc;
    const uint64_t groups = c->h_ts->n_groups;
    const uint64_t dropped = c->h_ts->n_dropped;
    c->last_groups = groups;
    const auto t_count = now();
    *n = (size_t)groups;
    if (groups > cap || (groups && !rows)) return FA_ERR_CAPACITY;
    if (groups >= (1ull << 31)) return FA_ERR_INVALID;
    if (groups) {
        rc = ensure_scratch(c, groups * sizeof(fa_row));
        if (rc) return rc;
        FA_CUDA(c, cudaMemsetAsync(&c->d_ts->flush_rows, 0, 8, c->stream));
        cudaError_t e;
#define CALL_COMPACT(K) launch_compact<K>(c, (fa_row *)c->d_scratch, groups)
        FA_DISPATCH_KW(c->kw, CALL_COMPACT)
#undef CALL_COMPACT
        FA_CUDA(c, e);
        fa_row *src = (fa_row *)c->d_scratch;
        // ORDER BY (Date, Timeslot, SrcAS, DstAS, ETypeMap.EType): create.sh:90
        if (!(flags & FA_FLUSH_UNSORTED)) {
            rc = sort_rows_device(c, groups, &src);
            if (rc) return rc;
        }
        const size_t bytes = groups * sizeof(fa_row);
        // a pinned destination (cudaHostAlloc / cudaHostRegister / a torch pinned tensor) takes the rows directly;
        // pageable memory goes through the context's pinned landing area and one host memcpy
        cudaPointerAttributes attr{};
        const bool pinned_dst = cudaPointerGetAttributes(&attr, rows) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();  // an unregistered pointer is reported as an error on some drivers: not ours to keep
        if (!pinned_dst && bytes <= (256u << 20)) {
            if (bytes > c->bounce_bytes) {
                if (c->h_bounce) cudaFreeHost(c->h_bounce);
                c->h_bounce = nullptr;
                c->bounce_bytes = 0;
                const size_t want = std::max<size_t>(bytes * 2, 1u << 20);
                FA_CUDA(c, cudaHostAlloc(&c->h_bounce, want, cudaHostAllocDefault));
                c->bounce_bytes = want;
            }
            FA_CUDA(c, cudaMemcpyAsync(c->h_bounce, src, bytes, cudaMemcpyDeviceToHost, c->stream));
            if (!(flags & FA_FLUSH_KEEP)) {
                rc = reset_table(c);  // overlaps the copy on the same stream order
                if (rc) return rc;
            }
            const auto t_enq = now();
            FA_CUDA(c, cudaStreamSynchronize(c->stream));
            const auto t_sync = now();
            memcpy(rows, c->h_bounce, bytes);
            if (dbg)
                fprintf(stderr, "[fa_flush] wait-for-kernels %ld us | merge+count %ld us | enqueue compact/sort/copy/reset %ld us | device work + sync %ld us | memcpy %ld us\n",
                        us(t_in, t_start), us(t_start, t_count), us(t_count, t_enq), us(t_enq, t_sync), us(t_sync, now()));
            return dropped ? FA_ERR_TABLE_FULL : FA_OK;
        }
        FA_CUDA(c, cudaMemcpyAsync(rows, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        FA_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (!(flags & FA_FLUSH_KEEP)) {
        rc = reset_table(c);
        if (rc) return rc;
    }
    return dropped ? FA_ERR_TABLE_FULL : FA_OK;
}

// ---------------------------------------------------------------------------------------------
// merging aggregates: one context's rows into another's table; the box-wide exchange
// ---------------------------------------------------------------------------------------------

template <int KW>
static cudaError_t launch_add_rows(fa_ctx *c, const fa_row *d_rows, unsigned long long n, uint32_t owner, uint32_t n_owners)
{
    SubmitParams p{};
    fill_table_params(c, p);
    const int grid = (int)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)c->num_sms * 16);
    k_add_rows<KW><<<grid, 256, 0, c->stream>>>(p, d_rows, n, owner, n_owners);
    c->n_kernels++;
    return cudaGetLastError();
}

static int add_rows_device(fa_ctx *c, const fa_row *d_rows, size_t n, uint32_t owner, uint32_t n_owners)
{
    if (!n) return FA_OK;
    cudaError_t e;
#define CALL_ADD(K) launch_add_rows<K>(c, d_rows, n, owner, n_owners)
    FA_DISPATCH_KW(c->kw, CALL_ADD)
#undef CALL_ADD
    FA_CUDA(c, e);
    return FA_OK;
}

extern "C" int fa_merge_rows(fa_ctx *c, const fa_row *rows, size_t n, uint32_t owner, uint32_t n_owners)
{
    if (c && c->flush_pending) return FA_ERR_INVALID;  // fa_flush_end first
    if (!c || !c->d_slots || (n && !rows) || (n_owners > 1 && owner >= n_owners)) return FA_ERR_INVALID;
    if (!n) return FA_OK;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    cudaPointerAttributes attr{};
    const bool on_device = cudaPointerGetAttributes(&attr, rows) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (on_device) {
        if (attr.device != c->cfg.device) {  // a peer GPU's rows: read in place when the topology allows it
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->cfg.device, attr.device);
            if (can) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                cudaGetLastError();
            }
            if (!can) {
                fa_row *tmp = nullptr;
                FA_CUDA(c, cudaMalloc(&tmp, n * sizeof(fa_row)));
                cudaError_t ce = cudaMemcpyPeerAsync(tmp, c->cfg.device, rows, attr.device, n * sizeof(fa_row), c->stream);
                int rc = ce == cudaSuccess ? add_rows_device(c, tmp, n, owner, n_owners) : FA_ERR_CUDA;
                cudaStreamSynchronize(c->stream);
                cudaFree(tmp);
                FA_CUDA(c, ce);
                return rc;
            }
        }
        return add_rows_device(c, rows, n, owner, n_owners);
    }
    // host rows: through a device copy that lives until the kernel has read it
    fa_row *tmp = nullptr;
    FA_CUDA(c, cudaMalloc(&tmp, n * sizeof(fa_row)));
    cudaError_t ce = cudaMemcpyAsync(tmp, rows, n * sizeof(fa_row), cudaMemcpyHostToDevice, c->stream);
    int rc = ce == cudaSuccess ? add_rows_device(c, tmp, n, owner, n_owners) : FA_ERR_CUDA;
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    FA_CUDA(c, ce);
    return rc;
}

static int key_words_of_mode(int key_mode)
{
    switch (key_mode) {
    case FA_KEY_FLOWS5M: return KeyTraits<FA_KEY_FLOWS5M>::KW;
    case FA_KEY_ASPAIR: return KeyTraits<FA_KEY_ASPAIR>::KW;
    case FA_KEY_SRCADDR: return KeyTraits<FA_KEY_SRCADDR>::KW;
    case FA_KEY_DSTADDR: return KeyTraits<FA_KEY_DSTADDR>::KW;
    case FA_KEY_5TUPLE: return KeyTraits<FA_KEY_5TUPLE>::KW;
    case FA_KEY_SRCPORT: return KeyTraits<FA_KEY_SRCPORT>::KW;
    case FA_KEY_DSTPORT: return KeyTraits<FA_KEY_DSTPORT>::KW;
    default: return 0;
    }
}

extern "C" int fa_row_owner(int key_mode, const fa_row *rows, size_t n, uint32_t n_owners, uint32_t *owner)
{
    const int kw = key_words_of_mode(key_mode);
    if (!kw || !n_owners || (n && (!rows || !owner))) return FA_ERR_INVALID;
    for (size_t i = 0; i < n; i++) {
        unsigned long long h;
        switch (kw) {
        case 1: h = hash64<1>(rows[i].key); break;
        case 2: h = hash64<2>(rows[i].key); break;
        case 4: h = hash64<4>(rows[i].key); break;
        default: h = hash64<11>(rows[i].key); break;
        }
        owner[i] = key_owner(h, n_owners);
    }
    return FA_OK;
}

static bool row_key_less(const fa_row &a, const fa_row &b, int kw)
{
    for (int k = 0; k < kw; k++)
        if (a.key[k] != b.key[k]) return a.key[k] < b.key[k];
    return false;
}

extern "C" int fa_flush_box(fa_ctx *const *ctxs, int n_ctx, fa_row *rows, size_t cap, size_t *n, uint32_t flags)
{
    if (!ctxs || n_ctx < 1 || !n || (flags & FA_FLUSH_KEEP)) return FA_ERR_INVALID;
    for (int i = 0; i < n_ctx; i++)
        if (!ctxs[i] || !ctxs[i]->d_slots || ctxs[i]->cfg.key_mode != ctxs[0]->cfg.key_mode || ctxs[i]->flush_pending) return FA_ERR_INVALID;
    if (n_ctx == 1) return fa_flush(ctxs[0], rows, cap, n, flags);
    for (int i = 0; i < n_ctx; i++) FA_DRAIN(ctxs[i]);
    // 1. every context: fold the replicas, compact its rows (unsorted) into its scratch block, empty its table
    std::vector<uint64_t> groups(n_ctx, 0);
    uint64_t dropped = 0;
    for (int i = 0; i < n_ctx; i++) {
        fa_ctx *c = ctxs[i];
        FA_CUDA(c, cudaSetDevice(c->cfg.device));
        int rc = merge_hot(c);
        if (rc) return rc;
        rc = read_counters(c);
        if (rc) return rc;
        groups[i] = c->h_ts->n_groups;
        dropped += c->h_ts->n_dropped;
        if (groups[i]) {
            rc = ensure_scratch(c, groups[i] * sizeof(fa_row));
            if (rc) return rc;
            FA_CUDA(c, cudaMemsetAsync(&c->d_ts->flush_rows, 0, 8, c->stream));
            cudaError_t e;
#define CALL_COMPACT(K) launch_compact<K>(c, (fa_row *)c->d_scratch, groups[i])
            FA_DISPATCH_KW(c->kw, CALL_COMPACT)
#undef CALL_COMPACT
            FA_CUDA(c, e);
        }
        rc = reset_table(c);
        if (rc) return rc;
    }
    for (int i = 0; i < n_ctx; i++) {
        FA_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->cfg.device));
        FA_CUDA(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
    }
    // 2. the exchange: context j sums the rows it owns out of every context's compacted rows
    for (int j = 0; j < n_ctx; j++)
        for (int i = 0; i < n_ctx; i++) {
            const int src = (i + j) % n_ctx;  // stagger the sources so that the peers are not all read in the same order
            int rc = fa_merge_rows(ctxs[j], (const fa_row *)ctxs[src]->d_scratch, (size_t)groups[src], (uint32_t)j, (uint32_t)n_ctx);
            if (rc) return rc;
        }
    uint64_t total = 0;
    for (int j = 0; j < n_ctx; j++) {  // every reader is done before any scratch block is reused below
        fa_ctx *c = ctxs[j];
        FA_CUDA(c, cudaSetDevice(c->cfg.device));
        int rc = read_counters(c);
        if (rc) return rc;
        total += c->h_ts->n_groups;
        dropped += c->h_ts->n_dropped;
    }
    *n = (size_t)total;
    if (total > cap || (total && !rows)) return FA_ERR_CAPACITY;
    // 3. every context emits its share; the shares are disjoint, so the box order is a merge of sorted runs
    std::vector<size_t> cut(1, 0);
    size_t at = 0;
    for (int j = 0; j < n_ctx; j++) {
        size_t m = 0;
        int rc = fa_flush(ctxs[j], rows + at, cap - at, &m, flags & FA_FLUSH_UNSORTED);
        if (rc && rc != FA_ERR_TABLE_FULL) return rc;
        at += m;
        cut.push_back(at);
    }
    if (!(flags & FA_FLUSH_UNSORTED)) {
        const int kw = ctxs[0]->kw;
        auto less = [kw](const fa_row &a, const fa_row &b) { return row_key_less(a, b, kw); };
        for (size_t width = 1; width < (size_t)n_ctx; width *= 2)  // pairwise merges of neighbouring runs
            for (size_t j = 0; j + width < (size_t)n_ctx; j += 2 * width)
                std::inplace_merge(rows + cut[j], rows + cut[j + width], rows + cut[std::min<size_t>(j + 2 * width, (size_t)n_ctx)], less);
    }
    return dropped ? FA_ERR_TABLE_FULL : FA_OK;
}

extern "C" int fa_reset(fa_ctx *c)
{
    if (c && c->flush_pending) return FA_ERR_INVALID;  // fa_flush_end first
    if (!c) return FA_ERR_INVALID;
    int rc = fa_sync(c);
    if (rc) return rc;
    rc = reset_table(c);
    if (rc) return rc;
    if (c->d_cms) FA_CUDA(c, cudaMemsetAsync(c->d_cms, 0, c->cms_words * 8, c->stream));
    FA_CUDA(c, cudaMemsetAsync(c->d_counters, 0, sizeof(Counters), c->stream));
    FA_CUDA(c, cudaMemsetAsync(c->d_ts, 0, sizeof(TableState), c->stream));
    FA_CUDA(c, cudaMemsetAsync(&c->d_counters->hint[1][0], 1, 8, c->stream));
    c->n_records = c->n_submits = c->bytes_in = 0;
    c->n_kernels = 0;
    busy_collect(c, true);
    c->busy_us = 0;
    return fa_sync(c);
}

// ---------------------------------------------------------------------------------------------
// sketch / heavy hitters
// ---------------------------------------------------------------------------------------------

extern "C" int fa_cms_read(fa_ctx *c, uint64_t *out, size_t cap_words)
{
    if (!c || !out || !c->d_cms) return FA_ERR_INVALID;
    if (cap_words < c->cms_words) return FA_ERR_CAPACITY;
    int rc = fa_sync(c);
    if (rc) return rc;
    FA_CUDA(c, cudaMemcpy(out, c->d_cms, c->cms_words * 8, cudaMemcpyDeviceToHost));
    return FA_OK;
}

static int ensure_cms_global(fa_ctx *c)
{
    if (!c->d_cms) return FA_ERR_INVALID;
    if (!c->d_cms_global) {
        FA_CUDA(c, cudaMalloc(&c->d_cms_global, c->cms_words * 8));
        FA_CUDA(c, cudaMemsetAsync(c->d_cms_global, 0, c->cms_words * 8, c->stream));
    }
    return FA_OK;
}

extern "C" int fa_cms_device(fa_ctx *c, int which, void **d_ptr, size_t *n_words)
{
    if (!c || !d_ptr || !c->d_cms) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    if (which == FA_CMS_GLOBAL) {
        int rc = ensure_cms_global(c);
        if (rc) return rc;
        *d_ptr = c->d_cms_global;
    } else {
        *d_ptr = c->d_cms;
    }
    if (n_words) *n_words = c->cms_words;
    return FA_OK;
}

struct HhLess {
    int kw;
    bool operator()(const fa_hh &a, const fa_hh &b) const
    {
        if (a.estimate != b.estimate) return a.estimate > b.estimate;
        for (int i = 0; i < kw; i++) {
            if (a.key[i] != b.key[i]) return a.key[i] < b.key[i];
        }
        return false;
    }
};

// ---- top-K selection on the device: radix sort of the (inverted) estimates, gather the head ----
__global__ void k_hh_sort_keys(const fa_hh *hh, uint32_t n, unsigned long long *keys, uint32_t *idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = ~hh[i].estimate;  // ascending sort of the complement = descending estimates
        idx[i] = i;
    }
}
__global__ void k_hh_gather(const fa_hh *hh, const uint32_t *idx, uint32_t m, fa_hh *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = hh[idx[i]];
}

extern "C" int fa_topk_local(fa_ctx *c, int which, size_t k, fa_hh *out, size_t *n)
{
    if (c && c->flush_pending) return FA_ERR_INVALID;  // fa_flush_end first
    if (!c || !n || !c->d_cms || !c->d_slots || (k && !out)) return FA_ERR_INVALID;
    int rc = fa_sync(c);
    if (rc) return rc;
    rc = merge_hot(c);
    if (rc) return rc;
    rc = read_counters(c);
    if (rc) return rc;
    const uint64_t groups = c->h_ts->n_groups;
    *n = 0;
    if (!groups || !k) return FA_OK;
    if (groups >= (1ull << 31)) return FA_ERR_INVALID;
    const unsigned long long *cms = c->d_cms;
    if (which == FA_CMS_GLOBAL) {
        rc = ensure_cms_global(c);
        if (rc) return rc;
        cms = c->d_cms_global;
    }
    // scratch: hh[groups] | keys_a | keys_b (u64) | idx_a | idx_b (u32) | head[m] | cub temp
    const uint32_t ng = (uint32_t)groups;
    const size_t kk = std::min<size_t>(k, groups);
    const uint32_t m = (uint32_t)std::min<uint64_t>(groups, (uint64_t)kk + 4096);  // head + room for ties at the cut
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)ng, 0, 64, c->stream);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t hh_b = al(groups * sizeof(fa_hh)), k_b = al(groups * 8), i_b = al(groups * 4), head_b = al((size_t)m * sizeof(fa_hh));
    rc = ensure_scratch(c, hh_b + 2 * k_b + 2 * i_b + head_b + al(cub_bytes));
    if (rc) return rc;
    uint8_t *base = (uint8_t *)c->d_scratch;
    fa_hh *d_hh = (fa_hh *)base;
    unsigned long long *keys_a = (unsigned long long *)(base + hh_b), *keys_b = (unsigned long long *)(base + hh_b + k_b);
    uint32_t *idx_a = (uint32_t *)(base + hh_b + 2 * k_b), *idx_b = (uint32_t *)(base + hh_b + 2 * k_b + i_b);
    fa_hh *d_head = (fa_hh *)(base + hh_b + 2 * k_b + 2 * i_b);
    void *cub_tmp = base + hh_b + 2 * k_b + 2 * i_b + head_b;
    FA_CUDA(c, cudaMemsetAsync(&c->d_ts->flush_rows, 0, 8, c->stream));
    cudaError_t e;
#define CALL_EST(K) launch_estimate<K>(c, cms, d_hh, groups)
    FA_DISPATCH_KW(c->kw, CALL_EST)
#undef CALL_EST
    FA_CUDA(c, e);
    const int g = (int)((ng + 255) / 256);
    k_hh_sort_keys<<<g, 256, 0, c->stream>>>(d_hh, ng, keys_a, idx_a);
    c->n_kernels++;
    FA_CUDA(c, cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, idx_a, idx_b, (int)ng, 0, 64, c->stream));
    k_hh_gather<<<(int)((m + 255) / 256), 256, 0, c->stream>>>(d_hh, idx_b, m, d_head);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    std::vector<fa_hh> head(m);
    FA_CUDA(c, cudaMemcpyAsync(head.data(), d_head, (size_t)m * sizeof(fa_hh), cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    if (m < groups && head[m - 1].estimate == head[kk - 1].estimate) {
        // more than 4096 candidates tie with the k-th estimate: order ALL of them on the host (exact, slow, pathological)
        std::vector<fa_hh> all(groups);
        FA_CUDA(c, cudaMemcpy(all.data(), d_hh, groups * sizeof(fa_hh), cudaMemcpyDeviceToHost));
        std::partial_sort(all.begin(), all.begin() + kk, all.end(), HhLess{c->kw});
        memcpy(out, all.data(), kk * sizeof(fa_hh));
        *n = kk;
        return FA_OK;
    }
    // the head holds every candidate that can make the cut; final order (estimate desc, key asc) on <= k+4096 rows
    std::sort(head.begin(), head.end(), HhLess{c->kw});
    memcpy(out, head.data(), kk * sizeof(fa_hh));
    *n = kk;
    return FA_OK;
}

extern "C" int fa_topk_merge(const fa_hh *lists, size_t n_total, int key_words, size_t k, fa_hh *out, size_t *n)
{
    if (!n || (n_total && !lists) || key_words < 1 || key_words > FA_MAX_KEY_WORDS) return FA_ERR_INVALID;
    std::vector<fa_hh> v(lists, lists + n_total);
    HhLess less{key_words};
    std::sort(v.begin(), v.end(), less);
    // the same key may come from several contexts (same global estimate): keep one
    size_t w = 0;
    for (size_t i = 0; i < v.size(); i++) {
        if (w && memcmp(v[w - 1].key, v[i].key, sizeof(uint32_t) * key_words) == 0 && v[w - 1].estimate == v[i].estimate) continue;
        v[w++] = v[i];
    }
    const size_t kk = std::min(k, w);
    if (kk && !out) return FA_ERR_INVALID;
    memcpy(out, v.data(), kk * sizeof(fa_hh));
    *n = kk;
    return FA_OK;
}

// ---- NCCL (dlopen'ed so hosts that never ask for a box-wide top-K need no NCCL) ----
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool nccl_load(std::string &err)
{
    if (g_nccl.lib) return true;
    // resolve everything into a local copy and publish it only when complete: a half-loaded table must never be
    // mistaken for a loaded one by the next call
    NcclApi api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        err = std::string("dlopen libnccl.so.2 failed: ") + dlerror();
        return false;
    }
#define SYM(f, name)                                              \
    *(void **)(&api.f) = dlsym(api.lib, name);                    \
    if (!api.f) {                                                 \
        err = std::string("missing NCCL symbol ") + name;         \
        dlclose(api.lib);                                         \
        return false;                                             \
    }
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl = api;
    return true;
}
}  // namespace

extern "C" int fa_topk(fa_ctx *const *ctxs, int n_ctx, size_t k, fa_hh *out, size_t *n)
{
    if (!ctxs || n_ctx < 1 || !n) return FA_ERR_INVALID;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i] || !ctxs[i]->d_cms || ctxs[i]->kw != ctxs[0]->kw || ctxs[i]->cms_words != ctxs[0]->cms_words) return FA_ERR_INVALID;
    }
    int which = FA_CMS_LOCAL;
    if (n_ctx > 1) {
        std::lock_guard<std::mutex> lk(g_nccl_mu);
        fa_ctx *c0 = ctxs[0];
        std::string err;
        if (!nccl_load(err)) {
            c0->last_error = err;
            return FA_ERR_NCCL;
        }
        std::vector<int> devs(n_ctx);
        for (int i = 0; i < n_ctx; i++) devs[i] = ctxs[i]->cfg.device;
        if (g_nccl.devs != devs) {
            for (ncclComm_t cm : g_nccl.comms)
                if (cm) g_nccl.CommDestroy(cm);
            g_nccl.devs.clear();
            g_nccl.comms.assign(n_ctx, nullptr);
            int r = g_nccl.CommInitAll(g_nccl.comms.data(), n_ctx, devs.data());
            if (r != 0) {
                c0->last_error = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r);
                for (ncclComm_t cm : g_nccl.comms)  // a partial init may have produced some communicators
                    if (cm) g_nccl.CommDestroy(cm);
                g_nccl.comms.clear();
                g_nccl.devs.clear();
                return FA_ERR_NCCL;
            }
            g_nccl.devs = devs;
        }
        for (int i = 0; i < n_ctx; i++) {
            int rc = fa_sync(ctxs[i]);
            if (rc) return rc;
            cudaSetDevice(ctxs[i]->cfg.device);
            rc = ensure_cms_global(ctxs[i]);
            if (rc) return rc;
        }
        // CMS is linear: the sum of the sketches is the sketch of the union of the partitions
        g_nccl.GroupStart();
        int r = 0;
        for (int i = 0; i < n_ctx && r == 0; i++) {
            cudaSetDevice(ctxs[i]->cfg.device);
            r = g_nccl.AllReduce(ctxs[i]->d_cms, ctxs[i]->d_cms_global, ctxs[i]->cms_words, /*ncclUint64*/ 5, /*ncclSum*/ 0,
                                 g_nccl.comms[i], ctxs[i]->stream);
        }
        int r2 = g_nccl.GroupEnd();
        if (r != 0 || r2 != 0) {
            c0->last_error = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r ? r : r2);
            return FA_ERR_NCCL;
        }
        which = FA_CMS_GLOBAL;
    }
    std::vector<fa_hh> lists;
    for (int i = 0; i < n_ctx; i++) {
        std::vector<fa_hh> part(k);
        size_t got = 0;
        int rc = fa_topk_local(ctxs[i], which, k, part.data(), &got);
        if (rc) return rc;
        lists.insert(lists.end(), part.begin(), part.begin() + got);
    }
    return fa_topk_merge(lists.data(), lists.size(), ctxs[0]->kw, k, out, n);
}

// ---------------------------------------------------------------------------------------------
// kernel-1 columns
// ---------------------------------------------------------------------------------------------

extern "C" int fa_columns(fa_ctx *c, fa_columns_view *v)
{
    if (!c || !v || !(c->cfg.flags & FA_CFG_COLUMNS)) return FA_ERR_INVALID;
    FA_DRAIN(c);
    const Columns &k = c->cols;
    v->n_records = c->cols_n;
    v->valid = k.valid;
    v->time_received = (const uint64_t *)k.time_received;
    v->time_flow_start = (const uint64_t *)k.time_flow_start;
    v->sampling_rate = (const uint64_t *)k.sampling_rate;
    v->bytes = (const uint64_t *)k.bytes;
    v->packets = (const uint64_t *)k.packets;
    v->type = k.type;
    v->sequence_num = k.sequence_num;
    v->src_as = k.src_as;
    v->dst_as = k.dst_as;
    v->etype = k.etype;
    v->proto = k.proto;
    v->src_port = k.src_port;
    v->dst_port = k.dst_port;
    v->src_addr = (const uint8_t *)k.src_addr;
    v->dst_addr = (const uint8_t *)k.dst_addr;
    v->sampler_addr = (const uint8_t *)k.sampler_addr;
    v->src_addr_len = k.src_addr_len;
    v->dst_addr_len = k.dst_addr_len;
    v->sampler_addr_len = k.sampler_addr_len;
    return FA_OK;
}

extern "C" int fa_columns_read(fa_ctx *c, const char *col, void *out, size_t cap_bytes)
{
    if (!c || !col || !out || !(c->cfg.flags & FA_CFG_COLUMNS)) return FA_ERR_INVALID;
    FA_DRAIN(c);
    const Columns &k = c->cols;
    struct Ent { const char *name; const void *ptr; size_t elem; };
    const Ent ents[] = {
        {"valid", k.valid, 1}, {"time_received", k.time_received, 8}, {"time_flow_start", k.time_flow_start, 8},
        {"sampling_rate", k.sampling_rate, 8}, {"bytes", k.bytes, 8}, {"packets", k.packets, 8}, {"type", k.type, 4},
        {"sequence_num", k.sequence_num, 4}, {"src_as", k.src_as, 4}, {"dst_as", k.dst_as, 4}, {"etype", k.etype, 4},
        {"proto", k.proto, 4}, {"src_port", k.src_port, 4}, {"dst_port", k.dst_port, 4}, {"src_addr", k.src_addr, 16},
        {"dst_addr", k.dst_addr, 16}, {"sampler_addr", k.sampler_addr, 16}, {"src_addr_len", k.src_addr_len, 1},
        {"dst_addr_len", k.dst_addr_len, 1}, {"sampler_addr_len", k.sampler_addr_len, 1}};
    for (const Ent &e : ents) {
        if (strcmp(e.name, col) == 0) {
            const size_t bytes = e.elem * c->cols_n;
            if (cap_bytes < bytes) return FA_ERR_CAPACITY;
            int rc = fa_sync(c);
            if (rc) return rc;
            FA_CUDA(c, cudaMemcpy(out, e.ptr, bytes, cudaMemcpyDeviceToHost));
            return FA_OK;
        }
    }
    return FA_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------
// timing
// ---------------------------------------------------------------------------------------------

extern "C" int fa_timer_start(fa_ctx *c)
{
    if (!c) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    FA_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
    return FA_OK;
}

extern "C" int fa_timer_stop(fa_ctx *c, float *ms)
{
    if (!c || !ms) return FA_ERR_INVALID;
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    FA_DRAIN(c);
    FA_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    FA_CUDA(c, cudaEventSynchronize(c->ev_t1));
    FA_CUDA(c, cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return FA_OK;
}

// ---------------------------------------------------------------------------------------------
// mocker (mocker/mocker.go:57-102)
// ---------------------------------------------------------------------------------------------

extern "C" int fa_mocker_host(const fa_mocker_config *cfg, uint64_t first, uint32_t n, uint8_t *buf, size_t cap, uint32_t *offsets,
                              size_t *bytes)
{
    if (!cfg || !bytes) return FA_ERR_INVALID;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += fa_mocker_record_len(*cfg, first + i);
    *bytes = (size_t)total;
    if (total > 0xFFFFFFF0ull) return FA_ERR_INVALID;
    if (total > cap || !buf || !offsets) return FA_ERR_CAPACITY;
    uint8_t *p = buf;
    for (uint32_t i = 0; i < n; i++) {
        offsets[i] = (uint32_t)(p - buf);
        p = fa_mocker_record_put(*cfg, first + i, p);
    }
    offsets[n] = (uint32_t)(p - buf);
    return FA_OK;
}

__global__ void k_mocker_len(fa_mocker_config cfg, unsigned long long first, uint32_t n, uint32_t *offsets)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) offsets[0] = 0;
    if (i < n) offsets[i + 1] = fa_mocker_record_len(cfg, first + i);
}

__global__ void k_mocker_put(fa_mocker_config cfg, unsigned long long first, uint32_t n, const uint32_t *offsets, uint8_t *buf)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fa_mocker_record_put(cfg, first + i, buf + offsets[i]);
}

extern "C" int fa_mocker_device(fa_ctx *c, const fa_mocker_config *cfg, uint64_t first, uint32_t n, uint8_t *d_buf, size_t cap,
                                uint32_t *d_offsets, size_t *bytes)
{
    if (!c || !cfg || !d_buf || !d_offsets || !bytes) return FA_ERR_INVALID;
    if (n > (1u << 25)) return FA_ERR_INVALID;  // keeps the u32 running sum below 2^32
    FA_CUDA(c, cudaSetDevice(c->cfg.device));
    *bytes = 0;
    if (n == 0) return FA_OK;
    const int grid = (int)((n + 255) / 256);
    k_mocker_len<<<grid, 256, 0, c->stream>>>(*cfg, first, n, d_offsets);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    size_t tmp = 0;
    FA_CUDA(c, cub::DeviceScan::InclusiveSum(nullptr, tmp, d_offsets + 1, d_offsets + 1, (int)n, c->stream));
    int rc = ensure_scratch(c, tmp);
    if (rc) return rc;
    FA_CUDA(c, cub::DeviceScan::InclusiveSum(c->d_scratch, tmp, d_offsets + 1, d_offsets + 1, (int)n, c->stream));
    uint32_t total = 0;
    FA_CUDA(c, cudaMemcpyAsync(&total, d_offsets + n, 4, cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    *bytes = total;
    if (total > cap) return FA_ERR_CAPACITY;
    k_mocker_put<<<grid, 256, 0, c->stream>>>(*cfg, first, n, d_offsets, d_buf);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    return FA_OK;
}
