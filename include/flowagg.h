/*
 * flowagg.h -- C ABI of libflowagg.so, the B200-native flow-aggregation stage.
 *
 * The reference (cloudflare/flow-pipeline) has no FFI, plugin or operator
 * interface; its only extension seam is the sarama.ConsumerGroupHandler that
 * inserter/inserter.go:167-196 implements.  This header is what a cgo shim
 * inside a rewritten ConsumeClaim binds (INTEGRATION.md shows the binding).
 * Each entry point names the reference code it replaces, paths relative to
 * the reference root.
 *
 * Conventions: plain C, plain pointers and sizes, no CUDA or torch types.
 * Every function returns FA_OK (0) or a negative fa_status and never aborts
 * the process (the reference log.Fatal's instead, inserter.go:104,249).
 * A fa_ctx is single-threaded -- one per ConsumeClaim goroutine / Kafka
 * partition / GPU stream (inserter.go:176); distinct contexts are independent
 * and may live on the same or on different GPUs.  There is NO CPU fallback:
 * if the CUDA device or the sm_100a kernels are unavailable fa_create fails
 * with FA_ERR_CUDA.
 */
#ifndef FLOWAGG_H
#define FLOWAGG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_ABI_VERSION 2

typedef enum fa_status {
    FA_OK = 0,
    FA_ERR_INVALID = -1,  /* bad argument / bad configuration                  */
    FA_ERR_CUDA = -2,     /* CUDA runtime error; see fa_last_error             */
    FA_ERR_NOMEM = -3,    /* host or device allocation failed                  */
    FA_ERR_CAPACITY = -4, /* caller's output array too small; *n holds the need */
    FA_ERR_TABLE_FULL = -5, /* group table overflowed; rows were dropped       */
    FA_ERR_NCCL = -6,     /* NCCL unavailable or failed                        */
    FA_ERR_FRAMING = -7   /* offsets == NULL and the stream ends inside a record */
} fa_status;

/* ---- group-by key (the GROUP BY list) ----------------------------------- */
typedef enum fa_key_mode {
    /* (Date, Timeslot, SrcAS, DstAS, EType): compose/clickhouse/create.sh:92-110.
     * key words: Timeslot (seconds, multiple of 300), SrcAS, DstAS, EType;
     * Date = Timeslot / 86400 is derived (fa_row_date). */
    FA_KEY_FLOWS5M = 0,
    FA_KEY_ASPAIR = 1,  /* (SrcAS, DstAS) -- BASELINE.json configs[1]            */
    FA_KEY_SRCADDR = 2, /* SrcAddr as FixedString(16): viz-ch.json:233           */
    FA_KEY_DSTADDR = 3, /* viz-ch.json:479                                       */
    FA_KEY_5TUPLE = 4,  /* SrcAddr, DstAddr, SrcPort, DstPort, Proto (configs[4]) */
    FA_KEY_SRCPORT = 5, /* viz-ch.json:358                                       */
    FA_KEY_DSTPORT = 6, /* viz-ch.json:604                                       */
    FA_KEY_MODES = 7
} fa_key_mode;

#define FA_MAX_KEY_WORDS 12

/* fa_config.flags */
#define FA_CFG_CMS 0x1u          /* keep a count-min sketch of Bytes*SamplingRate per key */
#define FA_CFG_SCALE_SAMPLING 0x2u /* sum(Bytes*SamplingRate), sum(Packets*SamplingRate):
                                      the dashboards' weighting (viz-ch.json:74)          */
#define FA_CFG_COLUMNS 0x4u      /* also materialise the decoded columns (kernel 1 output) */
#define FA_CFG_NO_AGGREGATE 0x8u /* decode only (with FA_CFG_COLUMNS): no group table      */
#define FA_CFG_TOPK_ONLY 0x20u   /* with FA_CFG_CMS and an address key: heavy hitters ONLY (the dashboards' top-N,
                                    viz-ch.json:233,479).  The sketch sees every flow; a key enters the group table -- now a
                                    bounded CANDIDATE table -- only once its sketch estimate reaches 1/(64*topk_k) of the weight
                                    seen so far (at most 64*topk_k keys can truly weigh that much), and leaves again when the
                                    bar has outgrown it.  Memory no longer depends on the number of distinct keys; fa_topk* work
                                    as before; fa_flush then returns the candidates with their sums SINCE ADMISSION, not an
                                    exact roll-up.  table_capacity: at least 4 * 64 * topk_k slots (0 = that).              */
#define FA_CFG_CALLER_STREAM 0x10u /* run on fa_config.stream even when it is NULL (the legacy
                                      default stream) instead of a private stream           */

typedef struct fa_config {
    uint32_t abi_version;    /* FA_ABI_VERSION */
    int32_t device;          /* CUDA device ordinal */
    uint32_t key_mode;       /* fa_key_mode */
    uint32_t flags;          /* FA_CFG_* */
    uint64_t table_capacity; /* group-table slots, rounded up to 2^k; 0 = 2^17.
                                Keep load <= 0.5: configs[4] needs 2^28. */
    uint32_t cms_depth;      /* 0 = 4   (configs[2]) */
    uint32_t cms_width_log2; /* 0 = 20  (configs[2]) */
    uint64_t max_batch_bytes;   /* staging for host submits; 0 = 256 MiB */
    uint32_t max_batch_records; /* 0 = 4 Mi */
    uint32_t topk_k;         /* FA_CFG_TOPK_ONLY: the K the context will be asked for; 0 = 1000 */
    void *stream; /* cudaStream_t to run on; NULL = a private non-blocking stream.
                     Passing the caller's stream lets the caller order/time work
                     with its own events. */
} fa_config;

typedef struct fa_ctx fa_ctx;

/* fa_submit flags */
#define FA_FRAMED 0x1u   /* each record is varint(len) || FlowMessage -- the Clickhouse
                            path, mocker -proto.fixedlen (mocker/mocker.go:98-101).
                            Otherwise records are bare messages (mocker.go:96-97)
                            and offsets are mandatory. */

/* One output row == one row of the fully merged flows_5m table
 * (create.sh:70-90): key, sum(Bytes), sum(Packets), count(). */
typedef struct fa_row {
    uint32_t key[FA_MAX_KEY_WORDS]; /* words beyond the mode's width are 0 */
    uint64_t bytes;
    uint64_t packets;
    uint64_t count;
} fa_row;

typedef struct fa_hh { /* heavy hitter: key + sketch estimate */
    uint32_t key[FA_MAX_KEY_WORDS];
    uint64_t estimate;
} fa_hh;

typedef struct fa_stats {
    uint64_t n_records;  /* records submitted since create/reset (msgCount, inserter.go:116) */
    uint64_t n_bad;      /* undecodable, skipped: inserter.go:125-126 */
    uint64_t n_nokey;    /* decoded but an address exceeds FixedString(16) (create.sh:15-16) */
    uint64_t n_dropped;  /* rows lost to a full group table (0 unless FA_ERR_TABLE_FULL) */
    uint64_t n_groups;   /* occupied group-table slots */
    uint64_t n_submits;  /* kernel launches of the decode/aggregate kernel */
    uint64_t bytes_in;   /* input bytes consumed */
    uint64_t n_kernels;  /* launches of the library's own CUDA kernels (all kinds) */
    uint64_t gpu_busy_us; /* device time of the decode/aggregate kernels so far (CUDA events around
                             every launch; what a GPU-busy gauge divides by wall time)            */
} fa_stats;

/* Decoded columns of the LAST submit (FA_CFG_COLUMNS).  Device pointers, valid
 * until the next submit.  Column set = the inserter's row (inserter.go:142-157)
 * united with flows_raw (create.sh:36-59).  Addresses are 16 bytes per record,
 * zero right-padded like FixedString(16) (README.md:186-202). */
typedef struct fa_columns_view {
    uint64_t n_records;
    const uint8_t *valid;      /* 1 = decoded, 0 = skipped (bad record) */
    const uint64_t *time_received, *time_flow_start, *sampling_rate, *bytes, *packets;
    const uint32_t *type, *sequence_num, *src_as, *dst_as, *etype, *proto, *src_port, *dst_port;
    const uint8_t *src_addr, *dst_addr, *sampler_addr; /* n * 16 bytes */
    const uint8_t *src_addr_len, *dst_addr_len, *sampler_addr_len; /* true length, saturated at 255 */
} fa_columns_view;

/* ---- lifecycle ----------------------------------------------------------- */
/* Replaces main()'s state construction (inserter.go:204-210). */
int fa_create(const fa_config *cfg, fa_ctx **out);
void fa_destroy(fa_ctx *ctx);
const char *fa_strerror(int status);
const char *fa_last_error(const fa_ctx *ctx); /* detail of the last FA_ERR_CUDA/NCCL */

/* ---- ingest: replaces (*state).buffer, inserter.go:113-165 ---------------- */
/* Library-owned pinned host slabs the Go side memcpy's msg.Value into
 * (sarama owns msg.Value only until the next loop iteration, inserter.go:179).
 * Two slabs (slot 0/1) so one fills while the other is in flight; the call
 * blocks until the slab's previous host-to-device copy has finished.  Pointers
 * stay valid until fa_destroy.  Sizes are fixed at create (max_batch_bytes,
 * max_batch_records+1 offsets). */
int fa_host_buffer(fa_ctx *ctx, int slot, uint8_t **buf, size_t *cap_bytes, uint32_t **offsets,
                   size_t *cap_records);

/* Submit n_records records held in HOST memory (ideally the pinned slab above):
 * record i is buf[offsets[i] .. offsets[i+1]).  offsets may be NULL only with
 * FA_FRAMED (the library then finds the boundaries on the GPU); n_records is
 * then ignored and discovered.  Inputs larger than max_batch_bytes /
 * max_batch_records are cut into batches at record boundaries and pipelined
 * (copy of batch i+1 overlaps the kernel of batch i).  Asynchronous; the host
 * buffers may be reused after fa_sync, or -- for the library's slabs -- after
 * fa_host_buffer returns them again.
 * offsets == NULL: len <= max_batch_bytes, the stream must start and end on a
 * record boundary, and the call runs ONE BEHIND: it enqueues this batch's
 * host-to-device copy and then indexes and launches the batch the previous such
 * call staged (the index needs one host look at three device counters, which now
 * waits while the next copy is already running).  Whatever reads the context
 * (fa_sync, fa_flush*, fa_stats_get, fa_topk*, ...) first finishes the batch
 * still pending; an error of a deferred batch is reported by the call that
 * finishes it. */
int fa_submit(fa_ctx *ctx, const uint8_t *buf, size_t len, const uint32_t *offsets, uint32_t n_records,
              uint32_t flags);

/* Same, with buf/offsets already in DEVICE memory (16-byte aligned buf).  Used
 * when another GPU stage produced the bytes and by the benchmark's
 * HBM-resident leg. */
int fa_submit_device(fa_ctx *ctx, const uint8_t *d_buf, size_t len, const uint32_t *d_offsets,
                     uint32_t n_records, uint32_t flags);

int fa_sync(fa_ctx *ctx); /* wait for everything submitted so far */
int fa_stats_get(fa_ctx *ctx, fa_stats *out); /* implies fa_sync */

/* ---- emit: replaces (*state).flush, inserter.go:90-111 -------------------- */
#define FA_FLUSH_KEEP 0x1u     /* do not reset the table (peek) */
#define FA_FLUSH_UNSORTED 0x2u /* skip the ORDER BY (create.sh:90) sort */
/* Synchronous.  Writes up to cap rows in canonical key order and resets the
 * group table (not the sketch).  *n = rows available; FA_ERR_CAPACITY if
 * cap < *n (nothing is reset then). */
int fa_flush(fa_ctx *ctx, fa_row *rows, size_t cap, size_t *n, uint32_t flags);

/* The same flush in two halves, so that it never stalls the stream ((*state).flush holds the inserter's global mutex
 * while it talks to the database, inserter.go:90-111, and buffer() waits behind it; here the next submit does not):
 * fa_flush_begin folds the hot-key replicas, swaps the filled group table for a spare, empty one -- every later
 * fa_submit already aggregates into the spare -- and enqueues compaction, ORDER BY, the copies to pinned host memory and
 * the emptying of the old table on a side stream; it returns at once.  fa_flush_end waits for that work (the thread
 * sleeps on a blocking event, it does not spin) and copies the rows out; same results and errors as fa_flush (with
 * FA_ERR_CAPACITY the rows are kept: call fa_flush_end again with a larger array).  One flush in flight per context;
 * between the two halves only fa_submit*, fa_sync, fa_stats_get and fa_host_buffer may be called on it.  Group tables
 * of more than 2^22 slots, FA_FLUSH_KEEP and FA_FLUSH_UNSORTED are drained by fa_flush_end itself (no spare table). */
int fa_flush_begin(fa_ctx *ctx, uint32_t flags);
int fa_flush_end(fa_ctx *ctx, fa_row *rows, size_t cap, size_t *n);

/* The merge step of the SummingMergeTree behind flows_5m (compose/clickhouse/create.sh:88-90: rows with equal
 * ORDER BY keys are summed) for rows that are already aggregates -- an earlier window's or another context's
 * fa_flush output.  `rows` may be host or device memory (a peer GPU's included).  n_owners > 1 adds only the rows
 * whose key this context owns (fa_row_owner(...) == owner); n_owners <= 1 adds all.  Asynchronous on the
 * context's stream for device rows; host rows are copied before the call returns. */
int fa_merge_rows(fa_ctx *ctx, const fa_row *rows, size_t n, uint32_t owner, uint32_t n_owners);

/* owner[i] in [0, n_owners) of rows[i] (host memory) under `key_mode`: the hash partition fa_flush_box and a
 * one-process-per-GPU host use for the exchange step of an exact box-wide roll-up.  Pure host function. */
int fa_row_owner(int key_mode, const fa_row *rows, size_t n, uint32_t n_owners, uint32_t *owner);

/* Exact box-wide roll-up of n_ctx contexts (same key mode; normally one per GPU of the box = one per Kafka partition
 * group, inserter/inserter.go:176).  Kafka does not partition by group key, so every context holds partial sums of
 * the same keys: one hash-partitioned exchange (context j pulls the rows it owns out of every context's compacted
 * rows -- peer loads over NVLink when the devices allow it, a peer copy otherwise -- and sums them in its own
 * table), then every context emits its share.  rows: host array, canonical order unless FA_FLUSH_UNSORTED.
 * Every context is reset.  FA_ERR_CAPACITY: *n = rows needed; the contexts then hold the exchanged (disjoint)
 * roll-ups and the call can be repeated with a larger array.  FA_FLUSH_KEEP is not supported (FA_ERR_INVALID). */
int fa_flush_box(fa_ctx *const *ctxs, int n_ctx, fa_row *rows, size_t cap, size_t *n, uint32_t flags);
int fa_reset(fa_ctx *ctx); /* clear table, sketch and statistics */

/* Date column of a FLOWS5M row: toDate(TimeReceived), days since epoch. */
static inline uint32_t fa_row_date(const fa_row *r) { return r->key[0] / 86400u; }

/* ---- sketch / heavy hitters (north star; semantic of viz-ch.json:233) ----- */
/* Copy the d*w uint64 counters to host (row-major, row j at j<<wlog2). */
int fa_cms_read(fa_ctx *ctx, uint64_t *out, size_t cap_words);
/* Device pointers of the counters so a multi-process host can all-reduce them
 * with its own communicator (torch.distributed / NCCL).  which = FA_CMS_LOCAL:
 * this context's sketch (never overwritten by a collective, so repeated top-K
 * queries do not double count); FA_CMS_GLOBAL: the reduced copy -- the host
 * all-reduces LOCAL into GLOBAL (out of place) before a box-wide query. */
#define FA_CMS_LOCAL 0
#define FA_CMS_GLOBAL 1
int fa_cms_device(fa_ctx *ctx, int which, void **d_ptr, size_t *n_words);
/* Top-k keys of THIS context's group table by sketch estimate
 * (estimate desc, key asc), estimated from the LOCAL or the GLOBAL sketch.
 * Merge the per-context lists with fa_topk_merge for a box-wide answer. */
int fa_topk_local(fa_ctx *ctx, int which, size_t k, fa_hh *out, size_t *n);
/* Merge per-context lists (concatenated in `lists`, n_total entries, duplicates
 * allowed) into the global top-k.  Pure host arithmetic on <= n_ctx*k rows. */
int fa_topk_merge(const fa_hh *lists, size_t n_total, int key_words, size_t k, fa_hh *out, size_t *n);
/* Single-process, n_ctx contexts (one per GPU): ncclAllReduce(sum, uint64) of
 * the LOCAL sketches into the GLOBAL ones over NVLink when n_ctx > 1, then
 * fa_topk_local(GLOBAL) + merge. */
int fa_topk(fa_ctx *const *ctxs, int n_ctx, size_t k, fa_hh *out, size_t *n);

/* ---- kernel-1 output ------------------------------------------------------- */
int fa_columns(fa_ctx *ctx, fa_columns_view *view);
/* Convenience for tests and row sinks: copy one column of the last submit to
 * host.  col names: "valid","time_received","time_flow_start","sampling_rate",
 * "bytes","packets","type","sequence_num","src_as","dst_as","etype","proto",
 * "src_port","dst_port","src_addr","dst_addr","sampler_addr","src_addr_len",
 * "dst_addr_len","sampler_addr_len". */
int fa_columns_read(fa_ctx *ctx, const char *col, void *out, size_t cap_bytes);

/* ---- timing on the context's stream (CUDA events) -------------------------- */
int fa_timer_start(fa_ctx *ctx);
int fa_timer_stop(fa_ctx *ctx, float *ms); /* synchronises the stop event */

/* ---- synthetic input: mocker/mocker.go:57-102 ------------------------------ */
#define FA_ADDR_MOCKER 0 /* 2001:db8:0:1::XX, XX uniform byte (mocker.go:64-71)  */
#define FA_ADDR_ZIPF24 1 /* SrcAddr low 24 bits ~ Zipf(s~1.1) rank (configs[2])  */
#define FA_ADDR_UNIQUE 2 /* SrcAddr low 64 bits = record index (configs[4])       */

typedef struct fa_mocker_config {
    uint64_t seed;             /* default 1 */
    uint64_t t0;               /* TimeReceived of record 0; 1584912398 = README.md:155 */
    uint64_t flows_per_second; /* TimeReceived = t0 + index / fps; 0 = constant */
    uint32_t n_src_as;         /* SrcAS = 65000 + U[0,n); mocker: 3 (mocker.go:61,80) */
    uint32_t n_dst_as;
    uint32_t addr_mode;        /* FA_ADDR_* */
    uint32_t framed;           /* 1 = -proto.fixedlen=true (mocker.go:98-101) */
} fa_mocker_config;

/* Host generator: records [first, first+n) into buf; offsets gets n+1 entries.
 * *bytes = total size (also returned when buf is too small: FA_ERR_CAPACITY). */
int fa_mocker_host(const fa_mocker_config *cfg, uint64_t first, uint32_t n, uint8_t *buf, size_t cap,
                   uint32_t *offsets, size_t *bytes);
/* Device generator (same bytes for the same (cfg, index)); d_buf 16-byte aligned. */
int fa_mocker_device(fa_ctx *ctx, const fa_mocker_config *cfg, uint64_t first, uint32_t n, uint8_t *d_buf,
                     size_t cap, uint32_t *d_offsets, size_t *bytes);

/* Library facts for harnesses: compiled arch string ("sm_100a"), ABI version. */
const char *fa_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
