/*
 * flow_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the one hot path of cloudflare/flow-pipeline:
 *   Kafka value bytes -> proto3 decode of flowprotob.FlowMessage
 *                        (inserter/inserter.go:124, proto.Unmarshal)
 *                     -> row extraction (inserter/inserter.go:129-157)
 *                     -> the flows_5m roll-up (compose/clickhouse/create.sh:92-110)
 * plus the count-min sketch / top-K the north star asks for (no reference
 * counterpart; the semantic it approximates is viz-ch.json:233).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may link or call this.  The product (libflowagg.so)
 * never does; it fails loudly without its CUDA kernels.
 *
 * PARITY PINNING.  The reference holds no tests, fixtures or golden vectors
 * for this path and its decode arithmetic lives in an
 * un-vendored dependency (github.com/golang/protobuf v1.4.3 ->
 * google.golang.org/protobuf v1.26.0-rc.1, go.mod:7 / go.sum:47,171), and no Go
 * toolchain exists here.  The oracle is therefore pinned against the
 * reference's embedded FileDescriptorProto (pb-ext/flow.pb.go:650-714) driven
 * through Python protobuf (upb): tests/golden/ holds the vectors and the
 * script that made them.  Where upb and protobuf-go are known to differ
 * (see fo_decode) the oracle follows protobuf-go and the vectors say so.
 * The Clickhouse roll-up cannot be executed here at all: that half is
 * "parity unpinned" beyond the README's sample rows (README.md:155-183).
 */
#ifndef FLOW_ORACLE_H
#define FLOW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- decode ------------------------------------------------------------ */

/* error classes, all mean "proto.Unmarshal returned an error" */
enum {
    FO_OK = 0,
    FO_ERR_TRUNCATED = -1,  /* varint / fixed / bytes runs past the end       */
    FO_ERR_OVERFLOW = -2,   /* varint longer than 10 bytes or 10th byte >= 2  */
    FO_ERR_FIELDNUM = -3,   /* field number 0 or > 2^29-1 (2^31-1 in groups)  */
    FO_ERR_WIRETYPE = -4,   /* wire type 6 or 7                               */
    FO_ERR_ENDGROUP = -5,   /* unmatched / mismatched / missing end-group     */
    FO_ERR_UTF8 = -6,       /* proto3 string (fields 100,101) not valid UTF-8 */
    FO_ERR_DEPTH = -7,      /* groups nested deeper than FO_MAX_GROUP_DEPTH   */
    FO_ERR_FRAMING = -8     /* length prefix does not match the record span   */
};

#define FO_MAX_GROUP_DEPTH 32

/* The 16 fields the two reference consumers keep: the inserter's 14-column
 * row (inserter/inserter.go:142-157) and the 15 flows_raw columns
 * (compose/clickhouse/create.sh:36-59).  Address bytes are stored as
 * FixedString(16) does: first 16 bytes, zero right-padded (README.md:186-202);
 * the true length is kept beside them. */
typedef struct fo_flow {
    uint64_t time_received;   /* field 2  */
    uint64_t sampling_rate;   /* field 3  */
    uint64_t time_flow_start; /* field 38 */
    uint64_t bytes;           /* field 9  */
    uint64_t packets;         /* field 10 */
    uint32_t type;            /* field 1, enum -> int32, kept as its bit pattern */
    uint32_t sequence_num;    /* field 4  */
    uint32_t src_as;          /* field 14 */
    uint32_t dst_as;          /* field 15 */
    uint32_t etype;           /* field 30 */
    uint32_t proto;           /* field 20 */
    uint32_t src_port;        /* field 21 */
    uint32_t dst_port;        /* field 22 */
    uint32_t src_addr_len;    /* true length of field 6  */
    uint32_t dst_addr_len;    /* true length of field 7  */
    uint32_t sampler_addr_len;/* true length of field 11 */
    uint32_t src_addr_off;    /* offset of field 6 payload inside the message  */
    uint32_t dst_addr_off;
    uint8_t src_addr[16];
    uint8_t dst_addr[16];
    uint8_t sampler_addr[16];
} fo_flow;

/* proto.Unmarshal(msg, &fmsg) for the kept fields.  Returns FO_OK or FO_ERR_*.
 * On error *out is unspecified (the inserter skips the row, inserter.go:125). */
int fo_decode(const uint8_t *msg, size_t len, fo_flow *out);

/* One record of a batch: span = [begin,end) of buf.  framed != 0: the span is
 * varint(len) || message (mocker/mocker.go:98-101); framed == 0: bare message
 * (mocker/mocker.go:96-97). */
int fo_decode_record(const uint8_t *buf, size_t begin, size_t end, int framed, fo_flow *out);

/* Walk a length-delimited stream; writes record starts to offsets[0..n] (n+1
 * entries, last = end of the last complete record).  Returns the record count,
 * or -1 - (records found) if the stream ends inside a record. */
long fo_frame_walk(const uint8_t *buf, size_t len, uint32_t *offsets, size_t cap);

/* net.IP(b).String() with the inserter's "<nil>" -> "0.0.0.0" patch
 * (inserter/inserter.go:131-140).  out must hold 64 + 2*len bytes. */
void fo_ip_string(const uint8_t *addr, size_t len, char *out);

/* ---- keys, hash -------------------------------------------------------- */

enum {
    FO_KEY_FLOWS5M = 0, /* Timeslot, SrcAS, DstAS, EType   (create.sh:105-110) */
    FO_KEY_ASPAIR = 1,  /* SrcAS, DstAS                                        */
    FO_KEY_SRCADDR = 2, /* 4 big-endian words of SrcAddr   (viz-ch.json:233)   */
    FO_KEY_DSTADDR = 3, /*                                  (viz-ch.json:479)   */
    FO_KEY_5TUPLE = 4,  /* SrcAddr, DstAddr, SrcPort, DstPort, Proto           */
    FO_KEY_SRCPORT = 5, /* (viz-ch.json:358) */
    FO_KEY_DSTPORT = 6, /* (viz-ch.json:604) */
    FO_KEY_MODES = 7
};
#define FO_MAX_KEY_WORDS 12

int fo_key_words(int key_mode);
/* returns 0 if the flow cannot form this key (address longer than 16 bytes:
 * FixedString(16) would reject it), else 1 */
int fo_make_key(int key_mode, const fo_flow *f, uint32_t *key);
uint64_t fo_hash64(const uint32_t *key, int n_words);

/* ---- roll-up (flows_5m_view + fully merged SummingMergeTree) ------------ */

typedef struct fo_row {
    uint32_t key[FO_MAX_KEY_WORDS];
    uint64_t bytes, packets, count;
} fo_row;

typedef struct fo_agg fo_agg;
fo_agg *fo_agg_new(int key_mode, int scale_by_sampling_rate);
void fo_agg_free(fo_agg *a);
void fo_agg_add(fo_agg *a, const fo_flow *f);
void fo_agg_add_row(fo_agg *a, const fo_row *r); /* merge a partial aggregate */
size_t fo_agg_size(const fo_agg *a);
/* rows in canonical (ORDER BY key, create.sh:90) order */
size_t fo_agg_rows(const fo_agg *a, fo_row *rows, size_t cap);

/* ---- count-min sketch + top-K ------------------------------------------ */

typedef struct fo_hh {
    uint32_t key[FO_MAX_KEY_WORDS];
    uint64_t estimate;
} fo_hh;

/* idx_j = (lo32(h) + j * (hi32(h)|1)) & (2^wlog2 - 1), h = fo_hash64(key) */
void fo_cms_add(uint64_t *cms, int depth, int wlog2, const uint32_t *key, int n_words, uint64_t weight);
uint64_t fo_cms_estimate(const uint64_t *cms, int depth, int wlog2, const uint32_t *key, int n_words);
/* top-k of candidate rows by (estimate desc, key asc) */
size_t fo_topk(const uint64_t *cms, int depth, int wlog2, int n_words, const fo_row *cands,
               size_t n_cands, size_t k, fo_hh *out);

/* ---- whole-batch driver (CPU baseline) ---------------------------------- */

typedef struct fo_batch_result {
    uint64_t n_records, n_bad, n_nokey;
    double seconds; /* wall time of the decode+aggregate loop only */
} fo_batch_result;

/* Decode records [0,n) of buf (spans from offsets) and aggregate into `a`
 * (may be NULL), into cms (may be NULL).  n_threads > 1 shards records
 * contiguously over pthreads with per-thread tables merged at the end
 * (the "partition" sharding of inserter.go:176). */
int fo_run_batch(const uint8_t *buf, const uint32_t *offsets, size_t n, int framed, fo_agg *a,
                 uint64_t *cms, int depth, int wlog2, int n_threads, fo_batch_result *res);

/* Same over several slabs (each < 4 GiB) treated as one record stream. */
int fo_run_slabs(const uint8_t *const *bufs, const uint32_t *const *offsets, const size_t *ns, int n_slabs, int framed,
                 fo_agg *a, uint64_t *cms, int depth, int wlog2, int n_threads, fo_batch_result *res);

#ifdef __cplusplus
}
#endif
#endif
