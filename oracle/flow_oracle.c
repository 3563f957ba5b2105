/*
 * flow_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See flow_oracle.h for scope and the parity-pinning statement.
 *
 * Everything here is scalar, portable C99 + pthreads.  Each function cites the
 * reference lines (relative to /root/reference) it restates.
 */
#define _GNU_SOURCE
#include "flow_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ======================================================================== */
/* wire primitives: google.golang.org/protobuf/encoding/protowire           */
/* ======================================================================== */

/* protowire.ConsumeVarint: <= 10 bytes, the 10th byte must be 0 or 1.
 * Returns bytes consumed (>0) or FO_ERR_*. */
static int consume_varint(const uint8_t *b, size_t len, uint64_t *v)
{
    uint64_t x = 0;
    for (int i = 0; i < 10; i++) {
        if ((size_t)i >= len) return FO_ERR_TRUNCATED;
        uint64_t y = b[i];
        if (i == 9) {
            if (y >= 2) return FO_ERR_OVERFLOW;
            x |= y << 63;
            *v = x;
            return 10;
        }
        x |= (y & 0x7f) << (7 * i);
        if (y < 0x80) {
            *v = x;
            return i + 1;
        }
    }
    return FO_ERR_OVERFLOW; /* unreachable */
}

/* unicode/utf8.Valid: rejects overlongs, surrogates, > U+10FFFF, truncation */
static int utf8_valid(const uint8_t *s, size_t n)
{
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) {
            i++;
            continue;
        }
        size_t need;
        uint8_t lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) {
            need = 1;
        } else if (c == 0xE0) {
            need = 2;
            lo = 0xA0;
        } else if (c >= 0xE1 && c <= 0xEC) {
            need = 2;
        } else if (c == 0xED) {
            need = 2;
            hi = 0x9F;
        } else if (c >= 0xEE && c <= 0xEF) {
            need = 2;
        } else if (c == 0xF0) {
            need = 3;
            lo = 0x90;
        } else if (c >= 0xF1 && c <= 0xF3) {
            need = 3;
        } else if (c == 0xF4) {
            need = 3;
            hi = 0x8F;
        } else {
            return 0;
        }
        if (i + need >= n) return 0; /* sequence runs past the end */
        if (s[i + 1] < lo || s[i + 1] > hi) return 0;
        for (size_t k = 2; k <= need; k++)
            if (s[i + k] < 0x80 || s[i + k] > 0xBF) return 0;
        i += need + 1;
    }
    return 1;
}

/* protowire.ConsumeFieldValue for a start-group: skip to the matching
 * end-group.  Inside groups ConsumeTag accepts field numbers 1..2^31-1.
 * Nesting is bounded at FO_MAX_GROUP_DEPTH (protobuf-go: 10000 or unbounded,
 * depending on version -- documented deviation). */
static long skip_group(const uint8_t *b, size_t len, uint64_t start_num)
{
    uint64_t stack[FO_MAX_GROUP_DEPTH];
    int depth = 0;
    size_t p = 0;
    stack[depth++] = start_num;
    while (depth > 0) {
        uint64_t tag;
        int n = consume_varint(b + p, len - p, &tag);
        if (n < 0) return n;
        p += (size_t)n;
        uint64_t num = tag >> 3;
        unsigned wt = (unsigned)(tag & 7);
        if (num < 1 || num > 0x7fffffffull) return FO_ERR_FIELDNUM;
        uint64_t v;
        switch (wt) {
        case 0:
            n = consume_varint(b + p, len - p, &v);
            if (n < 0) return n;
            p += (size_t)n;
            break;
        case 1:
            if (len - p < 8) return FO_ERR_TRUNCATED;
            p += 8;
            break;
        case 2:
            n = consume_varint(b + p, len - p, &v);
            if (n < 0) return n;
            p += (size_t)n;
            if (v > (uint64_t)(len - p)) return FO_ERR_TRUNCATED;
            p += (size_t)v;
            break;
        case 3:
            if (depth >= FO_MAX_GROUP_DEPTH) return FO_ERR_DEPTH;
            stack[depth++] = num;
            break;
        case 4:
            if (stack[depth - 1] != num) return FO_ERR_ENDGROUP;
            depth--;
            break;
        case 5:
            if (len - p < 4) return FO_ERR_TRUNCATED;
            p += 4;
            break;
        default:
            return FO_ERR_WIRETYPE;
        }
    }
    return (long)p;
}

static void store_addr(uint8_t dst[16], uint32_t *dst_len, const uint8_t *src, uint64_t n)
{
    /* bytes fields are REPLACED, not appended (consumeBytes) */
    memset(dst, 0, 16);
    memcpy(dst, src, n < 16 ? (size_t)n : 16);
    *dst_len = n > 0xffffffffull ? 0xffffffffu : (uint32_t)n;
}

/*
 * proto.Unmarshal for flowprotob.FlowMessage (call site inserter/inserter.go:124;
 * field table pb-ext/flow.pb.go:58-143).  Restates
 * google.golang.org/protobuf/internal/impl.(*MessageInfo).unmarshalPointer:
 *   - tag = one varint (<=10 bytes); field number must be 1..2^29-1;
 *   - end-group at top level is always an error (groupTag == 0);
 *   - a known field whose wire type is not the declared one is "unknown" and is
 *     skipped by wire type (ConsumeFieldValue); same for unknown numbers;
 *   - scalars: last value wins; uint32/enum keep the low 32 bits; bool = v != 0;
 *   - bytes: replaced; proto3 strings (100 SrcCountry, 101 DstCountry) must be
 *     valid UTF-8;
 *   - wire types 6,7 and truncation are errors.
 * Known differences from upb (Python protobuf), which the golden vectors mark:
 *   upb rejects tags longer than 5 bytes (protobuf-go accepts any <=10-byte
 *   varint whose value passes the range check) and upb accepts a 10-byte varint
 *   whose last byte is >= 2 (protobuf-go: overflow error).
 */
int fo_decode(const uint8_t *msg, size_t len, fo_flow *out)
{
    memset(out, 0, sizeof(*out)); /* m.Reset() */
    size_t p = 0;
    while (p < len) {
        uint64_t tag;
        int n = consume_varint(msg + p, len - p, &tag);
        if (n < 0) return n;
        p += (size_t)n;
        uint64_t num = tag >> 3;
        unsigned wt = (unsigned)(tag & 7);
        if (num < 1 || num > 0x1fffffffull) return FO_ERR_FIELDNUM;
        uint64_t v;
        switch (wt) {
        case 0:
            n = consume_varint(msg + p, len - p, &v);
            if (n < 0) return n;
            p += (size_t)n;
            switch (num) {
            case 1: out->type = (uint32_t)v; break;
            case 2: out->time_received = v; break;
            case 3: out->sampling_rate = v; break;
            case 4: out->sequence_num = (uint32_t)v; break;
            case 9: out->bytes = v; break;
            case 10: out->packets = v; break;
            case 14: out->src_as = (uint32_t)v; break;
            case 15: out->dst_as = (uint32_t)v; break;
            case 20: out->proto = (uint32_t)v; break;
            case 21: out->src_port = (uint32_t)v; break;
            case 22: out->dst_port = (uint32_t)v; break;
            case 30: out->etype = (uint32_t)v; break;
            case 38: out->time_flow_start = v; break;
            default: break;
            }
            break;
        case 1:
            if (len - p < 8) return FO_ERR_TRUNCATED;
            p += 8;
            break;
        case 2:
            n = consume_varint(msg + p, len - p, &v);
            if (n < 0) return n;
            p += (size_t)n;
            if (v > (uint64_t)(len - p)) return FO_ERR_TRUNCATED;
            switch (num) {
            case 6:
                store_addr(out->src_addr, &out->src_addr_len, msg + p, v);
                out->src_addr_off = (uint32_t)p;
                break;
            case 7:
                store_addr(out->dst_addr, &out->dst_addr_len, msg + p, v);
                out->dst_addr_off = (uint32_t)p;
                break;
            case 11: store_addr(out->sampler_addr, &out->sampler_addr_len, msg + p, v); break;
            case 100:
            case 101:
                if (!utf8_valid(msg + p, (size_t)v)) return FO_ERR_UTF8;
                break;
            default: break;
            }
            p += (size_t)v;
            break;
        case 3: {
            long g = skip_group(msg + p, len - p, num);
            if (g < 0) return (int)g;
            p += (size_t)g;
            break;
        }
        case 4:
            return FO_ERR_ENDGROUP;
        case 5:
            if (len - p < 4) return FO_ERR_TRUNCATED;
            p += 4;
            break;
        default:
            return FO_ERR_WIRETYPE;
        }
    }
    return FO_OK;
}

/* framed: proto.Buffer.EncodeMessage framing (mocker/mocker.go:98-101) read back
 * the way Clickhouse's Protobuf format / proto.Buffer.DecodeMessage do: one
 * varint length, then exactly that many bytes, which must fill the span. */
int fo_decode_record(const uint8_t *buf, size_t begin, size_t end, int framed, fo_flow *out)
{
    if (end < begin) return FO_ERR_FRAMING;
    if (!framed) return fo_decode(buf + begin, end - begin, out);
    uint64_t mlen;
    int n = consume_varint(buf + begin, end - begin, &mlen);
    if (n < 0) return FO_ERR_FRAMING;
    if (mlen != (uint64_t)(end - begin - (size_t)n)) return FO_ERR_FRAMING;
    return fo_decode(buf + begin + n, (size_t)mlen, out);
}

long fo_frame_walk(const uint8_t *buf, size_t len, uint32_t *offsets, size_t cap)
{
    size_t p = 0;
    long n = 0;
    while (p < len) {
        uint64_t mlen;
        int k = consume_varint(buf + p, len - p, &mlen);
        if (k < 0 || mlen > (uint64_t)(len - p - (size_t)k)) {
            if ((size_t)n < cap) offsets[n] = (uint32_t)p;
            return -1 - n;
        }
        if ((size_t)n < cap) offsets[n] = (uint32_t)p;
        n++;
        p += (size_t)k + (size_t)mlen;
    }
    if ((size_t)n < cap) offsets[n] = (uint32_t)p;
    return n;
}

/* ======================================================================== */
/* net.IP.String  (Go net package; call site inserter/inserter.go:131-140)   */
/* ======================================================================== */

static char *append_hex16(char *o, unsigned v)
{
    static const char hx[] = "0123456789abcdef";
    if (v == 0) {
        *o++ = '0';
        return o;
    }
    int started = 0;
    for (int sh = 12; sh >= 0; sh -= 4) {
        unsigned d = (v >> sh) & 0xf;
        if (d || started) {
            *o++ = hx[d];
            started = 1;
        }
    }
    return o;
}

void fo_ip_string(const uint8_t *p, size_t len, char *out)
{
    static const char hx[] = "0123456789abcdef";
    if (len == 0) { /* "<nil>" -> "0.0.0.0", inserter.go:135-140 */
        strcpy(out, "0.0.0.0");
        return;
    }
    const uint8_t *p4 = NULL;
    if (len == 4) {
        p4 = p;
    } else if (len == 16) {
        int z = 1;
        for (int i = 0; i < 10; i++)
            if (p[i]) z = 0;
        if (z && p[10] == 0xff && p[11] == 0xff) p4 = p + 12;
    }
    if (p4) {
        sprintf(out, "%u.%u.%u.%u", p4[0], p4[1], p4[2], p4[3]);
        return;
    }
    if (len != 16) { /* "?" + hexString(ip) */
        char *o = out;
        *o++ = '?';
        for (size_t i = 0; i < len; i++) {
            *o++ = hx[p[i] >> 4];
            *o++ = hx[p[i] & 15];
        }
        *o = 0;
        return;
    }
    int e0 = -1, e1 = -1;
    for (int i = 0; i < 16; i += 2) {
        int j = i;
        while (j < 16 && p[j] == 0 && p[j + 1] == 0) j += 2;
        if (j > i && j - i > e1 - e0) {
            e0 = i;
            e1 = j;
            i = j;
        }
    }
    if (e1 - e0 <= 2) {
        e0 = -1;
        e1 = -1;
    }
    char *o = out;
    for (int i = 0; i < 16; i += 2) {
        if (i == e0) {
            *o++ = ':';
            *o++ = ':';
            i = e1;
            if (i >= 16) break;
        } else if (i > 0) {
            *o++ = ':';
        }
        o = append_hex16(o, ((unsigned)p[i] << 8) | p[i + 1]);
    }
    *o = 0;
}

/* ======================================================================== */
/* keys + hash                                                               */
/* ======================================================================== */

static const int k_key_words[FO_KEY_MODES] = {4, 2, 4, 4, 11, 1, 1};

int fo_key_words(int key_mode)
{
    return (key_mode >= 0 && key_mode < FO_KEY_MODES) ? k_key_words[key_mode] : -1;
}

static uint32_t be32(const uint8_t *b)
{
    return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
}

/*
 * FLOWS5M: compose/clickhouse/create.sh:94-110.  Timeslot =
 * toStartOfFiveMinute(TimeReceived) = t - t % 300 on the DateTime (UInt32)
 * column flows_raw.TimeReceived (create.sh:39), i.e. on the low 32 bits of the
 * UInt64 wire value; Date = toDate(TimeReceived) = Timeslot / 86400 days (UTC
 * server) and is therefore derived at output time, not stored in the key.
 * Word order makes numeric lexicographic order on the words equal the
 * table's ORDER BY (create.sh:90).  Addresses are FixedString(16): compared
 * bytewise, hence big-endian words.
 */
int fo_make_key(int key_mode, const fo_flow *f, uint32_t *key)
{
    switch (key_mode) {
    case FO_KEY_FLOWS5M: {
        uint32_t t = (uint32_t)f->time_received;
        key[0] = t - t % 300u;
        key[1] = f->src_as;
        key[2] = f->dst_as;
        key[3] = f->etype;
        return 1;
    }
    case FO_KEY_ASPAIR:
        key[0] = f->src_as;
        key[1] = f->dst_as;
        return 1;
    case FO_KEY_SRCADDR:
        if (f->src_addr_len > 16) return 0;
        for (int i = 0; i < 4; i++) key[i] = be32(f->src_addr + 4 * i);
        return 1;
    case FO_KEY_DSTADDR:
        if (f->dst_addr_len > 16) return 0;
        for (int i = 0; i < 4; i++) key[i] = be32(f->dst_addr + 4 * i);
        return 1;
    case FO_KEY_5TUPLE:
        if (f->src_addr_len > 16 || f->dst_addr_len > 16) return 0;
        for (int i = 0; i < 4; i++) key[i] = be32(f->src_addr + 4 * i);
        for (int i = 0; i < 4; i++) key[4 + i] = be32(f->dst_addr + 4 * i);
        key[8] = f->src_port;
        key[9] = f->dst_port;
        key[10] = f->proto;
        return 1;
    case FO_KEY_SRCPORT:
        key[0] = f->src_port;
        return 1;
    case FO_KEY_DSTPORT:
        key[0] = f->dst_port;
        return 1;
    default:
        return 0;
    }
}

/* The sketch/table hash.  Ours to define (the reference has no sketch); the
 * CUDA side restates the same arithmetic.  Words are folded in pairs. */
uint64_t fo_hash64(const uint32_t *key, int n_words)
{
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int i = 0; i < n_words; i += 2) {
        uint64_t w = key[i];
        if (i + 1 < n_words) w |= (uint64_t)key[i + 1] << 32;
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

/* ======================================================================== */
/* roll-up                                                                   */
/* ======================================================================== */

struct fo_agg {
    int key_mode, kw, scale;
    size_t cap, size; /* cap is a power of two */
    fo_row *slots;
    uint8_t *used;
};

fo_agg *fo_agg_new(int key_mode, int scale)
{
    if (fo_key_words(key_mode) < 0) return NULL;
    fo_agg *a = (fo_agg *)calloc(1, sizeof(*a));
    a->key_mode = key_mode;
    a->kw = fo_key_words(key_mode);
    a->scale = scale;
    a->cap = 1024;
    a->slots = (fo_row *)calloc(a->cap, sizeof(fo_row));
    a->used = (uint8_t *)calloc(a->cap, 1);
    return a;
}

void fo_agg_free(fo_agg *a)
{
    if (!a) return;
    free(a->slots);
    free(a->used);
    free(a);
}

static void agg_insert(fo_agg *a, const uint32_t *key, uint64_t b, uint64_t p, uint64_t c);

static void agg_grow(fo_agg *a)
{
    fo_row *old = a->slots;
    uint8_t *oldu = a->used;
    size_t oldcap = a->cap;
    a->cap *= 2;
    a->size = 0;
    a->slots = (fo_row *)calloc(a->cap, sizeof(fo_row));
    a->used = (uint8_t *)calloc(a->cap, 1);
    for (size_t i = 0; i < oldcap; i++)
        if (oldu[i]) agg_insert(a, old[i].key, old[i].bytes, old[i].packets, old[i].count);
    free(old);
    free(oldu);
}

static void agg_insert(fo_agg *a, const uint32_t *key, uint64_t b, uint64_t p, uint64_t c)
{
    if ((a->size + 1) * 2 > a->cap) agg_grow(a);
    size_t mask = a->cap - 1;
    size_t i = (size_t)(fo_hash64(key, a->kw) >> 32) & mask;
    for (;;) {
        if (!a->used[i]) {
            a->used[i] = 1;
            memset(&a->slots[i], 0, sizeof(fo_row));
            memcpy(a->slots[i].key, key, (size_t)a->kw * 4);
            a->size++;
            break;
        }
        if (memcmp(a->slots[i].key, key, (size_t)a->kw * 4) == 0) break;
        i = (i + 1) & mask;
    }
    /* sum()/count() over UInt64 wrap modulo 2^64 (create.sh:105-107) */
    a->slots[i].bytes += b;
    a->slots[i].packets += p;
    a->slots[i].count += c;
}

/* flows_5m_view: sum(Bytes), sum(Packets), count() GROUP BY key
 * (create.sh:105-110).  scale: the dashboards' sum(Bytes*SamplingRate)
 * (viz-ch.json:74,233). */
void fo_agg_add(fo_agg *a, const fo_flow *f)
{
    uint32_t key[FO_MAX_KEY_WORDS];
    if (!fo_make_key(a->key_mode, f, key)) return;
    uint64_t b = f->bytes, p = f->packets;
    if (a->scale) {
        b *= f->sampling_rate;
        p *= f->sampling_rate;
    }
    agg_insert(a, key, b, p, 1);
}

void fo_agg_add_row(fo_agg *a, const fo_row *r)
{
    agg_insert(a, r->key, r->bytes, r->packets, r->count);
}

size_t fo_agg_size(const fo_agg *a) { return a->size; }

static int g_cmp_kw;
static int row_cmp(const void *x, const void *y)
{
    const fo_row *a = (const fo_row *)x, *b = (const fo_row *)y;
    for (int i = 0; i < g_cmp_kw; i++) {
        if (a->key[i] < b->key[i]) return -1;
        if (a->key[i] > b->key[i]) return 1;
    }
    return 0;
}

/* SummingMergeTree fully merged == GROUP BY the ORDER BY key (create.sh:88-90);
 * canonical row order is that key ascending. */
size_t fo_agg_rows(const fo_agg *a, fo_row *rows, size_t cap)
{
    size_t n = 0;
    for (size_t i = 0; i < a->cap; i++)
        if (a->used[i]) {
            if (n < cap) rows[n] = a->slots[i];
            n++;
        }
    if (n <= cap) {
        g_cmp_kw = a->kw;
        qsort(rows, n, sizeof(fo_row), row_cmp);
    }
    return n;
}

/* ======================================================================== */
/* count-min sketch + top-K                                                  */
/* ======================================================================== */

void fo_cms_add(uint64_t *cms, int depth, int wlog2, const uint32_t *key, int n_words, uint64_t weight)
{
    uint64_t h = fo_hash64(key, n_words);
    uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32) | 1u;
    uint32_t mask = (1u << wlog2) - 1u;
    for (int j = 0; j < depth; j++) {
        uint32_t idx = (a + (uint32_t)j * b) & mask;
        cms[((size_t)j << wlog2) + idx] += weight;
    }
}

uint64_t fo_cms_estimate(const uint64_t *cms, int depth, int wlog2, const uint32_t *key, int n_words)
{
    uint64_t h = fo_hash64(key, n_words);
    uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32) | 1u;
    uint32_t mask = (1u << wlog2) - 1u;
    uint64_t est = ~0ull;
    for (int j = 0; j < depth; j++) {
        uint32_t idx = (a + (uint32_t)j * b) & mask;
        uint64_t c = cms[((size_t)j << wlog2) + idx];
        if (c < est) est = c;
    }
    return est;
}

static int hh_cmp(const void *x, const void *y)
{
    const fo_hh *a = (const fo_hh *)x, *b = (const fo_hh *)y;
    if (a->estimate > b->estimate) return -1;
    if (a->estimate < b->estimate) return 1;
    for (int i = 0; i < g_cmp_kw; i++) {
        if (a->key[i] < b->key[i]) return -1;
        if (a->key[i] > b->key[i]) return 1;
    }
    return 0;
}

size_t fo_topk(const uint64_t *cms, int depth, int wlog2, int n_words, const fo_row *cands,
               size_t n_cands, size_t k, fo_hh *out)
{
    fo_hh *all = (fo_hh *)malloc((n_cands ? n_cands : 1) * sizeof(fo_hh));
    for (size_t i = 0; i < n_cands; i++) {
        memset(&all[i], 0, sizeof(fo_hh));
        memcpy(all[i].key, cands[i].key, (size_t)n_words * 4);
        all[i].estimate = fo_cms_estimate(cms, depth, wlog2, cands[i].key, n_words);
    }
    g_cmp_kw = n_words;
    qsort(all, n_cands, sizeof(fo_hh), hh_cmp);
    size_t n = n_cands < k ? n_cands : k;
    memcpy(out, all, n * sizeof(fo_hh));
    free(all);
    return n;
}

/* ======================================================================== */
/* batch driver                                                              */
/* ======================================================================== */

typedef struct {
    const uint8_t *const *bufs;      /* slabs (each < 4 GiB: offsets are u32) */
    const uint32_t *const *offsets;
    const size_t *ns;
    int n_slabs;
    size_t lo, hi;                   /* global record range over the concatenated slabs */
    int framed;
    fo_agg *agg;      /* thread-private */
    uint64_t *cms;    /* thread-private or NULL */
    int depth, wlog2, key_mode, kw, scale;
    uint64_t n_bad, n_nokey;
} worker_t;

static void *worker_main(void *arg)
{
    worker_t *w = (worker_t *)arg;
    fo_flow f;
    uint32_t key[FO_MAX_KEY_WORDS];
    size_t base = 0;
    for (int sl = 0; sl < w->n_slabs; base += w->ns[sl], sl++) {
      size_t lo = w->lo > base ? w->lo - base : 0;
      size_t hi = w->hi > base ? w->hi - base : 0;
      if (hi > w->ns[sl]) hi = w->ns[sl];
      const uint8_t *buf = w->bufs[sl];
      const uint32_t *offs = w->offsets[sl];
      for (size_t r = lo; r < hi; r++) {
        int rc = fo_decode_record(buf, offs[r], offs[r + 1], w->framed, &f);
        if (rc != FO_OK) { /* inserter.go:125-126: log, skip the row */
            w->n_bad++;
            continue;
        }
        if (!w->agg && !w->cms) continue;
        if (!fo_make_key(w->key_mode, &f, key)) {
            w->n_nokey++;
            continue;
        }
        uint64_t b = f.bytes, p = f.packets;
        if (w->scale) {
            b *= f.sampling_rate;
            p *= f.sampling_rate;
        }
        if (w->agg) agg_insert(w->agg, key, b, p, 1);
        /* heavy-hitter weight: Bytes*SamplingRate (viz-ch.json:233) */
        if (w->cms) fo_cms_add(w->cms, w->depth, w->wlog2, key, w->kw, f.bytes * f.sampling_rate);
      }
    }
    return NULL;
}

typedef struct {
    worker_t *ws;
    int n_workers, me;
    fo_agg *out;
    uint64_t *cms_out;
    size_t cms_words;
} merge_t;

static void *merge_main(void *arg)
{
    merge_t *m = (merge_t *)arg;
    if (m->out) {
        for (int w = 0; w < m->n_workers; w++) {
            fo_agg *src = m->ws[w].agg;
            for (size_t i = 0; i < src->cap; i++) {
                if (!src->used[i]) continue;
                if ((fo_hash64(src->slots[i].key, src->kw) >> 40) % (uint64_t)m->n_workers != (uint64_t)m->me) continue;
                fo_agg_add_row(m->out, &src->slots[i]);
            }
        }
    }
    if (m->cms_out) {
        size_t lo = m->cms_words * (size_t)m->me / (size_t)m->n_workers;
        size_t hi = m->cms_words * (size_t)(m->me + 1) / (size_t)m->n_workers;
        for (int w = 0; w < m->n_workers; w++) {
            const uint64_t *src = m->ws[w].cms;
            for (size_t i = lo; i < hi; i++) m->cms_out[i] += src[i];
        }
    }
    return NULL;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int fo_run_batch(const uint8_t *buf, const uint32_t *offsets, size_t n, int framed, fo_agg *a,
                 uint64_t *cms, int depth, int wlog2, int n_threads, fo_batch_result *res)
{
    return fo_run_slabs(&buf, &offsets, &n, 1, framed, a, cms, depth, wlog2, n_threads, res);
}

int fo_run_slabs(const uint8_t *const *bufs, const uint32_t *const *offsets, const size_t *ns, int n_slabs, int framed,
                 fo_agg *a, uint64_t *cms, int depth, int wlog2, int n_threads, fo_batch_result *res)
{
    size_t n = 0;
    for (int i = 0; i < n_slabs; i++) n += ns[i];
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n && n > 0) n_threads = (int)n;
    worker_t *ws = (worker_t *)calloc((size_t)n_threads, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    size_t cms_words = cms ? ((size_t)depth << wlog2) : 0;
    int key_mode = a ? a->key_mode : FO_KEY_SRCADDR;
    double t0 = now_s();
    for (int t = 0; t < n_threads; t++) {
        worker_t *w = &ws[t];
        w->bufs = bufs;
        w->offsets = offsets;
        w->ns = ns;
        w->n_slabs = n_slabs;
        w->lo = n * (size_t)t / (size_t)n_threads;
        w->hi = n * (size_t)(t + 1) / (size_t)n_threads;
        w->framed = framed;
        w->key_mode = key_mode;
        w->kw = fo_key_words(key_mode);
        w->scale = a ? a->scale : 0;
        w->depth = depth;
        w->wlog2 = wlog2;
        if (n_threads == 1) {
            w->agg = a;
            w->cms = cms;
        } else {
            w->agg = a ? fo_agg_new(a->key_mode, a->scale) : NULL;
            w->cms = cms ? (uint64_t *)calloc(cms_words, 8) : NULL;
        }
    }
    if (n_threads == 1) {
        worker_main(&ws[0]);
    } else {
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker_main, &ws[t]);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        /* final merge of the per-thread tables, itself parallel: merger t
         * owns the keys with hash % T == t and the sketch columns [t*W/T, (t+1)*W/T) */
        merge_t *ms = (merge_t *)calloc((size_t)n_threads, sizeof(merge_t));
        for (int t = 0; t < n_threads; t++) {
            ms[t].ws = ws;
            ms[t].n_workers = n_threads;
            ms[t].me = t;
            ms[t].out = a ? fo_agg_new(a->key_mode, a->scale) : NULL;
            ms[t].cms_out = cms;
            ms[t].cms_words = cms_words;
            pthread_create(&th[t], NULL, merge_main, &ms[t]);
        }
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        for (int t = 0; t < n_threads; t++) {
            if (ms[t].out) { /* disjoint key sets: plain inserts */
                for (size_t i = 0; i < ms[t].out->cap; i++)
                    if (ms[t].out->used[i]) fo_agg_add_row(a, &ms[t].out->slots[i]);
                fo_agg_free(ms[t].out);
            }
        }
        free(ms);
        for (int t = 0; t < n_threads; t++) {
            if (ws[t].agg) fo_agg_free(ws[t].agg);
            if (ws[t].cms) free(ws[t].cms);
        }
    }
    double t1 = now_s();
    if (res) {
        res->n_records = n;
        res->n_bad = 0;
        res->n_nokey = 0;
        for (int t = 0; t < n_threads; t++) {
            res->n_bad += ws[t].n_bad;
            res->n_nokey += ws[t].n_nokey;
        }
        res->seconds = t1 - t0;
    }
    free(ws);
    free(th);
    return 0;
}
