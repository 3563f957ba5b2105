// inserter.cc -- host-side mirror of inserter/inserter.go on top of libflowagg.
//
// The reference host is Go; this image has no Go toolchain, so the host side above
// the C ABI is written in C++ (the Go/cgo version of the same loop is kept as source
// in host/go/inserter_b200.go and described in INTEGRATION.md).  It keeps the
// reference's shape:
//
//   flags            inserter.go:25-42 (same names, same defaults)
//   state            inserter.go:75-88
//   Setup/Cleanup    inserter.go:167-174
//   ConsumeClaim     inserter.go:176-196  one call per claimed partition; receive ->
//                    buffer -> MarkMessage, flush on a timer
//   buffer           inserter.go:113-165  here: memcpy msg.Value (+ a varint length
//                    prefix when the topic carries bare messages) into the pinned
//                    slab, record its offset; a full slab is handed to fa_submit
//   flush            inserter.go:90-111   here: fa_flush -> flows_5m rows
//                    (create.sh:70-110) written as TSV instead of one SQL INSERT
//                    per flow
//
// Kafka itself is not available here either.  A "claim" is therefore an abstract
// message source (ConsumerGroupClaim below); the shipped implementation reads one
// file per partition holding the Kafka values back to back, length-delimited the
// way mocker -proto.fixedlen writes them (mocker.go:98-101).  Offsets are marked
// only after the slab that holds the message has been submitted, which closes the
// reference's mark-before-durable gap (inserter.go:188).
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cstdarg>
#include <chrono>
#include <cinttypes>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/flowagg.h"

// ---- flags (inserter.go:25-42) ---------------------------------------------------------------
struct Flags {
    std::string LogLevel = "info";
    std::string MetricsAddr = ":8081";
    std::string MetricsPath = "/metrics";
    std::string KafkaVersion = "2.1.1";
    std::string KafkaTopic = "flows-processed";
    std::string KafkaBrk = "127.0.0.1:9092,[::1]:9092";
    std::string KafkaGroup = "postgres-inserter";
    double FlushTime = 5.0;  // -flush.dur, seconds
    long FlushCount = 100;   // -flush.count
    std::string PostgresUser = "postgres", PostgresPass, PostgresHost = "127.0.0.1", PostgresDbName = "postgres";
    int PostgresPort = 5432;
    // additions of this host (no reference counterpart)
    std::vector<std::string> ClaimFiles;  // -claim.file a,b,...: one per partition, stands in for the Kafka claim
    bool FixedLen = true;                 // -proto.fixedlen: values carry the varint length prefix (mocker.go:23)
    std::string Out = "-";                // -out: flows_5m rows (TSV); "-" = stdout
    std::string Key = "flows5m";          // -key: flows5m|aspair|srcaddr|dstaddr|5tuple|srcport|dstport
    int Devices = 1;                      // -gpus: partition p runs on GPU p mod gpus
    std::string Format = "tsv";           // -format tsv | rowbinary: Clickhouse RowBinary of the flows_5m schema (create.sh:70-87)
    std::string Sink = "flows5m";         // -sink flows5m: roll-up rows (create.sh:70-87); rows: the inserter's own
                                          //       14-column row per flow (inserter.go:51-66,142-157), TSV as Go would print it;
                                          //       copy: the same rows as a Postgres `COPY flows (...) FROM stdin` script
    bool FlushBox = false;                // -flush.box: the closing flush is ONE exact roll-up over every partition (fa_flush_box)
    bool OffsetsOnGPU = false;            // -offsets.gpu: hand the slab over WITHOUT its offsets array (fa_submit(offsets = NULL)): the library
                                          //       finds the record boundaries on the GPU; 4 bytes per flow less over PCIe
    bool DryRun = false;                  // -dry-run: walk the claims and fill slabs, no GPU, no aggregates
    bool Metrics = false;                 // -metrics: serve -metrics.addr (off by default in this mirror)
    double Linger = 0;                    // -linger: keep serving metrics this long after the claims are drained (or until SIGTERM)
};

static double parse_duration(const std::string &s)
{  // Go duration subset: 5s, 500ms, 2m, 1h
    char *end = nullptr;
    double v = strtod(s.c_str(), &end);
    std::string u = end ? end : "";
    if (u == "ms") return v / 1e3;
    if (u == "us") return v / 1e6;
    if (u == "m") return v * 60;
    if (u == "h") return v * 3600;
    return v;
}

static bool parse_flags(int argc, char **argv, Flags &f)
{
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        while (!a.empty() && a[0] == '-') a.erase(0, 1);
        std::string v;
        size_t eq = a.find('=');
        bool has = eq != std::string::npos;
        if (has) {
            v = a.substr(eq + 1);
            a = a.substr(0, eq);
        }
        auto val = [&]() -> std::string {
            if (has) return v;
            if (i + 1 < argc) return argv[++i];
            return "";
        };
        if (a == "loglevel") f.LogLevel = val();
        else if (a == "metrics.addr") f.MetricsAddr = val();
        else if (a == "metrics.path") f.MetricsPath = val();
        else if (a == "kafka.version") f.KafkaVersion = val();
        else if (a == "kafka.topic") f.KafkaTopic = val();
        else if (a == "kafka.brokers") f.KafkaBrk = val();
        else if (a == "kafka.group") f.KafkaGroup = val();
        else if (a == "flush.dur") f.FlushTime = parse_duration(val());
        else if (a == "flush.count") f.FlushCount = atol(val().c_str());
        else if (a == "postgres.user") f.PostgresUser = val();
        else if (a == "postgres.pass") f.PostgresPass = val();
        else if (a == "postgres.host") f.PostgresHost = val();
        else if (a == "postgres.port") f.PostgresPort = atoi(val().c_str());
        else if (a == "postgres.dbname") f.PostgresDbName = val();
        else if (a == "claim.file") {
            std::string s = val();
            size_t p = 0;
            while (p <= s.size()) {
                size_t q = s.find(',', p);
                if (q == std::string::npos) q = s.size();
                if (q > p) f.ClaimFiles.push_back(s.substr(p, q - p));
                p = q + 1;
            }
        } else if (a == "proto.fixedlen") f.FixedLen = !has || (v != "false" && v != "0");
        else if (a == "out") f.Out = val();
        else if (a == "key") f.Key = val();
        else if (a == "sink") f.Sink = val();
        else if (a == "format") f.Format = val();
        else if (a == "gpus") f.Devices = atoi(val().c_str());
        else if (a == "dry-run") f.DryRun = true;
        else if (a == "flush.box") f.FlushBox = true;
        else if (a == "offsets.gpu") f.OffsetsOnGPU = true;
        else if (a == "metrics") f.Metrics = true;
        else if (a == "linger") f.Linger = parse_duration(val());
        else {
            fprintf(stderr, "flag provided but not defined: -%s\n", a.c_str());
            return false;
        }
    }
    return true;
}

// ---- logging (logrus levels, inserter.go:201-202) ----------------------------------------------
static int g_level = 2;  // 0 error, 1 warn, 2 info, 3 debug
static std::mutex g_log_mu;
static void logf(int lvl, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
static void logf(int lvl, const char *fmt, ...)
{
    if (lvl > g_level) return;
    static const char *nm[] = {"error", "warning", "info", "debug"};
    std::lock_guard<std::mutex> lk(g_log_mu);
    fprintf(stderr, "level=%s msg=\"", nm[lvl]);
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\"\n");
}

// ---- the seam: sarama.ConsumerGroupSession / ConsumerGroupClaim ---------------------------------
struct ConsumerMessage {  // the fields of sarama.ConsumerMessage the reference reads (inserter.go:180)
    const uint8_t *Value;
    size_t Len;
    std::string Topic;
    int32_t Partition;
    int64_t Offset;
};

struct ConsumerGroupClaim {
    virtual ~ConsumerGroupClaim() {}
    virtual int32_t Partition() const = 0;
    // false = the claim is closed (rebalance / end of input)
    virtual bool Next(ConsumerMessage &m) = 0;
};

struct ConsumerGroupSession {
    std::mutex mu;
    std::map<int32_t, int64_t> marked;  // partition -> next offset to commit
    void MarkMessage(const ConsumerMessage &m)
    {
        std::lock_guard<std::mutex> lk(mu);
        int64_t &o = marked[m.Partition];
        if (m.Offset + 1 > o) o = m.Offset + 1;
    }
};

// One file per partition: Kafka values back to back, each varint(len) || FlowMessage.
struct FileClaim : ConsumerGroupClaim {
    int32_t part;
    std::string topic;
    std::vector<uint8_t> data;
    size_t pos = 0;
    int64_t offset = 0;
    bool fixedlen;
    FileClaim(const std::string &path, int32_t p, const std::string &t, bool fl) : part(p), topic(t), fixedlen(fl)
    {
        FILE *fp = fopen(path.c_str(), "rb");
        if (!fp) {
            logf(0, "cannot open claim file %s", path.c_str());
            exit(1);  // log.Fatal, as the reference does for infrastructure errors (inserter.go:249)
        }
        fseek(fp, 0, SEEK_END);
        long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        data.resize((size_t)n);
        if (n && fread(data.data(), 1, (size_t)n, fp) != (size_t)n) exit(1);
        fclose(fp);
    }
    int32_t Partition() const override { return part; }
    bool Next(ConsumerMessage &m) override
    {
        if (pos >= data.size()) return false;
        // the file is length-delimited; the Kafka VALUE is the framed record (-proto.fixedlen=true)
        // or the bare message (false)
        size_t p = pos;
        uint64_t len = 0;
        int sh = 0;
        while (p < data.size()) {
            uint8_t b = data[p++];
            len |= (uint64_t)(b & 0x7f) << sh;
            sh += 7;
            if (b < 0x80 || sh > 63) break;
        }
        if (len > data.size() - p) len = data.size() - p;  // truncated tail: hand it over, the decoder rejects it
        m.Topic = topic;
        m.Partition = part;
        m.Offset = offset++;
        if (fixedlen) {
            m.Value = data.data() + pos;
            m.Len = (p - pos) + (size_t)len;
        } else {
            m.Value = data.data() + p;
            m.Len = (size_t)len;
        }
        pos = p + (size_t)len;
        return true;
    }
};

// "?" + hexString(ip): what net.IP.String() prints for a length that is neither 4 nor 16
static std::string ip_string_hex(const uint8_t *p, size_t len)
{
    std::string o = "?";
    static const char hx[] = "0123456789abcdef";
    for (size_t i = 0; i < len; i++) {
        o += hx[p[i] >> 4];
        o += hx[p[i] & 15];
    }
    return o;
}

// ---- net.IP(b).String() with the "<nil>" -> "0.0.0.0" patch of inserter.go:131-140 ------------------
static std::string ip_string(const uint8_t *p, size_t len)
{
    char buf[64];
    if (len == 0) return "0.0.0.0";
    const uint8_t *p4 = nullptr;
    if (len == 4) p4 = p;
    if (len == 16) {
        bool z = true;
        for (int i = 0; i < 10; i++) z = z && p[i] == 0;
        if (z && p[10] == 0xff && p[11] == 0xff) p4 = p + 12;  // To4(): v4-mapped
    }
    if (p4) {
        snprintf(buf, sizeof buf, "%u.%u.%u.%u", p4[0], p4[1], p4[2], p4[3]);
        return buf;
    }
    if (len != 16) return ip_string_hex(p, len);
    int e0 = -1, e1 = -1;  // longest run of zero groups, leftmost on ties, at least two groups
    for (int i = 0; i < 16; i += 2) {
        int j = i;
        while (j < 16 && p[j] == 0 && p[j + 1] == 0) j += 2;
        if (j > i && j - i > e1 - e0) {
            e0 = i;
            e1 = j;
            i = j;
        }
    }
    if (e1 - e0 <= 2) e0 = e1 = -1;
    std::string o;
    for (int i = 0; i < 16; i += 2) {
        if (i == e0) {
            o += "::";
            i = e1;
            if (i >= 16) break;
        } else if (i > 0) {
            o += ':';
        }
        snprintf(buf, sizeof buf, "%x", (unsigned)((p[i] << 8) | p[i + 1]));
        o += buf;
    }
    return o;
}

// ---- state (inserter.go:75-88) --------------------------------------------------------------------
static const char *kKeyNames[] = {"flows5m", "aspair", "srcaddr", "dstaddr", "5tuple", "srcport", "dstport"};
// insert_count (inserter.go:44-49, "Inserts made to Postgres", never incremented upstream): one per flow that decoded and
// went to the sink or the roll-up = g_records - g_bad, refreshed wherever the context's counters are read
static std::atomic<uint64_t> g_rows{0}, g_bad{0}, g_flushes{0};  // roll-up rows written, undecodable messages, flushes
static std::atomic<uint64_t> g_records{0}, g_gpu_busy_us{0};    // messages handed to the GPU; device time of its kernels
static std::atomic<bool> g_stop{false};

struct PartitionState {  // what one ConsumeClaim goroutine owns
    fa_ctx *ctx = nullptr;
    int slot = 0;
    uint8_t *slab = nullptr;
    uint32_t *offs = nullptr;
    size_t slab_cap = 0, rec_cap = 0, fill = 0, nrec = 0;
    std::vector<ConsumerMessage> pending;  // marked once their slab is submitted
    uint64_t bad_seen = 0, records_seen = 0, busy_seen = 0;  // fa_stats figures already added to the metrics
};

struct state {
    Flags fl;
    std::atomic<bool> ready{false};  // close(s.ready) in Setup (inserter.go:168)
    std::atomic<long> msgCount{0};
    std::mutex out_mu;  // the reference's s.lock guarded the row buffer; here only the sink is shared
    FILE *out = stdout;
    int key_mode = 0;
    uint64_t rows_written = 0, bad = 0;

    bool per_flow_sink() const { return fl.Sink == "rows" || fl.Sink == "copy"; }

    int Setup(ConsumerGroupSession &)
    {
        ready = true;
        if (fl.Sink == "copy")  // the insert of inserter.go:99-106, as one COPY instead of one INSERT per flow
            fprintf(out, "COPY flows (date_inserted, time_flow, type, sampling_rate, src_ip, dst_ip, bytes, packets, src_port, dst_port, etype, "
                         "proto, src_as, dst_as) FROM stdin;\n");
        return 0;
    }
    int Cleanup(ConsumerGroupSession &)
    {
        if (fl.Sink == "copy") fprintf(out, "\\.\n");  // end-of-data marker of COPY ... FROM stdin
        return 0;
    }

    void acquire_slab(PartitionState &ps)
    {
        if (fl.DryRun) {
            static thread_local std::vector<uint8_t> b;
            static thread_local std::vector<uint32_t> o;
            b.resize(64u << 20);
            o.resize((1u << 20) + 1);
            ps.slab = b.data();
            ps.offs = o.data();
            ps.slab_cap = b.size();
            ps.rec_cap = 1u << 20;
        } else {
            int rc = fa_host_buffer(ps.ctx, ps.slot, &ps.slab, &ps.slab_cap, &ps.offs, &ps.rec_cap);
            if (rc) {
                logf(0, "fa_host_buffer: %s (%s)", fa_strerror(rc), fa_last_error(ps.ctx));
                exit(1);
            }
        }
        ps.fill = 0;
        ps.nrec = 0;
        ps.offs[0] = 0;
    }

    void submit_slab(PartitionState &ps, ConsumerGroupSession &sess)
    {
        if (ps.nrec == 0) return;
        if (!fl.DryRun) {
            // per-flow sinks read the decoded columns of THIS slab right after the submit: they keep the offsets (and the
            // synchronous hand-over); the roll-up sinks may run one slab behind (include/flowagg.h)
            const bool no_offsets = fl.OffsetsOnGPU && !per_flow_sink();
            int rc = fa_submit(ps.ctx, ps.slab, ps.fill, no_offsets ? nullptr : ps.offs, (uint32_t)ps.nrec, FA_FRAMED);
            if (rc) {
                logf(0, "fa_submit: %s (%s)", fa_strerror(rc), fa_last_error(ps.ctx));
                exit(1);
            }
        }
        if (!fl.DryRun && per_flow_sink()) write_flow_rows(ps);
        for (const ConsumerMessage &m : ps.pending) sess.MarkMessage(m);  // after the hand-over, not before (inserter.go:188)
        ps.pending.clear();
        ps.slot ^= 1;
        acquire_slab(ps);
    }

    // (*state).buffer, inserter.go:113-165: here the decode happens on the GPU, so buffering is a memcpy
    void buffer(PartitionState &ps, ConsumerGroupSession &sess, const ConsumerMessage &msg)
    {
        msgCount++;
        size_t need = msg.Len + 10;
        if (ps.fill + need > ps.slab_cap || ps.nrec >= ps.rec_cap) submit_slab(ps, sess);
        uint8_t *p = ps.slab + ps.fill;
        if (!fl.FixedLen) {  // bare value (Postgres path, mocker.go:96-97): add the length prefix ourselves
            uint64_t n = msg.Len;
            while (n >= 0x80) {
                *p++ = (uint8_t)(n | 0x80);
                n >>= 7;
            }
            *p++ = (uint8_t)n;
        }
        memcpy(p, msg.Value, msg.Len);
        p += msg.Len;
        ps.fill = (size_t)(p - ps.slab);
        ps.offs[++ps.nrec] = (uint32_t)ps.fill;
        ps.pending.push_back(msg);
        if (g_level >= 3) logf(3, "%s/%d/%" PRId64, msg.Topic.c_str(), msg.Partition, msg.Offset);
    }

    // (*state).flush, inserter.go:90-111: rows out, table reset
    bool flush(PartitionState &ps, ConsumerGroupSession &sess, bool closing = false)
    {
        logf(2, "Processed %ld records in the last iteration.", msgCount.exchange(0));
        submit_slab(ps, sess);
        if (fl.DryRun || per_flow_sink()) return true;
        if (closing && fl.FlushBox) return true;  // main() merges every partition's table in one fa_flush_box
        size_t n = 0;
        std::vector<fa_row> rows(1 << 16);
        int rc = fa_flush(ps.ctx, rows.data(), rows.size(), &n, 0);
        if (rc == FA_ERR_CAPACITY) {
            rows.resize(n);
            rc = fa_flush(ps.ctx, rows.data(), rows.size(), &n, 0);
        }
        if (rc && rc != FA_ERR_TABLE_FULL) {
            logf(0, "fa_flush: %s (%s)", fa_strerror(rc), fa_last_error(ps.ctx));
            exit(1);
        }
        fa_stats st;
        fa_stats_get(ps.ctx, &st);
        std::lock_guard<std::mutex> lk(out_mu);
        g_bad += st.n_bad - ps.bad_seen;
        ps.bad_seen = st.n_bad;
        g_records += st.n_records - ps.records_seen;
        ps.records_seen = st.n_records;
        g_gpu_busy_us += st.gpu_busy_us - ps.busy_seen;
        ps.busy_seen = st.gpu_busy_us;
        g_flushes++;
        bad = st.n_bad;
        for (size_t i = 0; i < n; i++) write_row(rows[i]);
        fflush(out);
        rows_written += n;
        return true;
    }

    // -sink rows: the reference's own output, one row per decoded flow with the columns of flow_fields
    // (inserter.go:51-66) in the order of inserter.go:142-157; undecodable messages produce no row (:125-126)
    void write_flow_rows(PartitionState &ps)
    {
        const size_t n = ps.nrec;
        std::vector<uint8_t> valid(n), sa(n * 16), da(n * 16), sal(n), dal(n);
        std::vector<uint64_t> tfs(n), sr(n), by(n), pk(n);
        std::vector<uint32_t> ty(n), sp(n), dp(n), et(n), pr(n), sas(n), das(n);
        struct { const char *name; void *dst; size_t bytes; } cols[] = {
            {"valid", valid.data(), n}, {"src_addr", sa.data(), n * 16}, {"dst_addr", da.data(), n * 16}, {"src_addr_len", sal.data(), n},
            {"dst_addr_len", dal.data(), n}, {"time_flow_start", tfs.data(), n * 8}, {"sampling_rate", sr.data(), n * 8}, {"bytes", by.data(), n * 8},
            {"packets", pk.data(), n * 8}, {"type", ty.data(), n * 4}, {"src_port", sp.data(), n * 4}, {"dst_port", dp.data(), n * 4},
            {"etype", et.data(), n * 4}, {"proto", pr.data(), n * 4}, {"src_as", sas.data(), n * 4}, {"dst_as", das.data(), n * 4}};
        for (auto &c : cols) {
            int rc = fa_columns_read(ps.ctx, c.name, c.dst, c.bytes);
            if (rc) {
                logf(0, "fa_columns_read(%s): %s (%s)", c.name, fa_strerror(rc), fa_last_error(ps.ctx));
                exit(1);
            }
        }
        std::lock_guard<std::mutex> lk(out_mu);
        for (size_t i = 0; i < n; i++) {
            if (!valid[i]) {
                bad++;
                continue;
            }
            // net.IP.String() of a value that is neither 4 nor 16 bytes long is "?" + hex of the whole value (inserter.go:131-140).
            // The columns keep the first 16 bytes and the true length (FixedString(16), create.sh:15-16), so a longer address
            // prints as "?" + hex of those 16 bytes + ".." (the tail is not kept anywhere downstream either).
            auto addr_text = [](const uint8_t *p, unsigned len) {
                return len > 16 ? ip_string_hex(p, 16) + ".." : ip_string(p, len);
            };
            const std::string s_ip = addr_text(&sa[i * 16], sal[i]), d_ip = addr_text(&da[i * 16], dal[i]);
            if (fl.Sink == "copy") {
                // Postgres COPY text format: tab-separated literals.  date_inserted is NOW() upstream (a SQL expression,
                // inserter.go:142): its value at the moment of the copy; time_flow = time.Unix(TimeFlowStart, 0) (:143), UTC.
                // The address texts hold only [0-9a-f.:?], nothing COPY would need escaped.
                char now_s[40], tf_s[40];
                const time_t now_t = time(nullptr), tf_t = (time_t)tfs[i];
                struct tm tmv;
                gmtime_r(&now_t, &tmv);
                strftime(now_s, sizeof now_s, "%Y-%m-%d %H:%M:%S+00", &tmv);
                gmtime_r(&tf_t, &tmv);
                strftime(tf_s, sizeof tf_s, "%Y-%m-%d %H:%M:%S+00", &tmv);
                fprintf(out, "%s\t%s\t%d\t%" PRIu64 "\t%s\t%s\t%" PRIu64 "\t%" PRIu64 "\t%u\t%u\t%u\t%u\t%u\t%u\n", now_s, tf_s, (int32_t)ty[i],
                        sr[i], s_ip.c_str(), d_ip.c_str(), by[i], pk[i], sp[i], dp[i], et[i], pr[i], sas[i], das[i]);
                rows_written++;
                continue;
            }
            fprintf(out, "NOW()\t%" PRIu64 "\t%d\t%" PRIu64 "\t%s\t%s\t%" PRIu64 "\t%" PRIu64 "\t%u\t%u\t%u\t%u\t%u\t%u\n", tfs[i], (int32_t)ty[i], sr[i],
                    s_ip.c_str(), d_ip.c_str(), by[i], pk[i], sp[i], dp[i], et[i], pr[i], sas[i], das[i]);
            rows_written++;
        }
        fflush(out);
    }

    // Clickhouse RowBinary of one flows_5m row: Date UInt16 (days), Timeslot DateTime UInt32, SrcAS, DstAS UInt32, the four
    // one-element ETypeMap arrays (LEB128 length 1 + element), Bytes, Packets, Count UInt64 -- 70 bytes, little-endian;
    // what `INSERT INTO flows_5m FORMAT RowBinary` takes (create.sh:70-87, :100-107)
    void write_row_binary(const fa_row &r)
    {
        uint8_t b[70];
        size_t o = 0;
        auto put = [&](uint64_t v, int n) {
            for (int i = 0; i < n; i++) b[o++] = (uint8_t)(v >> (8 * i));
        };
        put(r.key[0] / 86400u, 2);
        put(r.key[0], 4);
        put(r.key[1], 4);
        put(r.key[2], 4);
        put(1, 1);
        put(r.key[3], 4);
        put(1, 1);
        put(r.bytes, 8);
        put(1, 1);
        put(r.packets, 8);
        put(1, 1);
        put(r.count, 8);
        put(r.bytes, 8);
        put(r.packets, 8);
        put(r.count, 8);
        fwrite(b, 1, o, out);
    }

    void write_row(const fa_row &r)
    {
        g_rows++;
        if (key_mode == FA_KEY_FLOWS5M && fl.Format == "rowbinary") return write_row_binary(r);
        if (key_mode == FA_KEY_FLOWS5M) {
            // flows_5m columns (create.sh:70-87): Date, Timeslot, SrcAS, DstAS, ETypeMap.EType, .Bytes, .Packets,
            // .Count, Bytes, Packets, Count
            time_t t = (time_t)r.key[0];
            struct tm g;
            gmtime_r(&t, &g);
            char d[16], ts[32];
            strftime(d, sizeof d, "%Y-%m-%d", &g);
            strftime(ts, sizeof ts, "%Y-%m-%d %H:%M:%S", &g);
            fprintf(out, "%s\t%s\t%u\t%u\t[%u]\t[%" PRIu64 "]\t[%" PRIu64 "]\t[%" PRIu64 "]\t%" PRIu64 "\t%" PRIu64 "\t%" PRIu64 "\n", d, ts,
                    r.key[1], r.key[2], r.key[3], r.bytes, r.packets, r.count, r.bytes, r.packets, r.count);
        } else {
            static const int kw[] = {4, 2, 4, 4, 11, 1, 1};
            for (int i = 0; i < kw[key_mode]; i++) fprintf(out, "%u\t", r.key[i]);
            fprintf(out, "%" PRIu64 "\t%" PRIu64 "\t%" PRIu64 "\n", r.bytes, r.packets, r.count);
        }
    }

    // (*state).ConsumeClaim, inserter.go:176-196
    int ConsumeClaim(ConsumerGroupSession &sess, ConsumerGroupClaim &claim, PartitionState &ps)
    {
        acquire_slab(ps);
        auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(fl.FlushTime);
        ConsumerMessage m;
        while (!g_stop) {
            if (!claim.Next(m)) break;  // claim.Messages() closed
            buffer(ps, sess, m);
            if (fl.FlushCount > 0 && msgCount >= fl.FlushCount) flush(ps, sess);  // inserter.go:118,161-163
            if (std::chrono::steady_clock::now() >= deadline) {                  // case <-s.flushTimer (inserter.go:189-191)
                flush(ps, sess);
                deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(fl.FlushTime);
            }
        }
        flush(ps, sess, true);
        return 0;
    }

    // -flush.box: Kafka does not partition by group key, so the partitions hold partial sums of the same keys.
    // A SummingMergeTree would merge the partial rows later (create.sh:88-90); this emits them merged already.
    void flush_box(std::vector<PartitionState> &parts)
    {
        std::vector<fa_ctx *> ctxs;
        for (auto &ps : parts)
            if (ps.ctx) ctxs.push_back(ps.ctx);
        if (ctxs.empty()) return;
        size_t n = 0;
        std::vector<fa_row> rows(1 << 16);
        int rc = fa_flush_box(ctxs.data(), (int)ctxs.size(), rows.data(), rows.size(), &n, 0);
        if (rc == FA_ERR_CAPACITY) {
            rows.resize(n);
            rc = fa_flush_box(ctxs.data(), (int)ctxs.size(), rows.data(), rows.size(), &n, 0);
        }
        if (rc && rc != FA_ERR_TABLE_FULL) {
            logf(0, "fa_flush_box: %s (%s)", fa_strerror(rc), fa_last_error(ctxs[0]));
            exit(1);
        }
        bad = 0;
        for (auto &ps : parts) {
            if (!ps.ctx) continue;
            fa_stats st;
            fa_stats_get(ps.ctx, &st);
            bad += st.n_bad;
            g_bad += st.n_bad - ps.bad_seen;
            ps.bad_seen = st.n_bad;
        }
        g_flushes++;
        for (size_t i = 0; i < n; i++) write_row(rows[i]);
        fflush(out);
        rows_written += n;
    }
};

// ---- metrics endpoint (inserter.go:69-73): insert_count, actually incremented here -------------------
static void metricsHTTP(const Flags &fl)
{
    int port = 8081;
    size_t c = fl.MetricsAddr.rfind(':');
    if (c != std::string::npos) port = atoi(fl.MetricsAddr.c_str() + c + 1);
    int s = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(s, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_ANY);
    a.sin_port = htons((uint16_t)port);
    if (bind(s, (sockaddr *)&a, sizeof a) || listen(s, 8)) {
        logf(1, "metrics: cannot listen on %s", fl.MetricsAddr.c_str());
        return;
    }
    while (!g_stop) {
        int cfd = accept(s, nullptr, nullptr);
        if (cfd < 0) continue;
        char req[1024];
        ssize_t n = read(cfd, req, sizeof req - 1);
        (void)n;
        char body[2048], resp[2400];
        int bl = snprintf(body, sizeof body,
                          "# HELP insert_count Inserts made to Postgres.\n# TYPE insert_count counter\ninsert_count %" PRIu64 "\n"
                          "# HELP flowagg_rollup_rows_total Aggregate rows written to the sink.\n# TYPE flowagg_rollup_rows_total counter\n"
                          "flowagg_rollup_rows_total %" PRIu64 "\n"
                          "# HELP flowagg_bad_records_total Messages proto.Unmarshal would reject (skipped, inserter.go:125).\n"
                          "# TYPE flowagg_bad_records_total counter\nflowagg_bad_records_total %" PRIu64 "\n"
                          "# HELP flowagg_flushes_total Roll-up flushes.\n# TYPE flowagg_flushes_total counter\nflowagg_flushes_total %" PRIu64 "\n"
                          "# HELP flowagg_records_total Kafka messages decoded on the GPU.\n# TYPE flowagg_records_total counter\n"
                          "flowagg_records_total %" PRIu64 "\n"
                          "# HELP flowagg_gpu_busy_seconds_total Device time of the decode/aggregate kernels (CUDA events).\n"
                          "# TYPE flowagg_gpu_busy_seconds_total counter\nflowagg_gpu_busy_seconds_total %.6f\n",
                          g_records.load() - g_bad.load(), g_rows.load(), g_bad.load(), g_flushes.load(), g_records.load(),
                          (double)g_gpu_busy_us.load() / 1e6);
        int rl = snprintf(resp, sizeof resp, "HTTP/1.1 200 OK\r\nContent-Type: text/plain; version=0.0.4\r\nContent-Length: %d\r\n\r\n%s", bl, body);
        if (write(cfd, resp, (size_t)rl) < 0) {}
        close(cfd);
    }
    close(s);
}

static void on_signal(int) { g_stop = true; }

int main(int argc, char **argv)
{
    state s;
    if (!parse_flags(argc, argv, s.fl)) return 2;
    const char *lv[] = {"error", "warning", "info", "debug"};
    for (int i = 0; i < 4; i++)
        if (s.fl.LogLevel == lv[i]) g_level = i;
    for (int i = 0; i < 7; i++)
        if (s.fl.Key == kKeyNames[i]) s.key_mode = i;
    if (s.fl.ClaimFiles.empty()) {
        logf(0, "no Kafka client in this build: give the claimed partitions as files with -claim.file a,b,...");
        return 2;
    }
    if (s.fl.Out != "-") s.out = fopen(s.fl.Out.c_str(), "w");
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);
    std::thread metrics;
    if (s.fl.Metrics) metrics = std::thread(metricsHTTP, std::cref(s.fl));

    ConsumerGroupSession sess;
    s.Setup(sess);
    const size_t np = s.fl.ClaimFiles.size();
    std::vector<std::unique_ptr<FileClaim>> claims;
    std::vector<PartitionState> parts(np);
    for (size_t p = 0; p < np; p++) {
        claims.emplace_back(new FileClaim(s.fl.ClaimFiles[p], (int32_t)p, s.fl.KafkaTopic, s.fl.FixedLen));
        if (!s.fl.DryRun) {
            fa_config cfg{};
            cfg.abi_version = FA_ABI_VERSION;
            cfg.device = (int32_t)(p % (size_t)(s.fl.Devices > 0 ? s.fl.Devices : 1));
            cfg.key_mode = (uint32_t)s.key_mode;
            cfg.max_batch_bytes = 64u << 20;
            cfg.max_batch_records = 1u << 20;
            if (s.per_flow_sink()) cfg.flags = FA_CFG_COLUMNS | FA_CFG_NO_AGGREGATE;  // kernel 1 alone
            int rc = fa_create(&cfg, &parts[p].ctx);
            if (rc) {
                logf(0, "fa_create: %s (%s)", fa_strerror(rc), parts[p].ctx ? fa_last_error(parts[p].ctx) : "");
                return 1;  // no CPU fallback: the stage needs its GPU
            }
        }
    }
    // sarama runs one ConsumeClaim goroutine per claimed partition (inserter.go:176); so do we
    std::vector<std::thread> th;
    for (size_t p = 0; p < np; p++) th.emplace_back([&, p] { s.ConsumeClaim(sess, *claims[p], parts[p]); });
    for (auto &t : th) t.join();
    if (s.fl.FlushBox && !s.fl.DryRun && !s.per_flow_sink()) s.flush_box(parts);
    s.Cleanup(sess);
    uint64_t total = 0;
    for (auto &kv : sess.marked) total += (uint64_t)kv.second;
    logf(2, "done: %" PRIu64 " messages marked over %zu partitions, %" PRIu64 " rows written, %" PRIu64 " undecodable", total, np, s.rows_written,
         s.bad);
    for (auto &ps : parts) {
        if (!ps.ctx) continue;
        fa_stats st;
        if (fa_stats_get(ps.ctx, &st) == FA_OK) {  // final figures for the metrics endpoint
            g_bad += st.n_bad - ps.bad_seen;
            ps.bad_seen = st.n_bad;
            g_records += st.n_records - ps.records_seen;
            ps.records_seen = st.n_records;
            g_gpu_busy_us += st.gpu_busy_us - ps.busy_seen;
            ps.busy_seen = st.gpu_busy_us;
        }
    }
    if (s.fl.Metrics && s.fl.Linger > 0) {  // a scraper gets to see the final counters
        const auto until = std::chrono::steady_clock::now() + std::chrono::duration<double>(s.fl.Linger);
        while (!g_stop && std::chrono::steady_clock::now() < until) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    for (auto &ps : parts)
        if (ps.ctx) fa_destroy(ps.ctx);
    g_stop = true;
    if (metrics.joinable()) metrics.detach();
    if (s.out != stdout) fclose(s.out);
    return 0;
}
