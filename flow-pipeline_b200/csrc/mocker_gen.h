// mocker_gen.h -- the synthetic FlowMessage producer, shared by host and device.
//
// Follows the record constructor of mocker/mocker.go:57-90 field for field and
// the two framings of mocker.go:95-102.  The reference draws from Go's global
// math/rand (never seeded in code, mocker.go:9) and throttles to ~4 msg/s
// (mocker.go:56), so no reproducible stream exists upstream; here every field is
// a pure function of (seed, record index) -- a counter-based generator -- so the
// CPU and the GPU emit identical bytes and 10^8..10^10 flows can be produced
// where they are consumed.
//
// Encoding = what proto.Marshal emits for this message (golang/protobuf v1.4.3):
// ascending field number, proto3 zero-omission, minimal varints.
#pragma once
#include <stdint.h>

#include "../../include/flowagg.h"

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD static inline
#endif

FA_HD uint64_t fa_mix64(uint64_t x)
{
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

FA_HD uint64_t fa_mocker_draw(uint64_t seed, uint64_t index, uint32_t k)
{
    uint64_t s = fa_mix64(seed + 0x9E3779B97F4A7C15ull * (index + 1));
    return fa_mix64(s + 0xD1B54A32D192ED03ull * (uint64_t)(k + 1));
}

// One decoded mocker record (before encoding).
struct fa_mocker_fields {
    uint64_t time_received, time_flow_start, sampling_rate, bytes, packets;
    uint32_t sequence_num, src_as, dst_as, src_port, dst_port, etype;
    uint8_t src_addr[16], dst_addr[16];
};

// Cumulative octave weights for the Zipf-like address rank: octave k (ranks
// [2^k, 2^(k+1))) has mass ~ 2^(-0.1 k), i.e. p(rank r) ~ r^-1.1.  Thresholds are
// floor(2^32 * sum_{i<=k} 2^(-0.1 i) / sum_{i<24} 2^(-0.1 i)); integer-only so CPU
// and GPU agree bit for bit.
FA_HD uint32_t fa_zipf24_rank(uint64_t r)
{
    const uint32_t thr[24] = {0x1526a169u, 0x28e2a8cfu, 0x3b4c5e7au, 0x4c7a6a69u, 0x5c81f031u, 0x6b76a903u,
                              0x796afbeeu, 0x86701485u, 0x9295f7ffu, 0x9deb98eeu, 0xa87ee9a3u, 0xb25ced56u,
                              0xbb91c82bu, 0xc428ce23u, 0xcc2c9107u, 0xd3a6ed70u, 0xdaa116e5u, 0xe123a331u,
                              0xe73694eeu, 0xece16565u, 0xf22b0dbfu, 0xf71a0f99u, 0xfbb47d04u, 0xffffffffu};
    uint32_t u = (uint32_t)r;
    uint32_t k = 0;
    while (k < 23 && u > thr[k]) k++;
    uint32_t within = (uint32_t)(r >> 32) & ((1u << k) - 1u);
    return (1u << k) + within - 1u;  // 0 .. 2^24-2
}

FA_HD void fa_mocker_make(const fa_mocker_config &c, uint64_t index, fa_mocker_fields &f)
{
    const uint64_t r0 = fa_mocker_draw(c.seed, index, 0);
    const uint64_t r1 = fa_mocker_draw(c.seed, index, 1);
    const uint64_t r2 = fa_mocker_draw(c.seed, index, 2);
    // mocker.go:57 ts := now (seconds); here a deterministic clock
    const uint64_t ts = c.t0 + (c.flows_per_second ? index / c.flows_per_second : 0);
    f.time_received = ts;    // mocker.go:85
    f.time_flow_start = ts;  // mocker.go:84
    f.sampling_rate = 1;     // mocker.go:76
    f.bytes = (uint32_t)r0 % 1500u;          // mocker.go:59
    f.packets = (uint32_t)(r0 >> 32) % 100u; // mocker.go:60
    const uint32_t nsa = c.n_src_as ? c.n_src_as : 3u, nda = c.n_dst_as ? c.n_dst_as : 3u;
    f.src_as = 65000u + (uint32_t)r1 % nsa;          // mocker.go:61,79
    f.dst_as = 65000u + (uint32_t)(r1 >> 32) % nda;  // mocker.go:62,80
    f.etype = 0x86dd;                                // mocker.go:81
    f.src_port = (uint32_t)r2 & 0xFFFFu;             // mocker.go:73,86
    f.dst_port = (uint32_t)(r2 >> 16) & 0xFFFFu;     // mocker.go:74,87
    f.sequence_num = (uint32_t)index;                // mocker.go:88,90
    // mocker.go:64-71: 2001:0db8:0000:0001:0000:0000:0000:00XX
    const uint8_t pfx[8] = {0x20, 0x01, 0x0d, 0xb8, 0x00, 0x00, 0x00, 0x01};
    for (int i = 0; i < 8; i++) {
        f.src_addr[i] = pfx[i];
        f.dst_addr[i] = pfx[i];
        f.src_addr[8 + i] = 0;
        f.dst_addr[8 + i] = 0;
    }
    f.dst_addr[15] = (uint8_t)(r2 >> 40);
    if (c.addr_mode == FA_ADDR_ZIPF24) {
        const uint32_t rank = fa_zipf24_rank(fa_mocker_draw(c.seed, index, 3));
        f.src_addr[13] = (uint8_t)(rank >> 16);
        f.src_addr[14] = (uint8_t)(rank >> 8);
        f.src_addr[15] = (uint8_t)rank;
    } else if (c.addr_mode == FA_ADDR_UNIQUE) {
        for (int i = 0; i < 8; i++) f.src_addr[8 + i] = (uint8_t)(index >> (56 - 8 * i));
    } else {
        f.src_addr[15] = (uint8_t)(r2 >> 32);
    }
}

FA_HD uint32_t fa_varint_len(uint64_t v)
{
    uint32_t n = 1;
    while (v >= 0x80) {
        v >>= 7;
        n++;
    }
    return n;
}

FA_HD uint8_t *fa_put_varint(uint8_t *p, uint64_t v)
{
    while (v >= 0x80) {
        *p++ = (uint8_t)(v | 0x80);
        v >>= 7;
    }
    *p++ = (uint8_t)v;
    return p;
}

// size of the bare message
FA_HD uint32_t fa_mocker_msg_len(const fa_mocker_fields &f)
{
    uint32_t n = 0;
    if (f.time_received) n += 1 + fa_varint_len(f.time_received);      // field 2
    if (f.sampling_rate) n += 1 + fa_varint_len(f.sampling_rate);      // field 3
    if (f.sequence_num) n += 1 + fa_varint_len(f.sequence_num);        // field 4
    n += 2 + 16;                                                        // field 6
    n += 2 + 16;                                                        // field 7
    if (f.bytes) n += 1 + fa_varint_len(f.bytes);                       // field 9
    if (f.packets) n += 1 + fa_varint_len(f.packets);                   // field 10
    if (f.src_as) n += 1 + fa_varint_len(f.src_as);                     // field 14
    if (f.dst_as) n += 1 + fa_varint_len(f.dst_as);                     // field 15
    if (f.src_port) n += 2 + fa_varint_len(f.src_port);                 // field 21
    if (f.dst_port) n += 2 + fa_varint_len(f.dst_port);                 // field 22
    if (f.etype) n += 2 + fa_varint_len(f.etype);                       // field 30
    if (f.time_flow_start) n += 2 + fa_varint_len(f.time_flow_start);   // field 38
    return n;
}

FA_HD uint8_t *fa_mocker_put_msg(const fa_mocker_fields &f, uint8_t *p)
{
    if (f.time_received) { *p++ = 0x10; p = fa_put_varint(p, f.time_received); }
    if (f.sampling_rate) { *p++ = 0x18; p = fa_put_varint(p, f.sampling_rate); }
    if (f.sequence_num) { *p++ = 0x20; p = fa_put_varint(p, f.sequence_num); }
    *p++ = 0x32; *p++ = 16;
    for (int i = 0; i < 16; i++) *p++ = f.src_addr[i];
    *p++ = 0x3a; *p++ = 16;
    for (int i = 0; i < 16; i++) *p++ = f.dst_addr[i];
    if (f.bytes) { *p++ = 0x48; p = fa_put_varint(p, f.bytes); }
    if (f.packets) { *p++ = 0x50; p = fa_put_varint(p, f.packets); }
    if (f.src_as) { *p++ = 0x70; p = fa_put_varint(p, f.src_as); }
    if (f.dst_as) { *p++ = 0x78; p = fa_put_varint(p, f.dst_as); }
    if (f.src_port) { *p++ = 0xa8; *p++ = 0x01; p = fa_put_varint(p, f.src_port); }
    if (f.dst_port) { *p++ = 0xb0; *p++ = 0x01; p = fa_put_varint(p, f.dst_port); }
    if (f.etype) { *p++ = 0xf0; *p++ = 0x01; p = fa_put_varint(p, f.etype); }
    if (f.time_flow_start) { *p++ = 0xb0; *p++ = 0x02; p = fa_put_varint(p, f.time_flow_start); }
    return p;
}

// framed (mocker.go:98-101) or bare (mocker.go:96-97) record size
FA_HD uint32_t fa_mocker_record_len(const fa_mocker_config &c, uint64_t index)
{
    fa_mocker_fields f;
    fa_mocker_make(c, index, f);
    const uint32_t m = fa_mocker_msg_len(f);
    return c.framed ? m + fa_varint_len(m) : m;
}

FA_HD uint8_t *fa_mocker_record_put(const fa_mocker_config &c, uint64_t index, uint8_t *p)
{
    fa_mocker_fields f;
    fa_mocker_make(c, index, f);
    if (c.framed) p = fa_put_varint(p, fa_mocker_msg_len(f));
    return fa_mocker_put_msg(f, p);
}
