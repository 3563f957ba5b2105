// frame.cuh -- record-boundary discovery on the GPU for submits that carry no
// offsets (a length-delimited stream as Clickhouse's Kafka engine reads it,
// compose/clickhouse/create.sh:28-34; framing of mocker/mocker.go:98-101).
//
// Finding the boundaries of varint(len)||message records is a sequential
// dependency (the next boundary is only known once the previous length is
// read).  The stream is cut into fixed tiles and solved by speculate-and-verify:
//
//   1. speculate  one warp per tile looks for the tile's entry -- the first record boundary at or behind the
//                 tile's first byte -- among the next 256 byte positions: a position qualifies when THREE
//                 consecutive non-empty records parse from it as protobuf messages whose fields chain exactly to
//                 the announced lengths, field numbers ascending.  (Walking length prefixes from a random byte,
//                 round 1's heuristic, does not work on flow records: they are so regular that a false chain
//                 keeps its phase from record to record and never meets the true one -- measured: 1 of 32 walkers
//                 on the true chain after 8 KiB, one verify round PER TILE, 33 s for 1.4 GB.)  The smallest
//                 qualifying position is the guess (99 % right on mocker streams); none: the tile's first byte.
//                 Tile 0 enters at 0.
//   2. walk       every dirty tile walks from its entry to its end, producing its
//                 exit (= the next tile's true entry IF its own entry was true)
//                 and its record count.
//   3. verify     entry[i] != exit[i-1] -> entry[i] = exit[i-1], tile i dirty.
//                 Repeat 2-3 until no tile is dirty (two rounds are enqueued without looking; the host
//                 synchronises once, after the scan of step 4).  By induction from tile 0
//                 every entry is then exact, whatever the guesses were: the
//                 speculation only decides how many rounds it takes (typically
//                 two; adversarial input degrades to one tile per round).
//   4. emit       exclusive scan of the counts (cub) and a last walk that writes
//                 offsets[].
//
// A stream that ends inside a record yields one final span [start, len) which
// the decoder then rejects and counts as a bad record (inserter.go:125-126).
#pragma once

namespace fa {

constexpr uint32_t kFrameTile = 16384;
constexpr uint32_t kFrameSearch = 256;   // byte positions behind a tile's start that are tried as its entry
constexpr uint32_t kFrameConfirm = 3;    // consecutive records that must parse from a candidate

// Read the record header at pos: returns the position of the next record, or
// `len` if the header is malformed / the record runs past the end.  One aligned 32-bit load covers the usual 1-2
// byte length prefix (the buffer is readable up to len rounded up to 16, include/flowagg.h); longer prefixes take
// the byte loop.
__device__ __forceinline__ unsigned long long frame_next(const uint8_t *buf, unsigned long long pos, unsigned long long len)
{
    if (pos >= len) return len;
    const uint32_t sh = (uint32_t)(pos & 3ull) * 8u;
    const uint32_t word = __ldg(reinterpret_cast<const uint32_t *>(buf + (pos & ~3ull))) >> sh;  // 4 - (pos & 3) bytes
    const uint32_t b0 = word & 0xffu;
    if (!(b0 & 0x80u)) {
        const unsigned long long body = pos + 1;
        return b0 > len - body ? len : body + b0;
    }
    if (sh < 24u && pos + 1 < len) {
        const uint32_t b1 = (word >> 8) & 0xffu;
        if (!(b1 & 0x80u)) {
            const unsigned long long body = pos + 2, v = (b0 & 0x7fu) | (b1 << 7);
            return v > len - body ? len : body + v;
        }
    }
    unsigned long long v = 0;
    uint32_t i = 0;
    for (; i < 10; i++) {
        if (pos + i >= len) return len;
        const unsigned long long y = __ldg(buf + pos + i);
        if (i == 9) {
            if (y >= 2) return len;
            v |= y << 63;
            i++;
            break;
        }
        v |= (y & 0x7f) << (7 * i);
        if (y < 0x80) {
            i++;
            break;
        }
    }
    const unsigned long long body = pos + i;
    if (v > len - body) return len;
    return body + v;
}

// <= 2-byte varint at p (p < end).  n = bytes taken, 0 if it is longer or cut off.
__device__ __forceinline__ uint32_t frame_varint2(const uint8_t *buf, unsigned long long p, unsigned long long end, uint32_t &n)
{
    const uint32_t b0 = __ldg(buf + p);
    if (b0 < 0x80u) {
        n = 1;
        return b0;
    }
    n = 0;
    if (p + 1 >= end) return 0;
    const uint32_t b1 = __ldg(buf + p + 1);
    if (b1 >= 0x80u) return 0;
    n = 2;
    return (b0 & 0x7fu) | (b1 << 7);
}

// Does a plausible record start at p?  varint(len) with len >= 2, then fields with ascending numbers whose values chain
// exactly to p + header + len.  Returns the position behind the record, or 0.  A heuristic for the speculation only.
__device__ __forceinline__ unsigned long long frame_plausible_record(const uint8_t *buf, unsigned long long p, unsigned long long len)
{
    if (p >= len) return 0;
    uint32_t hn;
    const uint32_t ln = frame_varint2(buf, p, len, hn);
    if (!hn || ln < 2u) return 0;
    unsigned long long q = p + hn;
    const unsigned long long end = q + ln;
    if (end > len) return 0;
    uint32_t last = 0;
    while (q < end) {
        uint32_t tn;
        const uint32_t tag = frame_varint2(buf, q, end, tn);
        if (!tn) return 0;
        q += tn;
        const uint32_t num = tag >> 3, wt = tag & 7u;
        if (num <= last) return 0;
        last = num;
        if (wt == 0u) {
            uint32_t k = 0;
            for (;; k++) {
                if (q + k >= end || k >= 10u) return 0;
                if (__ldg(buf + q + k) < 0x80u) break;
            }
            q += k + 1u;
        } else if (wt == 2u) {
            if (q >= end) return 0;
            uint32_t vn;
            const uint32_t v = frame_varint2(buf, q, end, vn);
            if (!vn) return 0;
            q += vn + v;
        } else if (wt == 1u) {
            q += 8u;
        } else if (wt == 5u) {
            q += 4u;
        } else {
            return 0;
        }
        if (q > end) return 0;
    }
    return end;
}

__global__ void k_frame_speculate(const uint8_t *buf, unsigned long long len, uint32_t n_tiles, uint32_t *entry, uint8_t *dirty)
{
    const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    const unsigned long long start = (unsigned long long)tile * kFrameTile;
    uint32_t best = 0xFFFFFFFFu;
    if (tile == 0) {
        best = 0;
    } else {
        for (uint32_t j = 0; j < kFrameSearch / 32u && best == 0xFFFFFFFFu; j++) {
            unsigned long long p = start + j * 32u + lane;
            const unsigned long long cand = p;
            bool ok = true;
            for (uint32_t k = 0; k < kFrameConfirm && ok && p < len; k++) {
                p = frame_plausible_record(buf, p, len);
                ok = p != 0;
            }
            ok = ok && cand < len;
            const uint32_t hits = __ballot_sync(0xFFFFFFFFu, ok);
            if (hits) best = (uint32_t)(start + j * 32u) + (uint32_t)(__ffs((int)hits) - 1);
        }
        if (best == 0xFFFFFFFFu) best = (uint32_t)start;  // nothing qualified: any guess will do, the verify rounds fix it
    }
    if (lane == 0) {
        entry[tile] = best;
        dirty[tile] = 1;
    }
}

__global__ void k_frame_walk(const uint8_t *buf, unsigned long long len, uint32_t n_tiles, const uint32_t *entry, uint32_t *exit_pos,
                             uint32_t *count, uint8_t *dirty)
{
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles || !dirty[tile]) return;
    dirty[tile] = 0;
    const unsigned long long end = min((unsigned long long)(tile + 1) * kFrameTile, len);
    unsigned long long pos = entry[tile];
    uint32_t n = 0;
    while (pos < end) {
        pos = frame_next(buf, pos, len);
        n++;
    }
    exit_pos[tile] = (uint32_t)pos;
    count[tile] = n;
}

__global__ void k_frame_verify(uint32_t n_tiles, uint32_t *entry, const uint32_t *exit_pos, uint8_t *dirty, uint32_t *n_dirty)
{
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile == 0 || tile >= n_tiles) return;
    const uint32_t want = exit_pos[tile - 1];
    if (entry[tile] != want) {
        entry[tile] = want;
        dirty[tile] = 1;
        atomicAdd(n_dirty, 1u);
    }
}

__global__ void k_frame_emit(const uint8_t *buf, unsigned long long len, uint32_t n_tiles, const uint32_t *entry, const uint32_t *base,
                             uint32_t total, uint32_t *offsets)
{
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    const unsigned long long end = min((unsigned long long)(tile + 1) * kFrameTile, len);
    unsigned long long pos = entry[tile];
    uint32_t at = base[tile];
    while (pos < end) {
        offsets[at++] = (uint32_t)pos;
        pos = frame_next(buf, pos, len);
    }
    if (tile == n_tiles - 1) offsets[total] = (uint32_t)len;
}

}  // namespace fa

// Builds c->d_frame_off (n+1 offsets) for the framed stream d_buf[0,len) on c->stream.
static int frame_index_device(fa_ctx *c, const uint8_t *d_buf, size_t len, uint32_t *n_found)
{
    using namespace fa;
    *n_found = 0;
    if (len == 0) return FA_OK;
    const uint32_t n_tiles = (uint32_t)((len + kFrameTile - 1) / kFrameTile);
    // scratch: entry, exit, count, base (u32 each) + dirty (u8) + n_dirty + cub temp
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_tiles, c->stream);
    const size_t words = (size_t)n_tiles * 4 + 16;
    const size_t need = words * 4 + n_tiles + 256 + cub_bytes + 256;
    int rc = ensure_scratch(c, need);
    if (rc) return rc;
    uint32_t *entry = (uint32_t *)c->d_scratch;
    uint32_t *exit_pos = entry + n_tiles;
    uint32_t *count = exit_pos + n_tiles;
    uint32_t *base = count + n_tiles;
    uint32_t *n_dirty = base + n_tiles;
    uint8_t *dirty = (uint8_t *)(n_dirty + 16);
    void *cub_tmp = (void *)(((uintptr_t)(dirty + n_tiles) + 255) & ~(uintptr_t)255);

    const int tpb = 256;
    k_frame_speculate<<<(n_tiles * 32 + tpb - 1) / tpb, tpb, 0, c->stream>>>(d_buf, len, n_tiles, entry, dirty);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    const int g = (int)((n_tiles + tpb - 1) / tpb);
    // Two walk/verify rounds are enqueued blind (the speculation is almost always right, so the second finds nothing dirty
    // and costs two near-empty launches), then the scan; the host looks ONCE at {tiles still dirty, record total}.  Only
    // adversarial input needs further rounds, each with its own look.
    auto round = [&]() -> int {
        k_frame_walk<<<g, tpb, 0, c->stream>>>(d_buf, len, n_tiles, entry, exit_pos, count, dirty);
        FA_CUDA(c, cudaMemsetAsync(n_dirty, 0, 4, c->stream));
        k_frame_verify<<<g, tpb, 0, c->stream>>>(n_tiles, entry, exit_pos, dirty, n_dirty);
        c->n_kernels += 2;
        FA_CUDA(c, cudaGetLastError());
        return FA_OK;
    };
    rc = round();
    if (rc) return rc;
    rc = round();
    if (rc) return rc;
    uint32_t total = 0;
    for (uint32_t extra = 0;; extra++) {
        FA_CUDA(c, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, count, base, (int)n_tiles, c->stream));
        uint32_t nd = 0, last_base = 0, last_count = 0;
        FA_CUDA(c, cudaMemcpyAsync(&nd, n_dirty, 4, cudaMemcpyDeviceToHost, c->stream));
        FA_CUDA(c, cudaMemcpyAsync(&last_base, base + n_tiles - 1, 4, cudaMemcpyDeviceToHost, c->stream));
        FA_CUDA(c, cudaMemcpyAsync(&last_count, count + n_tiles - 1, 4, cudaMemcpyDeviceToHost, c->stream));
        FA_CUDA(c, cudaStreamSynchronize(c->stream));
        total = last_base + last_count;
        static const bool dbg = getenv("FA_DEBUG_FRAME") != nullptr;
        if (dbg) fprintf(stderr, "[frame] %u tiles, look %u: %u dirty, %u records\n", n_tiles, extra, nd, total);
        if (nd == 0) break;
        if (extra > n_tiles + 1) {
            c->last_error = "frame index did not converge";
            return FA_ERR_FRAMING;
        }
        rc = round();
        if (rc) return rc;
    }
    if ((size_t)total + 1 > c->frame_off_cap) {  // grow-only
        if (c->d_frame_off) FA_CUDA(c, cudaFree(c->d_frame_off));
        c->d_frame_off = nullptr;
        c->frame_off_cap = 0;
        const size_t cap = std::max<size_t>((size_t)total + 1, 1024) * 5 / 4;
        FA_CUDA(c, cudaMalloc(&c->d_frame_off, cap * 4));
        c->frame_off_cap = cap;
    }
    k_frame_emit<<<g, tpb, 0, c->stream>>>(d_buf, len, n_tiles, entry, base, total, c->d_frame_off);
    c->n_kernels++;
    FA_CUDA(c, cudaGetLastError());
    *n_found = total;
    return FA_OK;
}
