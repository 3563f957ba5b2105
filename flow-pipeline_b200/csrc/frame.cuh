// frame.cuh -- record-boundary discovery on the GPU for submits that carry no
// offsets (a length-delimited stream as Clickhouse's Kafka engine reads it,
// compose/clickhouse/create.sh:28-34; framing of mocker/mocker.go:98-101).
//
// Finding the boundaries of varint(len)||message records is a sequential
// dependency (the next boundary is only known once the previous length is
// read).  The stream is cut into fixed tiles and solved by speculate-and-verify:
//
//   1. speculate  one warp per tile: 32 lanes start walking at 32 consecutive byte
//                 positions two tiles upstream.  A walk that ever lands on a true
//                 boundary stays on the true chain, so lanes that have
//                 synchronised agree on where they cross into the tile; the
//                 plurality answer (__match_any_sync) is the tile's guessed
//                 entry.  Tile 0 enters at 0.
//   2. walk       every dirty tile walks from its entry to its end, producing its
//                 exit (= the next tile's true entry IF its own entry was true)
//                 and its record count.
//   3. verify     entry[i] != exit[i-1] -> entry[i] = exit[i-1], tile i dirty.
//                 Repeat 2-3 until no tile is dirty.  By induction from tile 0
//                 every entry is then exact, whatever the guesses were: the
//                 speculation only decides how many rounds it takes (typically
//                 one; adversarial input degrades to one tile per round).
//   4. emit       exclusive scan of the counts (cub) and a last walk that writes
//                 offsets[].
//
// A stream that ends inside a record yields one final span [start, len) which
// the decoder then rejects and counts as a bad record (inserter.go:125-126).
#pragma once

namespace fa {

constexpr uint32_t kFrameTile = 4096;
constexpr uint32_t kFrameWarm = 2 * kFrameTile;

// Read the record header at pos: returns the position of the next record, or
// `len` if the header is malformed / the record runs past the end.
__device__ __forceinline__ unsigned long long frame_next(const uint8_t *buf, unsigned long long pos, unsigned long long len)
{
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

__global__ void k_frame_speculate(const uint8_t *buf, unsigned long long len, uint32_t n_tiles, uint32_t *entry, uint8_t *dirty)
{
    const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    const unsigned long long start = (unsigned long long)tile * kFrameTile;
    unsigned long long pos;
    if (start <= kFrameWarm) pos = 0;  // close to the head: walk the true chain
    else pos = start - kFrameWarm + lane;
    while (pos < start) pos = frame_next(buf, pos, len);
    const uint32_t e = (uint32_t)pos;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, e);
    const uint32_t votes = __popc(peers);
    // plurality; ties -> smallest position
    uint32_t best_votes = votes, best_e = e;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const uint32_t ov = __shfl_xor_sync(0xFFFFFFFFu, best_votes, d);
        const uint32_t oe = __shfl_xor_sync(0xFFFFFFFFu, best_e, d);
        if (ov > best_votes || (ov == best_votes && oe < best_e)) {
            best_votes = ov;
            best_e = oe;
        }
    }
    if (lane == 0) {
        entry[tile] = best_e;
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
    for (uint32_t round = 0;; round++) {
        k_frame_walk<<<g, tpb, 0, c->stream>>>(d_buf, len, n_tiles, entry, exit_pos, count, dirty);
        c->n_kernels++;
        FA_CUDA(c, cudaGetLastError());
        FA_CUDA(c, cudaMemsetAsync(n_dirty, 0, 4, c->stream));
        k_frame_verify<<<g, tpb, 0, c->stream>>>(n_tiles, entry, exit_pos, dirty, n_dirty);
        c->n_kernels++;
        FA_CUDA(c, cudaGetLastError());
        uint32_t nd = 0;
        FA_CUDA(c, cudaMemcpyAsync(&nd, n_dirty, 4, cudaMemcpyDeviceToHost, c->stream));
        FA_CUDA(c, cudaStreamSynchronize(c->stream));
        if (nd == 0) break;
        if (round > n_tiles + 1) {
            c->last_error = "frame index did not converge";
            return FA_ERR_FRAMING;
        }
    }
    FA_CUDA(c, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, count, base, (int)n_tiles, c->stream));
    uint32_t last_base = 0, last_count = 0;
    FA_CUDA(c, cudaMemcpyAsync(&last_base, base + n_tiles - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaMemcpyAsync(&last_count, count + n_tiles - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    FA_CUDA(c, cudaStreamSynchronize(c->stream));
    const uint32_t total = last_base + last_count;
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
