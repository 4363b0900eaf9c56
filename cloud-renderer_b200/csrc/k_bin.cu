// k_bin.cu — order-preserving two-level screen-space binning of billboard rectangles.
//
// The reference hands every instanced quad to the fixed-function rasteriser
// (glDrawArraysInstanced, src/Shaders/VoxelizeShader.cpp:63 and
// src/Shaders/ConeTraceShader.cpp:75) which visits fragments in instance order.  The
// per-pixel kernels need the same thing turned inside out: for every 16x16-pixel tile, the
// list of billboards whose rectangle touches it, in pass order.  Both levels are
// tile-centric ordered compactions (count, claim a segment with one atomicAdd, fill), so a
// tile's list content is deterministic and no per-entry atomics exist.
//
//   coarse: one CTA per 128x128-pixel tile scans all N rectangles;
//   fine:   one warp per 16x16-pixel tile scans its coarse tile's list.
//
// If a segment would not fit its pool the tile is left empty and the cursor still advances:
// the host sees cursor > capacity after the pass, grows the pool and re-runs the frame.
#include "crn_internal.cuh"

namespace crn {

namespace {

__device__ __forceinline__ bool overlaps(BoardRect r, int x0, int y0, int x1, int y1) {
    return r.i1 >= r.i0 && r.i0 <= x1 && r.i1 >= x0 && r.j0 <= y1 && r.j1 >= y0;
}

struct BinArgs {
    const BoardRect *rects;
    int n;
    int tilesX, tilesY, coarseX, coarseY;
    uint32_t *coarseOff, *coarseCnt, *coarseList;
    uint32_t coarseCap;
    uint32_t *tileOff, *tileCnt, *tileList;
    uint32_t tileCap;
    uint32_t *cursors;
};

__global__ void __launch_bounds__(256) bin_coarse_kernel(BinArgs a) {
    __shared__ uint32_t sBase;
    __shared__ uint32_t sWarp[8];
    const int ct = blockIdx.x;
    const int cx = ct % a.coarseX, cy = ct / a.coarseX;
    const int span = kTile * kCoarse;
    const int x0 = cx * span, y0 = cy * span, x1 = x0 + span - 1, y1 = y0 + span - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t cnt = 0;
    for (int base = 0; base < a.n; base += 256) {
        const int i = base + tid;
        const bool f = i < a.n && overlaps(a.rects[i], x0, y0, x1, y1);
        cnt += __syncthreads_count(f);
    }
    if (tid == 0) {
        uint32_t b = cnt ? atomicAdd(&a.cursors[0], cnt) : 0;
        const bool fits = (uint64_t)b + cnt <= a.coarseCap;
        sBase = b;
        a.coarseOff[ct] = b;
        a.coarseCnt[ct] = fits ? cnt : 0;
        if (!fits) cnt = 0;
        sWarp[0] = cnt;                      // reuse as the broadcast of "anything to fill"
    }
    __syncthreads();
    if (sWarp[0] == 0) return;
    const uint32_t segBase = sBase;
    __syncthreads();

    uint32_t running = 0;
    for (int base = 0; base < a.n; base += 256) {
        const int i = base + tid;
        const bool f = i < a.n && overlaps(a.rects[i], x0, y0, x1, y1);
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, f);
        if (lane == 0) sWarp[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint32_t c = sWarp[w];
            if (w < warp) before += c;
            total += c;
        }
        if (f) a.coarseList[segBase + running + before + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)i;
        running += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bin_fine_kernel(BinArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * 8 + warp;
    if (tile >= a.tilesX * a.tilesY) return;
    const int tx = tile % a.tilesX, ty = tile / a.tilesX;
    const int coarse = (ty / kCoarse) * a.coarseX + tx / kCoarse;
    const uint32_t off = a.coarseOff[coarse], cnt = a.coarseCnt[coarse];
    const int x0 = tx * kTile, y0 = ty * kTile, x1 = x0 + kTile - 1, y1 = y0 + kTile - 1;

    uint32_t c = 0;
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t e = base + lane;
        const bool f = e < cnt && overlaps(a.rects[a.coarseList[off + e]], x0, y0, x1, y1);
        c += __popc(__ballot_sync(0xFFFFFFFFu, f));
    }
    uint32_t segBase = 0;
    if (lane == 0 && c) segBase = atomicAdd(&a.cursors[1], c);
    segBase = __shfl_sync(0xFFFFFFFFu, segBase, 0);
    const bool fits = (uint64_t)segBase + c <= a.tileCap;
    if (lane == 0) {
        a.tileOff[tile] = segBase;
        a.tileCnt[tile] = fits ? c : 0;
    }
    if (!fits || c == 0) return;
    uint32_t running = 0;
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t e = base + lane;
        uint32_t k = 0;
        bool f = false;
        if (e < cnt) { k = a.coarseList[off + e]; f = overlaps(a.rects[k], x0, y0, x1, y1); }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, f);
        if (f) a.tileList[segBase + running + __popc(bal & ((1u << lane) - 1u))] = k;
        running += __popc(bal);
    }
}

} // namespace

int launch_bin(cudaStream_t st, const BoardRect *rects, int n, int W, int H, Bins &b) {
    BinArgs a;
    a.rects = rects; a.n = n;
    a.tilesX = b.tilesX; a.tilesY = b.tilesY; a.coarseX = b.coarseX; a.coarseY = b.coarseY;
    a.coarseOff = b.coarseOff; a.coarseCnt = b.coarseCnt; a.coarseList = b.coarseList; a.coarseCap = (uint32_t)b.coarseCap;
    a.tileOff = b.tileOff; a.tileCnt = b.tileCnt; a.tileList = b.tileList; a.tileCap = (uint32_t)b.tileCap;
    a.cursors = b.cursors;
    cudaMemsetAsync(b.cursors, 0, 2 * sizeof(uint32_t), st);
    bin_coarse_kernel<<<b.coarseX * b.coarseY, 256, 0, st>>>(a);
    const int tiles = b.tilesX * b.tilesY;
    bin_fine_kernel<<<(tiles + 7) / 8, 256, 0, st>>>(a);
    return 2;
}

} // namespace crn
