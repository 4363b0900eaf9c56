// k_bin.cu — order-preserving screen-space binning of billboard rectangles, one launch.
//
// The reference hands every instanced quad to the fixed-function rasteriser
// (glDrawArraysInstanced, src/Shaders/VoxelizeShader.cpp:63 and
// src/Shaders/ConeTraceShader.cpp:75) which visits fragments in instance order.  The
// per-pixel kernels need the same thing turned inside out: for every 16x16-pixel tile, the
// list of billboards whose rectangle touches it, in pass order.
//
// One CTA of 512 threads owns a coarse tile (kCoarse x kCoarse fine tiles of 16x16 pixels, one fine
// tile per warp at kCoarse = 4: small coarse tiles keep the busiest CTA's serial work short); coarse tiles that
// miss the bounding rectangle of all billboards (k_prep_sort.cu) return at once.
//   sweep 1: stream all N rectangles (pass order) in chunks of 512, four loads in flight per
//            thread; keep the ones touching the coarse tile with an ordered block compaction (one
//            barrier per chunk) into a shared-memory buffer.  Whenever the buffer fills (and at the
//            end) the 16 warps count, each for its own fine tile(s), and the buffer is appended to the coarse pool (one atomicAdd per flush).
//   claim:   every fine tile claims its segment of the list pool with ONE atomicAdd.
//   sweep 2: the warps re-read only the CTA's own coarse segments (no barriers) and fill.
// No per-entry atomics, deterministic list content.  If a segment would not fit a pool the
// tile is left empty and the cursor still advances: the host sees cursor > capacity after the
// pass, grows the pool and re-runs the frame.
#include "crn_internal.cuh"

namespace crn {

namespace {

constexpr int kBinThreads = 512;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinCap = 2048;                       // shared-memory buffer entries
constexpr int kBinUnroll = 4;                       // rectangles loaded per thread before compacting
constexpr int kFinePerWarp = (kCoarse * kCoarse) / kBinWarps;   // fine tiles per warp
static_assert(kFinePerWarp >= 1 && kFinePerWarp * kBinWarps == kCoarse * kCoarse, "one warp owns a whole number of fine tiles");

__device__ __forceinline__ bool overlaps(BoardRect r, int x0, int y0, int x1, int y1) {
    return r.i1 >= r.i0 && r.i0 <= x1 && r.i1 >= x0 && r.j0 <= y1 && r.j1 >= y0;
}

struct BinArgs {
    const BoardRect *rects;
    int n;
    int tilesX, tilesY, coarseX, coarseY;
    uint32_t *tileOff, *tileCnt, *tileList;
    uint32_t tileCap;
    uint4 *coarsePool;            // {rect (2 words), pass-order index, -}
    uint32_t coarseCap;
    uint32_t *cursors;            // [0] coarse pool, [1] fine pool
    const int32_t *bounds;        // [minI, maxI, minJ, maxJ] over all unclipped rectangles
};

constexpr int kMaxSegs = 96;      // coarse segments one CTA can remember (N up to ~147k boards on one coarse tile)

__global__ void __launch_bounds__(kBinThreads) bin_kernel(BinArgs a) {
    __shared__ BoardRect sRect[kBinCap];
    __shared__ uint32_t sIdx[kBinCap];
    __shared__ uint32_t sWarp[2][kBinWarps];
    __shared__ uint32_t sSegBase[kMaxSegs], sSegCnt[kMaxSegs];
    __shared__ uint32_t sSegs, sBad;
    const int ct = blockIdx.x;
    const int cx = ct % a.coarseX, cy = ct / a.coarseX;
    const int span = kTile * kCoarse;
    const int X0 = cx * span, Y0 = cy * span, X1 = X0 + span - 1, Y1 = Y0 + span - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // this warp's fine tile(s)
    int tIdx[kFinePerWarp], tx0[kFinePerWarp], ty0[kFinePerWarp];
    bool tOk[kFinePerWarp];
#pragma unroll
    for (int k = 0; k < kFinePerWarp; k++) {
        const int f = warp * kFinePerWarp + k;
        const int fx = cx * kCoarse + (f % kCoarse), fy = cy * kCoarse + (f / kCoarse);
        tOk[k] = fx < a.tilesX && fy < a.tilesY;
        tIdx[k] = fy * a.tilesX + fx;
        tx0[k] = fx * kTile; ty0[k] = fy * kTile;
    }
    // nothing can touch this coarse tile: publish empty fine tiles and leave
    if (a.n == 0 || a.bounds[0] > X1 || a.bounds[1] < X0 || a.bounds[2] > Y1 || a.bounds[3] < Y0) {
        if (lane < kFinePerWarp && tOk[lane]) { a.tileOff[tIdx[lane]] = 0; a.tileCnt[tIdx[lane]] = 0; }
        return;
    }
    if (tid == 0) { sSegs = 0; sBad = 0; }
    uint32_t cnt[kFinePerWarp];
#pragma unroll
    for (int k = 0; k < kFinePerWarp; k++) cnt[k] = 0;
    __syncthreads();

    // ---- sweep 1: coarse compaction + fine counts + coarse segments
    uint32_t buf = 0, parity = 0;
    for (int chunk = 0; chunk < a.n; chunk += kBinUnroll * kBinThreads) {
        BoardRect rr[kBinUnroll];
#pragma unroll
        for (int k = 0; k < kBinUnroll; k++) {
            const int i = chunk + k * kBinThreads + tid;
            rr[k] = {0, -1, 0, -1};
            if (i < a.n) rr[k] = a.rects[i];
        }
#pragma unroll
        for (int k = 0; k < kBinUnroll; k++) {
            const int first = chunk + k * kBinThreads;
            if (first >= a.n) break;
            const int i = first + tid;
            const bool f = overlaps(rr[k], X0, Y0, X1, Y1);
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, f);
            uint32_t *wc = sWarp[parity];                          // double-buffered: one barrier per step
            parity ^= 1u;
            if (lane == 0) wc[warp] = __popc(bal);
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kBinWarps; w++) {
                const uint32_t c = wc[w];
                if (w < warp) before += c;
                total += c;
            }
            if (f) {
                const uint32_t p = buf + before + __popc(bal & ((1u << lane) - 1u));
                sRect[p] = rr[k]; sIdx[p] = (uint32_t)i;
            }
            buf += total;
            const bool last = first + kBinThreads >= a.n;
            if (buf > kBinCap - kBinThreads || (last && buf > 0)) {
                if (tid == 0) {                                    // claim a coarse segment for this flush
                    const uint32_t sb = atomicAdd(&a.cursors[0], buf);
                    const uint32_t k2 = sSegs;
                    if (k2 < kMaxSegs && (uint64_t)sb + buf <= a.coarseCap) { sSegBase[k2] = sb; sSegCnt[k2] = buf; sSegs = k2 + 1; }
                    else { sBad = 1; atomicOr(&a.cursors[2], k2 >= kMaxSegs ? 2u : 1u); }   // sticky until the host clears it; bit 1: segment table full (not fixable by growing the pool)
                }
                __syncthreads();                                   // buffer + segment visible
                const uint32_t segBase = sBad ? 0u : sSegBase[sSegs - 1];
                for (uint32_t e0 = 0; e0 < buf; e0 += 32) {
                    const uint32_t e = e0 + lane;
                    BoardRect q = {0, -1, 0, -1};
                    if (e < buf) q = sRect[e];
#pragma unroll
                    for (int t = 0; t < kFinePerWarp; t++) {
                        const bool g = tOk[t] && overlaps(q, tx0[t], ty0[t], tx0[t] + kTile - 1, ty0[t] + kTile - 1);
                        cnt[t] += __popc(__ballot_sync(0xFFFFFFFFu, g));
                    }
                }
                if (!sBad)
                    for (uint32_t e = tid; e < buf; e += kBinThreads) {
                        const BoardRect q = sRect[e];
                        a.coarsePool[segBase + e] = make_uint4(((uint32_t)(uint16_t)q.i0) | ((uint32_t)(uint16_t)q.i1 << 16),
                                                              ((uint32_t)(uint16_t)q.j0) | ((uint32_t)(uint16_t)q.j1 << 16), sIdx[e], 0u);
                    }
                buf = 0;
                __syncthreads();                                   // buffer free again, pool writes ordered for this CTA
            }
        }
    }

    // ---- claim one segment of the list pool per fine tile
    uint32_t base[kFinePerWarp];
    bool fits[kFinePerWarp];
    bool any = false;
    const bool bad = sBad != 0;
#pragma unroll
    for (int k = 0; k < kFinePerWarp; k++) {
        uint32_t b = 0;
        if (lane == 0 && tOk[k] && cnt[k]) b = atomicAdd(&a.cursors[1], cnt[k]);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        base[k] = b;
        fits[k] = !bad && (uint64_t)b + cnt[k] <= a.tileCap;
        if (lane == 0 && tOk[k] && cnt[k] && !fits[k]) atomicOr(&a.cursors[2], 1u);
        if (lane == 0 && tOk[k]) {
            a.tileOff[tIdx[k]] = b;
            a.tileCnt[tIdx[k]] = fits[k] ? cnt[k] : 0;
        }
        any = any || (cnt[k] > 0 && fits[k]);
    }
    if (!any) return;

    // ---- sweep 2: fill from the CTA's own coarse segments (each warp on its own, no barriers)
    uint32_t filled[kFinePerWarp];
#pragma unroll
    for (int k = 0; k < kFinePerWarp; k++) filled[k] = 0;
    const uint32_t nSegs = sSegs;
    for (uint32_t sg = 0; sg < nSegs; sg++) {
        const uint32_t sb = sSegBase[sg], sc = sSegCnt[sg];
        for (uint32_t e0 = 0; e0 < sc; e0 += 32) {
            const uint32_t e = e0 + lane;
            BoardRect q = {0, -1, 0, -1};
            uint32_t qi = 0;
            if (e < sc) {
                const uint4 v = a.coarsePool[sb + e];
                q.i0 = (int16_t)(v.x & 0xFFFFu); q.i1 = (int16_t)(v.x >> 16); q.j0 = (int16_t)(v.y & 0xFFFFu); q.j1 = (int16_t)(v.y >> 16);
                qi = v.z;
            }
#pragma unroll
            for (int t = 0; t < kFinePerWarp; t++) {
                const bool g = tOk[t] && overlaps(q, tx0[t], ty0[t], tx0[t] + kTile - 1, ty0[t] + kTile - 1);
                const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, g);
                if (g && fits[t]) a.tileList[base[t] + filled[t] + __popc(b2 & ((1u << lane) - 1u))] = qi;
                filled[t] += __popc(b2);
            }
        }
    }
}

// Longest-first launch order for the per-tile kernels: tiles are bucketed by log2 of their list length and
// emitted from the longest bucket down, so the CTAs that walk hundreds of layers start first and the tail of
// the grid is made of empty tiles (order inside a bucket is arbitrary; results do not depend on it).
// (tile rows of another rank of an interleaved trace, crn_set_tile_row_interleave, are left out: the trace grid only holds owned tiles)
__global__ void __launch_bounds__(1024) tile_hist_kernel(const uint32_t *__restrict__ tileCnt, int tiles, uint32_t *gHist, int tilesX, int ilvIndex, int ilvCount) {
    __shared__ uint32_t sHist[33];
    const int t = threadIdx.x, i = blockIdx.x * 1024 + t;
    if (t < 33) sHist[t] = 0;
    __syncthreads();
    if (i < tiles && (i / tilesX) % ilvCount == ilvIndex) atomicAdd(&sHist[32 - __clz(tileCnt[i])], 1u);            // bucket 0: empty tile
    __syncthreads();
    if (t < 33 && sHist[t]) atomicAdd(&gHist[t], sHist[t]);
}

__global__ void __launch_bounds__(1024) tile_scatter_kernel(const uint32_t *__restrict__ tileCnt, int tiles, const uint32_t *__restrict__ gHist,
                                                            uint32_t *gCur, uint32_t *order, int tilesX, int ilvIndex, int ilvCount) {
    __shared__ uint32_t sBase[33];
    const int t = threadIdx.x, i = blockIdx.x * 1024 + t;
    if (t == 0) {
        uint32_t run = 0;
        for (int b = 32; b >= 0; b--) { sBase[b] = run; run += gHist[b]; }
    }
    __syncthreads();
    // warp-aggregated claim: one atomic per (warp, bucket) instead of one per tile
    const int b = (i < tiles && (i / tilesX) % ilvCount == ilvIndex) ? 32 - __clz(tileCnt[i]) : 33;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b);
    const int lane = t & 31, leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader && b < 33) base = atomicAdd(&gCur[b], (uint32_t)__popc(peers));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (b < 33) order[sBase[b] + base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;
}

} // namespace

// scratch: 66 words (histogram + cursors) right after the order array
int launch_tile_order(cudaStream_t st, const Bins &b, uint32_t *order, int ilvIndex, int ilvCount) {
    const int tiles = b.tilesX * b.tilesY;
    uint32_t *gHist = order + tiles, *gCur = gHist + 33;
    cudaMemsetAsync(gHist, 0, 66 * sizeof(uint32_t), st);
    const int grid = (tiles + 1023) / 1024;
    tile_hist_kernel<<<grid, 1024, 0, st>>>(b.tileCnt, tiles, gHist, b.tilesX, ilvIndex, ilvCount);
    tile_scatter_kernel<<<grid, 1024, 0, st>>>(b.tileCnt, tiles, gHist, gCur, order, b.tilesX, ilvIndex, ilvCount);
    return 2;
}

int launch_bin(cudaStream_t st, const BoardRect *rects, const int32_t *bounds, int n, int W, int H, Bins &b) {
    BinArgs a;
    a.rects = rects; a.n = n;
    a.tilesX = b.tilesX; a.tilesY = b.tilesY; a.coarseX = b.coarseX; a.coarseY = b.coarseY;
    a.tileOff = b.tileOff; a.tileCnt = b.tileCnt; a.tileList = b.tileList; a.tileCap = (uint32_t)b.tileCap;
    a.coarsePool = reinterpret_cast<uint4 *>(b.coarseList); a.coarseCap = (uint32_t)(b.coarseCap / 4);
    a.cursors = b.cursors; a.bounds = bounds;
    cudaMemsetAsync(b.cursors, 0, 2 * sizeof(uint32_t), st);
    bin_kernel<<<b.coarseX * b.coarseY, kBinThreads, 0, st>>>(a);
    return 1;
}

} // namespace crn
