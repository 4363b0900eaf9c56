// k_prep_sort.cu — per-billboard set-up for both passes and the two orderings.
//
// Replaces, on the device:
//   * the billboard vertex stage's per-instance part (res/billboard_vert_instanced.glsl:20-37):
//     centre = volumePosition + boardPosition, view-space centre, and the window rectangle the
//     quad rasterises to under the light camera (pass 1) and the user camera (cone trace);
//   * CloudVolume::sortBoards (src/CloudVolume.cpp:65-82): back-to-front order by
//     distance(position + offset, cameraPosition).  The reference's O(N^2) CPU selection sort
//     becomes a bucketed rank sort: keys are spread over 2048 buckets between the frame's min
//     and max key, and an element's rank is its bucket's start plus the number of smaller keys
//     in its own bucket (64-bit compares, ~N^2/2048 of them).  Deterministic: same permutation
//     as a stable sort with ties broken by instance index;
//   * (new) a front-to-back order from the sun, by a conservative lower bound of the pass-1
//     depth, which lets the voxelize kernel stop walking a tile's list early.
//
// Compiled with -fmad=false: rectangles and keys are part of the exact-parity contract
// (tests compare them with the oracle bit for bit), so every operation is a single IEEE
// rounding in the written order.
#include "crn_internal.cuh"

namespace crn {

namespace {

__device__ __forceinline__ void mul_point(const float *m, float x, float y, float z, float out[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++) out[r] = ((m[0 + r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r];
}

// window rectangle of a view-facing quad (padded by one pixel; coverage is decided per pixel later)
__device__ __forceinline__ BoardRect quad_rect(const ViewParams &vp, const float cv[4], float scale) {
    BoardRect q;
    const float zv = cv[2];
    const float zc = vp.P[10] * zv + vp.P[14] * 1.0f;
    const float wc = vp.P[11] * zv + vp.P[15] * 1.0f;
    const bool clipped = !(zc >= -wc && zc <= wc) || !(wc > 0.0f);
    if (clipped) { q.i0 = q.j0 = 0; q.i1 = q.j1 = -1; return q; }
    const float W = (float)vp.W, H = (float)vp.H;
    float x0 = (vp.P[0] * (cv[0] - scale) + vp.P[12] * 1.0f) / wc;
    float x1 = (vp.P[0] * (cv[0] + scale) + vp.P[12] * 1.0f) / wc;
    float y0 = (vp.P[5] * (cv[1] - scale) + vp.P[13] * 1.0f) / wc;
    float y1 = (vp.P[5] * (cv[1] + scale) + vp.P[13] * 1.0f) / wc;
    float fx0 = (x0 + 1.0f) * 0.5f * W, fx1 = (x1 + 1.0f) * 0.5f * W;
    float fy0 = (y0 + 1.0f) * 0.5f * H, fy1 = (y1 + 1.0f) * 0.5f * H;
    fx0 = fminf(fmaxf(fx0, -2.0f), W + 2.0f); fx1 = fminf(fmaxf(fx1, -2.0f), W + 2.0f);
    fy0 = fminf(fmaxf(fy0, -2.0f), H + 2.0f); fy1 = fminf(fmaxf(fy1, -2.0f), H + 2.0f);
    q.i0 = (int16_t)max(0, (int)floorf(fx0) - 1); q.i1 = (int16_t)min(vp.W - 1, (int)floorf(fx1) + 1);
    q.j0 = (int16_t)max(0, (int)floorf(fy0) - 1); q.j1 = (int16_t)min(vp.H - 1, (int)floorf(fy1) + 1);
    return q;
}

__device__ __forceinline__ uint32_t sortable(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct PrepArgs {
    const float *pos, *scale;
    int n;
    float fluff;
    float volpos[3];
    float nearPlane[3];
    float clip;
    float camPos[3];
    uint64_t *keyL, *keyC;
    BoardRec *recL, *recC;
    BoardRect *rectL, *rectC;
    float *lb;
    uint32_t *range;          // [minL, maxL, minC, maxC] of the sortable key bits
    int32_t *bounds;          // [minI, maxI, minJ, maxJ] light, then camera: union of the unclipped rectangles
    NoiseLat lat;             // window of the combined-octave noise lattice (k_noiselat.cu): camera records carry an "inside" flag
    uint32_t *wbox;           // sortable bits of [min x, y, z of (centre - r) | max x, y, z of (centre + r)] over all billboards (camera pass)
};

__device__ __forceinline__ void warp_bounds(BoardRect q, bool valid, int32_t *b) {
    const bool ok = valid && q.i1 >= q.i0;
    const int lo_i = __reduce_min_sync(0xFFFFFFFFu, ok ? (int)q.i0 : 0x7FFFFFFF), hi_i = __reduce_max_sync(0xFFFFFFFFu, ok ? (int)q.i1 : -1);
    const int lo_j = __reduce_min_sync(0xFFFFFFFFu, ok ? (int)q.j0 : 0x7FFFFFFF), hi_j = __reduce_max_sync(0xFFFFFFFFu, ok ? (int)q.j1 : -1);
    if ((threadIdx.x & 31) == 0 && hi_i >= 0) { atomicMin(&b[0], lo_i); atomicMax(&b[1], hi_i); atomicMin(&b[2], lo_j); atomicMax(&b[3], hi_j); }
}

__device__ __forceinline__ void warp_minmax(uint32_t u, bool valid, uint32_t *range) {
    uint32_t lo = valid ? u : 0xFFFFFFFFu, hi = valid ? u : 0u;
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(&range[0], lo); atomicMax(&range[1], hi); }
}

// world-space bounding box of all spheres: every fragment's start position lies inside it (k_conebake.cu skips the rest)
__device__ __forceinline__ void warp_box(float cx, float cy, float cz, float r, bool valid, uint32_t *wbox) {
    const float c[3] = {cx, cy, cz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, valid ? sortable(c[k] - r) : 0xFFFFFFFFu);
        const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, valid ? sortable(c[k] + r) : 0u);
        if ((threadIdx.x & 31) == 0) { atomicMin(&wbox[k], lo); atomicMax(&wbox[3 + k], hi); }
    }
}

__global__ void __launch_bounds__(256) prep_kernel(PrepArgs a, ViewParams light, ViewParams cam) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i0 < a.n;
    const int i = live ? i0 : a.n - 1;                       // dead lanes redo the last board and store nothing
    const float ox = a.pos[3 * i], oy = a.pos[3 * i + 1], oz = a.pos[3 * i + 2];
    const float cx = a.volpos[0] + ox, cy = a.volpos[1] + oy, cz = a.volpos[2] + oz;
    const float s = a.scale[i];
    const float r = a.fluff != 1.0f ? s * a.fluff : s;       // CloudVolume::uploadBillboards, src/CloudVolume.cpp:153-161
    float cv[4];
    if (a.recL) {
        mul_point(light.V, cx, cy, cz, cv);
        BoardRec rec = {cx, cy, cz, r, cv[0], cv[1], cv[2], i};
        const BoardRect q = quad_rect(light, cv, r);
        if (live) { a.recL[i] = rec; a.rectL[i] = q; }
        warp_bounds(q, live, a.bounds);
        // conservative lower bound of first_voxelize's gl_FragDepth over the whole quad:
        // every surface point is within r of the centre.
        const float dx = a.nearPlane[0] - cx, dy = a.nearPlane[1] - cy, dz = a.nearPlane[2] - cz;
        const float dc = sqrtf((dx * dx + dy * dy) + dz * dz);
        const float lb = ((dc - r) / a.clip) * 0.9999f - 1e-5f;
        if (live) { a.lb[i] = lb; a.keyL[i] = ((uint64_t)sortable(lb) << 32) | (uint32_t)i; }
        warp_minmax(sortable(lb), live, a.range);
    }
    if (a.recC) {
        mul_point(cam.V, cx, cy, cz, cv);
        BoardRec rec = {cx, cy, cz, r, cv[0], cv[1], cv[2], i};
        // the whole noise march of this billboard (sphere stretched along the view ray by the march's overshoot) stays
        // inside the lattice window: the trace kernel reads the pre-summed octaves for it
        if (a.lat.on && r * a.lat.ext[0] + fabsf(cx - a.lat.winC[0]) <= a.lat.winH[0] && r * a.lat.ext[1] + fabsf(cy - a.lat.winC[1]) <= a.lat.winH[1] &&
            r * a.lat.ext[2] + fabsf(cz - a.lat.winC[2]) <= a.lat.winH[2])
            rec.idx |= kRecInLattice;
        const BoardRect q = quad_rect(cam, cv, r);
        if (live) { a.recC[i] = rec; a.rectC[i] = q; }
        warp_bounds(q, live, a.bounds + 4);
        // sortBoards key: glm::distance(position + offset, point) = length(point - p)
        const float dx = a.camPos[0] - cx, dy = a.camPos[1] - cy, dz = a.camPos[2] - cz;
        const float d = sqrtf((dx * dx + dy * dy) + dz * dz);
        // ascending composite = front to back; equal distances: later instance first, i.e. the
        // exact reverse of the draw order (far first, earlier instance first among equals)
        if (live) a.keyC[i] = ((uint64_t)sortable(d) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
        warp_minmax(sortable(d), live, a.range + 2);
        warp_box(cx, cy, cz, r, live, a.wbox);
    }
}

// ---- bucketed rank sort -------------------------------------------------------------------
constexpr int kBuckets = 2048;

__device__ __forceinline__ uint32_t bucket_of(uint64_t key, uint32_t lo, uint32_t hi) {
    const uint32_t u = (uint32_t)(key >> 32);
    const uint64_t span = (uint64_t)(hi - lo) + 1u;
    return (uint32_t)(((uint64_t)(u - lo) * kBuckets) / span);          // monotone in the key, < kBuckets
}

struct SortArgs {
    int n;
    const uint64_t *key;       // composite keys, instance order
    const uint32_t *range;     // [lo, hi]
    uint32_t *hist;            // kBuckets counters -> (after scan) bucket starts, kBuckets+1 entries
    uint32_t *slot;            // arrival slot of element i inside its bucket
    uint64_t *grouped;         // keys grouped by bucket (arbitrary order inside a bucket)
    uint32_t *rank;            // final position of element i
};

__global__ void __launch_bounds__(256) hist_kernel(SortArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    a.slot[i] = atomicAdd(&a.hist[bucket_of(a.key[i], a.range[0], a.range[1])], 1u);
}

__global__ void __launch_bounds__(1024) scan_kernel(SortArgs a) {       // exclusive scan of kBuckets counters, one CTA
    __shared__ uint32_t s[kBuckets];
    __shared__ uint32_t wsum[32];
    const int t = threadIdx.x;
    const uint32_t v0 = a.hist[2 * t], v1 = a.hist[2 * t + 1];
    uint32_t x = v0 + v1;
    const int lane = t & 31, w = t >> 5;
    uint32_t incl = x;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = wsum[lane], wi = ws;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += y; }
        wsum[lane] = wi - ws;
    }
    __syncthreads();
    const uint32_t excl = wsum[w] + incl - x;
    s[2 * t] = excl; s[2 * t + 1] = excl + v0;
    __syncthreads();
    a.hist[2 * t] = s[2 * t]; a.hist[2 * t + 1] = s[2 * t + 1];
    if (t == 1023) a.hist[kBuckets] = excl + x;
}

__global__ void __launch_bounds__(256) group_kernel(SortArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint64_t k = a.key[i];
    a.grouped[a.hist[bucket_of(k, a.range[0], a.range[1])] + a.slot[i]] = k;
}

// rank = bucket start + number of smaller keys in the same bucket (keys are unique: they embed the index)
__global__ void __launch_bounds__(256) bucket_rank_kernel(SortArgs a, int descendingIndex) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n) return;
    const uint64_t k = a.grouped[p];
    const uint32_t b = bucket_of(k, a.range[0], a.range[1]);
    const uint32_t s = a.hist[b], e = a.hist[b + 1];
    uint32_t r = 0;
    for (uint32_t q = s; q < e; q++) r += (a.grouped[q] < k);
    const uint32_t low = (uint32_t)k;
    a.rank[descendingIndex ? 0xFFFFFFFFu - low : low] = s + r;
}

struct ScatterArgs {
    int n;
    const uint32_t *rankL, *rankC;
    const BoardRec *recTmpL, *recTmpC;
    const BoardRect *rectTmpL, *rectTmpC;
    const float *lbTmp;
    BoardRec *recL, *recC;
    BoardRect *rectL, *rectC;
    float *lbSorted;
    int32_t *drawOrder;
};

__global__ void __launch_bounds__(256) scatter_kernel(ScatterArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.rankL) {
        const uint32_t k = a.rankL[i];
        a.recL[k] = a.recTmpL[i];
        a.rectL[k] = a.rectTmpL[i];
        a.lbSorted[k] = a.lbTmp[i];
    }
    if (a.rankC) {
        const uint32_t k = a.rankC[i];
        a.recC[k] = a.recTmpC[i];
        a.rectC[k] = a.rectTmpC[i];
        a.drawOrder[a.n - 1 - (int)k] = i;       // far -> near, what sortBoards leaves in place
    }
}

} // namespace

// scratch for one pass: [hist (kBuckets+64) | slot n] u32 in the first half, grouped n u64 in the second
size_t sort_tmp_bytes(int n) {
    const size_t half = ((size_t)(kBuckets + 64 + n) * 4 + 255) / 256 * 256;
    const size_t other = ((size_t)n * 8 + 255) / 256 * 256;
    return 2 * (half > other ? half : other);
}

const uint32_t *sort_tmp_world_box(const void *sortTmp, int n) {
    return reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(sortTmp) + sort_tmp_bytes(n)) + 12;
}

const int32_t *sort_tmp_bounds(const void *sortTmp, int n, int pass) {
    return reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(sortTmp) + sort_tmp_bytes(n)) + 4 + 4 * pass;
}

int launch_prep_sort(cudaStream_t st, const float *pos, const float *scale, int n, float fluff, const float volpos[3],
                     const ViewParams &light, const float nearPlane[3], float clip, const ViewParams &cam,
                     const float camPos[3], bool doLight, bool doCam, uint32_t *rankL, uint32_t *rankC,
                     uint64_t *keyL, uint64_t *keyC, BoardRec *recTmpL, BoardRec *recTmpC, BoardRect *rectTmpL,
                     BoardRect *rectTmpC, float *lbTmp, BoardRec *recL, BoardRec *recC, BoardRect *rectL,
                     BoardRect *rectC, float *lbSorted, int32_t *drawOrder, void *sortTmp, const NoiseLat *lat) {
    if (n <= 0) return 0;
    PrepArgs pa;
    if (lat) pa.lat = *lat; else pa.lat.on = 0;
    pa.pos = pos; pa.scale = scale; pa.n = n; pa.fluff = fluff;
    for (int k = 0; k < 3; k++) { pa.volpos[k] = volpos[k]; pa.nearPlane[k] = nearPlane[k]; pa.camPos[k] = camPos[k]; }
    pa.clip = clip;
    pa.keyL = doLight ? keyL : nullptr; pa.keyC = doCam ? keyC : nullptr;
    pa.recL = doLight ? recTmpL : nullptr; pa.recC = doCam ? recTmpC : nullptr;
    pa.rectL = rectTmpL; pa.rectC = rectTmpC; pa.lb = lbTmp;
    uint32_t *range = reinterpret_cast<uint32_t *>(sortTmp) + sort_tmp_bytes(n) / 4;      // 4 words after the scratch
    pa.range = range;
    pa.bounds = reinterpret_cast<int32_t *>(range + 4);
    pa.wbox = range + 12;
    static const uint32_t kRangeInit[18] = {0xFFFFFFFFu, 0u, 0xFFFFFFFFu, 0u, 0x7FFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu, 0xFFFFFFFFu,
                                            0x7FFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu, 0xFFFFFFFFu,
                                            0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    cudaMemcpyAsync(range, kRangeInit, sizeof kRangeInit, cudaMemcpyHostToDevice, st);
    const int blocks = (n + 255) / 256;
    prep_kernel<<<blocks, 256, 0, st>>>(pa, light, cam);

    // bucketed rank sort per pass (scratch: hist[kBuckets+1] | slot[n] | grouped[n] live in sortTmp)
    auto sort_pass = [&](const uint64_t *key, const uint32_t *range, uint32_t *rank, char *tmp, int descendingIndex) {
        SortArgs sa2;
        sa2.n = n; sa2.key = key; sa2.range = range; sa2.rank = rank;
        sa2.hist = reinterpret_cast<uint32_t *>(tmp);
        sa2.slot = sa2.hist + (kBuckets + 64);
        sa2.grouped = reinterpret_cast<uint64_t *>(tmp + sort_tmp_bytes(n) / 2);
        cudaMemsetAsync(sa2.hist, 0, sizeof(uint32_t) * (kBuckets + 1), st);
        hist_kernel<<<blocks, 256, 0, st>>>(sa2);
        scan_kernel<<<1, 1024, 0, st>>>(sa2);
        group_kernel<<<blocks, 256, 0, st>>>(sa2);
        bucket_rank_kernel<<<blocks, 256, 0, st>>>(sa2, descendingIndex);
    };
    int launches = 2;
    if (doLight) { sort_pass(keyL, range, rankL, (char *)sortTmp, 0); launches += 4; }
    if (doCam) { sort_pass(keyC, range + 2, rankC, (char *)sortTmp, 1); launches += 4; }
    ScatterArgs sa;
    sa.n = n;
    sa.rankL = doLight ? rankL : nullptr; sa.rankC = doCam ? rankC : nullptr;
    sa.recTmpL = recTmpL; sa.recTmpC = recTmpC; sa.rectTmpL = rectTmpL; sa.rectTmpC = rectTmpC; sa.lbTmp = lbTmp;
    sa.recL = recL; sa.recC = recC; sa.rectL = rectL; sa.rectC = rectC; sa.lbSorted = lbSorted; sa.drawOrder = drawOrder;
    scatter_kernel<<<blocks, 256, 0, st>>>(sa);
    return launches;
}

} // namespace crn
