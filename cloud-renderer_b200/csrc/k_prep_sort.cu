// k_prep_sort.cu — per-billboard set-up for both passes and the two orderings.
//
// Replaces, on the device:
//   * the billboard vertex stage's per-instance part (res/billboard_vert_instanced.glsl:20-37):
//     centre = volumePosition + boardPosition, view-space centre, and the window rectangle the
//     quad rasterises to under the light camera (pass 1) and the user camera (cone trace);
//   * CloudVolume::sortBoards (src/CloudVolume.cpp:65-82): back-to-front order by
//     distance(position + offset, cameraPosition).  The reference's O(N^2) CPU selection sort
//     becomes a counting rank sort spread over the whole GPU (one 64-bit compare per pair);
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
};

__global__ void __launch_bounds__(256) prep_kernel(PrepArgs a, ViewParams light, ViewParams cam) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float ox = a.pos[3 * i], oy = a.pos[3 * i + 1], oz = a.pos[3 * i + 2];
    const float cx = a.volpos[0] + ox, cy = a.volpos[1] + oy, cz = a.volpos[2] + oz;
    const float s = a.scale[i];
    const float r = a.fluff != 1.0f ? s * a.fluff : s;       // CloudVolume::uploadBillboards, src/CloudVolume.cpp:153-161
    float cv[4];
    if (a.recL) {
        mul_point(light.V, cx, cy, cz, cv);
        BoardRec rec = {cx, cy, cz, r, cv[0], cv[1], cv[2], i};
        a.recL[i] = rec;
        a.rectL[i] = quad_rect(light, cv, r);
        // conservative lower bound of first_voxelize's gl_FragDepth over the whole quad:
        // every surface point is within r of the centre.
        const float dx = a.nearPlane[0] - cx, dy = a.nearPlane[1] - cy, dz = a.nearPlane[2] - cz;
        const float dc = sqrtf((dx * dx + dy * dy) + dz * dz);
        const float lb = ((dc - r) / a.clip) * 0.9999f - 1e-5f;
        a.lb[i] = lb;
        a.keyL[i] = ((uint64_t)sortable(lb) << 32) | (uint32_t)i;
    }
    if (a.recC) {
        mul_point(cam.V, cx, cy, cz, cv);
        BoardRec rec = {cx, cy, cz, r, cv[0], cv[1], cv[2], i};
        a.recC[i] = rec;
        a.rectC[i] = quad_rect(cam, cv, r);
        // sortBoards key: glm::distance(position + offset, point) = length(point - p)
        const float dx = a.camPos[0] - cx, dy = a.camPos[1] - cy, dz = a.camPos[2] - cz;
        const float d = sqrtf((dx * dx + dy * dy) + dz * dz);
        // ascending composite = front to back; equal distances: later instance first, i.e. the
        // exact reverse of the draw order (far first, earlier instance first among equals)
        a.keyC[i] = ((uint64_t)sortable(d) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
    }
}

// rank[i] += #{ j in this CTA's j-range : key[j] < key[i] }   (keys are unique)
constexpr int kRankThreads = 256;
constexpr int kRankChunk = 1024;
__global__ void __launch_bounds__(kRankThreads) rank_kernel(const uint64_t *__restrict__ keyL, const uint64_t *__restrict__ keyC,
                                                            int n, int jPerSplit, uint32_t *rankL, uint32_t *rankC) {
    __shared__ uint64_t sL[kRankChunk];
    __shared__ uint64_t sC[kRankChunk];
    const int i = blockIdx.x * kRankThreads + threadIdx.x;
    const uint64_t myL = (keyL && i < n) ? keyL[i] : 0, myC = (keyC && i < n) ? keyC[i] : 0;
    const int jBeg = blockIdx.y * jPerSplit, jEnd = min(n, jBeg + jPerSplit);
    uint32_t cL = 0, cC = 0;
    for (int base = jBeg; base < jEnd; base += kRankChunk) {
        const int m = min(kRankChunk, jEnd - base);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += kRankThreads) {
            if (keyL) sL[t] = keyL[base + t];
            if (keyC) sC[t] = keyC[base + t];
        }
        __syncthreads();
        if (keyL) {
#pragma unroll 8
            for (int t = 0; t < m; t++) cL += (sL[t] < myL);
        }
        if (keyC) {
#pragma unroll 8
            for (int t = 0; t < m; t++) cC += (sC[t] < myC);
        }
    }
    if (i < n) {
        if (keyL && cL) atomicAdd(&rankL[i], cL);
        if (keyC && cC) atomicAdd(&rankC[i], cC);
    }
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

int launch_prep_sort(cudaStream_t st, const float *pos, const float *scale, int n, float fluff, const float volpos[3],
                     const ViewParams &light, const float nearPlane[3], float clip, const ViewParams &cam,
                     const float camPos[3], bool doLight, bool doCam, uint32_t *rankL, uint32_t *rankC,
                     uint64_t *keyL, uint64_t *keyC, BoardRec *recTmpL, BoardRec *recTmpC, BoardRect *rectTmpL,
                     BoardRect *rectTmpC, float *lbTmp, BoardRec *recL, BoardRec *recC, BoardRect *rectL,
                     BoardRect *rectC, float *lbSorted, int32_t *drawOrder) {
    if (n <= 0) return 0;
    PrepArgs pa;
    pa.pos = pos; pa.scale = scale; pa.n = n; pa.fluff = fluff;
    for (int k = 0; k < 3; k++) { pa.volpos[k] = volpos[k]; pa.nearPlane[k] = nearPlane[k]; pa.camPos[k] = camPos[k]; }
    pa.clip = clip;
    pa.keyL = doLight ? keyL : nullptr; pa.keyC = doCam ? keyC : nullptr;
    pa.recL = doLight ? recTmpL : nullptr; pa.recC = doCam ? recTmpC : nullptr;
    pa.rectL = rectTmpL; pa.rectC = rectTmpC; pa.lb = lbTmp;
    const int blocks = (n + 255) / 256;
    prep_kernel<<<blocks, 256, 0, st>>>(pa, light, cam);

    if (doLight) cudaMemsetAsync(rankL, 0, sizeof(uint32_t) * n, st);
    if (doCam) cudaMemsetAsync(rankC, 0, sizeof(uint32_t) * n, st);
    // split the j range so that the grid fills the machine (148 SMs x 8 CTAs of 256 threads)
    int splits = max(1, min((n + kRankChunk - 1) / kRankChunk, (148 * 8 + blocks - 1) / blocks));
    int jPerSplit = ((n + splits - 1) / splits + kRankChunk - 1) / kRankChunk * kRankChunk;
    splits = (n + jPerSplit - 1) / jPerSplit;
    rank_kernel<<<dim3(blocks, splits), kRankThreads, 0, st>>>(doLight ? keyL : nullptr, doCam ? keyC : nullptr, n, jPerSplit,
                                                              rankL, rankC);
    ScatterArgs sa;
    sa.n = n;
    sa.rankL = doLight ? rankL : nullptr; sa.rankC = doCam ? rankC : nullptr;
    sa.recTmpL = recTmpL; sa.recTmpC = recTmpC; sa.rectTmpL = rectTmpL; sa.rectTmpC = rectTmpC; sa.lbTmp = lbTmp;
    sa.recL = recL; sa.recC = recC; sa.rectL = rectL; sa.rectC = rectC; sa.lbSorted = lbSorted; sa.drawOrder = drawOrder;
    scatter_kernel<<<blocks, 256, 0, st>>>(sa);
    return 3;
}

} // namespace crn
