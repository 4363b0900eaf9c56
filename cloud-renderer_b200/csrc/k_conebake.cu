// k_conebake.cu — per-frame acceleration data for traceCone (res/conetrace_frag.glsl:64-79).
//
// traceCone's per-step height and LOD are the same for every fragment of a frame, so two things can be
// prepared once per (volume, cone parameters, light position) instead of per fragment — both EXACT:
//
// (1) Baked cone steps.  A step samples textureLod(volume, p, lod) = (1-f) * tri_L(p) + f * tri_{L+1}(p) with a
//     fixed (L, f).  tri_L is piecewise trilinear with its kinks on the planes through the texel centres of level L,
//     (c + 1/2) * 2^L voxels; those of level L+1 lie at (c + 1/2) * 2^(L+1).  Both sets are contained in the lattice
//     k * 2^(L-1), so the blend is trilinear inside every lattice cell and linear interpolation of its NODE values
//     reproduces it exactly.  One kernel evaluates the blend at the nodes (exact rational weights 0, 1/4, 1/2, 3/4, 1
//     from the linear chain) into a layered RG16 texture whose layer k holds node plane k and the step to plane k+1: the
//     trace kernel then needs ONE bilinear pass of the texture unit + one FMA per step instead of the four passes of a
//     mip-linear 3D fetch (measured: 288 G/s for tex3DLod at a fractional LOD, 1121 G/s for RG16 layered bilinear).
//     SNORM16 storage: 1.5e-5 absolute, far below the 8-bit filter weights of the texture unit.
//
// (2) Need codes.  Only the sun-facing shell of the cloud is lit, so most of the fine (textureLod) cone steps read nothing
//     but zero texels and contribute exactly 0.  The sample point of step i, pos + h_i * normalize(light - pos) / dim, is a
//     function of the start position only, so a coarse grid over the volume (cells of 2 voxels) stores, per cell, one bit
//     per group of consecutive steps: can a step of the group, started inside the cell, read a non-zero texel?  The test is
//     the footprint itself (cell box pushed along the cone, widened by the variation of the light direction over the cell,
//     converted to the 2x2x2-per-sample texel range) against the non-zero bits of the sampled level(s) (k_skipmask.cu).
//     A fragment reads one byte.  Conservative by construction, so still exact.
#include "crn_internal.cuh"

#include <algorithm>

namespace crn {

namespace {

// textures that share one lattice (same lower level) are evaluated together: the two trilinear samples are the same,
// only the blend fraction differs
struct BakeGroup {
    int level0, n, count;
    int NL, NU;                              // texels per axis of the lower / upper level (NU = 0: no upper level needed)
    uint32_t offL, offU;                     // byte offsets of the two levels in the chain
    float frac[kMaxBakedTex];
    cudaSurfaceObject_t surf[kMaxBakedTex];
};

struct BakeArgs {
    const uint32_t *bits;
    const uint8_t *chain;
    int nGroups;
    BakeGroup g[kMaxBakedTex];
};

// one axis of the trilinear footprint of a lattice node: the lattice spacing is half a texel of the lower level
// (shift 1) and a quarter of the upper one (shift 2), so the weights are exact multiples of 1/4
struct NodeAxis {
    int i0, i1;
    float w;
};

__device__ __forceinline__ NodeAxis node_axis(int k, int shift, int N) {
    NodeAxis a;
    const int i = (k - shift) >> shift;                              // floor(k / 2^shift - 1/2)
    a.w = shift == 1 ? ((k & 1) ? 0.0f : 0.5f) : (float)((k + 2) & 3) * 0.25f;
    a.i0 = min(max(i, 0), N - 1);                                    // CLAMP_TO_EDGE
    a.i1 = min(max(i + 1, 0), N - 1);
    return a;
}

// kFmt: 0 = R8 chain level (>= 1), 1 = R32F chain level (>= 1), 2 = the 1-bit level 0
// the bilinear (x, y) part of a node's trilinear sample on texel plane z: four loads
template <int kFmt>
__device__ __forceinline__ float plane_sample(const uint32_t *__restrict__ bits, const uint8_t *__restrict__ lvl, int N, const NodeAxis &X, const NodeAxis &Y, int z) {
    float c[4];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const uint32_t row = (uint32_t)(z * N + (q ? Y.i1 : Y.i0));
        if (kFmt == 2) {
            const uint32_t *r = bits + row * (uint32_t)(N >> 5);
            c[2 * q] = (float)((__ldg(r + (X.i0 >> 5)) >> (X.i0 & 31)) & 1u);
            c[2 * q + 1] = (float)((__ldg(r + (X.i1 >> 5)) >> (X.i1 & 31)) & 1u);
        } else if (kFmt == 1) {
            const float *r = reinterpret_cast<const float *>(lvl) + row * (uint32_t)N;
            c[2 * q] = __ldg(r + X.i0); c[2 * q + 1] = __ldg(r + X.i1);
        } else {
            const uint8_t *r = lvl + row * (uint32_t)N;
            c[2 * q] = (float)__ldg(r + X.i0); c[2 * q + 1] = (float)__ldg(r + X.i1);
        }
    }
    const float x0 = fmaf(X.w, c[1] - c[0], c[0]), x1 = fmaf(X.w, c[3] - c[2], c[2]);
    return fmaf(Y.w, x1 - x0, x0);
}

// A column of nodes (x, y, k0..) of one level: consecutive nodes share their texel planes (a new plane every 2 nodes of the
// lower level, every 4 of the upper one), so the two planes of the current node are kept and only a new one is fetched.
template <int kFmt>
struct ColumnSampler {
    const uint32_t *bits;
    const uint8_t *lvl;
    int N, shift;
    NodeAxis X, Y;
    int z0 = -1, z1 = -1;
    float p0 = 0.0f, p1 = 0.0f;
    __device__ __forceinline__ float at(int k) {
        const NodeAxis Z = node_axis(k, shift, N);
        if (Z.i0 != z0) {
            if (Z.i0 == z1) { z0 = z1; p0 = p1; }
            else { z0 = Z.i0; p0 = plane_sample<kFmt>(bits, lvl, N, X, Y, z0); }
        }
        if (Z.i1 != z1) {
            z1 = Z.i1;
            p1 = z1 == z0 ? p0 : plane_sample<kFmt>(bits, lvl, N, X, Y, z1);
        }
        const float v = fmaf(Z.w, p1 - p0, p0);
        return kFmt == 0 ? v * (1.0f / 255.0f) : v;
    }
};

// One thread per (x, y) and run of kBakeRun layers (kBakeRun + 1 node evaluations for kBakeRun texels, ~4 loads per
// layer).  Layer k holds (B(x,y,k), B(x,y,k+1) - B(x,y,k)) as SNORM16 codes: the trace kernel's z blend is one FMA, and
// because the step is the difference of the quantised planes, layer k + its step is exactly layer k+1.
constexpr int kBakeRun = 16;
template <bool kF32>
__global__ void __launch_bounds__(256) bake_steps_kernel(const __grid_constant__ BakeArgs a) {
    const BakeGroup &g = a.g[blockIdx.z];
    const int n = g.n;
    const int runs = (n - 1 + kBakeRun - 1) / kBakeRun;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yr = blockIdx.y * 8 + (threadIdx.x >> 5);              // y + n * run
    if (x >= n || yr >= n * runs) return;
    const int y = yr % n, k0 = (yr / n) * kBakeRun;
    ColumnSampler<2> lo0{a.bits, nullptr, g.NL, 1, node_axis(x, 1, g.NL), node_axis(y, 1, g.NL)};
    ColumnSampler<kF32 ? 1 : 0> lo{nullptr, a.chain + g.offL, g.NL, 1, node_axis(x, 1, g.NL), node_axis(y, 1, g.NL)};
    ColumnSampler<kF32 ? 1 : 0> up{nullptr, a.chain + g.offU, max(g.NU, 1), 2, node_axis(x, 2, max(g.NU, 1)), node_axis(y, 2, max(g.NU, 1))};
    int prev[kMaxBakedTex];
    for (int k = k0; k <= min(k0 + kBakeRun, n - 1); k++) {
        const float vlo = g.level0 == 0 ? lo0.at(k) : lo.at(k);
        const float vhi = g.NU ? up.at(k) : 0.0f;
#pragma unroll
        for (int t = 0; t < kMaxBakedTex; t++) {
            if (t < g.count) {
                const float v = g.frac[t] > 0.0f ? fmaf(g.frac[t], vhi - vlo, vlo) : vlo;
                const int q = (int)__float2uint_rn(__saturatef(v) * 32767.0f);
                if (k > k0) surf2DLayeredwrite(make_short2((short)prev[t], (short)(q - prev[t])), g.surf[t], x * 4, y, k - 1);
                prev[t] = q;
            }
        }
    }
}

struct CodeGroup {
    float hMin, hMax;                         // heights of the group's first and last step / dim: |P - pos| in normalized coordinates
    int lv0, nLv;                             // mip levels its steps sample: lv0 (and lv0 + 1 when nLv == 2)
};

struct CodeLevel {
    int size, wpr;                            // texels per axis, words per row of the non-zero bit volume
    uint32_t off;                             // word offset in `nz` (level 0: the occupancy set itself)
    float sizeF;
};

struct CodeArgs {
    int G;
    int nGroups;
    int coarse;                               // level of the early-out test
    uint32_t fineSet;                         // groups it covers: all levels <= coarse and at most three levels below it
    float fineMin, fineMax;                   // height range of those groups / dim
    CodeGroup g[kCodeGroups];
    CodeLevel lv[kMaxLevels];
    const uint32_t *bits;                     // level 0: one bit per voxel
    const uint32_t *nz;                       // levels >= 1: one bit per non-zero texel (k_skipmask.cu)
    const uint32_t *worldBox;                 // sortable bits of the billboards' world bounding box (k_prep_sort.cu), or nullptr
    float lightPos[3], b0[3], range[3];
    uint8_t *code;
    cudaSurfaceObject_t codeSurf;             // the same grid as skip bits (~code) in a 3D array, for the fast trace variant's texture lookup
};

__device__ __forceinline__ float unsortable(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

struct TexelBox {
    int lo[3], hi[3];
};

// texels of a level a LINEAR, CLAMP_TO_EDGE lookup can read when its normalized coordinate lies in [pmin, pmax]:
// floor(p * n - 1/2) and the one after, with 1/64 texel of slack for the texture unit's fixed-point coordinates
__device__ __forceinline__ TexelBox footprint(const float pmin[3], const float pmax[3], const CodeLevel &L) {
    TexelBox b;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        b.lo[k] = min(max(__float2int_rd(fmaf(pmin[k], L.sizeF, -0.515625f)), 0), L.size - 1);
        b.hi[k] = min(max(__float2int_rd(fmaf(pmax[k], L.sizeF, -0.484375f)) + 1, 0), L.size - 1);
    }
    return b;
}

// Warp-cooperative test of one level: every lane has its own texel box (live lanes only); the (y, z) rows are taken over the
// union of the lanes' ranges (a superset: still conservative), each lane loads one row word per word column, one OR
// reduction per column, then every lane looks at its own x range in the reduced words.  The 32 cells of a warp are
// neighbours in x, so their boxes overlap almost entirely: ~3 loads per lane instead of ~25.
__device__ __forceinline__ bool warp_any_set(const uint32_t *__restrict__ m, const CodeLevel &L, const TexelBox &b, bool live) {
    const uint32_t full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int ylo = __reduce_min_sync(full, live ? b.lo[1] : 0x7FFFFFFF), yhi = __reduce_max_sync(full, live ? b.hi[1] : -1);
    const int zlo = __reduce_min_sync(full, live ? b.lo[2] : 0x7FFFFFFF), zhi = __reduce_max_sync(full, live ? b.hi[2] : -1);
    const int wlo = __reduce_min_sync(full, live ? b.lo[0] >> 5 : 0x7FFFFFFF), whi = __reduce_max_sync(full, live ? b.hi[0] >> 5 : -1);
    if (yhi < ylo) return false;                                      // no live lane
    // lanes as an 8 (y) x 4 (z) patch of rows, stepped over the range (no integer division: the kernel is issue-bound)
    const int ly = lane & 7, lz = lane >> 3;
    bool any = false;
    for (int w = wlo; w <= whi; w++) {
        uint32_t acc = 0;
        for (int zb = zlo; zb <= zhi; zb += 4)
            for (int yb = ylo; yb <= yhi; yb += 8) {
                const int y = yb + ly, z = zb + lz;
                if (y <= yhi && z <= zhi) acc |= __ldg(m + (uint32_t)(z * L.size + y) * (uint32_t)L.wpr + w);
            }
        acc = __reduce_or_sync(full, acc);
        if (live && w >= (b.lo[0] >> 5) && w <= (b.hi[0] >> 5)) {
            uint32_t sel = full;
            if (w == (b.lo[0] >> 5)) sel &= full << (b.lo[0] & 31);
            if (w == (b.hi[0] >> 5)) sel &= full >> (31 - (b.hi[0] & 31));
            any = any || (acc & sel) != 0;
        }
    }
    return any;
}

// One thread per cell, one warp per run of 32 cells along x.  Bit g of the cell's code: some step of group g, started
// anywhere inside the cell, can read a non-zero texel.  The test is the footprint itself: the cell's box pushed along the
// cone by the group's height range (widened by the variation of the light direction over the cell), converted to the texel
// range a trilinear lookup touches, against the non-zero bits of the level(s) the group samples.  Conservative by
// construction, so skipping on a clear bit is exact.
__global__ void __launch_bounds__(256) need_code_kernel(const __grid_constant__ CodeArgs a) {
    const int G = a.G;
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = blockIdx.y * 8 + (threadIdx.x >> 5), iz = blockIdx.z;
    if (iy >= G) return;                                              // (whole warps)
    bool live = ix < G;
    const uint32_t cell = ((uint32_t)iz * G + iy) * G + min(ix, G - 1);
    const float invG = 1.0f / (float)G;
    if (a.worldBox) {
        // fragments start on the billboards' spheres: a cell that does not meet their bounding box is never looked up
        const int ic[3] = {ix, iy, iz};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lo = unsortable(__ldg(a.worldBox + k)), hi = unsortable(__ldg(a.worldBox + 3 + k));
            const float c0 = a.b0[k] + ((float)ic[k] - 0.01f) * invG * a.range[k], c1 = a.b0[k] + ((float)ic[k] + 1.01f) * invG * a.range[k];
            if (fmaxf(c0, c1) < lo || fminf(c0, c1) > hi) live = false;
        }
    }
    if (!__any_sync(0xFFFFFFFFu, live)) return;
    // half a cell, in normalized coordinates, with slack: the fast trace variant finds its cell with a point-sampled texture
    // fetch, whose fixed-point coordinate may land in the neighbouring cell within 1/256 of a cell border
    const float half = 0.5f * invG * (1.0f + 1.0f / 64.0f) + 1.0e-6f;
    const float nc[3] = {((float)ix + 0.5f) * invG, ((float)iy + 0.5f) * invG, ((float)iz + 0.5f) * invG};
    float toL[3], hw2 = 0.0f, d2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float w = a.b0[k] + nc[k] * a.range[k];            // world position of the cell centre
        toL[k] = a.lightPos[k] - w;
        d2 += toL[k] * toL[k];
        const float hw = half * fabsf(a.range[k]);
        hw2 += hw * hw;
    }
    const float dist = sqrtf(d2), hw = sqrtf(hw2);               // |light - centre|, half diagonal of the cell in world units
    // the unit vector towards the light turns by at most |dw| / (distance to the light) over the cell; a light inside
    // or next to the cell gives no useful bound: every group stays needed there
    const float dmin = dist - hw;
    const uint32_t all = (1u << a.nGroups) - 1u;
    const bool near = !(dmin > 4.0f * hw) || !(dist > 0.0f);
    if (live && near) {
        a.code[cell] = (uint8_t)all;
        if (a.codeSurf) surf3Dwrite((unsigned char)0, a.codeSurf, ix, iy, iz);
    }
    live = live && !near;
    if (!__any_sync(0xFFFFFFFFu, live)) return;
    const float invDist = near ? 0.0f : 1.0f / dist, turn = near ? 0.0f : (hw / dmin) * 1.01f;
    float dir[3];
#pragma unroll
    for (int k = 0; k < 3; k++) dir[k] = toL[k] * invDist;

    // sample positions a group's steps can reach from inside the cell
    auto reach = [&](float h0, float h1, float pmin[3], float pmax[3]) {
        const float spread = h1 * turn + 2.0e-5f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float p0 = h0 * dir[k], p1 = h1 * dir[k];
            pmin[k] = nc[k] - half + fminf(p0, p1) - spread;
            pmax[k] = nc[k] + half + fmaxf(p0, p1) + spread;
        }
    };
    // Early out for all the groups whose levels lie at or below a coarse level (at most three levels above their own): the
    // coarse level's own trilinear footprint of everything those groups can reach contains the ancestors of every finer
    // texel they can read, and a non-zero texel has non-zero ancestors up to three levels up (255 -> 32 -> 4 -> 1 through
    // (sum + 4) >> 3; floats never vanish).  Nearly every cell leaves here.
    uint32_t candidates = live ? all : 0u;
    if (a.fineSet) {
        float pmin[3], pmax[3];
        reach(a.fineMin, a.fineMax, pmin, pmax);
        const CodeLevel &LC = a.lv[a.coarse];
        if (!warp_any_set(a.coarse == 0 ? a.bits : a.nz + LC.off, LC, footprint(pmin, pmax, LC), live)) candidates &= ~a.fineSet;
    }
    uint32_t bitsOut = 0;
    for (int g = 0; g < a.nGroups; g++) {
        const bool want = (candidates >> g) & 1u;
        if (!__any_sync(0xFFFFFFFFu, want)) continue;
        const CodeGroup &cg = a.g[g];
        float pmin[3], pmax[3];
        reach(cg.hMin, cg.hMax, pmin, pmax);
        bool any = false;
        for (int j = 0; j < cg.nLv; j++) {
            const CodeLevel &L = a.lv[cg.lv0 + j];
            any = warp_any_set(cg.lv0 + j == 0 ? a.bits : a.nz + L.off, L, footprint(pmin, pmax, L), want) || any;
        }
        if (any) bitsOut |= 1u << g;
    }
    if (live) {
        a.code[cell] = (uint8_t)bitsOut;
        if (a.codeSurf) surf3Dwrite((unsigned char)(~bitsOut & all), a.codeSurf, ix, iy, iz);
    }
}

} // namespace

int launch_bake_steps(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, const uint8_t *chain, const BakeTex *tex, int nTex) {
    if (nTex <= 0) return 0;
    BakeArgs a{};
    a.bits = bits; a.chain = chain;
    int most = 0;
    for (int i = 0; i < nTex; i++) {
        int g = -1;
        for (int j = 0; j < a.nGroups; j++)
            if (a.g[j].level0 == tex[i].level0) g = j;
        if (g < 0) {
            g = a.nGroups++;
            BakeGroup &bg = a.g[g];
            bg.level0 = tex[i].level0; bg.n = tex[i].n; bg.count = 0;
            bg.NL = vol.levelSize[bg.level0]; bg.offL = vol.levelOff[bg.level0];
            bg.NU = 0; bg.offU = 0;
            most = std::max(most, bg.n);
        }
        BakeGroup &bg = a.g[g];
        if (tex[i].frac > 0.0f) { bg.NU = vol.levelSize[bg.level0 + 1]; bg.offU = vol.levelOff[bg.level0 + 1]; }
        bg.frac[bg.count] = tex[i].frac; bg.surf[bg.count] = tex[i].surf; bg.count++;
    }
    const dim3 grid((most + 31) / 32, (most * ((most - 1 + kBakeRun - 1) / kBakeRun) + 7) / 8, a.nGroups);
    if (vol.texelBytes == 4) bake_steps_kernel<true><<<grid, 256, 0, st>>>(a);
    else bake_steps_kernel<false><<<grid, 256, 0, st>>>(a);
    return 1;
}

int launch_need_code(cudaStream_t st, const VolumeParams &vol, const TraceParams &tp, const uint32_t *bits, const uint32_t *nz, const uint32_t *worldBox,
                     uint8_t *code, cudaSurfaceObject_t codeSurf) {
    CodeArgs a{};
    a.G = tp.codeDim;
    a.nGroups = std::min(tp.nGroups, kCodeGroups);
    int top = 0;
    for (int g = 0; g < a.nGroups; g++) {
        const ConeGroup &gr = tp.groups[g];
        a.g[g].hMin = tp.steps[gr.first].height / (float)vol.dim; a.g[g].hMax = tp.steps[gr.first + gr.count - 1].height / (float)vol.dim;
        a.g[g].lv0 = gr.level; a.g[g].nLv = (gr.two && gr.level + 1 < vol.levels) ? 2 : 1;
        top = std::max(top, gr.level + a.g[g].nLv - 1);
    }
    a.coarse = std::min(std::min(std::max(top, 2), 3), vol.levels - 1);
    a.fineSet = 0; a.fineMin = 1e30f; a.fineMax = 0.0f;
    for (int g = 0; g < a.nGroups; g++)
        if (a.g[g].lv0 + a.g[g].nLv - 1 <= a.coarse && a.coarse - a.g[g].lv0 <= 3) {
            a.fineSet |= 1u << g;
            a.fineMin = std::min(a.fineMin, a.g[g].hMin); a.fineMax = std::max(a.fineMax, a.g[g].hMax);
        }
    uint32_t off[kMaxLevels] = {};
    skipmask_words(vol, off);
    for (int l = 0; l < vol.levels; l++) {
        a.lv[l].size = vol.levelSize[l]; a.lv[l].sizeF = (float)vol.levelSize[l];
        a.lv[l].wpr = vol.levelSize[l] >= 32 ? vol.levelSize[l] / 32 : 1; a.lv[l].off = off[l];
    }
    a.bits = bits; a.nz = nz; a.code = code; a.worldBox = worldBox; a.codeSurf = codeSurf;
    a.b0[0] = vol.xB[0]; a.b0[1] = vol.yB[0]; a.b0[2] = vol.zB[0];
    a.range[0] = vol.xB[1] - vol.xB[0]; a.range[1] = vol.yB[1] - vol.yB[0]; a.range[2] = vol.zB[1] - vol.zB[0];
    for (int k = 0; k < 3; k++) a.lightPos[k] = tp.lightPos[k];
    need_code_kernel<<<dim3((a.G + 31) / 32, (a.G + 7) / 8, a.G), 256, 0, st>>>(a);
    return 1;
}

} // namespace crn
