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
//     from the linear chain) into a layered RG16 texture whose layer k holds the node planes k and k+1: the trace
//     kernel then needs ONE bilinear pass of the texture unit + one z blend per step instead of the four passes of a
//     mip-linear 3D fetch (measured: 288 G/s for tex3DLod at a fractional LOD, 1121 G/s for RG16 layered bilinear).
//     UNORM16 storage: 7.6e-6 absolute, far below the 8-bit filter weights of the texture unit.
//
// (2) Need codes.  The empty-space masks M_l (k_skipmask.cu) decide whether a GROUP of cone steps can contribute;
//     the old kernel tested every group per fragment (8 lookups of ~30 instructions).  The lookup point of group g,
//     P_g(pos) = pos + h_g * normalize(light - pos) / dim, is a function of the start position only, so a coarse grid
//     over the volume (cells of 2 voxels) stores, per cell, one bit per group: the OR of M_l over every texel P_g can
//     reach from a start position inside the cell (cell box pushed along the cone, widened by the variation of the
//     light direction over the cell).  A fragment reads one byte.  Conservative by construction, so still exact.
#include "crn_internal.cuh"

#include <algorithm>

namespace crn {

namespace {

struct BakeArgs {
    VolumeParams vol;
    const uint32_t *bits;
    const uint8_t *chain;
    int nTex;
    BakeTex tex[kMaxBakedTex];
};

// trilinear sample of level l at the lattice node (kx,ky,kz); `scale` = lattice spacing / texel size of level l
// (1/2 for the lower level, 1/4 for the upper one; 1/2 and 1 when the lattice is the half-voxel one of level 0)
__device__ __forceinline__ float node_sample(const BakeArgs &a, int l, float scale, int kx, int ky, int kz) {
    const int N = a.vol.levelSize[l];
    int i0[3], i1[3];
    float w[3];
    const int k[3] = {kx, ky, kz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float t = (float)k[d] * scale - 0.5f;              // exact: multiples of 1/4
        const float fl = floorf(t);
        w[d] = t - fl;
        const int i = (int)fl;
        i0[d] = min(max(i, 0), N - 1);                           // CLAMP_TO_EDGE
        i1[d] = min(max(i + 1, 0), N - 1);
    }
    float c[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int x = (q & 1) ? i1[0] : i0[0], y = (q & 2) ? i1[1] : i0[1], z = (q & 4) ? i1[2] : i0[2];
        if (l == 0) {
            const uint32_t word = __ldg(a.bits + ((size_t)z * N + y) * (N >> 5) + (x >> 5));
            c[q] = (float)((word >> (x & 31)) & 1u);
        } else if (a.vol.texelBytes == 4) {
            c[q] = __ldg(reinterpret_cast<const float *>(a.chain + a.vol.levelOff[l]) + ((size_t)z * N + y) * N + x);
        } else {
            c[q] = (float)__ldg(a.chain + a.vol.levelOff[l] + ((size_t)z * N + y) * N + x) * (1.0f / 255.0f);
        }
    }
    const float x00 = fmaf(w[0], c[1] - c[0], c[0]), x10 = fmaf(w[0], c[3] - c[2], c[2]);
    const float x01 = fmaf(w[0], c[5] - c[4], c[4]), x11 = fmaf(w[0], c[7] - c[6], c[6]);
    const float y0 = fmaf(w[1], x10 - x00, x00), y1 = fmaf(w[1], x11 - x01, x01);
    return fmaf(w[2], y1 - y0, y0);
}

__device__ __forceinline__ float node_value(const BakeArgs &a, const BakeTex &t, int kx, int ky, int kz) {
    // lattice spacing is 2^(L-1) voxels (half a voxel for L = 0): half a texel of level L, a quarter of level L+1
    const float lo = node_sample(a, t.level0, 0.5f, kx, ky, kz);
    if (!(t.frac > 0.0f)) return lo;
    const float hi = node_sample(a, t.level0 + 1, 0.25f, kx, ky, kz);
    return fmaf(t.frac, hi - lo, lo);
}

__global__ void __launch_bounds__(256) bake_steps_kernel(const __grid_constant__ BakeArgs a) {
    const BakeTex &t = a.tex[blockIdx.y];
    const int n = t.n;
    const size_t total = (size_t)n * n * (n - 1);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e % n), y = (int)((e / n) % n), k = (int)(e / ((size_t)n * n));
        const float v0 = node_value(a, t, x, y, k), v1 = node_value(a, t, x, y, k + 1);
        ushort2 o;
        o.x = (unsigned short)__float2uint_rn(__saturatef(v0) * 65535.0f);
        o.y = (unsigned short)__float2uint_rn(__saturatef(v1) * 65535.0f);
        surf2DLayeredwrite(o, t.surf, x * 4, y, k);
    }
}

struct CodeArgs {
    int G;
    int nGroups;
    float height[kCodeGroups];
    int size[kCodeGroups], wpr[kCodeGroups];
    uint32_t maskOff[kCodeGroups];
    const uint32_t *mask;
    float lightPos[3], b0[3], range[3];
    float invDim;
    uint8_t *code;
};

__global__ void __launch_bounds__(256) need_code_kernel(const __grid_constant__ CodeArgs a) {
    const int G = a.G;
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)G * G * G) return;
    const int ix = (int)(cell % G), iy = (int)((cell / G) % G), iz = (int)(cell / ((size_t)G * G));
    const float invG = 1.0f / (float)G;
    const float half = 0.5f * invG * 1.001f + 1.0e-6f;           // half a cell, in normalized coordinates, with slack
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
    uint32_t bitsOut = 0;
    // the unit vector towards the light turns by at most |dw| / (distance to the light) over the cell; a light inside
    // or next to the cell gives no useful bound: every group stays needed there
    const float dmin = dist - hw;
    if (!(dmin > 4.0f * hw) || !(dist > 0.0f)) {
        a.code[cell] = 0xFF;
        return;
    }
    const float invDist = 1.0f / dist;
    for (int g = 0; g < a.nGroups; g++) {
        const float reach = a.height[g] * a.invDim;              // |P - pos| in normalized coordinates
        const float ext = half + reach * (hw / dmin) * 1.01f + 2.0e-5f;
        const int n = a.size[g];
        int lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float p = nc[k] + reach * toL[k] * invDist;
            // the trace kernel looks the point up with CLAMP_TO_EDGE semantics: clamp the texel range the same way
            lo[k] = min(max((int)floorf((p - ext) * (float)n), 0), n - 1);
            hi[k] = min(max((int)floorf((p + ext) * (float)n), 0), n - 1);
        }
        const uint32_t *m = a.mask + a.maskOff[g];
        uint32_t any = 0;
        const int w0 = lo[0] >> 5, w1 = hi[0] >> 5;
        for (int z = lo[2]; z <= hi[2] && !any; z++)
            for (int y = lo[1]; y <= hi[1] && !any; y++) {
                const uint32_t *row = m + ((size_t)z * n + y) * a.wpr[g];
                for (int w = w0; w <= w1; w++) {
                    uint32_t sel = 0xFFFFFFFFu;
                    if (w == w0) sel &= 0xFFFFFFFFu << (lo[0] & 31);
                    if (w == w1) sel &= 0xFFFFFFFFu >> (31 - (hi[0] & 31));
                    any |= __ldg(row + w) & sel;
                }
            }
        if (any) bitsOut |= 1u << g;
    }
    a.code[cell] = (uint8_t)bitsOut;
}

} // namespace

int launch_bake_steps(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, const uint8_t *chain, const BakeTex *tex, int nTex) {
    if (nTex <= 0) return 0;
    BakeArgs a;
    a.vol = vol; a.bits = bits; a.chain = chain; a.nTex = nTex;
    size_t most = 0;
    for (int i = 0; i < nTex; i++) {
        a.tex[i] = tex[i];
        most = std::max(most, (size_t)tex[i].n * tex[i].n * (tex[i].n - 1));
    }
    const unsigned blocks = (unsigned)std::min<size_t>((most + 255) / 256, 148 * 16);
    bake_steps_kernel<<<dim3(blocks, nTex), 256, 0, st>>>(a);
    return 1;
}

int launch_need_code(cudaStream_t st, const VolumeParams &vol, const TraceParams &tp, const uint32_t *mask, uint8_t *code) {
    CodeArgs a{};
    a.G = tp.codeDim;
    a.nGroups = std::min(tp.nGroups, kCodeGroups);
    for (int g = 0; g < a.nGroups; g++) {
        a.height[g] = tp.groups[g].height; a.size[g] = tp.groups[g].size; a.wpr[g] = tp.groups[g].wpr; a.maskOff[g] = tp.groups[g].maskOff;
    }
    a.mask = mask; a.code = code;
    a.b0[0] = vol.xB[0]; a.b0[1] = vol.yB[0]; a.b0[2] = vol.zB[0];
    a.range[0] = vol.xB[1] - vol.xB[0]; a.range[1] = vol.yB[1] - vol.yB[0]; a.range[2] = vol.zB[1] - vol.zB[0];
    for (int k = 0; k < 3; k++) a.lightPos[k] = tp.lightPos[k];
    a.invDim = 1.0f / (float)vol.dim;
    const size_t cells = (size_t)a.G * a.G * a.G;
    need_code_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(a);
    return 1;
}

} // namespace crn
