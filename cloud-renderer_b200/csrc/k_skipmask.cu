// k_skipmask.cu — conservative empty-space masks for the cone trace.
//
// traceCone (res/conetrace_frag.glsl:64-79) always takes all `steps` samples, but in a cloud
// only the sun-facing shell is lit (1 % of the voxels at C3), so ~80 % of the samples read
// nothing but zero texels and contribute exactly 0 to the sum.  Skipping those is EXACT as long
// as the skip test is conservative.  For every level l this file builds one bit per texel:
//
//     M_l(t) = 1  iff  some texel of level l within the 5x5x5 neighbourhood of t is non-zero,
//                 or   some texel of level l+1 within the 5x5x5 neighbourhood of parent(t) is.
//
// The trace kernel groups consecutive cone steps that share the lower mip level and lie within
// one level-l texel of a common point; if M_l at the texel containing that point is 0, every
// trilinear footprint (level l and level l+1, CLAMP_TO_EDGE) of every step in the group lies
// inside the all-zero neighbourhood, so all of them return exactly 0 and are skipped.
//
// Three tiny launches per frame (bit-parallel, word = 32 texels along x):
//   nonzero_kernel : R8 levels >= 1 -> bit volumes (level 0 is already the occupancy set)
//   dilate_kernel  : S_l = 5x5x5 dilation of the non-zero bits of level l
//   combine_kernel : M_l = S_l | upsample(S_{l+1})
#include "crn_internal.cuh"

namespace crn {

namespace {

struct MaskArgs {
    int levels;
    int size[kMaxLevels];          // texels per axis
    int wpr[kMaxLevels];           // words per row = max(1, size/32)
    uint32_t off[kMaxLevels];      // word offset of the level in each mask buffer
    uint32_t chainOff[kMaxLevels];
    uint32_t totalWords;
    const uint32_t *bits;          // level-0 occupancy
    const uint8_t *chain;
    int texelBytes;
    uint32_t *nz;                  // non-zero bits, levels >= 1 (level 0 slot unused)
    uint32_t *dil;                 // S_l
    uint32_t *mask;                // M_l
    uint32_t *fill;                // fill[l] = number of set bits of M_l (the trace kernel skips tests that cannot pay off)
};

__device__ __forceinline__ bool locate(const MaskArgs &a, uint32_t w, int &l, int &z, int &y, int &wx) {
    if (w >= a.totalWords) return false;
    l = 0;
    while (l + 1 < a.levels && w >= a.off[l + 1]) l++;
    uint32_t r = w - a.off[l];
    const int n = a.size[l], wpr = a.wpr[l];
    wx = r % wpr; r /= wpr;
    y = r % n; z = r / n;
    return true;
}

__global__ void __launch_bounds__(256) nonzero_kernel(MaskArgs a) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x + a.off[1];     // levels >= 1 only
    int l, z, y, wx;
    if (!locate(a, w, l, z, y, wx)) return;
    const int n = a.size[l];
    const int cnt = min(32, n);
    uint32_t m = 0;
    if (a.texelBytes == 4) {
        const float *rowf = reinterpret_cast<const float *>(a.chain + a.chainOff[l]) + ((size_t)z * n + y) * n + wx * 32;
        for (int k = 0; k < cnt; k++) m |= (rowf[k] != 0.0f ? 1u : 0u) << k;
        a.nz[w] = m;
        return;
    }
    const uint8_t *row = a.chain + a.chainOff[l] + ((size_t)z * n + y) * n + wx * 32;
    if (cnt == 32) {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(row), v1 = *reinterpret_cast<const uint4 *>(row + 16);
        const uint32_t q[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t x = q[k];
            m |= ((x & 0xFFu) ? 1u : 0u) << (4 * k) | ((x & 0xFF00u) ? 1u : 0u) << (4 * k + 1) |
                 ((x & 0xFF0000u) ? 1u : 0u) << (4 * k + 2) | ((x & 0xFF000000u) ? 1u : 0u) << (4 * k + 3);
        }
    } else {
        for (int k = 0; k < cnt; k++) m |= (row[k] ? 1u : 0u) << k;
    }
    a.nz[w] = m;
}

__global__ void __launch_bounds__(256) dilate_kernel(MaskArgs a) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    int l, z, y, wx;
    if (!locate(a, w, l, z, y, wx)) return;
    const int n = a.size[l], wpr = a.wpr[l];
    const uint32_t *src = (l == 0 ? a.bits : a.nz + a.off[l]);
    // dilation commutes with OR: gather the 5x5 (y,z) rows first, spread along x once
    uint32_t c = 0, lo = 0, hi = 0;
    const bool hasLo = wx > 0, hasHi = wx + 1 < wpr;
    for (int dz = -2; dz <= 2; dz++) {
        const int zz = z + dz;
        if (zz < 0 || zz >= n) continue;
#pragma unroll
        for (int dy = -2; dy <= 2; dy++) {
            const int yy = y + dy;
            if (yy < 0 || yy >= n) continue;
            const uint32_t *row = src + ((size_t)zz * n + yy) * wpr + wx;
            c |= row[0];
            if (hasLo) lo |= row[-1];
            if (hasHi) hi |= row[1];
        }
    }
    uint32_t acc = c | (c << 1) | (c << 2) | (c >> 1) | (c >> 2) | (lo >> 31) | (lo >> 30) | (hi << 31) | (hi << 30);
    if (n < 32) acc &= (1u << n) - 1u;
    a.dil[w] = acc;
}

__device__ __forceinline__ uint32_t double_bits16(uint32_t h) {     // 16 bits -> each bit twice (32 bits)
    uint32_t x = h & 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}

__global__ void __launch_bounds__(256) combine_kernel(MaskArgs a) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    int l = -1, z, y, wx;
    uint32_t m = 0;
    if (locate(a, w, l, z, y, wx)) {
        m = a.dil[w];
        if (l + 1 < a.levels) {
            const int np = a.size[l + 1], wprp = a.wpr[l + 1];
            const uint32_t *prow = a.dil + a.off[l + 1] + ((size_t)(z >> 1) * np + (y >> 1)) * wprp;
            const uint32_t pw = prow[wx >> 1];
            m |= double_bits16((wx & 1) ? (pw >> 16) : pw);
            if (a.size[l] < 32) m &= (1u << a.size[l]) - 1u;
        }
        a.mask[w] = m;
    }
    // per-level population of M_l: one atomic per (warp, level)
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, l);
    const uint32_t sum = __reduce_add_sync(peers, (uint32_t)__popc(m));
    if (l >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1 && sum) atomicAdd(&a.fill[l], sum);
}

} // namespace

size_t skipmask_words(const VolumeParams &vol, uint32_t *off) {
    size_t total = 0;
    for (int l = 0; l < vol.levels; l++) {
        const int n = vol.levelSize[l];
        if (off) off[l] = (uint32_t)total;
        total += (size_t)n * n * (n >= 32 ? n / 32 : 1);
    }
    return total;
}

int launch_skipmask(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, const uint8_t *chain, uint32_t *nz,
                    uint32_t *dil, uint32_t *mask, uint32_t *fill) {
    MaskArgs a{};
    a.levels = vol.levels;
    for (int l = 0; l < vol.levels; l++) {
        a.size[l] = vol.levelSize[l];
        a.wpr[l] = vol.levelSize[l] >= 32 ? vol.levelSize[l] / 32 : 1;
        a.chainOff[l] = vol.levelOff[l];
    }
    a.totalWords = (uint32_t)skipmask_words(vol, a.off);
    a.bits = bits; a.chain = chain; a.nz = nz; a.dil = dil; a.mask = mask; a.fill = fill;
    cudaMemsetAsync(fill, 0, kMaxLevels * sizeof(uint32_t), st);
    a.texelBytes = vol.texelBytes;
    int launches = 0;
    if (vol.levels > 1) {
        const uint32_t upper = a.totalWords - a.off[1];
        nonzero_kernel<<<(upper + 255) / 256, 256, 0, st>>>(a);
        launches++;
    }
    dilate_kernel<<<(a.totalWords + 255) / 256, 256, 0, st>>>(a);
    combine_kernel<<<(a.totalWords + 255) / 256, 256, 0, st>>>(a);
    return launches + 2;
}

} // namespace crn
