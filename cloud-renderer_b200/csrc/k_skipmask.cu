// k_skipmask.cu — non-zero bit volumes of the mip levels, for the cone trace's empty-space skipping.
//
// traceCone (res/conetrace_frag.glsl:64-79) always takes all `steps` samples, but in a cloud
// only the sun-facing shell is lit (1 % of the voxels at C3), so most of the samples read
// nothing but zero texels and contribute exactly 0 to the sum.  Skipping those is EXACT as long
// as the skip test is conservative.  The test itself lives in k_conebake.cu (need codes: the
// texel footprint a group of cone steps can reach from a cell of start positions); this file
// supplies what it is tested against: one bit per texel of every level >= 1, set iff the texel
// is non-zero (level 0 is already the 1-bit occupancy set).  One tiny launch per frame
// (bit-parallel, word = 32 texels along x).
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
    const uint8_t *chain;
    int texelBytes;
    uint32_t *nz;                  // non-zero bits, levels >= 1 (level 0 slot unused: the occupancy set is that level's)
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

int launch_skipmask(cudaStream_t st, const VolumeParams &vol, const uint8_t *chain, uint32_t *nz) {
    if (vol.levels <= 1) return 0;
    MaskArgs a{};
    a.levels = vol.levels;
    for (int l = 0; l < vol.levels; l++) {
        a.size[l] = vol.levelSize[l];
        a.wpr[l] = vol.levelSize[l] >= 32 ? vol.levelSize[l] / 32 : 1;
        a.chainOff[l] = vol.levelOff[l];
    }
    a.totalWords = (uint32_t)skipmask_words(vol, a.off);
    a.chain = chain; a.nz = nz;
    a.texelBytes = vol.texelBytes;
    const uint32_t upper = a.totalWords - a.off[1];
    nonzero_kernel<<<(upper + 255) / 256, 256, 0, st>>>(a);
    return 1;
}

} // namespace crn
