// k_scene.cu — device-side scene generation and animation.
//
// Replaces CloudVolume::regenerateBillboards (src/CloudVolume.cpp:120-137) and the per-frame
// host work + re-upload of CloudVolume::uploadBillboards (src/CloudVolume.cpp:139-164) for the
// animated configs: the billboard arrays are produced and advected where they are consumed,
// so a frame needs no host->device copy at all.
//
// The reference draws offsets/scales from Util::genRandom (rand() seeded by time(0),
// src/Util.hpp:19-45, src/main.cpp:70), which is not reproducible; the stream here is the
// splitmix64 generator of the fixed-seed fixtures (cloud-renderer_b200/scene.py), evaluated
// counter-based so every billboard is one independent thread:
//     u_k   = top 53 bits of splitmix64(seed, 4*i + k + 1) * 2^-53           (float64, [0,1))
//     off_c = (float)(u_c * (maxOffset_c - minOffset_c) + minOffset_c)          c = x,y,z
//     scale = (float)(u_3 * (maxScale - minScale) + minScale)
// Float64 products and sums are rounded separately (-fmad=false), so the bytes equal what the
// numpy fixture code produces.
#include "crn_internal.cuh"

namespace crn {

namespace {

__device__ __forceinline__ double splitmix_u01(uint64_t seed, uint64_t i) {
    uint64_t z = seed + i * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct GenArgs {
    uint64_t seed;
    double lo[4], span[4];            // x, y, z, scale
    double radiusFactor;
    int n;
};

__global__ void __launch_bounds__(256) generate_kernel(GenArgs a, float *pos0, float *pos, float *scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double u = splitmix_u01(a.seed, (uint64_t)i * 4 + k + 1);
        double d = u * a.span[k] + a.lo[k];
        if (k == 3) d = d * a.radiusFactor;
        v[k] = (float)d;
    }
    pos0[3 * i + 0] = v[0]; pos0[3 * i + 1] = v[1]; pos0[3 * i + 2] = v[2];
    pos[3 * i + 0] = v[0];  pos[3 * i + 1] = v[1];  pos[3 * i + 2] = v[2];
    scale[i] = v[3];
}

// rigid rotation of the base offsets about +Y (the analytic field of the animated configs,
// SURVEY.md §8d): float32, one rounding per operation
__global__ void __launch_bounds__(256) rotate_kernel(const float *pos0, float *pos, int n, float c, float s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pos0[3 * i + 0], y = pos0[3 * i + 1], z = pos0[3 * i + 2];
    pos[3 * i + 0] = c * x + s * z;
    pos[3 * i + 1] = y;
    pos[3 * i + 2] = (-s) * x + c * z;
}

} // namespace

int launch_generate_boards(cudaStream_t st, int n, const float minOff[3], const float maxOff[3], float minScale,
                           float maxScale, double radiusFactor, uint64_t seed, float *pos0, float *pos, float *scale) {
    if (n <= 0) return 0;
    GenArgs a;
    a.seed = seed; a.n = n; a.radiusFactor = radiusFactor;
    for (int k = 0; k < 3; k++) { a.lo[k] = (double)minOff[k]; a.span[k] = (double)maxOff[k] - (double)minOff[k]; }
    a.lo[3] = (double)minScale; a.span[3] = (double)maxScale - (double)minScale;
    generate_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, pos0, pos, scale);
    return 1;
}

int launch_rotate_boards(cudaStream_t st, int n, const float *pos0, float *pos, float c, float s) {
    if (n <= 0) return 0;
    rotate_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos0, pos, n, c, s);
    return 1;
}

} // namespace crn
