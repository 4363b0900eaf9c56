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

// ---- voxel export: VoxelShader::updateVoxelData (src/Shaders/VoxelShader.cpp:102-133) ---------------------------
// The reference reads level 0 back and loops over all D^3 texels on the CPU, collecting for every non-empty one
// position + reverseVoxelIndex(get3DIndices(i)) (src/CloudVolume.cpp:103-118) and its value.  Here: an ordered stream
// compaction of the occupancy bits, three tiny launches (per-block counts, scan of the block counts, write).
constexpr int kExpWordsPerThread = 4, kExpThreads = 256, kExpWordsPerBlock = kExpWordsPerThread * kExpThreads;

__global__ void __launch_bounds__(kExpThreads) export_count_kernel(const uint32_t *__restrict__ bits, uint32_t words,
                                                                   uint32_t *__restrict__ blockSum) {
    const uint32_t w0 = blockIdx.x * kExpWordsPerBlock + threadIdx.x * kExpWordsPerThread;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kExpWordsPerThread; k++) if (w0 + k < words) c += __popc(bits[w0 + k]);
    __shared__ uint32_t sWarp[kExpThreads / 32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < kExpThreads / 32; k++) t += sWarp[k];
        blockSum[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts in place (one CTA; at most 32768 blocks for 1024^3), total -> *total
__global__ void __launch_bounds__(1024) export_scan_kernel(uint32_t *__restrict__ blockSum, uint32_t nBlocks,
                                                           unsigned long long *__restrict__ total) {
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sCarry;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nBlocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nBlocks ? blockSum[i] : 0;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sWarp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sWarp[lane];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o); if (lane >= o) w += t; }
            sWarp[lane] = w;                                   // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = sCarry, warpOff = warp ? sWarp[warp - 1] : 0;
        if (i < nBlocks) blockSum[i] = carry + warpOff + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) sCarry = carry + sWarp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = sCarry;
}

struct ExportArgs {
    const uint32_t *bits;         // which voxels to emit
    const uint32_t *lit;          // value channel: w = 1 where this bit is set, else 0
    uint32_t words;
    const uint32_t *blockOff;
    float4 *out;
    unsigned long long capacity;
    int D;
    float pos[3], lo[3], range[3];
};

__global__ void __launch_bounds__(kExpThreads) export_write_kernel(ExportArgs a) {
    const uint32_t w0 = blockIdx.x * kExpWordsPerBlock + threadIdx.x * kExpWordsPerThread;
    uint32_t w[kExpWordsPerThread], c = 0;
#pragma unroll
    for (int k = 0; k < kExpWordsPerThread; k++) { w[k] = (w0 + k < a.words) ? a.bits[w0 + k] : 0u; c += __popc(w[k]); }
    // exclusive scan of c over the CTA
    __shared__ uint32_t sWarp[kExpThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    uint32_t warpOff = 0;
    for (int k = 0; k < warp; k++) warpOff += sWarp[k];
    unsigned long long o = (unsigned long long)a.blockOff[blockIdx.x] + warpOff + inc - c;
    const int wpr = a.D >> 5;
    const float fdim = (float)a.D;
#pragma unroll
    for (int k = 0; k < kExpWordsPerThread; k++) {
        uint32_t m = w[k];
        if (!m) continue;
        const uint32_t word = w0 + k;
        const uint32_t litw = a.lit[word];
        const int row = word / wpr, y = row % a.D, z = row / a.D, xb = (word % wpr) << 5;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (o < a.capacity) {
                const int x = xb + b;
                // position + reverseVoxelIndex: float(idx) * range / dimension + bounds.x  (src/CloudVolume.cpp:112-118)
                const float vx = (float)x * a.range[0] / fdim + a.lo[0];
                const float vy = (float)y * a.range[1] / fdim + a.lo[1];
                const float vz = (float)z * a.range[2] / fdim + a.lo[2];
                a.out[o] = make_float4(a.pos[0] + vx, a.pos[1] + vy, a.pos[2] + vz, ((litw >> b) & 1u) ? 1.0f : 0.0f);
            }
            o++;
        }
    }
}

} // namespace

size_t export_scratch_words(size_t words) { return (words + kExpWordsPerBlock - 1) / kExpWordsPerBlock; }

int launch_export_voxels(cudaStream_t st, const uint32_t *bits, const uint32_t *lit, size_t words, int D, const float pos[3],
                         const float lo[3], const float range[3], uint32_t *blockScratch, unsigned long long *total, float4 *out,
                         unsigned long long capacity) {
    const uint32_t nBlocks = (uint32_t)export_scratch_words(words);
    export_count_kernel<<<nBlocks, kExpThreads, 0, st>>>(bits, (uint32_t)words, blockScratch);
    export_scan_kernel<<<1, 1024, 0, st>>>(blockScratch, nBlocks, total);
    if (!out || !capacity) return 2;
    ExportArgs a;
    a.bits = bits; a.lit = lit; a.words = (uint32_t)words; a.blockOff = blockScratch; a.out = out; a.capacity = capacity; a.D = D;
    for (int k = 0; k < 3; k++) { a.pos[k] = pos[k]; a.lo[k] = lo[k]; a.range[k] = range[k]; }
    export_write_kernel<<<nBlocks, kExpThreads, 0, st>>>(a);
    return 3;
}

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
