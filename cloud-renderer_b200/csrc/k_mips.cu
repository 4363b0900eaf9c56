// k_mips.cu — stage 2: level-0 materialisation + the whole 3D mip chain in ONE launch.
//
// Replaces glGenerateMipmap(GL_TEXTURE_3D) (src/Shaders/VoxelizeShader.cpp:105) on the R8
// volume (src/CloudVolume.cpp:18): a 2x2x2 box filter per level, every level re-quantised to
// UNORM8.  The driver-defined tie rounding is decreed round-half-up: (sum of 8 + 4) >> 3.
//
// B200 design.  Input is the 1-bit occupancy set k_voxelize.cu produced (D^3/8 bytes), so
// the kernel is write-bound: it reads 1/8 byte and writes 8/7 bytes per voxel.
//   * one CTA owns a BX x 16 x 16 brick (BX = min(D,128)); every thread fetches one row of
//     the brick with a single 128-bit load and parks it in shared memory;
//   * level 0: bits -> bytes with a multiply-spread, written as full 128-byte lines
//     (8 lanes x 16 B per row);
//   * level 1 straight from the bits: 2-bit pair sums, widened to nibbles, added over the
//     four rows of a 2x2 (y,z) block -> 16 texels per thread, one 128-bit store;
//   * levels 2..4 from shared memory (the brick collapses to BX/16 x 1 x 1);
//   * levels 5.. (<= 32 KB of input up to 512^3) by whichever CTA retires last
//     (fence + ticket), so the chain is complete when the launch ends.
// Integer arithmetic throughout: results are bit-exact against the oracle.
#include "crn_internal.cuh"

#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cstdlib>

namespace crn {

namespace {

struct MipArgs {
    VolumeParams vol;
    const uint32_t *bits;
    uint8_t *chain;
    uint32_t *ticket;
    int bx;                 // brick x extent in voxels: 32, 64 or 128
    int writeLevel0;
    int tail;               // compute levels >= 5 in the last CTA (off in slab mode)
    int zBrick0;            // first brick row in z (slab mode)
    int useSurf;            // also write every level into the texture-unit copy (CRN_SAMPLER_TEXTURE)
    cudaSurfaceObject_t surf[kMaxLevels];
};

__device__ __forceinline__ uint32_t spread4(uint32_t nib) {    // 4 bits -> 4 bytes of 0x00/0xFF
    return ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;
}

__device__ __forceinline__ uint32_t box8(uint32_t sum) { return (sum + 4u) >> 3; }

// value of a level-1 texel whose 2x2x2 block holds `cnt` occupied (255) voxels: (255*cnt+4)>>3
__device__ __forceinline__ uint32_t level1_value(uint32_t cnt) { return cnt * 32u - (cnt > 4u ? 1u : 0u); }

// kTma (A/B variant, CRN_MIPS_TMA=1, WX = 4 only): the linear level-0 brick (128 x 16 x 16 bytes) is expanded into shared
// memory and written with ONE 3-D TMA store (cp.async.bulk.tensor.3d.global.shared::cta) instead of 2048 st.global.v4
template <int WX, bool kTma>   // words per brick row: 1, 2 or 4
__global__ void __launch_bounds__(256) mip_chain_kernel(MipArgs a, const __grid_constant__ CUtensorMap tmL0) {
    constexpr int BX = WX * 32;
    __shared__ uint32_t sBits[256 * WX];                 // [row = z*16+y][word]
    __shared__ __align__(128) uint8_t sL0[kTma ? 256 * BX : 16];      // [z][y][x], the TMA box
    __shared__ __align__(16) uint8_t sL1[(BX / 2) * 8 * 8];
    __shared__ uint8_t sL2[(BX / 4) * 4 * 4];
    __shared__ uint8_t sL3[(BX / 8) * 2 * 2];
    __shared__ uint8_t sL4[(BX / 16)];
    __shared__ uint32_t sLast;

    const int D = a.vol.dim, L = a.vol.levels;
    const int tid = threadIdx.x;
    const int bxi = blockIdx.x, byi = blockIdx.y, bzi = blockIdx.z + a.zBrick0;
    const int x0 = bxi * BX, y0 = byi * 16, z0 = bzi * 16;
    const int wordsPerRow = D >> 5;

    {   // one row of the brick per thread, 128-bit when the brick is 128 wide
        const int y = tid & 15, z = tid >> 4;
        const uint32_t *src = a.bits + ((size_t)(z0 + z) * D + (y0 + y)) * wordsPerRow + (x0 >> 5);
        if constexpr (WX == 4) {
            *reinterpret_cast<uint4 *>(&sBits[tid * 4]) = __ldg(reinterpret_cast<const uint4 *>(src));
        } else if constexpr (WX == 2) {
            *reinterpret_cast<uint2 *>(&sBits[tid * 2]) = __ldg(reinterpret_cast<const uint2 *>(src));
        } else {
            sBits[tid] = __ldg(src);
        }
    }
    __syncthreads();

    if (a.writeLevel0 || a.useSurf) {                    // linear level 0 only on request (nothing in the pipeline reads it)
        uint8_t *l0 = a.chain + a.vol.levelOff[0];
        constexpr int SEG = BX / 16;                     // 16-byte segments per row
#pragma unroll
        for (int it = 0; it < SEG; it++) {
            const int s = it * 256 + tid;
            const int row = s / SEG, seg = s % SEG;
            const uint32_t w = sBits[row * WX + (seg >> 1)];
            const uint32_t h = (seg & 1) ? (w >> 16) : (w & 0xFFFFu);
            uint4 o;
            o.x = spread4(h & 15u); o.y = spread4((h >> 4) & 15u); o.z = spread4((h >> 8) & 15u); o.w = spread4(h >> 12);
            const int y = row & 15, z = row >> 4;
            if (a.writeLevel0) {
                if constexpr (kTma) *reinterpret_cast<uint4 *>(&sL0[row * BX + seg * 16]) = o;
                else *reinterpret_cast<uint4 *>(l0 + ((size_t)(z0 + z) * D + (y0 + y)) * D + x0 + seg * 16) = o;
            }
            if (a.useSurf) surf3Dwrite(o, a.surf[0], x0 + seg * 16, y0 + y, z0 + z);
        }
        if constexpr (kTma) {
            if (a.writeLevel0) {
                // make the generic-proxy writes to shared memory visible to the async proxy, then one thread issues the store
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sL0);
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                 :: "l"(reinterpret_cast<uint64_t>(&tmL0)), "r"(x0), "r"(y0), "r"(z0), "r"(src) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // sL0 is not reused, but the CTA must not retire under the copy
                }
            }
        }
    }

    if (L > 1) {   // level 1 from the bits: thread <-> (level-1 row, word); 64 rows x WX words
        constexpr int N = 64 * WX;
        if (tid < N) {
            const int w = tid % WX, r1 = tid / WX;
            const int y1 = r1 & 7, z1 = r1 >> 3;
            uint32_t ev = 0, od = 0;                     // nibble sums of the even / odd level-1 texels
#pragma unroll
            for (int dz = 0; dz < 2; dz++)
#pragma unroll
                for (int dy = 0; dy < 2; dy++) {
                    const uint32_t b = sBits[((2 * z1 + dz) * 16 + (2 * y1 + dy)) * WX + w];
                    const uint32_t p = (b & 0x55555555u) + ((b >> 1) & 0x55555555u);
                    ev += p & 0x33333333u;
                    od += (p >> 2) & 0x33333333u;
                }
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {                // 4 output bytes per word: texels 4q..4q+3
                const uint32_t c0 = (ev >> (8 * q)) & 15u, c1 = (od >> (8 * q)) & 15u;
                const uint32_t c2 = (ev >> (8 * q + 4)) & 15u, c3 = (od >> (8 * q + 4)) & 15u;
                o[q] = level1_value(c0) | (level1_value(c1) << 8) | (level1_value(c2) << 16) | (level1_value(c3) << 24);
            }
            const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4 *>(&sL1[(r1 * WX + w) * 16]) = ov;
            const int D1 = D >> 1;
            uint8_t *l1 = a.chain + a.vol.levelOff[1];
            *reinterpret_cast<uint4 *>(l1 + ((size_t)((z0 >> 1) + z1) * D1 + ((y0 >> 1) + y1)) * D1 + (x0 >> 1) + w * 16) = ov;
            if (a.useSurf) surf3Dwrite(ov, a.surf[1], (x0 >> 1) + w * 16, (y0 >> 1) + y1, (z0 >> 1) + z1);
        }
    }
    __syncthreads();

    // levels 2..4 inside the brick, from shared memory
    auto reduce = [&](const uint8_t *src, uint8_t *dst, int sx, int sy, int lvl) {
        // src dims (sx, sy, sy) -> dst dims (sx/2, sy/2, sy/2)
        const int dx = sx >> 1, dy = sy >> 1;
        const int n = dx * dy * dy;
        const int Dl = D >> lvl;
        uint8_t *out = a.chain + a.vol.levelOff[lvl];
        for (int o = tid; o < n; o += 256) {
            const int x = o % dx, y = (o / dx) % dy, z = o / (dx * dy);
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < 8; k++)
                s += src[((2 * z + (k >> 2)) * sy + (2 * y + ((k >> 1) & 1))) * sx + 2 * x + (k & 1)];
            const uint8_t v = (uint8_t)box8(s);
            dst[o] = v;
            out[((size_t)((z0 >> lvl) + z) * Dl + ((y0 >> lvl) + y)) * Dl + (x0 >> lvl) + x] = v;
            if (a.useSurf) surf3Dwrite(v, a.surf[lvl], (x0 >> lvl) + x, (y0 >> lvl) + y, (z0 >> lvl) + z);
        }
    };
    if (L > 2) { reduce(sL1, sL2, BX / 2, 8, 2); __syncthreads(); }
    if (L > 3) { reduce(sL2, sL3, BX / 4, 4, 3); __syncthreads(); }
    if (L > 4) { reduce(sL3, sL4, BX / 8, 2, 4); }

    if (!a.tail || L <= 5) return;

    // levels 5.. : last CTA to retire reduces level 4 (D/16)^3 down to the top
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
        sLast = (atomicAdd(a.ticket, 1u) == total - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    for (int lvl = 5; lvl < L; lvl++) {
        const int Ds = D >> (lvl - 1), Dd = D >> lvl;
        const uint8_t *src = a.chain + a.vol.levelOff[lvl - 1];
        uint8_t *dst = a.chain + a.vol.levelOff[lvl];
        const int n = Dd * Dd * Dd;
        for (int o = tid; o < n; o += 256) {
            const int x = o % Dd, y = (o / Dd) % Dd, z = o / (Dd * Dd);
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {                // 4 (y,z) rows, 2 adjacent bytes each
                const uint8_t *p = src + ((size_t)(2 * z + (k >> 1)) * Ds + (2 * y + (k & 1))) * Ds + 2 * x;
                const uint16_t two = __ldcg(reinterpret_cast<const uint16_t *>(p));
                s += (two & 0xFFu) + (two >> 8);
            }
            const uint8_t v = (uint8_t)box8(s);
            dst[o] = v;
            if (a.useSurf) surf3Dwrite(v, a.surf[lvl], x, y, z);
        }
        __threadfence();
        __syncthreads();
    }
    if (tid == 0) *a.ticket = 0;                         // re-arm for the next frame
}

// ---- R32F variant (CRN_VOLUME_R32F): level 0 = 0.0/1.0, every coarser texel the plain mean of its 8 children ----
// Same brick decomposition; this kernel is write-bound for real (4 B/voxel: 64 MB of level 0 at 256^3, 512 MB at 512^3).
template <int WX>
__global__ void __launch_bounds__(256) mip_chain_f32_kernel(MipArgs a) {
    constexpr int BX = WX * 32;
    __shared__ uint32_t sBits[256 * WX];
    __shared__ float sL1[(BX / 2) * 8 * 8];
    __shared__ float sL2[(BX / 4) * 4 * 4];
    __shared__ float sL3[(BX / 8) * 2 * 2];
    __shared__ float sL4[(BX / 16)];
    __shared__ uint32_t sLast;
    const int D = a.vol.dim, L = a.vol.levels;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * BX, y0 = blockIdx.y * 16, z0 = (blockIdx.z + a.zBrick0) * 16;
    const int wordsPerRow = D >> 5;
    float *chain = reinterpret_cast<float *>(a.chain);
    auto level = [&](int l) { return chain + a.vol.levelOff[l] / 4; };
    {
        const int y = tid & 15, z = tid >> 4;
        const uint32_t *src = a.bits + ((size_t)(z0 + z) * D + (y0 + y)) * wordsPerRow + (x0 >> 5);
#pragma unroll
        for (int w = 0; w < WX; w++) sBits[tid * WX + w] = __ldg(src + w);
    }
    __syncthreads();
    if (a.writeLevel0 || a.useSurf) {            // 4 voxels -> one 128-bit store; a warp writes 512 contiguous bytes
        float *l0 = level(0);
        constexpr int QUADS = BX / 4;                       // float4 per row
        for (int s = tid; s < 256 * QUADS; s += 256) {
            const int row = s / QUADS, q = s % QUADS;
            const uint32_t nib = (sBits[row * WX + (q >> 3)] >> ((q & 7) * 4)) & 15u;
            const float4 o = make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f, (nib & 8u) ? 1.0f : 0.0f);
            const int y = row & 15, z = row >> 4;
            if (a.writeLevel0) *reinterpret_cast<float4 *>(l0 + ((size_t)(z0 + z) * D + (y0 + y)) * D + x0 + q * 4) = o;
            if (a.useSurf) surf3Dwrite(o, a.surf[0], (x0 + q * 4) * 4, y0 + y, z0 + z);
        }
    }
    if (L > 1) {                    // level 1 = (number of lit children) / 8, straight from the bits
        const int D1 = D >> 1;
        float *l1 = level(1);
        for (int o = tid; o < (BX / 2) * 64; o += 256) {
            const int x = o % (BX / 2), r1 = o / (BX / 2), y1 = r1 & 7, z1 = r1 >> 3;
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t b = sBits[((2 * z1 + (k >> 1)) * 16 + (2 * y1 + (k & 1))) * WX + (x >> 4)];
                cnt += (b >> ((2 * x) & 31)) & 1u;
                cnt += (b >> ((2 * x + 1) & 31)) & 1u;
            }
            const float v = (float)cnt * 0.125f;
            sL1[o] = v;
            l1[((size_t)((z0 >> 1) + z1) * D1 + ((y0 >> 1) + y1)) * D1 + (x0 >> 1) + x] = v;
            if (a.useSurf) surf3Dwrite(v, a.surf[1], ((x0 >> 1) + x) * 4, (y0 >> 1) + y1, (z0 >> 1) + z1);
        }
    }
    __syncthreads();
    auto reduce = [&](const float *src, float *dst, int sx, int sy, int lvl) {
        const int dx = sx >> 1, dy = sy >> 1, n = dx * dy * dy, Dl = D >> lvl;
        float *out = level(lvl);
        for (int o = tid; o < n; o += 256) {
            const int x = o % dx, y = (o / dx) % dy, z = o / (dx * dy);
            auto S = [&](int ddx, int ddy, int ddz) { return src[((2 * z + ddz) * sy + (2 * y + ddy)) * sx + 2 * x + ddx]; };
            const float v = (((S(0, 0, 0) + S(1, 0, 0)) + (S(0, 1, 0) + S(1, 1, 0))) + ((S(0, 0, 1) + S(1, 0, 1)) + (S(0, 1, 1) + S(1, 1, 1)))) * 0.125f;
            dst[o] = v;
            out[((size_t)((z0 >> lvl) + z) * Dl + ((y0 >> lvl) + y)) * Dl + (x0 >> lvl) + x] = v;
            if (a.useSurf) surf3Dwrite(v, a.surf[lvl], ((x0 >> lvl) + x) * 4, (y0 >> lvl) + y, (z0 >> lvl) + z);
        }
    };
    if (L > 2) { reduce(sL1, sL2, BX / 2, 8, 2); __syncthreads(); }
    if (L > 3) { reduce(sL2, sL3, BX / 4, 4, 3); __syncthreads(); }
    if (L > 4) { reduce(sL3, sL4, BX / 8, 2, 4); }
    if (!a.tail || L <= 5) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) sLast = (atomicAdd(a.ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1u) ? 1u : 0u;
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    for (int lvl = 5; lvl < L; lvl++) {
        const int Ds = D >> (lvl - 1), Dd = D >> lvl;
        const float *src = level(lvl - 1);
        float *dst = level(lvl);
        for (int o = tid; o < Dd * Dd * Dd; o += 256) {
            const int x = o % Dd, y = (o / Dd) % Dd, z = o / (Dd * Dd);
            auto S = [&](int ddx, int ddy, int ddz) { return __ldcg(src + ((size_t)(2 * z + ddz) * Ds + (2 * y + ddy)) * Ds + 2 * x + ddx); };
            const float v = (((S(0, 0, 0) + S(1, 0, 0)) + (S(0, 1, 0) + S(1, 1, 0))) + ((S(0, 0, 1) + S(1, 0, 1)) + (S(0, 1, 1) + S(1, 1, 1)))) * 0.125f;
            dst[o] = v;
            if (a.useSurf) surf3Dwrite(v, a.surf[lvl], x * 4, y, z);
        }
        __threadfence();
        __syncthreads();
    }
    if (tid == 0) *a.ticket = 0;
}

__global__ void __launch_bounds__(256) mip_level_f32_kernel(const float *__restrict__ src, float *__restrict__ dst, int Ds) {
    const int Dd = Ds >> 1;
    const size_t n = (size_t)Dd * Dd * Dd;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(o % Dd), y = (int)((o / Dd) % Dd), z = (int)(o / ((size_t)Dd * Dd));
        auto S = [&](int ddx, int ddy, int ddz) { return src[((size_t)(2 * z + ddz) * Ds + (2 * y + ddy)) * Ds + 2 * x + ddx]; };
        dst[o] = (((S(0, 0, 0) + S(1, 0, 0)) + (S(0, 1, 0) + S(1, 1, 0))) + ((S(0, 0, 1) + S(1, 0, 1)) + (S(0, 1, 1) + S(1, 1, 1)))) * 0.125f;
    }
}

__global__ void __launch_bounds__(256) level_to_surface_f32_kernel(const float *__restrict__ src, cudaSurfaceObject_t surf, int n) {
    const size_t total = (size_t)n * n * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e % n), y = (int)((e / n) % n), z = (int)(e / ((size_t)n * n));
        surf3Dwrite(src[e], surf, x * 4, y, z);
    }
}

// one level from the previous one (used after a slab exchange, crn_finish_mips)
__global__ void __launch_bounds__(256) mip_level_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int Ds) {
    const int Dd = Ds >> 1;
    const size_t n = (size_t)Dd * Dd * Dd;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(o % Dd), y = (int)((o / Dd) % Dd), z = (int)(o / ((size_t)Dd * Dd));
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint8_t *p = src + ((size_t)(2 * z + (k >> 1)) * Ds + (2 * y + (k & 1))) * Ds + 2 * x;
            const uint16_t two = *reinterpret_cast<const uint16_t *>(p);
            s += (two & 0xFFu) + (two >> 8);
        }
        dst[o] = (uint8_t)box8(s);
    }
}

// linear chain level -> its texture-unit copy (after a slab exchange / crn_finish_mips)
__global__ void __launch_bounds__(256) level_to_surface_kernel(const uint8_t *__restrict__ src, cudaSurfaceObject_t surf, int n) {
    const size_t total = (size_t)n * n * n;
    if (n >= 16) {
        const size_t vecs = total / 16;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += (size_t)gridDim.x * blockDim.x) {
            const size_t e = i * 16;
            const int x = (int)(e % n), y = (int)((e / n) % n), z = (int)(e / ((size_t)n * n));
            surf3Dwrite(*reinterpret_cast<const uint4 *>(src + e), surf, x, y, z);
        }
    } else {
        for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
            const int x = (int)(e % n), y = (int)((e / n) % n), z = (int)(e / ((size_t)n * n));
            surf3Dwrite(src[e], surf, x, y, z);
        }
    }
}

// level 0 from the occupancy bits, on demand: the linear copy (crn_read_volume / crn_volume_level_ptr) and / or the
// texture-unit copy (after a slab exchange, which ships the bits, not 8x as many bytes)
template <bool kF32>
__global__ void __launch_bounds__(256) expand_level0_kernel(const uint32_t *__restrict__ bits, uint8_t *__restrict__ linear, cudaSurfaceObject_t surf,
                                                            int D, int useSurf) {
    const size_t words = (size_t)D * D * (D >> 5);
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = __ldg(bits + w);
        const int wpr = D >> 5;
        const int x0 = (int)(w % wpr) * 32, y = (int)((w / wpr) % D), z = (int)(w / ((size_t)wpr * D));
        if (kF32) {
            float *dst = linear ? reinterpret_cast<float *>(linear) + ((size_t)z * D + y) * D + x0 : nullptr;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t nib = (b >> (4 * q)) & 15u;
                const float4 o = make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f, (nib & 8u) ? 1.0f : 0.0f);
                if (dst) reinterpret_cast<float4 *>(dst)[q] = o;
                if (useSurf) surf3Dwrite(o, surf, (x0 + 4 * q) * 4, y, z);
            }
        } else {
            uint8_t *dst = linear ? linear + ((size_t)z * D + y) * D + x0 : nullptr;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t hw = (b >> (16 * h)) & 0xFFFFu;
                uint4 o;
                o.x = spread4(hw & 15u); o.y = spread4((hw >> 4) & 15u); o.z = spread4((hw >> 8) & 15u); o.w = spread4(hw >> 12);
                if (dst) reinterpret_cast<uint4 *>(dst)[h] = o;
                if (useSurf) surf3Dwrite(o, surf, x0 + 16 * h, y, z);
            }
        }
    }
}

__global__ void __launch_bounds__(256) count_bits_kernel(const uint32_t *__restrict__ bits, size_t words, unsigned long long *out) {
    unsigned long long c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x)
        c += __popc(bits[i]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

} // namespace

// tensor map of the linear level 0 (D^3 bytes, x fastest) with a 128 x 16 x 16 box, for the TMA-store variant
static bool make_level0_tensor_map(CUtensorMap *tm, void *level0, int D) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)D, (cuuint64_t)D};
    const cuuint64_t strides[2] = {(cuuint64_t)D, (cuuint64_t)D * D};            // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {128, 16, 16}, estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, level0, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_mips(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, uint8_t *chain, uint32_t *ticket,
                bool writeLevel0, const TexSet *ts) {
    MipArgs a;
    a.vol = vol; a.bits = bits; a.chain = chain; a.ticket = ticket;
    a.bx = vol.dim >= 128 ? 128 : vol.dim;
    a.writeLevel0 = writeLevel0 ? 1 : 0;
    const bool whole = vol.z0 == 0 && vol.z1 == vol.dim;
    a.tail = whole ? 1 : 0;
    a.zBrick0 = vol.z0 / 16;
    a.useSurf = (ts && ts->enabled) ? 1 : 0;
    for (int l = 0; l < kMaxLevels; l++) a.surf[l] = a.useSurf ? ts->surf[l] : 0;
    dim3 grid(vol.dim / a.bx, vol.dim / 16, (vol.z1 - vol.z0) / 16);
    if (vol.texelBytes == 4) {
        if (a.bx == 128) mip_chain_f32_kernel<4><<<grid, 256, 0, st>>>(a);
        else if (a.bx == 64) mip_chain_f32_kernel<2><<<grid, 256, 0, st>>>(a);
        else mip_chain_f32_kernel<1><<<grid, 256, 0, st>>>(a);
        return 1;
    }
    CUtensorMap tm{};
    static const bool wantTma = getenv("CRN_MIPS_TMA") && atoi(getenv("CRN_MIPS_TMA")) != 0;
    if (wantTma && a.bx == 128 && a.writeLevel0 && make_level0_tensor_map(&tm, chain + vol.levelOff[0], vol.dim)) {
        mip_chain_kernel<4, true><<<grid, 256, 0, st>>>(a, tm);
        return 1;
    }
    if (a.bx == 128) mip_chain_kernel<4, false><<<grid, 256, 0, st>>>(a, tm);
    else if (a.bx == 64) mip_chain_kernel<2, false><<<grid, 256, 0, st>>>(a, tm);
    else mip_chain_kernel<1, false><<<grid, 256, 0, st>>>(a, tm);
    return 1;
}

int launch_expand_level0(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, uint8_t *chain, const TexSet *ts) {
    const size_t words = (size_t)vol.dim * vol.dim * (vol.dim >> 5);
    const int blocks = (int)std::min<size_t>((words + 255) / 256, 148 * 16);
    uint8_t *linear = chain ? chain + vol.levelOff[0] : nullptr;
    const cudaSurfaceObject_t surf = ts ? ts->surf[0] : 0;
    if (vol.texelBytes == 4) expand_level0_kernel<true><<<blocks, 256, 0, st>>>(bits, linear, surf, vol.dim, ts ? 1 : 0);
    else expand_level0_kernel<false><<<blocks, 256, 0, st>>>(bits, linear, surf, vol.dim, ts ? 1 : 0);
    return 1;
}

int launch_chain_to_surfaces(cudaStream_t st, const VolumeParams &vol, const uint8_t *chain, const TexSet &ts, int firstLevel) {
    int launches = 0;
    for (int l = firstLevel; l < vol.levels; l++) {
        const int n = vol.levelSize[l];
        const size_t work = (size_t)n * n * n / (n >= 16 ? 16 : 1);
        const int blocks = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
        if (vol.texelBytes == 4) {
            const int b4 = (int)(((size_t)n * n * n + 255) / 256 < 148 * 8 ? ((size_t)n * n * n + 255) / 256 : 148 * 8);
            level_to_surface_f32_kernel<<<b4, 256, 0, st>>>(reinterpret_cast<const float *>(chain + vol.levelOff[l]), ts.surf[l], n);
        } else {
            level_to_surface_kernel<<<blocks, 256, 0, st>>>(chain + vol.levelOff[l], ts.surf[l], n);
        }
        launches++;
    }
    return launches;
}

int launch_finish_mips(cudaStream_t st, const VolumeParams &vol, uint8_t *chain, int firstLevel) {
    int launches = 0;
    for (int l = firstLevel; l < vol.levels; l++) {
        const int Ds = vol.levelSize[l - 1], Dd = Ds / 2;
        const size_t n = (size_t)Dd * Dd * Dd;
        const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
        if (vol.texelBytes == 4)
            mip_level_f32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float *>(chain + vol.levelOff[l - 1]),
                                                        reinterpret_cast<float *>(chain + vol.levelOff[l]), Ds);
        else
            mip_level_kernel<<<blocks, 256, 0, st>>>(chain + vol.levelOff[l - 1], chain + vol.levelOff[l], Ds);
        launches++;
    }
    return launches;
}

int launch_count_bits(cudaStream_t st, const uint32_t *bits, size_t words, unsigned long long *out) {
    cudaMemsetAsync(out, 0, sizeof(unsigned long long), st);
    const int blocks = (int)((words + 255) / 256 < 148 * 8 ? (words + 255) / 256 : 148 * 8);
    count_bits_kernel<<<blocks, 256, 0, st>>>(bits, words, out);
    return 1;
}

} // namespace crn
