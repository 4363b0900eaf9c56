// k_microbench.cu — measures the three ceilings MEASURED_PEAKS.json does not carry and that the
// stages are bounded by (BASELINE.json north_star: "texture/L1 fetch rate for tracing, atomic
// rate for splatting"):
//   0  tex3D trilinear fetch rate, RGBA8_SNORM 32^3 (L1-resident), warp-coherent coordinates
//   1  tex3D trilinear fetch rate, R8 UNORM 256^3 (L2-resident), warp-coherent coordinates
//   2  LDG.32 L1-hit rate, warp-coherent addresses (what explicit filtering issues)
//   3  global atomicOr (RED) rate, addresses spread over a 2 MB bitset
//   4  shared-memory atomicOr rate, spread addresses
//   5  FFMA issue rate (sanity: the instruction-issue ceiling)
//   6  tex2DLayered bilinear, RGBA8_SNORM 32x32x32 layers (the noise texture's layout)
//   7  tex2DLayered bilinear, RG16 UNORM 129x129x128 layers (baked cone-step textures, slice pairs)
//   8  tex2DLayered bilinear, RG8 UNORM 129x129x128 layers
//   9  tex2DLayered bilinear, RGBA8_SNORM, f16x2 return (two registers instead of four)
//  10  tex3D trilinear, R16 UNORM 129^3
//  11  tex3DLod on the mipmapped R8 256^3 volume at LOD 4.5 (two levels blended: "quadrilinear")
//  12  tex3DLod on the mipmapped R8 256^3 volume at LOD 2.5
//  13  tex2DLayered bilinear, RG16F 129x129x128 layers
//  14  tex3D RGBA8_SNORM 32^3 sampled exactly AT slice centres in z (does the filter skip the zero-weight slice?)
//  15  tex2DLayered bilinear, RGBA16_SNORM 161x161x160 layers (64-bit texels, 33 MB: L2-resident), strided: every fetch of a warp
//      lands several texels from the last one (the combined-octave noise lattice's access pattern)
//  16  the same pattern on RGBA8_SNORM 161x161x160 (32-bit texels)
//  17  tex3D trilinear, RG16_SNORM 161^3, the same pattern
//  18  as 15 but L1-friendly steps (the format's filter rate without misses)
// Each returns giga-operations per second (lane-level operations), timed with CUDA events.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../../include/cloud_renderer_b200.h"

namespace {

constexpr int kIters = 256;

__global__ void __launch_bounds__(256) tex_noise_kernel(cudaTextureObject_t tex, float *out, float step) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = (lane & 7) * 0.006f + warp * 0.013f, v = (lane >> 3) * 0.006f + warp * 0.007f, w = warp * 0.003f;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        const float4 t = tex3D<float4>(tex, u, v, w);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        u += step; v += step * 0.7f; w += step * 0.3f;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// 3D fetch whose z coordinate sits exactly on a slice centre: the second slice has weight 0
__global__ void __launch_bounds__(256) tex_noise_zc_kernel(cudaTextureObject_t tex, float *out, float step) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = (lane & 7) * 0.006f + warp * 0.013f, v = (lane >> 3) * 0.006f + warp * 0.007f;
    int layer = warp & 31;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        const float4 t = tex3D<float4>(tex, u, v, ((float)layer + 0.5f) * (1.0f / 32.0f));
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        u += step; v += step * 0.7f; layer = (layer + (i & 1)) & 31;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// bilinear fetch from a layered 2D RGBA8_SNORM texture (the noise texture's layout in the trace kernel)
__global__ void __launch_bounds__(256) tex_layered_kernel(cudaTextureObject_t tex, float *out, float step) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = (lane & 7) * 0.006f + warp * 0.013f, v = (lane >> 3) * 0.006f + warp * 0.007f;
    int layer = warp & 31;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        const float4 t = tex2DLayered<float4>(tex, u, v, layer);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        u += step; v += step * 0.7f; layer = (layer + (i & 1)) & 31;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// two-channel layered texture (baked cone-step slice pairs)
__global__ void __launch_bounds__(256) tex_layered2_kernel(cudaTextureObject_t tex, float *out, float step, int layers) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = 0.1f + (lane & 7) * 0.0008f + (warp % 97) * 0.008f, v = 0.1f + (lane >> 3) * 0.0012f + (warp % 89) * 0.009f;
    int layer = warp % layers;
    float2 acc = make_float2(0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        const float2 t = tex2DLayered<float2>(tex, u, v, layer);
        acc.x += t.x; acc.y += t.y;
        u += step; v += step * 0.7f; layer += (i & 15) == 15; if (layer >= layers) layer = 0;
    }
    if (acc.x + acc.y == 12345.678f) out[0] = acc.x;
}

// RGBA8_SNORM layered, result as two f16x2 registers
__global__ void __launch_bounds__(256) tex_layered_h2_kernel(cudaTextureObject_t tex, float *out, float step) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = (lane & 7) * 0.006f + warp * 0.013f, v = (lane >> 3) * 0.006f + warp * 0.007f;
    int layer = warp & 31;
    uint32_t a0 = 0, a1 = 0;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        uint32_t lo, hi;
        asm volatile("tex.a2d.v2.f16x2.f32 {%0, %1}, [%2, {%3, %4, %5, %5}];" : "=r"(lo), "=r"(hi) : "l"(tex), "r"(layer), "f"(u), "f"(v));
        asm("add.f16x2 %0, %0, %1;" : "+r"(a0) : "r"(lo));
        asm("add.f16x2 %0, %0, %1;" : "+r"(a1) : "r"(hi));
        u += step; v += step * 0.7f; layer = (layer + (i & 1)) & 31;
    }
    if (a0 + a1 == 0x12345678u) out[0] = 1.0f;
}

// four-channel layered texture, every iteration jumps `step` (normalized) in u/v and to another layer
__global__ void __launch_bounds__(256) tex_layered4_kernel(cudaTextureObject_t tex, float *out, float step, int layers) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = 0.1f + (lane & 7) * 0.0013f + (warp % 97) * 0.008f, v = 0.1f + (lane >> 3) * 0.0013f + (warp % 89) * 0.009f;
    int layer = warp % layers;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        const float4 t = tex2DLayered<float4>(tex, u, v, layer);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        u += step; v += step * 0.7f; layer += 3; if (layer >= layers) layer -= layers;
        if (u > 0.9f) u -= 0.8f;
        if (v > 0.9f) v -= 0.8f;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

__global__ void __launch_bounds__(256) tex_vol2_kernel(cudaTextureObject_t tex, float *out, float step) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = 0.1f + (lane & 7) * 0.0013f + (warp % 97) * 0.008f, v = 0.1f + (lane >> 3) * 0.0013f + (warp % 89) * 0.009f,
          w = 0.1f + (warp % 83) * 0.0095f;
    float2 acc = make_float2(0, 0);
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        const float2 t = tex3D<float2>(tex, u, v, w);
        acc.x += t.x; acc.y += t.y;
        u += step; v += step * 0.7f; w += 0.019f;
        if (u > 0.9f) u -= 0.8f;
        if (v > 0.9f) v -= 0.8f;
        if (w > 0.9f) w -= 0.8f;
    }
    if (acc.x + acc.y == 12345.678f) out[0] = acc.x;
}

__global__ void __launch_bounds__(256) tex_lod_kernel(cudaTextureObject_t tex, float *out, float step, float lod) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = 0.1f + (lane & 7) * 0.0008f + (warp % 97) * 0.008f, v = 0.1f + (lane >> 3) * 0.0012f + (warp % 89) * 0.009f,
          w = 0.1f + (warp % 83) * 0.0095f;
    float acc = 0;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        acc += tex3DLod<float>(tex, u, v, w, lod);
        u += step; v += step * 0.7f; w += step * 0.3f;
    }
    if (acc == 12345.678f) out[0] = acc;
}

__global__ void __launch_bounds__(256) tex_vol_kernel(cudaTextureObject_t tex, float *out, float step) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float u = 0.1f + (lane & 7) * 0.0008f + (warp % 97) * 0.008f, v = 0.1f + (lane >> 3) * 0.0012f + (warp % 89) * 0.009f,
          w = 0.1f + (warp % 83) * 0.0095f;
    float acc = 0;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        acc += tex3D<float>(tex, u, v, w);
        u += step; v += step * 0.7f; w += step * 0.3f;
    }
    if (acc == 12345.678f) out[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_kernel(const uint32_t *__restrict__ tab, uint32_t mask, uint32_t *out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t idx = (uint32_t)(warp * 131 + (lane >> 2));      // 8 distinct words per warp: a few sectors per request
    uint32_t acc = 0;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        acc += __ldg(tab + (idx & mask));
        idx += 37;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(256) red_kernel(uint32_t *bits, uint32_t mask) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t idx = t * 2654435761u;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        atomicOr(bits + ((idx >> 7) & mask), 1u << (idx & 31));
        idx += 0x9E3779B9u;
    }
}

__global__ void __launch_bounds__(256) satom_kernel(uint32_t *out) {
    __shared__ uint32_t s[4096];
    for (int i = threadIdx.x; i < 4096; i += 256) s[i] = 0;
    __syncthreads();
    uint32_t idx = threadIdx.x * 2654435761u;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        atomicOr(&s[(idx >> 9) & 4095], 1u << (idx & 31));
        idx += 0x9E3779B9u;
    }
    __syncthreads();
    if (s[threadIdx.x] == 0x12345u) out[0] = 1;
}

__global__ void __launch_bounds__(256) ffma_kernel(float *out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
    for (int i = 0; i < kIters; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) out[0] = s;
}

template <class F>
double time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

cudaTextureObject_t make_tex(cudaArray_t arr, bool wrap) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = wrap ? cudaAddressModeWrap : cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    cudaTextureObject_t t = 0;
    cudaCreateTextureObject(&t, &rd, &td, nullptr);
    return t;
}

} // namespace

extern "C" int crn_microbench(int device, int which, double *gops) {
    if (!gops) return CRN_ERR_INVALID_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count) return CRN_ERR_NO_DEVICE;
    cudaSetDevice(device);
    const int blocks = 148 * 32, threads = 256;
    const double lanes = (double)blocks * threads * kIters;
    float *dOut = nullptr;
    cudaMalloc(&dOut, 256);
    double ms = 0, ops = lanes;
    if (which == 0 || which == 1 || which == 14) {
        const int n = which != 1 ? 32 : 256;
        cudaChannelFormatDesc cd = which != 1 ? cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindSigned)
                                              : cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
        cudaArray_t arr = nullptr;
        cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, n));
        const size_t texel = which != 1 ? 4 : 1;
        std::vector<uint8_t> h((size_t)n * n * n * texel);
        for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(h.data(), n * texel, n, n);
        cp.dstArray = arr; cp.extent = make_cudaExtent(n, n, n); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        cudaTextureObject_t tex = make_tex(arr, which != 1);
        if (which == 14) ms = time_ms([&] { tex_noise_zc_kernel<<<blocks, threads>>>(tex, dOut, 0.004f); }, 5);
        else if (which == 0) ms = time_ms([&] { tex_noise_kernel<<<blocks, threads>>>(tex, dOut, 0.004f); }, 5);
        else ms = time_ms([&] { tex_vol_kernel<<<blocks, threads>>>(tex, dOut, 0.0009f); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    } else if (which == 6) {
        const int n = 32;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindSigned);
        cudaArray_t arr = nullptr;
        cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, n), cudaArrayLayered);
        std::vector<uint8_t> h((size_t)n * n * n * 4);
        for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(h.data(), n * 4, n, n);
        cp.dstArray = arr; cp.extent = make_cudaExtent(n, n, n); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        cudaTextureObject_t tex = make_tex(arr, true);
        ms = time_ms([&] { tex_layered_kernel<<<blocks, threads>>>(tex, dOut, 0.004f); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    } else if (which == 7 || which == 8 || which == 13) {
        const int n = 129, layers = 128;
        const int bits = which == 8 ? 8 : 16;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(bits, bits, 0, 0, which == 13 ? cudaChannelFormatKindFloat : cudaChannelFormatKindUnsigned);
        cudaArray_t arr = nullptr;
        cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, layers), cudaArrayLayered);
        const size_t texel = bits / 4;
        std::vector<uint8_t> h((size_t)n * n * layers * texel);
        for (size_t i = 0; i < h.size(); i++) h[i] = which == 13 ? (uint8_t)((i & 1) ? 0x3c : (i * 2654435761u >> 24)) : (uint8_t)(i * 2654435761u >> 24);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(h.data(), n * texel, n, n);
        cp.dstArray = arr; cp.extent = make_cudaExtent(n, n, layers); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear; td.readMode = which == 13 ? cudaReadModeElementType : cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
        cudaTextureObject_t tex = 0;
        cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        ms = time_ms([&] { tex_layered2_kernel<<<blocks, threads>>>(tex, dOut, 0.0009f, layers); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    } else if (which >= 15 && which <= 18) {
        const int n = 161, layers = 160;
        const bool vol = which == 17;
        const int bits = which == 16 ? 8 : 16, chans = vol ? 2 : 4;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(bits, bits, chans == 4 ? bits : 0, chans == 4 ? bits : 0, cudaChannelFormatKindSigned);
        cudaArray_t arr = nullptr;
        const int depth = vol ? n : layers;
        if (cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, depth), vol ? 0 : cudaArrayLayered) != cudaSuccess) { cudaFree(dOut); return CRN_ERR_CUDA; }
        const size_t texel = (size_t)bits / 8 * chans;
        std::vector<uint8_t> h((size_t)n * n * depth * texel);
        for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(h.data(), n * texel, n, n);
        cp.dstArray = arr; cp.extent = make_cudaExtent(n, n, depth); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        cudaTextureObject_t tex = make_tex(arr, false);
        if (vol) ms = time_ms([&] { tex_vol2_kernel<<<blocks, threads>>>(tex, dOut, 0.031f); }, 5);
        else ms = time_ms([&] { tex_layered4_kernel<<<blocks, threads>>>(tex, dOut, which == 18 ? 0.0009f : 0.031f, which == 18 ? 1 : layers); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    } else if (which == 9) {
        const int n = 32;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindSigned);
        cudaArray_t arr = nullptr;
        cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, n), cudaArrayLayered);
        std::vector<uint8_t> h((size_t)n * n * n * 4);
        for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(h.data(), n * 4, n, n);
        cp.dstArray = arr; cp.extent = make_cudaExtent(n, n, n); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        cudaTextureObject_t tex = make_tex(arr, true);
        ms = time_ms([&] { tex_layered_h2_kernel<<<blocks, threads>>>(tex, dOut, 0.004f); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    } else if (which == 10) {
        const int n = 129;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
        cudaArray_t arr = nullptr;
        cudaMalloc3DArray(&arr, &cd, make_cudaExtent(n, n, n));
        std::vector<uint8_t> h((size_t)n * n * n * 2);
        for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(h.data(), n * 2, n, n);
        cp.dstArray = arr; cp.extent = make_cudaExtent(n, n, n); cp.kind = cudaMemcpyHostToDevice;
        cudaMemcpy3D(&cp);
        cudaTextureObject_t tex = make_tex(arr, false);
        ms = time_ms([&] { tex_vol_kernel<<<blocks, threads>>>(tex, dOut, 0.0009f); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeArray(arr);
    } else if (which == 11 || which == 12) {
        const int n = 256, L = 9;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
        cudaMipmappedArray_t marr = nullptr;
        cudaMallocMipmappedArray(&marr, &cd, make_cudaExtent(n, n, n), L);
        for (int l = 0; l < L; l++) {
            const int s = n >> l;
            cudaArray_t lvl = nullptr;
            cudaGetMipmappedArrayLevel(&lvl, marr, l);
            std::vector<uint8_t> h((size_t)s * s * s);
            for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
            cudaMemcpy3DParms cp{};
            cp.srcPtr = make_cudaPitchedPtr(h.data(), s, s, s);
            cp.dstArray = lvl; cp.extent = make_cudaExtent(s, s, s); cp.kind = cudaMemcpyHostToDevice;
            cudaMemcpy3D(&cp);
        }
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeMipmappedArray; rd.res.mipmap.mipmap = marr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear; td.mipmapFilterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
        td.minMipmapLevelClamp = 0.0f; td.maxMipmapLevelClamp = (float)(L - 1);
        cudaTextureObject_t tex = 0;
        cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        const float lod = which == 11 ? 4.5f : 2.5f;
        ms = time_ms([&] { tex_lod_kernel<<<blocks, threads>>>(tex, dOut, 0.0009f, lod); }, 5);
        cudaDestroyTextureObject(tex);
        cudaFreeMipmappedArray(marr);
    } else if (which == 2) {
        uint32_t *tab = nullptr;
        cudaMalloc(&tab, 64 << 10);
        cudaMemset(tab, 1, 64 << 10);
        ms = time_ms([&] { ldg_kernel<<<blocks, threads>>>(tab, (64 << 10) / 4 - 1, (uint32_t *)dOut); }, 5);
        cudaFree(tab);
    } else if (which == 3) {
        uint32_t *bits = nullptr;
        cudaMalloc(&bits, 2 << 20);
        cudaMemset(bits, 0, 2 << 20);
        ms = time_ms([&] { red_kernel<<<blocks, threads>>>(bits, (2 << 20) / 4 - 1); }, 5);
        cudaFree(bits);
    } else if (which == 4) {
        ms = time_ms([&] { satom_kernel<<<blocks, threads>>>((uint32_t *)dOut); }, 5);
    } else if (which == 5) {
        ms = time_ms([&] { ffma_kernel<<<blocks, threads>>>(dOut, 1.0001f, 0.5f); }, 5);
        ops = lanes * 8;
    } else {
        cudaFree(dOut);
        return CRN_ERR_INVALID_ARG;
    }
    cudaFree(dOut);
    if (cudaGetLastError() != cudaSuccess) return CRN_ERR_CUDA;
    *gops = ops / (ms * 1e-3) / 1e9;
    return CRN_OK;
}
