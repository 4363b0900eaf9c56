// k_trace.cu — stage 3: voxel cone tracing of the noise-marched billboards, per pixel.
//
// Replaces ConeTraceShader::coneTrace's instanced, alpha-blended draw
// (src/Shaders/ConeTraceShader.cpp:22-75), its GPU program (res/conetrace_frag.glsl:122-201
// with helpers :48-120, vertex stage res/billboard_vert_instanced.glsl) and the frame state
// around it: the clear (src/main.cpp:112-113), the optional sun pass
// (src/Shaders/SunShader.cpp:7-42, res/sun_frag.glsl:14-28) and the fixed-function
// SRC_ALPHA / ONE_MINUS_SRC_ALPHA blend (src/main.cpp:94-95).
//
// The reference shades one fragment per (billboard, covered pixel) and lets the ROP blend
// them back to front.  Here one thread owns one pixel, a warp an 8x4 patch, a CTA a 16x16
// tile, and walks the tile's billboard list (k_bin.cu) FRONT TO BACK carrying the
// transmittance T of everything already composited:
//     C += T * a * src;   T *= (1 - a);       final = C + T * background
// which is the same sum the back-to-front blend produces, re-associated.  A warp stops as
// soon as every pixel's T is under the caller's cutoff (vote), so the deep interior of the
// cloud — hundreds of layers of overdraw at 20k billboards — is never shaded.  With
// cutoff = 0 every fragment is shaded, as in the reference.
//
// Two samplers, chosen per frame (crn_trace_params.sampler):
//   CRN_SAMPLER_TEXTURE  the texture units: one mipmapped R8/R32F texture sampled with tex3DLod at the
//                        step's fractional LOD (LINEAR_MIPMAP_LINEAR, as the reference's sampler state,
//                        src/CloudVolume.cpp:19-23) and the RGBA8_SNORM noise texture; the kernel is
//                        bound by the texture pipe.
//   CRN_SAMPLER_EXPLICIT no texture units: level 0 is read from the 1-bit occupancy set (2 MB at
//                        256^3), coarser levels from the linear chain, the noise from a pre-decoded
//                        (g,a) float2 copy; trilinear / mip-linear weights follow GL 4.4 §8.14 in full
//                        float precision (134 dB against the oracle, 3x slower: issue-bound).
// traceCone's per-step height, LOD split and weight are identical for every fragment and are
// precomputed on the host (TraceParams::steps).  With the texture sampler the coarse steps read
// per-frame baked step textures (k_conebake.cu: one bilinear pass + a z blend instead of a mip-linear
// 3D fetch, exact); the remaining fine steps are grouped, and a group whose bit in the need-code grid
// (k_conebake.cu: the group's texel footprint against the non-zero bit volumes of k_skipmask.cu) is clear is skipped
// exactly.  noise3D's octaves 1..3 come pre-summed from the combined-octave lattice (k_noiselat.cu, exact) in the fast variant.
#include "crn_internal.cuh"

#include <cstdlib>

namespace crn {

namespace {

struct TraceArgs {
    VolumeParams vol;
    const BoardRec *recs;
    const uint32_t *tileOff, *tileCnt, *tileList;
    int tilesX, tilesY;
    const uint32_t *bits;
    const uint8_t *chain;
    const uint32_t *bitsA;        // CRN_VOLUME_RG8: occupancy channel (level 0 bits + R8 chain), else nullptr
    const uint8_t *chainA;
    const float2 *noise;          // (g, a) decoded, dim^3
    void *image;
    int format;
    unsigned long long *stats;
    float invRange[3];            // 1 / range per axis
    float invDim;                 // 1 / voxelDim
    const uint8_t *code;          // need-code grid (k_conebake.cu): bit g = group g may contribute from this cell; or nullptr
    const uint32_t *order;        // launch order of the tiles, longest list first (k_bin.cu)
    float4 *segPartial;           // small frames: (colour, alpha, transmittance) of every list segment, [tile][segment][256 pixels]
    uint32_t *segArrived;         // ... and how many segments of a (tile, quarter) have finished
};

__device__ __forceinline__ float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

struct Lin {                       // one axis of a linear filter footprint
    int i0, i1;
    float a;
};

__device__ __forceinline__ Lin clamp_axis(float t, int n) {        // CLAMP_TO_EDGE
    Lin l;
    t -= 0.5f;
    const float fl = floorf(t);
    l.a = t - fl;
    const float c = fminf(fmaxf(fl, -1.0f), (float)n);              // keeps NaN/inf away from the cast
    const int i = (int)c;
    l.i0 = min(max(i, 0), n - 1);
    l.i1 = min(max(i + 1, 0), n - 1);
    return l;
}

__device__ __forceinline__ Lin repeat_axis(float s, int n) {        // REPEAT
    Lin l;
    s -= floorf(s);
    const float t = s * (float)n - 0.5f;
    const float fl = floorf(t);
    l.a = t - fl;
    int i = (int)fl;                                                // [-1, n-1]
    l.i0 = i < 0 ? i + n : i;
    l.i1 = l.i0 + 1 == n ? 0 : l.i0 + 1;
    return l;
}

__device__ __forceinline__ float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }

// level 0 from the occupancy bits; returns the filtered UNORM value in [0,1]
__device__ __forceinline__ float sample_bits(const uint32_t *__restrict__ bits, int D, float px, float py, float pz) {
    const Lin X = clamp_axis(px, D), Y = clamp_axis(py, D), Z = clamp_axis(pz, D);
    const int wpr = D >> 5;
    const int w0 = X.i0 >> 5, w1 = X.i1 >> 5, s0 = X.i0 & 31, s1 = X.i1 & 31;
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = (k & 1) ? Y.i1 : Y.i0, z = (k & 2) ? Z.i1 : Z.i0;
        const uint32_t *row = bits + (size_t)(z * D + y) * wpr;
        const uint32_t a = __ldg(row + w0);
        const uint32_t b = (w1 == w0) ? a : __ldg(row + w1);
        const float v0 = (float)((a >> s0) & 1u), v1 = (float)((b >> s1) & 1u);
        c[k] = lerpf(v0, v1, X.a);
    }
    return lerpf(lerpf(c[0], c[1], Y.a), lerpf(c[2], c[3], Y.a), Z.a);
}

// level >= 1 from the R8 chain; returns the filtered UNORM value in [0,1]
__device__ __forceinline__ float sample_bytes(const uint8_t *__restrict__ t, int n, float px, float py, float pz) {
    const Lin X = clamp_axis(px, n), Y = clamp_axis(py, n), Z = clamp_axis(pz, n);
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = (k & 1) ? Y.i1 : Y.i0, z = (k & 2) ? Z.i1 : Z.i0;
        const uint8_t *row = t + (size_t)(z * n + y) * n;
        c[k] = lerpf((float)__ldg(row + X.i0), (float)__ldg(row + X.i1), X.a);
    }
    return lerpf(lerpf(c[0], c[1], Y.a), lerpf(c[2], c[3], Y.a), Z.a) * (1.0f / 255.0f);
}

// level >= 1 of an R32F chain
__device__ __forceinline__ float sample_floats(const float *__restrict__ t, int n, float px, float py, float pz) {
    const Lin X = clamp_axis(px, n), Y = clamp_axis(py, n), Z = clamp_axis(pz, n);
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = (k & 1) ? Y.i1 : Y.i0, z = (k & 2) ? Z.i1 : Z.i0;
        const float *row = t + (size_t)(z * n + y) * n;
        c[k] = lerpf(__ldg(row + X.i0), __ldg(row + X.i1), X.a);
    }
    return lerpf(lerpf(c[0], c[1], Y.a), lerpf(c[2], c[3], Y.a), Z.a);
}

__device__ __forceinline__ float sample_level(const TraceArgs &a, const uint32_t *__restrict__ bits, const uint8_t *__restrict__ chain,
                                              int l, float px, float py, float pz) {
    if (l == 0) return sample_bits(bits, a.vol.dim, px, py, pz);            // level 0 is 0/1 in every format
    const float s = 1.0f / (float)(1 << l);
    if (a.vol.texelBytes == 4)
        return sample_floats(reinterpret_cast<const float *>(chain + a.vol.levelOff[l]), a.vol.levelSize[l], px * s, py * s, pz * s);
    return sample_bytes(chain + a.vol.levelOff[l], a.vol.levelSize[l], px * s, py * s, pz * s);
}

// texture(noiseMap, uvw): green and alpha channels only (the shader uses nothing else)
__device__ __forceinline__ float2 sample_noise(const float2 *__restrict__ nz, int n, float u, float v, float w) {
    const Lin X = repeat_axis(u, n), Y = repeat_axis(v, n), Z = repeat_axis(w, n);
    float2 c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = (k & 1) ? Y.i1 : Y.i0, z = (k & 2) ? Z.i1 : Z.i0;
        const float2 *row = nz + (size_t)(z * n + y) * n;
        const float2 t0 = __ldg(row + X.i0), t1 = __ldg(row + X.i1);
        c[k] = make_float2(lerpf(t0.x, t1.x, X.a), lerpf(t0.y, t1.y, X.a));
    }
    const float2 c0 = make_float2(lerpf(c[0].x, c[1].x, Y.a), lerpf(c[0].y, c[1].y, Y.a));
    const float2 c1 = make_float2(lerpf(c[2].x, c[3].x, Y.a), lerpf(c[2].y, c[3].y, Y.a));
    return make_float2(lerpf(c0.x, c1.x, Z.a), lerpf(c0.y, c1.y, Z.a));
}

// sun_frag.glsl:14-28 on the sun quad; returns false when the pixel is not covered / discarded
__device__ __forceinline__ bool sun_fragment(const TraceParams &tp, const ViewParams &cam, float ndcx, float ndcy, int px, int py,
                                             float col[4]) {
    const BoardRect r = tp.sunRect;
    if (!(r.i1 >= r.i0) || px < r.i0 || px > r.i1 || py < r.j0 || py > r.j1) return false;
    const BoardRec &q = tp.sunRec;
    const float wz = cam.ortho ? 1.0f : -q.zv;
    const float xv = cam.ortho ? (ndcx - cam.P[12]) / cam.P[0] : ndcx * wz / cam.P[0];
    const float yv = cam.ortho ? (ndcy - cam.P[13]) / cam.P[5] : ndcy * wz / cam.P[5];
    const float u = xv - q.xv, v = yv - q.yv;
    if (!(fabsf(u) < q.r && fabsf(v) < q.r)) return false;
    const float dist = sqrtf(u * u + v * v);                 // |fragPos - center| on the quad's plane
    if (dist < tp.sun.innerRadius) {
        col[0] = tp.sun.innerColor[0]; col[1] = tp.sun.innerColor[1]; col[2] = tp.sun.innerColor[2]; col[3] = 1.0f;
        return true;
    }
    const float scale = (dist - tp.sun.innerRadius) / (tp.sun.outerRadius - tp.sun.innerRadius);
    if (scale > 0.99f) return false;
#pragma unroll
    for (int k = 0; k < 3; k++) col[k] = tp.sun.outerColor[k] * scale + tp.sun.innerColor[k] * (1.0f - scale);
    col[3] = 1.0f - scale;
    return true;
}

// kTex = false: explicit filtering (CRN_SAMPLER_EXPLICIT);  kTex = true: texture units (CRN_SAMPLER_TEXTURE)
constexpr int kFastBaked = 16;        // the fast variant unrolls (and handles at most) this many baked steps (32^3 / 64^3 volumes bake all 16)
// resident CTAs per SM the kernel is compiled for (register cap = 65536 / 64 / MINB).  Measured at C3 (trace ms):
// round 1: 12: 4.82, 14: 4.73, 16: 4.39, 18: 4.36, 20: 4.52, 24: 4.60; with the noise lattice: 10: 2.73, 12: 2.74, 16: 2.49,
// 20: 2.52; final kernel: 14: 2.31, 16: 2.22, 18: 2.19, 20: 2.24 -> 18 (56 registers) for the fast variant, 16 for the generic one
#ifndef CRN_TRACE_MINB
#define CRN_TRACE_MINB 18
#endif
#ifndef CRN_TRACE_GENERIC_MINB
#define CRN_TRACE_GENERIC_MINB 16
#endif
#ifndef CRN_TRACE_THREADS
#define CRN_TRACE_THREADS 64
#endif
constexpr int kTraceThreads = CRN_TRACE_THREADS;     // 2 warp patches per CTA: measured best (256: 5.87 ms, 128: 5.83, 64: 5.79 at C3)

// kGate: the paper variant's `if (sampleColor.a > 0)` on a second (occupancy) chain, CRN_VOLUME_RG8 only
template <bool kTex, bool kStats, bool kGate>
__global__ void __launch_bounds__(kTraceThreads, CRN_TRACE_GENERIC_MINB) trace_kernel(const __grid_constant__ TraceArgs a, const __grid_constant__ ViewParams cam,
                                                    const __grid_constant__ TraceParams tp, const __grid_constant__ TexSet ts) {
    // a 16x16 tile is 8 warp patches; a CTA carries kTraceThreads/32 of them, so a slow patch holds up fewer warps
    constexpr int kWarps = kTraceThreads / 32, kSplit = 8 / kWarps;
    const int tile = (int)a.order[blockIdx.x / kSplit];
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) + (blockIdx.x % kSplit) * kWarps;
    const int tx = tile % a.tilesX, ty = tile / a.tilesX;
    const int px = tx * kTile + (warp & 1) * 8 + (lane & 7);
    const int py = ty * kTile + (warp >> 1) * 4 + (lane >> 3);
    // rows outside [row0,row1) belong to another rank (image-space sharding)
    if (tp.ilvCount > 1 && (ty % tp.ilvCount) != tp.ilvIndex) return;
    const bool valid = px < cam.W && py < cam.H && py >= tp.row0 && py < tp.row1;
    if (__all_sync(0xFFFFFFFFu, !valid)) return;

    const float ndcx = ((float)px + 0.5f) / (float)cam.W * 2.0f - 1.0f;
    const float ndcy = ((float)py + 0.5f) / (float)cam.H * 2.0f - 1.0f;
    const float invP0 = 1.0f / cam.P[0], invP5 = 1.0f / cam.P[5];

    // background = clear colour, then the sun pass blended over it
    float bg[4] = {tp.bg[0], tp.bg[1], tp.bg[2], tp.bg[3]};
    if (tp.p.drawSun) {
        float sc[4];
        if (sun_fragment(tp, cam, ndcx, ndcy, px, py, sc)) {
            const float sa = saturatef(sc[3]);
#pragma unroll
            for (int k = 0; k < 4; k++) bg[k] = saturatef(sc[k]) * sa + bg[k] * (1.0f - sa);
        }
    }

    // quantizeFramebuffer: the reference's 8-bit window framebuffer.  dst starts as the quantised background and every
    // blend result goes back through 8 bits, so the list is walked BACK TO FRONT (the reference's draw order) and there
    // is no transmittance to terminate on.
    const bool fb8 = tp.p.quantizeFramebuffer != 0;
    auto q8 = [](float x) { return floorf(saturatef(x) * 255.0f + 0.5f) / 255.0f; };
    float dst8[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (fb8) {
#pragma unroll
        for (int k = 0; k < 4; k++) dst8[k] = q8(tp.bg[k]);
        if (tp.p.drawSun) {
            float sc[4];
            if (sun_fragment(tp, cam, ndcx, ndcy, px, py, sc)) {
                const float sa = saturatef(sc[3]);
#pragma unroll
                for (int k = 0; k < 4; k++) dst8[k] = q8(saturatef(sc[k]) * sa + dst8[k] * (1.0f - sa));
            }
        }
    }
    float C[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float T = 1.0f;
    unsigned long long nFrag = 0, nCone = 0, nNoise = 0, nSkip = 0, nFetch = 0, nBakedF = 0, nLat = 0, nCode = 0;

    const uint32_t cnt = tp.active ? a.tileCnt[tile] : 0;
    const uint32_t off = cnt ? a.tileOff[tile] : 0;
    const float cutoff = tp.p.transmittanceCutoff;
    const int D = a.vol.dim;
    (void)D;
    const int nzDim = tp.noiseDim;
    const float invAdjust = 1.0f / tp.p.adjustSize, invStep = 1.0f / tp.p.stepSize;
    const float noiseSpan = (float)(tp.p.maxNoiseSteps - tp.p.minNoiseSteps);
    // viewRay = normalize(V[0][2], V[1][2], V[2][2]) = cam.nrm (res/conetrace_frag.glsl:138)
    const float rx = cam.nrm[0], ry = cam.nrm[1], rz = cam.nrm[2];

    for (uint32_t e = 0; e < cnt; e++) {
        const bool live = valid && (fb8 || T > cutoff);                 // T starts at 1: cutoff 0 keeps everything live
        if (__all_sync(0xFFFFFFFFu, !live)) break;                      // early ray termination, whole patch
        const uint32_t k = a.tileList[off + (fb8 ? cnt - 1u - e : e)];  // the list is front to back; the framebuffer mode needs far first
        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(&a.recs[k].cx));
        const float4 r1 = __ldg(reinterpret_cast<const float4 *>(&a.recs[k].xv));
        const float radius = r0.w;
        // pixel centre on the quad's plane, relative to the quad's centre
        const float wz = cam.ortho ? 1.0f : -r1.z;
        const float xv = cam.ortho ? (ndcx - cam.P[12]) * invP0 : ndcx * wz * invP0;
        const float yv = cam.ortho ? (ndcy - cam.P[13]) * invP5 : ndcy * wz * invP5;
        const float u = xv - r1.x, v = yv - r1.y;
        const float d2 = u * u + v * v, r2 = radius * radius;
        const float h2 = r2 - d2;                                       // half chord squared
        const float invr = 1.0f / radius;

        float col[4];
        bool shade = live && fabsf(u) < radius && fabsf(v) < radius;    // inside the quad
        if (tp.p.showQuad) {                                            // conetrace_frag.glsl:124-134
            if (shade) {
                const float sc = sqrtf(fmaxf(0.0f, h2)) * invr;              // sqrt(1 - d^2/r^2)
                const float ftx = (u * invr + 1.0f) * 0.5f, fty = (v * invr + 1.0f) * 0.5f;
                const bool border = ftx < 0.01f || fty < 0.01f || ftx > 0.99f || fty > 0.99f;
                col[0] = col[1] = col[2] = col[3] = border ? 1.0f : sc;
            }
        } else {
            // discards: raySphereIntersect's disc = 4*(r^2 - d^2) < 0.01 (noise on), and
            // sphereContrib = sqrt(1 - d^2/r^2) < 0.01 (cone trace on)
            if (tp.p.doNoiseSample) shade = shade && !(4.0f * h2 < 0.01f);
            if (tp.p.doConeTrace) shade = shade && !(h2 < 1.0e-4f * r2);   // sqrt(1 - d2/r2) < 0.01
            if (!__any_sync(0xFFFFFFFFu, shade)) continue;
            const float h = sqrtf(fmaxf(h2, 0.0f));
            // fragPos - center, and the two sphere hits along the view ray: near = -h, far = +h
            const float ox = u * cam.right[0] + v * cam.up[0];
            const float oy = u * cam.right[1] + v * cam.up[1];
            const float oz = u * cam.right[2] + v * cam.up[2];
            col[0] = col[1] = col[2] = col[3] = 0.0f;

            if (tp.p.doNoiseSample) {                                   // conetrace_frag.glsl:137-174
                float ux = (ox - rx * h) * invr, uy = (oy - ry * h) * invr, uz = (oz - rz * h) * invr;        // unitTex
                float tx = (r0.x + ox - rx * h) * invAdjust, tyy = (r0.y + oy - ry * h) * invAdjust,
                      tz = (r0.z + oz - rz * h) * invAdjust;                                                // localTexNear
                const float len = 2.0f * h * invAdjust;
                float iSteps = fminf(len * invStep, noiseSpan) + (float)tp.p.minNoiseSteps;
                const float inv = 1.0f / (iSteps - 1.0f);
                const float dxs = rx * len * inv, dys = ry * len * inv, dzs = rz * len * inv;              // localTexDelta
                float opacity = 0.0f, light = 0.0f;
                const int nIter = shade ? (int)ceilf(iSteps) : 0;
                const int nMax = __reduce_max_sync(0xFFFFFFFFu, nIter);
                for (int i = 0; i < nMax; i++) {
                    if (i < nIter) {
                        float ng = 0.0f, na = 0.0f;
                        const float cx = tx, cy = tyy, cz = tz, vx = ux, vy = uy, vz = uz;
                        for (int o = 0; o < tp.p.numOctaves; o++) {     // noise3D, :103-120: texture(noiseMap, (uv + offset_o) * freq_o) * pers_o
                            const float f = tp.octFreq[o], b = tp.octBias[o];
                            float2 s;
                            if constexpr (kTex) {
                                // one bilinear pass on layer floor(z) returns (g,a) of slices z and z+1; blend them here
                                const float wz = fmaf(cz, tp.octFreqZ[o], tp.octBiasZ[o]);
                                const float fl = floorf(wz), az = wz - fl;
                                int layer = (int)fl;
                                if (tp.noiseMask >= 0) layer &= tp.noiseMask;
                                else { layer %= nzDim; if (layer < 0) layer += nzDim; }
                                const float4 t = tex2DLayered<float4>(ts.noise, fmaf(cx, f, b), fmaf(cy, f, b), layer);
                                s = make_float2(fmaf(az, t.z - t.x, t.x), fmaf(az, t.w - t.y, t.y));
                            } else {
                                s = sample_noise(a.noise, nzDim, fmaf(cx, f, b), fmaf(cy, f, b), fmaf(cz, f, b));
                            }
                            ng = fmaf(tp.octPers[o], s.x, ng);
                            na = fmaf(tp.octPers[o], s.y, na);
                        }
                        na = fabsf(na);
                        const float uu = vx * vx + vy * vy + vz * vz;
                        ng += vy * rsqrtf(uu);                          // noiseCell.xyz += normalize(unitTex)
                        opacity = fmaf(na, 1.0f - uu, opacity);
                        light += saturatef(ng * 0.5f + 0.5f);
                        tx += dxs; tyy += dys; tz += dzs;
                        ux += dxs; uy += dys; uz += dzs;            // (sic) tex-space delta on the unit-sphere coord
                        if (kStats) { nNoise += tp.p.numOctaves; if (__float_as_int(r1.w) & kRecInLattice) nLat++; }
                    }
                }
                const float c = tp.p.minNoiseColor + tp.p.noiseColorScale * light * inv;
                const float alpha = 1.0f - sqrtf(d2) * invr;                   // 1 - length(fragTex - 0.5) * 2
                col[0] = col[1] = col[2] = c;
                col[3] = saturatef(opacity * tp.p.noiseOpacity * inv) * alpha;
            }

            if (tp.p.doConeTrace) {                                     // conetrace_frag.glsl:176-200, traceCone :64-79
                // start on the camera-facing sphere surface: fragPos + n * r * sphereContrib = center + o + n*h
                const float wx3 = r0.x + ox + rx * h, wy3 = r0.y + oy + ry * h, wz3 = r0.z + oz + rz * h;
                // calculateVoxelLerp / voxelDim: the normalized texture coordinate, identical on every level
                const float nx = (wx3 - a.vol.xB[0]) * a.invRange[0];
                const float ny = (wy3 - a.vol.yB[0]) * a.invRange[1];
                const float nz = (wz3 - a.vol.zB[0]) * a.invRange[2];
                float ex = tp.lightPos[0] - wx3, ey = tp.lightPos[1] - wy3, ez = tp.lightPos[2] - wz3;
                const float il = rsqrtf(ex * ex + ey * ey + ez * ez) * a.invDim;      // normalize(dir) / voxelDim
                ex *= il; ey *= il; ez *= il;
                float indirect = 0.0f;
                // One byte decides every group of fine steps: bit g of the start cell's need code is clear when no start
                // position inside that cell can reach a non-zero texel with group g (all-zero footprints contribute exactly 0).
                // Start positions outside the volume have no cell: everything is fetched (CLAMP_TO_EDGE lookups).
                uint32_t code = 0xFFFFFFFFu;
                if (a.code && tp.nGroups > 0) {
                    const int ix = __float2int_rd(nx * tp.codeDimF), iy = __float2int_rd(ny * tp.codeDimF), iz = __float2int_rd(nz * tp.codeDimF);
                    const uint32_t G = (uint32_t)tp.codeDim;
                    if (shade && (uint32_t)ix < G && (uint32_t)iy < G && (uint32_t)iz < G)
                        code = __ldg(a.code + ((uint32_t)iz * G + (uint32_t)iy) * G + (uint32_t)ix) | ~((1u << kCodeGroups) - 1u);
                }
                // group and step loops are warp-uniform (constants come from uniform registers); lanes take part by predicate
                auto coneGroup = [&](const int g) {
                    const ConeGroup &gr = tp.groups[g];
                    const bool need = shade && ((code >> (g < 31 ? g : 31)) & 1u);
                    if (kStats && shade && !need) nSkip += gr.count;
                    if (!__any_sync(0xFFFFFFFFu, need)) return;
#pragma unroll 2
                    for (int i = gr.first; i < gr.first + gr.count; i++) {
                        const ConeStep &st = tp.steps[i];
                        if (need) {
                            const float sx = fmaf(st.height, ex, nx), sy = fmaf(st.height, ey, ny), sz = fmaf(st.height, ez, nz);
                            float s;
                            if constexpr (kTex) {
                                s = tex3DLod<float>(ts.vol, sx, sy, sz, st.lod);        // quadrilinear in one fetch
                            } else {
                                const float fd = (float)D;
                                s = sample_level(a, a.bits, a.chain, st.level0, sx * fd, sy * fd, sz * fd);
                                if (st.frac != 0.0f) s = lerpf(s, sample_level(a, a.bits, a.chain, st.level0 + 1, sx * fd, sy * fd, sz * fd), st.frac);
                            }
                            if constexpr (kGate) {                      // paper/tex/conetracing.tex:36-39
                                float al;
                                if constexpr (kTex) {
                                    al = tex3DLod<float>(ts.volA, sx, sy, sz, st.lod);
                                } else {
                                    const float fd = (float)D;
                                    al = sample_level(a, a.bitsA, a.chainA, st.level0, sx * fd, sy * fd, sz * fd);
                                    if (st.frac != 0.0f) al = lerpf(al, sample_level(a, a.bitsA, a.chainA, st.level0 + 1, sx * fd, sy * fd, sz * fd), st.frac);
                                }
                                if (!(al > 0.0f)) s = 0.0f;
                                if (kStats) nFetch += st.frac != 0.0f ? 2 : 1;
                            }
                            indirect = fmaf(s, st.weight, indirect);
                            if (kStats) nFetch += st.frac != 0.0f ? 2 : 1;
                        }
                    }
                };
                for (int g = 0; g < tp.nGroups; g++) coneGroup(g);
                // baked steps: the mip-linear lookup of this step, pre-blended on a lattice that holds the texel centres of
                // both levels (exact), as slice pairs: one bilinear pass returns the node planes floor(z) and floor(z)+1
                if constexpr (kTex && !kGate) {
                    auto bakedStep = [&](const int b) {
                        const BakedStep &bs = tp.baked[b];
                        const float sx = fmaf(bs.height, ex, nx), sy = fmaf(bs.height, ey, ny);
                        const float zl = __saturatef(fmaf(bs.height, ez, nz)) * bs.zScale;      // [0, n-1)
                        const float m = __fadd_rd(zl, 12582912.0f);                             // floor in the low mantissa bits
                        const float az = __fadd_rn(zl, -__fadd_rn(m, -12582912.0f));
                        const int layer = __float_as_int(m) & 0xFF;
                        float4 t;
                        asm("tex.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                            : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                            : "l"(bs.tex), "r"(layer), "f"(fmaf(sx, bs.A, bs.B)), "f"(fmaf(sy, bs.A, bs.B)));
                        if (shade) indirect = fmaf(fmaf(az, t.y, t.x), bs.weight, indirect);
                        if (kStats && shade) nBakedF++;
                    };
                    for (int b = 0; b < tp.nBaked; b++) bakedStep(b);
                }
                if (kStats && shade) { nCone += tp.nSteps; if (a.code && tp.nGroups > 0) nCode++; }
                if (tp.p.doNoiseSample) { col[0] *= indirect; col[1] *= indirect; col[2] *= indirect; }
                else col[0] = col[1] = col[2] = col[3] = indirect;
            }
        }

        if (shade && fb8) {                                             // SRC_ALPHA / ONE_MINUS_SRC_ALPHA onto the 8-bit target
            const float sa = saturatef(col[3]);
#pragma unroll
            for (int k = 0; k < 4; k++) dst8[k] = q8(saturatef(col[k]) * sa + dst8[k] * (1.0f - sa));
            nFrag++;
        } else if (shade) {                                             // blend, front to back
            const float sa = saturatef(col[3]);
            const float w = T * sa;
            C[0] = fmaf(w, saturatef(col[0]), C[0]);
            C[1] = fmaf(w, saturatef(col[1]), C[1]);
            C[2] = fmaf(w, saturatef(col[2]), C[2]);
            C[3] = fmaf(w, sa, C[3]);
            T *= (1.0f - sa);
            nFrag++;
        }
    }

    if (valid) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = fb8 ? dst8[k] : fmaf(T, bg[k], C[k]);
        const size_t p = (size_t)py * cam.W + px;
        if (a.format == CRN_IMAGE_RGBA32F) {
            reinterpret_cast<float4 *>(a.image)[p] = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            uchar4 q;
            q.x = (unsigned char)floorf(saturatef(o[0]) * 255.0f + 0.5f);
            q.y = (unsigned char)floorf(saturatef(o[1]) * 255.0f + 0.5f);
            q.z = (unsigned char)floorf(saturatef(o[2]) * 255.0f + 0.5f);
            q.w = (unsigned char)floorf(saturatef(o[3]) * 255.0f + 0.5f);
            reinterpret_cast<uchar4 *>(a.image)[p] = q;
        }
    }

    if (kStats) {
        for (int s = 16; s > 0; s >>= 1) {
            nFrag += __shfl_down_sync(0xFFFFFFFFu, nFrag, s);
            nCone += __shfl_down_sync(0xFFFFFFFFu, nCone, s);
            nNoise += __shfl_down_sync(0xFFFFFFFFu, nNoise, s);
            nSkip += __shfl_down_sync(0xFFFFFFFFu, nSkip, s);
            nFetch += __shfl_down_sync(0xFFFFFFFFu, nFetch, s);
            nBakedF += __shfl_down_sync(0xFFFFFFFFu, nBakedF, s);
            nLat += __shfl_down_sync(0xFFFFFFFFu, nLat, s);
            nCode += __shfl_down_sync(0xFFFFFFFFu, nCode, s);
        }
        if (lane == 0) {
            if (nFrag) atomicAdd(&a.stats[0], nFrag);
            if (nCone) atomicAdd(&a.stats[1], nCone);
            if (nNoise) atomicAdd(&a.stats[2], nNoise);
            if (nSkip) atomicAdd(&a.stats[3], nSkip);
            if (nFetch) atomicAdd(&a.stats[4], nFetch);
            if (nBakedF) atomicAdd(&a.stats[5], nBakedF);
            if (nLat) atomicAdd(&a.stats[6], nLat);
            if (nCode) atomicAdd(&a.stats[7], nCode);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// The fast variant: the reference's own configuration (src/Shaders/ConeTraceShader.hpp:15-36 defaults in kind, any
// values) — texture sampler, noise on with 4 octaves on a 32^3 texture, cone trace on, showQuad off, perspective
// camera, R8/R32F volume — with everything that does not depend on the fragment folded into constants, the octave /
// group / baked-step loops unrolled, and the per-fragment state arranged so that nothing is recomputed inside the
// noise march (the generic kernel above re-materialises a dozen values per step under its 64-register cap):
//   * colour is grey in every mode of conetrace_frag.glsl, so ONE colour accumulator + alpha;
//   * the noise march runs along the view ray, a kernel constant perpendicular to the billboard plane:
//     texture coordinate = QA + s1 * ray, |unitTex|^2 = |o|^2 / r^2 + s2^2 with two running scalars;
//   * background (clear colour + sun disc) is evaluated after the list walk, not carried through it;
//   * billboards flagged by the prep kernel (k_prep_sort.cu: the whole march stays inside the lattice window) take TWO
//     lookups per march step — octave 0 from the noise texture's RGBA16 copy, octaves 1..3 from the combined-octave lattice
//     (k_noiselat.cu) — with all six lookup coordinates as running sums; the others one lookup per octave;
//   * every slice-pair texel holds (plane, step to the next plane), so a z blend is one FMA per channel, and the noise
//     value is carried at half scale (the textures store v / 2), which turns saturate(ng * 0.5 + 0.5) into one FADD.SAT;
//   * the need code comes from one point-sampled fetch of a 3D texture of skip bits.
// Same fragments and the same order of composition as the generic kernel; the lookups are the reference's own or exact
// re-tabulations of them.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// bilinear fetch of texel pair planes from a layered texture (explicit level 0: no LOD operand in SASS)
__device__ __forceinline__ float4 tex_layer4(unsigned long long tex, int layer, float u, float v) {
    float4 t;
    asm("tex.level.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}], 0f00000000;"
        : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(tex), "r"(layer), "f"(u), "f"(v));
    return t;
}

// NB: number of baked cone steps (compile-time: no predication of the unused slots).
// kSeg: small frames (1280x720: a few thousand tiles) do not fill 148 SMs and are bound by the ONE warp that walks the
// longest list; and on mid-sized frames (C3) shorter-lived CTAs let the next frame's set-up kernels and the read-back in
// sooner.  There the list of a tile is cut into tp.segCount contiguous depth segments, each composited on its own
// by a separate CTA into (C_s, T_s); the last CTA to arrive merges them front to back: C = C_0 + T_0 C_1 + T_0 T_1 C_2 ...
// Same fragments, same order; the sum is re-associated (differences of a few ulp), and the early ray termination only
// sees its own segment's transmittance (still bounded by the cutoff).
template <int NB, bool kSeg>
__global__ void __launch_bounds__(kTraceThreads, CRN_TRACE_MINB) trace_fast_kernel(const __grid_constant__ TraceArgs a, const __grid_constant__ ViewParams cam,
                                                    const __grid_constant__ TraceParams tp, const __grid_constant__ TexSet ts) {
    constexpr int kWarps = kTraceThreads / 32, kSplit = 8 / kWarps;
    const int S = kSeg ? tp.segCount : 1;
    const int seg = kSeg ? (int)(blockIdx.x % S) : 0;
    const int cta = kSeg ? (int)(blockIdx.x / S) : (int)blockIdx.x;           // (tile, quarter)
    const int tile = (int)a.order[cta / kSplit];
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) + (cta % kSplit) * kWarps;
    const int tx = tile % a.tilesX, ty = tile / a.tilesX;
    // 2x2-pixel quads: the texture unit works on groups of four consecutive lanes, and a disc edge leaves fewer
    // partially filled 2x2 blocks than 4x1 strips (measured at C3: trace 3.35 -> 3.30 ms)
    const int lx = (lane & 1) | ((lane >> 1) & 6), ly = ((lane >> 1) & 1) | ((lane >> 3) & 2);
    // (the pixel's coordinates are recomputed after the list walk: only xs, ys and `valid` stay live through it)
    const int pxOf = (warp & 1) * 8 + lx, pyOf = (warp >> 1) * 4 + ly;
    bool valid;
    float xs, ys;                                                       // view-space x, y of the pixel ray at unit depth
    {
        const int px = tx * kTile + pxOf, py = ty * kTile + pyOf;
        if (tp.ilvCount > 1 && (ty % tp.ilvCount) != tp.ilvIndex) return;
        valid = px < cam.W && py < cam.H && py >= tp.row0 && py < tp.row1;
        if (!kSeg && __all_sync(0xFFFFFFFFu, !valid)) return;       // (segmented: every thread stays for the barrier below)
        xs = (((float)px + 0.5f) / (float)cam.W * 2.0f - 1.0f) * tp.f.invP0;
        ys = (((float)py + 0.5f) / (float)cam.H * 2.0f - 1.0f) * tp.f.invP5;
    }

    float Cg = 0.0f, Ca = 0.0f, T = 1.0f;
    const uint32_t cnt = a.tileCnt[tile];
    const uint32_t *list = a.tileList + (cnt ? a.tileOff[tile] : 0);
    const float cutoff = tp.p.transmittanceCutoff;
    const float rx = cam.nrm[0], ry = cam.nrm[1], rz = cam.nrm[2];     // viewRay (res/conetrace_frag.glsl:138)
    // this CTA's share of the list: short lists are not cut (segment 0 takes them whole, the other CTAs leave)
    // ... and neither are very long ones: hundreds of layers of large billboards end in the early ray termination, which a
    // later segment cannot see (reference radii at C3: ~4700 entries per tile, 7.8 ms uncut against 12.3 ms cut in two)
    const int nSeg = (kSeg && cnt >= (uint32_t)(S * tp.segMin) && cnt <= (uint32_t)tp.segMax) ? S : 1;
    if (kSeg && seg >= nSeg) return;
    const uint32_t eBegin = kSeg ? (uint32_t)((uint64_t)cnt * seg / nSeg) : 0u;
    const uint32_t eEnd = kSeg ? (uint32_t)((uint64_t)cnt * (seg + 1) / nSeg) : cnt;

    for (const uint32_t *lp = list + eBegin, *const le = list + eEnd; lp < le; lp++) {
        const bool live = valid && T > cutoff;
        if (__all_sync(0xFFFFFFFFu, !live)) break;                      // early ray termination, whole patch
        const uint32_t k = __ldg(lp);
        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(&a.recs[k].cx));
        const float4 r1 = __ldg(reinterpret_cast<const float4 *>(&a.recs[k].xv));
        const float radius = r0.w;
        const float u = fmaf(xs, -r1.z, -r1.x), v = fmaf(ys, -r1.z, -r1.y);
        const float d2 = fmaf(u, u, v * v), r2 = radius * radius;
        const float h2 = r2 - d2;
        // both discards (4 h^2 < 0.01 and h^2 < 1e-4 r^2) and the quad test (implied by h^2 > 0) in one compare
        const bool shade = live && h2 >= fmaxf(1.0e-4f * r2, 0.0025f);
        if (!__any_sync(0xFFFFFFFFu, shade)) continue;

        const float h = sqrtf(fmaxf(h2, 0.0f));
        const float invr = 1.0f / radius;
        const float ox = fmaf(u, cam.right[0], v * cam.up[0]);
        const float oy = fmaf(u, cam.right[1], v * cam.up[1]);
        const float oz = fmaf(u, cam.right[2], v * cam.up[2]);
        // point of the billboard plane under this pixel
        const float qx = r0.x + ox, qy = r0.y + oy, qz = r0.z + oz;

        // ---- noise march (res/conetrace_frag.glsl:137-174)
        const float invAdjust = tp.f.invAdjust;
        const float len = 2.0f * h * invAdjust;
        const float iSteps = fminf(len * tp.f.invStep, tp.f.span) + tp.f.minSteps;
        const float inv = 1.0f / (iSteps - 1.0f);
        const float sStep = len * inv;
        const int nIter = shade ? (int)ceilf(iSteps) : 0;
        const int nMax = __reduce_max_sync(0xFFFFFFFFu, nIter);
        const float qax = qx * invAdjust, qay = qy * invAdjust, qaz = qz * invAdjust;
        // (the noise textures return HALF the noise value, see TexSet::noiseD: the march carries ng / 2 and na / 2)
        const float uu0 = d2 * invr * invr, vy0 = 0.5f * oy * invr;
        float s1 = -h * invAdjust;                                      // texture-space distance from the plane point
        float s2 = -h * invr;                                           // unit-sphere distance from the plane (sic: advanced by the texture-space step)
        float opacity = 0.0f, light = 0.0f;
        // the combined-octave lattice (k_noiselat.cu) covers this billboard's whole march: decided per billboard by the
        // camera-side prep kernel (k_prep_sort.cu), warp-uniform
        const bool inLat = (__float_as_int(r1.w) & kRecInLattice) != 0;
        if (inLat) {
            // every lookup coordinate is affine in the march distance: six running sums, no constants inside the loop
            const float K = tp.lat.K;
            const float c0x = fmaf(s1, rx, qax), c0y = fmaf(s1, ry, qay), c0z = fmaf(s1, rz, qaz);
            float u0 = c0x + tp.octBias[0], v0 = c0y + tp.octBias[0], z0 = fmaf(c0z, tp.octFreqZ[0], tp.octBiasZ[0]);
            float u1 = fmaf(c0x, K, tp.lat.B[0]), v1 = fmaf(c0y, K, tp.lat.B[1]), z1 = fmaf(c0z, K, tp.lat.B[2]);
            const float du0 = sStep * rx, dv0 = sStep * ry, dz0 = sStep * (rz * tp.octFreqZ[0]);
            const float du1 = du0 * K, dv1 = dv0 * K, dz1 = sStep * (rz * K);
            const float ryh = 0.5f * ry;
            for (int i = 0; i < nMax; i++) {
                if (i < nIter) {
                    // octave 0 (it alone carries the wind offset): plane floor(z) of the noise texture and the step to the next
                    const float m = __fadd_rd(z0, 12582912.0f);
                    const float az = __fadd_rn(z0, -__fadd_rn(m, -12582912.0f));
                    const float4 t = tex_layer4(ts.noiseD, __float_as_int(m) & 31, u0, v0);
                    // octaves 1..3, pre-summed on the finest octave's texel lattice: node plane floor(z) and the step to the next
                    const float ml = __fadd_rd(z1, 12582912.0f);
                    const float al = __fadd_rn(z1, -__fadd_rn(ml, -12582912.0f));
                    const float4 q = tex_layer4(tp.lat.tex, __float_as_int(ml) & 0x7FF, u1, v1);
                    float ng = fmaf(tp.lat.scale, fmaf(al, q.z, q.x), fmaf(az, t.z, t.x));
                    const float na = fmaf(tp.lat.scale, fmaf(al, q.w, q.y), fmaf(az, t.w, t.y));
                    const float uu = fmaf(s2, s2, uu0);
                    ng = fmaf(fmaf(s2, ryh, vy0), rsqrtf(uu), ng);      // (noiseCell.y + normalize(unitTex).y) / 2
                    opacity = fmaf(fabsf(na), 1.0f - uu, opacity);
                    light += __saturatef(ng + 0.5f);
                    u0 += du0; v0 += dv0; z0 += dz0; u1 += du1; v1 += dv1; z1 += dz1;
                    s2 += sStep;
                }
            }
        } else
        for (int i = 0; i < nMax; i++) {
            if (i < nIter) {
                const float cx = fmaf(s1, rx, qax), cy = fmaf(s1, ry, qay), cz = fmaf(s1, rz, qaz);
                float ng = 0.0f, na = 0.0f;
#pragma unroll
                for (int o = 0; o < 4; o++) {
                    // layer = floor(z texel coordinate) on the FADD pipe: adding 1.5*2^23 rounding DOWN leaves the floor in
                    // the low mantissa bits (two's complement, so the & also wraps negative layers); octave 3's offset is 0
                    // by decree (octaveOffsets[3]): literal bias.  Octave 0 has frequency and persistence 1 (noise3D starts there).
                    const float zc = o == 3 ? fmaf(cz, tp.octFreqZ[o], -0.5f) : fmaf(cz, tp.octFreqZ[o], tp.octBiasZ[o]);
                    const float m = __fadd_rd(zc, 12582912.0f);
                    const float az = __fadd_rn(zc, -__fadd_rn(m, -12582912.0f));
                    const int layer = __float_as_int(m) & 31;
                    const float tu = o == 0 ? cx + tp.octBias[0] : o == 3 ? cx * tp.octFreq[o] : fmaf(cx, tp.octFreq[o], tp.octBias[o]);
                    const float tv = o == 0 ? cy + tp.octBias[0] : o == 3 ? cy * tp.octFreq[o] : fmaf(cy, tp.octFreq[o], tp.octBias[o]);
                    const float4 t = tex_layer4(ts.noiseD, layer, tu, tv);
                    const float sg = fmaf(az, t.z, t.x), sa = fmaf(az, t.w, t.y);
                    if (o == 0) { ng = sg; na = sa; }
                    else { ng = fmaf(tp.octPers[o], sg, ng); na = fmaf(tp.octPers[o], sa, na); }
                }
                const float uu = fmaf(s2, s2, uu0);                     // |unitTex|^2: the plane offset is perpendicular to the ray
                ng = fmaf(fmaf(s2, 0.5f * ry, vy0), rsqrtf(uu), ng);    // (noiseCell.y + normalize(unitTex).y) / 2
                opacity = fmaf(fabsf(na), 1.0f - uu, opacity);
                light += __saturatef(ng + 0.5f);
                s1 += sStep; s2 += sStep;
            }
        }
        const float grey = fmaf(tp.p.noiseColorScale * light, inv, tp.p.minNoiseColor);
        const float alpha = __saturatef(opacity * (2.0f * tp.p.noiseOpacity) * inv) * (1.0f - sqrtf(d2) * invr);

        // ---- cone trace towards the sun (res/conetrace_frag.glsl:176-200, traceCone :64-79)
        const float wx3 = fmaf(rx, h, qx), wy3 = fmaf(ry, h, qy), wz3 = fmaf(rz, h, qz);   // camera-facing sphere surface
        const float nx = fmaf(wx3, tp.f.nScale[0], tp.f.nBias[0]);
        const float ny = fmaf(wy3, tp.f.nScale[1], tp.f.nBias[1]);
        const float nz = fmaf(wz3, tp.f.nScale[2], tp.f.nBias[2]);
        float ex = tp.lightPos[0] - wx3, ey = tp.lightPos[1] - wy3, ez = tp.lightPos[2] - wz3;
        const float il = rsqrtf(fmaf(ex, ex, fmaf(ey, ey, ez * ez))) * tp.f.invDim;        // normalize(dir) / voxelDim
        ex *= il; ey *= il; ez *= il;
        float indirect = 0.0f;
        if (tp.nGroups > 0) {
            // bit g of the start cell's SKIP code: no start position in this cell can reach a non-zero texel with group g.  One
            // point-sampled fetch of the code grid (k_conebake.cu) at the start position; outside the volume the border
            // returns 0: everything is fetched.  Groups beyond kCodeGroups have no bit.
            uint32_t code = tp.nGroups >= 32 ? 0xFFFFFFFFu : (1u << tp.nGroups) - 1u;
            if (a.code) {
                uint32_t skip, u1_, u2_, u3_;
                asm("tex.level.3d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}], 0f00000000;"
                    : "=r"(skip), "=r"(u1_), "=r"(u2_), "=r"(u3_) : "l"(ts.code), "f"(nx), "f"(ny), "f"(nz));
                code &= ~skip;
            }
            code = shade ? code : 0u;
            if (__any_sync(0xFFFFFFFFu, code != 0)) {                   // most patches need no fine step at all
                for (int g = 0; g < tp.nGroups; g++) {
                    const ConeGroup &gr = tp.groups[g];
                    const bool need = (code >> (g < 31 ? g : 31)) & 1u;
                    if (!__any_sync(0xFFFFFFFFu, need)) continue;
                    for (int i = gr.first; i < gr.first + gr.count; i++) {
                        const ConeStep &st = tp.steps[i];
                        if (need) {
                            const float s = tex3DLod<float>(ts.vol, fmaf(st.height, ex, nx), fmaf(st.height, ey, ny), fmaf(st.height, ez, nz), st.lod);
                            indirect = fmaf(s, st.weight, indirect);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const BakedStep &bs = tp.baked[b];
            const float zl = fma_sat(bs.height, ez, nz) * bs.zScale;                        // [0, n-1)
            const float m = __fadd_rd(zl, 12582912.0f);                                     // floor in the low mantissa bits
            const float az = __fadd_rn(zl, -__fadd_rn(m, -12582912.0f));
            const float4 t = tex_layer4(bs.tex, __float_as_int(m) & 0xFF, fmaf(bs.hA, ex, fmaf(nx, bs.A, bs.B)), fmaf(bs.hA, ey, fmaf(ny, bs.A, bs.B)));
            indirect = fmaf(fmaf(az, t.y, t.x), bs.weight, indirect);
        }

        if (shade) {                                                    // blend, front to back
            const float w = T * alpha;                                  // alpha is already in [0,1]
            Cg = fmaf(w, __saturatef(grey * indirect), Cg);
            Ca = fmaf(w, alpha, Ca);
            T *= (1.0f - alpha);
        }
    }

    if constexpr (kSeg) {
        if (nSeg > 1) {
            __shared__ uint32_t sLast;
            float4 *part = a.segPartial + ((size_t)tile * S) * 256 + warp * 32 + lane;
            __stcg(part + (size_t)seg * 256, make_float4(Cg, Ca, T, 0.0f));
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) sLast = atomicAdd(&a.segArrived[cta], 1u) == (uint32_t)(nSeg - 1) ? 1u : 0u;
            __syncthreads();
            if (!sLast) return;                                         // the last segment to finish merges all of them
            __threadfence();
            if (threadIdx.x == 0) a.segArrived[cta] = 0;                // re-arm for the next frame
            Cg = 0.0f; Ca = 0.0f; T = 1.0f;
            for (int s2 = 0; s2 < nSeg; s2++) {                         // front to back
                const float4 p = __ldcg(part + (size_t)s2 * 256);
                Cg = fmaf(T, p.x, Cg); Ca = fmaf(T, p.y, Ca); T *= p.z;
            }
        }
    }

    if (valid) {
        const int px = (tile % a.tilesX) * kTile + pxOf, py = (tile / a.tilesX) * kTile + pyOf;
        const float ndcx = ((float)px + 0.5f) / (float)cam.W * 2.0f - 1.0f;
        const float ndcy = ((float)py + 0.5f) / (float)cam.H * 2.0f - 1.0f;
        // background = clear colour, then the sun pass blended over it
        float bg[4] = {tp.bg[0], tp.bg[1], tp.bg[2], tp.bg[3]};
        if (tp.p.drawSun) {
            float sc[4];
            if (sun_fragment(tp, cam, ndcx, ndcy, px, py, sc)) {
                const float sa = saturatef(sc[3]);
#pragma unroll
                for (int k = 0; k < 4; k++) bg[k] = saturatef(sc[k]) * sa + bg[k] * (1.0f - sa);
            }
        }
        const float o0 = fmaf(T, bg[0], Cg), o1 = fmaf(T, bg[1], Cg), o2 = fmaf(T, bg[2], Cg), o3 = fmaf(T, bg[3], Ca);
        const size_t p = (size_t)py * cam.W + px;
        if (a.format == CRN_IMAGE_RGBA32F) {
            reinterpret_cast<float4 *>(a.image)[p] = make_float4(o0, o1, o2, o3);
        } else {
            uchar4 q;
            q.x = (unsigned char)floorf(saturatef(o0) * 255.0f + 0.5f);
            q.y = (unsigned char)floorf(saturatef(o1) * 255.0f + 0.5f);
            q.z = (unsigned char)floorf(saturatef(o2) * 255.0f + 0.5f);
            q.w = (unsigned char)floorf(saturatef(o3) * 255.0f + 0.5f);
            reinterpret_cast<uchar4 *>(a.image)[p] = q;
        }
    }
}

} // namespace

int launch_trace(cudaStream_t st, const ViewParams &cam, const VolumeParams &vol, const TraceParams &tp,
                 const BoardRec *recs, const Bins &b, const uint32_t *bits, const uint8_t *chain,
                 const uint32_t *bitsA, const uint8_t *chainA, const int8_t *noise, const TexSet *ts, const uint8_t *needCode, const uint32_t *tileOrder, void *image,
                 int format, unsigned long long *stats, float4 *segPartial, uint32_t *segArrived, int ownedTiles) {
    TraceArgs a;
    a.vol = vol;
    a.recs = recs;
    a.tileOff = b.tileOff; a.tileCnt = b.tileCnt; a.tileList = b.tileList;
    a.tilesX = b.tilesX; a.tilesY = b.tilesY;
    a.bits = bits; a.chain = chain; a.bitsA = bitsA; a.chainA = chainA;
    a.noise = reinterpret_cast<const float2 *>(noise);
    a.image = image; a.format = format; a.stats = stats;
    a.code = (needCode && tp.p.skipEmptySpace && tp.codeDim > 0) ? needCode : nullptr;
    a.order = tileOrder;
    a.segPartial = segPartial; a.segArrived = segArrived;
    a.invRange[0] = 1.0f / (vol.xB[1] - vol.xB[0]);
    a.invRange[1] = 1.0f / (vol.yB[1] - vol.yB[0]);
    a.invRange[2] = 1.0f / (vol.zB[1] - vol.zB[0]);
    a.invDim = 1.0f / (float)vol.dim;
    TexSet none{};
    const bool useTex = ts && ts->enabled && tp.p.sampler == CRN_SAMPLER_TEXTURE;
    const int grid = ownedTiles * (256 / kTraceThreads);          // the tile order lists owned tiles only
    const bool gate = bitsA != nullptr;
    // the reference's configuration: see trace_fast_kernel
    // The fast variant takes floor() of the noise z coordinate by adding 1.5 * 2^23, valid for |z| < 2^22 texels.  z is
    // (world / adjustSize * freq + offset) * 32; the host does not know the billboard positions, so it bounds |world| by
    // 4x the volume's extent from the origin (+64): billboards further out than that need the generic variant
    // (CRN_NO_FAST=1), which uses floorf.
    float extent = 0.0f;
    for (int k = 0; k < 2; k++) extent = fmaxf(extent, fmaxf(fabsf(vol.xB[k]), fmaxf(fabsf(vol.yB[k]), fabsf(vol.zB[k]))));
    float zmax = 0.0f;
    for (int o = 0; o < 4; o++) zmax = fmaxf(zmax, (4.0f * extent + 64.0f) / fmaxf(fabsf(tp.p.adjustSize), 1e-20f) * fabsf(tp.octFreqZ[o]) + fabsf(tp.octBiasZ[o]));
    const bool zOk = zmax < 2097152.0f;
    const bool fast = zOk && useTex && !gate && !tp.stats && tp.active && !tp.p.quantizeFramebuffer && tp.p.numOctaves == 4 && tp.noiseDim == 32 && tp.p.doNoiseSample &&
                      tp.p.doConeTrace && !tp.p.showQuad && !cam.ortho && tp.nBaked <= kFastBaked && vol.texelBytes <= 4 && !getenv("CRN_NO_FAST");
    if (gate) {                                                   // opt-in paper variant: stats variant only when asked
        if (useTex && tp.stats) trace_kernel<true, true, true><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, *ts);
        else if (useTex) trace_kernel<true, false, true><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, *ts);
        else if (tp.stats) trace_kernel<false, true, true><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, none);
        else trace_kernel<false, false, true><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, none);
    }
    else if (fast) {
        const bool segmented = tp.segCount > 1 && segPartial && segArrived;
        const int gridF = segmented ? grid * tp.segCount : grid;
        switch (tp.nBaked) {
#define CRN_FAST(NB) case NB: if (segmented) trace_fast_kernel<NB, true><<<gridF, kTraceThreads, 0, st>>>(a, cam, tp, *ts); \
                              else trace_fast_kernel<NB, false><<<gridF, kTraceThreads, 0, st>>>(a, cam, tp, *ts); break;
            CRN_FAST(0) CRN_FAST(1) CRN_FAST(2) CRN_FAST(3) CRN_FAST(4) CRN_FAST(5) CRN_FAST(6) CRN_FAST(7) CRN_FAST(8)
            CRN_FAST(9) CRN_FAST(10) CRN_FAST(11) CRN_FAST(12) CRN_FAST(13) CRN_FAST(14) CRN_FAST(15) CRN_FAST(16)
#undef CRN_FAST
        }
    }
    else if (useTex && tp.stats) trace_kernel<true, true, false><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, *ts);
    else if (useTex) trace_kernel<true, false, false><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, *ts);
    else if (tp.stats) trace_kernel<false, true, false><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, none);
    else trace_kernel<false, false, false><<<grid, kTraceThreads, 0, st>>>(a, cam, tp, none);
    return 1;
}

} // namespace crn
