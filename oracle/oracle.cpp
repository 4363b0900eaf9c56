/*
 * oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A behavioural restatement, in plain C++17 (+OpenMP over image rows), of the per-frame
 * volumetric pipeline of jaafersheriff/Cloud-Renderer:
 *     voxelize pass 1 + pass 2  ->  3D mip chain  ->  billboard cone trace.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object, and only as the checker / reported baseline.
 * Nothing under cloud-renderer_b200/ includes, links or dlopens it.
 *
 * PARITY STATUS.  The reference ships no tests, golden vectors or fixtures (SURVEY.md §4,
 * §8c) and its GL 4.4 + GLFW + GLM build cannot run in this image, so the reference does
 * not pin this oracle by itself.  What pins it instead:
 *   (1) oracle/ref_glsl/ compiles the reference's OWN shader sources
 *       (res/first_voxelize.glsl, res/second_voxelize.glsl, res/conetrace_frag.glsl,
 *       res/sun_frag.glsl, res/billboard_vert*.glsl), read in place from /root/reference,
 *       as C++ through a small GLSL-vocabulary shim into oracle/_ref/, and
 *       tests/test_oracle_vs_ref_glsl_cpu.py (live) and tests/test_golden_cpu.py (committed
 *       vectors, tests/golden/) compare this file's per-fragment arithmetic against them;
 *   (2) closed-form anchors (tests/test_oracle_closed_form_cpu.py).
 * The fixed-function stages a GL driver supplies (rasterisation, depth test, texture
 * filtering, mip generation, blending) exist nowhere in /root/reference and are restated
 * here from the OpenGL 4.4 core specification; driver-defined choices are decreed once,
 * below, and listed in DESIGN.md ("Decrees").
 *
 * Every function cites the reference file:line it follows.  All arithmetic is float32 in
 * the shader's written operation order; build with -ffp-contract=off (oracle/Makefile).
 */
#include "../include/cloud_renderer_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------
// GLM 0.9.8.5 subset (un-vendored dependency, Clouds.vcxproj:118).  Restated from the
// published formulas: right-handed, -1..1 clip depth, no GLM_FORCE_* defines
// (Clouds.vcxproj:129-181).  dot() sums left to right like glm::detail::compute_dot.
// ---------------------------------------------------------------------------------
struct vec3 { float x, y, z; };
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(vec3 v) { return sqrtf(dot(v, v)); }
inline float distance(vec3 a, vec3 b) { return length(b - a); }
inline vec3 normalize(vec3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
inline vec3 cross(vec3 a, vec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline vec3 v3(const float *p) { return {p[0], p[1], p[2]}; }
inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// column-major: m[c*4+r]
inline void mat_identity(float *m) { std::memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.0f; }

void glm_lookAt(vec3 eye, vec3 center, vec3 up, float *m) {
    vec3 f = normalize(center - eye);
    vec3 s = normalize(cross(f, up));
    vec3 u = cross(s, f);
    mat_identity(m);
    m[0] = s.x; m[4] = s.y; m[8] = s.z;
    m[1] = u.x; m[5] = u.y; m[9] = u.z;
    m[2] = -f.x; m[6] = -f.y; m[10] = -f.z;
    m[12] = -dot(s, eye);
    m[13] = -dot(u, eye);
    m[14] = dot(f, eye);
}

void glm_ortho(float l, float r, float b, float t, float n, float f, float *m) {
    mat_identity(m);
    m[0] = 2.0f / (r - l);
    m[5] = 2.0f / (t - b);
    m[10] = -2.0f / (f - n);
    m[12] = -(r + l) / (r - l);
    m[13] = -(t + b) / (t - b);
    m[14] = -(f + n) / (f - n);
}

void glm_perspective(float fovy, float aspect, float n, float f, float *m) {
    float tanHalf = tanf(fovy / 2.0f);
    std::memset(m, 0, 64);
    m[0] = 1.0f / (aspect * tanHalf);
    m[5] = 1.0f / tanHalf;
    m[11] = -1.0f;
    m[10] = -(f + n) / (f - n);
    m[14] = -(2.0f * f * n) / (f - n);
}

// mat4 * (v,1), GLSL/GLM column-vector convention, terms summed left to right
inline void mat_mul_point(const float *m, vec3 v, float out[4]) {
    for (int r = 0; r < 4; r++)
        out[r] = ((m[0 + r] * v.x + m[4 + r] * v.y) + m[8 + r] * v.z) + m[12 + r];
}

// ---------------------------------------------------------------------------------
// Uniform blocks exactly as the pass drivers upload them.
// ---------------------------------------------------------------------------------
struct VolumeUniforms {          // VoxelizeShader::bindVolume, src/Shaders/VoxelizeShader.cpp:120-129
    float xB[2], yB[2], zB[2];   //   ConeTraceShader::bindVolume, src/Shaders/ConeTraceShader.cpp:84-93
    int voxelDim;
    float stepSize;              // min(voxelSize.xyz), voxelSize = range/(float)dimension (src/CloudVolume.cpp:88-92)
};

VolumeUniforms volume_uniforms(const crn_volume_desc &v) {
    VolumeUniforms u;
    for (int k = 0; k < 2; k++) {
        u.xB[k] = v.position[0] + v.xBounds[k];
        u.yB[k] = v.position[1] + v.yBounds[k];
        u.zB[k] = v.position[2] + v.zBounds[k];
    }
    u.voxelDim = v.dimension;
    float rx = v.xBounds[1] - v.xBounds[0], ry = v.yBounds[1] - v.yBounds[0], rz = v.zBounds[1] - v.zBounds[0];
    float d = (float)v.dimension;
    u.stepSize = fminf(rx / d, fminf(ry / d, rz / d));
    return u;
}

// calculateVoxelLerp, res/second_voxelize.glsl:18-28 == res/conetrace_frag.glsl:48-58
inline vec3 voxel_lerp(const VolumeUniforms &u, vec3 pos) {
    float rangeX = u.xB[1] - u.xB[0];
    float rangeY = u.yB[1] - u.yB[0];
    float rangeZ = u.zB[1] - u.zB[0];
    float fd = (float)u.voxelDim;
    return {fd * ((pos.x - u.xB[0]) / rangeX), fd * ((pos.y - u.yB[0]) / rangeY), fd * ((pos.z - u.zB[0]) / rangeZ)};
}

// The billboard vertex stage for a view-facing instanced quad
// (res/billboard_vert_instanced.glsl:20-37 with Vi = transpose(V with zero translation),
// src/Shaders/VoxelizeShader.cpp:50-53 / src/Shaders/ConeTraceShader.cpp:66-68).
// M*Vi*(vx,vy,0,1) = center + scale*(vx*right + vy*up); all four corners share one
// view-space depth, so attribute interpolation is affine for both the ortho light camera
// and the perspective user camera, and fragPos at a pixel centre is the un-projection of
// that centre onto the quad's plane.  fragNor = mat3(transpose(inverse(M*Vi)))*(0,0,1)
// = back/scale; the fragment stages only ever use normalize(fragNor) = back.
struct ViewBasis {
    vec3 right, up, back;   // rows 0..2 of V's rotation block
    vec3 nrm;               // normalize(back) == normalize(fragNor)
    const float *V, *P;
    bool ortho;
};

ViewBasis make_basis(const float *P, const float *V) {
    ViewBasis b;
    b.right = {V[0], V[4], V[8]};
    b.up = {V[1], V[5], V[9]};
    b.back = {V[2], V[6], V[10]};
    b.nrm = normalize(b.back);
    b.V = V; b.P = P;
    b.ortho = (P[15] == 1.0f);
    return b;
}

// Pixel-centre un-projection (fixed-function viewport + rasteriser, GL 4.4 §13.6/§14.6):
// window x = (ndc+1)/2*W, samples at i+0.5.  Returns view-space x,y of the centre on the
// plane view-z = zv.
inline float ndc_of(int i, int n) { return ((float)i + 0.5f) / (float)n * 2.0f - 1.0f; }
inline float view_x(const ViewBasis &b, float ndcx, float zv) {
    return b.ortho ? (ndcx - b.P[12]) / b.P[0] : (ndcx * (0.0f - zv)) / b.P[0];
}
inline float view_y(const ViewBasis &b, float ndcy, float zv) {
    return b.ortho ? (ndcy - b.P[13]) / b.P[5] : (ndcy * (0.0f - zv)) / b.P[5];
}

struct QuadSetup {
    vec3 center; float scale;
    float cv[4];            // V*(center,1)
    bool clipped;           // whole quad outside near/far
    int i0, i1, j0, j1;     // conservative pixel rect, inclusive, already clamped
};

// window-space rectangle of the quad, padded by one pixel so rounding can never lose a
// covered centre; coverage itself is decided per pixel by the fragment stage.
QuadSetup quad_setup(const ViewBasis &b, vec3 center, float scale, int W, int H) {
    QuadSetup q;
    q.center = center; q.scale = scale;
    mat_mul_point(b.V, center, q.cv);
    float zv = q.cv[2];
    // clip-space z and w of every vertex of this quad
    float zc = b.P[10] * zv + b.P[14] * 1.0f;
    float wc = b.P[11] * zv + b.P[15] * 1.0f;
    q.clipped = !(zc >= -wc && zc <= wc) || !(wc > 0.0f);
    if (q.clipped) { q.i0 = q.j0 = 0; q.i1 = q.j1 = -1; return q; }
    float x0 = (b.P[0] * (q.cv[0] - scale) + b.P[12] * 1.0f) / wc;
    float x1 = (b.P[0] * (q.cv[0] + scale) + b.P[12] * 1.0f) / wc;
    float y0 = (b.P[5] * (q.cv[1] - scale) + b.P[13] * 1.0f) / wc;
    float y1 = (b.P[5] * (q.cv[1] + scale) + b.P[13] * 1.0f) / wc;
    float fx0 = (x0 + 1.0f) * 0.5f * (float)W, fx1 = (x1 + 1.0f) * 0.5f * (float)W;
    float fy0 = (y0 + 1.0f) * 0.5f * (float)H, fy1 = (y1 + 1.0f) * 0.5f * (float)H;
    // clamp in float first so huge quads cannot overflow the int conversion
    fx0 = fminf(fmaxf(fx0, -2.0f), (float)W + 2.0f); fx1 = fminf(fmaxf(fx1, -2.0f), (float)W + 2.0f);
    fy0 = fminf(fmaxf(fy0, -2.0f), (float)H + 2.0f); fy1 = fminf(fmaxf(fy1, -2.0f), (float)H + 2.0f);
    q.i0 = std::max(0, (int)floorf(fx0) - 1); q.i1 = std::min(W - 1, (int)floorf(fx1) + 1);
    q.j0 = std::max(0, (int)floorf(fy0) - 1); q.j1 = std::min(H - 1, (int)floorf(fy1) + 1);
    return q;
}

// fragPos for pixel (i,j) of this quad + its in-plane offsets (u,v) from the centre
inline vec3 frag_pos(const ViewBasis &b, const QuadSetup &q, int i, int j, int W, int H, float *uo, float *vo) {
    float xv = view_x(b, ndc_of(i, W), q.cv[2]);
    float yv = view_y(b, ndc_of(j, H), q.cv[2]);
    float u = xv - q.cv[0], v = yv - q.cv[1];
    *uo = u; *vo = v;
    return {q.center.x + (u * b.right.x + v * b.up.x),
            q.center.y + (u * b.right.y + v * b.up.y),
            q.center.z + (u * b.right.z + v * b.up.z)};
}

inline float board_radius(const crn_volume_desc &vol, float s) {
    // CloudVolume::uploadBillboards, src/CloudVolume.cpp:153-161
    return vol.fluffiness != 1.0f ? s * vol.fluffiness : s;
}

// ---------------------------------------------------------------------------------
// Texture sampling restated from GL 4.4 §8.14 (linear filter, §8.14.3 mipmapping).
// ---------------------------------------------------------------------------------
struct Chain {                  // R8 immutable 3D texture with `levels` mips (src/CloudVolume.cpp:18-23);
    const uint8_t *data;        // or, for CRN_VOLUME_R32F, the same chain with float texels (fdata)
    const float *fdata;
    const uint8_t *alpha;       // CRN_VOLUME_RG8 (paper variant): the occupancy channel's own R8 chain, else NULL
    int dim, levels;
    size_t off[16]; int size[16];   // element offsets
};

Chain make_chain(const void *data, int dim, int levels, bool is_float = false) {
    Chain c; c.dim = dim; c.levels = levels; c.alpha = nullptr;
    c.data = is_float ? nullptr : (const uint8_t *)data;
    c.fdata = is_float ? (const float *)data : nullptr;
    size_t o = 0; int s = dim;
    for (int l = 0; l < levels; l++) { c.off[l] = o; c.size[l] = s; o += (size_t)s * s * s; s = std::max(1, s / 2); }
    return c;
}

// the chain a scene's cone trace reads; CRN_VOLUME_RG8 is planar: the rgb chain, then the alpha chain
Chain scene_chain(const crn_volume_desc &vol, const uint8_t *chain_bytes) {
    Chain c = make_chain(chain_bytes, vol.dimension, vol.levels, vol.format == CRN_VOLUME_R32F);
    if (vol.format == CRN_VOLUME_RG8) {
        size_t total = 0;
        for (int l = 0; l < c.levels; l++) total += (size_t)c.size[l] * c.size[l] * c.size[l];
        c.alpha = chain_bytes + total;
    }
    return c;
}

// LINEAR, CLAMP_TO_EDGE x3, UNORM8 decode c/255
float sample_level(const Chain &c, int l, vec3 uvw) {
    int n = c.size[l];
    const uint8_t *t = c.data ? c.data + c.off[l] : nullptr;
    const float *tf = c.fdata ? c.fdata + c.off[l] : nullptr;
    float fx = uvw.x * (float)n - 0.5f, fy = uvw.y * (float)n - 0.5f, fz = uvw.z * (float)n - 0.5f;
    float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    float ax = fx - flx, ay = fy - fly, az = fz - flz;
    // clamp in float: NaN/inf coordinates must not reach the int conversion
    auto cl = [n](float f) { f = fminf(fmaxf(f, -1.0f), (float)n); int i = (int)f; return std::min(std::max(i, 0), n - 1); };
    int x0 = cl(flx), x1 = cl(flx + 1.0f), y0 = cl(fly), y1 = cl(fly + 1.0f), z0 = cl(flz), z1 = cl(flz + 1.0f);
    auto T = [&](int x, int y, int z) {
        const size_t i = ((size_t)z * n + y) * n + x;
        return tf ? tf[i] : (float)t[i] / 255.0f;                 // R32F texel, or UNORM8 decode
    };
    float c00 = T(x0, y0, z0) * (1.0f - ax) + T(x1, y0, z0) * ax;
    float c10 = T(x0, y1, z0) * (1.0f - ax) + T(x1, y1, z0) * ax;
    float c01 = T(x0, y0, z1) * (1.0f - ax) + T(x1, y0, z1) * ax;
    float c11 = T(x0, y1, z1) * (1.0f - ax) + T(x1, y1, z1) * ax;
    float c0 = c00 * (1.0f - ay) + c10 * ay;
    float c1 = c01 * (1.0f - ay) + c11 * ay;
    return c0 * (1.0f - az) + c1 * az;
}

// textureLod with MIN = LINEAR_MIPMAP_LINEAR, MAG = LINEAR (src/CloudVolume.cpp:19-20)
float texture_lod(const Chain &c, vec3 uvw, float lod) {
    float q = (float)(c.levels - 1);
    if (!(lod > 0.0f)) return sample_level(c, 0, uvw);          // magnification (and NaN)
    if (lod >= q) return sample_level(c, c.levels - 1, uvw);
    float fl = floorf(lod);
    int d0 = (int)fl;
    float f = lod - fl;
    float s0 = sample_level(c, d0, uvw);
    if (f == 0.0f) return s0;
    float s1 = sample_level(c, d0 + 1, uvw);
    return s0 * (1.0f - f) + s1 * f;
}

struct Noise { const int8_t *rgba; int dim; };                 // RGBA8_SNORM, REPEAT, LINEAR, no mips
                                                               // (src/Shaders/ConeTraceShader.cpp:152-158)
inline float snorm(int8_t c) { return fmaxf((float)c / 127.0f, -1.0f); }

void sample_noise(const Noise &nz, vec3 uvw, float out[4]) {
    int n = nz.dim;
    float f[3] = {uvw.x, uvw.y, uvw.z};
    int i0[3], i1[3]; float a[3];
    for (int k = 0; k < 3; k++) {
        float s = f[k] - floorf(f[k]);                          // REPEAT: fractional part
        float t = s * (float)n - 0.5f;
        float fl = floorf(t);
        a[k] = t - fl;
        int i = (int)fl;                                        // in [-1, n-1]
        i0[k] = ((i % n) + n) % n;
        i1[k] = (i0[k] + 1) % n;
    }
    for (int ch = 0; ch < 4; ch++) {
        auto T = [&](int x, int y, int z) { return snorm(nz.rgba[(((size_t)z * n + y) * n + x) * 4 + ch]); };
        float c00 = T(i0[0], i0[1], i0[2]) * (1.0f - a[0]) + T(i1[0], i0[1], i0[2]) * a[0];
        float c10 = T(i0[0], i1[1], i0[2]) * (1.0f - a[0]) + T(i1[0], i1[1], i0[2]) * a[0];
        float c01 = T(i0[0], i0[1], i1[2]) * (1.0f - a[0]) + T(i1[0], i0[1], i1[2]) * a[0];
        float c11 = T(i0[0], i1[1], i1[2]) * (1.0f - a[0]) + T(i1[0], i1[1], i1[2]) * a[0];
        float c0 = c00 * (1.0f - a[1]) + c10 * a[1];
        float c1 = c01 * (1.0f - a[1]) + c11 * a[1];
        out[ch] = c0 * (1.0f - a[2]) + c1 * a[2];
    }
}

// ---------------------------------------------------------------------------------
// conetrace_frag.glsl helpers
// ---------------------------------------------------------------------------------
struct TraceUniforms {
    crn_trace_params p;
    VolumeUniforms vol;
    vec3 lightPos;               // Sun::position, src/Shaders/ConeTraceShader.cpp:26
    float octaveOffsets[4];      // see noise3D below
    vec3 viewRay;                // normalize(V[0][2],V[1][2],V[2][2]), res/conetrace_frag.glsl:138
};

// noise3D, res/conetrace_frag.glsl:103-120.  `octaveOffsets` is a uniform vec3 indexed
// by the octave number, so octave i adds the SCALAR component i to all three coordinates;
// the host uploads windVel*runTime into it (src/Shaders/ConeTraceShader.cpp:55-61).
// DECREE: component 3 (read when numOctaves == 4, out of bounds in GLSL) is 0.
void noise3D(const Noise &nz, const TraceUniforms &u, vec3 uv, int octaves, float out[4], uint64_t *taps) {
    float acc[4] = {0, 0, 0, 0};
    float freq = 1.0f, pers = 1.0f;
    for (int i = 0; i < octaves; i++) {
        float off = i < 3 ? u.octaveOffsets[i] : 0.0f;
        vec3 uvOffset = {uv.x + off, uv.y + off, uv.z + off};
        float o[4];
        sample_noise(nz, uvOffset * freq, o);
        (*taps)++;
        for (int k = 0; k < 4; k++) acc[k] += pers * o[k];
        freq *= u.p.freqStep;
        pers *= u.p.persStep;
    }
    acc[3] = fabsf(acc[3]);
    for (int k = 0; k < 4; k++) out[k] = acc[k];
}

// traceCone, res/conetrace_frag.glsl:64-79
float trace_cone(const Chain &c, const TraceUniforms &u, vec3 position, vec3 direction, uint64_t *taps) {
    float fd = (float)u.vol.voxelDim;
    direction = normalize(direction);
    direction = direction / fd;
    position = position / fd;
    int steps = u.p.vctSteps;
    float coneHeight = u.p.vctConeInitialHeight;
    float color = 0.0f;
    float tanHalf = tanf(u.p.vctConeAngle / 2.0f);
    for (int i = 1; i <= steps; i++) {
        float coneRadius = coneHeight * tanHalf;
        float lod = log2f(fmaxf(1.0f, 2.0f * coneRadius));
        float s = texture_lod(c, position + direction * coneHeight, lod + u.p.vctLodOffset);
        (*taps)++;
        if (c.alpha) {                                   // paper/tex/conetracing.tex:36-39: if (sampleColor.a > 0.f)
            Chain ca = c; ca.data = c.alpha; ca.fdata = nullptr; ca.alpha = nullptr;
            float al = texture_lod(ca, position + direction * coneHeight, lod + u.p.vctLodOffset);
            if (!(al > 0.0f)) s = 0.0f;
        }
        color += s * (float)i / ((float)steps * u.p.vctDownScaling);
        coneHeight += coneRadius;
    }
    return color;
}

// raySphereIntersect, res/conetrace_frag.glsl:85-101
bool ray_sphere(vec3 rO, vec3 rD, vec3 sO, float sR, float *tnear, float *tfar) {
    vec3 delta = rO - sO;
    float A = dot(rD, rD);
    float B = 2.0f * dot(delta, rD);
    float C = dot(delta, delta) - sR * sR;
    float disc = B * B - 4.0f * A * C;
    if (disc < 0.01f) return false;
    float sq = sqrtf(disc);
    *tnear = (-B - sq) / (2.0f * A);
    *tfar = (-B + sq) / (2.0f * A);
    return true;
}

// main(), res/conetrace_frag.glsl:122-201.  Returns false on discard.
bool conetrace_fragment(const TraceUniforms &u, const Chain &chain, const Noise &nz, const ViewBasis &cam,
                        vec3 fragPos, float fragTexX, float fragTexY, vec3 center, float radius,
                        float color[4], uint64_t *coneTaps, uint64_t *noiseTaps) {
    color[0] = color[1] = color[2] = color[3] = 0.0f;
    if (u.p.showQuad) {                                                   // :124-134
        float sc = distance(center, fragPos) / radius;
        sc = sqrtf(fmaxf(0.0f, 1.0f - sc * sc));
        color[0] = color[1] = color[2] = color[3] = sc;
        if (fragTexX < 0.01f || fragTexY < 0.01f || fragTexX > 0.99f || fragTexY > 0.99f)
            color[0] = color[1] = color[2] = color[3] = 1.0f;
        return true;
    }
    if (u.p.doNoiseSample) {                                              // :137-174
        float tnear, tfar;
        if (!ray_sphere(fragPos, u.viewRay, center, radius, &tnear, &tfar)) return false;
        vec3 worldNear = fragPos + u.viewRay * tnear;
        vec3 worldFar = fragPos + u.viewRay * tfar;
        vec3 unitTex = (worldNear - center) / radius;
        float fNoiseSizeAdjust = 1.0f / u.p.adjustSize;
        vec3 localTexNear = worldNear * fNoiseSizeAdjust;
        vec3 localTexFar = worldFar * fNoiseSizeAdjust;
        float iSteps = length(localTexFar - localTexNear) / u.p.stepSize;
        iSteps = fminf(iSteps, (float)(u.p.maxNoiseSteps - u.p.minNoiseSteps)) + (float)u.p.minNoiseSteps;
        vec3 currentTex = localTexNear;
        vec3 localTexDelta = (localTexFar - localTexNear) / (iSteps - 1.0f);
        float opacityAdjust = u.p.noiseOpacity / (iSteps - 1.0f);
        float lightAdjust = 1.0f / (iSteps - 1.0f);
        float runningOpacity = 0.0f, runningLight = 0.0f;
        for (int i = 0; (float)i < iSteps; i++) {
            float cell[4];
            noise3D(nz, u, currentTex, u.p.numOctaves, cell, noiseTaps);
            vec3 n = normalize(unitTex);
            cell[0] += n.x; cell[1] += n.y; cell[2] += n.z;
            runningOpacity += cell[3] * (1.0f - dot(unitTex, unitTex));
            runningLight += saturate(((cell[0] * 0.0f + cell[1] * 1.0f) + cell[2] * 0.0f) * 0.5f + 0.5f);
            currentTex = currentTex + localTexDelta;
            unitTex = unitTex + localTexDelta;
        }
        float col = u.p.minNoiseColor + (u.p.noiseColorScale * runningLight * lightAdjust);
        float dx = fragTexX - 0.5f, dy = fragTexY - 0.5f;
        float alpha = 1.0f - sqrtf(dx * dx + dy * dy) * 2.0f;
        runningOpacity = saturate(runningOpacity * opacityAdjust);
        color[0] = color[1] = color[2] = col;
        color[3] = runningOpacity * alpha;
    }
    if (u.p.doConeTrace) {                                                // :176-200
        float sc = distance(center, fragPos) / radius;
        sc = sqrtf(fmaxf(0.0f, 1.0f - sc * sc));
        if (sc < 0.01f) return false;
        vec3 pos = fragPos + cam.nrm * radius * sc;
        vec3 voxelPosition = voxel_lerp(u.vol, pos);
        vec3 dir = u.lightPos - pos;
        float indirect = trace_cone(chain, u, voxelPosition, dir, coneTaps);
        if (u.p.doNoiseSample) { color[0] *= indirect; color[1] *= indirect; color[2] *= indirect; }
        else color[0] = color[1] = color[2] = color[3] = indirect;
    }
    return true;
}

// sun_frag.glsl:14-28 drawn through billboard_vert.glsl with M = T(sun)*S(outerRadius)
// (src/Shaders/SunShader.cpp:7-42). Returns false on discard.
bool sun_fragment(const crn_sun &s, vec3 fragPos, float color[4]) {
    float dist = distance(v3(s.position), fragPos);
    if (dist < s.innerRadius) {
        color[0] = s.innerColor[0]; color[1] = s.innerColor[1]; color[2] = s.innerColor[2]; color[3] = 1.0f;
        return true;
    }
    float scale = (dist - s.innerRadius) / (s.outerRadius - s.innerRadius);
    if (scale > 0.99f) return false;
    for (int k = 0; k < 3; k++) color[k] = s.outerColor[k] * scale + s.innerColor[k] * (1.0f - scale);
    color[3] = 1.0f - scale;
    return true;
}

// Fixed-function blend, glBlendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA) on all four channels
// (src/main.cpp:94-95); a fixed-point colour buffer clamps the source to [0,1] first
// (GL 4.4 §17.3.8).  quantize8 models the reference's 8-bit window framebuffer being
// written back after every blend.
inline float q8(float x) { return floorf(saturate(x) * 255.0f + 0.5f) / 255.0f; }
inline void blend(float dst[4], const float src_in[4], bool quantize8) {
    float src[4];
    for (int k = 0; k < 4; k++) src[k] = saturate(src_in[k]);
    float a = src[3];
    for (int k = 0; k < 4; k++) {
        float r = src[k] * a + dst[k] * (1.0f - a);
        dst[k] = quantize8 ? q8(r) : r;
    }
}

} // namespace

// =================================================================================
// C entry points (ctypes-friendly).  Struct types come from include/cloud_renderer_b200.h
// so that the oracle and the library are driven with byte-identical inputs.
// =================================================================================
/* res/first_voxelize.glsl:53-58 (commented out as shipped, live in paper/tex/voxelization.tex:13-27): walk the chord of
 * the sphere under this fragment from the far side towards the light in steps of stepSize, one image store per step.
 * Pinned against the shader itself compiled with those lines switched back on (oracle/ref_glsl: first_voxelize_paper). */
template <class Store>
static inline void paper_march(const VolumeUniforms &vu, int D, vec3 fragPos, vec3 dir, float dist, Store &&store) {
    vec3 start = fragPos - dir * dist;
    for (float s = 0.0f; s < 2.0f * dist; s += vu.stepSize) {
        vec3 f = voxel_lerp(vu, start + dir * s);
        if (!(f.x > -1.0f && f.x < (float)D && f.y > -1.0f && f.y < (float)D && f.z > -1.0f && f.z < (float)D)) continue;   // GL drops the store
        store((int)f.x, (int)f.y, (int)f.z);
    }
}

extern "C" {

typedef struct orc_scene {
    crn_volume_desc vol;
    crn_sun sun;
    crn_camera cam;
    crn_trace_params tp;
    int32_t width, height;
    int32_t n_boards;
    const float *board_pos;      /* n*3 offsets relative to vol.position */
    const float *board_scale;    /* n */
    const int8_t *noise;         /* noise_dim^3 * 4 */
    int32_t noise_dim;
} orc_scene;

typedef struct orc_trace_stats {
    uint64_t fragments, coneSamples, noiseSamples, rectPixels;
} orc_trace_stats;

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline legs ask for every host core explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Sun::update, src/Sun.hpp:26-43 */
void orc_sun_update(const crn_volume_desc *vol, const crn_sun *sun, crn_sun_derived *out) {
    vec3 mn = {vol->xBounds[0], vol->yBounds[0], vol->zBounds[0]};
    vec3 mx = {vol->xBounds[1], vol->yBounds[1], vol->zBounds[1]};
    vec3 vp = v3(vol->position), sp = v3(sun->position);
    vec3 lookDir = normalize(vp - sp);
    float minLen = length(mn), maxLen = length(mx);
    float L = fmaxf(minLen, maxLen);
    vec3 lookPos = vp - lookDir * L;
    glm_lookAt(lookPos, vp, {0.0f, 1.0f, 0.0f}, out->V);
    float minmin = 2.0f * fminf(mn.x, fminf(mn.y, mn.z));
    float maxmax = 2.0f * fmaxf(mx.x, fmaxf(mx.y, mx.z));
    vec3 nearP = lookPos + lookDir * 0.01f;
    vec3 farP = lookPos + lookDir * 2.0f * L;
    out->nearPlane[0] = nearP.x; out->nearPlane[1] = nearP.y; out->nearPlane[2] = nearP.z;
    out->farPlane[0] = farP.x; out->farPlane[1] = farP.y; out->farPlane[2] = farP.z;
    out->clipDistance = distance(nearP, farP);
    glm_ortho(minmin, maxmax, minmin, maxmax, 0.01f, 0.01f + out->clipDistance, out->P);
}

/* Camera::update's matrix block, src/Camera.cpp:59-60: fovy 45 is passed where GLM 0.9.8.5
 * expects radians, and Window::width / Window::height is an integer division. */
void orc_camera_update(int32_t width, int32_t height, const float eye[3], const float lookAt[3], crn_camera *out) {
    float aspect = (float)(width / height);
    glm_perspective(45.0f, aspect, 0.01f, 2500.0f, out->P);
    glm_lookAt(v3(eye), v3(lookAt), {0.0f, 1.0f, 0.0f}, out->V);
    out->position[0] = eye[0]; out->position[1] = eye[1]; out->position[2] = eye[2];
}

/* CloudVolume::sortBoards, src/CloudVolume.cpp:65-82 — the selection sort, literally:
 * picks the farthest remaining board, swaps it into place. In-place on both arrays. */
void orc_sort_boards(float *pos, float *scale, int32_t n, const float volpos[3], const float point[3]) {
    vec3 vp = v3(volpos), pt = v3(point);
    for (int i = 0; i < n; i++) {
        int minIdx = i;
        float dmin = distance(vp + v3(pos + 3 * minIdx), pt);
        for (int j = i + 1; j < n; j++) {
            float dj = distance(vp + v3(pos + 3 * j), pt);
            if (dmin < dj) { minIdx = j; dmin = dj; }
        }
        if (i != minIdx) {
            for (int k = 0; k < 3; k++) std::swap(pos[3 * i + k], pos[3 * minIdx + k]);
            std::swap(scale[i], scale[minIdx]);
        }
    }
}

/* sort key of one board, exported so tests can detect ties */
float orc_board_distance(const float off[3], const float volpos[3], const float point[3]) {
    return distance(v3(volpos) + v3(off), v3(point));
}

/* ConeTraceShader::initNoiseMap, src/Shaders/ConeTraceShader.cpp:100-159, minus the rand()
 * fill (the alpha bytes are the fixture input).  Keeps the operator-precedence quirk
 * (only the second density is divided by heightAdjust).
 * DECREE: (char)(normal*128) saturates to [-128,127] (the C cast of 128.0f is undefined);
 * a zero gradient (NaN normal) stores 0. */
void orc_build_noise(const int8_t *alpha, int32_t dim, int8_t *rgba) {
    auto idx = [dim](int x, int y, int z) {
        if (x < 0) x += dim; if (y < 0) y += dim; if (z < 0) z += dim;
        x %= dim; y %= dim; z %= dim;
        return x + y * dim + z * dim * dim;
    };
    auto density = [&](int i) { return (float)alpha[i] / 128.0f; };
    auto to_char = [](float f) {
        if (!(f == f)) return (int8_t)0;
        float v = f * 128.0f;
        v = fminf(fmaxf(v, -128.0f), 127.0f);
        return (int8_t)(int)v;                       // truncation toward zero like a C cast
    };
    const float heightAdjust = 0.5f;
    for (int z = 0; z < dim; z++)
        for (int y = 0; y < dim; y++)
            for (int x = 0; x < dim; x++) {
                vec3 g;
                g.x = density(idx(x + 1, y, z)) - density(idx(x - 1, y, z)) / heightAdjust;
                g.y = density(idx(x, y + 1, z)) - density(idx(x, y - 1, z)) / heightAdjust;
                g.z = density(idx(x, y, z + 1)) - density(idx(x, y, z - 1)) / heightAdjust;
                vec3 n = normalize(g);
                int i = idx(x, y, z);
                rgba[4 * i + 0] = to_char(n.x);
                rgba[4 * i + 1] = to_char(n.y);
                rgba[4 * i + 2] = to_char(n.z);
                rgba[4 * i + 3] = alpha[i];
            }
}

/* Voxelize pass 1 (src/Shaders/VoxelizeShader.cpp:36-70, res/first_voxelize.glsl:39-64)
 * + pass 2 (src/Shaders/VoxelizeShader.cpp:75-102, res/second_voxelize.glsl:34-51).
 * posmap: W*H*4 floats (xyz world position, a = 1 where a fragment landed), may be NULL.
 * depth:  W*H floats, may be NULL.  level0: D^3 bytes, 0 or 255.
 * DECREES: the depth buffer is float32 (the reference asks for an unsized
 * GL_DEPTH_COMPONENT); GL_LESS against a 1.0 clear, so equal depths keep the earlier
 * instance and a clamped depth of 1 never lands; ivec3() truncates toward zero;
 * imageStore outside [0,D)^3 is dropped (GL 4.4 §8.26). */
static void voxelize_impl(const orc_scene *sc, float *posmap_out, float *depth_out, uint8_t *level0, uint8_t *alpha0);

void orc_voxelize(const orc_scene *sc, float *posmap_out, float *depth_out, uint8_t *level0) {
    voxelize_impl(sc, posmap_out, depth_out, level0, nullptr);
}

/* Paper variant (opt-in, CRN_VOLUME_RG8).  The shipped first pass has its interior march commented
 * out (res/first_voxelize.glsl:53-58); paper/tex/voxelization.tex:13-27 describes it live: every non-discarded
 * fragment of every billboard (image stores are not depth-tested) walks the chord of its sphere from the far side
 * towards the light in steps of `stepSize` and stores (0,0,0,1); pass 2 then stores (1,1,1,1) on the lit shell.
 * level0: the rgb channel (identical to orc_voxelize's), alpha0: the a channel = interior | lit, both D^3 bytes. */
void orc_voxelize_paper(const orc_scene *sc, uint8_t *level0, uint8_t *alpha0) {
    voxelize_impl(sc, nullptr, nullptr, level0, alpha0);
}

static void voxelize_impl(const orc_scene *sc, float *posmap_out, float *depth_out, uint8_t *level0, uint8_t *alpha0) {
    const int W = sc->width, H = sc->height, D = sc->vol.dimension;
    crn_sun_derived sd;
    orc_sun_update(&sc->vol, &sc->sun, &sd);
    ViewBasis lb = make_basis(sd.P, sd.V);
    VolumeUniforms vu = volume_uniforms(sc->vol);
    vec3 nearP = v3(sd.nearPlane);
    float clip = sd.clipDistance;
    vec3 vp = v3(sc->vol.position);

    if (alpha0) std::memset(alpha0, 0, (size_t)D * D * D);
    std::vector<float> posmap((size_t)W * H * 4, 0.0f);       // glClearColor(0,0,0,0)
    std::vector<float> depth((size_t)W * H, 1.0f);            // glClear depth = 1
    std::vector<QuadSetup> quads(sc->n_boards);
    for (int b = 0; b < sc->n_boards; b++)
        quads[b] = quad_setup(lb, vp + v3(sc->board_pos + 3 * b), board_radius(sc->vol, sc->board_scale[b]), W, H);

#pragma omp parallel for schedule(dynamic, 8)
    for (int j = 0; j < H; j++) {
        for (int b = 0; b < sc->n_boards; b++) {               // instance order
            const QuadSetup &q = quads[b];
            if (q.clipped || j < q.j0 || j > q.j1) continue;
            float radius = q.scale;
            for (int i = q.i0; i <= q.i1; i++) {
                float u, v;
                vec3 fragPos = frag_pos(lb, q, i, j, W, H, &u, &v);
                float sphereContrib = distance(q.center, fragPos) / radius;
                sphereContrib = sqrtf(fmaxf(0.0f, 1.0f - sphereContrib * sphereContrib));
                if (sphereContrib < 0.01f) continue;            // discard
                vec3 dir = lb.nrm;
                float dist = radius * sphereContrib;
                if (alpha0)                                     // every thread stores the same constant: the race is benign (and is the shader's own)
                    paper_march(vu, D, fragPos, dir, dist, [&](int x, int y, int z) { alpha0[((size_t)z * D + y) * D + x] = 255; });
                vec3 worldPos = fragPos + dir * dist;
                float d = distance(nearP, worldPos) / clip;
                d = saturate(d);
                size_t t = (size_t)j * W + i;
                if (d < depth[t]) {
                    depth[t] = d;
                    posmap[4 * t + 0] = worldPos.x; posmap[4 * t + 1] = worldPos.y;
                    posmap[4 * t + 2] = worldPos.z; posmap[4 * t + 3] = 1.0f;
                }
            }
        }
    }

    std::memset(level0, 0, (size_t)D * D * D);                 // CloudVolume::clearGPU, src/CloudVolume.cpp:96-100
    const float k = 1.0f / sqrtf(3.0f);                        // normalize(vec3(+-1)) component
    const float delta = vu.stepSize * k;
    auto store = [&](vec3 p) {
        vec3 f = voxel_lerp(vu, p);
        // truncate toward zero; compare in float first so NaN/huge never hit the cast
        if (!(f.x > -1.0f && f.x < (float)D && f.y > -1.0f && f.y < (float)D && f.z > -1.0f && f.z < (float)D)) return;
        int x = (int)f.x, y = (int)f.y, z = (int)f.z;
        level0[((size_t)z * D + y) * D + x] = 255;
        if (alpha0) alpha0[((size_t)z * D + y) * D + x] = 255;             // (1,1,1,1)
    };
    for (size_t t = 0; t < (size_t)W * H; t++) {
        if (!(posmap[4 * t + 3] > 0.0f)) continue;
        vec3 p = {posmap[4 * t], posmap[4 * t + 1], posmap[4 * t + 2]};
        store(p);
        for (int s = 0; s < 8; s++) {                          // order of res/second_voxelize.glsl:42-49
            float sx = (s & 4) ? -delta : delta, sy = (s & 2) ? -delta : delta, sz = (s & 1) ? -delta : delta;
            store({p.x + sx, p.y + sy, p.z + sz});
        }
    }
    if (posmap_out) std::memcpy(posmap_out, posmap.data(), posmap.size() * sizeof(float));
    if (depth_out) std::memcpy(depth_out, depth.data(), depth.size() * sizeof(float));
}

/* glGenerateMipmap(GL_TEXTURE_3D), src/Shaders/VoxelizeShader.cpp:105.
 * DECREE: 2x2x2 box filter, each level re-quantised to UNORM8 with round-half-up:
 * (sum of 8 bytes + 4) >> 3.  chain: levels concatenated, level 0 first. */
void orc_mips(const uint8_t *level0, int32_t D, int32_t levels, uint8_t *chain) {
    size_t n0 = (size_t)D * D * D;
    std::memmove(chain, level0, n0);
    uint8_t *src = chain; int s = D;
    for (int l = 1; l < levels; l++) {
        uint8_t *dst = src + (size_t)s * s * s;
        int h = s / 2;
        for (int z = 0; z < h; z++)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < h; x++) {
                    unsigned sum = 0;
                    for (int dz = 0; dz < 2; dz++)
                        for (int dy = 0; dy < 2; dy++)
                            for (int dx = 0; dx < 2; dx++)
                                sum += src[((size_t)(2 * z + dz) * s + (2 * y + dy)) * s + (2 * x + dx)];
                    dst[((size_t)z * h + y) * h + x] = (uint8_t)((sum + 4) >> 3);
                }
        src = dst; s = h;
    }
}

/* same box filter on float texels, no re-quantisation (CRN_VOLUME_R32F) */
void orc_mips_f32(const float *level0, int32_t D, int32_t levels, float *chain) {
    size_t n0 = (size_t)D * D * D;
    std::memmove(chain, level0, n0 * sizeof(float));
    float *src = chain; int s = D;
    for (int l = 1; l < levels; l++) {
        float *dst = src + (size_t)s * s * s;
        int h = s / 2;
        for (int z = 0; z < h; z++)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < h; x++) {
                    auto S = [&](int dx, int dy, int dz) { return src[((size_t)(2 * z + dz) * s + (2 * y + dy)) * s + (2 * x + dx)]; };
                    float a = (S(0, 0, 0) + S(1, 0, 0)) + (S(0, 1, 0) + S(1, 1, 0));
                    float b = (S(0, 0, 1) + S(1, 0, 1)) + (S(0, 1, 1) + S(1, 1, 1));
                    dst[((size_t)z * h + y) * h + x] = (a + b) * 0.125f;
                }
        src = dst; s = h;
    }
}

/* The frame's colour passes: glClear (src/main.cpp:112-113), optional sun pass
 * (src/main.cpp:116), then ConeTraceShader::coneTrace's instanced draw in ARRAY ORDER
 * (src/Shaders/ConeTraceShader.cpp:22-75; call orc_sort_boards first, as coneTrace does
 * at :20) with depth test off and alpha blending.
 * chain_bytes: the mip chain, levels concatenated — uint8 texels, or float texels when vol.format is
 * CRN_VOLUME_R32F (an extension: the shipped reference only has R8, src/CloudVolume.cpp:18).
 * image_f32: W*H*4 floats (always written).  image_u8: W*H*4 bytes or NULL.
 * quantize_fb8: write every blend result back through 8 bits like the window framebuffer.
 * rows [row0,row1) only (others left untouched) — lets the baseline time a crop. */
void orc_cone_trace(const orc_scene *sc, const uint8_t *chain_bytes, float *image_f32, uint8_t *image_u8,
                    int32_t quantize_fb8, int32_t row0, int32_t row1, orc_trace_stats *stats) {
    const int W = sc->width, H = sc->height;
    row0 = std::max(0, row0); row1 = std::min(H, row1);
    ViewBasis cam = make_basis(sc->cam.P, sc->cam.V);
    Chain chain = scene_chain(sc->vol, chain_bytes);
    Noise nz = {sc->noise, sc->noise_dim};
    TraceUniforms u;
    u.p = sc->tp;
    u.vol = volume_uniforms(sc->vol);
    u.lightPos = v3(sc->sun.position);
    for (int k = 0; k < 3; k++) u.octaveOffsets[k] = sc->tp.windVel[k] * sc->tp.runTime;
    u.octaveOffsets[3] = 0.0f;
    u.viewRay = normalize({sc->cam.V[2], sc->cam.V[6], sc->cam.V[10]});
    const bool active = sc->tp.doConeTrace || sc->tp.doNoiseSample || sc->tp.showQuad;   // ConeTraceShader.cpp:16-18
    vec3 vp = v3(sc->vol.position);

    std::vector<QuadSetup> quads(sc->n_boards);
    for (int b = 0; b < sc->n_boards; b++)
        quads[b] = quad_setup(cam, vp + v3(sc->board_pos + 3 * b), board_radius(sc->vol, sc->board_scale[b]), W, H);
    QuadSetup sunq = quad_setup(cam, v3(sc->sun.position), sc->sun.outerRadius, W, H);

    uint64_t nfrag = 0, ncone = 0, nnoise = 0, nrect = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nfrag, ncone, nnoise, nrect)
    for (int j = row0; j < row1; j++) {
        float *row = image_f32 + (size_t)j * W * 4;
        for (int i = 0; i < W; i++)
            for (int k = 0; k < 4; k++) row[4 * i + k] = quantize_fb8 ? q8(sc->tp.clearColor[k]) : sc->tp.clearColor[k];
        if (sc->tp.drawSun && !sunq.clipped && j >= sunq.j0 && j <= sunq.j1) {
            for (int i = sunq.i0; i <= sunq.i1; i++) {
                float uu, vv;
                vec3 fragPos = frag_pos(cam, sunq, i, j, W, H, &uu, &vv);
                if (!(fabsf(uu) < sunq.scale && fabsf(vv) < sunq.scale)) continue;     // outside the quad
                float c[4];
                if (sun_fragment(sc->sun, fragPos, c)) blend(row + 4 * i, c, quantize_fb8 != 0);
            }
        }
        if (!active) continue;
        for (int b = 0; b < sc->n_boards; b++) {
            const QuadSetup &q = quads[b];
            if (q.clipped || j < q.j0 || j > q.j1) continue;
            for (int i = q.i0; i <= q.i1; i++) {
                nrect++;
                float uu, vv;
                vec3 fragPos = frag_pos(cam, q, i, j, W, H, &uu, &vv);
                if (!(fabsf(uu) < q.scale && fabsf(vv) < q.scale)) continue;           // outside the quad
                float tx = (uu / q.scale + 1.0f) / 2.0f, ty = (vv / q.scale + 1.0f) / 2.0f;   // fragTex
                float c[4];
                uint64_t ct = 0, nt = 0;
                if (!conetrace_fragment(u, chain, nz, cam, fragPos, tx, ty, q.center, q.scale, c, &ct, &nt)) continue;
                nfrag++; ncone += ct; nnoise += nt;
                blend(row + 4 * i, c, quantize_fb8 != 0);
            }
        }
    }
    if (image_u8)
        for (int j = row0; j < row1; j++)
            for (int i = 0; i < W * 4; i++)
                image_u8[(size_t)j * W * 4 + i] = (uint8_t)floorf(saturate(image_f32[(size_t)j * W * 4 + i]) * 255.0f + 0.5f);
    if (stats) { stats->fragments = nfrag; stats->coneSamples = ncone; stats->noiseSamples = nnoise; stats->rectPixels = nrect; }
}

/* The fixed-function texture units, exported for oracle/ref_glsl (the reference's compiled
 * shaders call back into these; GL 4.4 §8.14 restated above). */
void orc_sample_volume(const uint8_t *chain_bytes, int32_t dim, int32_t levels, float u, float v, float w, float lod, float out[4]) {
    Chain c = make_chain(chain_bytes, dim, levels);
    float r = texture_lod(c, {u, v, w}, lod);
    out[0] = r; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f;              /* GL_R8: (r,0,0,1) */
}
void orc_sample_noise(const int8_t *rgba, int32_t dim, float u, float v, float w, float out[4]) {
    Noise nz = {rgba, dim};
    sample_noise(nz, {u, v, w}, out);
}

/* The LOD textureLod() is called with at every traceCone step (res/conetrace_frag.glsl:70-76);
 * identical for every fragment.  lods[vctSteps], heights[vctSteps]. */
void orc_cone_lods(const crn_trace_params *tp, float *lods, float *heights) {
    float coneHeight = tp->vctConeInitialHeight;
    float tanHalf = tanf(tp->vctConeAngle / 2.0f);
    for (int i = 1; i <= tp->vctSteps; i++) {
        float coneRadius = coneHeight * tanHalf;
        lods[i - 1] = log2f(fmaxf(1.0f, 2.0f * coneRadius)) + tp->vctLodOffset;
        heights[i - 1] = coneHeight;
        coneHeight += coneRadius;
    }
}

/* One fragment of conetrace_frag.glsl at an explicit (fragPos, fragTex, center, radius):
 * the probe tests/test_oracle_vs_ref_glsl.py compares against the compiled reference
 * shader.  Returns 0 on discard. */
int32_t orc_conetrace_fragment(const orc_scene *sc, const uint8_t *chain_bytes, const float fragPos[3],
                               const float fragTex[2], const float center[3], float radius, float color[4]) {
    ViewBasis cam = make_basis(sc->cam.P, sc->cam.V);
    Chain chain = scene_chain(sc->vol, chain_bytes);
    Noise nz = {sc->noise, sc->noise_dim};
    TraceUniforms u;
    u.p = sc->tp;
    u.vol = volume_uniforms(sc->vol);
    u.lightPos = v3(sc->sun.position);
    for (int k = 0; k < 3; k++) u.octaveOffsets[k] = sc->tp.windVel[k] * sc->tp.runTime;
    u.octaveOffsets[3] = 0.0f;
    u.viewRay = normalize({sc->cam.V[2], sc->cam.V[6], sc->cam.V[10]});
    uint64_t a = 0, b = 0;
    return conetrace_fragment(u, chain, nz, cam, v3(fragPos), fragTex[0], fragTex[1], v3(center), radius, color, &a, &b) ? 1 : 0;
}

/* One fragment of first_voxelize.glsl at an explicit fragPos: writes worldPos + depth. */
int32_t orc_first_voxelize_fragment(const orc_scene *sc, const float fragPos[3], const float center[3], float radius,
                                    float worldPos[3], float *depth) {
    crn_sun_derived sd;
    orc_sun_update(&sc->vol, &sc->sun, &sd);
    ViewBasis lb = make_basis(sd.P, sd.V);
    float s = distance(v3(center), v3(fragPos)) / radius;
    s = sqrtf(fmaxf(0.0f, 1.0f - s * s));
    if (s < 0.01f) return 0;
    vec3 wp = v3(fragPos) + lb.nrm * (radius * s);
    worldPos[0] = wp.x; worldPos[1] = wp.y; worldPos[2] = wp.z;
    *depth = distance(v3(sd.nearPlane), wp) / sd.clipDistance;
    return 1;
}

/* The image stores of the paper variant's interior march for one fragment of first_voxelize.glsl at an explicit fragPos
 * (in-range stores only, in march order); returns their number, at most `cap` are written to idx_out (3 ints each). */
int32_t orc_first_voxelize_march(const orc_scene *sc, const float fragPos[3], const float center[3], float radius, int32_t *idx_out, int32_t cap) {
    crn_sun_derived sd;
    orc_sun_update(&sc->vol, &sc->sun, &sd);
    ViewBasis lb = make_basis(sd.P, sd.V);
    VolumeUniforms vu = volume_uniforms(sc->vol);
    float s = distance(v3(center), v3(fragPos)) / radius;
    s = sqrtf(fmaxf(0.0f, 1.0f - s * s));
    if (s < 0.01f) return -1;
    int32_t n = 0;
    paper_march(vu, sc->vol.dimension, v3(fragPos), lb.nrm, radius * s, [&](int x, int y, int z) {
        if (n < cap) { idx_out[3 * n] = x; idx_out[3 * n + 1] = y; idx_out[3 * n + 2] = z; }
        n++;
    });
    return n;
}

/* The 9 voxel indices second_voxelize.glsl stores for one position-map texel; -1 marks a
 * dropped (out of range) store.  out: 9*3 ints. */
void orc_second_voxelize_indices(const crn_volume_desc *vol, const float worldPos[3], int32_t *out) {
    VolumeUniforms vu = volume_uniforms(*vol);
    const int D = vol->dimension;
    const float delta = vu.stepSize * (1.0f / sqrtf(3.0f));
    vec3 p = v3(worldPos);
    for (int s = 0; s < 9; s++) {
        vec3 q = p;
        if (s > 0) {
            int m = s - 1;
            q = {p.x + ((m & 4) ? -delta : delta), p.y + ((m & 2) ? -delta : delta), p.z + ((m & 1) ? -delta : delta)};
        }
        vec3 f = voxel_lerp(vu, q);
        bool ok = f.x > -1.0f && f.x < (float)D && f.y > -1.0f && f.y < (float)D && f.z > -1.0f && f.z < (float)D;
        out[3 * s + 0] = ok ? (int)f.x : -1; out[3 * s + 1] = ok ? (int)f.y : -1; out[3 * s + 2] = ok ? (int)f.z : -1;
    }
}

/* Enumerates, in draw order, the fragments the rasteriser hands to the fragment stage for the
 * billboard draw of pass `which` (0: light camera / first_voxelize, 1: user camera / conetrace):
 * every pixel centre inside each quad, with the interpolated attributes.  Lets a test run the
 * reference's compiled shader on exactly the fragments the oracle shades.  Returns the count
 * (fills at most `cap`).  rec: {i, j, board, fragPos[3], fragTex[2]} as 8 floats per fragment. */
int64_t orc_list_fragments(const orc_scene *sc, int32_t which, int64_t cap, float *rec) {
    crn_sun_derived sd;
    ViewBasis vb;
    if (which == 0) { orc_sun_update(&sc->vol, &sc->sun, &sd); vb = make_basis(sd.P, sd.V); }
    else vb = make_basis(sc->cam.P, sc->cam.V);
    const int W = sc->width, H = sc->height;
    int64_t n = 0;
    for (int b = 0; b < sc->n_boards; b++) {
        QuadSetup q = quad_setup(vb, v3(sc->vol.position) + v3(sc->board_pos + 3 * b), board_radius(sc->vol, sc->board_scale[b]), W, H);
        if (q.clipped) continue;
        for (int j = q.j0; j <= q.j1; j++)
            for (int i = q.i0; i <= q.i1; i++) {
                float uu, vv;
                vec3 fp = frag_pos(vb, q, i, j, W, H, &uu, &vv);
                if (!(fabsf(uu) < q.scale && fabsf(vv) < q.scale)) continue;
                if (n < cap) {
                    float *r = rec + 8 * n;
                    r[0] = (float)i; r[1] = (float)j; r[2] = (float)b; r[3] = fp.x; r[4] = fp.y; r[5] = fp.z;
                    r[6] = (uu / q.scale + 1.0f) / 2.0f; r[7] = (vv / q.scale + 1.0f) / 2.0f;
                }
                n++;
            }
    }
    return n;
}

/* Conservative window-space rectangle of one billboard quad as both passes rasterise it
 * (which = 0: light camera from Sun::update; 1: user camera).  rect = {i0,i1,j0,j1}
 * inclusive, or i1<i0 when the quad is clipped.  The binning contract is defined on it. */
void orc_board_rect(const orc_scene *sc, int32_t which, int32_t b, int32_t rect[4]) {
    crn_sun_derived sd;
    ViewBasis vb;
    if (which == 0) { orc_sun_update(&sc->vol, &sc->sun, &sd); vb = make_basis(sd.P, sd.V); }
    else vb = make_basis(sc->cam.P, sc->cam.V);
    QuadSetup q = quad_setup(vb, v3(sc->vol.position) + v3(sc->board_pos + 3 * b),
                             board_radius(sc->vol, sc->board_scale[b]), sc->width, sc->height);
    rect[0] = q.i0; rect[1] = q.i1; rect[2] = q.j0; rect[3] = q.j1;
}

} // extern "C"
