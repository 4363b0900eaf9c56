// glsl_shim.hpp — just enough GLSL vocabulary, in C++, for the reference's shader SOURCES
// (/root/reference/res/*.glsl, read in place, never copied) to compile as C++ functions.
// TEST INFRASTRUCTURE ONLY (part of the oracle; see oracle/oracle.cpp's header).
//
// What comes from the reference: every statement of every shader main() and helper.
// What this file supplies: vector/matrix types with the swizzles those shaders use, the GLSL
// built-ins they call, `discard`, and the storage-qualifier keywords as no-ops.  Texture and
// image access go through callbacks (the fixed-function units are not in the reference).
// Float semantics: float32 everywhere, one rounding per operation (build with
// -ffp-contract=off); normalize(v) = v * (1/sqrt(dot(v,v))), distance(a,b) = length(a-b).
#pragma once
#include <cmath>
#include <cstdint>

namespace glsl {

struct vec2; struct vec3; struct vec4;

// ---- swizzle proxies: trivially-copyable views over the parent's storage -----------------
template <int A, int B> struct swz2 {
    float d[4];
    operator vec2() const;
    swz2 &operator=(const vec2 &v);
};
template <int A, int B, int C> struct swz3 {
    float d[4];
    operator vec3() const;
    swz3 &operator=(const vec3 &v);
    swz3 &operator+=(const vec3 &v);
    swz3 &operator*=(float s);
};

struct vec2 {
    union { struct { float x, y; }; struct { float r, g; }; };
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
};
struct vec3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; swz2<0, 1> xy; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    // `uniform vec3 octaveOffsets` is indexed by the octave counter in conetrace_frag.glsl:110;
    // index 3 is out of bounds in GLSL (undefined): DECREE 4 in DESIGN.md -> reads 0.
    float operator[](int i) const { return i == 0 ? x : i == 1 ? y : i == 2 ? z : 0.0f; }
    vec3 &operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
    vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3 &operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};
struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<0, 1> xy;
        swz3<0, 1, 2> xyz;
        swz3<0, 1, 2> rgb;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3 &v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    vec4 &operator+=(const vec4 &o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
};
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} explicit ivec2(const vec2 &v) : x((int)v.x), y((int)v.y) {} };
struct ivec3 { int x, y, z; ivec3() : x(0), y(0), z(0) {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
               explicit ivec3(const vec3 &v) : x((int)v.x), y((int)v.y), z((int)v.z) {} };   // truncation toward zero
struct ivec4 { int x, y, z, w; ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {} };

template <int A, int B> swz2<A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int A, int B> swz2<A, B> &swz2<A, B>::operator=(const vec2 &v) { d[A] = v.x; d[B] = v.y; return *this; }
template <int A, int B, int C> swz3<A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int A, int B, int C> swz3<A, B, C> &swz3<A, B, C>::operator=(const vec3 &v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }
template <int A, int B, int C> swz3<A, B, C> &swz3<A, B, C>::operator+=(const vec3 &v) { d[A] += v.x; d[B] += v.y; d[C] += v.z; return *this; }
template <int A, int B, int C> swz3<A, B, C> &swz3<A, B, C>::operator*=(float s) { d[A] *= s; d[B] *= s; d[C] *= s; return *this; }

// ---- arithmetic ---------------------------------------------------------------------------
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec4 operator*(float s, vec4 a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float log2(float x) { return ::log2f(x); }
inline float abs(float x) { return ::fabsf(x); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float clamp(float x, float lo, float hi) { return ::fminf(::fmaxf(x, lo), hi); }
inline float length(vec2 v) { return ::sqrtf(dot(v, v)); }
inline float length(vec3 v) { return ::sqrtf(dot(v, v)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline vec3 normalize(vec3 v) { return v * (1.0f / ::sqrtf(dot(v, v))); }

// ---- matrices (column-major, m[col][row]) --------------------------------------------------
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(const mat4 &m) { for (int i = 0; i < 3; i++) c[i] = vec3(m[i].x, m[i].y, m[i].z); }
};
inline vec4 operator*(const mat4 &m, const vec4 &v) {
    vec4 r;
    for (int i = 0; i < 4; i++) r[i] = ((m[0][i] * v.x + m[1][i] * v.y) + m[2][i] * v.z) + m[3][i] * v.w;
    return r;
}
inline vec4 operator*(const vec4 &v, const mat4 &m) { return vec4(dot(v, m[0]), dot(v, m[1]), dot(v, m[2]), dot(v, m[3])); }
inline mat4 operator*(const mat4 &a, const mat4 &b) { mat4 r; for (int j = 0; j < 4; j++) r[j] = a * b[j]; return r; }
inline vec3 operator*(const mat3 &m, const vec3 &v) {
    return vec3((m.c[0].x * v.x + m.c[1].x * v.y) + m.c[2].x * v.z, (m.c[0].y * v.x + m.c[1].y * v.y) + m.c[2].y * v.z,
                (m.c[0].z * v.x + m.c[1].z * v.y) + m.c[2].z * v.z);
}
inline mat4 transpose(const mat4 &m) { mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = m[j][i]; return r; }
inline mat4 inverse(const mat4 &m) {          // cofactor expansion (what GLSL inverse() specifies mathematically)
    float a[16], inv[16];
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) a[c * 4 + r] = m[c][r];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    mat4 r;
    for (int c = 0; c < 4; c++) for (int rr = 0; rr < 4; rr++) r[c][rr] = inv[c * 4 + rr] / det;
    return r;
}

// ---- texture / image units: callbacks into the fixed-function restatement --------------------
struct sampler3D {
    void *user;
    void (*fetch)(void *user, float u, float v, float w, float lod, int has_lod, float out[4]);
};
inline vec4 texture(const sampler3D &s, vec3 uvw) { float o[4]; s.fetch(s.user, uvw.x, uvw.y, uvw.z, 0.0f, 0, o); return vec4(o[0], o[1], o[2], o[3]); }
inline vec4 textureLod(const sampler3D &s, vec3 uvw, float lod) { float o[4]; s.fetch(s.user, uvw.x, uvw.y, uvw.z, lod, 1, o); return vec4(o[0], o[1], o[2], o[3]); }
struct image3D {
    int n;                      // stores recorded so far
    int idx[1024][3];           // (the paper variant's first-pass march stores one voxel per step of the chord)
    float val[1024][4];
};
struct image2D { const float *texels; int width, height; };
inline void imageStore(image3D &img, ivec3 p, vec4 v) {
    if (img.n < 1024) { img.idx[img.n][0] = p.x; img.idx[img.n][1] = p.y; img.idx[img.n][2] = p.z;
                      img.val[img.n][0] = v.x; img.val[img.n][1] = v.y; img.val[img.n][2] = v.z; img.val[img.n][3] = v.w; img.n++; }
}
inline void imageStore(image3D &img, ivec3 p, ivec4 v) { imageStore(img, p, vec4((float)v.x, (float)v.y, (float)v.z, (float)v.w)); }
inline vec4 imageLoad(const image2D &img, ivec2 p) {
    const float *t = img.texels + ((size_t)p.y * img.width + p.x) * 4;
    return vec4(t[0], t[1], t[2], t[3]);
}

} // namespace glsl

