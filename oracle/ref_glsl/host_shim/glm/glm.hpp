// host_shim/glm/glm.hpp — the slice of GLM 0.9.8.5 (Clouds.vcxproj:118; not vendored in /root/reference) that the
// reference's host code on the hot path calls, so that its OWN sources (src/Sun.hpp, src/Camera.cpp,
// src/CloudVolume.cpp, src/Shaders/ConeTraceShader.cpp) compile unmodified.  TEST INFRASTRUCTURE ONLY.
// Arithmetic follows GLM's published definitions, float32, one rounding per written operation:
//   dot = x*x' + y*y' + z*z' (left to right), length = sqrt(dot), normalize = v * (1 / sqrt(dot)),
//   cross as in detail/func_geometric.inl, lookAt/ortho/perspective = the RH, -1..1-depth variants
//   (no GLM_FORCE_* define in Clouds.vcxproj:129-181).
#pragma once
#include <cmath>

namespace glm {

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
};

struct vec3 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    template <class A, class B, class C> vec3(A a, B b_, C c) : x((float)a), y((float)b_), z((float)c) {}
    float &operator[](int i) { return i == 0 ? x : i == 1 ? y : z; }
    const float &operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
};

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
};

struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};

struct mat4 {                               // column-major, m[col][row]
    vec4 c[4];
    mat4() {}
    explicit mat4(float d) { c[0].x = d; c[1].y = d; c[2].z = d; c[3].w = d; }
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};

inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 &operator+=(vec3 &a, const vec3 &b) { a = a + b; return a; }
inline vec3 &operator-=(vec3 &a, const vec3 &b) { a = a - b; return a; }

inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline float cos(float a) { return std::cos(a); }
inline double cos(double a) { return std::cos(a); }
inline float sin(float a) { return std::sin(a); }
inline double sin(double a) { return std::sin(a); }

inline float dot(const vec3 &a, const vec3 &b) { const vec3 t(a.x * b.x, a.y * b.y, a.z * b.z); return t.x + t.y + t.z; }
inline float length(const vec3 &v) { return std::sqrt(dot(v, v)); }
inline float distance(const vec3 &a, const vec3 &b) { return length(b - a); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3 &v) { return v * inversesqrt(dot(v, v)); }
inline vec3 cross(const vec3 &x, const vec3 &y) {
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}

} // namespace glm
