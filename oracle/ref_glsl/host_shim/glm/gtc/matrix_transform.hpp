// host_shim/glm/gtc/matrix_transform.hpp — GLM 0.9.8.5 lookAt / ortho / perspective (RH, -1..1 depth), as published
// in glm/gtc/matrix_transform.inl (lookAtRH, ortho with zNear/zFar, perspectiveRH).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../glm.hpp"

namespace glm {

inline mat4 lookAt(const vec3 &eye, const vec3 &center, const vec3 &up) {
    const vec3 f(normalize(center - eye));
    const vec3 s(normalize(cross(f, up)));
    const vec3 u(cross(s, f));
    mat4 R(1.0f);
    R[0][0] = s.x; R[1][0] = s.y; R[2][0] = s.z;
    R[0][1] = u.x; R[1][1] = u.y; R[2][1] = u.z;
    R[0][2] = -f.x; R[1][2] = -f.y; R[2][2] = -f.z;
    R[3][0] = -dot(s, eye); R[3][1] = -dot(u, eye); R[3][2] = dot(f, eye);
    return R;
}

inline mat4 ortho(float left, float right, float bottom, float top, float zNear, float zFar) {
    mat4 R(1.0f);
    R[0][0] = 2.0f / (right - left);
    R[1][1] = 2.0f / (top - bottom);
    R[2][2] = -2.0f / (zFar - zNear);
    R[3][0] = -(right + left) / (right - left);
    R[3][1] = -(top + bottom) / (top - bottom);
    R[3][2] = -(zFar + zNear) / (zFar - zNear);
    return R;
}

inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
    const float tanHalfFovy = std::tan(fovy / 2.0f);
    mat4 R(0.0f);
    R[0][0] = 1.0f / (aspect * tanHalfFovy);
    R[1][1] = 1.0f / tanHalfFovy;
    R[2][3] = -1.0f;
    R[2][2] = -(zFar + zNear) / (zFar - zNear);
    R[3][2] = -(2.0f * zFar * zNear) / (zFar - zNear);
    return R;
}

} // namespace glm
