// ref_host.cpp — C entry points around the reference's own HOST code, compiled from where it lies under
// /root/reference/src (see build_ref.sh): Sun::update (src/Sun.hpp:26-43), Camera::update (src/Camera.cpp:12-61),
// CloudVolume's constructor head / update / sortBoards / get3DIndices / reverseVoxelIndex (src/CloudVolume.cpp),
// the normal loop of ConeTraceShader::initNoiseMap (src/Shaders/ConeTraceShader.cpp:100-151).  TEST INFRASTRUCTURE ONLY.
// What is NOT reference code here: GLM itself (host_shim/glm: un-vendored third party, restated from its published
// definitions), the GLFW/glad stubs, and the input stubs below (no key or mouse button is ever down).
#define protected public        // Camera keeps phi/theta/position/lookAt protected; the harness has to set them
#include "Camera.hpp"
#include "Sun.hpp"
#include "CloudVolume.hpp"
#undef protected

#include <cstdint>
#include <cstring>
#include <new>

// ---- statics the reference defines in translation units that are not compiled here (src/IO/*.cpp) ----
int Window::width = 1280, Window::height = 720;
double Mouse::dx = 0.0, Mouse::dy = 0.0;
bool Mouse::isDown(int) { return false; }
bool Keyboard::isKeyPressed(int) { return false; }
void CloudVolume::uploadBillboards() {}             // GL buffer upload: not part of the arithmetic

struct CHAR4 { char r, g, b, a; };
void ref_noise_normals_impl(CHAR4 *pData, int dimension);   // host_noise translation unit

namespace {
void store_mat(const glm::mat4 &m, float *o) { for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) o[c * 4 + r] = m[c][r]; }
CloudVolume *make_volume(int dim, const float pos[3], const float xb[2], const float yb[2], const float zb[2], int levels) {
    // the constructor's own assignments (src/CloudVolume.cpp:7-13) run; bounds may differ per axis afterwards
    CloudVolume *v = new CloudVolume(dim, glm::vec2(xb[0], xb[1]), glm::vec3(pos[0], pos[1], pos[2]), levels);
    v->xBounds = glm::vec2(xb[0], xb[1]); v->yBounds = glm::vec2(yb[0], yb[1]); v->zBounds = glm::vec2(zb[0], zb[1]);
    return v;
}
} // namespace

extern "C" {

// Sun::update(vol): V, P, nearPlane, farPlane, clipDistance
void ref_host_sun_update(const float volpos[3], const float xb[2], const float yb[2], const float zb[2], const float sunpos[3],
                         float V[16], float P[16], float nearPlane[3], float farPlane[3], float *clip) {
    CloudVolume *v = make_volume(32, volpos, xb, yb, zb, 1);
    Sun::position = glm::vec3(sunpos[0], sunpos[1], sunpos[2]);
    Sun::update(v);
    store_mat(Sun::V, V); store_mat(Sun::P, P);
    for (int k = 0; k < 3; k++) { nearPlane[k] = Sun::nearPlane[k]; farPlane[k] = Sun::farPlane[k]; }
    *clip = Sun::clipDistance;
    delete v;
}

// the defaults the reference starts with (src/main.cpp:37-46)
void ref_host_sun_defaults(float position[3], float innerColor[3], float outerColor[3], float *innerRadius, float *outerRadius) {
    for (int k = 0; k < 3; k++) { position[k] = Sun::position[k]; innerColor[k] = Sun::innerColor[k]; outerColor[k] = Sun::outerColor[k]; }
    *innerRadius = Sun::innerRadius; *outerRadius = Sun::outerRadius;
}

// Camera::update() with the given eye and view angles (no input): P, V and the look-at point it derives
void ref_host_camera_update(int width, int height, const float position[3], double phi, double theta, float P[16], float V[16], float lookAt[3]) {
    Window::width = width; Window::height = height;
    Camera::position = glm::vec3(position[0], position[1], position[2]);
    Camera::phi = phi; Camera::theta = theta;
    Camera::update();
    store_mat(Camera::getP(), P); store_mat(Camera::getV(), V);
    const glm::vec3 l = Camera::getLookAt();
    lookAt[0] = l.x; lookAt[1] = l.y; lookAt[2] = l.z;
}

// CloudVolume::sortBoards(point), in place on the caller's arrays
void ref_host_sort_boards(float *pos3, float *scale, int n, const float volpos[3], const float point[3]) {
    const float b[2] = {-5.0f, 5.0f};
    CloudVolume *v = make_volume(32, volpos, b, b, b, 1);
    for (int i = 0; i < n; i++) { glm::vec3 p(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]); float s = scale[i]; v->addCloudBoard(p, s); }
    v->sortBoards(glm::vec3(point[0], point[1], point[2]));
    for (int i = 0; i < n; i++) {
        pos3[3 * i] = v->billboards.positions[i].x; pos3[3 * i + 1] = v->billboards.positions[i].y; pos3[3 * i + 2] = v->billboards.positions[i].z;
        scale[i] = v->billboards.scales[i];
    }
    delete v;
}

// CloudVolume::update's range / voxelSize, get3DIndices and reverseVoxelIndex for one linear index
void ref_host_voxel_index(int dim, const float volpos[3], const float xb[2], const float yb[2], const float zb[2], int index,
                          int ijk[3], float world[3], float voxelSize[3]) {
    CloudVolume *v = make_volume(dim, volpos, xb, yb, zb, 1);
    v->update();
    const glm::ivec3 i = v->get3DIndices(index);
    const glm::vec3 w = v->reverseVoxelIndex(i);
    ijk[0] = i.x; ijk[1] = i.y; ijk[2] = i.z;
    world[0] = w.x; world[1] = w.y; world[2] = w.z;
    voxelSize[0] = v->voxelSize.x; voxelSize[1] = v->voxelSize.y; voxelSize[2] = v->voxelSize.z;
    delete v;
}

// the normal loop of ConeTraceShader::initNoiseMap on a texture whose alpha channel is given (rgba[dim^3][4], in place)
void ref_host_noise_normals(int8_t *rgba, int dim) { ref_noise_normals_impl(reinterpret_cast<CHAR4 *>(rgba), dim); }

} // extern "C"
