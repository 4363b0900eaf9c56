// cloud_renderer_b200.hpp — header-only C++ mirror of the reference's parameter surface on top of
// the C-ABI (cloud_renderer_b200.h).  Class, method and field names are the reference's, so the
// frame loop of src/main.cpp:98-124 compiles against it unchanged apart from the GL/GLFW lines:
//
//     Camera::update(); Sun::update(volume); volume->update();
//     voxelizeShader->voxelize(volume); coneShader->coneTrace(volume);
//
// GLM is not a dependency: vec2/vec3/mat4 below are layout-compatible PODs (a glm::vec3* can be
// reinterpret_cast to crn::vec3*; mat4 is column-major like glm::value_ptr).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "cloud_renderer_b200.h"

namespace crn {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct ivec3 { int x, y, z; };
struct mat4 { float m[16]; };

inline void check(crn_ctx *ctx, int rc) {
    if (rc != CRN_OK) throw std::runtime_error(std::string("cloud_renderer_b200: ") + crn_last_error(ctx));
}

// Window::width / height / runTime (src/IO/Window.hpp:15-53): position-map + image size, wind clock
struct Window {
    static inline int width = 1280, height = 720;      // src/main.cpp:25-26
    static inline float runTime = 0.f;
};

// src/Camera.hpp:15-29 — only the matrices and the position are on the hot path
class Camera {
public:
    static void update() {                               // src/Camera.cpp:59-60
        crn_camera c;
        const float e[3] = {position.x, position.y, position.z}, l[3] = {lookAt.x, lookAt.y, lookAt.z};
        crn_camera_update(Window::width, Window::height, e, l, &c);
        for (int i = 0; i < 16; i++) { P.m[i] = c.P[i]; V.m[i] = c.V[i]; }
    }
    static mat4 &getP() { return P; }
    static mat4 &getV() { return V; }
    static vec3 getPosition() { return position; }
    static vec3 getLookAt() { return lookAt; }
    static inline vec3 position{0.f, 0.f, 0.f}, lookAt{1.f, 0.f, 0.f};   // zero-initialised statics + first look-at update
    static inline mat4 P{}, V{};
};

// src/CloudVolume.hpp:12-55
class CloudVolume {
public:
    struct Billboards {
        std::vector<vec3> positions;
        std::vector<float> scales;
        int count = 0;
        vec3 minOffset{-1.f, -1.f, -1.f}, maxOffset{1.f, 1.f, 1.f};
        float minScale = 1.f, maxScale = 1.f;
    };

    CloudVolume(int dim, vec2 bounds, vec3 position, int mips, int device = 0, void *stream = nullptr)
        : position(position), xBounds(bounds), yBounds(bounds), zBounds(bounds), dimension(dim), levels(mips) {
        int rc = crn_create(device, stream, &ctx);
        if (rc != CRN_OK) throw std::runtime_error(std::string("cloud_renderer_b200: ") + crn_last_error(nullptr));
        update();
    }
    ~CloudVolume() { crn_destroy(ctx); }
    CloudVolume(const CloudVolume &) = delete;
    CloudVolume &operator=(const CloudVolume &) = delete;

    void update() {                                      // src/CloudVolume.cpp:84-93
        uploadBillboards();
        range = {xBounds.y - xBounds.x, yBounds.y - yBounds.x, zBounds.y - zBounds.x};
        voxelSize = {range.x / (float)dimension, range.y / (float)dimension, range.z / (float)dimension};
        crn_volume_desc d{};
        d.dimension = dimension; d.levels = levels;
        d.position[0] = position.x; d.position[1] = position.y; d.position[2] = position.z;
        d.xBounds[0] = xBounds.x; d.xBounds[1] = xBounds.y; d.yBounds[0] = yBounds.x; d.yBounds[1] = yBounds.y;
        d.zBounds[0] = zBounds.x; d.zBounds[1] = zBounds.y;
        d.fluffiness = fluffiness; d.format = format;
        check(ctx, crn_set_volume(ctx, &d));
    }
    void clearGPU() {}                                   // the clear is part of crn_voxelize (src/CloudVolume.cpp:96-100)
    void addCloudBoard(vec3 &pos, float &scale) {        // src/CloudVolume.cpp:58-62
        billboards.count++; billboards.positions.push_back(pos); billboards.scales.push_back(scale);
    }
    // src/CloudVolume.cpp:65-82.  The O(N^2) host selection sort is replaced by the device sort inside
    // crn_cone_trace; calling this applies the resulting order (far -> near) to the host arrays, which is
    // what the reference's sort leaves behind.  `point` is implied: the camera of the last coneTrace.
    void sortBoards(vec3 /*point*/) {
        if (!billboards.count) return;
        std::vector<int32_t> order(billboards.count);
        if (crn_read_sorted_order(ctx, order.data()) != CRN_OK) return;       // nothing traced yet
        Billboards b = billboards;
        for (int i = 0; i < billboards.count; i++) { billboards.positions[i] = b.positions[order[i]]; billboards.scales[i] = b.scales[order[i]]; }
    }
    void uploadBillboards() {                            // src/CloudVolume.cpp:139-164 (fluffiness is applied on the device)
        check(ctx, crn_set_billboards(ctx, billboards.count ? &billboards.positions[0].x : nullptr,
                                      billboards.count ? billboards.scales.data() : nullptr, billboards.count, CRN_MEM_HOST));
    }
    void regenerateBillboards(int count, vec3 minOffset, vec3 maxOffset, float minScale, float maxScale) {   // src/CloudVolume.cpp:120-133
        billboards.minOffset = minOffset; billboards.maxOffset = maxOffset; billboards.minScale = minScale; billboards.maxScale = maxScale;
        billboards.positions.clear(); billboards.scales.clear(); billboards.count = 0;
        auto rnd = [](float lo, float hi) { return rand() / (float)RAND_MAX * (hi - lo) + lo; };   // Util::genRandom, src/Util.hpp:31-33
        for (int i = 0; i < count; i++) {
            vec3 p{rnd(minOffset.x, maxOffset.x), rnd(minOffset.y, maxOffset.y), rnd(minOffset.z, maxOffset.z)};
            float s = rnd(minScale, maxScale);
            addCloudBoard(p, s);
        }
    }
    // Device-side variant (§8 f3): same distribution, generated where it is consumed from a counter-based
    // splitmix64 stream; the host arrays are refreshed from the device so caller code can still read them.
    void regenerateBillboardsOnDevice(int count, vec3 minOffset, vec3 maxOffset, float minScale, float maxScale, uint64_t seed) {
        billboards.minOffset = minOffset; billboards.maxOffset = maxOffset; billboards.minScale = minScale; billboards.maxScale = maxScale;
        check(ctx, crn_regenerate_billboards(ctx, count, &minOffset.x, &maxOffset.x, minScale, maxScale, 1.0, seed));
        billboards.positions.resize(count); billboards.scales.resize(count); billboards.count = count;
        if (count) check(ctx, crn_read_billboards(ctx, &billboards.positions[0].x, billboards.scales.data()));
    }
    void resetBillboards() { regenerateBillboards(billboards.count, billboards.minOffset, billboards.maxOffset, billboards.minScale, billboards.maxScale); }

    ivec3 get3DIndices(const int index) const {          // src/CloudVolume.cpp:103-110
        const int line = dimension, slice = dimension * line;
        const int z = index / slice, y = (index - z * slice) / line, x = index - z * slice - y * line;
        return {x, y, z};
    }
    vec3 reverseVoxelIndex(const ivec3 &v) const {       // src/CloudVolume.cpp:112-118
        return {float(v.x) * range.x / dimension + xBounds.x, float(v.y) * range.y / dimension + yBounds.x,
                float(v.z) * range.z / dimension + zBounds.x};
    }
    // what VoxelShader::updateVoxelData reads back (src/Shaders/VoxelShader.cpp:102-133)
    std::vector<uint8_t> readLevel(int level) const {
        const size_t s = std::max(1, dimension >> level);
        std::vector<uint8_t> v(s * s * s);
        check(ctx, crn_read_volume(ctx, level, v.data()));
        return v;
    }
    uint64_t activeVoxels() const { uint64_t n = 0; check(ctx, crn_count_active_voxels(ctx, &n)); return n; }
    // voxelPositions / voxelData of the debug voxel view, compacted on the device: (x, y, z, value) per non-empty voxel
    struct Voxel { float x, y, z, value; };
    std::vector<Voxel> exportVoxels(int channel = 0) const {
        uint64_t n = 0;
        check(ctx, crn_export_voxels(ctx, channel, nullptr, 0, &n));
        std::vector<Voxel> v(n);
        if (n) check(ctx, crn_export_voxels(ctx, channel, &v[0].x, n, &n));
        return v;
    }

    vec3 position;
    vec2 xBounds, yBounds, zBounds;
    int dimension;
    vec3 range{}, voxelSize{};
    int levels;
    Billboards billboards;
    float fluffiness = 1.f;
    int format = CRN_VOLUME_R8;          // CRN_VOLUME_R32F / CRN_VOLUME_RG8 (paper variant) are opt-in extensions
    crn_ctx *ctx = nullptr;                              // replaces volId / instancedQuad VBOs
};

// src/Sun.hpp:12-43 (defaults src/main.cpp:37-43)
class Sun {
public:
    static inline vec3 position{5.f, 20.f, -5.f};
    static inline mat4 P{}, V{};
    static inline vec3 nearPlane{}, farPlane{};
    static inline float clipDistance = 0.f;
    static inline vec3 innerColor{1.f, 1.f, 1.f}, outerColor{1.f, 1.f, 0.f};
    static inline float innerRadius = 1.f, outerRadius = 2.f;

    static crn_sun desc() {
        crn_sun s{};
        s.position[0] = position.x; s.position[1] = position.y; s.position[2] = position.z;
        s.innerColor[0] = innerColor.x; s.innerColor[1] = innerColor.y; s.innerColor[2] = innerColor.z;
        s.outerColor[0] = outerColor.x; s.outerColor[1] = outerColor.y; s.outerColor[2] = outerColor.z;
        s.innerRadius = innerRadius; s.outerRadius = outerRadius;
        return s;
    }
    static void update(CloudVolume *vol) {               // src/Sun.hpp:26-43
        crn_volume_desc d{};
        d.dimension = vol->dimension; d.levels = vol->levels;
        d.position[0] = vol->position.x; d.position[1] = vol->position.y; d.position[2] = vol->position.z;
        d.xBounds[0] = vol->xBounds.x; d.xBounds[1] = vol->xBounds.y; d.yBounds[0] = vol->yBounds.x; d.yBounds[1] = vol->yBounds.y;
        d.zBounds[0] = vol->zBounds.x; d.zBounds[1] = vol->zBounds.y;
        const crn_sun s = desc();
        crn_sun_derived o;
        crn_sun_update(&d, &s, &o);
        for (int i = 0; i < 16; i++) { P.m[i] = o.P[i]; V.m[i] = o.V[i]; }
        nearPlane = {o.nearPlane[0], o.nearPlane[1], o.nearPlane[2]};
        farPlane = {o.farPlane[0], o.farPlane[1], o.farPlane[2]};
        clipDistance = o.clipDistance;
        check(vol->ctx, crn_set_sun(vol->ctx, &s));
    }
};

// src/Shaders/VoxelizeShader.hpp
class VoxelizeShader {
public:
    void voxelize(CloudVolume *volume) {                 // src/Shaders/VoxelizeShader.cpp:18-31
        check(volume->ctx, crn_set_window(volume->ctx, Window::width, Window::height));
        check(volume->ctx, crn_keep_position_map(volume->ctx, keepPositionMap ? 1 : 0));
        check(volume->ctx, crn_voxelize(volume->ctx));
    }
    // the reference's debug views sample `positionMap` (src/main.cpp:172-198)
    std::vector<float> readPositionMap(CloudVolume *volume) const {
        std::vector<float> v((size_t)Window::width * Window::height * 4);
        check(volume->ctx, crn_read_position_map(volume->ctx, v.data()));
        return v;
    }
    bool keepPositionMap = false;
};

// src/Shaders/ConeTraceShader.hpp:15-36 — same public fields, same defaults
class ConeTraceShader {
public:
    explicit ConeTraceShader(unsigned noiseSeed = 1u) { initNoiseMap(32, noiseSeed); }

    void coneTrace(CloudVolume *volume) {                // src/Shaders/ConeTraceShader.cpp:15-82
        if (!doConeTrace && !doNoiseSample && !showQuad) return;
        crn_ctx *c = volume->ctx;
        if (noiseUploadedTo != c) { check(c, crn_set_noise(c, noise.data(), noiseDim)); noiseUploadedTo = c; }   // per context: a second CloudVolume gets its own copy
        crn_camera cam;
        for (int i = 0; i < 16; i++) { cam.P[i] = Camera::getP().m[i]; cam.V[i] = Camera::getV().m[i]; }
        const vec3 p = Camera::getPosition();
        cam.position[0] = p.x; cam.position[1] = p.y; cam.position[2] = p.z;
        check(c, crn_set_camera(c, &cam));
        check(c, crn_set_window(c, Window::width, Window::height));
        crn_trace_params t;
        crn_default_trace_params(&t);
        t.stepSize = stepSize; t.noiseOpacity = noiseOpacity; t.numOctaves = numOctaves; t.freqStep = freqStep; t.persStep = persStep;
        t.adjustSize = adjustSize; t.minNoiseSteps = minNoiseSteps; t.maxNoiseSteps = maxNoiseSteps; t.minNoiseColor = minNoiseColor;
        t.noiseColorScale = noiseColorScale; t.windVel[0] = windVel.x; t.windVel[1] = windVel.y; t.windVel[2] = windVel.z;
        t.vctSteps = vctSteps; t.vctConeAngle = vctConeAngle; t.vctConeInitialHeight = vctConeInitialHeight; t.vctLodOffset = vctLodOffset;
        t.vctDownScaling = vctDownScaling; t.showQuad = showQuad; t.doConeTrace = doConeTrace; t.doNoiseSample = doNoiseSample;
        t.runTime = Window::runTime;
        t.transmittanceCutoff = transmittanceCutoff; t.sampler = sampler; t.drawSun = drawSun;
        check(c, crn_set_trace_params(c, &t));
        framebuffer.resize((size_t)Window::width * Window::height * 4);
        check(c, crn_cone_trace(c, framebuffer.data(), CRN_MEM_HOST, CRN_IMAGE_RGBA8));
        volume->sortBoards(p);                           // the reference sorts the host arrays in this call (:20)
    }

    /* Noise map parameters */
    float stepSize = 0.01f, noiseOpacity = 4.0f;
    int numOctaves = 4;
    float freqStep = 3.f, persStep = 0.5f, adjustSize = 40.f;
    int minNoiseSteps = 2, maxNoiseSteps = 8;
    float minNoiseColor = 0.2f, noiseColorScale = 0.45f;
    vec3 windVel{0.01f, 0.f, 0.f};
    /* Cone trace parameters */
    int vctSteps = 16;
    float vctConeAngle = 0.9f, vctConeInitialHeight = 0.1f, vctLodOffset = 0.f, vctDownScaling = 1.f;
    bool showQuad = false, doConeTrace = true, doNoiseSample = true;
    /* extensions (defaults keep the reference's behaviour) */
    float transmittanceCutoff = 0.f;
    int sampler = CRN_SAMPLER_TEXTURE;
    int drawSun = 1;
    /* the window framebuffer: RGBA8, row 0 at the bottom */
    std::vector<uint8_t> framebuffer;

private:
    void initNoiseMap(int dimension, unsigned seed) {    // src/Shaders/ConeTraceShader.cpp:127-159
        noiseDim = dimension;
        std::vector<int8_t> alpha((size_t)dimension * dimension * dimension);
        srand(seed);
        for (auto &a : alpha) {
            const float v = rand() / (float)RAND_MAX * 256.f - 128.f;      // Util::genRandom(-128.f, 128.f)
            a = (int8_t)std::min(127.f, std::max(-128.f, v));
        }
        noise.resize(alpha.size() * 4);
        crn_build_noise(alpha.data(), dimension, noise.data());
    }
    std::vector<int8_t> noise;
    int noiseDim = 0;
    const crn_ctx *noiseUploadedTo = nullptr;
};

} // namespace crn
