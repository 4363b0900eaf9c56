// ref_api.cpp — C entry points around the reference's shader main()s compiled as C++
// (see build_ref.sh).  TEST INFRASTRUCTURE ONLY.  Texture fetches are delegated to the
// fixed-function restatement in liboracle.so (orc_sample_volume / orc_sample_noise): the GL
// texture unit is not part of /root/reference.
#include "glsl_shim.hpp"

#include <cstring>

using namespace glsl;

extern "C" {
void orc_sample_volume(const uint8_t *chain, int dim, int levels, float u, float v, float w, float lod, float out[4]);
void orc_sample_noise(const int8_t *rgba, int dim, float u, float v, float w, float out[4]);
}

#define COMMON_BUILTINS extern vec4 gl_FragCoord; extern float gl_FragDepth; extern vec4 gl_Position; extern bool gl_Discarded; void shader_main();

namespace conetrace_frag {      // res/conetrace_frag.glsl
COMMON_BUILTINS
extern vec3 fragPos, fragNor; extern vec2 fragTex; extern vec3 center; extern float scale;
extern mat4 V; extern vec3 lightPos; extern bool showQuad; extern int voxelDim; extern vec2 xBounds, yBounds, zBounds;
extern sampler3D volumeTexture; extern bool doConeTrace; extern int vctSteps; extern float vctConeAngle, vctConeInitialHeight, vctLodOffset, vctDownScaling;
extern bool doNoise; extern sampler3D noiseMap; extern vec3 octaveOffsets; extern float stepSize, noiseOpacity; extern int numOctaves;
extern float freqStep, persStep, adjustSize; extern int minNoiseSteps, maxNoiseSteps; extern float minNoiseColor, noiseColorScale;
extern vec4 color;
}
namespace first_voxelize {      // res/first_voxelize.glsl
COMMON_BUILTINS
extern vec3 fragPos, fragNor, center; extern float scale; extern vec3 lightNearPlane; extern float clipDistance;
extern int voxelDim; extern vec2 xBounds, yBounds, zBounds; extern float stepSize; extern vec4 color;
}
namespace first_voxelize_paper {   // res/first_voxelize.glsl with lines 54-58 un-commented (build_ref.sh)
COMMON_BUILTINS
extern vec3 fragPos, fragNor, center; extern float scale; extern vec3 lightNearPlane; extern float clipDistance;
extern image3D volume; extern int voxelDim; extern vec2 xBounds, yBounds, zBounds; extern float stepSize; extern vec4 color;
}
namespace second_voxelize {     // res/second_voxelize.glsl
COMMON_BUILTINS
extern image3D volume; extern int voxelDim; extern vec2 xBounds, yBounds, zBounds; extern float stepSize; extern image2D positionMap;
}
namespace sun_frag {            // res/sun_frag.glsl
COMMON_BUILTINS
extern vec3 fragPos, center, innerColor, outerColor; extern float innerRadius, outerRadius; extern vec4 color;
}
namespace billboard_vert_instanced {   // res/billboard_vert_instanced.glsl
COMMON_BUILTINS
extern vec3 vertPos, vertNor, boardPosition; extern float boardScale; extern mat4 P, V, Vi; extern vec3 volumePosition;
extern vec3 fragPos, fragNor; extern vec2 fragTex; extern vec3 center; extern float scale;
}

namespace {
struct VolTex { const uint8_t *chain; int dim, levels; };
struct NoiseTex { const int8_t *rgba; int dim; };
void fetch_volume(void *u, float x, float y, float z, float lod, int has_lod, float out[4]) {
    const VolTex *t = (const VolTex *)u;
    orc_sample_volume(t->chain, t->dim, t->levels, x, y, z, has_lod ? lod : 0.0f, out);
}
void fetch_noise(void *u, float x, float y, float z, float, int, float out[4]) {
    const NoiseTex *t = (const NoiseTex *)u;
    orc_sample_noise(t->rgba, t->dim, x, y, z, out);
}
mat4 load_mat(const float *m) { mat4 r; for (int c = 0; c < 4; c++) for (int i = 0; i < 4; i++) r[c][i] = m[c * 4 + i]; return r; }
vec3 v3(const float *p) { return vec3(p[0], p[1], p[2]); }
} // namespace

extern "C" {

// uniform block of ConeTraceShader::coneTrace (src/Shaders/ConeTraceShader.cpp:26-69) as plain floats/ints
struct ref_conetrace_uniforms {
    float V[16]; float lightPos[3];
    int showQuad; int voxelDim; float xBounds[2], yBounds[2], zBounds[2];
    int doConeTrace; int vctSteps; float vctConeAngle, vctConeInitialHeight, vctLodOffset, vctDownScaling;
    int doNoise; float octaveOffsets[3]; float stepSize, noiseOpacity; int numOctaves; float freqStep, persStep, adjustSize;
    int minNoiseSteps, maxNoiseSteps; float minNoiseColor, noiseColorScale;
};

// one invocation of conetrace_frag.glsl's main(); returns 0 when the fragment was discarded
int ref_conetrace_fragment(const ref_conetrace_uniforms *u, const uint8_t *chain, int levels, const int8_t *noise, int noiseDim,
                           const float fragPos[3], const float fragNor[3], const float fragTex[2], const float center[3], float scale,
                           float color[4]) {
    namespace S = conetrace_frag;
    VolTex vt = {chain, u->voxelDim, levels};
    NoiseTex nt = {noise, noiseDim};
    S::V = load_mat(u->V); S::lightPos = v3(u->lightPos); S::showQuad = u->showQuad != 0; S::voxelDim = u->voxelDim;
    S::xBounds = vec2(u->xBounds[0], u->xBounds[1]); S::yBounds = vec2(u->yBounds[0], u->yBounds[1]); S::zBounds = vec2(u->zBounds[0], u->zBounds[1]);
    S::volumeTexture = sampler3D{&vt, fetch_volume}; S::noiseMap = sampler3D{&nt, fetch_noise};
    S::doConeTrace = u->doConeTrace != 0; S::vctSteps = u->vctSteps; S::vctConeAngle = u->vctConeAngle;
    S::vctConeInitialHeight = u->vctConeInitialHeight; S::vctLodOffset = u->vctLodOffset; S::vctDownScaling = u->vctDownScaling;
    S::doNoise = u->doNoise != 0; S::octaveOffsets = v3(u->octaveOffsets); S::stepSize = u->stepSize; S::noiseOpacity = u->noiseOpacity;
    S::numOctaves = u->numOctaves; S::freqStep = u->freqStep; S::persStep = u->persStep; S::adjustSize = u->adjustSize;
    S::minNoiseSteps = u->minNoiseSteps; S::maxNoiseSteps = u->maxNoiseSteps; S::minNoiseColor = u->minNoiseColor; S::noiseColorScale = u->noiseColorScale;
    S::fragPos = v3(fragPos); S::fragNor = v3(fragNor); S::fragTex = vec2(fragTex[0], fragTex[1]); S::center = v3(center); S::scale = scale;
    S::color = vec4(0.0f); S::gl_Discarded = false;
    S::shader_main();
    color[0] = S::color.x; color[1] = S::color.y; color[2] = S::color.z; color[3] = S::color.w;
    return S::gl_Discarded ? 0 : 1;
}

// first_voxelize.glsl main(): color (= world position, a) and gl_FragDepth; 0 on discard
int ref_first_voxelize_fragment(const float fragPos[3], const float fragNor[3], const float center[3], float scale,
                                const float lightNearPlane[3], float clipDistance, float color[4], float *depth) {
    namespace S = first_voxelize;
    S::fragPos = v3(fragPos); S::fragNor = v3(fragNor); S::center = v3(center); S::scale = scale;
    S::lightNearPlane = v3(lightNearPlane); S::clipDistance = clipDistance;
    S::color = vec4(0.0f); S::gl_FragDepth = 0.0f; S::gl_Discarded = false;
    S::shader_main();
    color[0] = S::color.x; color[1] = S::color.y; color[2] = S::color.z; color[3] = S::color.w;
    *depth = S::gl_FragDepth;
    return S::gl_Discarded ? 0 : 1;
}

// first_voxelize.glsl main() with its interior march switched back on: the image stores it issues, in order (count
// returned, -1 on discard; at most `cap` written), plus the colour / depth outputs
int ref_first_voxelize_paper_fragment(const float fragPos[3], const float fragNor[3], const float center[3], float scale,
                                      const float lightNearPlane[3], float clipDistance, int voxelDim, const float xB[2], const float yB[2],
                                      const float zB[2], float stepSize, int *idx_out, int cap, float color[4], float *depth) {
    namespace S = first_voxelize_paper;
    S::fragPos = v3(fragPos); S::fragNor = v3(fragNor); S::center = v3(center); S::scale = scale;
    S::lightNearPlane = v3(lightNearPlane); S::clipDistance = clipDistance;
    S::volume.n = 0;
    S::voxelDim = voxelDim; S::xBounds = vec2(xB[0], xB[1]); S::yBounds = vec2(yB[0], yB[1]); S::zBounds = vec2(zB[0], zB[1]); S::stepSize = stepSize;
    S::color = vec4(0.0f); S::gl_FragDepth = 0.0f; S::gl_Discarded = false;
    S::shader_main();
    if (S::gl_Discarded) return -1;
    for (int i = 0; i < S::volume.n && i < cap; i++) {
        idx_out[3 * i] = S::volume.idx[i][0]; idx_out[3 * i + 1] = S::volume.idx[i][1]; idx_out[3 * i + 2] = S::volume.idx[i][2];
    }
    color[0] = S::color.x; color[1] = S::color.y; color[2] = S::color.z; color[3] = S::color.w;
    *depth = S::gl_FragDepth;
    return S::volume.n;
}

// second_voxelize.glsl main() on one position-map texel: the image stores it issues (count returned)
int ref_second_voxelize_fragment(const float texel[4], int voxelDim, const float xB[2], const float yB[2], const float zB[2], float stepSize,
                                 int idx_out[9 * 3], float val_out[9]) {
    namespace S = second_voxelize;
    S::volume.n = 0;
    S::voxelDim = voxelDim; S::xBounds = vec2(xB[0], xB[1]); S::yBounds = vec2(yB[0], yB[1]); S::zBounds = vec2(zB[0], zB[1]); S::stepSize = stepSize;
    S::positionMap = image2D{texel, 1, 1};
    S::gl_FragCoord = vec4(0.5f, 0.5f, 0.0f, 1.0f);
    S::gl_Discarded = false;
    S::shader_main();
    for (int i = 0; i < S::volume.n && i < 9; i++) {
        idx_out[3 * i] = S::volume.idx[i][0]; idx_out[3 * i + 1] = S::volume.idx[i][1]; idx_out[3 * i + 2] = S::volume.idx[i][2];
        val_out[i] = S::volume.val[i][0];
    }
    return S::volume.n;
}

// sun_frag.glsl main(); 0 on discard
int ref_sun_fragment(const float fragPos[3], const float center[3], const float innerColor[3], const float outerColor[3],
                     float innerRadius, float outerRadius, float color[4]) {
    namespace S = sun_frag;
    S::fragPos = v3(fragPos); S::center = v3(center); S::innerColor = v3(innerColor); S::outerColor = v3(outerColor);
    S::innerRadius = innerRadius; S::outerRadius = outerRadius; S::color = vec4(0.0f); S::gl_Discarded = false;
    S::shader_main();
    color[0] = S::color.x; color[1] = S::color.y; color[2] = S::color.z; color[3] = S::color.w;
    return S::gl_Discarded ? 0 : 1;
}

// billboard_vert_instanced.glsl main() for one vertex of one instance
void ref_billboard_vertex(const float P[16], const float V[16], const float Vi[16], const float volumePosition[3], const float vertPos[3],
                          const float vertNor[3], const float boardPosition[3], float boardScale, float gl_Position[4], float fragPos[3],
                          float fragNor[3], float fragTex[2], float center[3], float *scale) {
    namespace S = billboard_vert_instanced;
    S::P = load_mat(P); S::V = load_mat(V); S::Vi = load_mat(Vi); S::volumePosition = v3(volumePosition);
    S::vertPos = v3(vertPos); S::vertNor = v3(vertNor); S::boardPosition = v3(boardPosition); S::boardScale = boardScale;
    S::shader_main();
    for (int i = 0; i < 4; i++) gl_Position[i] = S::gl_Position[i];
    fragPos[0] = S::fragPos.x; fragPos[1] = S::fragPos.y; fragPos[2] = S::fragPos.z;
    fragNor[0] = S::fragNor.x; fragNor[1] = S::fragNor.y; fragNor[2] = S::fragNor.z;
    fragTex[0] = S::fragTex.x; fragTex[1] = S::fragTex.y;
    center[0] = S::center.x; center[1] = S::center.y; center[2] = S::center.z;
    *scale = S::scale;
}

} // extern "C"
