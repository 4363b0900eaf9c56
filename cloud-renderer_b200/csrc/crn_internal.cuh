// crn_internal.cuh — shared declarations of the B200 Cloud-Renderer hot path (host + device).
// Not part of the public boundary (that is include/cloud_renderer_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/cloud_renderer_b200.h"

namespace crn {

// ----------------------------------------------------------------------------------------
// geometry of the two tilings
// ----------------------------------------------------------------------------------------
constexpr int kTile = 16;            // fine tile edge in pixels: one 256-thread CTA, 8 warps of 8x4
#ifndef CRN_COARSE
#define CRN_COARSE 4
#endif
constexpr int kCoarse = CRN_COARSE;  // coarse tile = kCoarse x kCoarse fine tiles (64 px)
constexpr int kMaxConeSteps = 128;   // vctSteps upper bound (reference UI slider tops out far lower)
constexpr int kMaxLevels = 16;
constexpr int kMaxOctaves = 8;
constexpr int kMaxBakedTex = 16;     // baked cone-step textures per frame (k_conebake.cu)
constexpr int kMaxBakedSteps = 32;   // cone steps that may read them (several steps can share one texture)
constexpr int kCodeGroups = 8;       // empty-space groups with a bit in the need-code grid (later groups are always fetched)

// One billboard as a pass sees it, stored in that pass's list order (32 B, two 128-bit loads).
struct BoardRec {
    float cx, cy, cz, r;     // world-space centre (volumePosition + boardPosition), radius (scale * fluffiness)
    float xv, yv, zv;        // V * (centre, 1)
    int32_t idx;             // instance index in the caller's arrays (low 30 bits) | kRecInLattice (camera pass)
};
constexpr int32_t kRecInLattice = 1 << 30;    // the billboard's noise march stays inside the noise-lattice window (k_noiselat.cu)
constexpr int32_t kRecIndexMask = kRecInLattice - 1;

// inclusive pixel rectangle of the quad (i1 < i0: clipped away)
struct BoardRect {
    int16_t i0, i1, j0, j1;
};

// The slice of a camera (light or user) the device code needs.
struct ViewParams {
    float right[3], up[3], back[3];   // rows of V's rotation block
    float nrm[3];                     // normalize(back) == normalize(fragNor)
    float V[16], P[16];
    int32_t ortho;                    // P[15] == 1
    int32_t W, H;
};

struct VolumeParams {                 // uniforms of VoxelizeShader::bindVolume / ConeTraceShader::bindVolume
    float xB[2], yB[2], zB[2];        // absolute bounds (position + relative bounds)
    int32_t dim, levels;
    float stepSize;
    uint32_t levelOff[kMaxLevels];    // byte offset of each level in the chain buffer
    int32_t levelSize[kMaxLevels];
    int32_t z0, z1;                   // slab (voxel slices) this context owns
    int32_t texelBytes;               // 1: R8 (UNORM8, re-quantised mips)   4: R32F (float mips)
};

struct ConeStep {                     // traceCone's per-step constants (identical for every fragment)
    float height;                     // coneHeight at this step, voxels
    float weight;                     // float(i) / (steps * vctDownScaling)
    int32_t level0;                   // lower mip level sampled
    float frac;                       // blend toward level0+1 (0: single level)
    float lod0, lod1;                 // level0 and level0+1 as floats
    float lod;                        // level0 + frac: the tex3DLod operand (mip-linear blend in the texture unit)
};

struct ConeGroup {                    // consecutive cone steps decided by ONE bit of the need code (k_conebake.cu)
    float height;                     // mid height, voxels
    int32_t first, count;             // steps [first, first+count)
    int32_t level;                    // lower mip level its steps sample
    int32_t two;                      // some step also blends in level + 1
};

// A cone step whose (lower level, mip fraction) pair was baked into its own texture for this frame (k_conebake.cu):
// one bilinear fetch on a layered slice-pair texture + a z blend in the kernel replaces the two trilinear fetches of
// textureLod.  The lattice has n nodes per axis over [0,1] (texel centres ON the nodes).
struct BakedStep {
    float height, weight;
    float A, B;                       // x' = s * A + B: normalized volume coordinate -> normalized texture coordinate, A = (n-1)/n, B = 0.5/n
    float zScale;                     // (n-1) * (1 - 2^-20): lattice z of saturate(s), always < n-1 so that floor() <= n-2
    float hA;                         // height * A
    unsigned long long tex;           // cudaTextureObject_t of the step's layered RG16 slice-pair texture
};

// constants of the fast trace variant, derived on the host once per frame
struct FastConst {
    float invP0, invP5;               // 1 / P[0], 1 / P[5]
    float invAdjust, invStep;         // 1 / adjustSize, 1 / stepSize
    float span, minSteps;             // maxNoiseSteps - minNoiseSteps, minNoiseSteps (as floats)
    float nScale[3], nBias[3];        // normalized volume coordinate of a world position: w * nScale + nBias
    float invDim;
};

// The combined-octave noise lattice (k_noiselat.cu): octaves [1, numOctaves) of noise3D pre-summed on the texel lattice of the
// finest octave, inside a world-space window.  A billboard whose noise march stays inside the window (centre +- radius *
// ext[k] within winC[k] +- winH[k]) reads it with one bilinear pass per march step instead of one per octave.
struct NoiseLat {
    int32_t on;
    float K;                          // lattice units per unit of uv (= world / adjustSize): freq_F * noiseDim
    float B[3];                       // x, y: unnormalized texel coordinate = uv * K + B (node i at i + 1/2);  z: node coordinate = uv * K + B
    float scale;                      // stored value * 2 * scale = the octave sum (texels hold (g, a, dg, da) / (2 scale), d = next plane - this one)
    float winC[3], winH[3];           // window centre and half size, world units
    float ext[3];                     // reach of a billboard's march per unit radius: sqrt(1 + 9 viewRay_k^2) (the march overshoots the far hit by up to a chord)
    unsigned long long tex;           // layered RGBA16_SNORM, LINEAR, CLAMP, unnormalized coordinates
};

struct TraceParams {
    crn_trace_params p;
    float lightPos[3];
    float octaveOffsets[4];
    float bg[4];                      // clear colour
    crn_sun sun;
    BoardRect sunRect;
    BoardRec sunRec;
    int32_t nSteps;
    int32_t noiseDim;
    int32_t row0, row1;
    int32_t ilvIndex, ilvCount;       // tile-row interleave (image-space sharding)
    int32_t active;                   // doConeTrace || doNoiseSample || showQuad
    int32_t stats;
    int32_t nGroups;
    float octFreq[kMaxOctaves];       // freq of octave o (freqStep^o)
    float octBias[kMaxOctaves];       // octaveOffsets[o] * freq: uv*freq + bias == (uv + offset)*freq
    float octPers[kMaxOctaves];       // persStep^o
    float octFreqZ[kMaxOctaves];      // freq * noiseDim and bias * noiseDim - 0.5: the z texel coordinate of the layered
    float octBiasZ[kMaxOctaves];      //   noise texture (z filtered in the kernel, see TexSet::noise)
    int32_t noiseMask;                // noiseDim - 1 if noiseDim is a power of two, else -1
    int32_t nFine;                    // steps[0..nFine) are fetched with textureLod (grouped for the empty-space test)
    int32_t nBaked;                   // baked[0..nBaked) are the remaining steps, in step order
    int32_t segCount, segMin, segMax; // tile lists of segCount * segMin .. segMax entries are cut into segCount depth segments (1: off)
    int32_t codeDim;                  // cells per axis of the need-code grid (0: no empty-space skipping)
    float codeDimF;
    FastConst f;
    NoiseLat lat;
    BakedStep baked[kMaxBakedSteps];
    ConeStep steps[kMaxConeSteps];
    ConeGroup groups[kMaxConeSteps];
};

// bins of one pass
struct Bins {
    int32_t tilesX = 0, tilesY = 0, coarseX = 0, coarseY = 0;
    uint32_t *coarseOff = nullptr, *coarseCnt = nullptr;   // per coarse tile
    uint32_t *coarseList = nullptr; size_t coarseCap = 0;
    uint32_t *tileOff = nullptr, *tileCnt = nullptr;       // per fine tile
    uint32_t *tileList = nullptr; size_t tileCap = 0;
    uint32_t *cursors = nullptr;                            // [0] coarse cursor, [1] fine cursor (device)
    size_t tilesAlloc = 0, coarseAlloc = 0;
};

// kernels' launch wrappers (each returns the number of kernels it launched)
int launch_prep_sort(cudaStream_t st, const float *pos, const float *scale, int n, float fluff, const float volpos[3],
                     const ViewParams &light, const float nearPlane[3], float clip, const ViewParams &cam,
                     const float camPos[3], bool doLight, bool doCam, uint32_t *rankL, uint32_t *rankC,
                     uint64_t *keyL, uint64_t *keyC, BoardRec *recTmpL, BoardRec *recTmpC, BoardRect *rectTmpL,
                     BoardRect *rectTmpC, float *lbTmp, BoardRec *recL, BoardRec *recC, BoardRect *rectL,
                     BoardRect *rectC, float *lbSorted, int32_t *drawOrder, void *sortTmp, const NoiseLat *lat = nullptr);
size_t sort_tmp_bytes(int n);      // sortTmp must hold sort_tmp_bytes(n) + 16 bytes
int launch_bin(cudaStream_t st, const BoardRect *rects, const int32_t *bounds, int n, int W, int H, Bins &b);
int launch_tile_order(cudaStream_t st, const Bins &b, uint32_t *order, int ilvIndex = 0, int ilvCount = 1);     // order[owned tiles]: longest lists first (owned: tile row % ilvCount == ilvIndex)
const int32_t *sort_tmp_bounds(const void *sortTmp, int n, int pass);
const uint32_t *sort_tmp_world_box(const void *sortTmp, int n);         // camera pass: sortable bits of the spheres' world bounding box   // rect bounds written by prep_kernel (pass 0 light, 1 camera)
int launch_voxelize(cudaStream_t st, const ViewParams &light, const VolumeParams &vol, const float nearPlane[3],
                    float clip, const BoardRec *recs, const float *lbSorted, const Bins &b, uint32_t *bits,
                    float4 *posmap, uint32_t *bitsA);
// texture-unit view of the chain + noise (CRN_SAMPLER_TEXTURE)
struct TexSet {
    cudaSurfaceObject_t surf[kMaxLevels];   // one per level, written by the mip kernel
    cudaTextureObject_t tex[kMaxLevels];    // one per level: LINEAR, CLAMP, normalized coords, UNORM8 -> float
    cudaTextureObject_t vol;                // the whole mipmapped array: LINEAR in-level and between levels (tex3DLod)
    cudaTextureObject_t volA;               // CRN_VOLUME_RG8: the occupancy channel's chain, same sampling state
    cudaTextureObject_t noise;              // layered 2D RGBA8_SNORM, LINEAR, REPEAT: layer z holds (g_z, a_z, g_z+1, a_z+1)
    cudaTextureObject_t noiseD;             // the same as RGBA16_SNORM differences: (g_z, a_z, g_z+1 - g_z, a_z+1 - a_z) / 2 (fast trace variant)
    cudaTextureObject_t code;               // 3D R8UI, POINT, BORDER: the need-code grid as SKIP bits (bit g set: group g cannot contribute from this cell; outside the volume: 0)
    cudaTextureObject_t baked[kMaxBakedTex]; // layered 2D RG16 UNORM, LINEAR, CLAMP: layer k holds (B(x,y,k), B(x,y,k+1)) of one baked step
    int32_t enabled;
};

// one baked cone-step texture: value(node) = (1-frac) * trilinear(level0) + frac * trilinear(level0+1) at the nodes of
// a lattice that contains the texel centres of both levels (spacing 2^(level0-1) voxels), so that linear interpolation
// between nodes reproduces the mip-linear lookup exactly
struct BakeTex {
    int32_t level0;
    float frac;
    int32_t n;                        // nodes per axis = dim / spacing + 1
    cudaSurfaceObject_t surf;         // layered RG16, n x n x (n-1)
};

int launch_mips(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, uint8_t *chain, uint32_t *ticket,
                bool writeLevel0, const TexSet *ts);
int launch_expand_level0(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, uint8_t *chain, const TexSet *ts);   // chain / ts may be null
int launch_chain_to_surfaces(cudaStream_t st, const VolumeParams &vol, const uint8_t *chain, const TexSet &ts, int firstLevel);
int launch_finish_mips(cudaStream_t st, const VolumeParams &vol, uint8_t *chain, int firstLevel);
int launch_trace(cudaStream_t st, const ViewParams &cam, const VolumeParams &vol, const TraceParams &tp,
                 const BoardRec *recs, const Bins &b, const uint32_t *bits, const uint8_t *chain,
                 const uint32_t *bitsA, const uint8_t *chainA, const int8_t *noise, const TexSet *ts, const uint8_t *needCode, const uint32_t *tileOrder, void *image,
                 int format, unsigned long long *stats, float4 *segPartial, uint32_t *segArrived, int ownedTiles);
int launch_bake_steps(cudaStream_t st, const VolumeParams &vol, const uint32_t *bits, const uint8_t *chain, const BakeTex *tex, int nTex);
int launch_need_code(cudaStream_t st, const VolumeParams &vol, const TraceParams &tp, const uint32_t *bits, const uint32_t *nz, const uint32_t *worldBox,
                     uint8_t *code, cudaSurfaceObject_t codeSurf);
int launch_noise_lattice(cudaStream_t st, const float2 *noise, int dim, const int n[3], const long long base[3], int first, int last,
                         const double *m, const float *pers, float invScale, cudaSurfaceObject_t surf);
size_t skipmask_words(const VolumeParams &vol, uint32_t *off);
int launch_skipmask(cudaStream_t st, const VolumeParams &vol, const uint8_t *chain, uint32_t *nz);
int launch_generate_boards(cudaStream_t st, int n, const float minOff[3], const float maxOff[3], float minScale,
                           float maxScale, double radiusFactor, uint64_t seed, float *pos0, float *pos, float *scale);
size_t export_scratch_words(size_t words);
int launch_export_voxels(cudaStream_t st, const uint32_t *bits, const uint32_t *lit, size_t words, int D, const float pos[3],
                         const float lo[3], const float range[3], uint32_t *blockScratch, unsigned long long *total, float4 *out,
                         unsigned long long capacity);
int launch_rotate_boards(cudaStream_t st, int n, const float *pos0, float *pos, float c, float s);
int launch_count_bits(cudaStream_t st, const uint32_t *bits, size_t words, unsigned long long *out);

} // namespace crn
