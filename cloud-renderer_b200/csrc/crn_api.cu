// crn_api.cu — the C-ABI (include/cloud_renderer_b200.h): context, host-side uniform
// derivation (what the reference's pass drivers do on the CPU before each draw), buffer
// management and the stream-ordered orchestration of the kernels in k_*.cu.
//
// Reference call sites replaced (src/main.cpp:98-124):
//   Sun::update(volume)                -> crn_set_sun + derive_sun()          (src/Sun.hpp:26-43)
//   volume->update()                   -> crn_set_volume/crn_set_billboards   (src/CloudVolume.cpp:84-93,139-164)
//   voxelizeShader->voxelize(volume)   -> crn_voxelize                        (src/Shaders/VoxelizeShader.cpp:18-31)
//   coneShader->coneTrace(volume)      -> crn_cone_trace                      (src/Shaders/ConeTraceShader.cpp:15-82)
// There is no CPU fallback: every entry point that computes needs a CUDA device and fails
// with CRN_ERR_NO_DEVICE / CRN_ERR_CUDA otherwise.
#include "crn_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

using namespace crn;

// ------------------------------------------------------------------------------------------
// host math: the GLM 0.9.8.5 calls the reference's drivers make (RH, -1..1 depth), float32,
// one rounding per operation, sums left to right.
// ------------------------------------------------------------------------------------------
namespace {

struct F3 { float v[3]; };
inline F3 f3(const float *p) { return {{p[0], p[1], p[2]}}; }
inline F3 sub3(F3 a, F3 b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
inline F3 add3(F3 a, F3 b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
inline F3 scale3(F3 a, float s) { return {{a.v[0] * s, a.v[1] * s, a.v[2] * s}}; }
inline float dot3(F3 a, F3 b) { float t0 = a.v[0] * b.v[0], t1 = a.v[1] * b.v[1], t2 = a.v[2] * b.v[2]; return (t0 + t1) + t2; }
inline float len3(F3 a) { return sqrtf(dot3(a, a)); }
inline F3 norm3(F3 a) { return scale3(a, 1.0f / sqrtf(dot3(a, a))); }
inline F3 cross3(F3 x, F3 y) {
    return {{x.v[1] * y.v[2] - y.v[1] * x.v[2], x.v[2] * y.v[0] - y.v[2] * x.v[0], x.v[0] * y.v[1] - y.v[0] * x.v[1]}};
}

void look_at(F3 eye, F3 center, F3 up, float *m) {
    const F3 f = norm3(sub3(center, eye));
    const F3 s = norm3(cross3(f, up));
    const F3 u = cross3(s, f);
    for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    for (int c = 0; c < 3; c++) { m[c * 4 + 0] = s.v[c]; m[c * 4 + 1] = u.v[c]; m[c * 4 + 2] = -f.v[c]; }
    m[12] = -dot3(s, eye); m[13] = -dot3(u, eye); m[14] = dot3(f, eye);
}

void ortho(float l, float r, float b, float t, float n, float f, float *m) {
    for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    m[0] = 2.0f / (r - l); m[5] = 2.0f / (t - b); m[10] = -2.0f / (f - n);
    m[12] = -(r + l) / (r - l); m[13] = -(t + b) / (t - b); m[14] = -(f + n) / (f - n);
}

void perspective(float fovy, float aspect, float n, float f, float *m) {
    const float th = tanf(fovy / 2.0f);
    for (int i = 0; i < 16; i++) m[i] = 0.0f;
    m[0] = 1.0f / (aspect * th); m[5] = 1.0f / th; m[11] = -1.0f;
    m[10] = -(f + n) / (f - n); m[14] = -(2.0f * f * n) / (f - n);
}

void derive_sun(const crn_volume_desc &vol, const crn_sun &sun, crn_sun_derived *out) {     // src/Sun.hpp:26-43
    const F3 mn = {{vol.xBounds[0], vol.yBounds[0], vol.zBounds[0]}}, mx = {{vol.xBounds[1], vol.yBounds[1], vol.zBounds[1]}};
    const F3 vp = f3(vol.position);
    const F3 lookDir = norm3(sub3(vp, f3(sun.position)));
    const float L = fmaxf(len3(mn), len3(mx));
    const F3 lookPos = sub3(vp, scale3(lookDir, L));
    const F3 upv = {{0.0f, 1.0f, 0.0f}};
    look_at(lookPos, vp, upv, out->V);
    const float minmin = 2.0f * fminf(mn.v[0], fminf(mn.v[1], mn.v[2]));
    const float maxmax = 2.0f * fmaxf(mx.v[0], fmaxf(mx.v[1], mx.v[2]));
    const F3 nearP = add3(lookPos, scale3(lookDir, 0.01f));
    const F3 farP = add3(lookPos, scale3(scale3(lookDir, 2.0f), L));
    for (int k = 0; k < 3; k++) { out->nearPlane[k] = nearP.v[k]; out->farPlane[k] = farP.v[k]; }
    out->clipDistance = len3(sub3(farP, nearP));
    ortho(minmin, maxmax, minmin, maxmax, 0.01f, 0.01f + out->clipDistance, out->P);
}

ViewParams make_view(const float *P, const float *V, int W, int H) {
    ViewParams v;
    for (int k = 0; k < 3; k++) { v.right[k] = V[4 * k + 0]; v.up[k] = V[4 * k + 1]; v.back[k] = V[4 * k + 2]; }
    const F3 n = norm3(f3(v.back));
    for (int k = 0; k < 3; k++) v.nrm[k] = n.v[k];
    std::memcpy(v.V, V, 64); std::memcpy(v.P, P, 64);
    v.ortho = (P[15] == 1.0f) ? 1 : 0;
    v.W = W; v.H = H;
    return v;
}

// host twin of k_prep_sort.cu's quad_rect, for the one sun quad
void host_quad(const ViewParams &vp, const float c[3], float scale, BoardRec *rec, BoardRect *q) {
    float cv[4];
    for (int r = 0; r < 4; r++) cv[r] = ((vp.V[r] * c[0] + vp.V[4 + r] * c[1]) + vp.V[8 + r] * c[2]) + vp.V[12 + r];
    *rec = {c[0], c[1], c[2], scale, cv[0], cv[1], cv[2], -1};
    const float zv = cv[2];
    const float zc = vp.P[10] * zv + vp.P[14] * 1.0f, wc = vp.P[11] * zv + vp.P[15] * 1.0f;
    if (!(zc >= -wc && zc <= wc) || !(wc > 0.0f)) { *q = {0, -1, 0, -1}; return; }
    const float W = (float)vp.W, H = (float)vp.H;
    float x0 = (vp.P[0] * (cv[0] - scale) + vp.P[12] * 1.0f) / wc, x1 = (vp.P[0] * (cv[0] + scale) + vp.P[12] * 1.0f) / wc;
    float y0 = (vp.P[5] * (cv[1] - scale) + vp.P[13] * 1.0f) / wc, y1 = (vp.P[5] * (cv[1] + scale) + vp.P[13] * 1.0f) / wc;
    float fx0 = (x0 + 1.0f) * 0.5f * W, fx1 = (x1 + 1.0f) * 0.5f * W, fy0 = (y0 + 1.0f) * 0.5f * H, fy1 = (y1 + 1.0f) * 0.5f * H;
    fx0 = fminf(fmaxf(fx0, -2.0f), W + 2.0f); fx1 = fminf(fmaxf(fx1, -2.0f), W + 2.0f);
    fy0 = fminf(fmaxf(fy0, -2.0f), H + 2.0f); fy1 = fminf(fmaxf(fy1, -2.0f), H + 2.0f);
    q->i0 = (int16_t)std::max(0, (int)floorf(fx0) - 1); q->i1 = (int16_t)std::min(vp.W - 1, (int)floorf(fx1) + 1);
    q->j0 = (int16_t)std::max(0, (int)floorf(fy0) - 1); q->j1 = (int16_t)std::min(vp.H - 1, (int)floorf(fy1) + 1);
}

std::mutex g_err_mutex;
std::string g_create_error;

} // namespace

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct crn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    std::string err;
    uint64_t launches = 0;

    bool haveVol = false, haveSun = false, haveCam = false, haveNoise = false, haveWindow = false;
    crn_volume_desc vol{};
    crn_sun sun{};
    crn_camera cam{};
    crn_trace_params tp{};
    int W = 0, H = 0;
    int nBoards = 0;
    int noiseDim = 0;
    int row0 = 0, row1 = 1 << 30;
    int ilvIndex = 0, ilvCount = 1;
    int z0 = 0, z1 = -1;                 // -1: whole volume
    bool keepPosmap = false, statsOn = false, timingOn = false;
    int imageFormat = CRN_IMAGE_RGBA8;   // format of the frame in `image`
    size_t slabPoolCap = 0;              // light-pass pool capacity that has been validated for slab voxelizes
    bool wantLinear0 = false;            // the caller took crn_volume_level_ptr(0): keep the linear level 0 current
    bool linear0Valid = false, linear0ValidA = false;   // chain level 0 holds the expansion of the current bits (it is produced on demand)
    bool noBake = false;                 // CRN_NO_BAKE=1: every cone step through textureLod (A/B and tests)
    bool voxelized = false, traced = false;

    VolumeParams vparams{};
    size_t chainBytes = 0;

    bool havePos0 = false;               // pos0 holds the un-advected offsets of the current billboard set
    DevBuf exportTmp, exportOut;         // crn_export_voxels scratch
    DevBuf bitsA, chainA;                // CRN_VOLUME_RG8: the occupancy (alpha) channel's level-0 set and R8 chain
    DevBuf pos0, pos, scale, keyL, keyC, rankL, rankC, recTmpL, recTmpC, rectTmpL, rectTmpC, lbTmp, recL, rectL,
        lbSorted, drawOrder, bits, chain, noise, posmap, image, misc, maskNz, sortTmp;
    bool maskCurrent = false;
    size_t poolMin = (size_t)1 << 20;    // initial bin-pool entries (CRN_BIN_POOL_MIN overrides: tests force the growth path)
    Bins binsL;
    uint32_t *hCursors = nullptr;        // pinned: [0..2] light cursors + flags, [4..6] camera cursors + flags
    unsigned long long *hStats = nullptr;

    // Everything a trace READS and a voxelize / trace set-up WRITES exists twice, so that the set-up of frame k+1 (light
    // side, mips, masks, baked steps, need codes on the light stream; camera-side prep / sort / bin on the side stream)
    // overlaps the trace of frame k on the caller's stream.  The chain, the masks and the bits are read only by that
    // set-up itself (texture sampler) and stay single.
    struct BakeKey { uint64_t gen; int nTex; int level0[kMaxBakedTex]; float frac[kMaxBakedTex]; int n[kMaxBakedTex]; };
    struct CodeKey { uint64_t gen; int G, nGroups; float height[kCodeGroups]; int level[kCodeGroups], first[kCodeGroups], count[kCodeGroups]; float light[3]; float bounds[6]; };
    struct VolSet {                      // texture-unit copies (CRN_SAMPLER_TEXTURE) + per-frame cone acceleration data (k_conebake.cu)
        cudaMipmappedArray_t volArray = nullptr, volArrayA = nullptr;
        TexSet ts{}, tsA{};              // tsA: CRN_VOLUME_RG8, texture-unit copy of the occupancy chain
        int volArrayDim = 0, volArrayLevels = 0, volArrayFormat = -1;
        bool texCurrent = false;         // the arrays hold the chain of the last voxelize
        cudaArray_t bakedArr[kMaxBakedTex] = {};
        cudaSurfaceObject_t bakedSurf[kMaxBakedTex] = {};
        int bakedN[kMaxBakedTex] = {};
        DevBuf needCode;
        cudaArray_t codeArr = nullptr;   // the need codes as a 3D R8UI texture of skip bits (fast trace variant)
        cudaSurfaceObject_t codeSurf = 0;
        int codeG = 0;
        BakeKey bakeKey{};
        CodeKey codeKey{};
        bool bakeValid = false, codeValid = false;
        uint64_t codeBoardsGen = ~0ull;
        cudaEvent_t evFree = nullptr;    // recorded after the last trace that read this set
        bool freeValid = false;
    } vset[2];
    int vs = 0;                          // the set the last voxelize wrote = the one traces read
    struct CamSet {                      // camera-side records, bins and tile order of one trace
        DevBuf recC, rectC, tileOrder, sortTmpC;
        Bins binsC;
        cudaEvent_t evFree = nullptr;
        bool freeValid = false;
    } cset[2];
    int cs = 0;
    cudaArray_t noiseArray = nullptr, noiseArrayD = nullptr;
    cudaTextureObject_t noiseTex = 0, noiseTexD = 0;
    // combined-octave noise lattice (k_noiselat.cu): baked when the noise texture, the octave parameters or the window change
    struct LatKey { uint64_t noiseGen; int dim, octaves; float freqStep, persStep, adjust; int n[3]; long long base[3]; };
    cudaArray_t latArray = nullptr;
    cudaSurfaceObject_t latSurf = 0;
    cudaTextureObject_t latTex = 0;
    int latN[3] = {0, 0, 0};
    LatKey latKey{};
    bool latValid = false;
    bool noLattice = false;              // CRN_NO_LATTICE=1: every octave through its own lookup (A/B and tests)
    uint64_t noiseGen = 0;
    // host-known bounds over the billboards: max(|(x,z)|, |y|) of the offsets (unchanged by crn_animate_billboards' rotation
    // about Y, so an animated set keeps its window) and the largest scale (< 0: unknown, device-resident source)
    float boardPosMax = -1.0f, boardScaleMax = 0.0f;
    BakeTex bakePlan[kMaxBakedTex] = {};
    int nBakePlan = 0;
    DevBuf segPartial, segArrived;       // small frames: per-segment partial composites of the cut tile lists (k_trace.cu)
    int segOverride = -1;                // CRN_TRACE_SEGMENTS=n: force the segment count (experiments)
    int segMaxOverride = -1;             // CRN_TRACE_SEGMAX=n: longest tile list that is still cut (experiments)
    uint64_t volumeGen = 0;              // bumped whenever the chain changes (voxelize, finish_mips)
    uint64_t boardsGen = 0;              // bumped whenever the billboard arrays change
    bool lastTraceExplicit = false;      // the last trace read the bits / the chain directly (explicit sampler)
    cudaEvent_t evPrepared = nullptr, evVoxDone = nullptr, evAccT[2] = {};

    // pipelined read-back (crn_cone_trace_async): second image buffer, copy stream, frame/copy events
    DevBuf image2;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evFrame[2] = {}, evCopy[2] = {};
    cudaEvent_t evBin[2] = {};           // bin cursors ready (light, camera): their read-back rides the copy stream
    // the camera-side set-up (prep, sort, bin, tile order) depends only on the billboards and the camera, so it runs on
    // a side stream and overlaps the voxelize + mip kernels of the same frame, which leave most SMs idle
    cudaStream_t auxStream = nullptr;
    cudaEvent_t evBoards = nullptr, evAuxDone = nullptr, evTraceEnd = nullptr, evAuxT[3] = {};
    bool traceEndValid = false;
    bool copyPending[2] = {false, false};
    int imgSel = 0;

    // the light-side set-up of frame k+1 (billboard upload, prep, sort, bin, voxelize kernel) touches nothing the trace of
    // frame k reads (texture sampler: texture arrays + masks; the occupancy bits are consumed by the mip/mask kernels
    // before the trace starts), so it runs on its own stream and overlaps that trace; the mip + mask kernels, which
    // overwrite what the trace samples, stay on the caller's stream behind it
    cudaStream_t lightStream = nullptr;
    cudaEvent_t evLightDone = nullptr, evBitsFree = nullptr, evMainMark = nullptr;
    bool bitsFreeValid = false, auxDoneValid = false;
    cudaEvent_t evCur[2] = {};           // cursor read-back done (light, camera): the next bin pass resets the cursors
    bool curValid[2] = {false, false};
    bool bitsExposed = false;            // crn_volume_bits_ptr handed the set to the caller: keep strict stream order
    bool levelExposed = false;           // ... or crn_volume_level_ptr did

    cudaEvent_t evV[6] = {}, evT[4] = {};
    bool evVValid = false, evTValid = false;
};

namespace {

inline crn_ctx::VolSet &VS(crn_ctx *c) { return c->vset[c->vs]; }
inline crn_ctx::CamSet &CS(crn_ctx *c) { return c->cset[c->cs]; }

int fail(crn_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else { std::lock_guard<std::mutex> g(g_err_mutex); g_create_error = buf; }
    return code;
}

#define CRN_CUDA(c, call)                                                                              \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail((c), CRN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int reserve(crn_ctx *c, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return CRN_OK;
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copyStream) CRN_CUDA(c, cudaStreamSynchronize(c->copyStream));
    if (c->auxStream) CRN_CUDA(c, cudaStreamSynchronize(c->auxStream));
    if (c->lightStream) CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
    if (b.p) CRN_CUDA(c, cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    CRN_CUDA(c, cudaMalloc(&b.p, want));
    CRN_CUDA(c, cudaMemsetAsync(b.p, 0, want, c->stream));      // tickets / counters start at zero
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));              // the first user may be another of the context's streams
    b.cap = want;
    return CRN_OK;
}

int alloc_u32(crn_ctx *c, uint32_t *&p, size_t &have, size_t want) {
    if (want <= have) return CRN_OK;
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->auxStream) CRN_CUDA(c, cudaStreamSynchronize(c->auxStream));
    if (c->lightStream) CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
    if (p) CRN_CUDA(c, cudaFree(p));
    p = nullptr; have = 0;
    CRN_CUDA(c, cudaMalloc(&p, want * sizeof(uint32_t)));
    have = want;
    return CRN_OK;
}

int ensure_bins(crn_ctx *c, Bins &b, int W, int H, int n) {
    b.tilesX = (W + kTile - 1) / kTile; b.tilesY = (H + kTile - 1) / kTile;
    b.coarseX = (b.tilesX + kCoarse - 1) / kCoarse; b.coarseY = (b.tilesY + kCoarse - 1) / kCoarse;
    const size_t tiles = (size_t)b.tilesX * b.tilesY, coarse = (size_t)b.coarseX * b.coarseY;
    if (tiles > b.tilesAlloc) {
        size_t h1 = b.tilesAlloc, h2 = b.tilesAlloc;
        int r = alloc_u32(c, b.tileOff, h1, tiles); if (r) return r;
        r = alloc_u32(c, b.tileCnt, h2, tiles); if (r) return r;
        b.tilesAlloc = tiles;
    }
    (void)coarse;
    if (!b.cursors) { CRN_CUDA(c, cudaMalloc(&b.cursors, 4 * sizeof(uint32_t))); CRN_CUDA(c, cudaMemsetAsync(b.cursors, 0, 4 * sizeof(uint32_t), c->stream)); CRN_CUDA(c, cudaStreamSynchronize(c->stream)); }
    const bool forced = c->poolMin != ((size_t)1 << 20);
    int r = alloc_u32(c, b.coarseList, b.coarseCap, 4 * (forced ? c->poolMin : std::max<size_t>((size_t)1 << 16, (size_t)n * 8))); if (r) return r;   // uint4 entries
    // first guess: ~100 tiles per billboard, or 32 entries per tile, whichever is larger; grown on demand
    r = alloc_u32(c, b.tileList, b.tileCap, forced ? c->poolMin : std::max<size_t>(std::max<size_t>(c->poolMin, (size_t)n * 96), tiles * 32)); if (r) return r;
    return CRN_OK;
}

void free_bins(Bins &b) {
    cudaFree(b.coarseOff); cudaFree(b.coarseCnt); cudaFree(b.coarseList);
    cudaFree(b.tileOff); cudaFree(b.tileCnt); cudaFree(b.tileList); cudaFree(b.cursors);
    b = Bins();
}

// grow a pool whose cursor ran past its capacity; returns true if anything grew
int grow_if_overflowed(crn_ctx *c, Bins &b, const uint32_t *cur, bool *grew) {
    *grew = false;
    if ((size_t)cur[0] * 4 > b.coarseCap) {                 // cursor counts uint4 entries
        int r = alloc_u32(c, b.coarseList, b.coarseCap, 4 * ((size_t)cur[0] + cur[0] / 2)); if (r) return r;
        *grew = true;
    }
    if (cur[1] > b.tileCap) {
        int r = alloc_u32(c, b.tileList, b.tileCap, (size_t)cur[1] + cur[1] / 2); if (r) return r;
        *grew = true;
    }
    return CRN_OK;
}

void sync_all(crn_ctx *c) {
    cudaStreamSynchronize(c->stream);
    if (c->lightStream) cudaStreamSynchronize(c->lightStream);
    if (c->auxStream) cudaStreamSynchronize(c->auxStream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
}

void free_chain_textures(cudaMipmappedArray_t &arr, TexSet &ts) {
    for (int l = 0; l < kMaxLevels; l++) {
        if (ts.tex[l]) cudaDestroyTextureObject(ts.tex[l]);
        if (ts.surf[l]) cudaDestroySurfaceObject(ts.surf[l]);
        ts.tex[l] = 0; ts.surf[l] = 0;
    }
    if (ts.vol) cudaDestroyTextureObject(ts.vol);
    ts.vol = 0;
    if (arr) cudaFreeMipmappedArray(arr);
    arr = nullptr;
}

void free_baked(crn_ctx *c, int i) {
    if (VS(c).ts.baked[i]) cudaDestroyTextureObject(VS(c).ts.baked[i]);
    if (VS(c).bakedSurf[i]) cudaDestroySurfaceObject(VS(c).bakedSurf[i]);
    if (VS(c).bakedArr[i]) cudaFreeArray(VS(c).bakedArr[i]);
    VS(c).ts.baked[i] = 0; VS(c).bakedSurf[i] = 0; VS(c).bakedArr[i] = nullptr; VS(c).bakedN[i] = 0;
}

void free_code_texture(crn_ctx::VolSet &v) {
    if (v.ts.code) cudaDestroyTextureObject(v.ts.code);
    if (v.codeSurf) cudaDestroySurfaceObject(v.codeSurf);
    if (v.codeArr) cudaFreeArray(v.codeArr);
    v.ts.code = 0; v.codeSurf = 0; v.codeArr = nullptr; v.codeG = 0;
}

void free_lattice(crn_ctx *c) {
    if (c->latTex) cudaDestroyTextureObject(c->latTex);
    if (c->latSurf) cudaDestroySurfaceObject(c->latSurf);
    if (c->latArray) cudaFreeArray(c->latArray);
    c->latTex = 0; c->latSurf = 0; c->latArray = nullptr; c->latValid = false;
    c->latN[0] = c->latN[1] = c->latN[2] = 0;
}

void free_vol_textures(crn_ctx *c) {
    free_chain_textures(VS(c).volArray, VS(c).ts);
    free_chain_textures(VS(c).volArrayA, VS(c).tsA);
    VS(c).ts.volA = 0; VS(c).tsA.enabled = 0;
    VS(c).volArrayDim = VS(c).volArrayLevels = 0; VS(c).volArrayFormat = -1; VS(c).ts.enabled = 0; VS(c).texCurrent = false;
}

// the R8 immutable 3D texture with `levels` mips of the reference (src/CloudVolume.cpp:18-23):
// LINEAR within a level, CLAMP_TO_EDGE x3; the mip-linear blend is done in the kernel.
int create_chain_textures(crn_ctx *c, cudaMipmappedArray_t &arr, TexSet &ts);

int ensure_vol_textures(crn_ctx *c) {
    const int D = c->vol.dimension, L = c->vol.levels;
    if (VS(c).volArray && VS(c).volArrayDim == D && VS(c).volArrayLevels == L && VS(c).volArrayFormat == c->vol.format) return CRN_OK;
    sync_all(c);
    free_vol_textures(c);
    int r;
    if ((r = create_chain_textures(c, VS(c).volArray, VS(c).ts))) return r;
    if (c->vol.format == CRN_VOLUME_RG8) {
        if ((r = create_chain_textures(c, VS(c).volArrayA, VS(c).tsA))) return r;
        VS(c).ts.volA = VS(c).tsA.vol; VS(c).tsA.enabled = 1;
    }
    VS(c).volArrayDim = D; VS(c).volArrayLevels = L; VS(c).volArrayFormat = c->vol.format; VS(c).ts.enabled = 1; VS(c).texCurrent = false;
    return CRN_OK;
}

int create_chain_textures(crn_ctx *c, cudaMipmappedArray_t &arr, TexSet &ts) {
    const int D = c->vol.dimension, L = c->vol.levels;
    const bool f32 = c->vol.format == CRN_VOLUME_R32F;
    cudaChannelFormatDesc cd = f32 ? cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat)
                                   : cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
    const cudaTextureReadMode readMode = f32 ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
    CRN_CUDA(c, cudaMallocMipmappedArray(&arr, &cd, make_cudaExtent(D, D, D), L, cudaArraySurfaceLoadStore));
    for (int l = 0; l < L; l++) {
        cudaArray_t lvl = nullptr;
        CRN_CUDA(c, cudaGetMipmappedArrayLevel(&lvl, arr, l));
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray; rd.res.array.array = lvl;
        CRN_CUDA(c, cudaCreateSurfaceObject(&ts.surf[l], &rd));
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear; td.readMode = readMode; td.normalizedCoords = 1;
        CRN_CUDA(c, cudaCreateTextureObject(&ts.tex[l], &rd, &td, nullptr));
    }
    {   // one object over the whole chain for tex3DLod: LINEAR inside a level and between levels (LINEAR_MIPMAP_LINEAR)
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeMipmappedArray; rd.res.mipmap.mipmap = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear; td.mipmapFilterMode = cudaFilterModeLinear;
        td.readMode = readMode; td.normalizedCoords = 1;
        td.minMipmapLevelClamp = 0.0f; td.maxMipmapLevelClamp = (float)(L - 1);
        CRN_CUDA(c, cudaCreateTextureObject(&ts.vol, &rd, &td, nullptr));
    }
    return CRN_OK;
}

int check_volume(crn_ctx *c, const crn_volume_desc *d) {
    const int D = d->dimension;
    // 1024^3 R8 + mips = 1.23 GB: level offsets are 32-bit
    if (D < 32 || D > 1024 || (D & (D - 1))) return fail(c, CRN_ERR_UNSUPPORTED, "dimension %d: need a power of two in [32, 1024]", D);
    int maxL = 1; for (int s = D; s > 1; s >>= 1) maxL++;
    if (d->levels < 1 || d->levels > maxL || d->levels > kMaxLevels) return fail(c, CRN_ERR_INVALID_ARG, "levels %d out of range [1,%d]", d->levels, maxL);
    if (!(d->xBounds[1] > d->xBounds[0] && d->yBounds[1] > d->yBounds[0] && d->zBounds[1] > d->zBounds[0]))
        return fail(c, CRN_ERR_INVALID_ARG, "bounds must satisfy min < max on every axis");
    if (d->format != CRN_VOLUME_R8 && d->format != CRN_VOLUME_R32F && d->format != CRN_VOLUME_RG8) return fail(c, CRN_ERR_UNSUPPORTED, "unknown volume format %d", d->format);
    if (d->format == CRN_VOLUME_R32F && D > 512) return fail(c, CRN_ERR_UNSUPPORTED, "R32F volumes are limited to 512^3 (32-bit level offsets)");
    return CRN_OK;
}

void fill_vparams(crn_ctx *c) {
    VolumeParams &v = c->vparams;
    const crn_volume_desc &d = c->vol;
    for (int k = 0; k < 2; k++) {
        v.xB[k] = d.position[0] + d.xBounds[k]; v.yB[k] = d.position[1] + d.yBounds[k]; v.zB[k] = d.position[2] + d.zBounds[k];
    }
    v.dim = d.dimension; v.levels = d.levels;
    const float fd = (float)d.dimension;
    const float rx = d.xBounds[1] - d.xBounds[0], ry = d.yBounds[1] - d.yBounds[0], rz = d.zBounds[1] - d.zBounds[0];
    v.stepSize = fminf(rx / fd, fminf(ry / fd, rz / fd));          // src/Shaders/VoxelizeShader.cpp:128
    v.texelBytes = d.format == CRN_VOLUME_R32F ? 4 : 1;
    size_t off = 0; int s = d.dimension;
    for (int l = 0; l < kMaxLevels; l++) { v.levelOff[l] = 0; v.levelSize[l] = 0; }
    for (int l = 0; l < d.levels; l++) {
        v.levelOff[l] = (uint32_t)off; v.levelSize[l] = s;
        off += ((size_t)s * s * s * v.texelBytes + 255) / 256 * 256;
        s = std::max(1, s / 2);
    }
    c->chainBytes = off;
    v.z0 = c->z1 < 0 ? 0 : c->z0;
    v.z1 = c->z1 < 0 ? d.dimension : c->z1;
}

// Billboard arrays are written on the light stream, so the upload of frame k+1 does not queue behind the trace of
// frame k.  It has to follow (a) the camera-side set-up of the last trace, which reads the arrays on the side stream,
// (b) the light-side set-up of the last voxelize (same stream: in order) and (c), for a device-resident source,
// whatever the caller queued on their stream to produce it.
cudaStream_t upload_stream(crn_ctx *c, bool sourceOnCallerStream) {
    cudaStream_t up = c->lightStream;
    if (c->auxDoneValid) cudaStreamWaitEvent(up, c->evAuxDone, 0);
    if (sourceOnCallerStream) {
        cudaEventRecord(c->evMainMark, c->stream);
        cudaStreamWaitEvent(up, c->evMainMark, 0);
    }
    return up;
}

int require(crn_ctx *c, bool ok, const char *what) {
    return ok ? CRN_OK : fail(c, CRN_ERR_STATE, "%s has not been set", what);
}

// non-zero bits of the levels >= 1 (level 0 is the occupancy set itself): what the need codes are tested against
int build_masks(crn_ctx *c) {
    const size_t words = skipmask_words(c->vparams, nullptr);
    int r;
    if ((r = reserve(c, c->maskNz, words * 4))) return r;
    c->launches += launch_skipmask(c->lightStream, c->vparams, (const uint8_t *)c->chain.p, (uint32_t *)c->maskNz.p);
    c->maskCurrent = true;
    return CRN_OK;
}

// layered RG16 array of one baked cone step: n x n texels, n-1 layers (layer k = node planes k and k+1)
int ensure_baked_array(crn_ctx *c, int i, int n) {
    if (VS(c).bakedArr[i] && VS(c).bakedN[i] == n) return CRN_OK;
    sync_all(c);
    free_baked(c, i);
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(16, 16, 0, 0, cudaChannelFormatKindSigned);     // (value, step to the next plane)
    CRN_CUDA(c, cudaMalloc3DArray(&VS(c).bakedArr[i], &cd, make_cudaExtent(n, n, n - 1), cudaArrayLayered | cudaArraySurfaceLoadStore));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = VS(c).bakedArr[i];
    CRN_CUDA(c, cudaCreateSurfaceObject(&VS(c).bakedSurf[i], &rd));
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    CRN_CUDA(c, cudaCreateTextureObject(&VS(c).ts.baked[i], &rd, &td, nullptr));
    VS(c).bakedN[i] = n;
    VS(c).bakeValid = false;
    return CRN_OK;
}

// (re)build whatever of the cone acceleration data is stale for this frame's volume / cone parameters / light
int build_cone_accel(crn_ctx *c, TraceParams &tp) {
    int r;
    if (c->nBakePlan > 0) {
        crn_ctx::BakeKey k{};
        k.gen = c->volumeGen; k.nTex = c->nBakePlan;
        for (int i = 0; i < c->nBakePlan; i++) {
            if ((r = ensure_baked_array(c, i, c->bakePlan[i].n))) return r;
            c->bakePlan[i].surf = VS(c).bakedSurf[i];
            k.level0[i] = c->bakePlan[i].level0; k.frac[i] = c->bakePlan[i].frac; k.n[i] = c->bakePlan[i].n;
        }
        if (!VS(c).bakeValid || std::memcmp(&k, &VS(c).bakeKey, sizeof k) != 0) {
            if (VS(c).freeValid) cudaStreamWaitEvent(c->lightStream, VS(c).evFree, 0);     // a trace may still be reading this set's baked textures
            c->launches += launch_bake_steps(c->lightStream, c->vparams, (const uint32_t *)c->bits.p, (const uint8_t *)c->chain.p, c->bakePlan, c->nBakePlan);
            VS(c).bakeKey = k; VS(c).bakeValid = true;
        }
    }
    for (int b = 0; b < tp.nBaked; b++) tp.baked[b].tex = (unsigned long long)VS(c).ts.baked[tp.baked[b].tex];
    return CRN_OK;
}

// the need codes are only computed inside the world bounding box of this frame's billboards, which the camera-side
// prep kernel produces: this runs after that kernel (the caller's stream has waited for the side stream)
int build_need_codes(crn_ctx *c, TraceParams &tp, const uint32_t *worldBox) {
    int r;
    if (tp.codeDim > 0) {
        const size_t cells = (size_t)tp.codeDim * tp.codeDim * tp.codeDim;
        if ((r = reserve(c, VS(c).needCode, cells))) return r;
        if (VS(c).codeG != tp.codeDim) {
            sync_all(c);
            free_code_texture(VS(c));
            cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
            CRN_CUDA(c, cudaMalloc3DArray(&VS(c).codeArr, &cd, make_cudaExtent(tp.codeDim, tp.codeDim, tp.codeDim), cudaArraySurfaceLoadStore));
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeArray; rd.res.array.array = VS(c).codeArr;
            CRN_CUDA(c, cudaCreateSurfaceObject(&VS(c).codeSurf, &rd));
            cudaTextureDesc td{};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;     // outside the volume: no skip bit
            td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
            CRN_CUDA(c, cudaCreateTextureObject(&VS(c).ts.code, &rd, &td, nullptr));
            VS(c).codeG = tp.codeDim;
            VS(c).codeValid = false;
        }
        crn_ctx::CodeKey k{};
        k.gen = c->volumeGen; k.G = tp.codeDim; k.nGroups = std::min(tp.nGroups, kCodeGroups);
        for (int g = 0; g < k.nGroups; g++) { k.height[g] = tp.groups[g].height; k.level[g] = tp.groups[g].level; k.first[g] = tp.groups[g].first; k.count[g] = tp.groups[g].count; }
        for (int i = 0; i < 3; i++) k.light[i] = tp.lightPos[i];
        const float b[6] = {c->vparams.xB[0], c->vparams.xB[1], c->vparams.yB[0], c->vparams.yB[1], c->vparams.zB[0], c->vparams.zB[1]};
        std::memcpy(k.bounds, b, sizeof b);
        // (the box changes with the billboards, which the key cannot see: rebuilt every frame unless the billboards are the
        //  ones of the last build)
        if (!VS(c).codeValid || VS(c).codeBoardsGen != c->boardsGen || std::memcmp(&k, &VS(c).codeKey, sizeof k) != 0) {
            if (VS(c).freeValid) cudaStreamWaitEvent(c->lightStream, VS(c).evFree, 0);
            c->launches += launch_need_code(c->lightStream, c->vparams, tp, (const uint32_t *)c->bits.p, (const uint32_t *)c->maskNz.p, worldBox, (uint8_t *)VS(c).needCode.p, VS(c).codeSurf);
            VS(c).codeKey = k; VS(c).codeValid = true; VS(c).codeBoardsGen = c->boardsGen;
        }
    }
    return CRN_OK;
}

// the caller holds device pointers into the volume (crn_volume_bits_ptr / crn_volume_level_ptr) or shards it by Z-slabs:
// its stream and the light stream are joined at every hand-over instead of running ahead of each other
bool strict_order(const crn_ctx *c) {
    return c->bitsExposed || c->levelExposed || (c->z1 >= 0 && !(c->z0 == 0 && c->z1 == c->vol.dimension));
}

int enqueue_voxelize(crn_ctx *c) {
    const int n = c->nBoards;
    crn_sun_derived sd;
    derive_sun(c->vol, c->sun, &sd);
    const ViewParams light = make_view(sd.P, sd.V, c->W, c->H);
    const ViewParams dummy = light;
    fill_vparams(c);
    int r;
    const size_t nn = std::max(n, 1);
    if ((r = reserve(c, c->keyL, nn * 8))) return r;
    if ((r = reserve(c, c->rankL, nn * 4))) return r;
    if ((r = reserve(c, c->recTmpL, nn * sizeof(BoardRec)))) return r;
    if ((r = reserve(c, c->rectTmpL, nn * sizeof(BoardRect)))) return r;
    if ((r = reserve(c, c->lbTmp, nn * 4))) return r;
    if ((r = reserve(c, c->recL, nn * sizeof(BoardRec)))) return r;
    if ((r = reserve(c, c->rectL, nn * sizeof(BoardRect)))) return r;
    if ((r = reserve(c, c->lbSorted, nn * 4))) return r;
    if ((r = reserve(c, c->sortTmp, sort_tmp_bytes((int)nn) + 128))) return r;
    const size_t D = c->vol.dimension;
    if ((r = reserve(c, c->bits, D * D * D / 8))) return r;
    if ((r = reserve(c, c->chain, c->chainBytes))) return r;
    const bool paper = c->vol.format == CRN_VOLUME_RG8;
    if (paper) {
        if (c->vparams.z0 != 0 || c->vparams.z1 != c->vol.dimension)
            return fail(c, CRN_ERR_UNSUPPORTED, "CRN_VOLUME_RG8 (paper variant) does not support Z-slab sharding");
        if ((r = reserve(c, c->bitsA, D * D * D / 8))) return r;
        if ((r = reserve(c, c->chainA, c->chainBytes))) return r;
    }
    if ((r = reserve(c, c->misc, 256))) return r;
    if (c->keepPosmap && (r = reserve(c, c->posmap, (size_t)c->W * c->H * 16))) return r;
    if ((r = ensure_bins(c, c->binsL, c->W, c->H, n))) return r;
    const bool toTex = c->tp.sampler == CRN_SAMPLER_TEXTURE;
    c->vs ^= 1;                                                          // this frame's textures go into the other set
    if (toTex && (r = ensure_vol_textures(c))) return r;

    // ---- light stream: the whole set-up of the frame.  Nothing here is read by a trace in flight: the bits, the chain
    //      and the masks are consumed on this stream (texture sampler), the textures written below belong to the set
    //      the trace before last used.
    cudaStream_t st = c->lightStream;
    cudaStreamWaitEvent(st, c->evBoards, 0);
    if (c->bitsFreeValid) cudaStreamWaitEvent(st, c->evBitsFree, 0);     // an explicit-sampler trace reads the bits and the chain directly
    if (strict_order(c)) {                                               // the caller may have queued work on the bits / the chain
        cudaEventRecord(c->evMainMark, c->stream);
        cudaStreamWaitEvent(st, c->evMainMark, 0);
    }
    if (c->curValid[0]) cudaStreamWaitEvent(st, c->evCur[0], 0);
    if (c->timingOn) cudaEventRecord(c->evV[0], st);
    const float zero3[3] = {0, 0, 0};
    c->launches += launch_prep_sort(st, (const float *)c->pos.p, (const float *)c->scale.p, n, c->vol.fluffiness, c->vol.position,
                                    light, sd.nearPlane, sd.clipDistance, dummy, zero3, true, false, (uint32_t *)c->rankL.p, nullptr,
                                    (uint64_t *)c->keyL.p, nullptr, (BoardRec *)c->recTmpL.p, nullptr, (BoardRect *)c->rectTmpL.p,
                                    nullptr, (float *)c->lbTmp.p, (BoardRec *)c->recL.p, nullptr, (BoardRect *)c->rectL.p, nullptr,
                                    (float *)c->lbSorted.p, nullptr, c->sortTmp.p);
    if (c->timingOn) cudaEventRecord(c->evV[1], st);
    c->launches += launch_bin(st, (const BoardRect *)c->rectL.p, sort_tmp_bounds(c->sortTmp.p, (int)nn, 0), n, c->W, c->H, c->binsL);
    // read the cursors back on the copy stream: a D2H on the compute stream would queue behind an image copy in
    // flight on the same copy engine and stall the kernels of this frame
    cudaEventRecord(c->evBin[0], st);
    cudaStreamWaitEvent(c->copyStream, c->evBin[0], 0);
    cudaMemcpyAsync(c->hCursors, c->binsL.cursors, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copyStream);
    cudaEventRecord(c->evCur[0], c->copyStream); c->curValid[0] = true;
    if (c->timingOn) cudaEventRecord(c->evV[2], st);
    c->launches += launch_voxelize(st, light, c->vparams, sd.nearPlane, sd.clipDistance, (const BoardRec *)c->recL.p,
                                   (const float *)c->lbSorted.p, c->binsL, (uint32_t *)c->bits.p,
                                   c->keepPosmap ? (float4 *)c->posmap.p : nullptr, paper ? (uint32_t *)c->bitsA.p : nullptr);
    if (c->timingOn) cudaEventRecord(c->evV[3], st);
    // ---- expand the bits into the chain / this set's textures / the masks
    if (VS(c).freeValid) cudaStreamWaitEvent(st, VS(c).evFree, 0);       // the last trace that sampled this set's textures
    if (c->timingOn) cudaEventRecord(c->evV[5], st);
    // the linear copy of level 0 (8x the bits) is not read by anything in the pipeline: it is written only for a caller
    // that holds crn_volume_level_ptr(0), and expanded on demand for crn_read_volume
    const bool whole = c->vparams.z0 == 0 && c->vparams.z1 == c->vol.dimension;
    // (a slab's texture copy would be overwritten after the exchange anyway: the textures are filled then)
    c->launches += launch_mips(st, c->vparams, (const uint32_t *)c->bits.p, (uint8_t *)c->chain.p, (uint32_t *)c->misc.p, c->wantLinear0,
                               (toTex && whole) ? &VS(c).ts : nullptr);
    c->linear0Valid = c->wantLinear0; c->linear0ValidA = false;
    if (paper)
        c->launches += launch_mips(st, c->vparams, (const uint32_t *)c->bitsA.p, (uint8_t *)c->chainA.p, (uint32_t *)c->misc.p + 8,
                                   false, toTex ? &VS(c).tsA : nullptr);
    VS(c).texCurrent = toTex && whole;
    c->maskCurrent = false;
    c->volumeGen++;
    if (whole && c->tp.skipEmptySpace) {
        if ((r = build_masks(c))) return r;
    }
    if (c->timingOn) { cudaEventRecord(c->evV[4], st); c->evVValid = true; }
    cudaEventRecord(c->evVoxDone, st);
    c->bitsFreeValid = false;
    // a caller that holds pointers into the volume (slab exchange) consumes it on ITS stream
    if (strict_order(c)) cudaStreamWaitEvent(c->stream, c->evVoxDone, 0);
    CRN_CUDA(c, cudaGetLastError());
    c->voxelized = true;
    return CRN_OK;
}

void build_trace_params(crn_ctx *c, const ViewParams &cam, TraceParams *tp) {
    std::memset(tp, 0, sizeof *tp);
    tp->p = c->tp;
    for (int k = 0; k < 3; k++) { tp->lightPos[k] = c->sun.position[k]; tp->octaveOffsets[k] = c->tp.windVel[k] * c->tp.runTime; }
    tp->octaveOffsets[3] = 0.0f;
    for (int k = 0; k < 4; k++) tp->bg[k] = c->tp.clearColor[k];
    tp->sun = c->sun;
    host_quad(cam, c->sun.position, c->sun.outerRadius, &tp->sunRec, &tp->sunRect);
    tp->nSteps = c->tp.vctSteps;
    tp->noiseDim = c->noiseDim;
    tp->noiseMask = (c->noiseDim & (c->noiseDim - 1)) == 0 ? c->noiseDim - 1 : -1;
    tp->row0 = std::max(0, c->row0); tp->row1 = std::min(c->H, c->row1);
    tp->ilvIndex = c->ilvIndex; tp->ilvCount = c->ilvCount;
    tp->active = (c->tp.doConeTrace || c->tp.doNoiseSample || c->tp.showQuad) ? 1 : 0;   // ConeTraceShader.cpp:16-18
    tp->stats = c->statsOn ? 1 : 0;
    // traceCone (res/conetrace_frag.glsl:64-79): per-step constants, float ops as written
    float coneHeight = c->tp.vctConeInitialHeight;
    const float tanHalf = tanf(c->tp.vctConeAngle / 2.0f);
    const int L = c->vol.levels;
    for (int i = 1; i <= c->tp.vctSteps; i++) {
        const float coneRadius = coneHeight * tanHalf;
        const float lod = log2f(fmaxf(1.0f, 2.0f * coneRadius)) + c->tp.vctLodOffset;
        ConeStep &s = tp->steps[i - 1];
        s.height = coneHeight;
        s.weight = (float)i / ((float)c->tp.vctSteps * c->tp.vctDownScaling);
        if (!(lod > 0.0f)) { s.level0 = 0; s.frac = 0.0f; }
        else if (lod >= (float)(L - 1)) { s.level0 = L - 1; s.frac = 0.0f; }
        else { const float fl = floorf(lod); s.level0 = (int)fl; s.frac = lod - fl; }
        s.lod0 = (float)s.level0; s.lod1 = (float)(s.level0 + 1); s.lod = s.lod0 + s.frac;
        coneHeight += coneRadius;
    }
    // Baked steps (k_conebake.cu): the longest suffix of the step list whose (level, fraction) lattices are small enough.
    // Only the texture sampler of the shipped formats uses them: the explicit sampler stays the full-precision path and
    // the paper variant needs its alpha gate per sample.
    const int D = c->vol.dimension, S = c->tp.vctSteps;
    tp->nFine = S; tp->nBaked = 0;
    c->nBakePlan = 0;
    if (c->tp.sampler == CRN_SAMPLER_TEXTURE && c->vol.format != CRN_VOLUME_RG8 && c->tp.doConeTrace && !c->noBake) {
        int first = S;
        size_t bytes = 0;
        int texOf[kMaxConeSteps];
        for (int i = S - 1; i >= 0 && S - i <= kMaxBakedSteps; i--) {
            const int l0 = tp->steps[i].level0;
            const float fr = tp->steps[i].frac;
            const int n = l0 == 0 ? 2 * D + 1 : (D >> (l0 - 1)) + 1;
            if (n > 129) break;
            int t = -1;
            for (int j = 0; j < c->nBakePlan; j++)
                if (c->bakePlan[j].level0 == l0 && c->bakePlan[j].frac == fr) t = j;
            if (t < 0) {
                const size_t need = (size_t)n * n * (n - 1) * 4;
                if (c->nBakePlan == kMaxBakedTex || bytes + need > ((size_t)128 << 20)) break;
                t = c->nBakePlan++;
                c->bakePlan[t].level0 = l0; c->bakePlan[t].frac = fr; c->bakePlan[t].n = n; c->bakePlan[t].surf = 0;
                bytes += need;
            }
            texOf[i] = t;
            first = i;
        }
        tp->nFine = first; tp->nBaked = S - first;
        for (int i = first; i < S; i++) {
            BakedStep &b = tp->baked[i - first];
            const int n = c->bakePlan[texOf[i]].n;
            b.height = tp->steps[i].height; b.weight = tp->steps[i].weight;
            b.A = (float)(n - 1) / (float)n; b.B = 0.5f / (float)n;
            b.zScale = (float)(n - 1) * (1.0f - 1.0f / 1048576.0f);
            b.hA = b.height * b.A;
            b.tex = (unsigned long long)texOf[i];         // plan index here; build_cone_accel swaps in the texture object
        }
        // drop plan entries no surviving step refers to (a break above may have left the last one unused)
        bool used[kMaxBakedTex] = {};
        for (int i = first; i < S; i++) used[texOf[i]] = true;
        while (c->nBakePlan > 0 && !used[c->nBakePlan - 1]) c->nBakePlan--;
    }
    // groups for the empty-space test (k_conebake.cu need codes): one per textureLod step; while there are more than the
    // code has bits, the two neighbours of the same level(s) that lie closest together are merged
    tp->nGroups = 0;
    for (int i = 0; i < tp->nFine; i++) {
        ConeGroup &g = tp->groups[tp->nGroups++];
        g.first = i; g.count = 1; g.level = tp->steps[i].level0; g.two = tp->steps[i].frac != 0.0f ? 1 : 0;
        g.height = tp->steps[i].height;
    }
    while (tp->nGroups > kCodeGroups) {
        int best = -1;
        float bestSpan = 0.0f;
        for (int g = 0; g + 1 < tp->nGroups; g++) {
            const ConeGroup &p = tp->groups[g], &q = tp->groups[g + 1];
            if (p.level != q.level) continue;
            const float span = tp->steps[q.first + q.count - 1].height - tp->steps[p.first].height;
            if (best < 0 || span < bestSpan) { best = g; bestSpan = span; }
        }
        if (best < 0) break;                               // (groups past kCodeGroups have no bit: always fetched)
        ConeGroup &p = tp->groups[best];
        const ConeGroup &q = tp->groups[best + 1];
        p.count += q.count; p.two |= q.two;
        for (int g = best + 1; g + 1 < tp->nGroups; g++) tp->groups[g] = tp->groups[g + 1];
        tp->nGroups--;
    }
    for (int g = 0; g < tp->nGroups; g++) {
        ConeGroup &p = tp->groups[g];
        p.height = 0.5f * (tp->steps[p.first].height + tp->steps[p.first + p.count - 1].height);
    }
    // need-code grid: cells of two voxels; only worth a kernel when some step is still fetched with textureLod
    // (at most 128^3 cells: the kernel's cost goes with the cell count, and a 512^3 volume's cones reach half as far in world units)
    tp->codeDim = (c->tp.skipEmptySpace && tp->nGroups > 0 && c->tp.doConeTrace) ? std::max(16, std::min(D / 2, 128)) : 0;
    tp->codeDimF = (float)tp->codeDim;
    {   // derived constants of the fast variant
        FastConst &f = tp->f;
        f.invP0 = 1.0f / cam.P[0]; f.invP5 = 1.0f / cam.P[5];
        f.invAdjust = 1.0f / c->tp.adjustSize; f.invStep = 1.0f / c->tp.stepSize;
        f.span = (float)(c->tp.maxNoiseSteps - c->tp.minNoiseSteps); f.minSteps = (float)c->tp.minNoiseSteps;
        const float lo[3] = {c->vparams.xB[0], c->vparams.yB[0], c->vparams.zB[0]}, hi[3] = {c->vparams.xB[1], c->vparams.yB[1], c->vparams.zB[1]};
        for (int k = 0; k < 3; k++) { f.nScale[k] = 1.0f / (hi[k] - lo[k]); f.nBias[k] = -lo[k] * f.nScale[k]; }
        f.invDim = 1.0f / (float)c->vol.dimension;
    }
    // noise3D's per-octave constants (res/conetrace_frag.glsl:107-114)
    float freq = 1.0f, pers = 1.0f;
    for (int o = 0; o < kMaxOctaves; o++) {
        tp->octFreq[o] = freq; tp->octPers[o] = pers;
        tp->octBias[o] = (o < 3 ? tp->octaveOffsets[o] : 0.0f) * freq;
        tp->octFreqZ[o] = tp->octFreq[o] * (float)c->noiseDim;
        tp->octBiasZ[o] = tp->octBias[o] * (float)c->noiseDim - 0.5f;
        freq *= c->tp.freqStep; pers *= c->tp.persStep;
    }
}

// The combined-octave noise lattice (k_noiselat.cu) for this frame's parameters, or tp.lat.on = 0 when they do not
// qualify: texture sampler, 4 octaves (what the fast trace variant handles), freqStep an odd integer (every coarser
// octave's texel centres then fall on the finest octave's), no wind offset on octaves 1 and 2, adjustSize > 0, and a
// window that fits the memory cap.  The window is the volume's box grown by the billboards' reach when the host knows it
// (host uploads, generated sets), by 20 % otherwise; the kernel checks every billboard against it, so a set that pokes
// out of the window is still rendered exactly (those billboards take the per-octave lookups).
int ensure_noise_lattice(crn_ctx *c, const ViewParams &cam, TraceParams &tp) {
    NoiseLat &L = tp.lat;
    L.on = 0;
    const crn_trace_params &p = c->tp;
    // (only the fast trace variant reads it: the same conditions as launch_trace's choice, k_trace.cu)
    if (c->noLattice || p.sampler != CRN_SAMPLER_TEXTURE || !p.doNoiseSample || p.numOctaves != 4 || c->noiseDim != 32 || !p.doConeTrace || p.showQuad ||
        p.quantizeFramebuffer || cam.ortho || c->vol.format == CRN_VOLUME_RG8 || getenv("CRN_NO_FAST")) return CRN_OK;   // (stats frames keep it: they report what the fast variant does)
    const float fs = p.freqStep;
    if (!(fs >= 1.0f && fs <= 9.0f) || fs != floorf(fs) || ((int)fs & 1) == 0) return CRN_OK;
    if (!(p.adjustSize > 0.0f) || tp.octaveOffsets[1] != 0.0f || tp.octaveOffsets[2] != 0.0f) return CRN_OK;
    const int F = p.numOctaves - 1;
    const double K = (double)tp.octFreq[F] * c->noiseDim;               // lattice units per unit of uv
    // window, world units
    const float lo[3] = {c->vparams.xB[0], c->vparams.yB[0], c->vparams.zB[0]}, hi[3] = {c->vparams.xB[1], c->vparams.yB[1], c->vparams.zB[1]};
    int n[3]; long long base[3];
    double wlo[3], whi[3];
    size_t nodes = 1;
    for (int k = 0; k < 3; k++) {
        const double half = 0.5 * ((double)hi[k] - lo[k]), mid = 0.5 * ((double)hi[k] + lo[k]);
        double reach = 1.2 * half;
        if (c->boardPosMax >= 0.0f)       // the march runs from the near hit (-h along the ray) to at most 3h: |o_k + s ray_k| <= r sqrt(1 + 9 ray_k^2)
            reach = std::max(reach, std::ceil((double)c->boardPosMax + 3.2 * c->boardScaleMax * c->vol.fluffiness + std::fabs(mid - c->vol.position[k])));
        reach = std::min(reach, 2.5 * half);                               // (a cap: billboards beyond it take the per-octave path)
        wlo[k] = mid - reach; whi[k] = mid + reach;
        const long long b = (long long)std::floor(wlo[k] / p.adjustSize * K - 0.5) - 1;
        const long long t = (long long)std::ceil(whi[k] / p.adjustSize * K - 0.5) + 1;
        base[k] = b; n[k] = (int)(t - b + 1);
        nodes *= (size_t)n[k];
    }
    if (nodes * 8 > ((size_t)512 << 20)) {                                 // memory cap: shrink the window around the volume's centre
        const double f = std::cbrt((double)((size_t)512 << 20) / (double)(nodes * 8)) * 0.99;
        nodes = 1;
        for (int k = 0; k < 3; k++) {
            const double mid = 0.5 * (wlo[k] + whi[k]), reach = 0.5 * (whi[k] - wlo[k]) * f;
            wlo[k] = mid - reach; whi[k] = mid + reach;
            const long long b = (long long)std::floor(wlo[k] / p.adjustSize * K - 0.5) - 1;
            const long long t = (long long)std::ceil(whi[k] / p.adjustSize * K - 0.5) + 1;
            base[k] = b; n[k] = (int)(t - b + 1);
            nodes *= (size_t)n[k];
        }
    }
    for (int k = 0; k < 3; k++)
        if (n[k] < 4 || n[k] > 2000) return CRN_OK;
    if ((size_t)n[1] * n[2] / 8 + 1 > 65535) return CRN_OK;
    crn_ctx::LatKey key{};
    key.noiseGen = c->noiseGen; key.dim = c->noiseDim; key.octaves = p.numOctaves; key.freqStep = fs; key.persStep = p.persStep; key.adjust = p.adjustSize;
    for (int k = 0; k < 3; k++) { key.n[k] = n[k]; key.base[k] = base[k]; }
    float scale = 0.0f;
    for (int o = 1; o <= F; o++) scale += fabsf(tp.octPers[o]);
    if (!(scale > 0.0f)) return CRN_OK;
    if (!c->latValid || std::memcmp(&key, &c->latKey, sizeof key) != 0) {
        if (!c->latArray || c->latN[0] != n[0] || c->latN[1] != n[1] || c->latN[2] != n[2]) {
            sync_all(c);
            free_lattice(c);
            cudaChannelFormatDesc cd = cudaCreateChannelDesc(16, 16, 16, 16, cudaChannelFormatKindSigned);
            CRN_CUDA(c, cudaMalloc3DArray(&c->latArray, &cd, make_cudaExtent(n[0], n[1], n[2] - 1), cudaArrayLayered | cudaArraySurfaceLoadStore));
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeArray; rd.res.array.array = c->latArray;
            CRN_CUDA(c, cudaCreateSurfaceObject(&c->latSurf, &rd));
            cudaTextureDesc td{};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
            CRN_CUDA(c, cudaCreateTextureObject(&c->latTex, &rd, &td, nullptr));
            for (int k = 0; k < 3; k++) c->latN[k] = n[k];
        }
        // every trace in flight may be reading the old lattice
        if (c->traceEndValid) cudaStreamWaitEvent(c->lightStream, c->evTraceEnd, 0);
        double m[kMaxOctaves] = {};
        for (int o = 0; o <= F; o++) m[o] = (double)tp.octFreq[F] / (double)tp.octFreq[o];
        c->launches += launch_noise_lattice(c->lightStream, (const float2 *)c->noise.p, c->noiseDim, n, base, 1, F, m, tp.octPers, 1.0f / scale, c->latSurf);
        CRN_CUDA(c, cudaGetLastError());
        c->latKey = key; c->latValid = true;
    }
    L.on = 1;
    L.K = (float)K;
    for (int k = 0; k < 3; k++) {
        // uv = world / adjustSize; lattice coordinate = uv * K - 1/2; node index = that - base
        L.B[k] = (float)(-0.5 - (double)base[k] + (k < 2 ? 0.5 : 0.0));
        // the window the kernel tests against: one node inside the baked box on every side
        const double a = ((double)base[k] + 1.0 + 0.5) / K * p.adjustSize, b = ((double)base[k] + n[k] - 2.0 + 0.5) / K * p.adjustSize;
        L.winC[k] = (float)(0.5 * (a + b)); L.winH[k] = (float)(0.5 * (b - a) * (1.0 - 1e-5));
        L.ext[k] = sqrtf(1.0f + 9.0f * cam.nrm[k] * cam.nrm[k]) * 1.001f;
    }
    L.scale = scale;
    L.tex = (unsigned long long)c->latTex;
    return CRN_OK;
}

int enqueue_trace(crn_ctx *c, int format, DevBuf *target = nullptr) {
    DevBuf &img = target ? *target : c->image;
    const int n = c->nBoards;
    const ViewParams cam = make_view(c->cam.P, c->cam.V, c->W, c->H);
    fill_vparams(c);
    c->cs ^= 1;                                        // this trace's camera-side buffers: the other set
    int r;
    const size_t nn = std::max(n, 1);
    if ((r = reserve(c, c->keyC, nn * 8))) return r;
    if ((r = reserve(c, c->rankC, nn * 4))) return r;
    if ((r = reserve(c, c->recTmpC, nn * sizeof(BoardRec)))) return r;
    if ((r = reserve(c, c->rectTmpC, nn * sizeof(BoardRect)))) return r;
    if ((r = reserve(c, CS(c).recC, nn * sizeof(BoardRec)))) return r;
    if ((r = reserve(c, CS(c).rectC, nn * sizeof(BoardRect)))) return r;
    if ((r = reserve(c, c->drawOrder, nn * 4))) return r;
    if ((r = reserve(c, CS(c).sortTmpC, sort_tmp_bytes((int)nn) + 128))) return r;
    if ((r = reserve(c, c->misc, 256))) return r;
    const size_t texel = format == CRN_IMAGE_RGBA32F ? 16 : 4;
    if ((r = reserve(c, img, (size_t)c->W * c->H * texel))) return r;
    if ((r = ensure_bins(c, CS(c).binsC, c->W, c->H, n))) return r;
    if ((r = reserve(c, CS(c).tileOrder, ((size_t)CS(c).binsC.tilesX * CS(c).binsC.tilesY + 66) * 4))) return r;

    TraceParams tp;
    build_trace_params(c, cam, &tp);
    {   // small frames leave SMs idle behind the longest tile list: cut the lists (see trace_fast_kernel)
        const size_t tiles = (size_t)CS(c).binsC.tilesX * CS(c).binsC.tilesY;
        // measured trace ms at 1 / 2 / 4 segments: C1 (3600 tiles) 0.244 / 0.188 / 0.165, C2 (8160) 0.807 / 0.715 / 0.632,
        // C3 (32400) 3.015 / 2.955 / 3.086, C4 (129600) 9.46 / 10.12 / 10.65; a rank of an interleaved trace owns 1/count of the tiles
        const size_t active = tiles / (size_t)std::max(1, c->ilvCount);
        // final kernel: C1 0.280 / 0.185 / 0.143, C2 0.669 / 0.504 / 0.473, C3 2.173 / 2.180 / 2.324.  Stand-alone, C3 does not care
        // between 1 and 2 — but PIPELINED frames do: with the lists cut in two the CTAs live half as long, the high-priority
        // set-up kernels and the read-back of the neighbouring frames get their slots sooner: 380 -> 392 frames/s resident,
        // 347 -> 388 end to end (bench.py, segments 1 / 2 / 3 / 4: 380 / 392 / 383 / 377 and 347 / 388 / 381 / 372)
        // C2 pipelined: 2 / 4 / 6 segments 1685 / 1631 / 1531 frames/s; C1: 2 / 4 / 8 -> 3885 / 4564 / 3674; C4 (129600 tiles): 1 / 2 -> 128.5 / 122.9
        // (C4 on two GPUs, 64800 owned tiles: 1 / 2 segments -> 228 / 220 frames/s)
        tp.segCount = (active <= 4096 || (c->ilvCount > 1 && active <= 16384)) ? 4 : active <= 49152 ? 2 : 1;
        // Frames whose lists are very long on average (the camera-side bin entries of the last frame whose cursors have been
        // read back: reference radii at C3 hold 430 entries per tile of the WHOLE frame, the fill radii 23) gain nothing from
        // segments — their lists are over the cut limit anyway — and pay for the doubled grid: 90 against 96 frames/s
        if (c->hCursors && (size_t)c->hCursors[5] > 128 * tiles) tp.segCount = 1;
        if (c->segOverride >= 1) tp.segCount = std::min(c->segOverride, 16);
        tp.segMin = 6; tp.segMax = c->segMaxOverride > 0 ? c->segMaxOverride : 1024;
        if (tp.segCount > 1) {
            if ((r = reserve(c, c->segPartial, tiles * tp.segCount * 256 * sizeof(float4)))) return r;
            if ((r = reserve(c, c->segArrived, tiles * 4 * sizeof(uint32_t)))) return r;      // zeroed by reserve, re-armed by the kernel
        }
    }
    cudaStream_t st = c->stream, ls = c->lightStream;
    const bool useTex = c->tp.sampler == CRN_SAMPLER_TEXTURE;
    int ownedTiles = 0;                                // tile rows this context traces (crn_set_tile_row_interleave): only they get CTAs
    for (int ty = 0; ty < CS(c).binsC.tilesY; ty++)
        if (ty % std::max(1, c->ilvCount) == c->ilvIndex) ownedTiles += CS(c).binsC.tilesX;
    // ---- trace set-up on the light stream: after this frame's voxelize (same stream), before the next frame's
    if (strict_order(c)) {                             // whatever the caller queued on its stream (slab exchange, crn_finish_mips)
        cudaEventRecord(c->evMainMark, st);
        cudaStreamWaitEvent(ls, c->evMainMark, 0);
    }
    if (c->tp.skipEmptySpace && !c->maskCurrent) {     // chain came from an exchange, or the option was just switched on
        if ((r = build_masks(c))) return r;
    }
    if (useTex) {
        if ((r = ensure_vol_textures(c))) return r;
        if (!VS(c).texCurrent) {           // sampler switched after voxelize, or the chain came from an exchange
            if (VS(c).freeValid) cudaStreamWaitEvent(ls, VS(c).evFree, 0);
            // level 0 from the bits (they are what a slab exchange ships), the coarser levels from the linear chain
            c->launches += launch_expand_level0(ls, c->vparams, (const uint32_t *)c->bits.p, nullptr, &VS(c).ts);
            c->launches += launch_chain_to_surfaces(ls, c->vparams, (const uint8_t *)c->chain.p, VS(c).ts, 1);
            if (c->vol.format == CRN_VOLUME_RG8) {
                c->launches += launch_expand_level0(ls, c->vparams, (const uint32_t *)c->bitsA.p, nullptr, &VS(c).tsA);
                c->launches += launch_chain_to_surfaces(ls, c->vparams, (const uint8_t *)c->chainA.p, VS(c).tsA, 1);
            }
            VS(c).texCurrent = true;
        }
    }
    if (c->timingOn) cudaEventRecord(c->evT[2], ls);                // the cone acceleration data is part of the trace stage
    if ((r = ensure_noise_lattice(c, cam, tp))) return r;
    if ((r = build_cone_accel(c, tp))) return r;
    if (c->timingOn) cudaEventRecord(c->evT[1], ls);
    // ---- camera-side set-up on the side stream, into the camera set the trace before last used: concurrent with the
    //      light stream's work AND with the previous trace
    cudaStream_t ax = c->auxStream;
    cudaStreamWaitEvent(ax, c->evBoards, 0);
    if (CS(c).freeValid) cudaStreamWaitEvent(ax, CS(c).evFree, 0);
    if (c->curValid[1]) cudaStreamWaitEvent(ax, c->evCur[1], 0);
    if (c->timingOn) cudaEventRecord(c->evAuxT[0], ax);
    const float zero3[3] = {0, 0, 0};
    c->launches += launch_prep_sort(ax, (const float *)c->pos.p, (const float *)c->scale.p, n, c->vol.fluffiness, c->vol.position, cam,
                                    zero3, 1.0f, cam, c->cam.position, false, true, nullptr, (uint32_t *)c->rankC.p, nullptr,
                                    (uint64_t *)c->keyC.p, nullptr, (BoardRec *)c->recTmpC.p, nullptr, (BoardRect *)c->rectTmpC.p,
                                    nullptr, nullptr, (BoardRec *)CS(c).recC.p, nullptr, (BoardRect *)CS(c).rectC.p, nullptr,
                                    (int32_t *)c->drawOrder.p, CS(c).sortTmpC.p, &tp.lat);
    if (c->timingOn) cudaEventRecord(c->evAuxT[1], ax);
    c->launches += launch_bin(ax, (const BoardRect *)CS(c).rectC.p, sort_tmp_bounds(CS(c).sortTmpC.p, (int)nn, 1), n, c->W, c->H, CS(c).binsC);
    cudaEventRecord(c->evBin[1], ax);
    cudaStreamWaitEvent(c->copyStream, c->evBin[1], 0);
    cudaMemcpyAsync(c->hCursors + 4, CS(c).binsC.cursors, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copyStream);
    cudaEventRecord(c->evCur[1], c->copyStream); c->curValid[1] = true;
    c->launches += launch_tile_order(ax, CS(c).binsC, (uint32_t *)CS(c).tileOrder.p, c->ilvIndex, std::max(1, c->ilvCount));
    if (c->timingOn) cudaEventRecord(c->evAuxT[2], ax);
    cudaEventRecord(c->evAuxDone, ax);
    // ---- need codes: inside the billboards' world box, which the camera-side prep kernel has just produced
    cudaStreamWaitEvent(ls, c->evAuxDone, 0);
    if (c->timingOn) cudaEventRecord(c->evAccT[0], ls);
    if ((r = build_need_codes(c, tp, n > 0 ? sort_tmp_world_box(CS(c).sortTmpC.p, (int)nn) : nullptr))) return r;
    if (c->timingOn) cudaEventRecord(c->evAccT[1], ls);
    cudaEventRecord(c->evPrepared, ls);
    // ---- the trace itself, on the caller's stream
    cudaStreamWaitEvent(st, c->evPrepared, 0);                      // (the light stream has already waited for the side stream)
    unsigned long long *dStats = (unsigned long long *)((char *)c->misc.p + 64);
    if (c->statsOn) cudaMemsetAsync(dStats, 0, 8 * sizeof(unsigned long long), st);
    if (c->timingOn) cudaEventRecord(c->evT[0], st);
    c->launches += launch_trace(st, cam, c->vparams, tp, (const BoardRec *)CS(c).recC.p, CS(c).binsC, (const uint32_t *)c->bits.p,
                                (const uint8_t *)c->chain.p, c->vol.format == CRN_VOLUME_RG8 ? (const uint32_t *)c->bitsA.p : nullptr,
                                (const uint8_t *)c->chainA.p, (const int8_t *)c->noise.p, useTex ? &VS(c).ts : nullptr,
                                tp.codeDim > 0 ? (const uint8_t *)VS(c).needCode.p : nullptr, (const uint32_t *)CS(c).tileOrder.p,
                                img.p, format, dStats, tp.segCount > 1 ? (float4 *)c->segPartial.p : nullptr,
                                tp.segCount > 1 ? (uint32_t *)c->segArrived.p : nullptr, ownedTiles);
    cudaEventRecord(VS(c).evFree, st); VS(c).freeValid = true;
    cudaEventRecord(CS(c).evFree, st); CS(c).freeValid = true;
    cudaEventRecord(c->evTraceEnd, st);
    c->traceEndValid = true;
    c->auxDoneValid = true;
    if (!useTex) {                                     // the explicit sampler reads the bits and the chain under the next voxelize
        cudaEventRecord(c->evBitsFree, st);
        c->bitsFreeValid = true;
    }
    if (c->timingOn) { cudaEventRecord(c->evT[3], st); c->evTValid = true; }
    if (c->statsOn) cudaMemcpyAsync(c->hStats, dStats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    CRN_CUDA(c, cudaGetLastError());
    return CRN_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------
// exported
// ------------------------------------------------------------------------------------------
extern "C" {

const char *crn_version(void) { return "cloud-renderer_b200 0.1.0 sm_100a"; }

const char *crn_last_error(const crn_ctx *ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> g(g_err_mutex);
    return g_create_error.c_str();
}

int crn_create(int device, void *stream, crn_ctx **out) {
    if (!out) return fail(nullptr, CRN_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, CRN_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, CRN_ERR_INVALID_ARG, "device %d out of range [0,%d)", device, count);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, CRN_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    crn_ctx *c = new crn_ctx();
    c->device = device;
    if (stream) c->stream = (cudaStream_t)stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete c;
            return fail(nullptr, CRN_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->ownStream = true;
    }
    cudaMallocHost(&c->hCursors, 8 * sizeof(uint32_t));
    cudaMallocHost(&c->hStats, 8 * sizeof(unsigned long long));
    std::memset(c->hCursors, 0, 8 * sizeof(uint32_t));
    std::memset(c->hStats, 0, 8 * sizeof(unsigned long long));
    // The set-up streams run small kernels next to a trace kernel that has hundreds of thousands of CTAs queued: at equal
    // priority the block scheduler finishes dispatching the trace grid first and the "overlap" happens at its tail only
    // (measured: pipelined == serialised frame time).  Highest priority lets their CTAs in as soon as trace CTAs retire.
    int prLo = 0, prHi = 0;
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);                 // numerically lower = higher priority
    cudaStreamCreateWithPriority(&c->copyStream, cudaStreamNonBlocking, prHi);
    cudaStreamCreateWithPriority(&c->auxStream, cudaStreamNonBlocking, prHi);
    cudaStreamCreateWithPriority(&c->lightStream, cudaStreamNonBlocking, prHi);
    cudaEventCreateWithFlags(&c->evLightDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evPrepared, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evVoxDone, cudaEventDisableTiming);
    for (auto &ev : c->evAccT) cudaEventCreate(&ev);
    for (int k = 0; k < 2; k++) {
        cudaEventCreateWithFlags(&c->vset[k].evFree, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->cset[k].evFree, cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&c->evBitsFree, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evMainMark, cudaEventDisableTiming);
    for (auto &ev : c->evCur) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evBoards, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evAuxDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evTraceEnd, cudaEventDisableTiming);
    for (auto &ev : c->evAuxT) cudaEventCreate(&ev);
    for (int k = 0; k < 2; k++) {
        cudaEventCreateWithFlags(&c->evFrame[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->evCopy[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->evBin[k], cudaEventDisableTiming);
    }
    for (auto &ev : c->evV) cudaEventCreate(&ev);
    for (auto &ev : c->evT) cudaEventCreate(&ev);
    crn_default_trace_params(&c->tp);
    if (const char *nb = getenv("CRN_NO_BAKE")) c->noBake = atoi(nb) != 0;
    if (const char *nl = getenv("CRN_NO_LATTICE")) c->noLattice = atoi(nl) != 0;
    if (const char *sg = getenv("CRN_TRACE_SEGMENTS")) c->segOverride = atoi(sg);
    if (const char *sm = getenv("CRN_TRACE_SEGMAX")) c->segMaxOverride = atoi(sm);
    if (const char *pm = getenv("CRN_BIN_POOL_MIN")) { const long v = atol(pm); if (v > 0) c->poolMin = (size_t)v; }
    if ((e = cudaGetLastError()) != cudaSuccess) {
        crn_destroy(c);
        return fail(nullptr, CRN_ERR_CUDA, "context set-up: %s", cudaGetErrorString(e));
    }
    *out = c;
    return CRN_OK;
}

void crn_destroy(crn_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    if (c->auxStream) cudaStreamSynchronize(c->auxStream);
    if (c->lightStream) cudaStreamSynchronize(c->lightStream);
    DevBuf *bufs[] = {&c->exportTmp, &c->exportOut, &c->pos0, &c->bitsA, &c->chainA, &c->pos, &c->scale, &c->keyL, &c->keyC, &c->rankL, &c->rankC, &c->recTmpL, &c->recTmpC, &c->rectTmpL,
                      &c->rectTmpC, &c->lbTmp, &c->recL, &c->rectL, &c->lbSorted, &c->drawOrder, &c->bits, &c->segPartial, &c->segArrived,
                      &c->chain, &c->noise, &c->posmap, &c->image, &c->image2, &c->misc, &c->maskNz, &c->sortTmp,
                      &c->cset[0].recC, &c->cset[0].rectC, &c->cset[0].sortTmpC, &c->cset[0].tileOrder,
                      &c->cset[1].recC, &c->cset[1].rectC, &c->cset[1].sortTmpC, &c->cset[1].tileOrder, &c->vset[0].needCode, &c->vset[1].needCode};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    free_bins(c->binsL); free_bins(c->cset[0].binsC); free_bins(c->cset[1].binsC);
    for (c->vs = 0; c->vs < 2; c->vs++) {
        free_vol_textures(c);
        for (int i = 0; i < kMaxBakedTex; i++) free_baked(c, i);
        free_code_texture(c->vset[c->vs]);
    }
    c->vs = 0;
    for (int k = 0; k < 2; k++) {
        if (c->vset[k].evFree) cudaEventDestroy(c->vset[k].evFree);
        if (c->cset[k].evFree) cudaEventDestroy(c->cset[k].evFree);
    }
    if (c->evPrepared) cudaEventDestroy(c->evPrepared);
    if (c->evVoxDone) cudaEventDestroy(c->evVoxDone);
    for (auto &ev : c->evAccT) if (ev) cudaEventDestroy(ev);
    if (c->noiseTex) cudaDestroyTextureObject(c->noiseTex);
    if (c->noiseArray) cudaFreeArray(c->noiseArray);
    if (c->noiseTexD) cudaDestroyTextureObject(c->noiseTexD);
    if (c->noiseArrayD) cudaFreeArray(c->noiseArrayD);
    free_lattice(c);
    if (c->hCursors) cudaFreeHost(c->hCursors);
    if (c->hStats) cudaFreeHost(c->hStats);
    for (auto &ev : c->evV) if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->evT) if (ev) cudaEventDestroy(ev);
    if (c->auxStream) {
        cudaEventDestroy(c->evBoards); cudaEventDestroy(c->evAuxDone); cudaEventDestroy(c->evTraceEnd);
        for (auto &ev : c->evAuxT) if (ev) cudaEventDestroy(ev);
        cudaStreamDestroy(c->auxStream);
    }
    if (c->copyStream) {
        cudaStreamSynchronize(c->copyStream);
        for (int k = 0; k < 2; k++) {
            if (c->evFrame[k]) cudaEventDestroy(c->evFrame[k]);
            if (c->evCopy[k]) cudaEventDestroy(c->evCopy[k]);
            if (c->evBin[k]) cudaEventDestroy(c->evBin[k]);
        }
        cudaStreamDestroy(c->copyStream);
    }
    if (c->lightStream) {
        cudaEventDestroy(c->evLightDone); cudaEventDestroy(c->evBitsFree); cudaEventDestroy(c->evMainMark);
        for (auto &ev : c->evCur) if (ev) cudaEventDestroy(ev);
        cudaStreamDestroy(c->lightStream);
    }
    if (c->ownStream) cudaStreamDestroy(c->stream);
    delete c;
}

static int settle(crn_ctx *c, bool haveTrace, int format, bool slabNotShippedYet = false);

int crn_sync(crn_ctx *c) {
    if (!c) return CRN_ERR_INVALID_ARG;
    CRN_CUDA(c, cudaSetDevice(c->device));
    if (c->voxelized) return settle(c, false, 0);           // also re-runs a voxelize whose bin pool was too small
    if (c->lightStream) CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
    if (c->auxStream) CRN_CUDA(c, cudaStreamSynchronize(c->auxStream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copyStream) CRN_CUDA(c, cudaStreamSynchronize(c->copyStream));
    return CRN_OK;
}

void crn_default_trace_params(crn_trace_params *p) {          // src/Shaders/ConeTraceShader.hpp:15-36
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->stepSize = 0.01f; p->noiseOpacity = 4.0f; p->numOctaves = 4; p->freqStep = 3.0f; p->persStep = 0.5f;
    p->adjustSize = 40.0f; p->minNoiseSteps = 2; p->maxNoiseSteps = 8; p->minNoiseColor = 0.2f; p->noiseColorScale = 0.45f;
    p->windVel[0] = 0.01f; p->windVel[1] = 0.0f; p->windVel[2] = 0.0f;
    p->vctSteps = 16; p->vctConeAngle = 0.9f; p->vctConeInitialHeight = 0.1f; p->vctLodOffset = 0.0f; p->vctDownScaling = 1.0f;
    p->showQuad = 0; p->doConeTrace = 1; p->doNoiseSample = 1;
    p->runTime = 0.0f;
    p->clearColor[0] = 0.2f; p->clearColor[1] = 0.3f; p->clearColor[2] = 0.5f; p->clearColor[3] = 1.0f;   // src/main.cpp:112
    p->drawSun = 1;
    p->transmittanceCutoff = 0.0f;
    p->sampler = CRN_SAMPLER_EXPLICIT;
    p->skipEmptySpace = 1;
}

int crn_set_volume(crn_ctx *c, const crn_volume_desc *d) {
    if (!c || !d) return CRN_ERR_INVALID_ARG;
    int r = check_volume(c, d); if (r) return r;
    if (c->haveVol && (c->vol.dimension != d->dimension || c->vol.levels != d->levels || c->vol.format != d->format)) { c->voxelized = false; c->z1 = -1; c->z0 = 0; }
    c->vol = *d; c->haveVol = true;
    c->boardsGen++;                      // position / fluffiness move the spheres
    return CRN_OK;
}

int crn_set_billboards(crn_ctx *c, const float *positions3, const float *scales, int32_t count, int32_t mem) {
    if (!c || count < 0 || (count > 0 && (!positions3 || !scales))) return fail(c, CRN_ERR_INVALID_ARG, "bad billboard arrays");
    if (mem != CRN_MEM_HOST && mem != CRN_MEM_DEVICE) return fail(c, CRN_ERR_INVALID_ARG, "mem must be CRN_MEM_HOST or CRN_MEM_DEVICE");
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r;
    if ((r = reserve(c, c->pos, (size_t)std::max(count, 1) * 12))) return r;
    if ((r = reserve(c, c->scale, (size_t)std::max(count, 1) * 4))) return r;
    cudaStream_t up = upload_stream(c, mem == CRN_MEM_DEVICE);
    if (count) {
        const cudaMemcpyKind kind = mem == CRN_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        CRN_CUDA(c, cudaMemcpyAsync(c->pos.p, positions3, (size_t)count * 12, kind, up));
        CRN_CUDA(c, cudaMemcpyAsync(c->scale.p, scales, (size_t)count * 4, kind, up));
    }
    CRN_CUDA(c, cudaEventRecord(c->evBoards, up));
    // a device-resident source is read on the light stream: make the caller's stream wait for that read, so that the
    // caller may overwrite or free the source in stream order right after this call
    if (mem == CRN_MEM_DEVICE && count) CRN_CUDA(c, cudaStreamWaitEvent(c->stream, c->evBoards, 0));
    c->nBoards = count;
    c->havePos0 = false;
    c->boardsGen++;
    c->boardPosMax = -1.0f; c->boardScaleMax = 0.0f;
    if (mem == CRN_MEM_HOST) {                           // what the noise-lattice window is sized from (ensure_noise_lattice)
        float pm = 0.0f, sm = 0.0f;
        for (int i = 0; i < count; i++) {
            const float x = positions3[3 * i], y = positions3[3 * i + 1], z = positions3[3 * i + 2];
            pm = fmaxf(pm, fmaxf(x * x + z * z, y * y));
            sm = fmaxf(sm, fabsf(scales[i]));
        }
        c->boardPosMax = sqrtf(pm); c->boardScaleMax = sm;
    }
    return CRN_OK;
}

int crn_regenerate_billboards(crn_ctx *c, int32_t count, const float minOffset[3], const float maxOffset[3], float minScale,
                              float maxScale, double radiusFactor, uint64_t seed) {   // src/CloudVolume.cpp:120-137
    if (!c || count < 0 || !minOffset || !maxOffset) return fail(c, CRN_ERR_INVALID_ARG, "bad generator arguments");
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r;
    if ((r = reserve(c, c->pos0, (size_t)std::max(count, 1) * 12))) return r;
    if ((r = reserve(c, c->pos, (size_t)std::max(count, 1) * 12))) return r;
    if ((r = reserve(c, c->scale, (size_t)std::max(count, 1) * 4))) return r;
    cudaStream_t up = upload_stream(c, false);
    c->launches += launch_generate_boards(up, count, minOffset, maxOffset, minScale, maxScale, radiusFactor, seed,
                                          (float *)c->pos0.p, (float *)c->pos.p, (float *)c->scale.p);
    CRN_CUDA(c, cudaGetLastError());
    CRN_CUDA(c, cudaEventRecord(c->evBoards, up));
    c->nBoards = count;
    c->havePos0 = true;
    c->boardsGen++;
    {
        float m[3];
        for (int k = 0; k < 3; k++) m[k] = fmaxf(fabsf(minOffset[k]), fabsf(maxOffset[k]));
        c->boardPosMax = fmaxf(sqrtf(m[0] * m[0] + m[2] * m[2]), m[1]);
    }
    c->boardScaleMax = (float)(fmax(fabs((double)minScale), fabs((double)maxScale)) * fabs(radiusFactor));
    return CRN_OK;
}

int crn_animate_billboards(crn_ctx *c, double angle) {
    if (!c) return CRN_ERR_INVALID_ARG;
    CRN_CUDA(c, cudaSetDevice(c->device));
    const int n = c->nBoards;
    if (!c->havePos0) {                                  // first advection of an uploaded set: keep its base offsets
        int r;
        if ((r = reserve(c, c->pos0, (size_t)std::max(n, 1) * 12))) return r;
        if (n) CRN_CUDA(c, cudaMemcpyAsync(c->pos0.p, c->pos.p, (size_t)n * 12, cudaMemcpyDeviceToDevice, c->lightStream));
        c->havePos0 = true;
    }
    cudaStream_t up = upload_stream(c, false);
    c->launches += launch_rotate_boards(up, n, (const float *)c->pos0.p, (float *)c->pos.p, (float)std::cos(angle),
                                        (float)std::sin(angle));
    CRN_CUDA(c, cudaGetLastError());
    CRN_CUDA(c, cudaEventRecord(c->evBoards, up));
    c->boardsGen++;
    return CRN_OK;
}

int crn_read_billboards(crn_ctx *c, float *positions3_host, float *scales_host) {
    if (!c) return CRN_ERR_INVALID_ARG;
    CRN_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)c->nBoards;
    if (n && positions3_host) CRN_CUDA(c, cudaMemcpyAsync(positions3_host, c->pos.p, n * 12, cudaMemcpyDeviceToHost, c->lightStream));
    if (n && scales_host) CRN_CUDA(c, cudaMemcpyAsync(scales_host, c->scale.p, n * 4, cudaMemcpyDeviceToHost, c->lightStream));
    CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
    return CRN_OK;
}

int crn_set_sun(crn_ctx *c, const crn_sun *sun) {
    if (!c || !sun) return CRN_ERR_INVALID_ARG;
    c->sun = *sun; c->haveSun = true;
    return CRN_OK;
}

int crn_sun_update(const crn_volume_desc *vol, const crn_sun *sun, crn_sun_derived *out) {
    if (!vol || !sun || !out) return CRN_ERR_INVALID_ARG;
    derive_sun(*vol, *sun, out);
    return CRN_OK;
}

int crn_set_camera(crn_ctx *c, const crn_camera *cam) {
    if (!c || !cam) return CRN_ERR_INVALID_ARG;
    const float *P = cam->P;
    const bool persp = P[15] == 0.0f && P[11] == -1.0f, orth = P[15] == 1.0f && P[11] == 0.0f;
    if (!(persp || orth) || P[1] != 0.0f || P[4] != 0.0f || P[8] != 0.0f || P[9] != 0.0f || P[0] == 0.0f || P[5] == 0.0f)
        return fail(c, CRN_ERR_UNSUPPORTED, "P must be a glm::perspective or glm::ortho matrix (no skew, symmetric frustum)");
    c->cam = *cam; c->haveCam = true;
    return CRN_OK;
}

int crn_camera_update(int32_t width, int32_t height, const float eye[3], const float lookAt[3], crn_camera *out) {   // src/Camera.cpp:59-60
    if (!eye || !lookAt || !out || width <= 0 || height <= 0) return CRN_ERR_INVALID_ARG;
    const float aspect = (float)(width / height);             // integer division, as in the reference
    perspective(45.0f, aspect, 0.01f, 2500.0f, out->P);       // 45 *radians* under GLM 0.9.8.5
    const F3 upv = {{0.0f, 1.0f, 0.0f}};
    look_at(f3(eye), f3(lookAt), upv, out->V);
    for (int k = 0; k < 3; k++) out->position[k] = eye[k];
    return CRN_OK;
}

int crn_set_window(crn_ctx *c, int32_t width, int32_t height) {
    if (!c) return CRN_ERR_INVALID_ARG;
    if (width <= 0 || height <= 0 || width > 32767 || height > 32767) return fail(c, CRN_ERR_INVALID_ARG, "window %dx%d out of range", width, height);
    c->W = width; c->H = height; c->haveWindow = true;
    return CRN_OK;
}

int crn_set_trace_params(crn_ctx *c, const crn_trace_params *p) {
    if (!c || !p) return CRN_ERR_INVALID_ARG;
    if (p->vctSteps < 0 || p->vctSteps > kMaxConeSteps) return fail(c, CRN_ERR_UNSUPPORTED, "vctSteps %d > %d", p->vctSteps, kMaxConeSteps);
    if (p->numOctaves < 0 || p->numOctaves > kMaxOctaves) return fail(c, CRN_ERR_UNSUPPORTED, "numOctaves %d > %d", p->numOctaves, kMaxOctaves);
    if (!(p->transmittanceCutoff >= 0.0f && p->transmittanceCutoff < 1.0f)) return fail(c, CRN_ERR_INVALID_ARG, "transmittanceCutoff must be in [0,1)");
    if (p->sampler != CRN_SAMPLER_EXPLICIT && p->sampler != CRN_SAMPLER_TEXTURE) return fail(c, CRN_ERR_INVALID_ARG, "unknown sampler %d", p->sampler);
    c->tp = *p;
    return CRN_OK;
}

int crn_build_noise(const int8_t *alpha, int32_t dim, int8_t *rgba) {      // src/Shaders/ConeTraceShader.cpp:100-151
    if (!alpha || !rgba || dim <= 0) return CRN_ERR_INVALID_ARG;
    auto wrap = [dim](int v) { if (v < 0) v += dim; return v % dim; };
    auto at = [&](int x, int y, int z) { return wrap(x) + wrap(y) * dim + wrap(z) * dim * dim; };
    auto rho = [&](int i) { return (float)alpha[i] / 128.0f; };
    auto pack = [](float f) -> int8_t {
        if (f != f) return 0;
        const float v = fminf(fmaxf(f * 128.0f, -128.0f), 127.0f);
        return (int8_t)(int)v;
    };
    for (int z = 0; z < dim; z++)
        for (int y = 0; y < dim; y++)
            for (int x = 0; x < dim; x++) {
                F3 g;      // precedence quirk kept: only the second density is divided by heightAdjust (0.5)
                g.v[0] = rho(at(x + 1, y, z)) - rho(at(x - 1, y, z)) / 0.5f;
                g.v[1] = rho(at(x, y + 1, z)) - rho(at(x, y - 1, z)) / 0.5f;
                g.v[2] = rho(at(x, y, z + 1)) - rho(at(x, y, z - 1)) / 0.5f;
                const F3 nrm = norm3(g);
                const int i = at(x, y, z);
                rgba[4 * i + 0] = pack(nrm.v[0]); rgba[4 * i + 1] = pack(nrm.v[1]); rgba[4 * i + 2] = pack(nrm.v[2]);
                rgba[4 * i + 3] = alpha[i];
            }
    return CRN_OK;
}

int crn_set_noise(crn_ctx *c, const int8_t *rgba, int32_t dim) {
    if (!c || !rgba || dim <= 0 || dim > 512) return fail(c, CRN_ERR_INVALID_ARG, "bad noise texture");
    CRN_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)dim * dim * dim;
    // the fragment stage reads only .g and .a of the SNORM8 texel: decode them once, here
    std::vector<float> ga(n * 2);
    for (size_t i = 0; i < n; i++) {
        ga[2 * i + 0] = fmaxf((float)rgba[4 * i + 1] / 127.0f, -1.0f);
        ga[2 * i + 1] = fmaxf((float)rgba[4 * i + 3] / 127.0f, -1.0f);
    }
    int r = reserve(c, c->noise, n * 8); if (r) return r;
    CRN_CUDA(c, cudaMemcpyAsync(c->noise.p, ga.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    // texture-unit copy of the GL_RGBA8_SNORM, REPEAT x3, LINEAR, no-mips texture (src/Shaders/ConeTraceShader.cpp:152-158).
    // A 3D LINEAR fetch costs the texture unit two bilinear passes; only .g and .a are ever read, so the texture is
    // stored as a LAYERED 2D array whose layer z holds (g_z, a_z, g_z+1, a_z+1) (z+1 wrapped): one bilinear pass
    // returns both slices, the kernel blends them with the z weight in full float precision.  Same texel bytes, same
    // SNORM8 decode, REPEAT in x and y by the sampler, in z by the layer index.
    sync_all(c);
    if (c->noiseTex) { cudaDestroyTextureObject(c->noiseTex); c->noiseTex = 0; }
    if (c->noiseArray) { cudaFreeArray(c->noiseArray); c->noiseArray = nullptr; }
    if (c->noiseTexD) { cudaDestroyTextureObject(c->noiseTexD); c->noiseTexD = 0; }
    if (c->noiseArrayD) { cudaFreeArray(c->noiseArrayD); c->noiseArrayD = nullptr; }
    std::vector<int8_t> pairs(n * 4);
    for (int z = 0; z < dim; z++) {
        const size_t z0 = (size_t)z * dim * dim, z1 = (size_t)((z + 1) % dim) * dim * dim;
        for (size_t i = 0; i < (size_t)dim * dim; i++) {
            pairs[4 * (z0 + i) + 0] = rgba[4 * (z0 + i) + 1]; pairs[4 * (z0 + i) + 1] = rgba[4 * (z0 + i) + 3];
            pairs[4 * (z0 + i) + 2] = rgba[4 * (z1 + i) + 1]; pairs[4 * (z0 + i) + 3] = rgba[4 * (z1 + i) + 3];
        }
    }
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindSigned);
    CRN_CUDA(c, cudaMalloc3DArray(&c->noiseArray, &cd, make_cudaExtent(dim, dim, dim), cudaArrayLayered));
    cudaMemcpy3DParms cp{};
    cp.srcPtr = make_cudaPitchedPtr((void *)pairs.data(), (size_t)dim * 4, dim, dim);
    cp.dstArray = c->noiseArray; cp.extent = make_cudaExtent(dim, dim, dim); cp.kind = cudaMemcpyHostToDevice;
    CRN_CUDA(c, cudaMemcpy3D(&cp));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = c->noiseArray;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
    CRN_CUDA(c, cudaCreateTextureObject(&c->noiseTex, &rd, &td, nullptr));
    c->vset[0].ts.noise = c->vset[1].ts.noise = c->noiseTex;
    {   // the fast trace variant's copy: RGBA16_SNORM (g_z, a_z, g_z+1 - g_z, a_z+1 - a_z) / 2, so the z blend is one FMA per
        // channel.  Codes round(v * 16383.5) read back as v / 2 (to 1.5e-5); the step is the difference of the CODES, so
        // plane z + its step is exactly plane z+1.
        std::vector<int16_t> d(n * 4);
        auto code = [](int8_t v) { return std::max(-16383, std::min(16383, (int)lrintf(fmaxf((float)v / 127.0f, -1.0f) * 16383.5f))); };   // |step| <= 32766
        for (int z = 0; z < dim; z++) {
            const size_t z0 = (size_t)z * dim * dim, z1 = (size_t)((z + 1) % dim) * dim * dim;
            for (size_t i = 0; i < (size_t)dim * dim; i++) {
                const int g0 = code(rgba[4 * (z0 + i) + 1]), a0 = code(rgba[4 * (z0 + i) + 3]);
                const int g1 = code(rgba[4 * (z1 + i) + 1]), a1 = code(rgba[4 * (z1 + i) + 3]);
                d[4 * (z0 + i) + 0] = (int16_t)g0; d[4 * (z0 + i) + 1] = (int16_t)a0;
                d[4 * (z0 + i) + 2] = (int16_t)(g1 - g0); d[4 * (z0 + i) + 3] = (int16_t)(a1 - a0);
            }
        }
        cudaChannelFormatDesc cdD = cudaCreateChannelDesc(16, 16, 16, 16, cudaChannelFormatKindSigned);
        CRN_CUDA(c, cudaMalloc3DArray(&c->noiseArrayD, &cdD, make_cudaExtent(dim, dim, dim), cudaArrayLayered));
        cudaMemcpy3DParms cpD{};
        cpD.srcPtr = make_cudaPitchedPtr((void *)d.data(), (size_t)dim * 8, dim, dim);
        cpD.dstArray = c->noiseArrayD; cpD.extent = make_cudaExtent(dim, dim, dim); cpD.kind = cudaMemcpyHostToDevice;
        CRN_CUDA(c, cudaMemcpy3D(&cpD));
        cudaResourceDesc rdD{};
        rdD.resType = cudaResourceTypeArray; rdD.res.array.array = c->noiseArrayD;
        CRN_CUDA(c, cudaCreateTextureObject(&c->noiseTexD, &rdD, &td, nullptr));
        c->vset[0].ts.noiseD = c->vset[1].ts.noiseD = c->noiseTexD;
    }
    c->noiseDim = dim; c->haveNoise = true;
    c->noiseGen++;
    return CRN_OK;
}

int crn_voxelize(crn_ctx *c) {
    if (!c) return CRN_ERR_INVALID_ARG;
    int r;
    if ((r = require(c, c->haveVol, "volume")) || (r = require(c, c->haveSun, "sun")) || (r = require(c, c->haveWindow, "window"))) return r;
    CRN_CUDA(c, cudaSetDevice(c->device));
    if ((r = enqueue_voxelize(c))) return r;
    // A Z-slab is about to be handed to an exchange this library does not see.  The first slab after the bin pools were
    // (re)allocated is checked here, once, with a host round trip (grow + re-run if the first guess was too small); every
    // later frame stays asynchronous: a pool that turns out too small then is reported by the next synchronising call
    // (crn_cone_trace, crn_sync, crn_wait_images return CRN_ERR_STATE after growing it; re-submit the frame, exchange included).
    if (c->z1 >= 0 && !(c->z0 == 0 && c->z1 == c->vol.dimension) && c->slabPoolCap != c->binsL.tileCap + c->binsL.coarseCap) {
        if ((r = settle(c, false, 0, true))) return r;
        c->slabPoolCap = c->binsL.tileCap + c->binsL.coarseCap;
    }
    return CRN_OK;
}

// the rows this context owns (row range / tile-row interleave) of `src` -> `out`
static int copy_image(crn_ctx *c, void *out, cudaMemcpyKind kind, int format, const DevBuf &src, cudaStream_t st) {
    const size_t texel = format == CRN_IMAGE_RGBA32F ? 16 : 4;
    const int r0 = std::max(0, c->row0), r1 = std::min(c->H, c->row1);
    const size_t rowB = (size_t)c->W * texel;
    if (c->ilvCount > 1 && (r0 > 0 || r1 < c->H)) {
        // interleave AND a row range: the owned tile rows, each clipped to [row0,row1)
        const int tilesY = (c->H + kTile - 1) / kTile;
        for (int ty = c->ilvIndex; ty < tilesY; ty += c->ilvCount) {
            const int a = std::max(r0, ty * kTile), b = std::min(r1, std::min(c->H, (ty + 1) * kTile));
            if (b > a) CRN_CUDA(c, cudaMemcpyAsync((char *)out + (size_t)a * rowB, (char *)src.p + (size_t)a * rowB, (size_t)(b - a) * rowB, kind, st));
        }
    } else if (c->ilvCount > 1) {
        // only the tile rows this context owns: one strided 2-D copy (+ the partial last tile row)
        const int tilesY = (c->H + kTile - 1) / kTile;
        for (int ty = c->ilvIndex; ty < tilesY;) {
            const int full = (c->H - ty * kTile) / kTile > 0 ? ((c->H / kTile - 1 - ty) / c->ilvCount + 1) : 0;   // owned tile rows of full height
            if (full > 0) {
                const size_t off = (size_t)ty * kTile * rowB, pitch = (size_t)c->ilvCount * kTile * rowB;
                CRN_CUDA(c, cudaMemcpy2DAsync((char *)out + off, pitch, (char *)src.p + off, pitch, (size_t)kTile * rowB, full, kind, st));
                ty += full * c->ilvCount;
            } else {
                const size_t off = (size_t)ty * kTile * rowB;
                CRN_CUDA(c, cudaMemcpyAsync((char *)out + off, (char *)src.p + off, (size_t)(c->H - ty * kTile) * rowB, kind, st));
                ty += c->ilvCount;
            }
        }
    } else if (r1 > r0) {
        const size_t offB = (size_t)r0 * rowB, bytes = (size_t)(r1 - r0) * rowB;
        CRN_CUDA(c, cudaMemcpyAsync((char *)out + offB, (char *)src.p + offB, bytes, kind, st));
    }
    return CRN_OK;
}

// make sure neither pass ran with a truncated bin pool; re-run what did
static int settle(crn_ctx *c, bool haveTrace, int format, bool slabNotShippedYet) {
    for (int attempt = 0; attempt < 4; attempt++) {
        CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
        CRN_CUDA(c, cudaStreamSynchronize(c->auxStream));
        CRN_CUDA(c, cudaStreamSynchronize(c->stream));
        CRN_CUDA(c, cudaStreamSynchronize(c->copyStream));         // the cursors travel on the copy stream
        bool grewL = false, grewC = false;
        int r;
        if (c->voxelized && (r = grow_if_overflowed(c, c->binsL, c->hCursors, &grewL))) return r;
        if (haveTrace && (r = grow_if_overflowed(c, CS(c).binsC, c->hCursors + 4, &grewC))) return r;
        // bit 1 of the flags word: a coarse tile ran out of flush segments (k_bin.cu kMaxSegs: > ~147k rectangles over one
        // 64-pixel tile).  Growing the pools cannot fix that: report it instead of returning an incomplete frame.
        const bool segL = c->voxelized && (c->hCursors[2] & 2u), segC = haveTrace && (c->hCursors[6] & 2u);
        if (segL || segC) {
            if (segL) { CRN_CUDA(c, cudaMemsetAsync(c->binsL.cursors + 2, 0, sizeof(uint32_t), c->lightStream)); c->hCursors[2] = 0; }
            if (segC) { CRN_CUDA(c, cudaMemsetAsync(CS(c).binsC.cursors + 2, 0, sizeof(uint32_t), c->stream)); c->hCursors[6] = 0; }
            return fail(c, CRN_ERR_UNSUPPORTED, "more billboards overlap one 64-pixel tile than the binning pass can stage (%s pass); the frame is incomplete",
                        segL ? "light" : "camera");
        }
        if (grewC) {                                   // keep the other camera set's pools as large: the re-run below uses it
            Bins &o = c->cset[c->cs ^ 1].binsC;
            if ((r = alloc_u32(c, o.coarseList, o.coarseCap, CS(c).binsC.coarseCap))) return r;
            if ((r = alloc_u32(c, o.tileList, o.tileCap, CS(c).binsC.tileCap))) return r;
        }
        if (!grewL && !grewC) return CRN_OK;
        if (grewL && !slabNotShippedYet && c->z1 >= 0 && !(c->z0 == 0 && c->z1 == c->vol.dimension)) {
            // the truncated slab has already been shipped to the other ranks: this library cannot redo the exchange
            CRN_CUDA(c, cudaMemsetAsync(c->binsL.cursors + 2, 0, sizeof(uint32_t), c->lightStream));
            return fail(c, CRN_ERR_STATE, "the light-space bin pool overflowed while voxelizing a Z-slab; the pool has been grown, re-submit the frame (exchange included)");
        }
        // the truncated attempt left the sticky overflow flag behind; this frame is being redone, so clear it
        if (grewL) CRN_CUDA(c, cudaMemsetAsync(c->binsL.cursors + 2, 0, sizeof(uint32_t), c->lightStream));
        if (grewC) CRN_CUDA(c, cudaMemsetAsync(CS(c).binsC.cursors + 2, 0, sizeof(uint32_t), c->stream));
        if (grewL && (r = enqueue_voxelize(c))) return r;
        if (haveTrace && (r = enqueue_trace(c, format))) return r;
    }
    return fail(c, CRN_ERR_STATE, "bin pools did not settle");
}

int crn_cone_trace(crn_ctx *c, void *out, int32_t mem, int32_t format) {
    if (!c || !out) return fail(c, CRN_ERR_INVALID_ARG, "out is NULL");
    if (format != CRN_IMAGE_RGBA8 && format != CRN_IMAGE_RGBA32F) return fail(c, CRN_ERR_INVALID_ARG, "unknown image format %d", format);
    if (mem != CRN_MEM_HOST && mem != CRN_MEM_DEVICE) return fail(c, CRN_ERR_INVALID_ARG, "mem must be CRN_MEM_HOST or CRN_MEM_DEVICE");
    int r;
    if ((r = require(c, c->haveVol, "volume")) || (r = require(c, c->haveSun, "sun")) || (r = require(c, c->haveCam, "camera")) ||
        (r = require(c, c->haveWindow, "window")) || (r = require(c, c->haveNoise, "noise texture")))
        return r;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "crn_voxelize has not produced a volume yet");
    CRN_CUDA(c, cudaSetDevice(c->device));
    if (c->copyPending[0]) { CRN_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopy[0], 0)); c->copyPending[0] = false; }
    if ((r = enqueue_trace(c, format))) return r;
    if ((r = settle(c, true, format))) return r;
    c->traced = true;
    c->imageFormat = format;
    if ((r = copy_image(c, out, mem == CRN_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, format, c->image, c->stream))) return r;
    if (mem == CRN_MEM_HOST) CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    return CRN_OK;
}

int crn_cone_trace_async(crn_ctx *c, void *out, int32_t format) {
    if (!c || !out) return fail(c, CRN_ERR_INVALID_ARG, "out is NULL");
    if (format != CRN_IMAGE_RGBA8 && format != CRN_IMAGE_RGBA32F) return fail(c, CRN_ERR_INVALID_ARG, "unknown image format %d", format);
    int r;
    if ((r = require(c, c->haveVol, "volume")) || (r = require(c, c->haveSun, "sun")) || (r = require(c, c->haveCam, "camera")) ||
        (r = require(c, c->haveWindow, "window")) || (r = require(c, c->haveNoise, "noise texture")))
        return r;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "crn_voxelize has not produced a volume yet");
    CRN_CUDA(c, cudaSetDevice(c->device));
    const int k = c->imgSel;
    DevBuf &img = k ? c->image2 : c->image;
    // the frame two calls ago was copied out of this buffer: the kernels must not overwrite it before that copy is done
    if (c->copyPending[k]) CRN_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopy[k], 0));
    if ((r = enqueue_trace(c, format, &img))) return r;
    c->traced = true;
    CRN_CUDA(c, cudaEventRecord(c->evFrame[k], c->stream));
    CRN_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->evFrame[k], 0));
    if ((r = copy_image(c, out, cudaMemcpyDeviceToHost, format, img, c->copyStream))) return r;
    CRN_CUDA(c, cudaEventRecord(c->evCopy[k], c->copyStream));
    c->copyPending[k] = true;
    c->imgSel ^= 1;
    return CRN_OK;
}

int crn_cone_trace_enqueue(crn_ctx *c, int32_t format) {
    if (!c) return CRN_ERR_INVALID_ARG;
    if (format != CRN_IMAGE_RGBA8 && format != CRN_IMAGE_RGBA32F) return fail(c, CRN_ERR_INVALID_ARG, "unknown image format %d", format);
    int r;
    if ((r = require(c, c->haveVol, "volume")) || (r = require(c, c->haveSun, "sun")) || (r = require(c, c->haveCam, "camera")) ||
        (r = require(c, c->haveWindow, "window")) || (r = require(c, c->haveNoise, "noise texture")))
        return r;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "crn_voxelize has not produced a volume yet");
    CRN_CUDA(c, cudaSetDevice(c->device));
    if (c->copyPending[0]) { CRN_CUDA(c, cudaStreamWaitEvent(c->stream, c->evCopy[0], 0)); c->copyPending[0] = false; }
    if ((r = enqueue_trace(c, format))) return r;
    c->traced = true;
    c->imageFormat = format;
    return CRN_OK;
}

int crn_image_ptr(crn_ctx *c, void **dev_ptr, size_t *bytes) {
    if (!c || !dev_ptr || !bytes) return CRN_ERR_INVALID_ARG;
    if (!c->traced || !c->image.p) return fail(c, CRN_ERR_STATE, "no frame has been traced into the context's image yet");
    *dev_ptr = c->image.p;
    *bytes = (size_t)c->W * c->H * (c->imageFormat == CRN_IMAGE_RGBA32F ? 16 : 4);
    return CRN_OK;
}

int crn_wait_images(crn_ctx *c) {
    if (!c) return CRN_ERR_INVALID_ARG;
    CRN_CUDA(c, cudaSetDevice(c->device));
    if (c->copyStream) CRN_CUDA(c, cudaStreamSynchronize(c->copyStream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    c->copyPending[0] = c->copyPending[1] = false;
    // a frame enqueued without a host round trip may have run with a truncated bin pool: the kernels leave a sticky flag
    bool overflow = false;
    if (c->lightStream) CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
    Bins *bins[3] = {&c->binsL, &c->cset[0].binsC, &c->cset[1].binsC};
    for (int p = 0; p < 3; p++) {
        if (!bins[p]->cursors) continue;
        uint32_t cur[4] = {0, 0, 0, 0};
        CRN_CUDA(c, cudaMemcpy(cur, bins[p]->cursors, sizeof cur, cudaMemcpyDeviceToHost));
        if (cur[2] & 2u) {
            CRN_CUDA(c, cudaMemset(bins[p]->cursors + 2, 0, sizeof(uint32_t)));
            return fail(c, CRN_ERR_UNSUPPORTED, "more billboards overlap one 64-pixel tile than the binning pass can stage; the asynchronous frames are incomplete");
        }
        if (cur[2]) {
            overflow = true;
            CRN_CUDA(c, cudaMemset(bins[p]->cursors + 2, 0, sizeof(uint32_t)));
            int r = alloc_u32(c, bins[p]->coarseList, bins[p]->coarseCap, 2 * std::max<size_t>(bins[p]->coarseCap, (size_t)cur[0] * 4)); if (r) return r;
            r = alloc_u32(c, bins[p]->tileList, bins[p]->tileCap, 2 * std::max<size_t>(bins[p]->tileCap, (size_t)cur[1])); if (r) return r;
        }
    }
    if (overflow) return fail(c, CRN_ERR_STATE, "a bin pool overflowed during asynchronous frames: their images are incomplete; the pools have been grown, re-submit");
    return CRN_OK;
}

int crn_set_row_range(crn_ctx *c, int32_t row0, int32_t row1) {
    if (!c || row0 < 0 || row1 < row0) return fail(c, CRN_ERR_INVALID_ARG, "bad row range");
    c->row0 = row0; c->row1 = row1;
    return CRN_OK;
}

int crn_set_tile_row_interleave(crn_ctx *c, int32_t index, int32_t count) {
    if (!c || count < 1 || index < 0 || index >= count) return fail(c, CRN_ERR_INVALID_ARG, "bad interleave %d of %d", index, count);
    c->ilvIndex = index; c->ilvCount = count;
    return CRN_OK;
}

int crn_set_z_slab(crn_ctx *c, int32_t z0, int32_t z1) {
    if (!c) return CRN_ERR_INVALID_ARG;
    int r = require(c, c->haveVol, "volume"); if (r) return r;
    if (z0 < 0 || z1 <= z0 || z1 > c->vol.dimension || (z0 % 16) || (z1 % 16))
        return fail(c, CRN_ERR_INVALID_ARG, "slab [%d,%d) must be non-empty, inside the volume and aligned to 16 slices", z0, z1);
    c->z0 = z0; c->z1 = z1;
    return CRN_OK;
}

int crn_volume_level_ptr(crn_ctx *c, int32_t level, void **dev_ptr, size_t *bytes) {
    if (!c || !dev_ptr || !bytes) return CRN_ERR_INVALID_ARG;
    int r = require(c, c->haveVol, "volume"); if (r) return r;
    if (level < 0 || level >= c->vol.levels) return fail(c, CRN_ERR_INVALID_ARG, "level %d out of range", level);
    fill_vparams(c);
    if ((r = reserve(c, c->chain, c->chainBytes))) return r;
    const size_t s = c->vparams.levelSize[level];
    *dev_ptr = (char *)c->chain.p + c->vparams.levelOff[level];
    *bytes = s * s * s * c->vparams.texelBytes;
    c->levelExposed = true;             // the caller consumes / produces chain levels on its stream: strict stream order from now on
    if (level == 0) {                   // from now on every voxelize keeps the linear level 0 current
        c->wantLinear0 = true;
        if (c->voxelized && !c->linear0Valid) {
            CRN_CUDA(c, cudaSetDevice(c->device));
            c->launches += launch_expand_level0(c->stream, c->vparams, (const uint32_t *)c->bits.p, (uint8_t *)c->chain.p, nullptr);
            c->linear0Valid = true;
        }
    }
    return CRN_OK;
}

int crn_volume_bits_ptr(crn_ctx *c, void **dev_ptr, size_t *bytes) {
    if (!c || !dev_ptr || !bytes) return CRN_ERR_INVALID_ARG;
    int r = require(c, c->haveVol, "volume"); if (r) return r;
    const size_t D = c->vol.dimension;
    if ((r = reserve(c, c->bits, D * D * D / 8))) return r;
    *dev_ptr = c->bits.p; *bytes = D * D * D / 8;
    c->bitsExposed = true;
    return CRN_OK;
}

int crn_finish_mips(crn_ctx *c, int32_t first_level) {
    if (!c) return CRN_ERR_INVALID_ARG;
    int r = require(c, c->haveVol, "volume"); if (r) return r;
    if (first_level < 1 || first_level > c->vol.levels) return fail(c, CRN_ERR_INVALID_ARG, "first_level %d out of range", first_level);
    CRN_CUDA(c, cudaSetDevice(c->device));
    fill_vparams(c);
    CRN_CUDA(c, cudaStreamWaitEvent(c->stream, c->evVoxDone, 0));       // the chain is produced on the light stream
    c->levelExposed = true;                                             // ... and from here on consumed in the caller's order
    c->launches += launch_finish_mips(c->stream, c->vparams, (uint8_t *)c->chain.p, first_level);
    CRN_CUDA(c, cudaGetLastError());
    c->voxelized = true;
    VS(c).texCurrent = false;              // the texture-unit copy is refreshed from the chain at the next trace
    c->maskCurrent = false;
    c->volumeGen++;
    return CRN_OK;
}

int crn_read_volume(crn_ctx *c, int32_t level, void *dst) {
    if (!c || !dst) return CRN_ERR_INVALID_ARG;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "no volume has been produced yet");
    if (level < 0 || level >= c->vol.levels) return fail(c, CRN_ERR_INVALID_ARG, "level %d out of range", level);
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r = settle(c, false, 0); if (r) return r;
    if (level == 0 && !c->linear0Valid) {
        c->launches += launch_expand_level0(c->stream, c->vparams, (const uint32_t *)c->bits.p, (uint8_t *)c->chain.p, nullptr);
        c->linear0Valid = true;
    }
    const size_t s = c->vparams.levelSize[level];
    CRN_CUDA(c, cudaMemcpyAsync(dst, (char *)c->chain.p + c->vparams.levelOff[level], s * s * s * c->vparams.texelBytes, cudaMemcpyDeviceToHost,
                                c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    return CRN_OK;
}

int crn_read_volume_alpha(crn_ctx *c, int32_t level, void *dst) {
    if (!c || !dst) return CRN_ERR_INVALID_ARG;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "no volume has been produced yet");
    if (c->vol.format != CRN_VOLUME_RG8) return fail(c, CRN_ERR_STATE, "the volume has no occupancy channel (format is not CRN_VOLUME_RG8)");
    if (level < 0 || level >= c->vol.levels) return fail(c, CRN_ERR_INVALID_ARG, "level %d out of range", level);
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r = settle(c, false, 0); if (r) return r;
    if (level == 0 && !c->linear0ValidA) {
        c->launches += launch_expand_level0(c->stream, c->vparams, (const uint32_t *)c->bitsA.p, (uint8_t *)c->chainA.p, nullptr);
        c->linear0ValidA = true;
    }
    const size_t s = c->vparams.levelSize[level];
    CRN_CUDA(c, cudaMemcpyAsync(dst, (char *)c->chainA.p + c->vparams.levelOff[level], s * s * s, cudaMemcpyDeviceToHost, c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    return CRN_OK;
}

int crn_count_active_voxels(crn_ctx *c, uint64_t *count) {
    if (!c || !count) return CRN_ERR_INVALID_ARG;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "no volume has been produced yet");
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r = settle(c, false, 0); if (r) return r;
    const size_t D = c->vol.dimension;
    unsigned long long *d = (unsigned long long *)((char *)c->misc.p + 128);
    c->launches += launch_count_bits(c->stream, (const uint32_t *)c->bits.p, D * D * D / 32, d);
    unsigned long long h = 0;
    CRN_CUDA(c, cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    *count = h;
    return CRN_OK;
}

int crn_export_voxels(crn_ctx *c, int32_t channel, float *dst_host_xyzw, uint64_t capacity, uint64_t *count) {
    if (!c || !count || (capacity && !dst_host_xyzw)) return CRN_ERR_INVALID_ARG;
    if (!c->voxelized) return fail(c, CRN_ERR_STATE, "no volume has been produced yet");
    if (channel != 0 && channel != 1) return fail(c, CRN_ERR_INVALID_ARG, "channel must be 0 (lit) or 1 (occupancy)");
    if (channel == 1 && c->vol.format != CRN_VOLUME_RG8) return fail(c, CRN_ERR_STATE, "the volume has no occupancy channel (format is not CRN_VOLUME_RG8)");
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r = settle(c, false, 0); if (r) return r;
    const size_t D = c->vol.dimension, words = D * D * D / 32;
    if ((r = reserve(c, c->exportTmp, export_scratch_words(words) * 4 + 64))) return r;
    if (capacity && (r = reserve(c, c->exportOut, (size_t)capacity * 16))) return r;
    unsigned long long *dTotal = (unsigned long long *)((char *)c->exportTmp.p + (export_scratch_words(words) * 4 + 15) / 16 * 16);
    const float lo[3] = {c->vol.xBounds[0], c->vol.yBounds[0], c->vol.zBounds[0]};
    const float range[3] = {c->vol.xBounds[1] - c->vol.xBounds[0], c->vol.yBounds[1] - c->vol.yBounds[0], c->vol.zBounds[1] - c->vol.zBounds[0]};
    const uint32_t *which = (const uint32_t *)(channel == 1 ? c->bitsA.p : c->bits.p);
    c->launches += launch_export_voxels(c->stream, which, (const uint32_t *)c->bits.p, words, (int)D, c->vol.position, lo, range,
                                        (uint32_t *)c->exportTmp.p, dTotal, capacity ? (float4 *)c->exportOut.p : nullptr, capacity);
    CRN_CUDA(c, cudaGetLastError());
    unsigned long long h = 0;
    CRN_CUDA(c, cudaMemcpyAsync(&h, dTotal, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    *count = h;
    const uint64_t n = std::min<uint64_t>(h, capacity);
    if (n) {
        CRN_CUDA(c, cudaMemcpyAsync(dst_host_xyzw, c->exportOut.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
        CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return CRN_OK;
}

int crn_keep_position_map(crn_ctx *c, int32_t enable) {
    if (!c) return CRN_ERR_INVALID_ARG;
    c->keepPosmap = enable != 0;
    return CRN_OK;
}

int crn_read_position_map(crn_ctx *c, float *dst) {
    if (!c || !dst) return CRN_ERR_INVALID_ARG;
    if (!c->keepPosmap || !c->voxelized || !c->posmap.p) return fail(c, CRN_ERR_STATE, "enable crn_keep_position_map before crn_voxelize");
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r = settle(c, false, 0); if (r) return r;
    CRN_CUDA(c, cudaMemcpyAsync(dst, c->posmap.p, (size_t)c->W * c->H * 16, cudaMemcpyDeviceToHost, c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    return CRN_OK;
}

int crn_read_sorted_order(crn_ctx *c, int32_t *dst) {
    if (!c || !dst) return CRN_ERR_INVALID_ARG;
    if (!c->traced) return fail(c, CRN_ERR_STATE, "crn_cone_trace has not run yet");
    CRN_CUDA(c, cudaSetDevice(c->device));
    if (c->nBoards) CRN_CUDA(c, cudaMemcpyAsync(dst, c->drawOrder.p, (size_t)c->nBoards * 4, cudaMemcpyDeviceToHost, c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    return CRN_OK;
}

int crn_read_bins(crn_ctx *c, int32_t which, int32_t *tiles_x, int32_t *tiles_y, int32_t *tile_w, int32_t *tile_h,
                  int32_t *counts, int32_t *entries, uint64_t *total) {
    if (!c || (which != 0 && which != 1)) return CRN_ERR_INVALID_ARG;
    if ((which == 0 && !c->voxelized) || (which == 1 && !c->traced)) return fail(c, CRN_ERR_STATE, "that pass has not run yet");
    CRN_CUDA(c, cudaSetDevice(c->device));
    int r = settle(c, false, 0); if (r) return r;
    const Bins &b = which == 0 ? c->binsL : CS(c).binsC;
    const DevBuf &recs = which == 0 ? c->recL : CS(c).recC;
    const size_t tiles = (size_t)b.tilesX * b.tilesY;
    if (tiles_x) *tiles_x = b.tilesX;
    if (tiles_y) *tiles_y = b.tilesY;
    if (tile_w) *tile_w = kTile;
    if (tile_h) *tile_h = kTile;
    std::vector<uint32_t> cnt(tiles), off(tiles);
    CRN_CUDA(c, cudaMemcpy(cnt.data(), b.tileCnt, tiles * 4, cudaMemcpyDeviceToHost));
    CRN_CUDA(c, cudaMemcpy(off.data(), b.tileOff, tiles * 4, cudaMemcpyDeviceToHost));
    uint64_t tot = 0;
    for (size_t t = 0; t < tiles; t++) tot += cnt[t];
    if (total) *total = tot;
    if (counts) for (size_t t = 0; t < tiles; t++) counts[t] = (int32_t)cnt[t];
    if (entries && tot) {
        const uint32_t cur = c->hCursors[which == 0 ? 1 : 5];
        std::vector<uint32_t> list(cur);
        CRN_CUDA(c, cudaMemcpy(list.data(), b.tileList, (size_t)cur * 4, cudaMemcpyDeviceToHost));
        std::vector<BoardRec> rec(std::max(c->nBoards, 1));
        CRN_CUDA(c, cudaMemcpy(rec.data(), recs.p, (size_t)c->nBoards * sizeof(BoardRec), cudaMemcpyDeviceToHost));
        size_t o = 0;
        for (size_t t = 0; t < tiles; t++)
            for (uint32_t e = 0; e < cnt[t]; e++) entries[o++] = rec[list[off[t] + e]].idx & kRecIndexMask;
    }
    return CRN_OK;
}

int crn_get_trace_stats(crn_ctx *c, crn_trace_stats *out) {
    if (!c || !out) return CRN_ERR_INVALID_ARG;
    if (!c->traced || !c->statsOn) return fail(c, CRN_ERR_STATE, "enable crn_set_stats before crn_cone_trace");
    CRN_CUDA(c, cudaSetDevice(c->device));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    CRN_CUDA(c, cudaStreamSynchronize(c->copyStream));
    out->fragments = c->hStats[0]; out->coneSamples = c->hStats[1]; out->noiseSamples = c->hStats[2];
    out->binEntries = c->hCursors[5];
    out->coneSamplesSkipped = c->hStats[3];
    out->bakedFetches = c->hStats[5];
    out->noiseLatticeSteps = c->hStats[6]; out->codeLookups = c->hStats[7];
    out->filteredFetches = c->hStats[4] + c->hStats[5] + c->hStats[2];      // textureLod cone fetches + baked cone fetches + noise taps
    return CRN_OK;
}

int crn_set_stats(crn_ctx *c, int32_t enable) {
    if (!c) return CRN_ERR_INVALID_ARG;
    c->statsOn = enable != 0;
    return CRN_OK;
}

int crn_set_timing(crn_ctx *c, int32_t enable) {
    if (!c) return CRN_ERR_INVALID_ARG;
    c->timingOn = enable != 0;
    if (!c->timingOn) c->evVValid = c->evTValid = false;
    return CRN_OK;
}

int crn_get_timings(crn_ctx *c, crn_timings *out) {
    if (!c || !out) return CRN_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof *out);
    CRN_CUDA(c, cudaSetDevice(c->device));
    CRN_CUDA(c, cudaStreamSynchronize(c->lightStream));
    CRN_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->evVValid) {
        CRN_CUDA(c, cudaEventElapsedTime(&out->prepSortMs, c->evV[0], c->evV[1]));
        CRN_CUDA(c, cudaEventElapsedTime(&out->lightBinMs, c->evV[1], c->evV[2]));
        CRN_CUDA(c, cudaEventElapsedTime(&out->voxelizeMs, c->evV[2], c->evV[3]));
        CRN_CUDA(c, cudaEventElapsedTime(&out->mipMs, c->evV[5], c->evV[4]));
    }
    if (c->evTValid) {
        float t = 0;
        CRN_CUDA(c, cudaStreamSynchronize(c->auxStream));
        CRN_CUDA(c, cudaEventElapsedTime(&t, c->evAuxT[0], c->evAuxT[1]));        // side stream: overlaps the voxelize stage
        out->prepSortMs += t;
        CRN_CUDA(c, cudaEventElapsedTime(&out->camBinMs, c->evAuxT[1], c->evAuxT[2]));
        float t2 = 0;
        CRN_CUDA(c, cudaEventElapsedTime(&out->coneAccelMs, c->evT[2], c->evT[1]));          // baked steps
        CRN_CUDA(c, cudaEventElapsedTime(&t2, c->evAccT[0], c->evAccT[1]));                   // need codes
        out->coneAccelMs += t2;
        CRN_CUDA(c, cudaEventElapsedTime(&out->traceMs, c->evT[0], c->evT[3]));
    }
    return CRN_OK;
}

int crn_get_launch_count(crn_ctx *c, uint64_t *count) {
    if (!c || !count) return CRN_ERR_INVALID_ARG;
    *count = c->launches;
    return CRN_OK;
}

} // extern "C"
