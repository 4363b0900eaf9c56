/*
 * cloud_renderer_b200.h — C-ABI boundary of the B200-native Cloud-Renderer hot path.
 *
 * The reference (jaafersheriff/Cloud-Renderer) has no plugin/FFI layer: its boundary is
 * four C++ call sites per frame in src/main.cpp:106-123
 *     Sun::update(volume); volume->update(); voxelizeShader->voxelize(volume);
 *     coneShader->coneTrace(volume);
 * plus construction (src/main.cpp:81-88) and a set of global statics (Sun::*, Camera::*,
 * Window::{width,height,runTime}) and public shader-object fields.  This header is the
 * plain-C replacement for exactly that surface: POD structs with the reference's field
 * names, an opaque context that owns all device memory, integer status codes instead of
 * exit(), and explicit stream ordering.  No torch / C++ types appear in any signature.
 *
 * Matrices are column-major float[16] exactly as glm::value_ptr() yields them
 * (reference upload path: src/Shaders/Shader.cpp:208-210), i.e. m[col*4+row].
 *
 * Images follow the GL window convention the reference renders into: row 0 is the
 * BOTTOM row (gl_FragCoord.y = 0.5), x grows to the right, 4 channels RGBA.
 * Volumes are x-fastest linear arrays (src/CloudVolume.cpp:103-110).
 */
#ifndef CLOUD_RENDERER_B200_H
#define CLOUD_RENDERER_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (replace exitError(), src/main.cpp:63-67) ---- */
#define CRN_OK               0
#define CRN_ERR_INVALID_ARG  1
#define CRN_ERR_CUDA         2
#define CRN_ERR_STATE        3   /* a required input (volume, billboards, camera...) was never set */
#define CRN_ERR_UNSUPPORTED  4
#define CRN_ERR_NO_DEVICE    5

/* where a caller-provided buffer lives */
#define CRN_MEM_HOST    0
#define CRN_MEM_DEVICE  1

/* output image formats for crn_cone_trace */
#define CRN_IMAGE_RGBA8    0   /* what the reference's window framebuffer holds */
#define CRN_IMAGE_RGBA32F  1   /* un-quantised accumulator, for parity measurement */

/* how crn_cone_trace filters the volume chain and the noise texture */
#define CRN_SAMPLER_EXPLICIT 0  /* in-kernel trilinear / mip-linear in full float precision     */
#define CRN_SAMPLER_TEXTURE  1  /* B200 texture units (8-bit filter weights, like any GL GPU)   */

/* volume texel formats. R8 is the shipped reference format (src/CloudVolume.cpp:18). */
#define CRN_VOLUME_R8      0
#define CRN_VOLUME_R32F    1   /* float texels: level 0 = 0.0/1.0, mips = plain 2x2x2 means (no re-quantisation) */
#define CRN_VOLUME_RG8     2   /* paper variant (opt-in): channel 0 = the lit shell (the paper's rgb), channel 1 = its
                                * alpha = every voxel inside a billboard sphere (first-pass interior march,
                                * res/first_voxelize.glsl:53-58 as in paper/tex/voxelization.tex:13-27) | lit;
                                * two planar R8 chains; the cone sum is gated by alpha > 0
                                * (paper/tex/conetracing.tex:36-39) */

typedef struct crn_ctx crn_ctx;

/* CloudVolume's public parameter fields (src/CloudVolume.hpp:33-50). Bounds are relative
 * to `position`, as in the reference; `fluffiness` scales every billboard radius at
 * upload time (src/CloudVolume.cpp:153-161). */
typedef struct crn_volume_desc {
    int32_t dimension;      /* voxels per axis (D)                               */
    int32_t levels;         /* mip levels, 1 <= levels <= log2(D)+1               */
    float   position[3];
    float   xBounds[2];
    float   yBounds[2];
    float   zBounds[2];
    float   fluffiness;
    int32_t format;         /* CRN_VOLUME_*                                        */
} crn_volume_desc;

/* Sun's user-set statics (src/Sun.hpp:12-24, defaults src/main.cpp:37-43). */
typedef struct crn_sun {
    float position[3];
    float innerColor[3];
    float outerColor[3];
    float innerRadius;
    float outerRadius;
} crn_sun;

/* What Sun::update() derives (src/Sun.hpp:26-43). */
typedef struct crn_sun_derived {
    float P[16];
    float V[16];
    float nearPlane[3];
    float farPlane[3];
    float clipDistance;
} crn_sun_derived;

/* Camera::getP()/getV()/getPosition() (src/Camera.hpp:25-29). */
typedef struct crn_camera {
    float P[16];
    float V[16];
    float position[3];
} crn_camera;

/* Every public field of ConeTraceShader (src/Shaders/ConeTraceShader.hpp:15-36) plus
 * Window::runTime (wind animation, src/Shaders/ConeTraceShader.cpp:55-61) and the frame
 * state the reference sets around the pass (clear colour src/main.cpp:112, sun pass
 * src/main.cpp:116).  The last block are extensions with reference-neutral defaults. */
typedef struct crn_trace_params {
    float   stepSize;
    float   noiseOpacity;
    int32_t numOctaves;
    float   freqStep;
    float   persStep;
    float   adjustSize;
    int32_t minNoiseSteps;
    int32_t maxNoiseSteps;
    float   minNoiseColor;
    float   noiseColorScale;
    float   windVel[3];

    int32_t vctSteps;
    float   vctConeAngle;
    float   vctConeInitialHeight;
    float   vctLodOffset;
    float   vctDownScaling;

    int32_t showQuad;
    int32_t doConeTrace;
    int32_t doNoiseSample;

    float   runTime;

    float   clearColor[4];          /* (0.2,0.3,0.5,1) in the reference              */
    int32_t drawSun;                /* run the sun-disc pass before the cloud         */
    float   transmittanceCutoff;    /* early ray termination: stop shading a pixel once
                                       the transmittance of everything in front of the
                                       next billboard is below this. 0 = shade every
                                       fragment (reference behaviour).                */
    int32_t sampler;                /* CRN_SAMPLER_*                                   */
    int32_t skipEmptySpace;         /* skip cone samples whose whole filter footprint is
                                       provably zero (exact: they contribute 0). 1 = on  */
    int32_t quantizeFramebuffer;    /* 1: emulate the reference's 8-bit window framebuffer (src/main.cpp:94-95): the
                                       billboards are blended BACK TO FRONT and every blend result is written back
                                       through 8 bits per channel (no early ray termination in this mode).  0 (default):
                                       float accumulation, quantised once at the end.                                  */
} crn_trace_params;

/* Counters of one crn_cone_trace call (read back on demand). */
typedef struct crn_trace_stats {
    uint64_t fragments;        /* fragments that survived discard and were shaded      */
    uint64_t coneSamples;      /* textureLod() taps taken by traceCone                  */
    uint64_t noiseSamples;     /* texture(noiseMap) taps taken by noise3D               */
    uint64_t binEntries;       /* (tile, billboard) pairs in the camera-space bins      */
    uint64_t coneSamplesSkipped; /* of coneSamples: proven zero by the empty-space masks, not fetched */
    uint64_t filteredFetches;  /* filtered lookups actually issued: noise taps (bilinear, slice pairs) + 1 or 2 trilinear per fetched textureLod cone sample + bakedFetches */
    uint64_t bakedFetches;     /* of filteredFetches: cone samples served by a baked step texture (one bilinear pass each) */
    uint64_t noiseLatticeSteps; /* noise march steps of billboards inside the combined-octave noise lattice: the fast trace variant takes 2 lookups there (octave 0 + the pre-summed octaves) instead of numOctaves */
    uint64_t codeLookups;      /* fragments whose need code the fast trace variant fetches through the texture unit (one point-sampled pass) */
} crn_trace_stats;

/* Stage timings of the most recent frame, milliseconds, measured with CUDA events on
 * the context's stream (only recorded when crn_set_timing(ctx, 1)). */
typedef struct crn_timings {
    float prepSortMs;      /* per-billboard set-up + both sorts                        */
    float lightBinMs;      /* light-space binning                                      */
    float voxelizeMs;      /* clear + pass 1 + pass 2 (position map -> occupancy)      */
    float mipMs;           /* level-0 expand + all mip levels                          */
    float camBinMs;        /* camera-space binning                                     */
    float traceMs;         /* cone-trace kernel                                        */
    float coneAccelMs;     /* per-frame cone acceleration data (baked step textures, need codes), part of the trace stage */
} crn_timings;

/* ---- lifetime --------------------------------------------------------------------- */
/* `stream` is a cudaStream_t (or NULL for a private non-blocking stream). One ctx per
 * (device, stream); calls on a ctx are serialised by the caller. */
int  crn_create(int device, void *stream, crn_ctx **out);   /* note: every crn_* call leaves `device` current (cudaSetDevice) on the calling thread */
void crn_destroy(crn_ctx *ctx);
/* message for the last non-OK status on this ctx (ctx may be NULL: last create error) */
const char *crn_last_error(const crn_ctx *ctx);
int  crn_sync(crn_ctx *ctx);

/* ---- parameter surface ------------------------------------------------------------- */
/* replaces CloudVolume::CloudVolume(dim, bounds, position, mips) (src/CloudVolume.cpp:7-55)
 * and the per-frame CloudVolume::update() (src/CloudVolume.cpp:84-93) */
int crn_set_volume(crn_ctx *ctx, const crn_volume_desc *desc);
/* replaces CloudVolume::uploadBillboards() (src/CloudVolume.cpp:139-164): positions are
 * offsets relative to the volume position, scales are un-fluffed radii; array order is
 * the instance order of the voxelize draw. Data is copied asynchronously: a CRN_MEM_DEVICE
 * source is read in the order of the context's stream (after whatever the caller queued
 * there to produce it); a CRN_MEM_HOST source is read on an internal stream so that the
 * upload of the next frame does not wait for the trace of the current one - a pinned host
 * array must stay unchanged until the crn_voxelize that follows has completed (crn_sync,
 * or a synchronisation of the context's stream after that crn_voxelize); pageable memory
 * is staged before the call returns. */
int crn_set_billboards(crn_ctx *ctx, const float *positions3, const float *scales,
                       int32_t count, int32_t mem);
/* replaces the Sun statics + Sun::update(volume) (src/Sun.hpp:26-43); the derived light
 * camera is recomputed from the current volume at every crn_voxelize. */
int crn_set_sun(crn_ctx *ctx, const crn_sun *sun);
/* Device-side CloudVolume::regenerateBillboards (src/CloudVolume.cpp:120-137): `count`
 * billboards with offsets uniform in [minOffset, maxOffset) per axis and scales uniform in
 * [minScale, maxScale) times `radiusFactor` (1 = the reference's distribution), written
 * straight into the context's arrays - nothing crosses PCIe.  The reference's rand() stream
 * (seeded with time(0), src/main.cpp:70) is not reproducible; the stream here is
 * counter-based splitmix64 of `seed` (billboard i uses counters 4i+1..4i+4; float64
 * u*(max-min)+min rounded to float32), bit-identical to cloud-renderer_b200/scene.py. */
int crn_regenerate_billboards(crn_ctx *ctx, int32_t count, const float minOffset[3],
                              const float maxOffset[3], float minScale, float maxScale,
                              double radiusFactor, uint64_t seed);
/* Advect the current billboard set on the device: offsets = R_y(angle) * base offsets, where
 * the base is what crn_regenerate_billboards / crn_set_billboards last installed (absolute
 * angle, not incremental; float32 c*x+s*z, -s*x+c*z with c,s = (float)cos/sin(angle)).
 * This is the analytic animation field of the benchmark configs; it replaces the per-frame
 * host update + glBufferData re-upload of src/CloudVolume.cpp:139-164. */
int crn_animate_billboards(crn_ctx *ctx, double angle);
/* current billboard arrays (after generation / advection); either pointer may be NULL */
int crn_read_billboards(crn_ctx *ctx, float *positions3_host, float *scales_host);

/* the pure function behind Sun::update, exported so callers can inspect it */
int crn_sun_update(const crn_volume_desc *vol, const crn_sun *sun, crn_sun_derived *out);
/* replaces Camera::getP()/getV()/getPosition() as read by ConeTraceShader::coneTrace
 * (src/Shaders/ConeTraceShader.cpp:20,64-69) */
int crn_set_camera(crn_ctx *ctx, const crn_camera *cam);
/* the pure function behind Camera::update's matrix block (src/Camera.cpp:59-60),
 * including its integer-division aspect and radians-for-degrees fov */
int crn_camera_update(int32_t width, int32_t height, const float eye[3],
                      const float lookAt[3], crn_camera *out);
/* replaces Window::width/height as read by VoxelizeShader (position-map size,
 * src/Shaders/VoxelizeShader.cpp:15,20-22) and by the rasteriser (image size) */
int crn_set_window(crn_ctx *ctx, int32_t width, int32_t height);
int crn_set_trace_params(crn_ctx *ctx, const crn_trace_params *params);
/* fills `params` with the reference defaults (src/Shaders/ConeTraceShader.hpp:15-36) */
void crn_default_trace_params(crn_trace_params *params);
/* replaces ConeTraceShader::initNoiseMap's upload (src/Shaders/ConeTraceShader.cpp:127-159):
 * dim^3 RGBA8_SNORM texels, x-fastest, REPEAT + LINEAR */
int crn_set_noise(crn_ctx *ctx, const int8_t *rgba, int32_t dim);
/* the deterministic half of initNoiseMap (normals from the alpha/density channel);
 * alpha[dim^3] in, rgba[4*dim^3] out.  Host-side helper, no GPU. */
int crn_build_noise(const int8_t *alpha, int32_t dim, int8_t *rgba_out);

/* ---- the two passes ---------------------------------------------------------------- */
/* replaces VoxelizeShader::voxelize(CloudVolume*) (src/Shaders/VoxelizeShader.cpp:18-31):
 * clear, pass 1 (position map), pass 2 (scatter), mip chain.  Stream-ordered. */
int crn_voxelize(crn_ctx *ctx);
/* replaces ConeTraceShader::coneTrace(CloudVolume*) (src/Shaders/ConeTraceShader.cpp:15-82)
 * including the clear and (optionally) the sun pass that precede it in the frame
 * (src/main.cpp:112-116).  `out` receives width*height*4 texels of `format`.
 * With mem == CRN_MEM_HOST the call returns after the copy has completed. */
int crn_cone_trace(crn_ctx *ctx, void *out, int32_t mem, int32_t format);

/* Pipelined read-back: like crn_cone_trace(ctx, out_host, CRN_MEM_HOST, format) but returns as soon as the
 * frame and its device->host copy are enqueued.  The copy runs on a second stream and overlaps the next frame's
 * kernels (the context double-buffers the device image), so up to two frames may be in flight.  `out_host` must
 * stay valid — and should be pinned — until crn_wait_images() returns; that call also reports a bin-pool overflow
 * of any frame since the last wait (CRN_ERR_STATE: the pools have been grown, re-submit those frames). */
int crn_cone_trace_async(crn_ctx *ctx, void *out_host, int32_t format);
int crn_wait_images(crn_ctx *ctx);
/* Device-resident counterpart of crn_cone_trace_async: enqueue the frame and leave the image in the context's own
 * device buffer (the reference leaves it in the window framebuffer, src/main.cpp:123).  Nothing is copied and the host
 * is not synchronised; consume the image in stream order through crn_image_ptr, and call crn_sync / crn_wait_images
 * before trusting it (they report a bin pool that overflowed, as for the asynchronous path). */
int crn_cone_trace_enqueue(crn_ctx *ctx, int32_t format);
/* device address + size of the image written by the most recent crn_cone_trace / crn_cone_trace_enqueue (row 0 = bottom) */
int crn_image_ptr(crn_ctx *ctx, void **dev_ptr, size_t *bytes);

/* ---- sharding hooks (multi-GPU; results are invariant to them) --------------------- */
/* restrict crn_cone_trace to image rows [row0,row1); other rows of `out` are untouched */
int crn_set_row_range(crn_ctx *ctx, int32_t row0, int32_t row1);
/* load-balanced image-space sharding: this context traces only the 16-pixel tile rows ty with
 * ty % count == index (count = 1 turns it off); combines with crn_set_row_range */
int crn_set_tile_row_interleave(crn_ctx *ctx, int32_t index, int32_t count);
/* restrict crn_voxelize to voxel slices z in [z0,z1): only those slices of every
 * slab-local level are produced (levels whose texel spans more than the slab are left
 * for crn_finish_mips after the exchange).  crn_voxelize stays asynchronous: the exchange is
 * ordered after it on the context's stream (NCCL / peer copies enqueued on that stream), and
 * what has to travel is the occupancy bits (crn_volume_bits_ptr) plus chain levels 1.. of the
 * slab-local range — level 0 is re-expanded from the bits on every rank.  A bin pool that was
 * too small for the slab is reported by the next synchronising call (CRN_ERR_STATE; the pool has
 * been grown, re-submit the frame including the exchange). */
int crn_set_z_slab(crn_ctx *ctx, int32_t z0, int32_t z1);
/* device address + byte size of one level of the chain (for an external all-gather); written in stream order by
 * crn_voxelize.  The linear copy of level 0 (8x the bits) is produced only for a caller that has asked for it:
 * taking its address here switches that on for every later crn_voxelize (crn_read_volume expands it on demand). */
int crn_volume_level_ptr(crn_ctx *ctx, int32_t level, void **dev_ptr, size_t *bytes);
/* device address + size of the level-0 occupancy bitset (1 bit per voxel, x-fastest) */
int crn_volume_bits_ptr(crn_ctx *ctx, void **dev_ptr, size_t *bytes);
/* recompute levels [first_level, levels) from level first_level-1 (after a gather) */
int crn_finish_mips(crn_ctx *ctx, int32_t first_level);

/* ---- inspection (the reference's debug views: src/Shaders/VoxelShader.cpp:102-133,
 *      src/main.cpp:172-198) ---------------------------------------------------------- */
int crn_read_volume(crn_ctx *ctx, int32_t level, void *dst_host);          /* size_l^3 texels (uint8 or float) */
/* CRN_VOLUME_RG8 only: one level of the occupancy (alpha) channel, s^3 bytes, same order as
 * crn_read_volume (what the reference's debug voxel view draws as black cubes) */
int crn_read_volume_alpha(crn_ctx *ctx, int32_t level, void *dst_host);
int crn_count_active_voxels(crn_ctx *ctx, uint64_t *count);                  /* "Voxels in scene" */
/* VoxelShader::updateVoxelData (src/Shaders/VoxelShader.cpp:102-133): the cubes of the debug
 * voxel view.  For every non-empty level-0 texel, in ascending linear index (x fastest), one
 * float4 = (position + reverseVoxelIndex(x,y,z) [src/CloudVolume.cpp:112-118], value).
 * channel 0: lit voxels (value 1).  channel 1 (CRN_VOLUME_RG8): occupied voxels, value 1 where
 * lit else 0 (the paper's black interior cubes).  *count receives the total; at most
 * `capacity` entries are written (call with capacity 0 to size the buffer). */
int crn_export_voxels(crn_ctx *ctx, int32_t channel, float *dst_host_xyzw, uint64_t capacity,
                      uint64_t *count);
int crn_keep_position_map(crn_ctx *ctx, int32_t enable);                    /* default off   */
int crn_read_position_map(crn_ctx *ctx, float *dst_host_rgba32f);           /* W*H*4 floats  */
/* draw order of the last crn_cone_trace: indices into the billboard arrays, far -> near
 * (what CloudVolume::sortBoards leaves in place, src/CloudVolume.cpp:65-82) */
int crn_read_sorted_order(crn_ctx *ctx, int32_t *dst_host);
/* bins of the last pass: which = 0 light space (voxelize), 1 camera space (trace).
 * counts[tiles_x*tiles_y] (row-major, tile row 0 at the bottom); entries are billboard
 * indices concatenated tile by tile in list order.  Pass NULL to query sizes only. */
int crn_read_bins(crn_ctx *ctx, int32_t which, int32_t *tiles_x, int32_t *tiles_y,
                  int32_t *tile_w, int32_t *tile_h, int32_t *counts_host,
                  int32_t *entries_host, uint64_t *total_entries);
int crn_get_trace_stats(crn_ctx *ctx, crn_trace_stats *out);
int crn_set_stats(crn_ctx *ctx, int32_t enable);                             /* default off */
int crn_set_timing(crn_ctx *ctx, int32_t enable);
int crn_get_timings(crn_ctx *ctx, crn_timings *out);
/* number of kernels this ctx has launched since creation */
int crn_get_launch_count(crn_ctx *ctx, uint64_t *count);

/* Measures one hardware ceiling with a resident synthetic kernel; result in giga lane-operations
 * per second.  which: 0 tex3D trilinear RGBA8 32^3, 1 tex3D trilinear R8 256^3, 2 LDG.32 L1-hit,
 * 3 global atomicOr (RED) on a 2 MB set, 4 shared-memory atomicOr, 5 FFMA issue,
 * 6 tex2DLayered bilinear RGBA8 32x32x32 (the noise texture's layout), 7 tex2DLayered bilinear RG16 (baked cone
 * steps), 8 the same RG8, 9 RGBA8 with f16x2 return, 10 tex3D trilinear R16, 11/12 tex3DLod on a mipmapped R8 256^3
 * at LOD 4.5 / 2.5 (mip-linear: four bilinear passes), 13 tex2DLayered bilinear RG16F, 14 tex3D with z on slice centres,
 * 15-18 tex2DLayered bilinear RGBA16_SNORM / RGBA8_SNORM 161x161x160 and tex3D RG16_SNORM 161^3 with strided (L1-missing)
 * coordinates: the combined-octave noise lattice's format and access pattern. */
int crn_microbench(int device, int32_t which, double *giga_ops_per_s);

/* library identification: "cloud-renderer_b200 <version> sm_100a" */
const char *crn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CLOUD_RENDERER_B200_H */
