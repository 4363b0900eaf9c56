// k_voxelize.cu — stage 1: spherical-billboard voxelization (pass 1 + pass 2 fused).
//
// Replaces VoxelizeShader::firstVoxelize + secondVoxelize
// (src/Shaders/VoxelizeShader.cpp:36-118) and their GPU programs
// (res/billboard_vert_instanced.glsl, res/first_voxelize.glsl:39-64,
//  res/second_voxelize.glsl:34-51):
//
//   pass 1: per light-space texel, the sun-facing hemisphere point with the smallest radial
//           depth over all billboards covering the texel (GL_LESS z-buffer, earlier instance
//           wins ties);
//   pass 2: per valid texel, 9 constant stores into the volume (the point and its 8 diagonal
//           neighbours at stepSize/sqrt(3) per axis).
//
// B200 design.  One thread per texel, a 16x16 texel tile per CTA, each warp an 8x4 patch.
// The z-buffer is a register: the tile's billboard list (k_bin.cu) is ordered front-to-back
// from the sun by a conservative depth bound, so a warp stops walking the list as soon as no
// later billboard can win anywhere in its patch (one vote per entry).  The 133 MB RGBA32F
// position map + depth buffer of the reference never exist unless the caller asks for the
// debug copy.  The 9 stores are idempotent constants, so the volume's level 0 is kept as a
// 1-bit-per-voxel occupancy set (2 MB at 256^3, L2-resident): lanes of a warp that hit the
// same 32-voxel word merge their bits with match.any/redux.or and issue one atomicOr.
// k_mips.cu expands the set into the R8 texture.
//
// Compiled with -fmad=false: which voxel a texel lands in is part of the exact-parity
// contract, so the arithmetic is the shader's, one IEEE rounding per written operation.
#include "crn_internal.cuh"

namespace crn {

namespace {

struct VoxArgs {
    VolumeParams vol;
    float nearPlane[3];
    float clip;
    const BoardRec *recs;
    const float *lb;
    const uint32_t *tileOff, *tileCnt, *tileList;
    int tilesX;
    uint32_t *bits;
    uint32_t *bitsA;              // CRN_VOLUME_RG8: the occupancy (alpha) set also receives the lit stores, else nullptr
    float4 *posmap;
};

__global__ void __launch_bounds__(256) voxelize_kernel(VoxArgs a, ViewParams lp) {
    const int tile = blockIdx.x;
    const uint32_t cnt = a.tileCnt[tile];
    if (cnt == 0) return;                                       // untouched texels stay cleared
    const uint32_t off = a.tileOff[tile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = tile % a.tilesX, ty = tile / a.tilesX;
    const int px = tx * kTile + (warp & 1) * 8 + (lane & 7);
    const int py = ty * kTile + (warp >> 1) * 4 + (lane >> 3);
    const bool valid = px < lp.W && py < lp.H;

    // pixel-centre un-projection through the sun's ortho camera
    const float ndcx = ((float)px + 0.5f) / (float)lp.W * 2.0f - 1.0f;
    const float ndcy = ((float)py + 0.5f) / (float)lp.H * 2.0f - 1.0f;
    const float xv = (ndcx - lp.P[12]) / lp.P[0];
    const float yv = (ndcy - lp.P[13]) / lp.P[5];

    const float invClip = 1.0f / a.clip;
    float best = valid ? 1.0f : -1.0f;                          // depth clear value; dead lanes never block the vote
    int bestIdx = -1;
    float wx = 0.0f, wy = 0.0f, wz = 0.0f;

    for (uint32_t e = 0; e < cnt; e++) {
        const uint32_t k = a.tileList[off + e];
        const float lbk = a.lb[k];
        // list is ascending in lbk and lbk < every depth the billboard can produce
        if (__all_sync(0xFFFFFFFFu, !(lbk < best))) break;
        const float4 r0 = *reinterpret_cast<const float4 *>(&a.recs[k].cx);
        const float4 r1 = *reinterpret_cast<const float4 *>(&a.recs[k].xv);
        const float radius = r0.w;
        const float u = xv - r1.x, v = yv - r1.y;
        if (!(fabsf(u) < radius && fabsf(v) < radius) || !(lbk < best)) continue;
        {   // cheap conservative filter (MUFU maths, relative error < 1e-5) before the exact IEEE evaluation:
            // only candidates that are neither surely discarded nor surely behind the current winner go on
            const float r2 = radius * radius, h2 = r2 - (u * u + v * v);
            if (h2 < 0.9e-4f * r2) continue;                    // sphereContrib < 0.01 with margin: discarded for sure
            const float hh = h2 * rsqrtf(h2);                   // ~ radius * sphereContrib
            const float ax = (r0.x - a.nearPlane[0]) + (u * lp.right[0] + v * lp.up[0]) + lp.nrm[0] * hh;
            const float ay = (r0.y - a.nearPlane[1]) + (u * lp.right[1] + v * lp.up[1]) + lp.nrm[1] * hh;
            const float az = (r0.z - a.nearPlane[2]) + (u * lp.right[2] + v * lp.up[2]) + lp.nrm[2] * hh;
            const float q2 = ax * ax + ay * ay + az * az;
            const float dApprox = q2 * rsqrtf(q2) * invClip;
            if (dApprox > best * 1.0001f + 1.0e-6f) continue;   // cannot win (ties included) even with the error margin
        }
        // fragPos: interpolated quad position at this texel centre
        const float fx = r0.x + (u * lp.right[0] + v * lp.up[0]);
        const float fy = r0.y + (u * lp.right[1] + v * lp.up[1]);
        const float fz = r0.z + (u * lp.right[2] + v * lp.up[2]);
        // first_voxelize.glsl:42-48
        const float dx = fx - r0.x, dy = fy - r0.y, dz = fz - r0.z;
        float sc = sqrtf((dx * dx + dy * dy) + dz * dz) / radius;
        sc = sqrtf(fmaxf(0.0f, 1.0f - sc * sc));
        if (sc < 0.01f) continue;                               // discard
        // first_voxelize.glsl:50-63
        const float dist = radius * sc;
        const float px3 = fx + lp.nrm[0] * dist, py3 = fy + lp.nrm[1] * dist, pz3 = fz + lp.nrm[2] * dist;
        const float ex = px3 - a.nearPlane[0], ey = py3 - a.nearPlane[1], ez = pz3 - a.nearPlane[2];
        float d = sqrtf((ex * ex + ey * ey) + ez * ez) / a.clip;
        d = fminf(fmaxf(d, 0.0f), 1.0f);                        // gl_FragDepth clamp
        const int idx = __float_as_int(r1.w);
        if (d < best || (d == best && bestIdx >= 0 && idx < bestIdx)) {   // GL_LESS, first instance wins ties
            best = d; bestIdx = idx; wx = px3; wy = py3; wz = pz3;
        }
    }

    const bool hit = bestIdx >= 0;
    if (a.posmap && valid) a.posmap[(size_t)py * lp.W + px] = hit ? make_float4(wx, wy, wz, 1.0f) : make_float4(0, 0, 0, 0);

    // second_voxelize.glsl:34-51
    const int D = a.vol.dim;
    const float fd = (float)D;
    const float rangeX = a.vol.xB[1] - a.vol.xB[0], rangeY = a.vol.yB[1] - a.vol.yB[0], rangeZ = a.vol.zB[1] - a.vol.zB[0];
    const float delta = a.vol.stepSize * (1.0f / sqrtf(3.0f));
    const int wordsPerRow = D >> 5;
#pragma unroll 1
    for (int s = 0; s < 9; s++) {
        float qx = wx, qy = wy, qz = wz;
        if (s > 0) {
            const int m = s - 1;
            qx = wx + ((m & 4) ? -delta : delta);
            qy = wy + ((m & 2) ? -delta : delta);
            qz = wz + ((m & 1) ? -delta : delta);
        }
        const float vx = fd * ((qx - a.vol.xB[0]) / rangeX);
        const float vy = fd * ((qy - a.vol.yB[0]) / rangeY);
        const float vz = fd * ((qz - a.vol.zB[0]) / rangeZ);
        bool ok = hit && vx > -1.0f && vx < fd && vy > -1.0f && vy < fd && vz > -1.0f && vz < fd;
        uint32_t word = 0xFFFFFFFFu, bit = 0;
        if (ok) {
            const int ix = (int)vx, iy = (int)vy, iz = (int)vz;   // ivec3(): truncation toward zero
            ok = iz >= a.vol.z0 && iz < a.vol.z1;                 // slab ownership (whole volume by default)
            if (ok) {
                word = (uint32_t)((iz * D + iy) * wordsPerRow + (ix >> 5));
                bit = 1u << (ix & 31);
            }
        }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, word);
        const uint32_t merged = __reduce_or_sync(peers, bit);
        if (ok && lane == __ffs(peers) - 1) {
            atomicOr(&a.bits[word], merged);
            if (a.bitsA) atomicOr(&a.bitsA[word], merged);      // (1,1,1,1): the lit shell is occupied too
        }
    }
}

// Paper variant, first-pass interior march (res/first_voxelize.glsl:53-58, live in
// paper/tex/voxelization.tex:13-27): EVERY non-discarded fragment of EVERY billboard (image
// stores are not depth-tested, so there is no front-to-back early-out here) walks the chord of
// its sphere along the light direction in steps of stepSize and marks the voxels it lands in.
// Same tiling as voxelize_kernel; a warp's 8x4 texel patch spans about one voxel, so per step
// the 32 lanes hit 1-4 distinct words: match.any/redux.or merges them and the leader skips the
// atomic when the word already holds the bits it stored on the previous step.
__global__ void __launch_bounds__(256) interior_kernel(VoxArgs a, ViewParams lp) {
    const int tile = blockIdx.x;
    const uint32_t cnt = a.tileCnt[tile];
    if (cnt == 0) return;
    const uint32_t off = a.tileOff[tile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = tile % a.tilesX, ty = tile / a.tilesX;
    const int px = tx * kTile + (warp & 1) * 8 + (lane & 7);
    const int py = ty * kTile + (warp >> 1) * 4 + (lane >> 3);
    const bool valid = px < lp.W && py < lp.H;
    const float ndcx = ((float)px + 0.5f) / (float)lp.W * 2.0f - 1.0f;
    const float ndcy = ((float)py + 0.5f) / (float)lp.H * 2.0f - 1.0f;
    const float xv = (ndcx - lp.P[12]) / lp.P[0];
    const float yv = (ndcy - lp.P[13]) / lp.P[5];
    const int D = a.vol.dim;
    const float fd = (float)D;
    const float rangeX = a.vol.xB[1] - a.vol.xB[0], rangeY = a.vol.yB[1] - a.vol.yB[0], rangeZ = a.vol.zB[1] - a.vol.zB[0];
    const int wordsPerRow = D >> 5;
    const float step = a.vol.stepSize;
    const float kx = fd / rangeX, ky = fd / rangeY, kz = fd / rangeZ;      // approximate index scale (prefilter only)
    uint32_t lastWord = 0xFFFFFFFFu, lastBits = 0;

    for (uint32_t e = 0; e < cnt; e++) {
        const uint32_t k = a.tileList[off + e];
        const float4 r0 = *reinterpret_cast<const float4 *>(&a.recs[k].cx);
        const float4 r1 = *reinterpret_cast<const float4 *>(&a.recs[k].xv);
        const float radius = r0.w;
        const float u = xv - r1.x, v = yv - r1.y;
        bool act = valid && fabsf(u) < radius && fabsf(v) < radius;
        float two = 0.0f, sx = 0.0f, sy = 0.0f, sz = 0.0f;
        if (act) {
            const float fx = r0.x + (u * lp.right[0] + v * lp.up[0]);
            const float fy = r0.y + (u * lp.right[1] + v * lp.up[1]);
            const float fz = r0.z + (u * lp.right[2] + v * lp.up[2]);
            const float dx = fx - r0.x, dy = fy - r0.y, dz = fz - r0.z;
            float sc = sqrtf((dx * dx + dy * dy) + dz * dz) / radius;
            sc = sqrtf(fmaxf(0.0f, 1.0f - sc * sc));
            act = !(sc < 0.01f);                                            // discard
            const float dist = radius * sc;
            two = 2.0f * dist;
            sx = fx - lp.nrm[0] * dist; sy = fy - lp.nrm[1] * dist; sz = fz - lp.nrm[2] * dist;   // start
        }
        float s = 0.0f;
        while (__any_sync(0xFFFFFFFFu, act && s < two)) {
            bool ok = act && s < two;
            uint32_t word = 0xFFFFFFFFu, bit = 0;
            if (ok) {
                const float qx = sx + lp.nrm[0] * s, qy = sy + lp.nrm[1] * s, qz = sz + lp.nrm[2] * s;
                // one multiply per axis gives the voxel coordinate to ~4 ulp; only when that lands within 1e-3 of an
                // integer (where truncation could differ) is the shader's own divide-then-scale evaluated
                float vx = (qx - a.vol.xB[0]) * kx, vy = (qy - a.vol.yB[0]) * ky, vz = (qz - a.vol.zB[0]) * kz;
                const float nearInt = fminf(fminf(fabsf(vx - rintf(vx)), fabsf(vy - rintf(vy))), fabsf(vz - rintf(vz)));
                if (nearInt < 1.0e-3f) {
                    vx = fd * ((qx - a.vol.xB[0]) / rangeX);
                    vy = fd * ((qy - a.vol.yB[0]) / rangeY);
                    vz = fd * ((qz - a.vol.zB[0]) / rangeZ);
                }
                ok = vx > -1.0f && vx < fd && vy > -1.0f && vy < fd && vz > -1.0f && vz < fd;
                if (ok) {
                    const int ix = (int)vx, iy = (int)vy, iz = (int)vz;
                    ok = iz >= a.vol.z0 && iz < a.vol.z1;
                    if (ok) {
                        word = (uint32_t)((iz * D + iy) * wordsPerRow + (ix >> 5));
                        bit = 1u << (ix & 31);
                    }
                }
            }
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, word);
            const uint32_t merged = __reduce_or_sync(peers, bit);
            if (ok && lane == __ffs(peers) - 1) {
                if (word != lastWord || (merged & ~lastBits)) {
                    // overlapping spheres and neighbouring patches have usually set these bits already: a plain
                    // (possibly stale, hence only ever pessimistic) load saves the contended atomic
                    if ((a.bitsA[word] & merged) != merged) atomicOr(&a.bitsA[word], merged);
                    lastBits = (word == lastWord ? lastBits : 0u) | merged;
                    lastWord = word;
                }
            }
            s += step;
        }
    }
}

} // namespace

int launch_voxelize(cudaStream_t st, const ViewParams &light, const VolumeParams &vol, const float nearPlane[3],
                    float clip, const BoardRec *recs, const float *lbSorted, const Bins &b, uint32_t *bits,
                    float4 *posmap, uint32_t *bitsA) {
    VoxArgs a;
    a.vol = vol;
    for (int k = 0; k < 3; k++) a.nearPlane[k] = nearPlane[k];
    a.clip = clip;
    a.recs = recs; a.lb = lbSorted;
    a.tileOff = b.tileOff; a.tileCnt = b.tileCnt; a.tileList = b.tileList; a.tilesX = b.tilesX;
    a.bits = bits; a.posmap = posmap; a.bitsA = bitsA;
    const size_t words = (size_t)vol.dim * vol.dim * vol.dim / 32;
    const size_t w0 = (size_t)vol.z0 * vol.dim * (vol.dim / 32), w1 = (size_t)vol.z1 * vol.dim * (vol.dim / 32);
    (void)words;
    cudaMemsetAsync(bits + w0, 0, (w1 - w0) * sizeof(uint32_t), st);       // CloudVolume::clearGPU for the owned slab
    if (posmap) cudaMemsetAsync(posmap, 0, (size_t)light.W * light.H * sizeof(float4), st);   // clearPositionMap
    if (bitsA) {
        cudaMemsetAsync(bitsA + w0, 0, (w1 - w0) * sizeof(uint32_t), st);
        interior_kernel<<<b.tilesX * b.tilesY, 256, 0, st>>>(a, light);
    }
    voxelize_kernel<<<b.tilesX * b.tilesY, 256, 0, st>>>(a, light);
    return bitsA ? 2 : 1;
}

} // namespace crn
