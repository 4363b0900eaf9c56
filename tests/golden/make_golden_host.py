"""Generates tests/golden/ref_host_vectors.npz from the REFERENCE'S OWN HOST CODE and its first-pass march, compiled
from /root/reference (oracle/_ref/libref_glsl.so: src/Sun.hpp, src/Camera.cpp, src/CloudVolume.cpp,
src/Shaders/ConeTraceShader.cpp through stub GLM/GLFW headers; res/first_voxelize.glsl with lines 54-58 un-commented;
res/billboard_vert_instanced.glsl for interior fragment attributes).  Run in the build container:

    python tests/golden/make_golden_host.py

tests/test_golden_host_cpu.py replays the inputs through the oracle AND through the product's host entry points."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.import_package()
from cloud_renderer_b200 import scene as sc  # noqa: E402

orc = entry.import_oracle()

MARCH_CASES = [("tiny", "default"), ("small", "default"), ("small", "anisotropic")]


def main():
    assert orc.ref_available(), "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20261017)
    f32 = np.float32
    out = {}
    d = orc.ref_host_sun_defaults()                          # before anything moves the sun: the values of src/main.cpp:37-46
    out["sun_defaults"] = np.concatenate([d["position"], d["innerColor"], d["outerColor"], [d["innerRadius"], d["outerRadius"]]]).astype(f32)
    # ---- Sun::update: inputs 14 floats -> V16 P16 near3 far3 clip
    rec = []
    ins = [np.concatenate([[25, 0, 0], [-5, 5], [-5, 5], [-5, 5], [5, 20, -5]])]            # the reference's default scene
    for _ in range(40):
        lo, hi = -rng.uniform(1, 9, 3), rng.uniform(1, 9, 3)
        ins.append(np.concatenate([rng.uniform(-30, 30, 3), [lo[0], hi[0]], [lo[1], hi[1]], [lo[2], hi[2]], rng.uniform(-40, 40, 3)]))
    for v in ins:
        v = v.astype(f32)
        o = orc.ref_host_sun_update(v[0:3], v[3:5], v[5:7], v[7:9], v[9:12])
        rec.append(np.concatenate([v, o["V"], o["P"], o["nearPlane"], o["farPlane"], [o["clipDistance"]]]))
    out["sun_update"] = np.array(rec, dtype=f32)
    # ---- Camera::update: (W, H, position3, phi, theta) -> P16 V16 lookAt3   (phi/theta kept as float64 in a second array)
    rec, ang = [], []
    cams = [(1280, 720, (0, 0, 0), 0.0, 0.0), (1920, 1080, (0, 0, 0), 0.0, 0.0), (3840, 2160, (0, 0, 0), 0.0, 0.0), (7680, 4320, (0, 0, 0), 0.0, 0.0),
            (720, 1280, (1, 2, 3), 0.3, -0.7), (1000, 400, (-4, 0.5, 9), -0.4, 2.1)]
    for _ in range(30):
        cams.append((int(rng.integers(200, 4000)), int(rng.integers(200, 4000)), tuple(rng.uniform(-30, 30, 3)), float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-3.1, 3.1))))
    for W, H, pos, phi, theta in cams:
        pos = np.array(pos, dtype=f32)
        o = orc.ref_host_camera_update(W, H, pos, phi, theta)
        rec.append(np.concatenate([[W, H], pos, o["P"], o["V"], o["lookAt"]]))
        ang.append([phi, theta])
    out["camera_update"] = np.array(rec, dtype=f32)
    out["camera_angles"] = np.array(ang, dtype=np.float64)
    # ---- CloudVolume::sortBoards (selection sort, farthest first)
    for k, n in enumerate((24, 200, 500)):
        pos = rng.uniform(-2.5, 2.5, (n, 3)).astype(f32)
        scale = rng.uniform(1.0, 2.5, n).astype(f32)
        if k == 2:                                           # exact ties: the selection sort's behaviour on equal keys
            pos[100:110] = pos[50:60]
        volpos, point = np.array([25, 0, 0], f32), np.array([0.5 * k, 0, 0], f32)
        p, s = orc.ref_host_sort_boards(pos, scale, volpos, point)
        out[f"sort_{k}_in"] = np.concatenate([pos, scale[:, None]], axis=1)
        out[f"sort_{k}_out"] = np.concatenate([p, s[:, None]], axis=1)
        out[f"sort_{k}_pts"] = np.concatenate([volpos, point])
    # ---- CloudVolume::update / get3DIndices / reverseVoxelIndex
    rec = []
    for dim, xb, yb, zb in ((32, (-5, 5), (-5, 5), (-5, 5)), (64, (-6, 4), (-3, 5), (-5, 5.5)), (256, (-5, 5), (-5, 5), (-5, 5))):
        for index in [0, 1, dim - 1, dim, dim * dim - 1, dim * dim, dim ** 3 - 1] + [int(i) for i in rng.integers(0, dim ** 3, 12)]:
            ijk, w, vs = orc.ref_host_voxel_index(dim, (25, 0, 0), xb, yb, zb, index)
            rec.append(np.concatenate([[dim], xb, yb, zb, [index], ijk, w, vs]).astype(np.float64))
    out["voxel_index"] = np.array(rec, dtype=np.float64)
    # ---- ConeTraceShader::initNoiseMap's normal loop
    for dim in (8, 16, 32):
        alpha = rng.integers(-128, 128, dim ** 3).astype(np.int8)
        if dim == 8:
            alpha[:40] = 0                                   # flat neighbourhoods: the zero-gradient (NaN normal) case
        out[f"noise_{dim}_alpha"] = alpha
        out[f"noise_{dim}_rgba"] = orc.ref_host_noise_normals(alpha)
    # ---- the first pass with its interior march switched back on (paper variant)
    from make_golden import build_case           # same scene table as the shader vectors
    for ci, (name, which) in enumerate(MARCH_CASES):
        s = build_case(name, which)
        vp = f32(s.vol.position[:])
        sd = orc.sun_update(s.vol, s.sun, pkg.SunDerived)
        lback = np.array([sd.V[2], sd.V[6], sd.V[10]], dtype=f32)
        fl = orc.list_fragments(s, 0)
        pick = rng.choice(len(fl), size=min(120, len(fl)), replace=False)
        rows, seqs = [], []
        for k in pick:
            b = int(fl[k, 2])
            c, r = vp + s.board_pos[b], f32(s.board_scale[b])
            n, idx, col, dep = orc.ref_first_voxelize_paper_fragment(s.vol, fl[k, 3:6], lback / r, c, r, sd.nearPlane[:], sd.clipDistance)
            rows.append(np.concatenate([fl[k, 3:6], c, [r, n], col, [dep]]))
            seqs.append(idx)
        out[f"march_{ci}"] = np.array(rows, dtype=f32)                   # fragPos3 center3 radius n color4 depth
        out[f"march_{ci}_idx"] = np.concatenate(seqs).astype(np.int32) if seqs else np.zeros((0, 3), np.int32)
    # ---- interior fragment attributes: the compiled vertex shader at the 4 corners + affine interpolation at pixel centres
    s = sc.make_scene("small")
    sd = orc.sun_update(s.vol, s.sun, pkg.SunDerived)
    rec = []
    for cam_id, (P, V) in enumerate(((s.cam.P, s.cam.V), (sd.P, sd.V))):
        fr = orc.list_fragments(s, 1 if cam_id == 0 else 0)              # (i, j, board, fragPos3, fragTex2)
        pick = rng.choice(len(fr), size=200, replace=False)
        for k in pick:
            i, j, b = int(fr[k, 0]), int(fr[k, 1]), int(fr[k, 2])
            corners = {(vx, vy): orc.ref_billboard_vertex(P, V, s.vol.position[:], (vx, vy), s.board_pos[b], s.board_scale[b])
                       for vx, vy in ((-1, -1), (1, -1), (-1, 1), (1, 1))}
            # window coordinates of the corners (all four share clip w: the interpolation is affine in window space)
            def win(o):
                g = o["gl_Position"].astype(np.float64)
                return np.array([(g[0] / g[3] + 1) * 0.5 * s.width, (g[1] / g[3] + 1) * 0.5 * s.height])
            w00, w10, w01 = win(corners[(-1, -1)]), win(corners[(1, -1)]), win(corners[(-1, 1)])
            px = np.array([i + 0.5, j + 0.5])
            a = (px[0] - w00[0]) / (w10[0] - w00[0])                     # quad edges are axis-aligned in window space
            bb = (px[1] - w00[1]) / (w01[1] - w00[1])
            f00, f10, f01 = (corners[c]["fragPos"].astype(np.float64) for c in ((-1, -1), (1, -1), (-1, 1)))
            t00, t10, t01 = (corners[c]["fragTex"].astype(np.float64) for c in ((-1, -1), (1, -1), (-1, 1)))
            fp = f00 + a * (f10 - f00) + bb * (f01 - f00)
            ft = t00 + a * (t10 - t00) + bb * (t01 - t00)
            rec.append(np.concatenate([[cam_id, i, j, b], fp, ft]))
    out["interior_fragments"] = np.array(rec, dtype=np.float64)          # cam i j board fragPos3 fragTex2 (from the compiled vertex shader)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_host_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()
