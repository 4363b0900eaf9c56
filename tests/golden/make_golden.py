"""Generates tests/golden/ref_glsl_vectors.npz from the REFERENCE'S OWN SHADERS compiled as C++
(oracle/_ref/libref_glsl.so, built by `make -C oracle ref` from /root/reference/res/*.glsl).
Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each record is (inputs, outputs) of one shader main() invocation; tests/test_golden_cpu.py replays
the inputs through the oracle and compares.  Scenes are the fixed-seed fixtures of scene.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.import_package()
from cloud_renderer_b200 import scene as sc  # noqa: E402

orc = entry.import_oracle()

CASES = [                      # (scene name, parameter tweak id)
    ("tiny", "default"), ("small", "default"), ("small", "paper_params"), ("small", "no_noise"), ("small", "anisotropic"),
]


def tweak(s, which):
    if which == "paper_params":           # the overlay of paper/res/res3.png
        s.tp.vctSteps, s.tp.vctConeAngle, s.tp.vctConeInitialHeight, s.tp.vctDownScaling = 22, 0.76, 0.64, 1.52
        s.tp.freqStep, s.tp.persStep, s.tp.runTime = 1.475, 0.75, 7.25
    elif which == "no_noise":
        s.tp.doNoiseSample = 0
    elif which == "anisotropic":
        s.vol.xBounds[:], s.vol.yBounds[:], s.vol.zBounds[:] = (-6.0, 4.0), (-3.0, 5.0), (-5.0, 5.5)
        s.tp.numOctaves, s.tp.vctLodOffset = 3, 0.35
    return s


def build_case(name, which):
    s = tweak(sc.make_scene(name), which)
    s.board_pos, s.board_scale = orc.sort_boards(s.board_pos, s.board_scale, s.vol.position, s.cam.position)
    return s


def main():
    assert orc.ref_available(), "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(20260101)
    out = {}
    for ci, (name, which) in enumerate(CASES):
        s = build_case(name, which)
        posmap, _, l0 = orc.voxelize(s)
        chain = orc.mips(l0, s.vol.levels)
        vp = np.float32(s.vol.position[:])
        # ---- conetrace_frag.glsl
        fr = orc.list_fragments(s, 1)
        pick = rng.choice(len(fr), size=min(160, len(fr)), replace=False)
        u = orc.ref_conetrace_uniforms(s)
        back = np.array([s.cam.V[2], s.cam.V[6], s.cam.V[10]], dtype=np.float32)
        rec = []
        for k in pick:
            b = int(fr[k, 2])
            c, r = vp + s.board_pos[b], np.float32(s.board_scale[b])
            ok, col = orc.ref_conetrace_fragment(s, u, chain, fr[k, 3:6], back / r, fr[k, 6:8], c, r)
            rec.append(np.concatenate([fr[k, 3:8], c, [r, float(ok)], col]))
        out[f"conetrace_{ci}"] = np.array(rec, dtype=np.float32)          # fragPos3 fragTex2 center3 radius ok color4
        # ---- first_voxelize.glsl
        sd = orc.sun_update(s.vol, s.sun, pkg.SunDerived)
        lback = np.array([sd.V[2], sd.V[6], sd.V[10]], dtype=np.float32)
        fl = orc.list_fragments(s, 0)
        pick = rng.choice(len(fl), size=min(160, len(fl)), replace=False)
        rec = []
        for k in pick:
            b = int(fl[k, 2])
            c, r = vp + s.board_pos[b], np.float32(s.board_scale[b])
            ok, col, d = orc.ref_first_voxelize_fragment(fl[k, 3:6], lback / r, c, r, sd.nearPlane[:], sd.clipDistance)
            rec.append(np.concatenate([fl[k, 3:6], c, [r, float(ok)], col, [d]]))
        out[f"first_voxelize_{ci}"] = np.array(rec, dtype=np.float32)     # fragPos3 center3 radius ok color4 depth
        # ---- second_voxelize.glsl on valid position-map texels (+ a few just outside the volume)
        valid = np.argwhere(posmap[..., 3] > 0)
        pick = valid[rng.choice(len(valid), size=min(96, len(valid)), replace=False)]
        texels = [posmap[j, i] for j, i in pick]
        lo = vp + np.float32([s.vol.xBounds[0], s.vol.yBounds[0], s.vol.zBounds[0]])
        hi = vp + np.float32([s.vol.xBounds[1], s.vol.yBounds[1], s.vol.zBounds[1]])
        for t in range(16):
            p = np.where(rng.random(3) < 0.5, lo, hi) + np.float32(rng.uniform(-0.3, 0.3, 3))
            texels.append(np.float32([p[0], p[1], p[2], 1.0]))
        texels.append(np.float32([0, 0, 0, 0]))                          # invalid texel: no stores
        rec = []
        for t in texels:
            n, idx, val = orc.ref_second_voxelize_fragment(s.vol, t)
            rec.append(np.concatenate([t, [n], idx.ravel().astype(np.float32), val]))
        out[f"second_voxelize_{ci}"] = np.array(rec, dtype=np.float32)    # texel4 n idx27 val9
    # ---- sun_frag.glsl
    s = sc.make_scene("tiny")
    rec = []
    for d in np.linspace(0.0, 2.2, 45):
        p = np.float32(s.sun.position[:]) + np.float32([0.0, d * 0.6, d * 0.8])
        ok, col = orc.ref_sun_fragment(s.sun, p)
        rec.append(np.concatenate([p, [float(ok)], col]))
    out["sun"] = np.array(rec, dtype=np.float32)
    # ---- billboard_vert_instanced.glsl: the 4 corners of a few instances under both cameras
    rec = []
    sd = orc.sun_update(s.vol, s.sun, pkg.SunDerived)
    for cam_id, (P, V) in enumerate(((s.cam.P, s.cam.V), (sd.P, sd.V))):
        for b in range(6):
            for vx, vy in ((-1, -1), (1, -1), (-1, 1), (1, 1)):
                o = orc.ref_billboard_vertex(P, V, s.vol.position[:], (vx, vy), s.board_pos[b], s.board_scale[b])
                rec.append(np.concatenate([[cam_id, b, vx, vy], o["gl_Position"], o["fragPos"], o["fragNor"], o["fragTex"], o["center"], [o["scale"]]]))
    out["vertex"] = np.array(rec, dtype=np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_glsl_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
