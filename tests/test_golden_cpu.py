"""The oracle against golden vectors produced by the REFERENCE'S OWN SHADERS compiled as C++
(tests/golden/make_golden.py).  Runs anywhere: needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_glsl_vectors.npz"))

from golden_cases import CASES, build_case  # noqa: E402  (tests/golden_cases.py re-exports make_golden's table)


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_conetrace_fragments(ci, pkg, scenes, orc):
    s = build_case(*CASES[ci])
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    chain = orc.mips(l0, s.vol.levels)
    rec = G[f"conetrace_{ci}"]
    worst = 0.0
    for r in rec:
        ok, col = orc.conetrace_fragment(s, chain, r[0:3], r[3:5], r[5:8], r[8])
        assert ok == bool(r[9]), "discard decision differs from conetrace_frag.glsl"
        if ok:
            worst = max(worst, float(np.abs(col - r[10:14]).max()))
    # identical float32 operation order; the only freedom is normalize(fragNor) of an interpolated normal
    assert worst <= 2e-6, worst


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_first_voxelize_fragments(ci, pkg, scenes, orc):
    s = build_case(*CASES[ci])
    rec = G[f"first_voxelize_{ci}"]
    for r in rec:
        ok, wp, d = orc.first_voxelize_fragment(s, r[0:3], r[3:6], r[6])
        assert ok == bool(r[7])
        if ok:
            assert np.abs(wp - r[8:11]).max() <= 4e-6 and r[11] == 1.0     # color = vec4(worldPos, 1)
            assert abs(d - r[12]) <= 1e-6                                 # gl_FragDepth (before the fixed-function clamp)


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_second_voxelize_stores(ci, pkg, scenes, orc):
    s = build_case(*CASES[ci])
    D = s.vol.dimension
    for r in G[f"second_voxelize_{ci}"]:
        n, idx, val = int(r[4]), r[5:32].reshape(9, 3).astype(np.int32), r[32:41]
        if r[3] <= 0:
            assert n == 0                                                 # `if (worldPos.a > 0)`
            continue
        assert n == 9 and (val == 1.0).all()                             # nine imageStore(..., vec4(1))
        mine = orc.second_voxelize_indices(s.vol, r[0:3])
        for k in range(9):
            in_range = ((idx[k] >= 0) & (idx[k] < D)).all()
            if in_range:
                assert np.array_equal(mine[k], idx[k]), "voxel index differs from second_voxelize.glsl"
            else:
                # GL drops the out-of-range store; (-1,0) truncates to 0 and is therefore NOT out of range
                assert (mine[k] == -1).all()


def test_sun_fragments(pkg, scenes, orc):
    """sun_frag.glsl (res/sun_frag.glsl:14-28): inner disc opaque, linear falloff, discard beyond 0.99"""
    s = scenes.make_scene("tiny")
    for r in G["sun"]:
        # sun_frag.glsl is four lines; the golden rows are checked against its closed form
        dist = float(np.float32(np.linalg.norm((np.float32(r[0:3]) - np.float32(s.sun.position[:])).astype(np.float32))))
        if dist < s.sun.innerRadius:
            assert r[3] == 1.0 and np.allclose(r[4:8], [1, 1, 1, 1])
        else:
            sc_ = (dist - s.sun.innerRadius) / (s.sun.outerRadius - s.sun.innerRadius)
            if sc_ > 0.99 + 1e-6:
                assert r[3] == 0.0
            elif sc_ < 0.99 - 1e-6:
                assert r[3] == 1.0 and abs(r[7] - (1 - sc_)) < 1e-5 and abs(r[6] - (1 - sc_)) < 1e-5 and abs(r[4] - 1.0) < 1e-6


def test_vertex_stage_claims(pkg, scenes, orc):
    """billboard_vert_instanced.glsl: the facts the per-pixel formulation relies on (DESIGN.md decree 6)"""
    s = scenes.make_scene("tiny")
    sd = orc.sun_update(s.vol, s.sun, pkg.SunDerived)
    v = G["vertex"]
    for cam_id, V in enumerate((s.cam.V, sd.V)):
        right = np.array([V[0], V[4], V[8]]); up = np.array([V[1], V[5], V[9]]); back = np.array([V[2], V[6], V[10]])
        rows = v[v[:, 0] == cam_id]
        for b in range(6):
            q = rows[rows[:, 1] == b]
            c = np.float32(s.vol.position[:]) + s.board_pos[b]
            r = s.board_scale[b]
            # all four corners share clip z and w: attribute interpolation is affine for ortho AND perspective
            assert np.ptp(q[:, 6]) <= 2e-5 * max(1.0, abs(q[0, 6])) and np.ptp(q[:, 7]) <= 2e-5 * max(1.0, abs(q[0, 7]))
            for row in q:
                vx, vy = row[2], row[3]
                assert np.allclose(row[8:11], c + r * (vx * right + vy * up), atol=2e-5)      # fragPos = center + scale*(vx*right + vy*up)
                n = row[11:14] / np.linalg.norm(row[11:14])
                assert np.allclose(n, back / np.linalg.norm(back), atol=1e-5)                 # normalize(fragNor) = view back axis
                assert np.allclose(row[14:16], [(vx + 1) / 2, (vy + 1) / 2])                  # fragTex
                assert np.allclose(row[16:19], c, atol=1e-6) and abs(row[19] - r) < 1e-7       # flat center, scale
