"""Live comparison of the oracle with the reference's own shaders compiled as C++
(oracle/_ref/libref_glsl.so).  Needs the library `make -C oracle ref` builds from
/root/reference/res/*.glsl — present in the build container and shipped to the GPU box as a
binary; skipped where it does not exist (tests/test_golden_cpu.py covers that case)."""
import numpy as np
import pytest

from conftest import steady_state


@pytest.fixture(scope="module")
def ref(orc):
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libref_glsl.so not built (no /root/reference here)")
    orc.ref_lib()
    return orc


def test_whole_frame_through_the_reference_shader(pkg, scenes, ref):
    """every fragment of the 'tiny' frame shaded by conetrace_frag.glsl itself, blended in draw order:
    the image must be the oracle's image"""
    orc = ref
    s = steady_state(scenes.make_scene("tiny"), orc)
    s.tp.drawSun = 0
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    chain = orc.mips(l0, s.vol.levels)
    want, _, st = orc.cone_trace(s, chain, want_u8=False)
    u = orc.ref_conetrace_uniforms(s)
    back = np.array([s.cam.V[2], s.cam.V[6], s.cam.V[10]], dtype=np.float32)
    img = np.empty((s.height, s.width, 4), dtype=np.float32)
    img[:] = np.float32(s.tp.clearColor[:])
    vp = np.float32(s.vol.position[:])
    shaded = 0
    for f in orc.list_fragments(s, 1):                    # draw order: instance by instance
        i, j, b = int(f[0]), int(f[1]), int(f[2])
        ok, col = orc.ref_conetrace_fragment(s, u, chain, f[3:6], back / s.board_scale[b], f[6:8], vp + s.board_pos[b], s.board_scale[b])
        if not ok:
            continue
        shaded += 1
        src = np.clip(col, 0.0, 1.0).astype(np.float32)
        a = src[3]
        img[j, i] = src * a + img[j, i] * (np.float32(1.0) - a)
    assert shaded == st.fragments
    assert np.abs(img - want).max() <= 2e-6


def test_position_map_through_the_reference_shader(pkg, scenes, ref):
    """first_voxelize.glsl on every fragment + GL_LESS depth test = the oracle's position map;
    second_voxelize.glsl on every texel = the oracle's occupancy"""
    orc = ref
    s = steady_state(scenes.make_scene("tiny"), orc)
    posmap, depth, l0 = orc.voxelize(s)
    sd = orc.sun_update(s.vol, s.sun, pkg.SunDerived)
    lback = np.array([sd.V[2], sd.V[6], sd.V[10]], dtype=np.float32)
    vp = np.float32(s.vol.position[:])
    pm = np.zeros_like(posmap)
    dm = np.ones_like(depth)
    for f in orc.list_fragments(s, 0):
        i, j, b = int(f[0]), int(f[1]), int(f[2])
        ok, col, d = orc.ref_first_voxelize_fragment(f[3:6], lback / s.board_scale[b], vp + s.board_pos[b], s.board_scale[b], sd.nearPlane[:], sd.clipDistance)
        if not ok:
            continue
        d = np.float32(min(max(d, 0.0), 1.0))
        if d < dm[j, i]:
            dm[j, i], pm[j, i] = d, col
    assert np.array_equal(pm[..., 3], posmap[..., 3]), "coverage differs"
    assert np.abs(pm - posmap).max() <= 4e-6 and np.abs(dm - depth).max() <= 1e-6
    D = s.vol.dimension
    vol = np.zeros((D, D, D), dtype=np.uint8)
    for j, i in np.argwhere(posmap[..., 3] > 0):
        n, idx, _ = orc.ref_second_voxelize_fragment(s.vol, posmap[j, i])
        for k in range(n):
            x, y, z = idx[k]
            if 0 <= x < D and 0 <= y < D and 0 <= z < D:
                vol[z, y, x] = 255
    assert np.array_equal(vol, l0), "occupancy differs from second_voxelize.glsl run on the same position map"


def test_golden_vectors_are_current(pkg, scenes, ref):
    """the committed golden file is what the compiled reference produces today"""
    from golden_cases import CASES, build_case
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_glsl_vectors.npz"))
    orc = ref
    s = build_case(*CASES[1])
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    chain = orc.mips(l0, s.vol.levels)
    u = orc.ref_conetrace_uniforms(s)
    back = np.array([s.cam.V[2], s.cam.V[6], s.cam.V[10]], dtype=np.float32)
    for r in G["conetrace_1"][:40]:
        ok, col = orc.ref_conetrace_fragment(s, u, chain, r[0:3], back / r[8], r[3:5], r[5:8], r[8])
        assert ok == bool(r[9]) and np.array_equal(col, r[10:14])


def test_host_golden_vectors_are_current(pkg, scenes, ref):
    """tests/golden/ref_host_vectors.npz is what the reference's compiled HOST code produces today (src/Sun.hpp,
    src/Camera.cpp, src/CloudVolume.cpp, src/Shaders/ConeTraceShader.cpp through oracle/ref_glsl/host_shim)"""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host_vectors.npz"))
    orc = ref
    for r in G["sun_update"][:10]:
        o = orc.ref_host_sun_update(r[0:3], r[3:5], r[5:7], r[7:9], r[9:12])
        got = np.concatenate([o["V"], o["P"], o["nearPlane"], o["farPlane"], [o["clipDistance"]]])
        assert np.array_equal(got, r[12:])
    for r, (phi, theta) in list(zip(G["camera_update"], G["camera_angles"]))[:10]:
        o = orc.ref_host_camera_update(int(r[0]), int(r[1]), r[2:5], phi, theta)
        assert np.array_equal(o["P"], r[5:21]) and np.array_equal(o["V"], r[21:37]) and np.array_equal(o["lookAt"], r[37:40])
    a, b, pts = G["sort_1_in"], G["sort_1_out"], G["sort_1_pts"]
    p, s = orc.ref_host_sort_boards(a[:, :3], a[:, 3], pts[:3], pts[3:])
    assert np.array_equal(p, b[:, :3]) and np.array_equal(s, b[:, 3])
    assert np.array_equal(orc.ref_host_noise_normals(G["noise_16_alpha"]), G["noise_16_rgba"])
