"""Closed-form anchors for the CPU oracle: the reference has no golden vectors,
so the restatement is checked against answers that can be derived by hand, and against an
independent numpy statement of the integer stages."""
import numpy as np
import pytest


def np_mips(level0, levels):
    out, cur = [level0.ravel()], level0.astype(np.uint32)
    for _ in range(1, levels):
        n = cur.shape[0] // 2
        s = cur.reshape(n, 2, n, 2, n, 2).sum(axis=(1, 3, 5))
        cur = (s + 4) >> 3
        out.append(cur.astype(np.uint8).ravel())
    return np.concatenate(out)


def test_mips_all_ones_and_checkerboard(orc):
    D, L = 32, 6
    ones = np.full((D, D, D), 255, np.uint8)
    assert (orc.mips(ones, L) == 255).all()                      # (ii) all-ones level 0 -> every mip 255
    z, y, x = np.indices((D, D, D))
    checker = (((x + y + z) & 1) * 255).astype(np.uint8)
    chain = orc.mips(checker, L)
    assert (chain[D ** 3:] == 128).all()                         # 4 of 8 lit: (4*255+4)>>3 = 128, and 128 is a fixed point
    assert (orc.mips(np.zeros((D, D, D), np.uint8), L) == 0).all()


def test_mips_match_numpy(orc):
    rng = np.random.default_rng(11)
    for D, L in ((32, 4), (32, 6), (64, 7)):
        l0 = (rng.random((D, D, D)) < 0.3).astype(np.uint8) * 255
        assert np.array_equal(orc.mips(l0, L), np_mips(l0, L))
    l0 = rng.integers(0, 256, (32, 32, 32)).astype(np.uint8)     # arbitrary bytes, rounding ties included
    assert np.array_equal(orc.mips(l0, 6), np_mips(l0, 6))


def test_mips_f32_is_the_plain_box_mean(orc):
    rng = np.random.default_rng(5)
    l0 = rng.random((32, 32, 32)).astype(np.float32)
    chain = orc.mips_f32(l0, 3)
    l1 = l0.reshape(16, 2, 16, 2, 16, 2).mean(axis=(1, 3, 5))
    assert np.allclose(chain[32 ** 3:32 ** 3 + 16 ** 3].reshape(16, 16, 16), l1, rtol=1e-6)


def test_cone_lod_sequence(pkg, orc):
    tp = pkg.default_trace_params()
    lods, hs = orc.cone_lods(tp)
    # LOD 0 for steps 1-6, then 0.04, 0.61, 1.18, 1.75, 2.31, 2.88, 3.45, 4.02, 4.59, 5.16; reach ~37 voxels
    assert (lods[:6] == 0).all()
    assert np.allclose(lods[6:], [0.04, 0.61, 1.18, 1.75, 2.31, 2.88, 3.45, 4.02, 4.59, 5.16], atol=0.01)
    assert abs(hs[-1] - 36.9) < 0.2 and abs(hs[0] - 0.1) < 1e-7
    assert np.allclose(hs[1:] / hs[:-1], 1.0 + np.tan(0.45), rtol=1e-5)


def _probe_scene(pkg, scenes, noise_on, cone_on):
    s = scenes.make_scene("tiny")
    s.tp.doNoiseSample, s.tp.doConeTrace = int(noise_on), int(cone_on)
    return s


def test_cone_sum_in_a_full_volume_is_8_5(pkg, scenes, orc):
    """(ii) constant volume -> traceCone = sum_{i=1..S} i/(S*k) = (S+1)/(2k) = 8.5 at the defaults"""
    s = _probe_scene(pkg, scenes, False, True)
    full = np.full(orc.chain_size(32, 4), 255, np.uint8)
    centre = (25.0, 0.0, 0.0)
    ok, col = orc.conetrace_fragment(s, full, centre, (0.5, 0.5), centre, 1.0)
    assert ok and np.allclose(col, 8.5, rtol=1e-6)               # doNoise off: color = vec4(indirect)
    s.tp.vctSteps, s.tp.vctDownScaling = 22, 1.52                # the paper's res3.png overlay parameters
    ok, col = orc.conetrace_fragment(s, full, centre, (0.5, 0.5), centre, 1.0)
    assert np.allclose(col, 23 / (2 * 1.52), rtol=1e-6)


def test_empty_volume_blacks_the_colour_but_keeps_alpha(pkg, scenes, orc):
    """(iii) empty volume -> indirect = 0 -> rgb = 0, noise alpha intact"""
    s = _probe_scene(pkg, scenes, True, True)
    empty = np.zeros(orc.chain_size(32, 4), np.uint8)
    centre = (25.0, 0.0, 0.0)
    frag = (25.0, 0.3, 0.2)
    ok, col = orc.conetrace_fragment(s, empty, frag, (0.6, 0.65), centre, 1.5)
    assert ok and (col[:3] == 0).all() and col[3] > 0
    s2 = _probe_scene(pkg, scenes, True, False)
    ok2, col2 = orc.conetrace_fragment(s2, empty, frag, (0.6, 0.65), centre, 1.5)
    assert ok2 and col2[3] == col[3] and 0.2 <= col2[0] <= 0.65 + 1e-6     # minNoiseColor .. + noiseColorScale


def test_fragment_discards(pkg, scenes, orc):
    s = _probe_scene(pkg, scenes, True, True)
    chain = np.zeros(orc.chain_size(32, 4), np.uint8)
    centre = (25.0, 0.0, 0.0)
    # camera looks down +X, so the quad lies in the YZ plane; a point at distance >= r from the centre is discarded
    assert not orc.conetrace_fragment(s, chain, (25.0, 1.0, 0.0), (1.0, 0.5), centre, 1.0)[0]
    assert not orc.conetrace_fragment(s, chain, (25.0, 0.9999, 0.0), (1.0, 0.5), centre, 1.0)[0]      # disc = 4(r^2-d^2) < 0.01
    assert orc.conetrace_fragment(s, chain, (25.0, 0.99, 0.0), (0.995, 0.5), centre, 1.0)[0]


def test_single_billboard_lights_the_sun_facing_cap(pkg, scenes, orc):
    """(i) one billboard, sun on the +X axis: the lit set hugs the sun-facing hemisphere"""
    s = scenes.make_scene("tiny", boards=1)
    s.board_pos[:] = 0.0
    s.board_scale[:] = 2.0
    s.sun.position[:] = (125.0, 0.0, 0.0)
    s.width, s.height = 256, 256
    posmap, depth, l0 = orc.voxelize(s)
    D, vox = 32, 10.0 / 32
    lit = np.argwhere(l0 > 0)                                   # (z, y, x)
    assert len(lit) > 50
    centres = (lit[:, ::-1] + 0.5) * vox - 5.0                  # voxel centres relative to the billboard centre, (x,y,z)
    rad = np.linalg.norm(centres, axis=1)
    assert (np.abs(rad - 2.0) <= vox * (np.sqrt(3) / 2 + 1.0) + 1e-6).all(), "a lit voxel is off the sphere surface"
    assert (centres[:, 0] >= -2 * vox).all(), "a lit voxel is on the far side from the sun"
    pole = int((2.0 + 5.0) / vox)                               # voxel holding the pole point c + r*x
    assert l0[D // 2, D // 2, pole] == 255 or l0[D // 2 - 1, D // 2 - 1, pole] == 255
    # the valid texels form the disc of radius ~r: area pi*(r*sqrt(1-1e-4))^2 in a 20x20 ortho window
    frac = (posmap[..., 3] > 0).mean()
    assert abs(frac - np.pi * 4.0 / 400.0) < 2e-3
    # depth is the radial distance from Sun::nearPlane: smallest at the pole
    d = np.where(posmap[..., 3] > 0, depth, 1.0)
    j, i = np.unravel_index(np.argmin(d), d.shape)
    assert abs(i - 127.5) <= 1 and abs(j - 127.5) <= 1


def test_second_pass_store_pattern(pkg, scenes, orc):
    s = scenes.make_scene("tiny")
    idx = orc.second_voxelize_indices(s.vol, (25.0 + 0.01, 0.01, 0.01))     # just above the volume centre
    assert (idx[0] == 16).all()
    vox = 10.0 / 32
    step = vox / np.sqrt(3.0)                                   # stepSize * normalize(vec3(1,1,1)).x
    assert step < vox
    signs = np.array([[sx, sy, sz] for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)])
    exp = np.where(signs > 0, 16, 15)                           # +delta stays in voxel 16, -delta drops to 15
    assert np.array_equal(idx[1:], exp)
    # truncation toward zero: a point 0.5 voxel below the lower bound still lands in voxel 0; outside -> dropped
    idx = orc.second_voxelize_indices(s.vol, (20.0 - 0.4 * vox, 0.0, 0.0))
    assert idx[0][0] == 0
    idx = orc.second_voxelize_indices(s.vol, (20.0 - 2.5 * vox, 0.0, 0.0))
    assert (idx[0] == -1).all()
    idx = orc.second_voxelize_indices(s.vol, (30.0 + 0.01, 0.0, 0.0))       # == D -> dropped by the image store
    assert (idx[0] == -1).all()


def test_sort_boards_is_far_to_near(pkg, scenes, orc):
    s = scenes.make_scene("C1")
    p, sc_ = orc.sort_boards(s.board_pos, s.board_scale, s.vol.position, s.cam.position)
    d = orc.board_distances(p, s.vol.position, s.cam.position)
    assert (np.diff(d) <= 0).all()
    d0 = orc.board_distances(s.board_pos, s.vol.position, s.cam.position)
    assert len(np.unique(d0)) == len(d0), "fixture has distance ties; order would be ambiguous"
    order = np.argsort(-d0, kind="stable")
    assert np.array_equal(p, s.board_pos[order]) and np.array_equal(sc_, s.board_scale[order])


def test_empty_and_clipped_inputs(pkg, scenes, orc):
    s = scenes.make_scene("tiny", boards=0)
    posmap, depth, l0 = orc.voxelize(s)
    assert not posmap.any() and (depth == 1.0).all() and not l0.any()
    img, u8, st = orc.cone_trace(s, orc.mips(l0, 4))
    assert st.fragments == 0 and np.allclose(img, [0.2, 0.3, 0.5, 1.0])
    s = scenes.make_scene("tiny", boards=3)
    s.board_pos[:] = [[-40.0, 0, 0], [-30.0, 1, 1], [500.0, 0, 0]]        # behind the camera / beyond the light's far plane
    posmap, depth, l0 = orc.voxelize(s)
    img, _, st = orc.cone_trace(s, orc.mips(l0, 4))
    assert st.fragments == 0 or st.fragments < 200
    assert np.isfinite(img).all()


def test_r32f_chain_in_the_oracle(pkg, scenes, orc):
    """CRN_VOLUME_R32F (extension): a full float volume still sums to 8.5, and with 0/1 occupancy the float
    mips are exact dyadic rationals (count / 8^l), so R8 and R32F traces agree to the R8 quantisation step"""
    s = _probe_scene(pkg, scenes, False, True)
    s.vol.format = pkg.VOLUME_R32F
    full = np.ones(orc.chain_size(32, 4), np.float32)
    centre = (25.0, 0.0, 0.0)
    ok, col = orc.conetrace_fragment(s, full, centre, (0.5, 0.5), centre, 1.0)
    assert ok and np.allclose(col, 8.5, rtol=1e-6)
    rng = np.random.default_rng(2)
    l0 = (rng.random((32, 32, 32)) < 0.2)
    cf = orc.mips_f32(l0.astype(np.float32), 4)
    assert np.array_equal(cf * 8.0 ** 3 % 1.0, np.zeros_like(cf)), "float mips of a 0/1 volume are multiples of 8^-3"
    c8 = orc.mips((l0 * 255).astype(np.uint8), 4)
    assert np.abs(cf - c8 / 255.0).max() <= 0.5 / 255 * 3 + 1e-6        # one rounding per level


def test_paper_variant_fills_the_sphere(pkg, scenes, orc):
    """§8 f2 (CRN_VOLUME_RG8): one billboard of radius 2 — the occupancy channel is the solid sphere (chords marched
    in steps of one voxel), the lit channel is unchanged and contained in it, and the alpha gate leaves the image alone."""
    s = scenes.make_scene("tiny", boards=1)
    s.board_pos[:] = 0.0
    s.board_scale[:] = 2.0
    s.sun.position[:] = (125.0, 0.0, 0.0)
    s.width, s.height = 256, 256
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    l0p, a0 = orc.voxelize_paper(s)
    assert np.array_equal(l0p, l0)
    assert (a0 >= l0).all()
    D, vox = 32, 10.0 / 32
    z, y, x = np.indices((D, D, D))
    rad = np.linalg.norm((np.stack([x, y, z], -1) + 0.5) * vox - 5.0, axis=-1)
    inside = a0 > 0
    assert inside[rad < 2.0 - vox].all(), "a voxel well inside the sphere is not marked"
    # outside the sphere only the lit shell's +-stepSize/sqrt(3) neighbours (second pass) may be marked
    assert not inside[rad > 2.0 + vox * (np.sqrt(3) / 2 + 1.0) + 1e-6].any()
    vol = inside.sum() * vox ** 3                                # every voxel the solid sphere touches: r .. r + ~1 voxel
    assert 4.0 / 3.0 * np.pi * 2.0 ** 3 < vol < 4.0 / 3.0 * np.pi * (2.0 + 1.2 * vox) ** 3
    chain, chain_a = orc.mips(l0, s.vol.levels), orc.mips(a0, s.vol.levels)
    assert (chain_a >= chain).all()                              # box filter + round-half-up are monotone
    s = scenes.make_scene("tiny")
    l0, a0 = orc.voxelize_paper(s)
    chain, chain_a = orc.mips(l0, s.vol.levels), orc.mips(a0, s.vol.levels)
    plain, _, _ = orc.cone_trace(s, chain, want_u8=False)
    s.vol.format = pkg.VOLUME_RG8
    gated, _, _ = orc.cone_trace(s, np.concatenate([chain, chain_a]), want_u8=False)
    assert np.array_equal(plain.view(np.uint32), gated.view(np.uint32))
