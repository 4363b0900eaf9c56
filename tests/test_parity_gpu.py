"""Parity of the CUDA path (through the C-ABI) against the CPU oracle, same seeded inputs.
Bars: position map, voxel occupancy, mip chain, billboard order and bins bit-exact;
images PSNR >= 45 dB with the max per-channel error stated."""
import numpy as np
import pytest

from conftest import psnr, steady_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tiny", "small", "C1"])
def test_voxelize_exact(name, pkg, scenes, orc, renderer):
    s = steady_state(scenes.make_scene(name), orc)
    renderer.keep_position_map(True)
    renderer.set_scene(s)
    renderer.voxelize()
    posmap = renderer.read_position_map()
    ref_posmap, ref_depth, ref_l0 = orc.voxelize(s)
    assert np.array_equal(posmap.view(np.uint32), ref_posmap.view(np.uint32)), "position map differs"
    l0 = renderer.read_volume(0)
    assert np.array_equal(l0, ref_l0), f"occupancy differs in {(l0 != ref_l0).sum()} voxels"
    assert renderer.count_active_voxels() == int((ref_l0 > 0).sum())
    chain = renderer.read_chain()
    assert np.array_equal(chain, orc.mips(ref_l0, s.vol.levels)), "mip chain differs"
    renderer.keep_position_map(False)


@pytest.mark.parametrize("sampler", ["explicit", "texture"])
@pytest.mark.parametrize("name", ["tiny", "small", "C1"])
def test_image_psnr(name, sampler, pkg, scenes, orc, renderer):
    s = steady_state(scenes.make_scene(name), orc)
    s.tp.sampler = pkg.SAMPLER_TEXTURE if sampler == "texture" else pkg.SAMPLER_EXPLICIT
    renderer.set_scene(s)
    renderer.voxelize()
    img = renderer.cone_trace(fmt=pkg.IMAGE_RGBA32F)
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    ref, ref_u8, st = orc.cone_trace(s, orc.mips(l0, s.vol.levels))
    p = psnr(img, ref)
    err = float(np.abs(img - ref).max())
    print(f"{name}/{sampler}: PSNR {p:.2f} dB, max per-channel error {err:.3e}")
    assert p >= 45.0
    u8 = renderer.cone_trace(fmt=pkg.IMAGE_RGBA8)
    d = np.abs(u8.astype(np.int32) - ref_u8.astype(np.int32))
    print(f"{name}/{sampler}: RGBA8 max diff {d.max()} LSB, {100.0 * (d > 0).mean():.3f}% of channels differ")
    assert d.max() <= (2 if sampler == "explicit" else 6)
