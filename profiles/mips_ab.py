"""A/B of the mip kernel's level-0 write paths at 512^3 (C4) and 256^3 (C3): run under ncu for kernel time + DRAM bytes.
    python profiles/mips_ab.py MODE      MODE = tex | tex+linear | linear        (CRN_MIPS_TMA=1 selects the TMA store for the linear copy)
Also checks the linear level 0 against the bit set."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

mode = sys.argv[1] if len(sys.argv) > 1 else "tex"
cfg = sys.argv[2] if len(sys.argv) > 2 else "C4"
pkg = entry.import_package()
from cloud_renderer_b200 import scene as sc

s = sc.make_scene(cfg, frame=1)
s.tp.sampler = pkg.SAMPLER_EXPLICIT if mode == "linear" else pkg.SAMPLER_TEXTURE
r = pkg.Renderer(0)
r.set_scene(s)
if "linear" in mode:
    r.volume_level_ptr(0)                       # from now on every voxelize writes the linear level 0
r.set_timing(True)
for k in range(4):
    r.voxelize()
    r.sync()
t = r.timings()
print(f"{cfg} {mode} tma={os.environ.get('CRN_MIPS_TMA', '0')}: mips+masks stage {t.mipMs:.3f} ms")
if "linear" in mode:
    l0 = r.read_volume(0)
    lit = r.count_active_voxels()
    assert int((l0 == 255).sum()) == lit and int((l0 != 0).sum()) == lit, "linear level 0 does not match the bit set"
    import ctypes as C
    ex = r.export_voxels(0)
    D = s.vol.dimension
    print(f"   linear level 0 ok: {lit} lit voxels, all 255")
r.close()
