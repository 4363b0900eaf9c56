"""Small frames of every code path (both samplers, both volume formats, a Z-slab + crn_finish_mips) for
    compute-sanitizer --tool memcheck|racecheck python profiles/sanitizer_check.py
Round 1 result on B200: memcheck 0 errors, racecheck 0 hazards (profiles/r01_sanitizer.txt)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry
pkg = entry.import_package()
from cloud_renderer_b200 import scene as sc
for name, fmt in (("tiny", 0), ("small", 0), ("small", 1)):
    s = sc.make_scene(name)
    s.vol.format = fmt
    r = pkg.Renderer(0)
    for sampler in (0, 1):
        s.tp.sampler = sampler
        r.set_scene(s); r.voxelize(); img = r.cone_trace()
    r.set_z_slab(0, 16); r.voxelize(); r.finish_mips(min(5, s.vol.levels)) if s.vol.levels > 5 else None
    print(name, fmt, img.mean())
    r.close()
print("SAN_DONE")
