"""Small frames of every code path (both samplers, all volume formats, a Z-slab + crn_finish_mips, device-side
generation + animation, pipelined frames) for
    compute-sanitizer --tool memcheck|racecheck python profiles/sanitizer_check.py
Round 1 result on B200: memcheck 0 errors, racecheck 0 hazards (profiles/r01_sanitizer.txt); round 2: profiles/r02_sanitizer.txt."""
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
# 128^3 volume, fill radii: fine cone steps + need-code texture + baked steps + noise lattice + segmented tile lists
s = sc.make_scene("C2", boards=600, size=(640, 360))
s.tp.sampler = 1
r = pkg.Renderer(0)
r.set_scene(s)
for k in range(2):
    r.voxelize(); img = r.cone_trace()
print("C2crop", img.mean())
r.close()
# paper variant (interior march, second chain, gated trace) and device-side generation / animation, pipelined frames
s = sc.make_scene("small")
s.vol.format = 2
r = pkg.Renderer(0)
out = [np.empty((s.height, s.width, 4), np.uint8) for _ in range(2)]
for sampler in (0, 1):
    s.tp.sampler = sampler
    r.set_scene(s)
    r.regenerate_billboards(s.n_boards, (-2.5,) * 3, (2.5,) * 3, 1.0, 2.5, 1.0, 7)
    for k in range(3):
        r.animate_billboards(0.1 * k); r.voxelize(); r.cone_trace_async(out[k & 1])
    r.wait_images()
print("rg8", out[0].mean(), r.read_volume_alpha(0).mean())
r.close()
print("SAN_DONE")
