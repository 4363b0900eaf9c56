"""Quick A/B timer: C3 (or --config) frames through the library selected by CRN_LIB, per-stage CUDA-event times.
    CRN_LIB=.../libcloud_renderer_b200_<tag>.so python profiles/trace_time.py [--config C3] [--frames 8]"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--cutoff", type=float, default=1.0 / 1024.0)
ap.add_argument("--radius-mode", default="auto")
ap.add_argument("--view", type=int, default=None)
args = ap.parse_args()
pkg = entry.import_package()
from cloud_renderer_b200 import scene as sc

r = pkg.Renderer(0)
frames = [sc.make_scene(args.config, frame=k, radius_mode=args.radius_mode, view=args.view) for k in range(args.frames + 2)]
for f in frames:
    f.tp.transmittanceCutoff = args.cutoff
    f.tp.sampler = pkg.SAMPLER_TEXTURE
r.set_scene(frames[0])
import numpy as np
img = np.empty((frames[0].height, frames[0].width, 4), np.uint8)
r.set_timing(True)
tr, vx = [], []
for k, f in enumerate(frames):
    r.set_trace_params(f.tp); r.set_camera(f.cam); r.set_sun(f.sun)
    r.set_billboards(f.board_pos, f.board_scale)
    r.voxelize()
    r.cone_trace(img, pkg.IMAGE_RGBA8)
    t = r.timings()
    if k >= 2:
        tr.append(t.traceMs); vx.append(t.lightBinMs + t.voxelizeMs + t.mipMs)
print(f"{os.path.basename(os.environ.get('CRN_LIB', 'default')):40s} trace ms median {statistics.median(tr):.3f} min {min(tr):.3f}   voxelize+mip {statistics.median(vx):.3f}")
r.close()
