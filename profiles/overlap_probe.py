"""How much of the frame set-up hides behind the trace?  Pipelined C3 frames (device-resident animation) timed with one
event pair: (A) full frames, (B) the same without the L2 flush, (C) trace only (volume and billboards unchanged: the
light side, baked steps and need codes are reused; the camera-side prep/sort/bin still runs).
    python profiles/overlap_probe.py [--config C3] [--frames 40]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--frames", type=int, default=40)
args = ap.parse_args()
pkg = entry.import_package()
from cloud_renderer_b200 import scene as sc

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
r = pkg.Renderer(0, stream.cuda_stream)
f0 = sc.make_scene(args.config, frame=0)
f0.tp.transmittanceCutoff = 1.0 / 1024.0
f0.tp.sampler = pkg.SAMPLER_TEXTURE
r.set_scene(f0)
r.voxelize()
img = torch.empty((f0.height, f0.width, 4), dtype=torch.uint8, device="cuda")
r.cone_trace(img, pkg.IMAGE_RGBA8)
flush = torch.empty(int(1.25 * torch.cuda.get_device_properties(0).L2_cache_size), dtype=torch.uint8, device="cuda")

def leg(full, do_flush, n):
    r.set_billboards(f0.board_pos, f0.board_scale); r.sync()
    for k in range(6):
        if full: r.animate_billboards(0.2 * k / 60.0); r.voxelize()
        r.cone_trace_enqueue(pkg.IMAGE_RGBA8)
    r.wait_images(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(n):
        if do_flush: flush.fill_(k & 0xFF)
        if full: r.animate_billboards(0.2 * (k + 6) / 60.0); r.voxelize()
        r.cone_trace_enqueue(pkg.IMAGE_RGBA8)
    r.wait_images()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for _ in range(2):
    print(f"A full+flush {leg(True, True, args.frames):.3f} ms   B full {leg(True, False, args.frames):.3f} ms   C trace only {leg(False, False, args.frames):.3f} ms   "
          f"C+flush {leg(False, True, args.frames):.3f} ms")
r.close()
