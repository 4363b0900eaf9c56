#!/usr/bin/env python
"""bench.py — frames/s of the per-frame volumetric pipeline (voxelize + mip chain + cone trace).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C2|C1|C4]

One step = one full frame of BASELINE.json's headline configuration (C3: 256^3 animated
volume, 20k billboards, 3840x2160, sun shadow cones): new billboard positions -> voxelize
(pass 1 + pass 2) -> mip chain -> camera sort/bin -> cone trace -> RGBA8 image.
With N > 1 the frames (views of the C5 orbit / time steps) are sharded round-robin over ranks,
volume replicated, no data-path collective: weak scaling.

value  : whole-job frames/s with the step's inputs already in HBM and the image left in HBM.
e2e    : the same through the C-ABI with HOST buffers (pinned billboards in, RGBA8 image out).
--impl reference : the CPU oracle (restated reference, all host threads) on a bounded sample
                   of the same frame.  Not llvmpipe: see BASELINE.md §2.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

CUTOFF = 1.0 / 1024.0          # early-ray-termination transmittance cutoff used by the GPU arm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--cutoff", type=float, default=CUTOFF)
    ap.add_argument("--sampler", default="texture", choices=["texture", "explicit"])
    ap.add_argument("--no-skip", action="store_true", help="fetch every cone sample (no empty-space skipping)")
    ap.add_argument("--slab-exchange", default="chain", choices=["chain", "none"],
                    help="C4 with N>1: 'chain' = Z-slab voxelize+mips and ONE all-gather of the finished chain (north_star); "
                         "'none' = every rank voxelizes the whole volume (no collective), only the trace is sharded")
    ap.add_argument("--volume-format", default="r8", choices=["r8", "r32f", "rg8"],
                    help="r8 = the shipped reference format; rg8 = the paper variant with the occupancy channel")
    ap.add_argument("--radius-mode", default="auto", choices=["auto", "fill", "reference"],
                    help="billboard radii for N > 200: 'fill' keeps the cloud's fill (headline), 'reference' keeps U[1,2.5] (SURVEY 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="back-to-back steps (no L2 flush between them)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def metric_name(config):
    return "cone-traced frames/s @4K with 256^3 volume" if config in ("C3", "C5") else f"cone-traced frames/s ({config})"


def workload_name(sc, config, radius_mode):
    D, L, N, W, H = sc.CONFIGS[config]
    return (f"{config}: {D}^3 R8 volume ({L} levels), {N} billboards ({radius_mode} radii), {W}x{H}, "
            f"animated (one new frame per step), sun shadow cones 16 steps, noise 4 octaves")


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu capture of the dominant kernel (per launch)"""
    try:
        tot = 0.0
        for line in open(os.path.join(ROOT, "profiles", name)):
            f = line.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[f[2]]
        return tot or None
    except Exception:
        return None


def frame_inputs(sc, config, n_frames, rank, world, radius_mode="auto"):
    """billboard offsets for the frames this rank renders + the camera/sun of each"""
    base = sc.make_scene(config, cutoff=0.0, radius_mode=radius_mode)
    frames = []
    for k in range(n_frames):
        g = k * world + rank                          # global frame id, round-robin over ranks
        if world > 1:
            s = sc.make_scene(config, frame=g, view=g % 64, radius_mode=radius_mode)
        else:
            s = sc.make_scene(config, frame=g, radius_mode=radius_mode)
        frames.append(s)
    return base, frames


def cpu_baseline(config, orc, sc, budget_rows=64):
    """The oracle on a bounded sample of one frame: voxelize + mips in full, the trace on every
    `stride`-th image row (an unbiased sample of the frame), extrapolated to the frame."""
    s = sc.make_scene(config, frame=1)
    orc.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    t0 = time.perf_counter()
    _, _, l0 = orc.voxelize(s, want_posmap=False)
    t1 = time.perf_counter()
    chain = orc.mips(l0, s.vol.levels)
    t2 = time.perf_counter()
    s.board_pos, s.board_scale = orc.sort_boards(s.board_pos, s.board_scale, s.vol.position, s.cam.position)
    t3 = time.perf_counter()
    H = s.height
    stride = max(1, H // budget_rows)
    rows = list(range(stride // 2, H, stride))
    tt = 0.0
    frags = 0
    for r in rows:
        a = time.perf_counter()
        _, _, st = orc.cone_trace(s, chain, rows=(r, r + 1), want_u8=False)
        tt += time.perf_counter() - a
        frags += st.fragments
    trace_full = tt * H / len(rows)
    total = (t1 - t0) + (t2 - t1) + (t3 - t2) + trace_full
    return {
        "value": 1.0 / total, "unit": "frames/s", "cores": orc.num_threads(), "kind": "port",
        "sample": (f"{config} frame 1: oracle voxelize+mips+sort in full ({t1 - t0:.2f}+{t2 - t1:.2f}+{t3 - t2:.2f} s), cone trace on "
                   f"{len(rows)} of {H} rows ({tt:.2f} s) extrapolated x{H / len(rows):.1f}; every fragment shaded (no early termination)"),
        "voxelize_mip_ms": 1e3 * (t2 - t0), "trace_ms_extrapolated": 1e3 * trace_full,
        "fragments_per_frame_est": int(frags * H / len(rows)),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = entry.import_oracle()
    orc.build()
    entry.import_package()
    from cloud_renderer_b200 import scene as sc
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(args.config, orc, sc, budget_rows=32)
        if i >= args.warmup:
            vals.append(cb["value"])
        if i == 0 and 1.0 / cb["value"] > 60:        # keep the whole run within minutes
            pass
    v = float(np.mean(vals))
    cb["value"] = v
    D, L, N, W, H = sc.CONFIGS[args.config]
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.config), "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(sc, args.config, sc.make_scene(args.config).meta["radius_mode"]),
                   "implementation": "CPU oracle: the reference's algorithm restated in C++/OpenMP (oracle/oracle.cpp), all host threads; "
                                     "the reference's GL 4.4 executable cannot be built or run here (BASELINE.md 2), not llvmpipe"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cloud-renderer_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    real_stdout = os.dup(1)
    os.dup2(2, 1)                      # NCCL / libraries may print to stdout; the JSON line goes to the real one
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.import_package()
    from cloud_renderer_b200 import scene as sc
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(dev)            # torch's default stream has handle 0, which the C-ABI reads as "create your own":
    torch.cuda.set_stream(stream)              # use one explicit stream for torch's events / flush / NCCL waits AND the library's kernels
    r = pkg.Renderer(local, stream.cuda_stream)

    K, Wm = args.steps, args.warmup
    slab = args.config == "C4" and world > 1
    base, frames = frame_inputs(sc, args.config, K + Wm, 0 if slab else rank, 1 if slab else world, args.radius_mode)
    D, L, N, Wd, Ht = sc.CONFIGS[args.config]
    for f in frames:
        f.tp.transmittanceCutoff = args.cutoff
        f.tp.sampler = pkg.SAMPLER_TEXTURE if args.sampler == "texture" else pkg.SAMPLER_EXPLICIT
        f.tp.skipEmptySpace = 0 if args.no_skip else 1
        f.vol.format = {"r8": pkg.VOLUME_R8, "r32f": pkg.VOLUME_R32F, "rg8": pkg.VOLUME_RG8}[args.volume_format]
    r.set_scene(frames[0])

    # resident inputs (value leg) and pinned host inputs (e2e leg)
    d_pos = [torch.from_numpy(f.board_pos).to(dev) for f in frames]
    d_scale = torch.from_numpy(frames[0].board_scale).to(dev)
    h_pos = [torch.from_numpy(f.board_pos).pin_memory() for f in frames]
    h_scale = torch.from_numpy(frames[0].board_scale).pin_memory()
    d_img = torch.empty((Ht, Wd, 4), dtype=torch.uint8, device=dev)
    h_img = torch.empty((Ht, Wd, 4), dtype=torch.uint8).pin_memory()
    h_img2 = torch.empty((Ht, Wd, 4), dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    slab_mode = args.config == "C4" and world > 1
    gather_mode = slab_mode and args.slab_exchange == "chain"
    if slab_mode:
        r.set_tile_row_interleave(rank, world)              # balanced: 16-row tile rows dealt round-robin
    if gather_mode:
        # C4: ONE frame per step for the whole job.  Every rank voxelizes + mips its Z-slab, one
        # all-gather of the finished slab-local levels, replicated top levels, then each rank traces
        # its band of image rows (no image gather: the bands stay on their GPUs).
        from cloud_renderer_b200 import sharding as sh
        z0, z1 = sh.z_slab(D, rank, world)
        r.set_z_slab(z0, z1)
        r.voxelize()                                        # allocates bits + chain
        torch.cuda.synchronize()
        tens = sh.chain_tensors(torch, r, L, dev)
        nloc = sh.slab_local_levels(L)
        views = [sh.slab_view(t, rank, world) for t in tens[:1 + nloc]]     # bits + levels 0..4

    def step(i, host):
        f = frames[i]
        r.set_camera(f.cam); r.set_sun(f.sun); r.set_trace_params(f.tp)
        if host:
            r.set_billboards(h_pos[i].numpy(), h_scale.numpy())
        else:
            r.set_billboards(d_pos[i], d_scale)
        r.voxelize()
        if gather_mode:
            sh.all_gather_levels(dist, views)
            if L > nloc:
                r.finish_mips(nloc)
        r.cone_trace(h_img.numpy() if host else d_img, pkg.IMAGE_RGBA8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host, sampler=None):
        for i in range(Wm):
            step(i, host)
        barrier()
        if sampler:
            sampler.start()
        l0 = r.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        stage = {k: 0.0 for k in ("prepSortMs", "lightBinMs", "voxelizeMs", "mipMs", "camBinMs", "traceMs")}
        r.set_timing(True)
        for i in range(K):
            if not args.no_flush:
                flush.fill_(i & 0xFF)                  # > L2 (126 MB), outside the event pair
            ev[i][0].record(stream)
            step(Wm + i, host)
            ev[i][1].record(stream)
            t = r.timings()                            # syncs; per-stage CUDA-event times of this frame
            for k in stage:
                stage[k] += getattr(t, k)
        r.set_timing(False)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = sum(a.elapsed_time(b) for a, b in ev)
        launches = r.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], dtype=torch.int64, device=dev)
            dist.all_reduce(lt)
            launches = int(lt.item())
        return ms, {k: v / K for k, v in stage.items()}, launches, clocks

    def frame_host(i, out):
        """one frame through the public API with HOST buffers: pinned billboards in, RGBA8 image out (pipelined read-back)"""
        f = frames[i]
        r.set_camera(f.cam); r.set_sun(f.sun); r.set_trace_params(f.tp)
        r.set_billboards(h_pos[i].numpy(), h_scale.numpy())
        r.voxelize()
        if gather_mode:
            sh.all_gather_levels(dist, views)
            if L > nloc:
                r.finish_mips(nloc)
        r.cone_trace_async(out.numpy(), pkg.IMAGE_RGBA8)

    def timed_e2e():
        for i in range(Wm):
            frame_host(i, h_img if i & 1 else h_img2)
        r.wait_images()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(K):
            if not args.no_flush:
                flush.fill_(i & 0xFF)                  # inside the timed region here: conservative
            frame_host(Wm + i, h_img if i & 1 else h_img2)
        r.wait_images()                                # every image has landed in host memory
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    ms, stage, launches, clocks = timed(False, sampler)
    ms_e2e = timed_e2e()

    # fragments / samples actually shaded in one frame (stats pass, outside the timed region)
    r.set_stats(True)
    step(Wm, False)
    st = r.trace_stats()
    r.set_stats(False)
    frag, cone, noise, bins = st.fragments, st.coneSamples, st.noiseSamples, st.binEntries
    cone_skipped, fetches = st.coneSamplesSkipped, st.filteredFetches
    tex_peak = bil_peak = None
    if rank == 0 and args.sampler == "texture":
        try:
            tex_peak = min(pkg.microbench(0, local), pkg.microbench(1, local))     # G trilinear fetches/s, measured now
            bil_peak = pkg.microbench(6, local)                                     # G bilinear (layered 2D) fetches/s
        except Exception:
            tex_peak = bil_peak = None

    if rank == 0:
        job_frames = K if slab_mode else world * K          # C4 shards ONE frame per step over all ranks
        fps = job_frames / (ms * 1e-3)
        fps_e2e = job_frames / (ms_e2e * 1e-3)
        peak, peak_src = load_peaks()
        # dominant kernel = cone trace.  ALGORITHMIC bytes per launch: every cone tap
        # reads 8 texels per level (16 when two levels blend), every noise tap 8 RGBA8 texels, plus the image.
        trace_ms = stage["traceMs"]
        alg_bytes = (cone - cone_skipped) * 16 * 1 + noise * 8 * 4 + Wd * Ht * 4
        achieved = alg_bytes / (trace_ms * 1e-3) / 1e9
        traffic = ncu_traffic("r01_trace_texture_final.txt" if args.sampler == "texture" else "r01_trace_explicit.txt")
        out = {
            "metric": metric_name(args.config),
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong" if slab_mode else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(sc, args.config, frames[0].meta['radius_mode']),
                "sharding": ("Z-slab voxelize+mips, one all-gather of the finished chain (NCCL), tile-row-interleaved trace" if gather_mode else
                             "voxelize+mips replicated on every rank (no collective), tile-row-interleaved trace" if slab_mode else
                             "frames round-robin over ranks, volume replicated, no collective" if world > 1 else "single GPU"),
                "transmittance_cutoff": args.cutoff, "sampler": args.sampler, "skip_empty_space": not args.no_skip,
                "volume_format": args.volume_format,
                "l2": "none (back to back)" if args.no_flush else "256 MiB fill between steps, outside the per-step event pairs",
            },
            "clocks": clocks,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": N * 16, "d2h_bytes_per_step": Wd * Ht * 4,
                    "how": "crn_set_billboards(pinned host) + crn_voxelize + crn_cone_trace_async(pinned host RGBA8) per frame, "
                           "crn_wait_images at the end; the read-back of frame k overlaps the kernels of frame k+1"},
            "gpu_launches": launches,
            "stages_ms": stage, "voxelize_mip_ms": stage["lightBinMs"] + stage["voxelizeMs"] + stage["mipMs"],
            "per_frame": {"fragments_shaded": frag, "cone_samples": cone, "cone_samples_skipped_as_empty": cone_skipped,
                          "noise_samples": noise, "bin_entries": bins,
                          "cone_samples_per_s": cone / (trace_ms * 1e-3), "filtered_samples_per_s": (cone + noise) / (trace_ms * 1e-3)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "trace_kernel", "peak_source": peak_src,
                         "note": ("algorithmic texel bytes (SURVEY 8d: 16 B per fetched cone sample, 32 B per noise tap, + the image) are served "
                                  "by L1 at a 99.9% hit rate, so this fraction can exceed 1 and HBM is NOT the binding roof (traffic = DRAM bytes "
                                  "of one ncu capture, profiles/); the binding roof is the texture pipe: see roofline_tex")},
            # texture-pipe roofline: the cone trace's volume lookups are 3D trilinear (two bilinear passes in the texture
            # unit), the noise taps are single bilinear passes on the layered slice-pair texture; each kind is charged at
            # its own measured ceiling and the fraction is (time at the ceilings) / (measured kernel time)
            "roofline_tex": None if not tex_peak else {
                "bound": "tex", "unit": "G bilinear passes/s",
                "achieved": (2 * (fetches - noise) + noise) / (trace_ms * 1e-3) / 1e9, "peak": bil_peak,
                "frac": ((fetches - noise) / tex_peak + noise / bil_peak) / 1e9 / (trace_ms * 1e-3),
                "cone_trilinear_lookups_per_launch": fetches - noise, "noise_bilinear_lookups_per_launch": noise,
                "trilinear_peak": tex_peak,
                "peak_source": "crn_microbench, measured in this run: tex2DLayered RGBA8 bilinear (peak) and tex3D trilinear "
                               "(min of RGBA8 32^3 and R8 256^3) for the volume lookups"},
        }
        if not args.no_cpu_baseline and world == 1:
            orc = entry.import_oracle()
            orc.build()
            out["cpu_baseline"] = cpu_baseline(args.config, orc, sc)
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
