"""The committed bench lines (profiles/bench_r02_*.json, written by `bench.py --save` on the GPU box) carry every key the
bench contract names, and the reference arm (`bench.py --impl reference`, CPU only) prints a line of the same shape without
loading the product library.  No GPU needed."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_*.json"))), ids=os.path.basename)
def test_committed_bench_lines_follow_the_contract(path):
    d = json.loads(open(path).read())
    for k in BASE + ("clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, f"{k} missing"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - (d["n_gpus"] if d["scaling"] == "weak" else 1)) < 1e-6 * d["n_gpus"] + 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert abs(e["value"] - d["value"]) > 1e-9, "e2e must be measured, not copied from value"
    c = d["clocks"]
    if c:
        assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    if r:                                               # texture sampler: the texture-pipe roofline of the trace kernel
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r
        assert r["bound"] == "tex" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]


@pytest.mark.timeout(600)
def test_reference_arm_line_and_independence():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in BASE + ("impl", "cpu_baseline", "e2e"):
        assert k in d
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
