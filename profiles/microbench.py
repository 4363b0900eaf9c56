"""Measures the ceilings the stages are bounded by and that MEASURED_PEAKS.json lacks
(tex fetch rate, L1-hit load rate, atomic rates).  Run on the GPU box:
    python profiles/microbench.py > gpurun_out/microbench.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.import_package()
out = {pkg.MICROBENCH[k]: round(pkg.microbench(k), 2) for k in sorted(pkg.MICROBENCH)}
out["unit"] = "giga lane-operations/s, best of 5, CUDA events, 148x32 CTAs x 256 threads x 256 ops"
print(json.dumps(out, indent=1))
