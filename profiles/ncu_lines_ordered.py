"""Per-source-line instruction counts of an .ncu-rep in SOURCE ORDER (needs -lineinfo), with a per-unit normaliser.
    python profiles/ncu_lines_ordered.py gpurun_out/xxx.ncu-rep [units] [min share %]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
per = defaultdict(lambda: [0, 0, ""]); tot = 0; tots = 0
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[0].strip().isdigit(): continue
    try: n = int(float(r[ci] or 0)); s = int(float(r[cs] or 0))
    except ValueError: continue
    per[int(r[0])][0] += n; per[int(r[0])][1] += s; per[int(r[0])][2] = r[1]; tot += n; tots += s
print(f"total warp instructions {tot:,}, samples {tots:,}")
for k in sorted(per):
    n, s, src = per[k]
    if 100.0 * n / tot >= floor: print(f"{k:4d} {100*n/tot:5.2f}% inst {100*s/max(tots,1):5.2f}% samp {n/unit:8.1f}  {src.strip()[:110]}")
