"""Per-source-line instruction counts and stall samples from the source page of an .ncu-rep (needs -lineinfo).
    python profiles/ncu_lines.py gpurun_out/xxx.ncu-rep [top N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hdr_i]
    c_inst, c_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    per = defaultdict(lambda: [0, 0, ""])
    tot_i = tot_s = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= c_inst or not r[0].strip().isdigit():
            continue
        # rows are SASS instructions annotated with their CUDA line
        try:
            n, s = int(float(r[c_inst] or 0)), int(float(r[c_samp] or 0))
        except ValueError:
            continue
        key = int(r[0])
        per[key][0] += n; per[key][1] += s; per[key][2] = r[1]
        tot_i += n; tot_s += s
    print(f"total instructions {tot_i:,}  samples {tot_s:,}")
    for k, (n, s, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k:5d} {100.0 * n / max(tot_i, 1):6.2f}% inst {100.0 * s / max(tot_s, 1):6.2f}% samp  {src.strip()[:110]}")


if __name__ == "__main__":
    main()
