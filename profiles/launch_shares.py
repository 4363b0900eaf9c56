"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum).   python profiles/launch_shares.py gpurun_out/launches_x.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        agg[d['Kernel Name'][:70]][0] += 1; agg[d['Kernel Name'][:70]][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {v[0]:4d} launches {v[1]/1e3:10.1f} us total {v[1]/1e3/v[0]:9.1f} us each {100*v[1]/tot:5.1f}%")
