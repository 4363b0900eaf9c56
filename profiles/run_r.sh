#!/bin/bash
# round-2 re-entry: GPU tests, default bench (with extras), launch list of one bench step
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -3 gpurun_out/bench_r.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('serialized_ms_per_step'), d['stages_ms'], d['roofline'])
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk != 'workload'})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launches_r.log 2>&1
tail -2 gpurun_out/launches_r.log | head -c 400
