#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -3 gpurun_out/bench_p.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_p.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['serialized_ms_per_step'], d['stages_ms'])
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk != 'workload'})
PY
