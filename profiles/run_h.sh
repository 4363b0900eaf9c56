#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -5 gpurun_out/bench_h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_h.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['serialized_ms_per_step'], d['stages_ms'])
print(d['roofline'])
for k,v in d['extra'].items(): print(k, v)
print(d.get('cpu_baseline'))
PY
