#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print(d['value'], d['ms_per_step'], d['stages_ms'], d['per_frame'], d.get('roofline_tex'))
PY
