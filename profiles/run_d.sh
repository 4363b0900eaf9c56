#!/bin/bash
mkdir -p gpurun_out
python profiles/microbench.py > gpurun_out/microbench_r02.json 2> gpurun_out/microbench.err; grep -E "slice centres" gpurun_out/microbench_r02.json
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_d.json'))
print(d['value'], d['ms_per_step'], d['stages_ms'], d.get('roofline_tex'))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_fast -c 1 -f -o gpurun_out/trace_d \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
