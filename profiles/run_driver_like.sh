#!/bin/bash
# what the driver runs at round end, in its order: reference arm, smoke, default bench line
mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err; tail -4 gpurun_out/ref_arm.err | head -3; head -c 300 gpurun_out/ref_arm.json; echo
python __graft_entry__.py smoke 2>&1 | tail -1
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'traffic', d['roofline']['traffic'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['gpu_launches'])
PY
