#!/bin/bash
mkdir -p gpurun_out
for cap in 0 15 14 12; do
CRN_TRACE_CTAS=$cap python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q.json'))
print($cap, d['value'], d['ms_per_step'], d['e2e']['value'], d['serialized_ms_per_step'], d['stages_ms']['traceMs'])
PY
done
