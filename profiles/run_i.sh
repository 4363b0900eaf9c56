#!/bin/bash
mkdir -p gpurun_out
for e in 0 1; do
  if [ $e = 1 ]; then export BENCH_NO_SMI=1; fi
  python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_i$e.json 2> gpurun_out/bench_i$e.err; tail -3 gpurun_out/bench_i$e.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_i$e.json'))
print($e, d['value'], d['ms_per_step'], d['e2e']['value'], d['serialized_ms_per_step'], d['clocks'])
PY
done
