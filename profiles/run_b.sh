#!/bin/bash
# full bench line of the current build (with the extra configs) -> gpurun_out/bench_$1.json
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${1:-b}.json 2> gpurun_out/bench_${1:-b}.err; tail -3 gpurun_out/bench_${1:-b}.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${1:-b}.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('serialized_ms_per_step'), d['stages_ms'])
print({k: v for k, v in d['roofline'].items() if k in ('achieved','peak','frac','passes_per_launch','noise_bilinear','baked_cone_bilinear','textureLod_trilinear_equiv')})
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk != 'workload'})
PY
