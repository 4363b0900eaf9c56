#!/bin/bash
# final round-2 artefacts at N=1: GPU tests, microbenchmarks, the bench line (with extras + CPU baseline), launch list, full ncu capture of the trace kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
python profiles/microbench.py > gpurun_out/microbench_r02.json 2> gpurun_out/microbench.err
python bench.py --steps 20 --warmup 3 --save gpurun_out/bench_r02_C3_n1.json > /dev/null 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
for cfg in C1 C2 C4 C5; do python bench.py --config $cfg --steps 20 --warmup 3 --no-extras --no-cpu-baseline --save gpurun_out/bench_r02_${cfg}_n1.json > /dev/null 2>> gpurun_out/bench_final.err; done
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --radius-mode reference --save gpurun_out/bench_r02_C3refradii_n1.json > /dev/null 2>> gpurun_out/bench_final.err
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --sampler explicit --save gpurun_out/bench_r02_C3explicit_n1.json > /dev/null 2>> gpurun_out/bench_final.err
bash profiles/run_l.sh final
bash profiles/run_s.sh final
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_r02_*_n1.json')):
    d = json.load(open(f))
    print(f.split('/')[-1], round(d['value'], 1), 'fps', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), 'trace', round(d['stages_ms']['traceMs'], 3), 'vox+mip', round(d['voxelize_mip_ms'], 3), 'frac', d['roofline'] and round(d['roofline']['frac'], 3))
PY
