#!/bin/bash
# round-2 exploratory run: microbenchmarks, baseline bench, source-level ncu of the trace kernel
mkdir -p gpurun_out
python profiles/microbench.py > gpurun_out/microbench_r02.json 2> gpurun_out/microbench.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -c 1 -f -o gpurun_out/trace_base \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_base.log 2>&1
cat gpurun_out/microbench_r02.json; cat gpurun_out/bench_base.json | head -c 600
