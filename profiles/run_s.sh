#!/bin/bash
# full ncu capture (with source) of the C3 trace kernel -> gpurun_out/trace_$1.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_fast_kernel -s 3 -c 1 -f -o gpurun_out/trace_${1:-s} \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_${1:-s}.log 2>&1
tail -3 gpurun_out/ncu_${1:-s}.log | head -c 300
