#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_fast_kernel -s 3 -c 1 -f -o gpurun_out/trace_v \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_v.log 2>&1
tail -3 gpurun_out/ncu_v.log | head -c 300
