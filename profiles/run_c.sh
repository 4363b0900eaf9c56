#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -c 1 -f -o gpurun_out/trace_c \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c.log
