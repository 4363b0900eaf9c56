#!/bin/bash
# launch list (per-kernel durations, serialised) of two bench steps -> gpurun_out/launches_$1.csv
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${1:-l}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launches_${1:-l}.log 2>&1
tail -1 gpurun_out/launches_${1:-l}.log | head -c 200
