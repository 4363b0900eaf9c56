#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python profiles/trace_time.py 2>&1 | tail -1
python profiles/trace_time.py --config C4 --frames 4 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_l.csv \
    python profiles/trace_time.py --frames 1 > gpurun_out/ncu_l.log 2>&1
grep -E "bake_steps|need_code|trace_fast" gpurun_out/launches_l.csv | awk -F'","' '{print $5, $NF}' | tail -3
