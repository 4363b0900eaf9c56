#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bake_steps|need_code" -c 2 -f -o gpurun_out/bake_g \
    python profiles/trace_time.py --frames 1 > gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log
