#!/bin/bash
mkdir -p gpurun_out
for c in C1 C2 C3; do python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done
python -m pytest tests -m gpu -x -q -k "psnr or edge or paths_agree" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
