#!/bin/bash
# quick loop: trace times of the current build + the GPU tests
mkdir -p gpurun_out
for c in ${1:-C3 C2 C1}; do echo -n "$c: "; python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
