#!/bin/bash
mkdir -p gpurun_out
for s in 1 2 4 8; do for c in C1 C2; do echo -n "seg $s: "; CRN_TRACE_SEGMENTS=$s python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done; done
echo -n "C3 seg 2: "; CRN_TRACE_SEGMENTS=2 python profiles/trace_time.py --config C3 --frames 4 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
