#!/bin/bash
mkdir -p gpurun_out
for nl in 1 0; do for c in C3 C1 C2; do echo -n "NO_LATTICE=$nl $c: "; CRN_NO_LATTICE=$nl python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done; done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
