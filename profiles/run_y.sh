#!/bin/bash
mkdir -p gpurun_out
for s in 1 2 4; do for c in C3 C2 C1; do echo -n "segments $s $c: "; CRN_TRACE_SEGMENTS=$s python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done; done
