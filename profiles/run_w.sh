#!/bin/bash
mkdir -p gpurun_out
for v in "" _m20 _m12 _m10; do for c in C3 C2; do echo -n "variant '$v' $c: "; CRN_LIB=$PWD/cloud-renderer_b200/libcloud_renderer_b200$v.so python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done; done
