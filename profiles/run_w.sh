#!/bin/bash
# A/B of experiment builds (build.py build_variant): profiles/run_w.sh "<tag> <tag> ..." [configs]
mkdir -p gpurun_out
for v in "" $1; do for c in ${2:-C3 C2}; do s=${v:+_$v}; echo -n "variant '$v' $c: "; CRN_LIB=$PWD/cloud-renderer_b200/libcloud_renderer_b200$s.so python profiles/trace_time.py --config $c --frames 6 2>&1 | tail -1; done; done
