#!/bin/bash
mkdir -p gpurun_out
python profiles/trace_time.py 2>&1 | tail -1
for v in q2 mb12 mb14 mb18 mb20 mb24 q2mb20; do
  CRN_LIB=$PWD/cloud-renderer_b200/libcloud_renderer_b200_$v.so python profiles/trace_time.py 2>&1 | tail -1
done
