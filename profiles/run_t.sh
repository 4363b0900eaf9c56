#!/bin/bash
mkdir -p gpurun_out
python profiles/microbench.py > gpurun_out/microbench_r02b.json 2> gpurun_out/microbench.err; cat gpurun_out/microbench_r02b.json; tail -3 gpurun_out/microbench.err
