#!/bin/bash
# N-GPU run: the real-NCCL sharding test and the scaling bench (GPUS from the environment, default 2)
N=${GPUS:-2}
mkdir -p gpurun_out
python -m pytest tests/test_configs_gpu.py -m gpu -x -q -s -k two_rank > gpurun_out/pytest_nccl.log 2>&1; tail -6 gpurun_out/pytest_nccl.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -4 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['extra'].items(): print(k, {kk: vv for kk, vv in v.items() if kk != 'workload'})
PY
