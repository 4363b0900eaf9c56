#!/bin/bash
# multi-GPU run: profiles/run_mg.sh N  -> 2-rank NCCL parity test (N >= 2), C3 and C4 bench lines at N GPUs
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = 2 ]; then python -m pytest tests/test_configs_gpu.py -m gpu -q -k two_rank -s > gpurun_out/pytest_nccl_n2.log 2>&1; tail -6 gpurun_out/pytest_nccl_n2.log; fi
for cfg in C3 C4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 3 --config $cfg \
      --no-cpu-baseline > gpurun_out/bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err
  tail -2 gpurun_out/bench_${cfg}_n$N.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${cfg}_n$N.json'))
    print('$cfg N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['scaling'], d['config']['sharding'][:60])
    for k,v in d['extra'].items(): print('   ', k, {kk: vv for kk, vv in v.items() if kk in ('frames_per_s','e2e_frames_per_s','ms_per_frame','mode','exchange','all_gather_ms')})
except Exception as e: print('no line', e)
PY
done
