#!/bin/bash
# N-GPU bench line of the default config (extras: C5, C4 slab exchange vs replicated) -> gpurun_out/bench_r02_C3_n$N.json
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 20 --warmup 3 \
    --no-cpu-baseline --save gpurun_out/bench_r02_C3_n$N.json > /dev/null 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_C3_n$N.json'))
print('C3 N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d['extra'].items(): print('   ', k, {kk: vv for kk, vv in v.items() if kk in ('frames_per_s','e2e_frames_per_s','ms_per_frame','mode','exchange','trace_kernel_ms','voxelize_mip_ms','cone_accel_ms')})
PY
