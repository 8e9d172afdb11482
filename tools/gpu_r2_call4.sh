#!/bin/bash
# round 2, call 4 (8 GPUs): 8-GPU bit-identity, strong-scaling bench at N=8 and N=4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "8" > gpurun_out/pytest_multi8.log 2>&1; echo "pytest multi8 rc=$?"; tail -5 gpurun_out/pytest_multi8.log
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
c=d['config']
print($N,'ms/step',d['ms_per_step'],'x',c['update_x_ms'],'y',c['update_y_ms'],'comm',c['comm_ms'],'e2e',d['e2e']['value'],d['e2e']['seconds_per_call'])
print(c['per_rank_ms_x_y_comm_loop'])
PY
grep -E "e2e breakdown" gpurun_out/bench_${N}gpu.err | head -3
done
