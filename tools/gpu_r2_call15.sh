#!/bin/bash
# round 2, call 15 (2 GPUs): row-sharded fully observed fits (mode B) bit-identical to one GPU; C5/8 on 1 and 2 GPUs
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=20
timeout 300 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest single rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "2" --timeout=400 --timeout-method=thread > gpurun_out/pytest_multi2.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi2.log | cut -c1-400; cat gpurun_out/mgpu_worker_2.log | cut -c1-300 | head -30
timeout 300 python bench.py --config C5 --scale 8 --steps 5 --warmup 3 --extra '' --no-cpu > gpurun_out/bench_c5s8_1gpu.json 2> gpurun_out/bench_c5s8_1gpu.err; echo "bench 1gpu rc=$?"; cut -c1-900 gpurun_out/bench_c5s8_1gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config C5 --scale 8 --steps 5 --warmup 3 --extra '' --no-cpu > gpurun_out/bench_c5s8_2gpu.json 2> gpurun_out/bench_c5s8_2gpu.err; echo "bench 2gpu rc=$?"; cut -c1-1500 gpurun_out/bench_c5s8_2gpu.json; tail -4 gpurun_out/bench_c5s8_2gpu.err | cut -c1-300
