#!/bin/bash
# round 2, call 3 (2 GPUs): multi-GPU parity (NCCL + fused exchange with the peer flag barrier), 2-GPU bench, 1-GPU tier check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -8 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json; grep -E "e2e breakdown" gpurun_out/bench_2gpu.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/tune.py C2 1 "" "GLRMB200_CLUSTER16=32768" "GLRMB200_CLUSTER16=1000000" > gpurun_out/tune_c2.jsonl 2> gpurun_out/tune.err; echo "tune rc=$?"; cat gpurun_out/tune_c2.jsonl
