#!/bin/bash
# round 2, call 25 (2 GPUs): multi-GPU bit-identity on the final tree (new row-block partition of mode B, C1 twin on the gather kernels)
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=20
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "2" --timeout=350 --timeout-method=thread > gpurun_out/pytest_multi2.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi2.log | cut -c1-300; head -12 gpurun_out/mgpu_worker_2.log | cut -c1-200
