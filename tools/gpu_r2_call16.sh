#!/bin/bash
# round 2, call 16: phase clocks of the tensor-core X sweep (C5/8, C4/4), regression of the new row-block partition
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=15
timeout 300 python -m pytest tests -x -q -m gpu --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
GLRMB200_PHASE_TIMERS=1 DENSE_CHECK_SKIP=0,fma timeout 300 python tools/dense_check.py C5/8/0 C4/4/0 C5/64/3 > gpurun_out/dense_check8.jsonl 2> gpurun_out/dense_check8.err; echo "dense rc=$?"; cut -c1-420 gpurun_out/dense_check8.jsonl; cat gpurun_out/dense_check8.err | cut -c1-400
