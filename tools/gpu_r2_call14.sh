#!/bin/bash
# round 2, call 14: row-itself tie shortcut, pipelined operand fetch of the per-feature losses
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=15
timeout 600 python -m pytest tests -x -q -m gpu --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
DENSE_CHECK_SKIP=0,fma timeout 300 python tools/dense_check.py C5/64/3 C4/16/3 C5/8/0 C4/4/0 > gpurun_out/dense_check7.jsonl 2> gpurun_out/dense_check7.err; echo "dense rc=$?"; cut -c1-480 gpurun_out/dense_check7.jsonl; tail -3 gpurun_out/dense_check7.err
