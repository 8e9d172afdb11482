#!/bin/bash
# round 2, call 12: first run of the tensor-core (DMMA) kernels of the fully observed path
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=15
L=gpurun_out/mma_debug.log
: > $L
run() { echo "=== $*" >> $L; timeout 60 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/dense_debug.py 300 40 5 quad
run python tools/dense_debug.py 3000 140 100 quad
run python tools/dense_debug.py 3000 70 20 scalar_only
run python tools/dense_debug.py 3000 40 20
run python tools/dense_debug.py 30000 140 100 quad
cut -c1-300 $L
timeout 600 python -m pytest tests -x -q -m gpu --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
DENSE_CHECK_SKIP=0 timeout 300 python tools/dense_check.py C5/64/3 C4/16/3 C5/8/0 C4/4/0 > gpurun_out/dense_check5.jsonl 2> gpurun_out/dense_check5.err; echo "dense rc=$?"; cut -c1-480 gpurun_out/dense_check5.jsonl; tail -3 gpurun_out/dense_check5.err
