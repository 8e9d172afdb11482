#!/bin/bash
# round 2, call 13: repeated-trial shortcut in the tensor-core X sweep, full GPU suite, timings, ncu of the tensor-core kernels
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=15
timeout 600 python -m pytest tests -x -q -m gpu --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
DENSE_CHECK_SKIP=0,fma timeout 300 python tools/dense_check.py C5/64/3 C4/16/3 C5/8/0 C4/4/0 C1/1/3 > gpurun_out/dense_check6.jsonl 2> gpurun_out/dense_check6.err; echo "dense rc=$?"; cut -c1-480 gpurun_out/dense_check6.jsonl; tail -3 gpurun_out/dense_check6.err
DENSE_CHECK_SKIP=0,fma timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_mma -s 12 -c 6 -o gpurun_out/prof_mma_c5 \
   python tools/dense_check.py C5/64/0 > gpurun_out/ncu_mma_c5.log 2>&1; echo "ncu c5 rc=$?"
DENSE_CHECK_SKIP=0,fma timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_mma -s 12 -c 6 -o gpurun_out/prof_mma_c4 \
   python tools/dense_check.py C4/16/0 > gpurun_out/ncu_mma_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out | tail -4
