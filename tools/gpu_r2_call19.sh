#!/bin/bash
# round 2, call 19: level-count-specialised multinomial, compute-sanitizer on the tensor-core kernels and on the cluster tier
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=60
timeout 300 python -m pytest tests -x -q -m gpu --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
GLRMB200_PHASE_TIMERS=1 DENSE_CHECK_SKIP=0,fma timeout 300 python tools/dense_check.py C4/4/0 C4/16/3 > gpurun_out/dense_check9.jsonl 2> gpurun_out/dense_check9.err; echo "dense rc=$?"; cut -c1-420 gpurun_out/dense_check9.jsonl; cat gpurun_out/dense_check9.err | cut -c1-400
L=gpurun_out/sanitizer_mma.log
: > $L
run() { echo "=== $*" >> $L; timeout 240 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run compute-sanitizer --tool memcheck --error-exitcode 9 python tools/dense_debug.py 1000 40 20
run compute-sanitizer --tool memcheck --error-exitcode 9 python tools/dense_debug.py 700 140 100 quad
run compute-sanitizer --tool synccheck --error-exitcode 9 python tools/dense_debug.py 300 40 20
run compute-sanitizer --tool racecheck --error-exitcode 9 python tools/dense_debug.py 300 40 20
run compute-sanitizer --tool racecheck --error-exitcode 9 python tools/dense_debug.py 300 140 100 quad
grep -v "^\[dense\]" $L | cut -c1-250 | tail -40
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster" > gpurun_out/sanitizer_racecheck_cluster.log 2>&1; echo "racecheck cluster rc=$?"; tail -5 gpurun_out/sanitizer_racecheck_cluster.log | cut -c1-300
