#!/bin/bash
# round 2, call 10: after the barrier re-arm fix — hang repro cases, full GPU test-suite, dense timings, the full bench line
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=15
L=gpurun_out/dense_debug3.log
: > $L
run() { echo "=== $*" >> $L; timeout 40 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/dense_debug.py 30000 100 20 scalar_only
run python tools/dense_debug.py 30000 140 100 quad
run python tools/dense_debug.py 30000 40 20
cut -c1-250 $L
timeout 800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 240 python tools/dense_check.py C5/64/3 C4/16/3 C5/8/0 C4/4/0 > gpurun_out/dense_check4.jsonl 2> gpurun_out/dense_check4.err; echo "dense rc=$?"; cut -c1-420 gpurun_out/dense_check4.jsonl; tail -3 gpurun_out/dense_check4.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cut -c1-6000 gpurun_out/bench_full.json; tail -12 gpurun_out/bench_full.err | cut -c1-700
