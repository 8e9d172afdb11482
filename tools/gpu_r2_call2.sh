#!/bin/bash
# round 2, call 2: restructured engine (async loop, device-side stop rule, pooled allocations, staged factor transfers)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench.err
timeout 600 python tools/tune.py C2 1 "" "GLRMB200_CLUSTER16=32768" "GLRMB200_CLUSTER=4096" "GLRMB200_HEAVY=512" > gpurun_out/tune_c2.jsonl 2> gpurun_out/tune.err; echo "tune rc=$?"; cat gpurun_out/tune_c2.jsonl
