#!/bin/bash
# round 2, first GPU visit: measured roofs (microbench), tier-threshold sweep, C2/C3 bench lines, steady-state ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 300 ./tools/bin/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err; echo "microbench rc=$?"; cat gpurun_out/microbench.json
timeout 600 python tools/tune.py C2 1 > gpurun_out/tune_c2.jsonl 2> gpurun_out/tune.err; echo "tune rc=$?"; cat gpurun_out/tune_c2.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err; echo "bench c3 rc=$?"; cat gpurun_out/bench_c3.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 60 -c 7 -o gpurun_out/prof_steady \
   python bench.py --steps 2 --warmup 10 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
