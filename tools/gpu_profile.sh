#!/bin/bash
# bench + ncu evidence for profiles/: launch list of the bench command, one --set full capture of the sweep kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 10 -c 4 -o gpurun_out/prof_final \
   python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json
