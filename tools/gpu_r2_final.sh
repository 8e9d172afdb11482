#!/bin/bash
# round 2, final single-GPU visit: smoke, full GPU suite, the bench line (driver's flags), ncu launch lists and --set full captures
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log | cut -c1-200
timeout 400 python -m pytest tests -x -q -m gpu --timeout=120 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_final.json; tail -6 gpurun_out/bench_final.err | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --extra '' > gpurun_out/ncu_list_c2.log 2>&1; echo "ncu list c2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5s8.csv \
   python bench.py --config C5 --scale 8 --steps 2 --warmup 3 --no-cpu --extra '' > gpurun_out/ncu_list_c5.log 2>&1; echo "ncu list c5 rc=$?"
DENSE_CHECK_SKIP=0,fma timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_mma -s 10 -c 5 -o gpurun_out/prof_final_mma_c5 \
   python tools/dense_check.py C5/64/0 > gpurun_out/ncu_final_c5.log 2>&1; echo "ncu c5 rc=$?"
DENSE_CHECK_SKIP=0,fma timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_mma -s 10 -c 5 -o gpurun_out/prof_final_mma_c4 \
   python tools/dense_check.py C4/4/0 > gpurun_out/ncu_final_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out | tail -8
