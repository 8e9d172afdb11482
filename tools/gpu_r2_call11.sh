#!/bin/bash
# round 2, call 11: FP64 pipe rates (DFMA at low occupancy, DMMA m8n8k4), full GPU test-suite on the row-group reduction tree,
# ncu --set full of the fully observed kernels on the C5/64 twin
mkdir -p gpurun_out
timeout 120 ./tools/bin/microbench fp64 > gpurun_out/microbench_fp64.json 2> gpurun_out/microbench_fp64.err; echo "microbench rc=$?"; cat gpurun_out/microbench_fp64.json
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_ -s 30 -c 14 -o gpurun_out/prof_dense_r2 \
   python tools/dense_check.py C5/64/0 > gpurun_out/ncu_dense.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dense.log | cut -c1-300
ls -la gpurun_out | tail -5
