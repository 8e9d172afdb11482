#!/bin/bash
# round 2, call 6: TMA-fed dense kernels — parity tests, timings, sanitizer on small dense / cluster cases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense or config1" > gpurun_out/pytest_dense.log 2>&1; echo "pytest dense rc=$?"; tail -15 gpurun_out/pytest_dense.log
timeout 900 python tools/dense_check.py C1/1/5 C5/64/3 C4/16/3 C5/8/0 C4/4/0 > gpurun_out/dense_check2.jsonl 2> gpurun_out/dense_check2.err; echo "dense rc=$?"; cat gpurun_out/dense_check2.jsonl; tail -5 gpurun_out/dense_check2.err
GLRMB200_DENSE_NBUF=1 timeout 600 python tools/dense_check.py C5/64/0 C4/16/0 > gpurun_out/dense_check2_nbuf1.jsonl 2>&1; cat gpurun_out/dense_check2_nbuf1.jsonl
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense_path_every_rank_tile and (20 or 100) or dense_path_heterogeneous" > gpurun_out/sanitizer_memcheck_dense.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/sanitizer_memcheck_dense.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense_path_every_rank_tile and (20 or 100)" > gpurun_out/sanitizer_racecheck_dense.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/sanitizer_racecheck_dense.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_gpu.log
