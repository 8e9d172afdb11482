#!/bin/bash
# round 2, call 23: device impute / error_metric against the CPU restatement; full GPU suite on the final build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_evaluate.py -x -q -m gpu --timeout=90 --timeout-method=thread > gpurun_out/pytest_eval.log 2>&1; echo "pytest eval rc=$?"; tail -25 gpurun_out/pytest_eval.log | cut -c1-300
timeout 400 python -m pytest tests -q -m gpu --timeout=120 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-200
