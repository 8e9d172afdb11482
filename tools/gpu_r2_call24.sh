#!/bin/bash
# round 2, call 24: the full GPU suite on the final tree (new: NT=4 tile, tensor-core vs FMA kernels, impute consistency), smoke
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout=120 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-200
