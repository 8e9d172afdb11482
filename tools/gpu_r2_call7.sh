#!/bin/bash
# round 2, call 7: dense kernels v3 (cp.async staging, LDS), tier streams A/B + thresholds, full GPU test-suite
mkdir -p gpurun_out
export GLRMB200_WAIT_LIMIT_S=20
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense or config1" > gpurun_out/pytest_dense.log 2>&1; echo "pytest dense rc=$?"; tail -12 gpurun_out/pytest_dense.log
timeout 300 python tools/dense_check.py C1/1/5 C5/64/3 C4/16/3 C5/8/0 C4/4/0 > gpurun_out/dense_check3.jsonl 2> gpurun_out/dense_check3.err; echo "dense rc=$?"; cut -c1-420 gpurun_out/dense_check3.jsonl; tail -3 gpurun_out/dense_check3.err
timeout 500 python tools/tune.py C2 1 "" "GLRMB200_TIER_STREAMS=1" "GLRMB200_NO_PRIORITY=1" \
   "GLRMB200_CLUSTER=4096" "GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096" "GLRMB200_HEAVY=512" "GLRMB200_HEAVY=768 GLRMB200_CLUSTER=4096 GLRMB200_CLUSTER16=12288" \
   "SHARD=0/8" "SHARD=3/8" "SHARD=0/8 GLRMB200_TIER_STREAMS=1" "SHARD=0/8 GLRMB200_NO_PRIORITY=1" \
   "SHARD=0/8 GLRMB200_CLUSTER=4096" "SHARD=0/8 GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096" "SHARD=0/8 GLRMB200_HEAVY=512" \
   "SHARD=0/8 GLRMB200_HEAVY=768 GLRMB200_CLUSTER=4096 GLRMB200_CLUSTER16=12288" "SHARD=3/8 GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096" \
   "SHARD=0/4" "SHARD=0/4 GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096" "SHARD=0/2" \
   > gpurun_out/tune_streams.jsonl 2> gpurun_out/tune.err; echo "tune rc=$?"; cut -c1-260 gpurun_out/tune_streams.jsonl; tail -3 gpurun_out/tune.err
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent_tiles and 20-40000 or fixed_latent_features_every_tile and 50-17" > gpurun_out/sanitizer_memcheck_dense2.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck_dense2.log
