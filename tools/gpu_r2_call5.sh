#!/bin/bash
# round 2, call 5: dense path (tests + timings), pipeline-depth / tier variants incl. an emulated 1/8 shard
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/dense_check.py C1/1/5 C5/64/3 C4/16/3 C5/8/0 C4/4/0 > gpurun_out/dense_check.jsonl 2> gpurun_out/dense_check.err; echo "dense rc=$?"; cat gpurun_out/dense_check.jsonl; tail -5 gpurun_out/dense_check.err
V=lowrankmodels.jl_b200/csrc/variants
timeout 900 python tools/tune.py C2 1 "" "SHARD=0/8" "SHARD=3/8" "LIB=$V/libglrm_b200_hd4c1.so" "LIB=$V/libglrm_b200_hd4c1.so SHARD=0/8" \
   "LIB=$V/libglrm_b200_hd4c1t8.so" "LIB=$V/libglrm_b200_hd4c1t8.so SHARD=0/8" "LIB=$V/libglrm_b200_hd4c2.so" "LIB=$V/libglrm_b200_hd4c2.so SHARD=0/8" \
   "LIB=$V/libglrm_b200_ld4c3.so" "LIB=$V/libglrm_b200_ld4c3.so SHARD=0/8" \
   "SHARD=0/8 GLRMB200_HEAVY=384 GLRMB200_CLUSTER=3072 GLRMB200_CLUSTER16=12288" "GLRMB200_HEAVY=384 GLRMB200_CLUSTER=3072 GLRMB200_CLUSTER16=12288" \
   "LIB=$V/libglrm_b200_hd4c1.so SHARD=0/8 GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096" "LIB=$V/libglrm_b200_hd4c1.so GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096" \
   > gpurun_out/tune_variants.jsonl 2> gpurun_out/tune.err; echo "tune rc=$?"; cat gpurun_out/tune_variants.jsonl; tail -3 gpurun_out/tune.err
