#!/bin/bash
# round 2, call 22: smoke again; the heavy-unit tiers on one GPU (A/B of the cluster thresholds, 20 timed steps after 5)
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log | cut -c1-200
TUNE_STEPS=20 TUNE_WARM=5 timeout 400 python tools/tune.py C2 1 "" "GLRMB200_CLUSTER16=1000000000" "GLRMB200_CLUSTER=1000000000 GLRMB200_CLUSTER16=1000000000" "GLRMB200_CLUSTER=16384 GLRMB200_CLUSTER16=1000000000" "GLRMB200_CLUSTER16=32768" "" > gpurun_out/tune_tiers2.jsonl 2> gpurun_out/tune_tiers2.err; echo "tune rc=$?"; cut -c1-260 gpurun_out/tune_tiers2.jsonl
