#!/bin/bash
# round 2, call 20: what bounds one rank's shard of an 8-GPU C2 fit?  (SHARD=r/8 hook: the shard's sweeps on one GPU, no exchange)
mkdir -p gpurun_out
timeout 400 python tools/tune.py C2 1 "" "SHARD=0/8" "SHARD=1/8" "SHARD=4/8" "SHARD=7/8" "SHARD=0/4" "SHARD=0/2" > gpurun_out/tune_shards.jsonl 2> gpurun_out/tune_shards.err; echo "tune rc=$?"; cut -c1-330 gpurun_out/tune_shards.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_shard0of8.csv python tools/tune.py C2 1 "SHARD=0/8" > gpurun_out/ncu_shard.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_shard0of8.csv')))
h=None
for i,r in enumerate(rows):
    if r and r[0]=='ID': h=r; start=i+1; break
ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size'); ib=h.index('Block Size')
for r in rows[start:start+60]:
    print(r[ik][:60], r[ig], r[ib], r[iv])
PY
