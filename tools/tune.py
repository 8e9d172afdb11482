#!/usr/bin/env python
"""Tuning sweep on the GPU box: one problem, several engine configurations (env hooks read at create)."""
import itertools, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lowrankmodels_b200 as lrm
from bench import build_problem

config = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1
variants = sys.argv[3:] or ["", "GLRMB200_HEAVY=512", "GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096",
                            "GLRMB200_HEAVY=512 GLRMB200_CLUSTER=4096 GLRMB200_CLUSTER16=16384",
                            "GLRMB200_HEAVY=256 GLRMB200_CLUSTER=2048 GLRMB200_CLUSTER16=16384",
                            "GLRMB200_HEAVY=1024 GLRMB200_CLUSTER=4096 GLRMB200_CLUSTER16=16384"]
g, cfg = build_problem(config, scale)
ep = lrm.encode_problem(g, validate=False)
nnz = ep.nnz
STEPS, WARM = int(os.environ.get("TUNE_STEPS", "10")), int(os.environ.get("TUNE_WARM", "3"))
pw = lrm.ProxGradParams(max_iter=WARM, abs_tol=0, rel_tol=0)
pk = lrm.ProxGradParams(max_iter=STEPS, abs_tol=0, rel_tol=0)
ref = None
for v in variants:
    keys = []
    rank, nranks = 0, 1
    lrm._abi._lib = lrm._abi.load(lrm._abi.LIB_PATH)          # every variant starts from the default build
    for kv in v.split():
        k_, val = kv.split("=")
        if k_ == "LIB":                      # a differently compiled engine build, loaded side by side
            lrm._abi._lib = lrm._abi.load(os.path.join(ROOT, val))
            continue
        if k_ == "SHARD":                    # "r/N": time rank r's shard of an N-rank fit on this GPU (no exchange: timing only)
            rank, nranks = (int(x) for x in val.split("/"))
            os.environ["GLRMB200_NO_EXCHANGE"] = "1"
            keys.append("GLRMB200_NO_EXCHANGE")
            continue
        os.environ[k_] = val
        keys.append(k_)
    eng = lrm.Engine(ep, validate=False, rank=rank, nranks=nranks)
    eng.upload(g.X, g.Y)
    eng.fit_resident(pw)
    obj, _ = eng.fit_resident(pk)
    p = eng.last_profile
    eng.close()
    for k_ in keys:
        del os.environ[k_]
    if ref is None:
        ref = obj
    dev = float(np.max(np.abs(obj - ref) / np.abs(ref)))
    print(json.dumps({"variant": v or "default", "ms_per_step": p["loop_ms"] / STEPS, "x_ms": p["update_x_ms"] / STEPS,
                      "y_ms": p["update_y_ms"] / STEPS, "Gentries_s": nnz / (p["loop_ms"] / STEPS * 1e-3) / 1e9,
                      "x_trials": p["x_trials"], "y_trials": p["y_trials"], "obj_last": float(obj[-1]),
                      "max_rel_dev_vs_first": dev}), flush=True)
