#!/usr/bin/env python
"""Times the fully observed path on scaled twins of configs 4 and 5 (and checks 2 iterations against the oracle)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import lowrankmodels_b200 as lrm
from lowrankmodels_b200 import synth

def problem(name, scale):
    if name == "C5":
        c = synth.config5(scale=scale)
        return lrm.GLRM(c["A"], lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg(), c["k"], X=c["X0"], Y=c["Y0"])
    if name == "C4":
        c = synth.config4(scale=scale)
        losses = [lrm.QuadLoss()] * c["n_quad"] + [lrm.HingeLoss()] * c["n_hinge"] + [lrm.MultinomialLoss(c["levels"])] * c["n_multi"]
        return lrm.GLRM(c["A"], losses, lrm.QuadReg(0.1), lrm.QuadReg(0.1), c["k"], X=c["X0"], Y=c["Y0"])
    c = synth.config1()
    return lrm.GLRM(c["A"], lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), c["k"], X=c["X0"], Y=c["Y0"])

for name, scale, oracle_iters in [(a.split("/")[0], int(a.split("/")[1]), int(a.split("/")[2])) for a in sys.argv[1:]]:
    t = time.time()
    g = problem(name, scale)
    ep = lrm.encode_problem(g, validate=False)
    tgen = time.time() - t
    for dense in ("1", "fma", "0"):                   # tensor-core kernels, FP64-FMA kernels (glrm_dense.cuh), gather kernels
        os.environ["GLRMB200_DENSE"] = "0" if dense == "0" else "1"
        os.environ["GLRMB200_DENSE_MMA"] = "0" if dense == "fma" else "1"
        if dense in os.environ.get("DENSE_CHECK_SKIP", "").split(","):
            continue
        if dense == "0" and ep.nnz > 400_000_000:
            continue
        t = time.time()
        try:
            eng = lrm.Engine(ep, validate=False)
        except Exception as e:
            print(json.dumps({"config": f"{name}/{scale}", "dense": dense, "error": str(e)}), flush=True)
            continue
        tcreate = time.time() - t
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        eng.upload(X, Y)
        eng.fit_resident(lrm.ProxGradParams(max_iter=3, abs_tol=0, rel_tol=0))
        obj, _ = eng.fit_resident(lrm.ProxGradParams(max_iter=5, abs_tol=0, rel_tol=0))
        p = eng.last_profile
        out = {"config": f"{name}/{scale}", "dense": dense, "shape": list(g.shape), "k": g.k, "nnz": ep.nnz, "gen_s": round(tgen, 2),
               "create_s": round(tcreate, 2), "ms_per_step": p["loop_ms"] / 5, "x_ms": p["update_x_ms"] / 5, "y_ms": p["update_y_ms"] / 5,
               "x_trials_per_row": p["x_trials"] / 5 / g.shape[0], "y_trials_per_col": p["y_trials"] / 5 / g.shape[1],
               "launches": p["x_launches"] + p["y_launches"], "obj": [float(obj[0]), float(obj[-1])]}
        if oracle_iters and dense == "1":
            import oracle_py
            pk = lrm.ProxGradParams(max_iter=oracle_iters, abs_tol=0, rel_tol=0)
            Xe, Ye = g.X.copy(order="F"), g.Y.copy(order="F")
            got, _ = eng.fit(pk, Xe, Ye)
            Xo, Yo = g.X.copy(order="F"), g.Y.copy(order="F")
            t = time.time()
            want = oracle_py.fit(ep, lrm.encode_params(pk), Xo, Yo, mode=1)["objective"]
            fin = np.isfinite(want)
            out["oracle_s"] = round(time.time() - t, 1)
            out["oracle_max_rel_err"] = float(np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin])))
            out["inf_match"] = bool((np.isfinite(got) == fin).all())
        eng.close()
        print(json.dumps(out), flush=True)
