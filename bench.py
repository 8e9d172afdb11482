#!/usr/bin/env python
"""bench.py — observed-entries / sec / iteration of the GLRM prox-grad hot path on B200.

Headline workload (BASELINE.json configs[1], "C2"): MovieLens-20M-shaped sparse A (138 493 x 26 744,
20 000 263 observations), QuadLoss + QuadReg(0.1), k = 50, Float64, synthetic (hash-generated) data.
A *step* is one outer iteration of fit!(glrm, ProxGradParams) (proxgrad.jl:107-217): one X sweep,
one Y sweep, the objective record.

    python bench.py --gpus N --steps K --warmup W            our engine (N>1: launched by torchrun)
    python bench.py --impl reference ...                     the reference's algorithm on the host CPUs
                                                             (oracle port, faithful dense-XY form, OpenBLAS dgemm)

value  : whole-job entries/s/iter with the problem and the factors resident in HBM
e2e    : the same metric through the reference-facing call (glrmb200_create + glrmb200_fit with host
         buffers + read-back + destroy): every H2D / D2H copy is inside the timed region
extra_configs : the other BASELINE configurations at full size (C3; C4 and C5 through the fully observed
         path), each with its own timing, roofline fractions and a trajectory check against the CPU oracle
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "observed-entries/sec/iter (X+Y sweep) at k=50"
UNIT = "entries/s/iter"
PARITY_TOL = 1e-4            # north_star: objective trajectories within 1e-4 relative of the reference
FP64_PEAK_TFLOPS = 37.06     # tools/microbench.cu dmma_peak (FP64 mma.sync m8n8k4; DFMA: 36.58) on this pool's B200 (profiles/r2_microbench_fp64.json)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback 6650 GB/s (B200_PROFILING.md)", "fp64_tflops": FP64_PEAK_TFLOPS,
           "l2_gather_gbs": None}
    try:
        out["hbm_gbs"] = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        out["hbm_source"] = "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        pass
    try:
        mb = json.load(open(os.path.join(ROOT, "profiles", "r2_microbench.json")))
        out["fp64_tflops"] = float(mb["dfma_peak"]["TFLOPs"])
        try:
            mb2 = json.load(open(os.path.join(ROOT, "profiles", "r2_microbench_fp64.json")))
            out["fp64_tflops"] = max(out["fp64_tflops"], float(mb2["dmma_peak"]["TFLOPs"]))
        except Exception:
            pass
        out["l2_gather_gbs"] = {"y": float(mb["ldg_d4_y13MB"]["GBps"]), "x": float(mb["ldg_d4_x71MB"]["GBps"])}
    except Exception:
        pass
    return out


# ---- problems ------------------------------------------------------------------------------------------
def config_static(config):
    """The config dictionary both arms print (same workload, same words)."""
    return {
        "C2": {"workload": "C2: MovieLens-20M-shaped sparse 138493x26744, 20000263 obs, QuadLoss+QuadReg(0.1), k=50",
               "l2": "inputs larger than L2 (CSR+CSC index/value streams 480 MB + factors 66 MB vs 126 MB L2); no explicit flush"},
        "C3": {"workload": "C3: MovieLens-20M-shaped sparse 138493x26744, 20000263 obs, LogisticLoss+NonNegConstraint, k=50",
               "l2": "inputs larger than L2 (CSR+CSC index/value streams 480 MB + factors 66 MB vs 126 MB L2); no explicit flush"},
        "C4": {"workload": "C4: heterogeneous columns 1000000x1000 fully observed, 500 QuadLoss + 300 HingeLoss + 200 MultinomialLoss(5) "
                           "(d=1800), QuadReg(0.1), k=20",
               "l2": "inputs larger than L2 (A 8 GB, X 160 MB vs 126 MB L2); no explicit flush"},
        "C5": {"workload": "C5: k-means path 10000000x128 fully observed, QuadLoss + UnitOneSparseConstraint (rx) + ZeroReg (ry), k=100",
               "l2": "inputs larger than L2 (A 10.2 GB, X 10.2 GB vs 126 MB L2); no explicit flush"},
    }[config]


def build_problem(config, scale=1, rows=None):
    """-> (glrm, cfg) for one of the BASELINE configurations (scale > 1 / rows: CI twins)."""
    import lowrankmodels_b200 as lrm
    from lowrankmodels_b200 import synth
    if config in ("C2", "C3"):
        import scipy.sparse as sp
        if config == "C2":
            cfg, loss, reg = synth.config2(scale=scale), lrm.QuadLoss(), lrm.QuadReg(0.1)
        else:
            cfg, loss, reg = synth.config3(scale=scale), lrm.LogisticLoss(), lrm.NonNegConstraint()
        A = sp.csc_matrix((cfg["vals"], (cfg["rows"], cfg["cols"])), shape=(cfg["m"], cfg["n"]))
        return lrm.GLRM(A, loss, reg, reg, cfg["k"], X=cfg["X0"], Y=cfg["Y0"], checknan=False), cfg
    if config == "C4":
        c = synth.config4(scale=scale, rows=rows)
        losses = [lrm.QuadLoss()] * c["n_quad"] + [lrm.HingeLoss()] * c["n_hinge"] + [lrm.MultinomialLoss(c["levels"])] * c["n_multi"]
        return lrm.GLRM(c["A"], losses, lrm.QuadReg(0.1), lrm.QuadReg(0.1), c["k"], X=c["X0"], Y=c["Y0"], checknan=False), c
    if config == "C5":
        c = synth.config5(scale=scale, rows=rows)
        return lrm.GLRM(c["A"], lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg(), c["k"], X=c["X0"], Y=c["Y0"],
                        checknan=False), c
    raise SystemExit(f"unknown config {config}")


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        # samples inside the timed region; the region is tens of ms, so fall back to the closest samples around it
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or \
               [r for t, r in self.rows if t0 - 0.30 <= t <= t1 + 0.30 and len(r) >= 9] or \
               [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


# ---- CPU arm ---------------------------------------------------------------------------------------
def physical_cores():
    """Physical cores this process may run on (hyper-thread siblings counted once)."""
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return os.cpu_count() or 1
    cores, cur = set(), {}
    try:
        for line in open("/proc/cpuinfo"):
            if ":" not in line:
                if "processor" in cur and int(cur["processor"]) in allowed:
                    cores.add((cur.get("physical id", "0"), cur.get("core id", cur["processor"])))
                cur = {}
                continue
            key, v = line.split(":", 1)
            cur[key.strip()] = v.strip()
        if "processor" in cur and int(cur["processor"]) in allowed:
            cores.add((cur.get("physical id", "0"), cur.get("core id", cur["processor"])))
    except OSError:
        pass
    return max(1, len(cores)) if cores else len(allowed)


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    oracle_py.lib()
    return oracle_py


def cpu_faithful_sample(g, ep, steps, warmup, budget_s):
    """The reference's algorithm as written (proxgrad_multithread.jl: materialised XY through BLAS dgemm, dense x'Y / X'y in
    the line search) on the FULL-SIZE problem, one bounded sample per step: the X sweep over the first m/S rows and the Y
    sweep over the first n/S columns (all other units frozen), S chosen so that the run fits `budget_s`.  A full iteration
    costs S sample steps (rows and columns are i.i.d. in the generator, and the faithful form's cost per unit is the dense
    product, independent of the unit's degree).  -> dict(value, seconds_per_iter, S, threads, dgemm)."""
    import lowrankmodels_b200 as lrm
    orc = _oracle()
    threads = physical_cores()
    dgemm = orc.use_openblas_dgemm(threads)
    m, n = g.shape
    nnz = ep.nnz

    def run(S, iters):
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        p = lrm.ProxGradParams(max_iter=iters, abs_tol=0, rel_tol=0)
        res = orc.fit_units(ep, lrm.encode_params(p), X, Y, (0, max(1, m // S)), (0, max(1, n // S)), mode=0, nthreads=threads)
        return res["seconds"][1:]

    S = 64
    t1 = float(run(S, 1)[0])                                        # calibration step (also warms the BLAS threads)
    per_step_target = max(0.5, budget_s / max(1, steps + warmup))
    while S > 1 and t1 * 2 <= per_step_target:
        S //= 2
        t1 *= 2
    sec = run(S, warmup + steps)[warmup:]
    per_sample = float(np.mean(sec))
    per_iter = per_sample * S
    return {"value": nnz / per_iter, "seconds_per_iter": per_iter, "seconds_per_sample_step": per_sample, "S": S,
            "threads": threads, "dgemm": "OpenBLAS cblas_dgemm (NumPy's bundled libscipy_openblas64_)" if dgemm else
            "dot-product loop (no BLAS found)",
            "sample": f"full-size problem, each step sweeps rows [0, {max(1, m // S)}) and columns [0, {max(1, n // S)}) "
                      f"(1/{S} of the units; a full iteration = {S} such steps), faithful dense-XY form"}


def cpu_sparse_evaluated(g, ep, iters=10):
    """B2, best-effort CPU: the same arithmetic touching observed entries only (oracle mode 1), full size."""
    import lowrankmodels_b200 as lrm
    orc = _oracle()
    threads = physical_cores()
    X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
    p = lrm.ProxGradParams(max_iter=iters + 1, abs_tol=0, rel_tol=0)
    res = orc.fit(ep, lrm.encode_params(p), X, Y, mode=1, nthreads=threads)
    per_iter = float(np.mean(res["seconds"][2:]))
    return ep.nnz / per_iter, per_iter, res["objective"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import lowrankmodels_b200 as lrm
    g, cfg = build_problem(args.config, args.scale)
    ep = lrm.encode_problem(g, validate=False)
    r = cpu_faithful_sample(g, ep, args.steps, args.warmup, budget_s=150.0)
    val = r["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_iter"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_static(args.config) if args.scale == 1 else
        {"workload": config_static(args.config)["workload"] + f" - /{args.scale} twin", "l2": "twin: not a bench configuration"},
        "detail": {"algorithm": "proxgrad_multithread.jl as written (dense XY, dense line-search products), C/OpenMP port "
                                "(Julia is not installed on this box); XY through " + r["dgemm"],
                   "seconds_per_sample_step": r["seconds_per_sample_step"], "units_fraction": f"1/{r['S']}",
                   "ms_per_step_is": "the extrapolated full iteration (S sample steps)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -----------------------------------------------------------------------------------------
def pin_encoded(ep):
    """Re-home every array of the encoded problem in pinned host memory, so the e2e leg's H2D copies are
    DMA from page-locked buffers."""
    import torch
    from lowrankmodels_b200 import _abi
    ptrfn = {np.dtype(np.float64): _abi.dptr, np.dtype(np.int32): _abi.i32ptr, np.dtype(np.int64): _abi.i64ptr}
    hold = []
    for name, arr in list(ep.keep.items()):
        flat = np.ascontiguousarray(arr).reshape(-1)
        t = torch.empty(flat.shape[0], dtype=torch.from_numpy(flat[:1].copy()).dtype).pin_memory()
        v = t.numpy()
        v[:] = flat
        hold.append(t)
        ep.keep[name] = v
        setattr(ep.struct, name, ptrfn[v.dtype](v))
    ep.keep["_pinned_tensors"] = hold
    return ep


def pinned_like(a):
    import torch
    t = torch.empty(a.size, dtype=torch.float64).pin_memory()
    v = t.numpy().reshape(a.shape, order="F")
    v[...] = a
    return v, t


def traj_rel_err(got, want):
    got, want = np.asarray(got, float), np.asarray(want, float)
    if len(got) != len(want):
        return float("inf")
    fin = np.isfinite(want)
    if not (np.isfinite(got) == fin).all():
        return float("inf")
    return float(np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin]))) if fin.any() else 0.0


def parity_vs_oracle(config, g, ep, eng_factory, iters=3, twin=None):
    """3-iteration objective trajectory of the engine against the sparse-evaluated CPU oracle on the same bytes.  `twin`
    (config, kwargs, words) replaces the full-size problem where the oracle would take minutes."""
    import lowrankmodels_b200 as lrm
    orc = _oracle()
    words = "full size"
    if twin is not None:
        g, _ = build_problem(config, **twin[0])
        ep = lrm.encode_problem(g, validate=False)
        words = twin[1]
    p = lrm.ProxGradParams(max_iter=iters, abs_tol=0, rel_tol=0)
    X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
    t = time.time()
    want = orc.fit(ep, lrm.encode_params(p), X, Y, mode=1, nthreads=physical_cores())["objective"]
    t_orc = time.time() - t
    eng = eng_factory(ep)
    Xe, Ye = g.X.copy(order="F"), g.Y.copy(order="F")
    got, _ = eng.fit(p, Xe, Ye)
    eng.close()
    err = traj_rel_err(got, want)
    return {"against": "oracle/glrm_oracle.c sparse-evaluated form, same bytes", "problem": words, "iterations": iters,
            "max_rel_err": err, "tolerance": PARITY_TOL, "ok": bool(err <= PARITY_TOL), "oracle_seconds": round(t_orc, 2),
            "objective_last": [float(got[-1]), float(want[-1])]}


def dense_roofline(m, n, d, k, x_ms, y_ms, T_x, T_y, pk):
    """SURVEY.md section 8d, dense path: per pass m*n*8 + (m+d)*k*8 bytes; 4*m*d*k flops (gradient pass), 2*m*d*k (trial
    pass).  A sweep = one gradient pass + T trial passes (T = measured mean trials per unit)."""
    pass_bytes = m * n * 8.0 + (m + d) * k * 8.0
    out = {}
    for side, ms, T in (("update_x", x_ms, T_x), ("update_y", y_ms, T_y)):
        flops = (4.0 + 2.0 * T) * m * d * k
        byts = (1.0 + T) * pass_bytes
        out[side] = {"ms": ms, "mean_trials": T, "algorithmic_flops": flops, "algorithmic_bytes": byts,
                     "fp64_tflops": flops / (ms * 1e-3) / 1e12, "fp64_frac": flops / (ms * 1e-3) / 1e12 / pk["fp64_tflops"],
                     "hbm_gbs": byts / (ms * 1e-3) / 1e9, "hbm_frac": byts / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
    out["peaks"] = {"fp64_tflops": pk["fp64_tflops"], "fp64_source": "tools/microbench.cu dmma_peak / dfma_peak, the larger (profiles/r2_microbench_fp64.json)",
                    "hbm_gbs": pk["hbm_gbs"], "hbm_source": pk["hbm_source"]}
    return out


def run_extra_dense(config, args, local, pk):
    """C4 / C5 at full size on one GPU through the fully observed path: timing, both roofline fractions, oracle parity."""
    import lowrankmodels_b200 as lrm
    t0 = time.time()
    g, cfg = build_problem(config)
    ep = lrm.encode_problem(g, validate=False)
    m, n = g.shape
    k, d, nnz = g.k, int(ep.struct.d), ep.nnz
    t_gen = time.time() - t0
    steps = max(3, min(args.steps, 5))
    warm = max(3, min(args.warmup, 5))
    t0 = time.time()
    eng = lrm.Engine(ep, device=local, validate=False)
    t_create = time.time() - t0
    eng.upload(g.X, g.Y)
    eng.fit_resident(lrm.ProxGradParams(max_iter=warm, abs_tol=0, rel_tol=0))
    obj, _ = eng.fit_resident(lrm.ProxGradParams(max_iter=steps, abs_tol=0, rel_tol=0))
    prof = dict(eng.last_profile)
    ms = prof["loop_ms"] / steps
    x_ms, y_ms = prof["update_x_ms"] / steps, prof["update_y_ms"] / steps
    T_x, T_y = prof["x_trials"] / (m * steps), prof["y_trials"] / (n * steps)
    out = {"config": config_static(config)["workload"], "n_gpus": 1, "value": nnz / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "steps": steps, "warmup": warm, "update_x_ms": x_ms, "update_y_ms": y_ms, "mean_trials": {"x": T_x, "y": T_y},
           "objective_first_last": [float(obj[0]), float(obj[-1])], "generate_s": round(t_gen, 1), "create_s": round(t_create, 2),
           "gpu_launches": int(prof["x_launches"] + prof["y_launches"] + prof["other_launches"]),
           "roofline": dense_roofline(m, n, d, k, x_ms, y_ms, T_x, T_y, pk)}
    # full-size check: the engine's objective at the start point against one CPU pass of the oracle
    orc = _oracle()
    t0 = time.time()
    got0 = eng.objective(g.X, g.Y, include_regularization=False)
    want0 = orc.objective(ep, g.X, g.Y, include_reg=False)
    out["full_size_objective_at_start"] = {"engine": float(got0), "oracle": float(want0),
                                           "rel_err": float(abs(got0 - want0) / abs(want0)), "seconds": round(time.time() - t0, 1)}
    eng.close()
    del eng, g, ep
    gc.collect()
    twin = ({"rows": 62_500}, "same generator and columns (1000 features, d=1800), 62500 rows") if config == "C4" else \
           ({"rows": 156_250}, "same generator and centroids, 156250 rows x 128")
    out["parity"] = parity_vs_oracle(config, None, None, lambda e: lrm.Engine(e, device=local, validate=False), twin=twin)
    return out


def sparse_roofline(k, nnz_local, units_local, ms, T, pk, side="x"):
    """SURVEY.md section 8d, sparse path, one sweep: (1 + T) * nnz * (4 + 8 + 8k) + units * k * 16 algorithmic bytes."""
    alg = (1.0 + T) * nnz_local * (4 + 8 + 8 * k) + units_local * k * 16
    achieved = alg / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
           "peak_source": pk["hbm_source"], "algorithmic_bytes_per_launch_set": alg, "mean_trials_per_unit": T,
           "ms_per_launch_set": ms}
    if pk["l2_gather_gbs"]:
        # the physically binding roof: both factors are L2-resident, the sweep is an L2->SM gather of 416-byte columns
        l2 = pk["l2_gather_gbs"]["y" if side == "x" else "x"]      # the X sweep gathers columns of Y and vice versa
        out["l2"] = {"achieved": achieved, "peak_measured": l2, "frac": achieved / l2, "unit": "GB/s",
                     "what": "same algorithmic bytes against the measured L2->SM gather rate of random 416-byte rows "
                             "(tools/microbench.cu ldg_d4, profiles/r2_microbench.json)"}
    return out


class Job:
    """torch.distributed plumbing of one bench process (one rank per GPU)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=self.rank, world_size=self.world)
        self.comm_ready = False

    def barrier(self):
        if self.world > 1:
            self.dist.all_reduce(self.torch.zeros(1))
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def engine(self, ep):
        import lowrankmodels_b200 as lrm
        eng = lrm.Engine(ep, device=self.local, rank=self.rank, nranks=self.world, validate=False)
        if self.world > 1:
            if not self.comm_ready:
                uid = [lrm.Engine.unique_id() if self.rank == 0 else None]
                self.dist.broadcast_object_list(uid, src=0)
                eng.comm_init(uid[0])
                self.comm_ready = True
            else:
                eng.comm_init(None)         # the per-process NCCL communicator is cached inside the library
            if not os.environ.get("GLRMB200_NO_PEER") and not ep.struct.obs_full:
                eng.peer_init(self.dist)    # fused exchange: peer stores from the update kernels (CUDA IPC over NVLink)
        return eng


def value_leg(job, g, ep, pk, sample_clocks):
    """W warm-up + K timed steps with everything resident; -> dict of timings, trials, roofline (max over ranks)."""
    import lowrankmodels_b200 as lrm
    args, torch, dist = job.args, job.torch, job.dist
    nnz, k = ep.nnz, g.k
    m, n = g.shape
    dense = bool(ep.struct.obs_full)
    X0, _tx = pinned_like(g.X)
    Y0, _ty = pinned_like(g.Y)
    eng = job.engine(ep)
    eng.upload(X0, Y0)
    pw = lrm.ProxGradParams(max_iter=max(args.warmup, 1), abs_tol=0, rel_tol=0)
    pt = lrm.ProxGradParams(max_iter=args.steps, abs_tol=0, rel_tol=0)
    job.barrier()
    eng.fit_resident(pw)                      # W untimed warm-up steps
    job.barrier()
    sampler = ClockSampler(job.local)
    if sample_clocks and job.rank == 0:
        sampler.start()
        time.sleep(0.3)
    job.barrier()
    t0 = time.time()
    obj, _ = eng.fit_resident(pt)             # exactly K timed steps (loop_ms excludes the setup objective)
    job.barrier()
    t1 = time.time()
    prof = dict(eng.last_profile)
    clocks = sampler.stop(t0, t1) if (sample_clocks and job.rank == 0) else None
    loop_ms = job.max_over_ranks(prof["loop_ms"])
    ms_per_step = loop_ms / args.steps
    per_rank = None
    if job.world > 1:
        t = torch.tensor([prof["update_x_ms"], prof["update_y_ms"], prof["comm_ms"], prof["loop_ms"]], dtype=torch.float64)
        allr = [torch.zeros_like(t) for _ in range(job.world)]
        dist.all_gather(allr, t)
        per_rank = [[round(float(v) / args.steps, 4) for v in r] for r in allr]
    x_ms = job.max_over_ranks(prof["update_x_ms"]) / args.steps
    y_ms = job.max_over_ranks(prof["update_y_ms"]) / args.steps
    comm_ms = job.max_over_ranks(prof["comm_ms"]) / args.steps
    rb, re_, cb, ce = eng.shard()
    rows_local = re_ - rb
    T_x = prof["x_trials"] / max(1, rows_local * args.steps)
    T_y = prof["y_trials"] / max(1, (ce - cb) * args.steps)
    # ---- roofline of the dominant kernel (update-X), SURVEY.md section 8d ------------------------------
    if dense:
        # (several GPUs: rows are sharded, every rank passes over its own rows in both sweeps; Y's trials are counted on rank 0)
        dr = dense_roofline(rows_local, n, int(ep.struct.d), k, prof["update_x_ms"] / args.steps, prof["update_y_ms"] / args.steps,
                            T_x, prof["y_trials"] / max(1, n * args.steps), pk)
        roofline = {"bound": "tensor (FP64 DMMA pipe; the same 37 TFLOP/s as the FP64 FMA pipe on B200)", "kernel": "dense_mma_x_kernel (one X sweep)",
                    "achieved": dr["update_x"]["fp64_tflops"], "peak": pk["fp64_tflops"], "unit": "TFLOP/s",
                    "frac": dr["update_x"]["fp64_frac"], "traffic": None, "both": dr}
    else:
        nnz_x_local = int(ep.keep["row_ptr"][re_] - ep.keep["row_ptr"][rb])
        roofline = sparse_roofline(k, nnz_x_local, rows_local, prof["update_x_ms"] / args.steps, T_x, pk, side="x")
        roofline["kernel"] = "update-X (sweep_cluster_kernel + sweep_cta_kernel + sweep_warp_kernel launches of one X sweep)"
        roofline["traffic"] = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_update_x_traffic.json")))
            roofline["traffic"] = tr["dram_bytes_per_step"]
            roofline["traffic_source"] = tr.get("what")
        except Exception:
            pass
        roofline["note"] = ("algorithmic bytes are cache-oblivious (412 B per entry-pass at k=50); both factors fit the 126 MB L2, "
                            "so achieved can exceed the HBM peak (SURVEY.md section 8d caveat) - `l2` is the same traffic "
                            "against the measured L2->SM gather roof, the physically binding one")
    eng.close()
    return {"value": nnz / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "update_x_ms": x_ms, "update_y_ms": y_ms,
            "comm_ms": comm_ms, "mean_trials": {"x": T_x, "y": T_y}, "per_rank_ms_x_y_comm_loop": per_rank,
            "objective_first_last": [float(obj[0]), float(obj[-1])], "wall_seconds_timed_call": t1 - t0, "clocks": clocks,
            "roofline": roofline, "gpu_launches": int(prof["x_launches"] + prof["y_launches"] + prof["other_launches"])}


def run_ours(args):
    # exactly ONE line may reach stdout (NCCL prints its version banner there): park fd 1 on stderr until the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import lowrankmodels_b200 as lrm
    job = Job(args)
    rank, world = job.rank, job.world
    pk = peaks()
    t_gen = time.time()
    g, cfg = build_problem(args.config, args.scale)
    ep = lrm.encode_problem(g, validate=False)
    nnz = ep.nnz
    m, n = g.shape
    dense = bool(ep.struct.obs_full)
    log(f"[rank {rank}] problem ready in {time.time() - t_gen:.1f}s: {m}x{n}, nnz={nnz}, k={g.k}")
    if not dense:
        ep = pin_encoded(ep)
    v = value_leg(job, g, ep, pk, sample_clocks=True)

    # ---- e2e leg: the reference-facing call with host buffers ------------------------------------------
    pt = lrm.ProxGradParams(max_iter=args.steps, abs_tol=0, rel_tol=0)
    Xh, _t1 = pinned_like(g.X)
    Yh, _t2 = pinned_like(g.Y)
    job.barrier()
    e0 = time.time()
    eng2 = lrm.Engine(ep, device=job.local, rank=rank, nranks=world, validate=False)
    e1 = time.time()
    if world > 1:
        eng2.comm_init(None)
        if not os.environ.get("GLRMB200_NO_PEER") and not dense:
            eng2.peer_init(job.dist)
    e2 = time.time()
    obj2, _ = eng2.fit(pt, Xh, Yh)
    e3 = time.time()
    eng2.close()
    e4 = time.time()
    job.barrier()
    e2e_s = job.max_over_ranks(time.time() - e0)
    log(f"[rank {rank}] e2e breakdown: create {e1 - e0:.3f}s, comm+peer {e2 - e1:.3f}s, fit {e3 - e2:.3f}s, close {e4 - e3:.3f}s")
    prob_bytes = sum(a.nbytes for name, a in ep.keep.items() if isinstance(a, np.ndarray))
    fac_bytes = (g.X.nbytes + g.Y.nbytes)
    e2e = {"value": nnz / (e2e_s / args.steps), "unit": UNIT,
           "h2d_bytes_per_step": (prob_bytes + fac_bytes) / args.steps,
           "d2h_bytes_per_step": (fac_bytes + 8 * (args.steps + 1)) / args.steps,
           "call": "glrmb200_create + glrmb200_fit(host X,Y; max_iter=K) + glrmb200_destroy, pinned host buffers; "
                   "one call runs all K steps, so per-step bytes are the call's bytes / K",
           "seconds_per_call": e2e_s}

    # ---- this very configuration against the CPU oracle (one GPU; the sparse oracle finishes full size in seconds) ----
    parity = None
    if world == 1 and not dense:
        try:
            parity = parity_vs_oracle(args.config, g, ep, job.engine)
        except Exception as ex:   # never lose the bench line to the checker
            parity = {"error": repr(ex)}

    # ---- the other BASELINE configurations ----------------------------------------------------------------------------
    extra = {}
    wanted = [c for c in args.extra.split(",") if c and c != args.config] if args.scale == 1 else []
    cpu = None
    if world == 1 and not args.no_cpu and not dense:
        r = cpu_faithful_sample(g, ep, 2, 1, budget_s=20.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": r["sample"] + "; 2 timed sample steps after 1 warm-up; XY through " + r["dgemm"],
               "seconds_per_iter": r["seconds_per_iter"]}
        val2, per2, _ = cpu_sparse_evaluated(g, ep, iters=10)
        cpu["sparse_evaluated_value"] = val2
        cpu["sparse_evaluated_note"] = "best-effort CPU (observed entries only), full size, 10 timed iterations, same cores"
    del g, ep, Xh, Yh, _t1, _t2
    gc.collect()
    for c in wanted:
        try:
            t = time.time()
            if c in ("C2", "C3"):                                  # sparse: any number of GPUs
                g3, _ = build_problem(c)
                ep3 = pin_encoded(lrm.encode_problem(g3, validate=False))
                r = value_leg(job, g3, ep3, pk, sample_clocks=False)
                r.pop("clocks")
                r = dict(config=config_static(c)["workload"], n_gpus=world, unit=UNIT, steps=args.steps, warmup=args.warmup, **r)
                if world == 1:
                    r["parity"] = parity_vs_oracle(c, g3, ep3, job.engine)
                extra[c] = r
                del g3, ep3
            elif world == 1:
                extra[c] = run_extra_dense(c, args, job.local, pk)
            if rank == 0 and c in extra:
                log(f"[extra {c}] {time.time() - t:.1f}s: {json.dumps(extra[c])[:600]}")
        except Exception as ex:
            extra[c] = {"error": repr(ex)}
            log(f"[extra {c}] failed: {ex!r}")
        gc.collect()

    if rank != 0:
        job.dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": v["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": v["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_static(args.config) if args.scale == 1 else
        {"workload": config_static(args.config)["workload"] + f" - /{args.scale} twin ({m}x{n}, {nnz} obs)", "l2": "twin: not a bench configuration"},
        "detail": {"parallelism": (f"fully observed: rows of A and X sharded over {world} GPU(s) by whole row-block groups, Y replicated; "
                                   "partial G_Y / per-feature objectives all-gathered per line-search round (NCCL) and summed in group order"
                                   if dense and world > 1 else
                                   f"rows/columns sharded over {world} GPU(s); " + ("single GPU" if world == 1 else
                                   "NCCL all-gather per half-iteration" if os.environ.get("GLRMB200_NO_PEER") else
                                   "accepted columns stored into every peer from the update kernels (CUDA IPC over NVLink), "
                                   "peer-memory flag barrier per half-iteration")),
                   **{key: v[key] for key in ("update_x_ms", "update_y_ms", "comm_ms", "mean_trials", "per_rank_ms_x_y_comm_loop",
                                             "objective_first_last", "wall_seconds_timed_call")}},
        "clocks": v["clocks"], "e2e": e2e, "roofline": v["roofline"], "parity": parity,
        "gpu_launches": v["gpu_launches"],
    }
    if extra:
        line["extra_configs"] = extra
    if cpu:
        line["cpu_baseline"] = cpu
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        job.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=int, default=1, help="1 = the full BASELINE configuration")
    ap.add_argument("--extra", default="C3,C4,C5", help="other configurations reported under extra_configs (N=1; '' = none)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        log("note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
