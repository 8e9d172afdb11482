#!/usr/bin/env python
"""bench.py — observed-entries / sec / iteration of the GLRM prox-grad hot path on B200.

Workload (BASELINE.json configs[1], "C2"): MovieLens-20M-shaped sparse A (138 493 x 26 744,
20 000 263 observations), QuadLoss + QuadReg(0.1), k = 50, Float64, synthetic (hash-generated) data.
A *step* is one outer iteration of fit!(glrm, ProxGradParams) (proxgrad.jl:107-217): one X sweep,
one Y sweep, the objective record.

    python bench.py --gpus N --steps K --warmup W            our engine (N>1: launched by torchrun)
    python bench.py --impl reference ...                     the reference's algorithm on the host CPUs
                                                             (oracle port, faithful dense-XY form)

value  : whole-job entries/s/iter with the problem and the factors resident in HBM
e2e    : the same metric through the reference-facing call (glrmb200_create + glrmb200_fit with host
         buffers + read-back + destroy): every H2D / D2H copy is inside the timed region
"""
from __future__ import annotations

import argparse
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_problem(config, scale):
    """-> (glrm, loss/reg description) for C2 (QuadLoss+QuadReg) or C3 (LogisticLoss+NonNeg)."""
    import scipy.sparse as sp
    import lowrankmodels_b200 as lrm
    from lowrankmodels_b200 import synth
    if config == "C2":
        cfg, loss, reg = synth.config2(scale=scale), lrm.QuadLoss(), lrm.QuadReg(0.1)
    elif config == "C3":
        cfg, loss, reg = synth.config3(scale=scale), lrm.LogisticLoss(), lrm.NonNegConstraint()
    else:
        raise SystemExit(f"unknown config {config}")
    A = sp.csc_matrix((cfg["vals"], (cfg["rows"], cfg["cols"])), shape=(cfg["m"], cfg["n"]))
    g = lrm.GLRM(A, loss, reg, reg, cfg["k"], X=cfg["X0"], Y=cfg["Y0"], checknan=False)
    return g, cfg


def workload_name(config, scale, g, nnz):
    m, n = g.shape
    what = "QuadLoss+QuadReg(0.1)" if config == "C2" else "LogisticLoss+NonNegConstraint"
    tag = "" if scale == 1 else f" /{scale} twin"
    return f"{config}{tag}: MovieLens-20M-shaped sparse {m}x{n}, {nnz} obs, {what}, k={g.k}"


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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
def cpu_reference_run(config, steps, warmup, twin_scale=8, mode=0):
    """The reference's algorithm (oracle port) on the host cores, on the /twin_scale twin of the workload:
    the faithful form materialises XY (m x d Float64 — 29.6 GB at full C2) and costs ~12*m*d*k flop per
    iteration regardless of sparsity, so the bounded sample is the twin (same density => same
    entries-per-flop ratio).  Returns (entries/s/iter, seconds per iter, nnz, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import lowrankmodels_b200 as lrm
    g, cfg = build_problem(config, twin_scale)
    ep = lrm.encode_problem(g, validate=False)
    nnz = ep.nnz
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it for the CPU arm)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    oracle_py.lib()
    # thread count: the static row/column chunks of Threads.@threads stop scaling past the physical cores on
    # some hosts, so calibrate on one iteration (all threads vs half) and keep the faster setting
    p1 = lrm.ProxGradParams(max_iter=1, abs_tol=0, rel_tol=0)
    best = None
    for nt in sorted({threads, max(1, threads // 2)}, reverse=True):
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        t = oracle_py.fit(ep, lrm.encode_params(p1), X, Y, mode=mode, nthreads=nt)["seconds"][1]
        if best is None or t < best[0]:
            best = (t, nt)
    threads = best[1]
    X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
    p = lrm.ProxGradParams(max_iter=warmup + steps, abs_tol=0, rel_tol=0)
    res = oracle_py.fit(ep, lrm.encode_params(p), X, Y, mode=mode, nthreads=threads)
    sec = res["seconds"][1 + warmup:]
    per_iter = float(np.mean(sec))
    return nnz / per_iter, per_iter, nnz, threads, workload_name(config, twin_scale, g, nnz)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, per_iter, nnz, threads, wl = cpu_reference_run(args.config, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_iter * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, 1, *_shape_only(args.config)),
                   "sample": wl, "algorithm": "proxgrad_multithread.jl as written (dense XY, dense line-search "
                   "products), C/OpenMP port: Julia is not installed on this box"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": wl},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class _Shape:
    def __init__(self, m, n, k):
        self.shape, self.k = (m, n), k


def _shape_only(config):
    from lowrankmodels_b200 import synth
    return _Shape(synth.ML20M["m"], synth.ML20M["n"], 50), synth.ML20M["nnz"]


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


def run_ours(args):
    # exactly ONE line may reach stdout (NCCL prints its version banner there): park fd 1 on stderr until the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import lowrankmodels_b200 as lrm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.all_reduce(torch.zeros(1))
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    t_gen = time.time()
    g, cfg = build_problem(args.config, args.scale)
    ep = lrm.encode_problem(g, validate=False)
    nnz = ep.nnz
    m, n = g.shape
    k = g.k
    log(f"[rank {rank}] problem ready in {time.time() - t_gen:.1f}s: {m}x{n}, nnz={nnz}, k={k}")
    ep = pin_encoded(ep)
    X0, _tx = pinned_like(g.X)
    Y0, _ty = pinned_like(g.Y)

    # ---- value leg: everything resident ---------------------------------------------------------------
    eng = lrm.Engine(ep, device=local, rank=rank, nranks=world, validate=False)
    if world > 1:
        uid = [lrm.Engine.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0])
        if not os.environ.get("GLRMB200_NO_PEER"):
            eng.peer_init(dist)             # fused exchange: peer stores from the update kernels (CUDA IPC over NVLink)
    eng.upload(X0, Y0)
    pw = lrm.ProxGradParams(max_iter=max(args.warmup, 1), abs_tol=0, rel_tol=0)
    pk = lrm.ProxGradParams(max_iter=args.steps, abs_tol=0, rel_tol=0)
    barrier()
    eng.fit_resident(pw)                      # W untimed warm-up steps
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.time()
    obj, _ = eng.fit_resident(pk)             # exactly K timed steps (loop_ms excludes the setup objective)
    barrier()
    t1 = time.time()
    prof = dict(eng.last_profile)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    loop_ms = max_over_ranks(prof["loop_ms"])
    ms_per_step = loop_ms / args.steps
    value = nnz / (ms_per_step * 1e-3)
    per_rank = None
    if world > 1:
        t = torch.tensor([prof["update_x_ms"], prof["update_y_ms"], prof["comm_ms"], prof["loop_ms"]], dtype=torch.float64)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = [[round(float(v) / args.steps, 4) for v in r] for r in allr]
    x_ms = max_over_ranks(prof["update_x_ms"]) / args.steps
    y_ms = max_over_ranks(prof["update_y_ms"]) / args.steps
    comm_ms = max_over_ranks(prof["comm_ms"]) / args.steps

    # ---- roofline of the dominant kernel (update-X), SURVEY.md section 8d ------------------------------
    rb, re_, cb, ce = eng.shard()
    rows_local = re_ - rb
    nnz_x_local = int(ep.keep["row_ptr"][re_] - ep.keep["row_ptr"][rb])
    T_x = prof["x_trials"] / max(1, rows_local * args.steps)
    T_y = prof["y_trials"] / max(1, (ce - cb) * args.steps)
    bytes_per_entry_pass = 4 + 8 + 8 * k
    alg_bytes_x = (1.0 + T_x) * nnz_x_local * bytes_per_entry_pass + rows_local * k * 16
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes_x / (prof["update_x_ms"] / args.steps * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_update_x_traffic.json")))["dram_bytes_per_step"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "update-X (sweep_cta_kernel + sweep_warp_kernel launches of one X sweep)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch_set": alg_bytes_x, "mean_trials_per_row": T_x,
                "ms_per_launch_set": prof["update_x_ms"] / args.steps,
                "note": "algorithmic bytes are cache-oblivious (412 B per entry-pass at k=50); both factors fit "
                        "the 126 MB L2, so achieved can exceed the HBM peak (SURVEY.md section 8d caveat)"}
    eng.close()

    # ---- e2e leg: the reference-facing call with host buffers ------------------------------------------
    Xh, _t1 = pinned_like(g.X)
    Yh, _t2 = pinned_like(g.Y)
    barrier()
    e0 = time.time()
    eng2 = lrm.Engine(ep, device=local, rank=rank, nranks=world, validate=False)
    e1 = time.time()
    if world > 1:
        eng2.comm_init(None)                # the per-process NCCL communicator is cached inside the library
        if not os.environ.get("GLRMB200_NO_PEER"):
            eng2.peer_init(dist)
    e2 = time.time()
    obj2, _ = eng2.fit(pk, Xh, Yh)
    e3 = time.time()
    eng2.close()
    e4 = time.time()
    barrier()
    e2e_s = max_over_ranks(time.time() - e0)
    log(f"[rank {rank}] e2e breakdown: create {e1 - e0:.3f}s, comm+peer {e2 - e1:.3f}s, fit {e3 - e2:.3f}s, close {e4 - e3:.3f}s")
    prob_bytes = sum(v.nbytes for name, v in ep.keep.items() if isinstance(v, np.ndarray))
    fac_bytes = (g.X.nbytes + g.Y.nbytes)
    e2e = {"value": nnz / (e2e_s / args.steps), "unit": UNIT,
           "h2d_bytes_per_step": (prob_bytes + fac_bytes) / args.steps,
           "d2h_bytes_per_step": (fac_bytes + 8 * (args.steps + 1)) / args.steps,
           "call": "glrmb200_create + glrmb200_fit(host X,Y; max_iter=K) + glrmb200_destroy, pinned host buffers; "
                   "one call runs all K steps, so per-step bytes are the call's bytes / K",
           "seconds_per_call": e2e_s}

    if rank != 0:
        dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.scale, g, nnz),
                   "l2": "inputs larger than L2 (CSR+CSC index/value streams 480 MB + factors 66 MB vs 126 MB L2); "
                         "no explicit flush",
                   "parallelism": (f"rows/columns sharded over {world} GPU(s); " + ("single GPU" if world == 1 else
                                   "NCCL all-gather per half-iteration" if os.environ.get("GLRMB200_NO_PEER") else
                                   "accepted columns stored into every peer from the update kernels (CUDA IPC over NVLink), "
                                   "NCCL barrier per half-iteration")),
                   "update_x_ms": x_ms, "update_y_ms": y_ms, "comm_ms": comm_ms,
                   "mean_trials": {"x": T_x, "y": T_y}, "per_rank_ms_x_y_comm_loop": per_rank,
                   "objective_first_last": [float(obj[0]), float(obj[-1])],
                   "wall_seconds_timed_call": t1 - t0},
        "clocks": clocks, "e2e": e2e, "roofline": roofline,
        "gpu_launches": int(prof["x_launches"] + prof["y_launches"] + prof["other_launches"]),
    }
    if world == 1 and not args.no_cpu:
        val, per_iter, cnnz, threads, wl = cpu_reference_run(args.config, 2, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": wl + "; faithful dense-XY form (reference algorithm as written), "
                                                "2 timed iterations after 1 warm-up", "seconds_per_iter": per_iter}
        val2, per2, _, _, _ = cpu_reference_run(args.config, 3, 1, mode=1)
        line["cpu_baseline"]["sparse_evaluated_value"] = val2
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C3"])
    ap.add_argument("--scale", type=int, default=1, help="1 = the full BASELINE configuration")
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
