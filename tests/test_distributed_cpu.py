"""world_size-2 gloo tests (CPU): the host side of the N>1 path — shard planning through the C ABI, the
unique-id rendezvous, and the property the multi-GPU design rests on: "every rank sweeps its own cost-balanced
shard, then the updated factor is all-gathered" reproduces the unsharded sweep bit for bit (rows of X are
independent given Y, columns of Y given X: proxgrad.jl:118,162)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def allgather_columns(dist, mat, bounds, elems_per_unit=1):
    """In-place all-gather of a column-major (k, units) array whose unit ranges [bounds[r], bounds[r+1]) are
    owned by rank r — the host-side twin of the engine's allgather_units."""
    import torch
    world = len(bounds) - 1
    for r in range(world):
        lo, hi = int(bounds[r]) * elems_per_unit, int(bounds[r + 1]) * elems_per_unit
        if hi > lo:
            t = torch.from_numpy(np.ascontiguousarray(mat[..., lo:hi].T if mat.ndim == 2 else mat[lo:hi]))
            dist.broadcast(t, src=r)
            if mat.ndim == 2:
                mat[:, lo:hi] = t.numpy().T
            else:
                mat[lo:hi] = t.numpy()
    return mat


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lowrankmodels_b200 as lrm
        import oracle_py
        from helpers import glrm_from_config
        from lowrankmodels_b200 import distributed as D, synth

        cfg = synth.config2(scale=32)
        g = glrm_from_config(cfg, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
        ep = lrm.encode_problem(g)
        rb, cb = D.shard_bounds(ep, world)
        # every rank plans the same shards, they tile the ranges, and they balance the observations
        import torch
        t = torch.from_numpy(np.concatenate([rb, cb]).copy())
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert (t == ref).all()
        assert rb[0] == 0 and rb[-1] == cfg["m"] and cb[0] == 0 and cb[-1] == cfg["n"]
        # shards are balanced by predicted cost: 32-entry chunks (a partial chunk costs a whole one) + one chunk per unit
        cost = (np.diff(ep.keep["row_ptr"]) + 31) // 32 * 32 + 32
        cum = np.concatenate([[0], np.cumsum(cost)])
        cost_r = np.diff(cum[rb])
        assert cost_r.max() - cost_r.min() <= 2 * cost.max()
        # unique-id rendezvous (any 128 bytes stand in for ncclUniqueId on the CPU)
        uid = D.broadcast_unique_id(dist, rank, lambda: bytes(range(128)))
        assert uid == bytes(range(128))

        # sharded sweeps + all-gather == the unsharded oracle fit
        p = lrm.ProxGradParams(max_iter=3, abs_tol=0, rel_tol=0)
        ps = lrm.encode_params(p)
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        ar, ac = np.full(cfg["m"], p.stepsize), np.full(cfg["n"], p.stepsize)
        orow, ocol = np.zeros(cfg["m"]), np.zeros(cfg["n"])
        traj = []
        for _ in range(p.max_iter):
            oracle_py.half_sweep(ep, ps, X, Y, ar, 0, int(rb[rank]), int(rb[rank + 1]), orow)
            allgather_columns(dist, X, rb)
            oracle_py.half_sweep(ep, ps, X, Y, ac, 1, int(cb[rank]), int(cb[rank + 1]), ocol)
            allgather_columns(dist, Y, cb)
            allgather_columns(dist, ocol, cb)
            traj.append(float(np.sum(ocol)))
        Xo, Yo = g.X.copy(order="F"), g.Y.copy(order="F")
        want = oracle_py.fit(ep, ps, Xo, Yo, mode=1, nthreads=1)
        assert (X == Xo).all() and (Y == Yo).all()
        np.testing.assert_allclose(traj, want["objective"][1:], rtol=1e-13)

        # fully observed problems (mode B): rows are sharded by whole groups of row blocks (host-only planner), every rank sums
        # its groups, the 8 group sums are gathered and totalled in group order -> the same bits as one process
        m_dense = 19531
        b2, b8 = D.plan_dense_rows(m_dense, world), D.plan_dense_rows(m_dense, 8)
        assert b2[0] == 0 and b2[-1] == m_dense and set(b2.tolist()) <= set(b8.tolist())      # shards are unions of groups
        vals = synth.normal_matrix(7, 9, 5, m_dense)                                         # a (5, m) array summed over rows
        gsum = np.zeros((8, 5))
        for gidx in range(8):
            lo, hi = int(b8[gidx]), int(b8[gidx + 1])
            if int(b2[rank]) <= lo and hi <= int(b2[rank + 1]):                               # a group this rank owns
                for e in range(lo, hi):
                    gsum[gidx] += vals[:, e]
        owner = [next(r for r in range(world) if int(b2[r]) <= int(b8[gi]) and int(b8[gi + 1]) <= int(b2[r + 1])) for gi in range(8)]
        for gidx in range(8):
            tg = torch.from_numpy(gsum[gidx].copy())
            dist.broadcast(tg, src=owner[gidx])
            gsum[gidx] = tg.numpy()
        total = np.zeros(5)
        for gidx in range(8):
            total += gsum[gidx]
        alone = np.zeros(5)
        for gidx in range(8):
            part = np.zeros(5)
            for e in range(int(b8[gidx]), int(b8[gidx + 1])):
                part += vals[:, e]
            alone += part
        assert (total == alone).all()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_sweep_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
