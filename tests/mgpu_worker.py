"""Launched by tests/test_gpu_multi.py under torch.distributed.run (one rank per GPU): the sharded fit
(NCCL all-gather per half-iteration inside the engine) must reproduce the single-GPU fit bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import lowrankmodels_b200 as lrm
    from helpers import glrm_from_config
    from lowrankmodels_b200 import distributed as D, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    ok = True
    for name, cfg, loss, reg in (("C2/16", synth.config2(scale=16), lrm.QuadLoss(), lrm.QuadReg(0.1)),
                                 ("C3/16", synth.config3(scale=16), lrm.LogisticLoss(), lrm.NonNegConstraint()),
                                 ("C1", synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1))):
        g = glrm_from_config(cfg, loss, reg, reg)
        ep = lrm.encode_problem(g)
        p = lrm.ProxGradParams(max_iter=6, abs_tol=0, rel_tol=0)
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        # two exchange paths: NCCL all-gather per half-iteration, and the fused peer-store epilogue (CUDA IPC)
        results = []
        for fused in (False, True):
            Xf, Yf = g.X.copy(order="F"), g.Y.copy(order="F")
            eng = lrm.Engine(ep, device=local, rank=rank, nranks=world)
            eng.comm_init(D.broadcast_unique_id(dist, rank, lrm.Engine.unique_id))
            if fused:
                eng.peer_init(dist)
            rb, re_, cb, ce = eng.shard()
            brow, bcol = D.shard_bounds(ep, world)
            assert (rb, re_, cb, ce) == (brow[rank], brow[rank + 1], bcol[rank], bcol[rank + 1])
            objf, _ = eng.fit(p, Xf, Yf)
            arf, acf = eng.stepsizes()
            eng.close()
            results.append((objf, Xf, Yf, arf, acf))
        (obj, X, Y, ar, ac), (obj2, X2, Y2, ar2, ac2) = results
        fused_same = (obj == obj2).all() and (X == X2).all() and (Y == Y2).all() and (ar == ar2).all() and (ac == ac2).all()
        if rank == 0:
            print(f"{name}: fused peer-store exchange identical to NCCL exchange = {fused_same}", flush=True)
            if not fused_same:
                print("   obj", np.max(np.abs(obj - obj2)), "X", np.max(np.abs(X - X2)), "Y", np.max(np.abs(Y - Y2)), flush=True)
            ok = ok and bool(fused_same)
        if rank == 0:
            X1, Y1 = g.X.copy(order="F"), g.Y.copy(order="F")
            # (C1 is too small to shard by row groups: the N-GPU fit goes through the gather kernels, and so does its twin)
            with lrm.Engine(ep, device=local, gather_only=(name == "C1")) as e1:
                obj1, _ = e1.fit(p, X1, Y1)
                ar1, ac1 = e1.stepsizes()
            same = (obj == obj1).all() and (X == X1).all() and (Y == Y1).all() and (ar == ar1).all() and (ac == ac1).all()
            print(f"{name}: {world}-GPU vs 1-GPU identical={same} obj_last={obj[-1]:.9e}", flush=True)
            if not same:
                print("   obj", np.max(np.abs(obj - obj1)), "X", np.max(np.abs(X - X1)), "Y", np.max(np.abs(Y - Y1)),
                      "alpha", np.max(np.abs(ar - ar1)), np.max(np.abs(ac - ac1)), flush=True)
            ok = ok and bool(same)
    # SparseProxGradParams path (unconditional sweeps, global accept / revert incl. rejected iterations), both exchanges
    cfg = synth.config2(scale=16)
    g = glrm_from_config(cfg, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    ep = lrm.encode_problem(g)
    sp = lrm.SparseProxGradParams(12.0, max_iter=12, abs_tol=1e-6)
    outs = []
    for fused in (False, True):
        Xf, Yf = g.X.copy(order="F"), g.Y.copy(order="F")
        eng = lrm.Engine(ep, device=local, rank=rank, nranks=world)
        eng.comm_init(D.broadcast_unique_id(dist, rank, lrm.Engine.unique_id))
        if fused:
            eng.peer_init(dist)
        objf, _ = eng.fit_sparse(sp, Xf, Yf)
        eng.close()
        outs.append((objf, Xf, Yf))
    if rank == 0:
        X1, Y1 = g.X.copy(order="F"), g.Y.copy(order="F")
        with lrm.Engine(ep, device=local) as e1:
            obj1, _ = e1.fit_sparse(sp, X1, Y1)
        same = all(len(o) == len(obj1) and (o == obj1).all() and (x == X1).all() and (y == Y1).all() for o, x, y in outs)
        print(f"sparse-params C2/16: {world}-GPU (NCCL and fused) vs 1-GPU identical={same} recorded={len(obj1)}", flush=True)
        ok = ok and bool(same)
    # fully observed problems (mode B): rows sharded by whole groups of row blocks, Y replicated, partial
    # G_Y / objectives all-gathered per line-search round; every rank gets its own rows of X back
    c5 = synth.config5(scale=512)
    c4 = synth.config4(scale=64)
    dense_cases = (
        ("C5/512 dense", lrm.GLRM(c5["A"], lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg(), c5["k"], X=c5["X0"], Y=c5["Y0"])),
        ("C4/64 dense", lrm.GLRM(c4["A"], [lrm.QuadLoss()] * c4["n_quad"] + [lrm.HingeLoss()] * c4["n_hinge"]
                                 + [lrm.MultinomialLoss(c4["levels"])] * c4["n_multi"], lrm.QuadReg(0.1), lrm.QuadReg(0.1), c4["k"],
                                 X=c4["X0"], Y=c4["Y0"])))
    for name, g in dense_cases:
        ep = lrm.encode_problem(g)
        p = lrm.ProxGradParams(max_iter=4, abs_tol=0, rel_tol=0)
        Xf, Yf = g.X.copy(order="F"), g.Y.copy(order="F")
        eng = lrm.Engine(ep, device=local, rank=rank, nranks=world)
        eng.comm_init(D.broadcast_unique_id(dist, rank, lrm.Engine.unique_id))
        rb, re_, _, _ = eng.shard()
        objf, _ = eng.fit(p, Xf, Yf)
        arf, acf = eng.stepsizes()
        total = eng.objective(Xf, Yf)
        eng.close()
        parts = [None] * world
        dist.all_gather_object(parts, (rb, re_, np.ascontiguousarray(Xf[:, rb:re_])))
        if rank == 0:
            Xall = np.full_like(Xf, np.nan)
            for b, e, blk in parts:
                Xall[:, b:e] = blk
            X1, Y1 = g.X.copy(order="F"), g.Y.copy(order="F")
            with lrm.Engine(ep, device=local) as e1:
                obj1, _ = e1.fit(p, X1, Y1)
                ar1, ac1 = e1.stepsizes()
                total1 = e1.objective(X1, Y1)
            same = ((objf == obj1).all() and (Xall == X1).all() and (Yf == Y1).all() and (arf == ar1).all() and (acf == ac1).all()
                    and total == total1)
            print(f"{name}: {world}-GPU row-sharded vs 1-GPU identical={same} obj_last={objf[-1]:.9e} rows/rank={[e - b for b, e, _ in parts]}", flush=True)
            if not same:
                print("   obj", np.max(np.abs(objf - obj1)), "X", np.nanmax(np.abs(Xall - X1)), "Y", np.max(np.abs(Yf - Y1)),
                      "alpha", np.max(np.abs(arf - ar1)), np.max(np.abs(acf - ac1)), "objective()", total, total1, flush=True)
            ok = ok and bool(same)
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
