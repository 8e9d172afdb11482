"""Shared builders for the parity tests: small GLRMs on synthetic data (seeded, hash-based)."""
import numpy as np

import lowrankmodels_b200 as lrm
from lowrankmodels_b200 import synth


def glrm_from_config(cfg, loss, rx, ry, **kw):
    """cfg: a dict from lowrankmodels_b200.synth.config*()."""
    if cfg["full"]:
        return lrm.GLRM(cfg["A"], loss, rx, ry, cfg["k"], X=cfg["X0"].copy(), Y=cfg["Y0"].copy(), **kw)
    import scipy.sparse as sp
    A = sp.csc_matrix((cfg["vals"], (cfg["rows"], cfg["cols"])), shape=(cfg["m"], cfg["n"]))
    return lrm.GLRM(A, loss, rx, ry, cfg["k"], X=cfg["X0"].copy(), Y=cfg["Y0"].copy(), **kw)


def small_sparse(m=60, n=40, k=4, density=0.3, seed=3, labels=None, dup=False):
    """Random sparse pattern given as an explicit obs list (tests order/duplicate preservation)."""
    u = synth.uniform(seed, 51, np.arange(m * n)).reshape(m, n)
    ii, jj = np.nonzero(u < density)
    perm = np.argsort(synth.uniform(seed, 52, np.arange(len(ii))), kind="stable")  # shuffled obs order
    obs = np.stack([ii[perm], jj[perm]], axis=1)
    if dup:
        obs = np.concatenate([obs, obs[: len(obs) // 5]])                            # duplicates, as hello_world.jl:48
    P = synth.normal_matrix(seed, 53, m, 3)
    Q = synth.normal_matrix(seed, 54, 3, n)
    A = P @ Q + 0.1 * synth.normal_matrix(seed, 55, m, n)
    if labels == "bool":
        A = np.where(A >= 0, 1.0, -1.0)
    elif labels == "bool01":
        A = np.where(A >= 0, 1.0, 0.0)
    elif isinstance(labels, int):
        A = np.clip(np.floor((A - A.min()) / (A.max() - A.min() + 1e-9) * labels) + 1, 1, labels)
    elif labels == "count":
        A = np.floor(np.abs(A) * 2)
    X0 = synth.normal_matrix(seed, 56, k, m)
    return A, obs, X0


def run_oracle(orc, glrm, params, mode=1, nthreads=0):
    ep = lrm.encode_problem(glrm)
    X, Y = glrm.X.copy(order="F"), glrm.Y.copy(order="F")
    res = orc.fit(ep, lrm.encode_params(params), X, Y, mode=mode, nthreads=nthreads)
    res["X"], res["Y"] = X, Y
    return res


def assert_traj_close(a, b, rtol=1e-4, what=""):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert len(a) == len(b), f"{what}: trajectory lengths differ {len(a)} vs {len(b)}"
    for t, (x, y) in enumerate(zip(a, b)):
        if np.isinf(x) or np.isinf(y):
            assert x == y, f"{what}: entry {t}: {x} vs {y}"
        else:
            assert abs(x - y) <= rtol * abs(y), f"{what}: entry {t}: {x} vs {y} (rel {abs(x - y) / abs(y):.3e})"
