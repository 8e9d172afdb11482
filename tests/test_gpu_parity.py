"""GPU parity tests proper (run with -m gpu on the B200 box): the CUDA engine, called through the C ABI
(ctypes), against the CPU oracle on identical bytes.

Tolerance: BASELINE.json's north_star asks for objective trajectories within 1e-4 relative of the
reference; Float64 throughout lets these tests hold a much tighter band (1e-7) on the small and mid-size
cases, with the 1e-4 figure as the hard bound everywhere.  Observation bookkeeping is bit-exact
(tests/test_obs_bookkeeping.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

import lowrankmodels_b200 as lrm
from helpers import assert_traj_close, glrm_from_config, run_oracle, small_sparse
from lowrankmodels_b200 import _abi, synth

pytestmark = pytest.mark.gpu
TIGHT = 1e-7
SPEC = 1e-4


def engine_fit(g, params):
    X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
    with lrm.Engine(g) as eng:
        obj, sec = eng.fit(params, X, Y)
        ar, ac = eng.stepsizes()
        prof = eng.last_profile
    return dict(objective=obj, X=X, Y=Y, alpharow=ar, alphacol=ac, profile=prof)


def check(orc, g, params, rtol=TIGHT, factors=True):
    want = run_oracle(orc, g, params, mode=1)
    got = engine_fit(g, params)
    assert_traj_close(got["objective"], want["objective"], rtol, "engine vs oracle")
    if factors:
        np.testing.assert_allclose(got["alpharow"], want["alpharow"], rtol=1e-9)
        np.testing.assert_allclose(got["alphacol"], want["alphacol"], rtol=1e-9)
        np.testing.assert_allclose(got["X"], want["X"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(got["Y"], want["Y"], rtol=1e-5, atol=1e-8)
    assert got["profile"]["x_trials"] == want["trials"][0] or not factors
    assert got["profile"]["y_trials"] == want["trials"][1] or not factors
    return got, want


def test_config1_dense_quad_quadreg(orc):
    """BASELINE config 1 (examples/simple_glrms.jl fit_pca_nucnorm): dense 100x100, k=5, defaults."""
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    got, _ = check(orc, g, lrm.ProxGradParams(max_iter=40))
    assert got["profile"]["x_launches"] > 0 and got["profile"]["y_launches"] > 0


def test_basic_functionality_self_consistency(orc):
    """test/basic_functionality.jl:5-16 — ch.objective[end] equals ||A - X'Y||^2 under ZeroReg."""
    c = synth.config1(seed=5)
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg())
    got = engine_fit(g, lrm.Params(1, max_iter=60, abs_tol=1e-7, min_stepsize=1e-3))
    Ah = got["X"].T @ got["Y"]
    assert abs(np.linalg.norm(c["A"] - Ah) ** 2 - got["objective"][-1]) < 1e-9 * got["objective"][-1] + 1e-9


@pytest.mark.parametrize("dup", [False, True])
def test_sparse_obs_order_and_duplicates(orc, dup):
    A, obs, X0 = small_sparse(dup=dup)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.05), lrm.QuadReg(0.05), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 1, 4, A.shape[1]))
    check(orc, g, lrm.ProxGradParams(max_iter=15))


def test_nnmf_infeasible_start(orc):
    A, obs, X0 = small_sparse(seed=4)
    g = lrm.GLRM(np.abs(A), lrm.QuadLoss(), lrm.NonNegConstraint(), lrm.NonNegConstraint(), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 2, 4, A.shape[1]))
    got, _ = check(orc, g, lrm.ProxGradParams(max_iter=12))
    assert np.isinf(got["objective"][0]) and np.isfinite(got["objective"][1:]).all()


def test_logistic_nonneg(orc):
    A, obs, X0 = small_sparse(seed=6, labels="bool")
    g = lrm.GLRM(A, lrm.LogisticLoss(), lrm.NonNegConstraint(), lrm.NonNegConstraint(), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 3, 4, A.shape[1]))
    check(orc, g, lrm.ProxGradParams(max_iter=12))


def test_kmeans_unit_one_sparse_dense(orc):
    c = synth.config5(scale=20000, k=6, n=8, centroids=4)
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg())
    check(orc, g, lrm.ProxGradParams(max_iter=8))


SCALAR_LOSS_CASES = [
    ("huber", lambda: lrm.HuberLoss(0.7, crossover=0.6), None),
    ("l1", lambda: lrm.L1Loss(1.3), None),
    ("quantile", lambda: lrm.QuantileLoss(1.0, quantile=0.7), None),
    ("periodic", lambda: lrm.PeriodicLoss(2.5, 0.8), None),
    ("hinge", lambda: lrm.HingeLoss(1.2), "bool01"),
    ("whinge", lambda: lrm.WeightedHingeLoss(0.9, case_weight_ratio=2.0), "bool"),
    ("ordhinge", lambda: lrm.OrdinalHingeLoss(1, 5, 0.9), 5),
    ("poisson", lambda: lrm.PoissonLoss(), "count"),
]


@pytest.mark.parametrize("name,mk,labels", SCALAR_LOSS_CASES, ids=[c[0] for c in SCALAR_LOSS_CASES])
def test_each_scalar_loss(orc, name, mk, labels):
    A, obs, X0 = small_sparse(seed=12, labels=labels)
    g = lrm.GLRM(A, mk(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 4, obs=obs, X=0.3 * X0,
                 Y=0.3 * synth.normal_matrix(9, 5, 4, A.shape[1]))
    check(orc, g, lrm.ProxGradParams(max_iter=8), factors=name not in ("l1", "quantile", "hinge", "whinge", "ordhinge"),
          rtol=1e-6)


REG_CASES = [lrm.OneReg(0.05), lrm.QuadConstraint(1.5), lrm.NonNegOneReg(0.1), lrm.OneSparseConstraint(),
             lrm.KSparseConstraint(2), lrm.SimplexConstraint(), lrm.lastentry1(lrm.QuadReg(0.2)),
             lrm.lastentry_unpenalized(lrm.OneReg(0.1)), lrm.lastentry1(lrm.NonNegConstraint())]


@pytest.mark.parametrize("reg", REG_CASES, ids=lambda r: repr(r).replace(" ", ""))
def test_each_regularizer_on_rows(orc, reg):
    A, obs, X0 = small_sparse(seed=14, k=5)
    g = lrm.GLRM(A, lrm.QuadLoss(), reg, lrm.QuadReg(0.1), 5, obs=obs, X=np.abs(X0) * 0.4,
                 Y=synth.normal_matrix(9, 6, 5, A.shape[1]))
    check(orc, g, lrm.ProxGradParams(max_iter=8), rtol=1e-6)


def test_heterogeneous_scalar_columns_and_per_row_regs(orc):
    """test/hello_world.jl shape restricted to the scalar-embedding losses: per-column losses with
    different scales (per-entry loss lookup in the X sweep), per-row regularizers, duplicated obs."""
    m, n, k = 40, 8, 3
    u = synth.uniform(2, 61, np.arange(m * n)).reshape(m, n)
    z = synth.normal_matrix(2, 62, m, n)
    A = z.copy()
    A[:, 2] = np.where(z[:, 2] > 0, 1, -1)
    A[:, 3] = np.floor(u[:, 3] * 5) + 1
    A[:, 4] = np.where(z[:, 4] > 0, 1, 0)
    A[:, 7] = np.floor(u[:, 7] * 4)
    losses = [lrm.QuadLoss(1.5), lrm.HuberLoss(0.7), lrm.HingeLoss(1.2), lrm.OrdinalHingeLoss(1, 5, 0.9),
              lrm.LogisticLoss(1.1), lrm.QuadLoss(0.5), lrm.QuantileLoss(1.0, quantile=0.3), lrm.PoissonLoss()]
    rx = [lrm.QuadReg(0.1) if e % 4 == 0 else lrm.OneReg(0.05) if e % 4 == 1 else lrm.NonNegConstraint()
          if e % 4 == 2 else lrm.KSparseConstraint(2) for e in range(m)]
    ry = [lrm.QuadReg(0.1 + 0.01 * f) for f in range(n)]
    ii, jj = np.nonzero(u < 0.7)
    obs = np.concatenate([np.stack([ii, jj], axis=1), np.stack([ii, jj], axis=1)[:25]])
    g = lrm.GLRM(A, losses, rx, ry, k, obs=obs, X=np.abs(synth.normal_matrix(2, 63, k, m)) * 0.3,
                 Y=synth.normal_matrix(2, 64, k, n) * 0.3)
    check(orc, g, lrm.ProxGradParams(max_iter=10), rtol=1e-6, factors=False)


def mixed_problem(m=40, k=3, seed=2, dup=True):
    """test/hello_world.jl:5-45: real / boolean / ordinal / categorical columns with their own scales."""
    n = 11
    u = synth.uniform(seed, 61, np.arange(m * n)).reshape(m, n)
    z = synth.normal_matrix(seed, 62, m, n)
    A = z.copy()
    A[:, 2] = np.where(z[:, 2] > 0, 1, -1)
    A[:, 3] = np.floor(u[:, 3] * 5) + 1          # OrdinalHinge 1..5
    A[:, 4] = np.floor(u[:, 4] * 4) + 1          # BvS levels 1..4      (embedding dim 3)
    A[:, 5] = np.floor(u[:, 5] * 3) + 1          # Multinomial 1..3     (3)
    A[:, 6] = np.floor(u[:, 6] * 3) + 1          # OvA 1..3             (3)
    A[:, 7] = np.floor(u[:, 7] * 4)              # Poisson counts
    A[:, 9] = np.floor(u[:, 9] * 4) + 1          # Ordistic 1..4        (4)
    A[:, 10] = np.floor(u[:, 10] * 5) + 1        # MultinomialOrdinal 1..5 (4)
    losses = [lrm.QuadLoss(1.5), lrm.HuberLoss(0.7), lrm.HingeLoss(1.2), lrm.OrdinalHingeLoss(1, 5, 0.9),
              lrm.BvSLoss(4, 1.1), lrm.MultinomialLoss(3, 0.8), lrm.OvALoss(3, 1.3, bin_loss=lrm.HingeLoss(0.9)),
              lrm.PoissonLoss(), lrm.QuantileLoss(1.0, quantile=0.7), lrm.OrdisticLoss(4, 0.6),
              lrm.MultinomialOrdinalLoss(5, 0.7)]
    d = lrm.embedding_dim(losses)
    rx = [lrm.QuadReg(0.1) if e % 4 == 0 else lrm.OneReg(0.05) if e % 4 == 1 else lrm.NonNegConstraint()
          if e % 4 == 2 else lrm.KSparseConstraint(2) for e in range(m)]
    ry = [lrm.QuadReg(0.1 + 0.01 * f) if f % 2 == 0 else lrm.OneReg(0.03) for f in range(n)]
    ii, jj = np.nonzero(u < 0.7)
    obs = np.stack([ii, jj], axis=1)
    if dup:
        obs = np.concatenate([obs, obs[:25]])
    return lrm.GLRM(A, losses, rx, ry, k, obs=obs, X=np.abs(synth.normal_matrix(seed, 63, k, m)) * 0.3,
                    Y=synth.normal_matrix(seed, 64, k, d) * 0.3)


def test_vector_valued_losses_mixed_columns(orc):
    """Multinomial / OvA / BvS / Ordistic / MultinomialOrdinal block columns next to scalar ones (losses.jl:354-608)."""
    check(orc, mixed_problem(), lrm.ProxGradParams(max_iter=10), rtol=1e-6, factors=False)
    g = mixed_problem(m=25, k=5, seed=3, dup=False)
    lrm.add_offset(g)                                                   # lastentry1 rows, lastentry_unpenalized blocks
    check(orc, g, lrm.ProxGradParams(max_iter=6, inner_iter=2), rtol=1e-6, factors=False)


def test_ordinal_block_regularizers(orc):
    """OrdinalReg / MNLOrdinalReg on block columns with the offset wrappers on the rest (fit_dataframe.jl defaults)."""
    from test_oracle_trajectory import ordinal_problem
    check(orc, ordinal_problem(), lrm.ProxGradParams(max_iter=10), rtol=1e-6, factors=False)
    check(orc, ordinal_problem(m=60, k=7, seed=9), lrm.ProxGradParams(max_iter=6), rtol=1e-6, factors=False)


def test_config4_scaled_twin(orc):
    """BASELINE config 4 (/16 twin: 62 500 x 62, fully observed): 50 % QuadLoss, 30 % HingeLoss, 20 % MultinomialLoss(5), k=20."""
    c = synth.config4(scale=16)
    losses = ([lrm.QuadLoss()] * c["n_quad"] + [lrm.HingeLoss()] * c["n_hinge"] + [lrm.MultinomialLoss(c["levels"])] * c["n_multi"])
    g = lrm.GLRM(c["A"], losses, lrm.QuadReg(0.1), lrm.QuadReg(0.1), c["k"], X=c["X0"], Y=c["Y0"])
    got, want = check(orc, g, lrm.ProxGradParams(max_iter=3, abs_tol=0, rel_tol=0), rtol=1e-6, factors=False)
    assert_traj_close(got["objective"], want["objective"], SPEC)


def test_offset_and_inner_iterations(orc):
    A, obs, X0 = small_sparse(seed=8)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 4, 4, A.shape[1]), offset=True)
    check(orc, g, lrm.ProxGradParams(max_iter=6, inner_iter=2))


@pytest.mark.parametrize("k", [1, 3, 8, 13, 20, 33, 50, 64, 100, 130, 200])
def test_every_rank_tile(orc, k):
    """Each (G, R) lane-group tile: kp <= 8, 16, 32, 64, 128, 256."""
    A, obs, _ = small_sparse(seed=20, m=50, n=30)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, obs=obs,
                 X=0.5 * synth.normal_matrix(21, 1, k, 50), Y=0.5 * synth.normal_matrix(21, 2, k, 30))
    check(orc, g, lrm.ProxGradParams(max_iter=6))


@pytest.mark.parametrize("tile", ["16,2", "32,1", "32,2", "16,4"])
def test_alternative_tiles_same_result(orc, tile, monkeypatch):
    monkeypatch.setenv("GLRMB200_TILE", tile)
    A, obs, _ = small_sparse(seed=22, m=50, n=30)
    k = 50
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, obs=obs,
                 X=0.5 * synth.normal_matrix(23, 1, k, 50), Y=0.5 * synth.normal_matrix(23, 2, k, 30))
    check(orc, g, lrm.ProxGradParams(max_iter=5))


def test_heavy_units_cta_path(orc, monkeypatch):
    """Force every unit with >= 8 observations through the one-CTA-per-unit kernel."""
    monkeypatch.setenv("GLRMB200_HEAVY", "8")
    monkeypatch.setenv("GLRMB200_DENSE", "0")                          # fully observed cases below: gather kernels, implicit indices
    A, obs, X0 = small_sparse(seed=24, m=70, n=45, density=0.5)
    g = lrm.GLRM(A, lrm.LogisticLoss() if False else lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 4, obs=obs,
                 X=synth.normal_matrix(25, 1, 4, 70), Y=synth.normal_matrix(25, 2, 4, 45))
    check(orc, g, lrm.ProxGradParams(max_iter=8))
    c = synth.config1()
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.NonNegConstraint(), lrm.QuadReg(0.1))
    check(orc, g, lrm.ProxGradParams(max_iter=8))


def test_super_heavy_units_cluster_path(orc, monkeypatch):
    """Force units with >= 16 observations through the 8-CTA thread-block-cluster kernel (DSMEM reductions)."""
    monkeypatch.setenv("GLRMB200_HEAVY", "8")
    monkeypatch.setenv("GLRMB200_CLUSTER", "16")
    monkeypatch.setenv("GLRMB200_DENSE", "0")
    A, obs, X0 = small_sparse(seed=26, m=300, n=45, density=0.5)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.OneReg(0.05), 4, obs=obs,
                 X=synth.normal_matrix(27, 1, 4, 300), Y=synth.normal_matrix(27, 2, 4, 45))
    check(orc, g, lrm.ProxGradParams(max_iter=8))
    c = synth.config5(scale=5000, k=6, n=8, centroids=4)               # dense: 2000-entry columns, 8-entry rows
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg())
    check(orc, g, lrm.ProxGradParams(max_iter=6))
    Ab, obsb, Xb = small_sparse(seed=28, m=400, n=30, density=0.6, labels="bool")
    g = lrm.GLRM(Ab, lrm.LogisticLoss(), lrm.NonNegConstraint(), lrm.NonNegConstraint(), 4, obs=obsb, X=Xb,
                 Y=synth.normal_matrix(29, 3, 4, 30))
    check(orc, g, lrm.ProxGradParams(max_iter=8))


@pytest.mark.parametrize("cfgname", ["C2", "C3"])
def test_config2_and_3_scaled_twins(orc, cfgname):
    """The /8 twins of BASELINE configs 2 and 3 (17 311 x 3 343, 312 504 obs, k=50), fixed work."""
    if cfgname == "C2":
        cfg, loss, reg = synth.config2(scale=8), lrm.QuadLoss(), lrm.QuadReg(0.1)
    else:
        cfg, loss, reg = synth.config3(scale=8), lrm.LogisticLoss(), lrm.NonNegConstraint()
    g = glrm_from_config(cfg, loss, reg, reg)
    p = lrm.ProxGradParams(max_iter=10, abs_tol=0, rel_tol=0)
    got, want = check(orc, g, p, rtol=1e-6, factors=False)
    assert_traj_close(got["objective"], want["objective"], SPEC)
    assert got["profile"]["x_trials"] >= cfg["m"] * 10 * 0.9


def test_config5_scaled_twin_kmeans(orc):
    """BASELINE config 5 (/64 twin: 156 250 x 128 dense, k=100): QuadLoss, rx = UnitOneSparseConstraint, ry = ZeroReg
    (src/simple_glrms.jl:29-34).  Exercises the k<=128 tile, 128-entry rows and 156k-entry columns (cluster tier)."""
    c = synth.config5(scale=64)
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg())
    got, want = check(orc, g, lrm.ProxGradParams(max_iter=3, abs_tol=0, rel_tol=0), rtol=1e-6, factors=False)
    assert_traj_close(got["objective"], want["objective"], SPEC)
    assert np.isinf(got["objective"][0])                       # random start is infeasible for the one-hot constraint


@pytest.mark.parametrize("stepsize", [1.0, 12.0])
def test_sparse_proxgrad_params_semantics(orc, stepsize):
    """fit!(glrm, SparseProxGradParams()) (sparse_proxgrad.jl:21-130) on the engine: unconditional sweeps with one global
    step size, global accept / revert; stepsize 12 forces rejected iterations.  Series recorded as the reference does."""
    from test_oracle_sparse_params import problems
    for name, g in problems():
        p = lrm.SparseProxGradParams(stepsize, max_iter=14, abs_tol=1e-6)
        ep = lrm.encode_problem(g)
        Xo, Yo = g.X.copy(order="F"), g.Y.copy(order="F")
        want = orc.fit_sparse(ep, lrm.encode_sparse_params(p), Xo, Yo)
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        with lrm.Engine(ep, gather_only=True) as eng:
            obj, sec = eng.fit_sparse(p, X, Y)
        assert_traj_close(obj, want["objective"], 1e-7, name)
        np.testing.assert_allclose(X, Xo, rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(Y, Yo, rtol=1e-5, atol=1e-8)
        assert obj[-1] == obj[-2] and len(sec) == len(obj)
    cfg = synth.config2(scale=16)
    g = glrm_from_config(cfg, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    X0, Y0 = g.X.copy(), g.Y.copy()
    _, _, ch = lrm.fit_inplace(g, verbose=False)                       # sparse A -> SparseProxGradParams by default (fit.jl:13-15)
    assert ch.objective[-1] < ch.objective[0] and not (g.X == X0).all()


def test_objective_api_and_reg_scale(orc):
    A, obs, X0 = small_sparse(seed=30)
    g = lrm.GLRM(A, lrm.HuberLoss(), lrm.QuadReg(0.3), lrm.OneReg(0.2), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 7, 4, A.shape[1]))
    ep = lrm.encode_problem(g)
    with lrm.Engine(ep) as eng:
        for inc in (True, False):
            assert eng.objective(g.X, g.Y, inc) == pytest.approx(orc.objective(ep, g.X, g.Y, inc), rel=1e-12)
        eng.set_reg_scale(0.05)                                   # scale_regularizer! (glrm.jl:84-88)
        lrm.scale_regularizer(g, 0.05)
        ep2 = lrm.encode_problem(g)
        assert eng.objective(g.X, g.Y, True) == pytest.approx(orc.objective(ep2, g.X, g.Y, True), rel=1e-12)


def test_set_obs_swaps_lists_on_a_live_handle(orc):
    """cross_validate.jl:31-33 replaces the observation lists of a copy of the model and refits: glrmb200_set_obs does
    that on a live handle.  The result must equal a handle created from scratch on the training fold."""
    A, obs, X0 = small_sparse(seed=50, m=80, n=50, density=0.4)
    k = 4
    Y0 = synth.normal_matrix(51, 1, k, A.shape[1])
    full = lrm.GLRM(A, lrm.HuberLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, obs=obs, X=X0, Y=Y0)
    fold = synth.uniform(52, 1, np.arange(len(obs))) < 0.8
    train = lrm.GLRM(A, lrm.HuberLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, obs=obs[fold], X=X0, Y=Y0)
    p = lrm.ProxGradParams(max_iter=8)
    want = run_oracle(orc, train, p)
    Xa, Ya = train.X.copy(order="F"), train.Y.copy(order="F")
    with lrm.Engine(full) as eng:
        eng.fit(lrm.ProxGradParams(max_iter=2), full.X.copy(order="F"), full.Y.copy(order="F"))   # handle in use on all obs
        eng.set_obs(lrm.encode_problem(train))
        obj, _ = eng.fit(p, Xa, Ya)
    Xb, Yb = train.X.copy(order="F"), train.Y.copy(order="F")
    with lrm.Engine(train) as eng:
        obj2, _ = eng.fit(p, Xb, Yb)
    assert (obj == obj2).all() and (Xa == Xb).all() and (Ya == Yb).all()
    assert_traj_close(obj, want["objective"], TIGHT)
    # a fully observed handle switches to list mode
    c = synth.config1()
    gfull = glrm_from_config(c, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    ii, jj = np.nonzero(synth.uniform(53, 1, np.arange(100 * 100)).reshape(100, 100) < 0.5)
    gsub = lrm.GLRM(c["A"], lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 5, obs=np.stack([ii, jj], axis=1),
                    X=c["X0"], Y=c["Y0"])
    wsub = run_oracle(orc, gsub, p)
    Xc, Yc = gsub.X.copy(order="F"), gsub.Y.copy(order="F")
    with lrm.Engine(gfull) as eng:
        eng.set_obs(lrm.encode_problem(gsub))
        objc, _ = eng.fit(p, Xc, Yc)
    assert_traj_close(objc, wsub["objective"], TIGHT)


def test_fit_mutates_in_place_warm_starts_and_appends_ch(orc):
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    X_id, Y_id = id(g.X), id(g.Y)
    ch = lrm.ConvergenceHistory("t")
    X, Y, ch = lrm.fit_inplace(g, lrm.ProxGradParams(max_iter=5), ch=ch, verbose=False)
    assert X is g.X and Y is g.Y and id(g.X) == X_id and id(g.Y) == Y_id
    assert len(ch.objective) == 6 and len(ch.times) == 6 and ch.times[0] == 0
    first = ch.objective[-1]
    lrm.fit_inplace(g, lrm.ProxGradParams(max_iter=5), ch=ch, verbose=False)        # warm start, same ch
    assert len(ch.objective) == 12 and ch.objective[-1] <= first
    assert (np.diff(ch.times) >= 0).all()
    X0 = g.X.copy()
    Xt, Y2, ch2 = lrm.fit(g, lrm.ProxGradParams(max_iter=3), verbose=False)         # fit.jl:24-31
    assert (g.X == X0).all() and Xt.shape == (100, 5)


def test_errors_through_the_abi():
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    with lrm.Engine(g) as eng:
        with pytest.raises(_abi.GLRMB200Error) as ei:                               # norm(Y) == 0
            eng.fit(lrm.ProxGradParams(max_iter=2), g.X.copy(order="F"), np.zeros_like(g.Y, order="F"))
        assert ei.value.code == -1
    # device-side validation of the observations: NaN (glrm.jl:63-71), Boolean label domain (losses.jl:104),
    # index bounds — dense and list forms
    h = _abi.Handle()
    L = _abi.lib()
    ep = lrm.encode_problem(g)
    ep.keep["dense_A"][3, 4] = np.nan
    assert L.glrmb200_create(C.byref(h), C.byref(ep.struct), 0, 0, 1) == -7
    assert b"(4, 5) is NaN" in L.glrmb200_last_error()
    Ab, obs, X0 = small_sparse(seed=6, labels="bool")
    Ab[obs[5, 0], obs[5, 1]] = 2.0
    gb = lrm.GLRM(Ab, lrm.LogisticLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 4, obs=obs, X=X0)
    epb = lrm.encode_problem(gb, validate=False)
    assert L.glrmb200_create(C.byref(h), C.byref(epb.struct), 0, 0, 1) == -6
    assert f"({obs[5, 0] + 1}, {obs[5, 1] + 1})".encode() in L.glrmb200_last_error()
    gq = lrm.GLRM(Ab, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 4, obs=obs, X=X0)
    epq = lrm.encode_problem(gq)
    epq.keep["col_idx"][7] = 10_000
    assert L.glrmb200_create(C.byref(h), C.byref(epq.struct), 0, 0, 1) == -1
    assert b"col_idx[7]" in L.glrmb200_last_error()
    epq.keep["col_idx"][7] = 0
    epq.keep["row_val"][11] = np.nan
    assert L.glrmb200_create(C.byref(h), C.byref(epq.struct), 0, 0, 1) == -7
    A = np.floor(synth.uniform(1, 1, np.arange(60)).reshape(20, 3) * 3) + 1
    part = [(i, j) for i in range(20) for j in range(3) if (i, j) != (7, 1)]                    # not fully observed: gather kernels
    for g2 in (lrm.GLRM(A, lrm.MultinomialLoss(3), lrm.ZeroReg(), lrm.ZeroReg(), 40, obs=part),  # block columns need k <= 32 there
               lrm.GLRM(A * 3, lrm.MultinomialLoss(9), lrm.ZeroReg(), lrm.ZeroReg(), 2),        # embedding dim > 8
               lrm.GLRM(A, lrm.MultinomialLoss(3), lrm.ZeroReg(), lrm.UnitOneSparseConstraint(), 2)):  # non-separable block reg
        with pytest.raises(_abi.GLRMB200Error) as ei:
            lrm.Engine(g2)
        assert ei.value.code == -2                                                  # GLRMB200_E_UNSUPPORTED, never a CPU fallback


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (138 493 x 26 744, 20 000 263 obs, k=50): size-independent properties.
    (i) recorded objective decreases monotonically after entry 0; (ii) the last record equals
    objective(include_regularization=false) + sum_f ry(y_f) recomputed from the returned factors (quirk Q1);
    (iii) entry 0 equals the objective API on the start point; (iv) a second engine gives the same bits."""
    cfg = synth.config2(scale=1)
    g = glrm_from_config(cfg, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    p = lrm.ProxGradParams(max_iter=4, abs_tol=0, rel_tol=0)
    ep = lrm.encode_problem(g, validate=False)
    X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
    with lrm.Engine(ep) as eng:
        obj0 = eng.objective(X, Y, True)
        obj, _ = eng.fit(p, X, Y)
        assert obj[0] == pytest.approx(obj0, rel=1e-13)
        assert (np.diff(obj[1:]) < 0).all() and obj[1] < obj[0]
        loss_only = eng.objective(X, Y, False)
        assert obj[-1] == pytest.approx(loss_only + 0.1 * float(np.sum(Y * Y)), rel=1e-10)
    X2, Y2 = g.X.copy(order="F"), g.Y.copy(order="F")
    with lrm.Engine(ep) as eng:
        obj2, _ = eng.fit(p, X2, Y2)
    assert (obj2 == obj).all() and (X2 == X).all() and (Y2 == Y).all()           # deterministic reduction trees


# ---- the fully observed (dense) path: csrc/glrm_dense_mma.cuh (tensor cores); GLRMB200_DENSE_MMA=0: csrc/glrm_dense.cuh ----------------------------------------------------------------
def dense_problem(m=150, n=90, k=7, seed=40, losses=None, rx=None, ry=None, labels=None):
    P = synth.normal_matrix(seed, 1, m, 3)
    Q = synth.normal_matrix(seed, 2, 3, n)
    A = P @ Q + 0.1 * synth.normal_matrix(seed, 3, m, n)
    if labels == "bool":
        A = np.where(A >= 0, 1.0, -1.0)
    elif labels == "bool01":
        A = np.where(A >= 0, 1.0, 0.0)
    elif labels == "count":
        A = np.floor(np.abs(A) * 2)
    elif isinstance(labels, int):
        A = np.clip(np.floor((A - A.min()) / (A.max() - A.min() + 1e-9) * labels) + 1, 1, labels)
    losses = losses if losses is not None else lrm.QuadLoss()
    d = n if not isinstance(losses, list) else sum(l.embedding_dim() for l in losses)
    return lrm.GLRM(np.asfortranarray(A), losses, rx or lrm.QuadReg(0.1), ry or lrm.QuadReg(0.1), k,
                    X=0.4 * synth.normal_matrix(seed, 4, k, m), Y=0.4 * synth.normal_matrix(seed, 5, k, d))


@pytest.mark.parametrize("k", [1, 3, 8, 13, 20, 28, 33, 50, 64, 70, 96, 100, 104, 130])
def test_dense_path_every_rank_tile(orc, k):
    """Fully observed problems run the tensor-core kernels for k <= 104 (one instantiation per factor width NT = 1, 2, 3, 4,
    6, 8, 10, 12, 13 n-tiles of 8 indices: every one is hit here) and the gather kernels above; rows not a multiple of the
    128-row tile, 90 columns = three 32-column stages (resident in shared memory)."""
    check(orc, dense_problem(k=k), lrm.ProxGradParams(max_iter=5))


def test_dense_path_equals_gather_path(orc, monkeypatch):
    """The same fully observed fit through the dense kernels and (GLRMB200_DENSE=0) the gather kernels with implicit indices."""
    g = dense_problem(m=333, n=70, k=20)
    p = lrm.ProxGradParams(max_iter=8)
    a = engine_fit(g, p)
    monkeypatch.setenv("GLRMB200_DENSE", "0")
    b = engine_fit(g, p)
    assert_traj_close(a["objective"], b["objective"], 1e-9, "dense vs gather")
    np.testing.assert_allclose(a["alpharow"], b["alpharow"], rtol=1e-9)
    np.testing.assert_allclose(a["X"], b["X"], rtol=1e-6, atol=1e-9)
    assert a["profile"]["x_trials"] == b["profile"]["x_trials"] and a["profile"]["y_trials"] == b["profile"]["y_trials"]


def test_dense_tensor_core_and_fma_kernels_agree(orc, monkeypatch):
    """The two kernel families of the fully observed path (DMMA tensor-core kernels, and the FP64-FMA kernels that serve the
    shapes whose tensor-core tiles do not fit shared memory) on the same uniform and heterogeneous fits."""
    lv = 4
    m, n = 300, 40
    base = synth.normal_matrix(51, 1, m, 3) @ synth.normal_matrix(51, 2, 3, n)
    A = np.asfortranarray(base.copy())
    A[:, 20:30] = np.where(base[:, 20:30] >= 0, 1.0, -1.0)
    A[:, 30:] = np.clip(np.floor(np.abs(base[:, 30:]) * 2) + 1, 1, lv)
    mixed = [lrm.QuadLoss()] * 20 + [lrm.HingeLoss()] * 10 + [lrm.MultinomialLoss(lv)] * 10
    for g in (dense_problem(m=333, n=70, k=20, seed=50),
              lrm.GLRM(A, mixed, lrm.QuadReg(0.1), lrm.QuadReg(0.1), 12, X=0.3 * synth.normal_matrix(51, 4, 12, m),
                       Y=0.3 * synth.normal_matrix(51, 5, 12, 20 + 10 + 10 * lv))):
        p = lrm.ProxGradParams(max_iter=6)
        monkeypatch.setenv("GLRMB200_DENSE_MMA", "1")
        a = engine_fit(g, p)
        monkeypatch.setenv("GLRMB200_DENSE_MMA", "0")
        b = engine_fit(g, p)
        assert_traj_close(a["objective"], b["objective"], 1e-9, "tensor-core vs FMA dense kernels")
        np.testing.assert_allclose(a["alphacol"], b["alphacol"], rtol=1e-9)
        assert a["profile"]["x_trials"] == b["profile"]["x_trials"] and a["profile"]["y_trials"] == b["profile"]["y_trials"]
        assert_traj_close(a["objective"], run_oracle(orc, g, p, mode=1)["objective"], 1e-7, "tensor-core kernels vs oracle")


@pytest.mark.parametrize("reg", REG_CASES + [lrm.UnitOneSparseConstraint(), lrm.NonNegConstraint(), lrm.ZeroReg()],
                         ids=lambda r: repr(r).replace(" ", ""))
def test_dense_path_each_regularizer(orc, reg):
    g = dense_problem(m=100, n=40, k=5, seed=41, rx=reg, ry=lrm.QuadReg(0.05))
    check(orc, g, lrm.ProxGradParams(max_iter=6), rtol=1e-6, factors=False)
    g = dense_problem(m=80, n=30, k=5, seed=42, rx=lrm.QuadReg(0.05), ry=reg)
    check(orc, g, lrm.ProxGradParams(max_iter=6), rtol=1e-6, factors=False)


@pytest.mark.parametrize("name,mk,labels", SCALAR_LOSS_CASES + [("logistic", lambda: lrm.LogisticLoss(), "bool")],
                         ids=[c[0] for c in SCALAR_LOSS_CASES] + ["logistic"])
def test_dense_path_each_scalar_loss(orc, name, mk, labels):
    g = dense_problem(m=130, n=20, k=4, seed=43, losses=mk(), labels=labels)
    check(orc, g, lrm.ProxGradParams(max_iter=6), rtol=1e-6, factors=False)


def test_dense_path_heterogeneous_and_vector_losses(orc):
    """Scalar and vector-valued losses side by side, fully observed (the shape of config 4), incl. k above the gather
    path's k <= 32 limit for block columns."""
    m, n = 200, 24
    lv = 4
    A = np.empty((m, n), order="F")
    base = synth.normal_matrix(44, 1, m, n)
    A[:, :8] = base[:, :8]
    A[:, 8:14] = np.where(base[:, 8:14] >= 0, 1.0, -1.0)
    A[:, 14:] = np.clip(np.floor(np.abs(base[:, 14:]) * 2) + 1, 1, lv)
    losses = ([lrm.QuadLoss()] * 4 + [lrm.HuberLoss()] * 4 + [lrm.HingeLoss()] * 3 + [lrm.LogisticLoss()] * 3 +
              [lrm.MultinomialLoss(lv)] * 3 + [lrm.OvALoss(lv)] * 2 + [lrm.BvSLoss(lv)] * 2 + [lrm.OrdisticLoss(lv)] * 2 +
              [lrm.MultinomialOrdinalLoss(lv)])
    for k in (6, 40):
        d = sum(l.embedding_dim() for l in losses)
        g = lrm.GLRM(A, losses, lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, X=0.3 * synth.normal_matrix(44, 2, k, m),
                     Y=0.3 * synth.normal_matrix(44, 3, k, d))
        check(orc, g, lrm.ProxGradParams(max_iter=5), rtol=1e-6, factors=False)


@pytest.mark.parametrize("k,m,n", [(20, 40000, 21), (100, 24000, 70), (8, 50000, 130), (33, 30011, 64)])
def test_dense_path_persistent_tiles(orc, k, m, n):
    """More row tiles than resident CTAs (every CTA walks several tiles: the A-tile pipeline carries over from one tile to the
    next), mixed scalar / vector losses, one and two tile buffers, one and two CTAs per SM, a ragged last tile."""
    lv = 5
    base = synth.normal_matrix(49, 1, m, 3) @ synth.normal_matrix(49, 2, 3, n) + 0.3 * synth.normal_matrix(49, 3, m, n)
    nq, nh = n // 2, n // 4
    A = np.empty((m, n), order="F")
    A[:, :nq] = base[:, :nq]
    A[:, nq:nq + nh] = np.where(base[:, nq:nq + nh] >= 0, 1.0, -1.0)
    A[:, nq + nh:] = np.clip(np.floor(np.abs(base[:, nq + nh:]) * 2) + 1, 1, lv)
    losses = [lrm.QuadLoss()] * nq + [lrm.HingeLoss()] * nh + [lrm.MultinomialLoss(lv)] * (n - nq - nh)
    d = sum(l.embedding_dim() for l in losses)
    g = lrm.GLRM(A, losses, lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, X=0.3 * synth.normal_matrix(49, 4, k, m),
                 Y=0.3 * synth.normal_matrix(49, 5, k, d))
    check(orc, g, lrm.ProxGradParams(max_iter=3), rtol=1e-6, factors=False)


def test_dense_path_many_chunks_and_inner_iterations(orc):
    """d = 5 * 64 + 7 columns (six chunks), inner_iter = 2, offset wrappers."""
    g = dense_problem(m=70, n=327, k=9, seed=45)
    check(orc, g, lrm.ProxGradParams(max_iter=4, inner_iter=2), rtol=1e-6, factors=False)
    g = dense_problem(m=70, n=50, k=9, seed=46)
    lrm.add_offset(g)
    check(orc, g, lrm.ProxGradParams(max_iter=5), rtol=1e-6, factors=False)


def test_dense_handle_objective_and_set_obs(orc):
    """glrmb200_objective on a dense handle; set_obs turns it into a list-mode handle."""
    g = dense_problem(m=90, n=33, k=6, seed=47)
    ep = lrm.encode_problem(g)
    with lrm.Engine(g) as eng:
        for reg in (True, False):
            want = orc.objective(ep, g.X, g.Y, include_reg=reg)
            got = eng.objective(g.X, g.Y, include_regularization=reg)
            assert abs(got - want) <= 1e-12 * abs(want)
        A = np.asarray(g.A)
        ii, jj = np.nonzero(synth.uniform(48, 1, np.arange(90 * 33)).reshape(90, 33) < 0.5)
        sub = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 6, obs=np.stack([ii, jj], axis=1),
                       X=g.X.copy(), Y=g.Y.copy())
        eng.set_obs(lrm.encode_problem(sub))
        X, Y = sub.X.copy(order="F"), sub.Y.copy(order="F")
        obj, _ = eng.fit(lrm.ProxGradParams(max_iter=5), X, Y)
        want = run_oracle(orc, sub, lrm.ProxGradParams(max_iter=5))
        assert_traj_close(obj, want["objective"], TIGHT, "dense handle after set_obs")


# ---- regularizers with vector payloads (regularizers.jl:193-231,412-423) ---------------------------------------------------
def test_rem_quad_reg_mult_reg_shape(orc):
    """test/mult_reg.jl:5-29: one RemQuadReg(50, mean) per row and per column; trajectory vs oracle and the reference's own
    recovery thresholds."""
    n = m = 200
    r, eta, delta = 5, 0.01, 1e-3
    Um, Vm = synth.normal_matrix(71, 1, r, n), synth.normal_matrix(71, 2, r, m)
    U = Um + np.sqrt(eta) * synth.normal_matrix(71, 3, r, n)
    V = Vm + np.sqrt(eta) * synth.normal_matrix(71, 4, r, m)
    Yobs = U.T @ V + np.sqrt(delta) * synth.normal_matrix(71, 5, n, m)
    g = lrm.GLRM(np.asfortranarray(Yobs), lrm.QuadLoss(), [lrm.RemQuadReg(50, Um[:, i]) for i in range(n)],
                 [lrm.RemQuadReg(50, Vm[:, j]) for j in range(m)], r,
                 X=synth.normal_matrix(71, 6, r, n), Y=synth.normal_matrix(71, 7, r, m))
    got, _ = check(orc, g, lrm.ProxGradParams(max_iter=100))
    assert np.mean((U - got["X"]) ** 2) < 1e-3 and np.mean((V - got["Y"]) ** 2) < 1e-3


@pytest.mark.parametrize("k,nfix", [(4, 1), (5, 2), (7, 3), (20, 5), (20, 8), (50, 1), (50, 17), (100, 33), (100, 64)])
def test_fixed_latent_features_every_tile(orc, k, nfix):
    """fixed_latent_features on the columns and fixed_last_latent_features (restated with the reference's own indexing,
    regularizers.jl:223) on the rows, for every register tile shape and odd / even pinned lengths; the pinned entries
    never move (test/fixedfeatures_test.jl asserts exactly that)."""
    A, obs, X0 = small_sparse(m=70, n=40, k=k, seed=8)
    Y0 = 0.5 * synth.normal_matrix(72, 1, k, 40)
    X0 = 0.5 * X0
    ry = [lrm.fixed_latent_features(lrm.QuadReg(0.1), Y0[:nfix, j].copy()) for j in range(40)]
    rx = [lrm.fixed_last_latent_features(lrm.QuadReg(0.05), X0[k - nfix:, i].copy()) for i in range(70)]
    g = lrm.GLRM(A, lrm.QuadLoss(), rx, ry, k, obs=obs, X=X0.copy(), Y=Y0.copy())
    got, _ = check(orc, g, lrm.ProxGradParams(max_iter=8))
    assert (got["Y"][:nfix, :] == Y0[:nfix, :]).all() and (got["X"][k - nfix:, :] == X0[k - nfix:, :]).all()


@pytest.mark.parametrize("inner", [lrm.ZeroReg(), lrm.NonNegConstraint(), lrm.OneSparseConstraint(), lrm.UnitOneSparseConstraint(),
                                   lrm.KSparseConstraint(2), lrm.SimplexConstraint(), lrm.QuadConstraint(2.0), lrm.OneReg(0.1)],
                         ids=lambda r: type(r).__name__)
def test_fixed_latent_features_inner_regularizers(orc, inner):
    """The inner regularizer sees only the free entries (reductions, arg-max and sorting run over that sub-range)."""
    k = 6
    A, obs, X0 = small_sparse(m=50, n=30, k=k, seed=9)
    Y0 = np.abs(synth.normal_matrix(73, 1, k, 30))
    shared = lrm.FixedLatentFeaturesConstraint(np.array([0.25, 0.5]))            # one shared regularizer, inner = ZeroReg
    ry = [lrm.fixed_latent_features(inner.copy(), Y0[:2, j].copy()) for j in range(30)]
    g = lrm.GLRM(np.abs(A), lrm.QuadLoss(), shared, ry, k, obs=obs, X=np.abs(X0), Y=Y0.copy())
    g.X[:2, :] = np.array([[0.25], [0.5]])
    check(orc, g, lrm.ProxGradParams(max_iter=8), rtol=1e-6, factors=False)
    rx = [lrm.fixed_last_latent_features(inner.copy(), np.abs(X0[4:, i]).copy()) for i in range(50)]
    g = lrm.GLRM(np.abs(A), lrm.QuadLoss(), rx, lrm.QuadReg(0.1), k, obs=obs, X=np.abs(X0), Y=Y0.copy())
    check(orc, g, lrm.ProxGradParams(max_iter=8), rtol=1e-6, factors=False)


def test_payload_regularizers_objective_and_reg_scale(orc):
    """glrmb200_objective with payload regularizers (the penalty kernel walks mixed codes unit by unit) and
    scale_regularizer! reaching RemQuadReg / the inner regularizer of the fixed wrappers."""
    k = 5
    A, obs, X0 = small_sparse(m=40, n=25, k=k, seed=10)
    Y0 = synth.normal_matrix(74, 1, k, 25)
    rx = [lrm.RemQuadReg(0.3, X0[:, i] + 0.1) if i % 2 else lrm.QuadReg(0.2) for i in range(40)]
    ry = [lrm.fixed_latent_features(lrm.OneReg(0.1), np.abs(Y0[:2, j])) if j % 3 else lrm.NonNegOneReg(0.1) for j in range(25)]
    g = lrm.GLRM(A, lrm.QuadLoss(), rx, ry, k, obs=obs, X=X0.copy(), Y=np.abs(Y0))   # (Y starts on the pinned values: finite penalty)
    ep = lrm.encode_problem(g)
    with lrm.Engine(g) as eng:
        want = orc.objective(ep, g.X, g.Y, include_reg=True)
        got = eng.objective(g.X, g.Y, include_regularization=True)
        assert np.isfinite(want) and abs(got - want) <= 1e-12 * abs(want)
        eng.set_reg_scale(0.7)
        lrm.scale_regularizer(g, 0.7)
        ep2 = lrm.encode_problem(g)
        want2 = orc.objective(ep2, g.X, g.Y, include_reg=True)
        got2 = eng.objective(g.X, g.Y, include_regularization=True)
        assert abs(got2 - want2) <= 1e-12 * abs(want2) and abs(want2 - want) > 1e-6 * abs(want)
    check(orc, g, lrm.ProxGradParams(max_iter=6), rtol=1e-6, factors=False)
