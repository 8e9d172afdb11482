"""End-to-end restatement checks (CPU only): the C oracle's two forms against each other and against
the independent line-by-line Python restatement of proxgrad.jl, on small seeded problems covering the
reference's test shapes (test/basic_functionality.jl, test/hello_world.jl, examples/simple_glrms.jl)."""
import numpy as np
import pytest

import lowrankmodels_b200 as lrm
import proxgrad_ref as ref
from helpers import assert_traj_close, glrm_from_config, run_oracle, small_sparse
from lowrankmodels_b200 import synth


def check_all_forms(orc, glrm, params, rtol=1e-9):
    sparse = run_oracle(orc, glrm, params, mode=1, nthreads=1)
    faithful = run_oracle(orc, glrm, params, mode=0, nthreads=2)
    X, Y, ch, ar, ac = ref.fit_reference(glrm, params)
    assert_traj_close(sparse["objective"], ch, rtol, "C sparse vs python")
    assert_traj_close(faithful["objective"], ch, rtol, "C faithful vs python")
    np.testing.assert_allclose(sparse["alpharow"], ar, rtol=1e-12)
    np.testing.assert_allclose(sparse["alphacol"], ac, rtol=1e-12)
    np.testing.assert_allclose(sparse["X"], X, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(faithful["Y"], Y, rtol=1e-6, atol=1e-9)
    return sparse


def test_config1_dense_quad_quadreg(orc):
    """BASELINE config 1: dense 100x100 QuadLoss + QuadReg(0.1) k=5, ProxGradParams() defaults."""
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    res = check_all_forms(orc, g, lrm.ProxGradParams(max_iter=25))
    obj = res["objective"]
    assert obj[1] < obj[0] and (np.diff(obj[1:]) <= 1e-9).all()      # monotone after the first record
    # quirk Q1: entry 0 is the full objective, later entries omit rx
    ep = lrm.encode_problem(g)
    assert obj[0] == pytest.approx(orc.objective(ep, g.X, g.Y, True), rel=1e-12)


def test_basic_functionality_shape_zero_reg(orc):
    """test/basic_functionality.jl:5-16: exact rank-5, QuadLoss, ZeroReg, Params(1, max_iter=200,
    abs_tol=1e-7, min_stepsize=1e-3): ch.objective[end] == ||A - X'Y||^2 (self-consistency)."""
    c = synth.config1(seed=5)
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg())
    p = lrm.Params(1, max_iter=60, abs_tol=1e-7, min_stepsize=1e-3)
    res = run_oracle(orc, g, p, mode=1)
    Ah = res["X"].T @ res["Y"]
    assert abs(np.linalg.norm(c["A"] - Ah) ** 2 - res["objective"][-1]) < 1e-6 * res["objective"][-1] + 1e-7


@pytest.mark.parametrize("dup", [False, True])
def test_sparse_obs_with_order_and_duplicates(orc, dup):
    A, obs, X0 = small_sparse(dup=dup)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.05), lrm.QuadReg(0.05), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 1, 4, A.shape[1]))
    check_all_forms(orc, g, lrm.ProxGradParams(max_iter=12))


def test_nnmf_infeasible_start_first_objective_is_inf(orc):
    """examples/simple_glrms.jl fit_nnmf: NonNegConstraint with randn start -> ch[0] = Inf (quirk Q4)."""
    A, obs, X0 = small_sparse(seed=4)
    g = lrm.GLRM(np.abs(A), lrm.QuadLoss(), lrm.NonNegConstraint(), lrm.NonNegConstraint(), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 2, 4, A.shape[1]))
    res = check_all_forms(orc, g, lrm.ProxGradParams(max_iter=10))
    assert np.isinf(res["objective"][0]) and np.isfinite(res["objective"][1:]).all()


def test_logistic_nonneg_config3_shape(orc):
    A, obs, X0 = small_sparse(seed=6, labels="bool")
    g = lrm.GLRM(A, lrm.LogisticLoss(), lrm.NonNegConstraint(), lrm.NonNegConstraint(), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 3, 4, A.shape[1]))
    check_all_forms(orc, g, lrm.ProxGradParams(max_iter=10))


def test_kmeans_unit_one_sparse(orc):
    c = synth.config5(scale=50000, k=6, n=8, centroids=4)
    g = glrm_from_config(c, lrm.QuadLoss(), lrm.UnitOneSparseConstraint(), lrm.ZeroReg())
    check_all_forms(orc, g, lrm.ProxGradParams(max_iter=8))


def test_heterogeneous_losses_hello_world_shape(orc):
    """test/hello_world.jl:5-45 in spirit: mixed real / boolean / ordinal / categorical columns with
    different scales, mixed row regularizers, duplicated obs."""
    m, n, k = 30, 9, 3
    A = np.zeros((m, n))
    u = synth.uniform(2, 61, np.arange(m * n)).reshape(m, n)
    z = synth.normal_matrix(2, 62, m, n)
    A[:, 0:2] = z[:, 0:2]
    A[:, 2] = np.where(z[:, 2] > 0, 1, -1)
    A[:, 3] = np.floor(u[:, 3] * 5) + 1          # ordinal 1..5 (OrdinalHinge)
    A[:, 4] = np.floor(u[:, 4] * 4) + 1          # BvS levels 1..4
    A[:, 5] = np.floor(u[:, 5] * 3) + 1          # Multinomial 1..3
    A[:, 6] = np.floor(u[:, 6] * 3) + 1          # OvA 1..3
    A[:, 7] = np.floor(u[:, 7] * 4)              # Poisson counts
    A[:, 8] = z[:, 8]
    losses = [lrm.QuadLoss(1.5), lrm.HuberLoss(0.7), lrm.HingeLoss(1.2), lrm.OrdinalHingeLoss(1, 5, 0.9),
              lrm.BvSLoss(4, 1.1), lrm.MultinomialLoss(3, 0.8), lrm.OvALoss(3, 1.3), lrm.PoissonLoss(),
              lrm.QuantileLoss(1.0, quantile=0.7)]
    d = lrm.embedding_dim(losses)
    rx = [lrm.QuadReg(0.1) if e % 4 == 0 else lrm.OneReg(0.05) if e % 4 == 1 else lrm.NonNegConstraint()
          if e % 4 == 2 else lrm.KSparseConstraint(2) for e in range(m)]
    ry = lrm.QuadReg(0.1)
    ii, jj = np.nonzero(u < 0.7)
    obs = np.stack([ii, jj], axis=1)
    obs = np.concatenate([obs, obs[:25]])
    g = lrm.GLRM(A, losses, rx, ry, k, obs=obs, X=np.abs(synth.normal_matrix(2, 63, k, m)) * 0.3,
                 Y=synth.normal_matrix(2, 64, k, d) * 0.3)
    check_all_forms(orc, g, lrm.ProxGradParams(max_iter=8), rtol=1e-8)


def test_offset_wrappers_and_inner_iters(orc):
    A, obs, X0 = small_sparse(seed=8)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 4, 4, A.shape[1]), offset=True)
    assert type(g.rx[0]).__name__ == "lastentry1" and type(g.ry[0]).__name__ == "lastentry_unpenalized"
    check_all_forms(orc, g, lrm.ProxGradParams(max_iter=6, inner_iter=2))


def ordinal_problem(m=35, k=4, seed=5):
    """The DataFrame front-end's default for ordinal columns (fit_dataframe.jl:13-18,64-72): MultinomialOrdinalLoss with
    MNLOrdinalReg on the column block, offset on (rows lastentry1), next to real-valued columns; plus BvS + OrdinalReg."""
    n = 5
    u = synth.uniform(seed, 71, np.arange(m * n)).reshape(m, n)
    z = synth.normal_matrix(seed, 72, m, n)
    A = z.copy()
    A[:, 1] = np.floor(u[:, 1] * 5) + 1          # MultinomialOrdinal levels 1..5 (embedding 4)
    A[:, 2] = np.floor(u[:, 2] * 2) + 1          # MultinomialOrdinal levels 1..2 (embedding 1, still a block column)
    A[:, 3] = np.floor(u[:, 3] * 4) + 1          # BvS levels 1..4 (embedding 3)
    losses = [lrm.QuadLoss(), lrm.MultinomialOrdinalLoss(5), lrm.MultinomialOrdinalLoss(2), lrm.BvSLoss(4), lrm.HuberLoss()]
    ry = [lrm.QuadReg(0.1), lrm.MNLOrdinalReg(lrm.QuadReg(0.1)), lrm.MNLOrdinalReg(lrm.QuadReg(0.2)),
          lrm.OrdinalReg(lrm.OneReg(0.05)), lrm.QuadReg(0.1)]
    d = lrm.embedding_dim(losses)
    ii, jj = np.nonzero(u < 0.8)
    X0 = 0.3 * synth.normal_matrix(seed, 73, k, m)
    X0[-1, :] = 1.0
    Y0 = 0.3 * synth.normal_matrix(seed, 74, k, d)
    Y0[-1, :] = -0.3 - 0.2 * np.arange(d)        # decreasing negative last row: feasible for the ordinal rules
    return lrm.GLRM(A, losses, lrm.QuadReg(0.1), ry, k, obs=np.stack([ii, jj], axis=1), X=X0, Y=Y0, offset=True)


def test_ordinal_block_regularizers(orc):
    g = ordinal_problem()
    assert type(g.rx[0]).__name__ == "lastentry1" and type(g.ry[1]).__name__ == "MNLOrdinalReg"   # not re-wrapped (:409)
    assert type(g.ry[0]).__name__ == "lastentry_unpenalized"
    check_all_forms(orc, g, lrm.ProxGradParams(max_iter=8), rtol=1e-8)


def test_stopping_rule_fires_after_ten(orc):
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    res = run_oracle(orc, g, lrm.ProxGradParams(max_iter=100, rel_tol=1e-2))
    assert 12 <= len(res["objective"]) < 101          # i>10 && decrease/obj < rel_tol (proxgrad.jl:211)


def test_unit_sample_fit_matches_both_forms_and_freezes_the_rest(orc):
    """oracle_fit_units (bench.py's bounded CPU step): faithful and sparse-evaluated forms agree on the sampled units,
    every other column of X and Y is untouched, and the full ranges reproduce oracle_fit."""
    A, obs, X0 = small_sparse(seed=6)
    m, n = A.shape
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.05), lrm.QuadReg(0.05), 4, obs=obs, X=X0,
                 Y=synth.normal_matrix(9, 3, 4, n))
    ep = lrm.encode_problem(g)
    p = lrm.encode_params(lrm.ProxGradParams(max_iter=6, abs_tol=0, rel_tol=0))
    rows, cols = (m // 4, m // 2), (n // 3, n // 3 + max(2, n // 4))
    out = {}
    for mode in (0, 1):
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        res = orc.fit_units(ep, p, X, Y, rows, cols, mode=mode, nthreads=2)
        out[mode] = (X, Y, res)
        frozen_r = np.r_[0:rows[0], rows[1]:m]
        frozen_c = np.r_[0:cols[0], cols[1]:n]
        assert (X[:, frozen_r] == g.X[:, frozen_r]).all() and (Y[:, frozen_c] == g.Y[:, frozen_c]).all()
        assert not (X[:, rows[0]:rows[1]] == g.X[:, rows[0]:rows[1]]).all()
    assert_traj_close(out[0][2]["objective"][1:], out[1][2]["objective"][1:], 1e-9, "unit sample: faithful vs sparse")
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-7, atol=1e-10)
    assert out[0][2]["trials"] == out[1][2]["trials"]
    X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
    full = orc.fit_units(ep, p, X, Y, (0, m), (0, n), mode=0)
    want = run_oracle(orc, g, lrm.ProxGradParams(max_iter=6, abs_tol=0, rel_tol=0), mode=0)
    assert (full["objective"] == want["objective"]).all() and (X == want["X"]).all()


def test_rem_quad_reg_recovers_the_means(orc):
    """test/mult_reg.jl:5-29: QuadLoss with RemQuadReg(50, mean) per row / column pulls the factors to their means; the
    reference asserts mse(U), mse(V) < 1e-3 after fit!(glrm) with default parameters."""
    n = m = 200
    r, eta, delta = 5, 0.01, 1e-3
    Um, Vm = synth.normal_matrix(71, 1, r, n), synth.normal_matrix(71, 2, r, m)
    U = Um + np.sqrt(eta) * synth.normal_matrix(71, 3, r, n)
    V = Vm + np.sqrt(eta) * synth.normal_matrix(71, 4, r, m)
    Yobs = U.T @ V + np.sqrt(delta) * synth.normal_matrix(71, 5, n, m)
    g = lrm.GLRM(np.asfortranarray(Yobs), lrm.QuadLoss(), [lrm.RemQuadReg(50, Um[:, i]) for i in range(n)],
                 [lrm.RemQuadReg(50, Vm[:, j]) for j in range(m)], r,
                 X=synth.normal_matrix(71, 6, r, n), Y=synth.normal_matrix(71, 7, r, m))
    res = check_all_forms(orc, g, lrm.ProxGradParams(max_iter=100))
    assert np.mean((U - res["X"]) ** 2) < 1e-3 and np.mean((V - res["Y"]) ** 2) < 1e-3


@pytest.mark.parametrize("inner", [lrm.ZeroReg(), lrm.QuadReg(0.1), lrm.NonNegConstraint(), lrm.OneSparseConstraint()],
                         ids=lambda r: type(r).__name__)
def test_fixed_latent_features_both_sides(orc, inner):
    """test/fixedfeatures_test.jl shape: ry = fixed_latent_features(...) per column; plus the 'last' variant on the rows,
    restated with the reference's own indexing (regularizers.jl:223 feeds u[n+1:end] to the inner prox)."""
    A, obs, X0 = small_sparse(m=40, n=30, k=5, seed=8)
    k = 5
    Y0 = synth.normal_matrix(72, 1, k, 30)
    ry = [lrm.fixed_latent_features(inner.copy(), Y0[:2, j].copy()) for j in range(30)]
    rx = [lrm.fixed_last_latent_features(inner.copy(), X0[3:, i].copy()) for i in range(40)]
    g = lrm.GLRM(A, lrm.QuadLoss(), rx, ry, k, obs=obs, X=X0.copy(), Y=Y0.copy())
    res = check_all_forms(orc, g, lrm.ProxGradParams(max_iter=10))
    assert (res["Y"][:2, :] == Y0[:2, :]).all() and (res["X"][3:, :] == X0[3:, :]).all()     # pinned entries never move
    assert np.isfinite(res["objective"][1:]).all()
