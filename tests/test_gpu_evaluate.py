"""-m gpu: glrmb200_impute / glrmb200_error_metric (csrc/glrm_eval.cuh) against the CPU restatement of
src/impute_and_err.jl and src/evaluate_fit.jl:106-159 (oracle/impute_ref.py), through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import lowrankmodels_b200 as lrm
from lowrankmodels_b200 import synth

pytestmark = pytest.mark.gpu


def _ystart(g):
    ys = [0]
    for l in g.losses:
        ys.append(ys[-1] + l.embedding_dim())
    return ys


def mixed_problem(m=60, seed=5, sparse=False):
    """One column per (loss, natural domain) pair the reference defines, values drawn inside each domain."""
    k = 4
    losses = [lrm.QuadLoss(), lrm.L1Loss(2.0), lrm.HuberLoss(), lrm.QuantileLoss(), lrm.PeriodicLoss(3.0), lrm.PoissonLoss(20),
              lrm.OrdinalHingeLoss(1, 6), lrm.LogisticLoss(), lrm.WeightedHingeLoss(), lrm.MultinomialLoss(4), lrm.OvALoss(3),
              lrm.BvSLoss(4), lrm.OrdisticLoss(5), lrm.MultinomialOrdinalLoss(4)]
    n = len(losses)
    u = lambda j: synth.uniform(seed, 10 + j, np.arange(m))
    A = np.zeros((m, n))
    for j, l in enumerate(losses):
        name = type(l).__name__
        if name in ("QuadLoss", "L1Loss", "HuberLoss", "QuantileLoss"):
            A[:, j] = 4 * u(j) - 2
        elif name == "PeriodicLoss":
            A[:, j] = 3 * u(j)
        elif name == "PoissonLoss":
            A[:, j] = np.floor(6 * u(j))
        elif name == "OrdinalHingeLoss":
            A[:, j] = 1 + np.floor(6 * u(j))
        elif name in ("LogisticLoss", "WeightedHingeLoss"):
            A[:, j] = (u(j) > 0.5).astype(float)
        else:
            A[:, j] = 1 + np.floor(l.max * u(j)).clip(0, l.max - 1)
    d = sum(l.embedding_dim() for l in losses)
    X = 0.8 * synth.normal_matrix(seed, 1, k, m)
    Y = 0.8 * synth.normal_matrix(seed, 2, k, d)
    kw = {}
    if sparse:
        kw["obs"] = [(i, j) for i in range(m) for j in range(n) if (7 * i + 3 * j) % 5 != 0]
    return lrm.GLRM(A, losses, lrm.QuadReg(0.1), lrm.QuadReg(0.1), k, X=X, Y=Y, **kw)


@pytest.mark.parametrize("sparse", [False, True])
def test_impute_and_error_metric_natural_domains(sparse):
    import impute_ref
    g = mixed_problem(sparse=sparse)
    doms = [lrm.loss_domain(l) for l in g.losses]
    ys = _ystart(g)
    want = impute_ref.impute_table(g, doms, ys)
    got = lrm.impute(g)
    assert got.shape == want.shape
    fin = np.isfinite(want)
    assert (np.isfinite(got) == fin).all()
    # integer-valued domains must match exactly; real-valued imputations to rounding
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-12, atol=1e-12)
    for std in (False, True):
        e_want = impute_ref.error_metric(g, doms, ys, standardize=std)
        e_got = lrm.error_metric(g, standardize=std)
        assert abs(e_got - e_want) <= 1e-10 * max(1.0, abs(e_want)), (std, e_got, e_want)
    miss = lrm.impute_missing(g)
    A = np.asarray(g.A, dtype=float)
    for j in range(g.shape[1]):
        idx = np.asarray(g.observed_examples[j])
        assert (miss[idx, j] == A[idx, j]).all()


def test_domain_overrides_bool_and_ordinal_on_quadloss():
    """PCA on a binary / ordinal table scored over the data's own domain (the use case of impute_and_err.jl:7-13)."""
    import impute_ref
    m, n, k = 50, 6, 3
    A = np.zeros((m, n))
    A[:, :3] = (synth.uniform(9, 1, np.arange(3 * m)).reshape(m, 3) > 0.4).astype(float)
    A[:, 3:] = 1 + np.floor(5 * synth.uniform(9, 2, np.arange(3 * m)).reshape(m, 3))
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), k, X=synth.normal_matrix(9, 3, k, m), Y=synth.normal_matrix(9, 4, k, n))
    doms = [lrm.BoolDomain()] * 3 + [lrm.OrdinalDomain(1, 5)] * 2 + [lrm.CountDomain(7)]
    ys = _ystart(g)
    np.testing.assert_allclose(lrm.impute(g, doms), impute_ref.impute_table(g, doms, ys), rtol=0, atol=0)
    for std in (False, True):
        assert abs(lrm.error_metric(g, doms, standardize=std) - impute_ref.error_metric(g, doms, ys, standardize=std)) < 1e-9


def test_error_metric_on_a_fitted_fully_observed_handle():
    """The tensor-core handle of a fully observed problem scores its own fit without a second upload of A."""
    import impute_ref
    c = synth.config5(scale=20000)           # 500 x 128, k = 100
    g = lrm.GLRM(c["A"], lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 12, X=c["X0"][:12].copy(), Y=c["Y0"][:12].copy())
    with lrm.Engine(g) as eng:
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        eng.fit(lrm.ProxGradParams(max_iter=5), X, Y)
        g.X[...] = X
        g.Y[...] = Y
        got = lrm.error_metric(g, engine=eng)
    want = impute_ref.error_metric(g, [lrm.loss_domain(l) for l in g.losses], _ystart(g))
    assert abs(got - want) <= 1e-10 * abs(want)


def test_imputed_data_has_zero_error_metric_on_the_device():
    """The reference's own consistency test (test/err_test.jl:35,48-51) through the C ABI: a table made by glrmb200_impute
    scores exactly 0 under glrmb200_error_metric with the same factors and domains."""
    losses = [lrm.QuadLoss(), lrm.L1Loss(), lrm.HuberLoss(), lrm.PeriodicLoss(1.0), lrm.OrdinalHingeLoss(1, 10), lrm.LogisticLoss(),
              lrm.WeightedHingeLoss(), lrm.MultinomialLoss(4), lrm.OrdisticLoss(5)]
    m, n, k = 100, len(losses), 5
    d = sum(l.embedding_dim() for l in losses)
    X, Y = synth.normal_matrix(4, 1, k, m), synth.normal_matrix(4, 2, k, d)
    seed_A = np.ones((m, n))                       # any table inside every label domain: only its shape matters for impute
    g0 = lrm.GLRM(seed_A, losses, lrm.ZeroReg(), lrm.ZeroReg(), k, X=X, Y=Y)
    A = lrm.impute(g0)
    assert np.isfinite(A).all()
    g = lrm.GLRM(A, losses, lrm.ZeroReg(), lrm.ZeroReg(), k, X=X, Y=Y)
    assert lrm.error_metric(g, standardize=True) == 0.0
    assert lrm.error_metric(g, standardize=False) == 0.0
