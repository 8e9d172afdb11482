"""SparseProxGradParams (src/algorithms/sparse_proxgrad.jl:21-130 — what plain fit!(glrm) selects for a sparse A):
the C restatement against the line-by-line Python one, including rejected iterations (objective went up ->
alpha / max(1.5, -steps_in_a_row), revert) and the recorded series (initial, accepted only, final duplicate)."""
import numpy as np
import pytest

import lowrankmodels_b200 as lrm
import proxgrad_ref as ref
from helpers import small_sparse
from lowrankmodels_b200 import synth


def problems():
    A, obs, X0 = small_sparse(seed=40, dup=True)
    yield "quad", lrm.GLRM(A, lrm.QuadLoss(), lrm.QuadReg(0.05), lrm.QuadReg(0.05), 4, obs=obs, X=X0,
                           Y=synth.normal_matrix(41, 1, 4, A.shape[1]))
    Ab, obsb, Xb = small_sparse(seed=42, labels="bool")
    yield "logistic-nonneg", lrm.GLRM(Ab, lrm.LogisticLoss(), lrm.NonNegConstraint(), lrm.OneReg(0.02), 4, obs=obsb,
                                      X=np.abs(Xb), Y=np.abs(synth.normal_matrix(43, 1, 4, Ab.shape[1])))
    c = synth.config1(seed=7)
    yield "dense-huber", lrm.GLRM(c["A"], lrm.HuberLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1), 5, X=c["X0"], Y=c["Y0"])


@pytest.mark.parametrize("stepsize", [1.0, 12.0])     # 12: the first steps overshoot -> rejections and reverts
def test_c_oracle_matches_python_restatement(orc, stepsize):
    for name, g in problems():
        p = lrm.SparseProxGradParams(stepsize, max_iter=14, abs_tol=1e-6)
        ep = lrm.encode_problem(g)
        X, Y = g.X.copy(order="F"), g.Y.copy(order="F")
        r = orc.fit_sparse(ep, lrm.encode_sparse_params(p), X, Y)
        Xb, Yb, ch, alpha = ref.fit_sparse_reference(g, p)
        assert len(r["objective"]) == len(ch), name
        np.testing.assert_allclose(r["objective"], ch, rtol=1e-10, err_msg=name)
        assert r["alpha"] == pytest.approx(alpha, rel=1e-13)
        np.testing.assert_allclose(X, Xb, rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(Y, Yb, rtol=1e-7, atol=1e-10)
        assert r["objective"][-1] == r["objective"][-2]                 # the final duplicate (sparse_proxgrad.jl:126)
        assert (np.diff(r["objective"][:-1]) < 0).all()                 # only accepted iterations are recorded
        if stepsize > 5:
            assert len(ch) < p.max_iter + 2                            # some iterations were rejected


def test_default_dispatch_mirrors_fit_jl():
    """fit!(glrm) picks SparseProxGradParams for a SparseMatrixCSC, ProxGradParams otherwise (src/fit.jl:13-19)."""
    import scipy.sparse as sp
    seen = []

    class FakeEngine:
        def fit(self, p, X, Y):
            seen.append(type(p).__name__); return np.zeros(1), np.zeros(1)
        def fit_sparse(self, p, X, Y):
            seen.append(type(p).__name__); return np.zeros(1), np.zeros(1)
        def close(self):
            pass

    dense = lrm.GLRM(np.ones((4, 3)), lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2)
    sparse = lrm.GLRM(sp.csc_matrix(np.eye(4, 3)), lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2)
    lrm.fit_inplace(dense, verbose=False, engine=FakeEngine())
    lrm.fit_inplace(sparse, verbose=False, engine=FakeEngine())
    assert seen == ["ProxGradParams", "SparseProxGradParams"]
