"""The oracle (C restatement) and the line-by-line Python restatement against the known answers the
reference's own tests hold (tests/golden/reference_known_answers.json)."""
import json
import math
import os

import numpy as np
import pytest

import lowrankmodels_b200 as lrm
import proxgrad_ref as ref

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


def mk_loss(name, args):
    return getattr(lrm, name)(**args)


def mk_vec(x):
    if isinstance(x, str):
        n = int(x[x.index("(") + 1:x.index(")")])
        return np.ones(n)
    return np.array(x, dtype=float)


def val(v):
    if v == "Inf":
        return math.inf
    if isinstance(v, str):  # "ones(100)/sqrt(100)*7"
        return np.ones(100) / math.sqrt(100) * 7
    return v


@pytest.mark.parametrize("case", G["loss_evaluate"], ids=lambda c: c["src"])
def test_loss_evaluate(case, orc):
    l = mk_loss(case["loss"], case["args"])
    tol = max(case["tol"], 1e-15)
    assert abs(orc.loss_eval(l, case["u"], case["a"]) - case["value"]) <= tol
    assert abs(ref.evaluate(l, case["u"], case["a"]) - case["value"]) <= tol


@pytest.mark.parametrize("case", G["loss_grad"], ids=lambda c: c["src"])
def test_loss_grad(case, orc):
    l = mk_loss(case["loss"], case["args"])
    assert abs(orc.loss_grad(l, float(case["u"]), case["a"]) - case["value"]) <= case["tol"]
    assert abs(ref.grad(l, case["u"], case["a"]) - case["value"]) <= case["tol"]


@pytest.mark.parametrize("case", G["reg_evaluate"], ids=lambda c: c["src"])
def test_reg_evaluate(case, orc):
    r = getattr(lrm, case["reg"])(*case["args"])
    x = mk_vec(case["x"])
    assert orc.reg_eval(r, x) == val(case["value"])
    assert ref.reg_evaluate(r, x) == val(case["value"])


@pytest.mark.parametrize("case", G["reg_prox"], ids=lambda c: c["src"])
def test_reg_prox(case, orc):
    r = getattr(lrm, case["reg"])(*case["args"])
    x = mk_vec(case["x"])
    want = np.asarray(val(case["value"]), dtype=float)
    np.testing.assert_allclose(orc.reg_prox(r, x, case["alpha"]), want, rtol=0, atol=case["tol"])
    np.testing.assert_allclose(ref.prox(r, x, case["alpha"]), want, rtol=0, atol=case["tol"])


def test_bad_bool_label_is_an_error(orc):
    # myBool(2) throws InexactError (losses.jl:104)
    with pytest.raises(ValueError):
        orc.loss_eval(lrm.LogisticLoss(), 0.3, 2)
    with pytest.raises(ValueError):
        ref.evaluate(lrm.LogisticLoss(), 0.3, 2)


ALL_SCALAR = [lrm.QuadLoss(1.3), lrm.L1Loss(0.7), lrm.HuberLoss(1.1, crossover=0.8), lrm.QuantileLoss(0.9, quantile=0.3),
              lrm.PeriodicLoss(3.0, 1.2), lrm.PoissonLoss(), lrm.OrdinalHingeLoss(1, 6, 0.5)]
ALL_BOOL = [lrm.LogisticLoss(1.7), lrm.WeightedHingeLoss(0.6, case_weight_ratio=2.5), lrm.HingeLoss()]
ALL_VEC = [lrm.MultinomialLoss(4, 1.2), lrm.OvALoss(4, 0.8), lrm.BvSLoss(5, 1.1), lrm.OrdisticLoss(4, 0.9),
           lrm.MultinomialOrdinalLoss(5, 1.3), lrm.OvALoss(3, 1.0, bin_loss=lrm.HingeLoss(2.0))]


def test_c_oracle_matches_python_restatement_on_every_loss(orc):
    rng = np.random.default_rng(7)
    for l in ALL_SCALAR:
        for _ in range(40):
            u = float(rng.normal() * 3)
            a = float(rng.integers(1, 7)) if l.code in (6, 7) else float(rng.normal())
            assert orc.loss_eval(l, u, a) == pytest.approx(ref.evaluate(l, u, a), rel=1e-13, abs=1e-15)
            assert orc.loss_grad(l, u, a) == pytest.approx(ref.grad(l, u, a), rel=1e-13, abs=1e-15)
    for l in ALL_BOOL:
        for a in (1, 0, -1):
            for _ in range(20):
                u = float(rng.normal() * 3)
                assert orc.loss_eval(l, u, a) == pytest.approx(ref.evaluate(l, u, a), rel=1e-13, abs=1e-15)
                assert orc.loss_grad(l, u, a) == pytest.approx(ref.grad(l, u, a), rel=1e-13, abs=1e-15)
    for l in ALL_VEC:
        D = l.embedding_dim()
        for _ in range(40):
            u = rng.normal(size=D) * 2
            a = int(rng.integers(1, l.max + 1))
            assert orc.loss_eval(l, u, a) == pytest.approx(ref.evaluate(l, u, a), rel=1e-12, abs=1e-14)
            np.testing.assert_allclose(orc.loss_grad(l, u, a), ref.grad(l, u, a), rtol=1e-12, atol=1e-14)


ALL_REGS = [lrm.ZeroReg(), lrm.QuadReg(0.3), lrm.QuadConstraint(1.5), lrm.OneReg(0.4), lrm.NonNegConstraint(),
            lrm.NonNegOneReg(0.2), lrm.OneSparseConstraint(), lrm.KSparseConstraint(2), lrm.UnitOneSparseConstraint(),
            lrm.SimplexConstraint(), lrm.lastentry1(lrm.QuadReg(0.3)), lrm.lastentry_unpenalized(lrm.OneReg(0.2)),
            lrm.lastentry1(lrm.NonNegConstraint()), lrm.lastentry_unpenalized(lrm.QuadReg(0.5)),
            lrm.OrdinalReg(lrm.QuadReg(0.3)), lrm.MNLOrdinalReg(lrm.QuadReg(0.2)), lrm.OrdinalReg(lrm.OneReg(0.1)),
            lrm.MNLOrdinalReg(lrm.ZeroReg())]


def test_c_oracle_matches_python_restatement_on_every_regularizer(orc):
    rng = np.random.default_rng(11)
    for r in ALL_REGS:
        for shape in ((6,), (5, 3)):
            for _ in range(10):
                v = rng.normal(size=shape)
                alpha = float(rng.uniform(0.01, 0.5))
                p_c = orc.reg_prox(r, v, alpha)
                p_py = ref.prox(r, v, alpha)
                np.testing.assert_allclose(p_c, p_py, rtol=1e-13, atol=1e-15, err_msg=repr(r))
                for w in (v, p_py):
                    e_c, e_py = orc.reg_eval(r, w), ref.reg_evaluate(r, w)
                    assert (e_c == e_py) or e_c == pytest.approx(e_py, rel=1e-13), repr(r)
