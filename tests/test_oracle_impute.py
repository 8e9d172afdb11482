"""CPU: the restatement of src/impute_and_err.jl (oracle/impute_ref.py) against values read off the reference's definitions and
against the reference's own consistency test (test/err_test.jl:35,48-51: data made by `impute` has error_metric exactly 0)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import impute_ref
import lowrankmodels_b200 as lrm
from lowrankmodels_b200 import synth


def test_rules_read_off_the_reference():
    R, B, O = lrm.RealDomain(), lrm.BoolDomain(), lrm.OrdinalDomain(1, 5)
    assert impute_ref.impute(R, lrm.QuadLoss(), 0.37) == 0.37                       # impute_and_err.jl:40
    assert impute_ref.impute(R, lrm.PoissonLoss(), 1.5) == math.exp(1.5)            # :41
    assert impute_ref.impute(R, lrm.OrdinalHingeLoss(1, 4), 7.2) == 4.0             # :42 roundcutoff(u, l.min, l.max)
    assert impute_ref.impute(R, lrm.OrdinalHingeLoss(1, 4), 2.5) == 2.0             # round half to even, as Julia's round
    assert impute_ref.impute(B, lrm.LogisticLoss(), 0.0) == 1.0 and impute_ref.impute(B, lrm.LogisticLoss(), -1e-9) == 0.0   # :60
    assert impute_ref.impute(B, lrm.QuadLoss(), 0.49) == 0.0 and impute_ref.impute(B, lrm.QuadLoss(), 0.51) == 1.0           # :63
    assert impute_ref.impute(O, lrm.QuadLoss(), 9.0) == 5.0 and impute_ref.impute(O, lrm.L1Loss(), -3.0) == 1.0               # :75
    assert impute_ref.impute(O, lrm.LogisticLoss(), 0.1) == 5.0 and impute_ref.impute(O, lrm.LogisticLoss(), -0.1) == 1.0     # :78
    assert impute_ref.impute(lrm.CountDomain(3), lrm.PoissonLoss(3), 2.0) == 3.0    # :117 -> :76 roundcutoff(exp(u), 0, 3)
    assert impute_ref.impute(lrm.CategoricalDomain(1, 3), lrm.MultinomialLoss(3), np.array([0.1, 0.7, 0.7])) == 2.0           # :101 first max
    assert impute_ref.impute(lrm.OrdinalDomain(1, 4), lrm.OrdisticLoss(4), np.array([2.0, -0.5, 0.6, 3.0])) == 2.0            # :84 argmin(u.^2)
    assert impute_ref.error_metric_entry(B, lrm.LogisticLoss(), 2.0, 0.0) == 1.0    # :64-67 misclassification
    assert impute_ref.error_metric_entry(R, lrm.QuadLoss(), 2.0, 0.5) == 2.25       # :49-52 squared error
    P = lrm.PeriodicDomain(3.0)
    assert abs(impute_ref.error_metric_entry(P, lrm.PeriodicLoss(3.0), -0.5, 2.5)) < 1e-30   # :113-116 pos_mod(3,-0.5) == 2.5


def test_imputed_data_has_zero_error_metric():
    """test/err_test.jl:35,48-51 with the losses of its list (one column each) and their own domains."""
    losses = [lrm.QuadLoss(), lrm.L1Loss(), lrm.HuberLoss(), lrm.PeriodicLoss(1.0), lrm.OrdinalHingeLoss(1, 10), lrm.LogisticLoss(),
              lrm.WeightedHingeLoss()]
    m, n, k = 40, len(losses), 4
    X, Y = synth.normal_matrix(3, 1, k, m), synth.normal_matrix(3, 2, k, n)
    doms = [lrm.loss_domain(l) for l in losses]
    ys = list(range(n + 1))
    g0 = lrm.GLRM(np.zeros((m, n)), losses, lrm.ZeroReg(), lrm.ZeroReg(), k, X=X, Y=Y, checknan=False)
    A = impute_ref.impute_table(g0, doms, ys)
    g = lrm.GLRM(A, losses, lrm.ZeroReg(), lrm.ZeroReg(), k, X=X, Y=Y)
    assert impute_ref.error_metric(g, doms, ys, standardize=True) == 0.0
    assert impute_ref.error_metric(g, doms, ys, standardize=False) == 0.0
