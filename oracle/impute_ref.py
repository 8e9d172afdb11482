"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's imputation and error metrics:
/root/reference/src/impute_and_err.jl:31-168 (impute / error_metric per (domain, loss)) and
/root/reference/src/evaluate_fit.jl:106-159 (raw / standardised error_metric over observed_examples, impute_missing).
Plain Python loops over small cases; loss values come from oracle/proxgrad_ref.py's `evaluate` (losses.jl).
Imported by tests/ only; the product never loads it."""
import math

import numpy as np

import proxgrad_ref as R


def roundcutoff(x, a, b):  # impute_and_err.jl:31  (Julia round: ties to even == Python round)
    return float(min(max(round(x), a), b))


def squared_error(a_imputed, a):  # :34
    return (a_imputed - a) ** 2


def misclassification(a_imputed, a):  # :35
    return float(not (a_imputed == a))


def pos_mod(T, x):  # :112
    return math.fmod(x, T) if x > 0 else math.fmod(x, T) + T


_DIFF = ("QuadLoss", "L1Loss", "HuberLoss", "QuantileLoss", "PeriodicLoss")
_CLASSIF = ("LogisticLoss", "WeightedHingeLoss", "HingeLoss")


def _argmax(v):  # Julia argmax: first maximal element, NaN wins
    best = 0
    for j in range(1, len(v)):
        if (v[j] != v[j] and v[best] == v[best]) or (v[best] == v[best] and v[j] > v[best]):
            best = j
    return best


def impute(D, l, u):
    """impute(D::Domain, l::Loss, u) (impute_and_err.jl:40-120); u is a float or a vector (embedding_dim > 1)."""
    dn, ln = type(D).__name__, type(l).__name__
    if dn == "CountDomain":                                                        # :117
        class _O:  # OrdinalDomain(0, D.max_count)
            pass
        O = _O(); O.min, O.max = 0, D.max_count
        O.__class__.__name__ = "OrdinalDomain"
        return impute(O, l, u)
    if dn == "PeriodicDomain":                                                     # :109
        dn = "RealDomain"
    if np.ndim(u) == 0:
        u = float(u)
        if dn == "RealDomain":
            if ln in _DIFF: return u                                              # :40
            if ln == "PoissonLoss": return math.exp(u)                            # :41
            if ln == "OrdinalHingeLoss": return roundcutoff(u, l.min, l.max)      # :42
            if ln in ("WeightedHingeLoss", "HingeLoss"): return 1.0 / u           # :44-47
            return float("nan")                                                   # :43 error(...)
        if dn == "BoolDomain":
            if ln in _CLASSIF: return 1.0 if u >= 0 else 0.0                      # :60
            return 0.0 if R.evaluate(l, u, False) < R.evaluate(l, u, True) else 1.0   # :63
        if dn == "OrdinalDomain":
            if ln in _DIFF: return roundcutoff(u, D.min, D.max)                   # :75
            if ln == "PoissonLoss": return roundcutoff(math.exp(u), D.min, D.max)  # :76
            if ln == "OrdinalHingeLoss": return roundcutoff(u, D.min, D.max)      # :77
            if ln == "LogisticLoss": return float(D.max if u > 0 else D.min)      # :78
            return roundcutoff(math.ceil(1 / u) if u > 0 else math.floor(1 / u), D.min, D.max)   # :79-83
        return float("nan")
    u = [float(x) for x in u]
    if dn == "CategoricalDomain" and ln in ("MultinomialLoss", "OvALoss"):        # :101-102
        return float(_argmax(u) + 1)
    if dn == "OrdinalDomain":
        if ln == "OrdisticLoss":                                                  # :84
            return float(_argmax([-(x * x) for x in u]) + 1)
        if ln == "MultinomialOrdinalLoss":                                        # :85-90
            u = R.enforce_MNLOrdRules(list(u))
            eu = [math.exp(x) for x in u]
            p = [1 - eu[0]] + [eu[j - 1] - eu[j] for j in range(1, len(eu))] + [eu[-1]]
            return float(_argmax(p) + 1)
        vals = [R.evaluate(l, np.array(u), i) for i in range(D.min, D.max + 1)]   # :91-93
        return float(range(D.min, D.max + 1)[_argmax([-v for v in vals])])
    return float("nan")


def error_metric_entry(D, l, u, a):
    a_imp = impute(D, l, u)
    dn = type(D).__name__
    if dn in ("BoolDomain", "CategoricalDomain"):                                  # :64-67, :103-106
        return misclassification(a_imp, float(a))
    if dn == "PeriodicDomain":                                                     # :113-116
        return squared_error(pos_mod(D.T, a_imp), pos_mod(D.T, float(a)))
    return squared_error(a_imp, float(a))                                          # :49-52, :94-97, :118-121


def impute_table(glrm, domains, ystart):
    """impute(domains, losses, X'Y) (impute_and_err.jl:147-162)"""
    U = glrm.X.T @ glrm.Y
    m, n = glrm.shape
    out = np.zeros((m, n))
    for f in range(n):
        cols = range(ystart[f], ystart[f + 1])
        for i in range(m):
            u = U[i, cols[0]] if len(cols) == 1 else U[i, list(cols)]
            out[i, f] = impute(domains[f], glrm.losses[f], u)
    return out


def error_metric(glrm, domains, ystart, standardize=False):
    """raw_error_metric / std_error_metric (evaluate_fit.jl:106-137)"""
    U = glrm.X.T @ glrm.Y
    A = glrm.A.toarray() if hasattr(glrm.A, "toarray") else np.asarray(glrm.A, dtype=float)
    err = 0.0
    for j in range(glrm.shape[1]):
        cols = list(range(ystart[j], ystart[j + 1]))
        column_mean = column_err = 0.0
        obs = glrm.observed_examples[j]
        for i in obs:
            u = U[i, cols[0]] if len(cols) == 1 else U[i, cols]
            column_mean += A[i, j] ** 2
            column_err += error_metric_entry(domains[j], glrm.losses[j], u, A[i, j])
        if standardize:
            column_mean = column_mean / len(obs)
            if column_mean != 0:
                column_err = column_err / column_mean
        err += column_err
    return err
