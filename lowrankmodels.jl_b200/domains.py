"""Domains and the evaluation helpers that use them — host-side mirror of /root/reference/src/domains.jl and of the
callers in src/evaluate_fit.jl:106-168 / src/impute_and_err.jl (impute, impute_missing, error_metric).  The arithmetic runs
on the device (csrc/glrm_eval.cuh) through glrmb200_impute / glrmb200_error_metric; there is no CPU path here."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi, losses as L

DOMAIN_REAL, DOMAIN_BOOL, DOMAIN_ORDINAL, DOMAIN_CATEGORICAL, DOMAIN_PERIODIC, DOMAIN_COUNT = 1, 2, 3, 4, 5, 6


class Domain:  # domains.jl:21
    code = 0

    def params(self):
        return (0.0, 0.0)


@dataclass
class RealDomain(Domain):  # domains.jl:25
    code = DOMAIN_REAL


@dataclass
class BoolDomain(Domain):  # domains.jl:30
    code = DOMAIN_BOOL


@dataclass
class OrdinalDomain(Domain):  # domains.jl:35-45
    min: int = 1
    max: int = 10
    code = DOMAIN_ORDINAL

    def __post_init__(self):
        if self.max - self.min < 2:
            import warnings
            warnings.warn("The ordinal variable you've created is degenerate: it has only two levels. "
                          "Consider using a Boolean variable instead!")

    def params(self):
        return (float(self.min), float(self.max))


@dataclass
class CategoricalDomain(Domain):  # domains.jl:48-52
    min: int = 1
    max: int = 1
    code = DOMAIN_CATEGORICAL

    def params(self):
        return (float(self.min), float(self.max))


@dataclass
class PeriodicDomain(Domain):  # domains.jl:56-58
    T: float = 1.0
    code = DOMAIN_PERIODIC

    def params(self):
        return (float(self.T), 0.0)


@dataclass
class CountDomain(Domain):  # domains.jl:62-64
    max_count: int = 2**31
    code = DOMAIN_COUNT

    def params(self):
        return (float(self.max_count), 0.0)


def loss_domain(l):
    """l.domain as the reference's constructors set it (losses.jl:142,156,171,191,214,235,253,302,322,365,419,456,495,567);
    an explicit `domain` attribute on the loss object wins."""
    d = getattr(l, "domain", None)
    if d is not None:
        return d
    if isinstance(l, L.PeriodicLoss):
        return PeriodicDomain(l.T)
    if isinstance(l, L.DiffLoss):
        return RealDomain()
    if isinstance(l, L.PoissonLoss):
        return CountDomain(l.max_count)
    if isinstance(l, L.OrdinalHingeLoss):
        return OrdinalDomain(l.min, l.max)
    if isinstance(l, L.ClassificationLoss):
        return BoolDomain()
    if isinstance(l, (L.MultinomialLoss, L.OvALoss)):
        return CategoricalDomain(1, l.max)
    if isinstance(l, (L.BvSLoss, L.OrdisticLoss, L.MultinomialOrdinalLoss)):
        return OrdinalDomain(1, l.max)
    raise TypeError(f"no domain for {type(l).__name__}")


def _encode_domains(glrm, domains):
    doms = [loss_domain(l) for l in glrm.losses] if domains is None else list(domains)
    if len(doms) != len(glrm.losses):
        raise ValueError("one domain per column")
    code = np.array([d.code for d in doms], dtype=np.int32)
    par = np.array([d.params() for d in doms], dtype=np.float64).reshape(-1)
    return code, par


def _engine_for(glrm, engine):
    from .fit import Engine
    return (engine, False) if engine is not None else (Engine(glrm), True)


def impute(glrm, domains=None, engine=None):
    """impute(glrm) = impute(glrm.losses, glrm.X'*glrm.Y) (evaluate_fit.jl:150; impute_and_err.jl:147-168): the m x n table of
    a_u = argmin_a loss(u, a) over each column's domain."""
    eng, own = _engine_for(glrm, engine)
    try:
        code, par = _encode_domains(glrm, domains)
        out = np.zeros((eng.m, eng.n), order="F")
        X, Y = np.asfortranarray(glrm.X, dtype=np.float64), np.asfortranarray(glrm.Y, dtype=np.float64)
        _abi.check(_abi.lib().glrmb200_impute(eng.h, _abi.dptr(X), _abi.dptr(Y), _abi.i32ptr(code), _abi.dptr(par), _abi.dptr(out)))
        return out
    finally:
        if own:
            eng.close()


def impute_missing(glrm, domains=None, engine=None):
    """impute_missing(glrm) (evaluate_fit.jl:151-159): the imputed table with the observed entries of A put back."""
    Ahat = impute(glrm, domains, engine)
    A = glrm.A.toarray() if hasattr(glrm.A, "toarray") else np.asarray(glrm.A, dtype=np.float64)
    for j in range(glrm.shape[1]):
        idx = np.asarray(glrm.observed_examples[j], dtype=np.int64)
        Ahat[idx, j] = A[idx, j]
    return Ahat


def error_metric(glrm, domains=None, standardize=False, engine=None):
    """error_metric(glrm, domains; standardize) (evaluate_fit.jl:106-148) over the observed entries."""
    eng, own = _engine_for(glrm, engine)
    try:
        code, par = _encode_domains(glrm, domains)
        out = C.c_double(0)
        X, Y = np.asfortranarray(glrm.X, dtype=np.float64), np.asfortranarray(glrm.Y, dtype=np.float64)
        _abi.check(_abi.lib().glrmb200_error_metric(eng.h, _abi.dptr(X), _abi.dptr(Y), _abi.i32ptr(code), _abi.dptr(par),
                                                    int(bool(standardize)), C.byref(out)))
        return out.value
    finally:
        if own:
            eng.close()
