"""Regularizer vocabulary — host-side mirror of /root/reference/src/regularizers.jl.

Classes only describe a regularizer; `encode()` yields the (code, 4-double parameter row) of the C
ABI.  evaluate/prox run in the CUDA engine (csrc/glrm_device.cuh).  A regularizer without a device
implementation raises `ArgumentError`-style ValueError at encode time — there is no CPU fallback.
"""
from __future__ import annotations

import copy as _copy
from dataclasses import dataclass

import numpy as np

REG_ZERO, REG_QUAD, REG_QUAD_CONSTRAINT, REG_ONE, REG_NONNEG, REG_NONNEG_ONE = 0, 1, 2, 3, 4, 5
REG_ONE_SPARSE, REG_KSPARSE, REG_UNIT_ONE_SPARSE, REG_SIMPLEX, REG_REM_QUAD = 6, 7, 8, 9, 10
REG_LASTENTRY1, REG_LASTENTRY_UNPENALIZED = 0x100, 0x200
REG_ORDINAL, REG_MNL_ORDINAL = 0x400, 0x800
REG_FIXED_FIRST, REG_FIXED_LAST = 0x1000, 0x2000
REG_NPARAM = 4


class Regularizer:  # abstract type Regularizer (regularizers.jl:30)
    code = -1

    def _p0(self):
        return 0.0

    def encode(self):
        if self.code < 0:
            raise ValueError(f"{type(self).__name__} has no B200 device implementation "
                             "(no CPU fallback is provided)")
        p = np.zeros(REG_NPARAM)
        p[0] = self._p0()
        return int(self.code), p

    def payload(self):
        """Vector payload travelling next to the descriptor row (fixed_latent_features.y, RemQuadReg.m), or None."""
        return None

    def copy(self):
        return _copy.deepcopy(self)

    # scale / mul! (regularizers.jl:37-38)
    def get_scale(self):
        return getattr(self, "scale", 1.0)

    def mul(self, newscale):
        if hasattr(self, "scale"):
            self.scale = float(newscale)
        return self


@dataclass
class QuadReg(Regularizer):  # regularizers.jl:52-58
    scale: float = 1.0
    code = REG_QUAD

    def _p0(self):
        return self.scale


@dataclass
class QuadConstraint(Regularizer):  # regularizers.jl:68-76
    max_2norm: float = 1.0
    code = REG_QUAD_CONSTRAINT

    def _p0(self):
        return self.max_2norm


@dataclass
class OneReg(Regularizer):  # regularizers.jl:79-88
    scale: float = 1.0
    code = REG_ONE

    def _p0(self):
        return self.scale


@dataclass
class ZeroReg(Regularizer):  # regularizers.jl:91-97
    code = REG_ZERO


@dataclass
class NonNegConstraint(Regularizer):  # regularizers.jl:101-114
    code = REG_NONNEG


@dataclass
class NonNegOneReg(Regularizer):  # regularizers.jl:118-138
    scale: float = 1.0
    code = REG_NONNEG_ONE

    def _p0(self):
        return self.scale

    def get_scale(self):        # scale(r::NonNegOneReg) = 1 (regularizers.jl:137)
        return 1.0

    def mul(self, newscale):    # mul!(r::NonNegOneReg, newscale) = 1: a no-op (regularizers.jl:138)
        return self


@dataclass
class OneSparseConstraint(Regularizer):  # regularizers.jl:235-255
    code = REG_ONE_SPARSE


@dataclass
class KSparseConstraint(Regularizer):  # regularizers.jl:258-291
    k: int = 1
    code = REG_KSPARSE

    def _p0(self):
        return float(self.k)


@dataclass
class UnitOneSparseConstraint(Regularizer):  # regularizers.jl:295-318
    code = REG_UNIT_ONE_SPARSE


@dataclass
class SimplexConstraint(Regularizer):  # regularizers.jl:323-348
    code = REG_SIMPLEX


class _Wrapper(Regularizer):
    flag = 0

    def __init__(self, r: Regularizer = None):
        self.r = ZeroReg() if r is None else r

    def encode(self):
        if isinstance(self.r, _Wrapper):
            raise ValueError("nested offset wrappers have no device implementation")
        code, p = self.r.encode()
        return code | self.flag, p

    def get_scale(self):
        return self.r.get_scale()

    def mul(self, newscale):
        self.r.mul(newscale)
        return self

    def __repr__(self):
        return f"{type(self).__name__}({self.r!r})"


class lastentry1(_Wrapper):  # regularizers.jl:163-174
    flag = REG_LASTENTRY1


class lastentry_unpenalized(_Wrapper):  # regularizers.jl:178-189
    flag = REG_LASTENTRY_UNPENALIZED


class _FixedWrapper(_Wrapper):
    """fixed_latent_features / fixed_last_latent_features (regularizers.jl:193-231): n = len(y) entries of the factor column
    are pinned to y, the inner regularizer r sees the rest."""

    def __init__(self, r, y=None):
        if y is None:                       # FixedLatentFeaturesConstraint(y): standalone use, inner = ZeroReg (:199,220)
            r, y = ZeroReg(), r
        super().__init__(r)
        self.y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        self.n = len(self.y)

    def encode(self):
        if isinstance(self.r, _Wrapper) or self.r.payload() is not None:
            raise ValueError(f"{type(self).__name__} around {type(self.r).__name__} has no device implementation")
        code, p = self.r.encode()
        return code | self.flag, p

    def payload(self):
        return self.y

    def __repr__(self):
        return f"{type(self).__name__}({self.r!r}, n={self.n})"


class fixed_latent_features(_FixedWrapper):  # regularizers.jl:193-210
    flag = REG_FIXED_FIRST


class fixed_last_latent_features(_FixedWrapper):  # regularizers.jl:214-231
    flag = REG_FIXED_LAST


def FixedLatentFeaturesConstraint(y):  # regularizers.jl:199
    return fixed_latent_features(ZeroReg(), y)


def FixedLastLatentFeaturesConstraint(y):  # regularizers.jl:220
    return fixed_last_latent_features(ZeroReg(), y)


class OrdinalReg(_Wrapper):  # regularizers.jl:356-380 (block regularizer of the ordinal losses; ry only)
    flag = REG_ORDINAL


class MNLOrdinalReg(_Wrapper):  # regularizers.jl:385-407
    flag = REG_MNL_ORDINAL


class RemQuadReg(Regularizer):  # regularizers.jl:412-423: quadratic regularization around a non-zero mean m
    code = REG_REM_QUAD

    def __init__(self, scale_or_m, m=None):
        if m is None:                       # RemQuadReg(m) = RemQuadReg(1, m) (:416)
            scale_or_m, m = 1.0, scale_or_m
        self.scale = float(scale_or_m)
        self.m = np.ascontiguousarray(m, dtype=np.float64).ravel()

    def _p0(self):
        return self.scale

    def payload(self):
        return self.m

    def __repr__(self):
        return f"RemQuadReg({self.scale}, len(m)={len(self.m)})"


def encode_regs(regs):
    """Encode a list of regularizers; collapses to a single shared row when all entries are the same
    object or encode identically (the usual `fillcopies` case, conveniencemethods.jl:29-49)."""
    cnt = len(regs)
    cache = {}
    rows = []
    for r in regs:
        key = id(r)
        if key not in cache:
            cache[key] = r.encode()
        rows.append(cache[key])
    codes = np.fromiter((c for c, _ in rows), dtype=np.int32, count=cnt)
    params = np.stack([p for _, p in rows]) if cnt else np.zeros((0, REG_NPARAM))
    has_payload = any(r.payload() is not None for r in regs)
    if cnt > 1 and not has_payload and np.all(codes == codes[0]) and np.all(params == params[0]):
        return codes[:1].copy(), params[:1].copy()
    return codes, params


def encode_payloads(regs):
    """-> (ptr int64[count+1], values float64[...]) or (None, None) when no regularizer carries a vector payload."""
    pays = [r.payload() for r in regs]
    if all(p is None for p in pays):
        return None, None
    lens = np.fromiter((0 if p is None else len(p) for p in pays), dtype=np.int64, count=len(pays))
    ptr = np.zeros(len(pays) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    vals = np.concatenate([p for p in pays if p is not None]) if ptr[-1] else np.zeros(0)
    return ptr, np.ascontiguousarray(vals, dtype=np.float64)
