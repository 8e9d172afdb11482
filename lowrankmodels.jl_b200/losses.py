"""Loss vocabulary — host-side mirror of /root/reference/src/losses.jl (names, constructor
arguments and defaults follow the reference; the arithmetic itself lives in the CUDA engine,
csrc/glrm_device.cuh, and — for tests only — in oracle/).

Each class only *describes* a loss: `encode()` produces the (code, 8-double parameter row) the C ABI
takes (include/glrm_b200.h).  There is no Python/CPU evaluation path in the product.
"""
from __future__ import annotations

import copy as _copy
from dataclasses import dataclass, field

import numpy as np

# codes: include/glrm_b200.h
LOSS_QUAD, LOSS_L1, LOSS_HUBER, LOSS_QUANTILE, LOSS_PERIODIC, LOSS_POISSON = 1, 2, 3, 4, 5, 6
LOSS_ORDINAL_HINGE, LOSS_LOGISTIC, LOSS_WEIGHTED_HINGE = 7, 8, 9
LOSS_MULTINOMIAL, LOSS_OVA, LOSS_BVS, LOSS_ORDISTIC, LOSS_MULTINOMIAL_ORDINAL = 10, 11, 12, 13, 14
LOSS_NPARAM = 8


class Loss:
    """abstract type Loss (losses.jl:52).  Fields: `scale` (losses.jl:6-7)."""

    code = 0
    scale: float

    def embedding_dim(self) -> int:  # losses.jl:72
        return 1

    def _params(self):
        return ()

    def encode(self):
        p = np.zeros(LOSS_NPARAM)
        p[0] = self.scale
        for i, v in self._params():
            p[i] = v
        return self.code, p

    # label domain used by the encoder to reproduce the reference's dispatch-time errors
    # ("real" | "bool" | "level")
    label_kind = "real"

    def copy(self):
        return _copy.deepcopy(self)

    def __mul__(self, newscale):  # *(l::Loss, newscale) losses.jl:63-64 (sets, does not multiply)
        new = self.copy()
        new.scale = float(newscale)
        return new

    __rmul__ = __mul__


class DiffLoss(Loss):  # losses.jl:55
    pass


class ClassificationLoss(Loss):  # losses.jl:57
    label_kind = "bool"


@dataclass
class QuadLoss(DiffLoss):  # losses.jl:138-146
    scale: float = 1.0
    code = LOSS_QUAD


@dataclass
class L1Loss(DiffLoss):  # losses.jl:152-160
    scale: float = 1.0
    code = LOSS_L1


@dataclass
class HuberLoss(DiffLoss):  # losses.jl:166-177
    scale: float = 1.0
    crossover: float = 1.0
    code = LOSS_HUBER

    def _params(self):
        return ((1, self.crossover),)


@dataclass
class QuantileLoss(DiffLoss):  # losses.jl:186-201
    scale: float = 1.0
    quantile: float = 0.5
    code = LOSS_QUANTILE

    def _params(self):
        return ((1, self.quantile),)


@dataclass
class PeriodicLoss(DiffLoss):  # losses.jl:209-218  (T is the first positional argument)
    T: float = 1.0
    scale: float = 1.0
    code = LOSS_PERIODIC

    def _params(self):
        return ((1, self.T),)


@dataclass
class PoissonLoss(Loss):  # losses.jl:231-241 (constructor fixes scale = 1.0, :235)
    max_count: int = 2**31
    scale: float = 1.0
    code = LOSS_POISSON


@dataclass
class OrdinalHingeLoss(Loss):  # losses.jl:247-292
    min: int = 1
    max: int = 10
    scale: float = 1.0
    code = LOSS_ORDINAL_HINGE

    def _params(self):
        return ((1, self.min), (2, self.max))


@dataclass
class LogisticLoss(ClassificationLoss):  # losses.jl:298-306
    scale: float = 1.0
    code = LOSS_LOGISTIC


@dataclass
class WeightedHingeLoss(ClassificationLoss):  # losses.jl:317-341
    scale: float = 1.0
    case_weight_ratio: float = 1.0
    code = LOSS_WEIGHTED_HINGE

    def _params(self):
        return ((1, self.case_weight_ratio),)


def HingeLoss(scale: float = 1.0, **kw):  # losses.jl:324
    return WeightedHingeLoss(scale, **kw)


class _LevelLoss(Loss):
    label_kind = "level"
    max: int

    def _bin(self):
        return None

    def _params(self):
        out = [(2, self.max)]
        b = self._bin()
        if b is not None:
            if b.embedding_dim() != 1 or isinstance(b, _LevelLoss):
                raise ValueError("bin_loss must be a scalar loss")
            bcode, bp = b.encode()
            out += [(3, bcode), (4, bp[0]), (5, bp[1]), (6, bp[2])]
        return tuple(out)


@dataclass
class MultinomialLoss(_LevelLoss):  # losses.jl:360-398
    max: int = 2
    scale: float = 1.0
    code = LOSS_MULTINOMIAL

    def embedding_dim(self):  # :366
        return int(self.max)


@dataclass
class OvALoss(_LevelLoss):  # losses.jl:413-438; default bin_loss = LogisticLoss(scale) (:419)
    max: int = 1
    scale: float = 1.0
    bin_loss: Loss = None
    code = LOSS_OVA

    def __post_init__(self):
        if self.bin_loss is None:
            self.bin_loss = LogisticLoss(self.scale)

    def _bin(self):
        return self.bin_loss

    def embedding_dim(self):  # :421
        return int(self.max)


@dataclass
class BvSLoss(_LevelLoss):  # losses.jl:450-475
    max: int = 10
    scale: float = 1.0
    bin_loss: Loss = None
    code = LOSS_BVS

    def __post_init__(self):
        if self.bin_loss is None:
            self.bin_loss = LogisticLoss(self.scale)

    def _bin(self):
        return self.bin_loss

    def embedding_dim(self):  # :458
        return int(self.max) - 1


@dataclass
class OrdisticLoss(_LevelLoss):  # losses.jl:490-519
    max: int = 2
    scale: float = 1.0
    code = LOSS_ORDISTIC

    def embedding_dim(self):  # :496
        return int(self.max)


@dataclass
class MultinomialOrdinalLoss(_LevelLoss):  # losses.jl:562-608
    max: int = 10
    scale: float = 1.0
    code = LOSS_MULTINOMIAL_ORDINAL

    def embedding_dim(self):  # :569
        return int(self.max) - 1


def embedding_dim(losses) -> int:
    """embedding_dim(l::Array{Loss}) = sum(map(embedding_dim, l))  (losses.jl:73)"""
    if isinstance(losses, Loss):
        return losses.embedding_dim()
    return int(sum(l.embedding_dim() for l in losses))


def get_yidxs(losses):
    """get_yidxs (losses.jl:76-93): for each column of A the (0-based, half-open) span of columns of
    Y it owns.  Returns an int64 array `ystart` of length n+1; column f owns ystart[f]:ystart[f+1]."""
    ds = np.fromiter((l.embedding_dim() for l in losses), dtype=np.int64, count=len(losses))
    ystart = np.zeros(len(losses) + 1, dtype=np.int64)
    np.cumsum(ds, out=ystart[1:])
    return ystart


def encode_losses(losses):
    n = len(losses)
    codes = np.zeros(n, dtype=np.int32)
    params = np.zeros((n, LOSS_NPARAM))
    cache = {}
    for f, l in enumerate(losses):
        key = id(l)
        if key not in cache:
            cache[key] = l.encode()
        codes[f], params[f] = cache[key]
    return codes, params
