"""GLRM container and observation bookkeeping — mirror of /root/reference/src/glrm.jl:9-80,
src/modify_glrm.jl:5-24 and src/utilities/conveniencemethods.jl:29-49.

Indices held by this Python mirror are 0-based (the Julia shim converts Julia's 1-based lists at the
boundary); list order and duplicates are preserved exactly as `sort_observations` does.
"""
from __future__ import annotations

import numpy as np

from .losses import Loss, embedding_dim, get_yidxs
from .regularizers import Regularizer, lastentry1, lastentry_unpenalized

try:  # optional: scipy.sparse plays the role of SparseMatrixCSC
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None


class ObsLists:
    """A ragged `Vector{Vector{Int}}` stored flat: list i is idx[ptr[i]:ptr[i+1]] (0-based).
    `full` marks the `fill(1:n, m)` UnitRange default (glrm.jl:33-34) without materialising it."""

    def __init__(self, ptr=None, idx=None, full=None):
        self.full = full  # (count, length) when every list is 0..length-1
        self.ptr = ptr
        self.idx = idx

    @classmethod
    def from_lists(cls, lists):
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        ptr = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        idx = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if ptr[-1] > 0 \
            else np.zeros(0, dtype=np.int64)
        return cls(ptr, idx)

    def __len__(self):
        return self.full[0] if self.full is not None else len(self.ptr) - 1

    def __getitem__(self, i):
        if self.full is not None:
            return np.arange(self.full[1], dtype=np.int64)
        return self.idx[self.ptr[i]:self.ptr[i + 1]]

    def total(self):
        return self.full[0] * self.full[1] if self.full is not None else int(self.ptr[-1])


def sort_observations(obs, m, n, check_empty=False, return_perm=False):
    """sort_observations (modify_glrm.jl:5-18): unpack [(i,j), ...] into the two adjacency lists by
    `push!` in input order — order and duplicates preserved.  `obs` is an (nobs, 2) integer array or
    a list of pairs, 0-based.  Implemented as a stable counting sort (same result as the push! loop)."""
    obs = np.asarray(obs, dtype=np.int64).reshape(-1, 2)
    i, j = obs[:, 0], obs[:, 1]
    if obs.size and (i.min() < 0 or i.max() >= m or j.min() < 0 or j.max() >= n):
        raise IndexError("observation index out of bounds")  # Julia: BoundsError in push!
    order_r = np.argsort(i, kind="stable")
    order_c = np.argsort(j, kind="stable")
    rptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(i, minlength=m), out=rptr[1:])
    cptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(j, minlength=n), out=cptr[1:])
    feats = ObsLists(rptr, j[order_r].copy())
    exs = ObsLists(cptr, i[order_c].copy())
    if check_empty and ((np.diff(rptr) == 0).any() or (np.diff(cptr) == 0).any()):
        raise ValueError("Every row and column must contain at least one observation")
    if return_perm:
        return feats, exs, (order_r, order_c)
    return feats, exs


class Repeated:
    """`fillcopies(x, count)` (conveniencemethods.jl:29-49) without materialising `count` Python
    objects: a read-only sequence whose every element is the same descriptor.  (The reference makes
    independent copies; sharing is equivalent here because descriptors are only ever rescaled through
    `mul`, which sets an absolute scale.)"""

    def __init__(self, item, count):
        self.item = item
        self.count = int(count)

    def __len__(self):
        return self.count

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.item] * len(range(*i.indices(self.count)))
        if not -self.count <= i < self.count:
            raise IndexError(i)
        return self.item

    def __iter__(self):
        for _ in range(self.count):
            yield self.item


def _fillcopies(x, count, base):
    """conveniencemethods.jl:29-49: a single loss / regularizer is expanded to `count` copies."""
    if isinstance(x, base):
        return Repeated(x.copy(), count)
    if isinstance(x, Repeated):
        return x
    return list(x)


class GLRM:
    """GLRM(A, losses, rx, ry, k; X, Y, obs, observed_features, observed_examples, offset, scale,
    checknan, sparse_na) — glrm.jl:28-80.

    A: (m, n) array (float; labels as numbers) or scipy.sparse matrix (the SparseMatrixCSC case:
    with sparse_na the nonzeros are the observations in CSC order, glrm.jl:46-48).
    X is (k, m), Y is (k, d) Float64 — column-major ("F") like Julia, mutated in place by fit_inplace.
    """

    def __init__(self, A, losses, rx, ry, k, *, X=None, Y=None, obs=None, observed_features=None,
                 observed_examples=None, offset=False, scale=False, checknan=True, sparse_na=True,
                 rng=None):
        self._sparse = _sp is not None and _sp.issparse(A)
        if self._sparse:
            A = A.tocsc()
            A.sort_indices()
        else:
            A = np.asarray(A)
            if A.ndim != 2:
                raise ValueError("A must be a matrix")
        m, n = A.shape
        self.A = A
        self._obs_vals = None   # values aligned with `obs` when derived from a sparse A
        self._row_val = self._col_val = None
        self.k = int(k)
        self.losses = _fillcopies(losses, n, Loss)
        self.rx = _fillcopies(rx, m, Regularizer)
        self.ry = _fillcopies(ry, n, Regularizer)
        # glrm.jl:38-43
        if len(self.losses) != n:
            raise ValueError("There must be as many losses as there are columns in the data matrix")
        if len(self.rx) != m:
            raise ValueError("There must be either one X regularizer or as many X regularizers as there "
                             "are rows in the data matrix")
        if len(self.ry) != n:
            raise ValueError("There must be either one Y regularizer or as many Y regularizers as there "
                             "are columns in the data matrix")
        d = embedding_dim(self.losses)
        rng = np.random.default_rng() if rng is None else rng
        if X is None:
            X = rng.standard_normal((self.k, m))   # randn(k, m)   glrm.jl:31
        if Y is None:
            Y = rng.standard_normal((self.k, d))   # randn(k, embedding_dim(losses))
        X = np.asarray(X, dtype=np.float64)
        if X.shape != (self.k, m) and X.shape == (m, self.k):
            X = X.T                                                            # glrm.jl:57-61
        if X.shape != (self.k, m):
            raise ValueError("X must be of size (k,m) where m is the number of rows in the data matrix.")
        Y = np.asarray(Y, dtype=np.float64)
        if Y.shape != (self.k, d):
            raise ValueError("Y must be of size (k,d) where d is the sum of the embedding dimensions of "
                             "all the losses.")
        self.X = np.asfortranarray(X)
        self.Y = np.asfortranarray(Y)

        if obs is None and sparse_na and self._sparse:                         # glrm.jl:46-48
            coo_j = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr))
            nz = A.data != 0
            obs = np.stack([A.indices[nz].astype(np.int64), coo_j[nz]], axis=1)
            self._obs_vals = np.asarray(A.data[nz], dtype=np.float64)
        if obs is None:                                                        # glrm.jl:50-52
            self.observed_features = (ObsLists(full=(m, n)) if observed_features is None
                                      else self._as_obs(observed_features))
            self.observed_examples = (ObsLists(full=(n, m)) if observed_examples is None
                                      else self._as_obs(observed_examples))
        else:                                                                  # glrm.jl:53-55
            self.observed_features, self.observed_examples, perm = sort_observations(
                obs, m, n, return_perm=True)
            if self._obs_vals is not None:
                self._row_val = self._obs_vals[perm[0]]
                self._col_val = self._obs_vals[perm[1]]
                self._obs_vals = None

        if checknan:                                                           # glrm.jl:63-71
            self._check_nan()
        if scale:
            raise NotImplementedError("scale=true (equilibrate_variance!, modify_glrm.jl:34-58) is "
                                      "model rewriting outside the accelerated path")
        if offset:                                                             # glrm.jl:76-78
            add_offset(self)

    @staticmethod
    def _as_obs(x):
        return x if isinstance(x, ObsLists) else ObsLists.from_lists(x)

    @property
    def shape(self):
        return self.A.shape

    def values_at(self, rows, cols):
        """A[rows[t], cols[t]] as Float64 (labels are numbers)."""
        if self._sparse:
            return np.asarray(self.A[rows, cols], dtype=np.float64).ravel()
        return np.asarray(self.A[rows, cols], dtype=np.float64)

    def _check_nan(self):
        feats = self.observed_features
        if feats.full is not None:
            vals = self.A.toarray() if self._sparse else self.A
            bad = np.argwhere(np.isnan(np.asarray(vals, dtype=np.float64)))
            if len(bad):
                raise ValueError(f"Observed value in entry ({bad[0][0] + 1}, {bad[0][1] + 1}) is NaN.")
            return
        rows = np.repeat(np.arange(len(feats), dtype=np.int64), np.diff(feats.ptr))
        vals = self.values_at(rows, feats.idx)
        bad = np.flatnonzero(np.isnan(vals))
        if len(bad):
            raise ValueError(f"Observed value in entry ({rows[bad[0]] + 1}, {feats.idx[bad[0]] + 1}) is NaN.")


def add_offset(glrm: GLRM):
    """add_offset! (modify_glrm.jl:21-24)."""
    def wrap(regs, w):
        # lastentry_unpenalized(r::OrdinalReg) = r, same for MNLOrdinalReg (regularizers.jl:383,409)
        keep = lambda r: r if type(r).__name__ in ("OrdinalReg", "MNLOrdinalReg") and w is lastentry_unpenalized else w(r)
        if isinstance(regs, Repeated):
            return Repeated(keep(regs.item), regs.count)
        return [keep(r) for r in regs]
    glrm.rx, glrm.ry = wrap(glrm.rx, lastentry1), wrap(glrm.ry, lastentry_unpenalized)
    return glrm


def scale_regularizer(glrm: GLRM, newscale: float):
    """scale_regularizer! (glrm.jl:84-88)."""
    for regs in (glrm.rx, glrm.ry):
        for r in ([regs.item] if isinstance(regs, Repeated) else regs):
            r.mul(newscale)
    return glrm
