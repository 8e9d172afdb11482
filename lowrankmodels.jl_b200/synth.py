"""Deterministic synthetic inputs for the five BASELINE.json configs.

There is no network for datasets, and the reference's tests draw from Julia's RNG (not reproducible
outside Julia), so inputs come from a counter-based generator: SplitMix64 hash of
(seed, stream, index) -> U(0,1) -> Box-Muller.  Pure integer hashing, so Python / C++ / CUDA can
reproduce the same bytes from the same counters.

Deviation from SURVEY section 8d, recorded here and in DESIGN.md: column popularity is Zipf-Mandelbrot
p(rank) ~ 1/(rank + 40) rather than pure Zipf(1.0).  Pure Zipf(1.0) over 26 744 columns puts 9 % of
20 M observations on the top column, i.e. more than m = 138 493 distinct rows; the shifted law gives a
top column of ~0.38 % of the observations (MovieLens-20M's most rated film: 67 310 = 0.34 %).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SYNTH_LIB = None


def _synth_lib():
    """csrc/libglrm_synth.so (built by __graft_entry__.build()): distinct-pair sampling in C."""
    global _SYNTH_LIB
    if _SYNTH_LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libglrm_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
        _SYNTH_LIB = C.CDLL(path)
        _SYNTH_LIB.glrm_synth_pattern.restype = C.c_int
        for fn in ("glrm_synth_c4", "glrm_synth_c5", "glrm_synth_normal_fill"):
            getattr(_SYNTH_LIB, fn).restype = None
    return _SYNTH_LIB


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """One SplitMix64 output step on uint64 array `x` (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _key(seed, stream):
    return splitmix64(np.array([np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream)],
                               dtype=np.uint64))[0]


def uniform(seed, stream, idx):
    """U(0,1) for each counter in `idx` (any integer array); never exactly 0."""
    idx = np.asarray(idx).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64(idx ^ _key(seed, stream))
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(seed, stream, idx):
    """N(0,1) via Box-Muller on counters 2*idx, 2*idx+1."""
    idx = np.asarray(idx).astype(np.uint64)
    u1 = uniform(seed, stream, idx * np.uint64(2))
    u2 = uniform(seed, stream, idx * np.uint64(2) + np.uint64(1))
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def normal_matrix(seed, stream, rows, cols):
    """(rows, cols) Fortran-ordered N(0,1): element (r, c) uses counter c*rows + r.  Large matrices are filled by the C
    helper (same counters, OpenMP; libm's log / cos instead of numpy's, i.e. equal up to the last ulp)."""
    if rows * cols >= 4_000_000:
        out = np.empty((rows, cols), order="F")
        _synth_lib().glrm_synth_normal_fill(C.c_uint64(int(_key(seed, stream))), C.c_int64(0), C.c_int64(rows * cols),
                                            out.ctypes.data_as(C.POINTER(C.c_double)))
        return out
    return normal(seed, stream, np.arange(rows * cols, dtype=np.uint64)).reshape((rows, cols), order="F")


# ---- sparsity pattern: MovieLens-20M shaped --------------------------------------------------------
ML20M = dict(m=138_493, n=26_744, nnz=20_000_263, dmin=20, dmax=9_254)


def _row_degrees(m, n, nnz, lo, hi, seed):
    """lognormal(mu=ln 68, sigma=1.22) row degrees, clipped to [lo, hi], rescaled to sum to nnz."""
    z = normal(seed, 11, np.arange(m))
    raw = np.exp(1.22 * z)
    hi = min(hi, n)
    scale = nnz / raw.sum()
    for _ in range(60):                       # fixed-point on the clip
        deg = np.clip(raw * scale, lo, hi)
        s = deg.sum()
        if abs(s - nnz) < 0.5:
            break
        free = (deg > lo) & (deg < hi)
        if not free.any():
            break
        scale *= (nnz - (s - deg[free].sum())) / max(deg[free].sum(), 1e-300)
    deg = np.clip(np.floor(raw * scale), lo, hi).astype(np.int64)
    resid = int(nnz - deg.sum())
    order = np.argsort(-raw, kind="stable")
    i = 0
    while resid != 0 and i < 50 * m:          # hand out the rounding residual, heaviest rows first
        e = order[i % m]
        if resid > 0 and deg[e] < hi:
            deg[e] += 1
            resid -= 1
        elif resid < 0 and deg[e] > lo:
            deg[e] -= 1
            resid += 1
        i += 1
    assert deg.sum() == nnz, (deg.sum(), nnz)
    return deg


def sparse_pattern(m, n, nnz, lo, hi, seed=1, shift=40.0):
    """Distinct (row, col) pairs: row degrees lognormal, columns drawn by Zipf-Mandelbrot popularity
    (random column relabelling so popular columns are scattered).  Returned in CSC order (column-major,
    rows ascending) — what `findall(!iszero, A)` yields for a SparseMatrixCSC (glrm.jl:46-48)."""
    deg = np.ascontiguousarray(_row_degrees(m, n, nnz, lo, hi, seed), dtype=np.int64)
    w = 1.0 / (np.arange(1, n + 1) + shift * n / ML20M["n"])
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    relabel = np.argsort(uniform(seed, 12, np.arange(n)), kind="stable")
    rows = np.empty(nnz, dtype=np.int64)
    cols = np.empty(nnz, dtype=np.int64)
    L = _synth_lib()
    rc = L.glrm_synth_pattern(m, n, _p64(deg), cdf.ctypes.data_as(C.POINTER(C.c_double)),
                              _p64(np.ascontiguousarray(relabel, dtype=np.int64)),
                              C.c_uint64(int(_key(seed, 13))), _p64(rows), _p64(cols))
    if rc != 0:
        raise RuntimeError("pattern generator did not converge")
    return rows, cols


def _lowrank_at(rows, cols, m, n, r, seed, chunk=2_000_000):
    """(P Q)[rows, cols] / sqrt(r) for P (m x r), Q (r x n) ~ N(0,1), evaluated entry-wise in chunks."""
    P = normal_matrix(seed, 21, m, r)
    Q = normal_matrix(seed, 22, n, r)
    out = np.empty(rows.size)
    for s in range(0, rows.size, chunk):
        e = slice(s, s + chunk)
        out[e] = np.einsum("ij,ij->i", P[rows[e]], Q[cols[e]])
    return out / np.sqrt(r)


def _scaled_ml20m(scale):
    m = ML20M["m"] // scale
    n = ML20M["n"] // scale
    nnz = ML20M["nnz"] // (scale * scale)
    lo = max(1, round(ML20M["dmin"] / scale))
    hi = max(lo + 1, round(ML20M["dmax"] / scale))
    return m, n, nnz, lo, hi


def config1(seed=1):
    """C1: dense 100x100 QuadLoss + QuadReg(0.1), k=5 (examples/simple_glrms.jl:31-40)."""
    m = n = 100
    k = 5
    A = normal_matrix(seed, 1, m, k) @ normal_matrix(seed, 2, k, n)
    return dict(name="C1", m=m, n=n, k=k, A=np.asfortranarray(A), full=True,
                X0=normal_matrix(seed, 3, k, m), Y0=normal_matrix(seed, 4, k, n))


def config2(scale=1, seed=1, k=50):
    """C2: MovieLens-20M-shaped sparse ratings, QuadLoss + QuadReg(0.1), k=50.  scale=8 is the CI twin."""
    m, n, nnz, lo, hi = _scaled_ml20m(scale)
    rows, cols = sparse_pattern(m, n, nnz, lo, hi, seed)
    key = rows * n + cols
    base = _lowrank_at(rows, cols, m, n, 10, seed)
    vals = np.clip(np.round(2.0 * (base + 3.5 + 0.5 * normal(seed, 23, key))) / 2.0, 0.5, 5.0)
    return dict(name=f"C2/{scale}", m=m, n=n, k=k, rows=rows, cols=cols, vals=vals, full=False,
                X0=normal_matrix(seed, 3, k, m), Y0=normal_matrix(seed, 4, k, n))


def config3(scale=1, seed=1, k=50):
    """C3: C2's pattern with labels sign(p.q + eps) in {-1,+1}; LogisticLoss + NonNegConstraint."""
    m, n, nnz, lo, hi = _scaled_ml20m(scale)
    rows, cols = sparse_pattern(m, n, nnz, lo, hi, seed)
    key = rows * n + cols
    base = _lowrank_at(rows, cols, m, n, 10, seed) * np.sqrt(10)
    vals = np.where(base + normal(seed, 23, key) >= 0, 1.0, -1.0)
    return dict(name=f"C3/{scale}", m=m, n=n, k=k, rows=rows, cols=cols, vals=vals, full=False,
                X0=normal_matrix(seed, 3, k, m), Y0=normal_matrix(seed, 4, k, n))


def config4(scale=1, seed=1, k=20, levels=5, rows=None):
    """C4: heterogeneous columns, fully observed: 50 % QuadLoss, 30 % HingeLoss, 20 % MultinomialLoss(5).
    `rows`: number of rows instead of 1e6 / scale (same generator, same columns) — the parity twin of the full-size run."""
    m = rows if rows is not None else 1_000_000 // scale
    n = 1_000 // scale
    nq, nh = n // 2, (n * 3) // 10
    nm = n - nq - nh
    P = normal_matrix(seed, 31, m, 4)
    Q = normal_matrix(seed, 32, 4, n)
    A = np.empty((m, n), order="F")
    dp = C.POINTER(C.c_double)
    _synth_lib().glrm_synth_c4(C.c_int64(m), C.c_int64(n), C.c_int64(nq), C.c_int64(nh), C.c_int64(levels),
                               np.ascontiguousarray(P.ravel(order="F")).ctypes.data_as(dp),
                               np.ascontiguousarray(Q.ravel(order="F")).ctypes.data_as(dp),
                               C.c_uint64(int(_key(seed, 33))), C.c_uint64(int(_key(seed, 34))), A.ctypes.data_as(dp))
    d = nq + nh + nm * levels
    return dict(name=f"C4/{scale}", m=m, n=n, k=k, A=A, full=True, n_quad=nq, n_hinge=nh, n_multi=nm,
                levels=levels, d=d, X0=normal_matrix(seed, 3, k, m), Y0=normal_matrix(seed, 4, k, d))


def config5(scale=1, seed=1, k=100, n=128, centroids=100, rows=None):
    """C5: k-means path, A_i = c_{z_i} + 0.1 eps; QuadLoss, rx = UnitOneSparseConstraint, ry = ZeroReg.
    `rows`: number of rows instead of 1e7 / scale (same generator and centroids)."""
    m = rows if rows is not None else 10_000_000 // scale
    Cn = normal_matrix(seed, 41, centroids, n)
    z = np.minimum((uniform(seed, 42, np.arange(m)) * centroids).astype(np.int64), centroids - 1)
    A = np.empty((m, n), order="F")
    dp = C.POINTER(C.c_double)
    _synth_lib().glrm_synth_c5(C.c_int64(m), C.c_int64(n), C.c_int64(centroids),
                               np.ascontiguousarray(Cn.ravel(order="F")).ctypes.data_as(dp),
                               np.ascontiguousarray(z, dtype=np.int64).ctypes.data_as(C.POINTER(C.c_int64)),
                               C.c_uint64(int(_key(seed, 43))), A.ctypes.data_as(dp))
    return dict(name=f"C5/{scale}", m=m, n=n, k=k, A=A, full=True,
                X0=normal_matrix(seed, 3, k, m), Y0=normal_matrix(seed, 4, k, n))
