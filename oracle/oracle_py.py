"""ctypes wrapper of oracle/libglrm_oracle.so — TEST INFRASTRUCTURE, NOT PRODUCT.
Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglrm_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_loss_eval.restype = C.c_double
        L.oracle_loss_eval.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_double, C.POINTER(C.c_int)]
        L.oracle_loss_grad.restype = None
        L.oracle_loss_grad.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)]
        L.oracle_reg_eval.restype = C.c_double
        L.oracle_reg_eval.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int64]
        L.oracle_reg_prox.restype = None
        L.oracle_reg_prox.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int64, C.c_double]
        L.oracle_objective.restype = C.c_double
        L.oracle_objective.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.POINTER(C.c_int)]
        L.oracle_fit.restype = C.c_int
        L.oracle_fit.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp, C.c_int32,
                                 C.POINTER(C.c_int32), C.c_int32, C.c_int32, _dp, _dp,
                                 C.POINTER(C.c_int64)]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_fit_units.restype = C.c_int
        L.oracle_fit_units.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp, C.c_int32, C.POINTER(C.c_int32),
                                       C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        L.oracle_fit_sparse.restype = C.c_int
        L.oracle_fit_sparse.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp, C.c_int32, C.POINTER(C.c_int32), _dp]
        L.oracle_half_sweep.restype = C.c_int
        L.oracle_half_sweep.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, _dp, C.c_int32, C.c_int64, C.c_int64, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


_blas = None


def use_openblas_dgemm(nthreads=0):
    """Route the faithful form's XY = X'Y through cblas_dgemm of the OpenBLAS bundled with NumPy (ILP64 symbols
    `scipy_cblas_dgemm64_`), as the reference does through Julia's BLAS.  Returns True when the entry point was found."""
    global _blas
    import glob
    L = lib()
    L.oracle_set_dgemm.restype = None
    L.oracle_set_dgemm.argtypes = [C.c_void_p]
    if _blas is None:
        cands = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
        for path in cands:
            try:
                B = C.CDLL(path)
                fn = C.cast(B.scipy_cblas_dgemm64_, C.c_void_p).value
                _blas = (B, fn)
                break
            except (OSError, AttributeError):
                continue
    if _blas is None:
        return False
    if nthreads > 0:
        try:
            _blas[0].scipy_openblas_set_num_threads64_(C.c_int(nthreads))
        except AttributeError:
            pass
    L.oracle_set_dgemm(C.c_void_p(_blas[1]))
    return True


def loss_eval(loss, u, a):
    code, p = loss.encode()
    u = np.atleast_1d(np.asarray(u, dtype=np.float64)).copy()
    err = C.c_int(0)
    r = lib().oracle_loss_eval(code, _d(p), _d(u), len(u), float(a), C.byref(err))
    if err.value:
        raise ValueError(f"oracle loss error {err.value}")
    return r


def loss_grad(loss, u, a):
    code, p = loss.encode()
    scalar = np.ndim(u) == 0
    u = np.atleast_1d(np.asarray(u, dtype=np.float64)).copy()
    g = np.zeros_like(u)
    err = C.c_int(0)
    lib().oracle_loss_grad(code, _d(p), _d(u), len(u), float(a), _d(g), C.byref(err))
    if err.value:
        raise ValueError(f"oracle loss error {err.value}")
    return float(g[0]) if scalar else g


def reg_eval(reg, v):
    code, p = reg.encode()
    v = np.asfortranarray(np.asarray(v, dtype=np.float64))
    k, D = (v.shape[0], 1) if v.ndim == 1 else v.shape
    return lib().oracle_reg_eval(code, _d(p), _d(v), k, D)


def reg_prox(reg, v, alpha):
    code, p = reg.encode()
    out = np.array(v, dtype=np.float64, order="F")
    k, D = (out.shape[0], 1) if out.ndim == 1 else out.shape
    lib().oracle_reg_prox(code, _d(p), _d(out), k, D, float(alpha))
    return out


def objective(ep, X, Y, include_reg=True):
    """ep: lowrankmodels_b200.encode.EncodedProblem"""
    X = np.asfortranarray(X, dtype=np.float64)
    Y = np.asfortranarray(Y, dtype=np.float64)
    err = C.c_int(0)
    r = lib().oracle_objective(C.addressof(ep.struct), _d(X), _d(Y), int(include_reg), C.byref(err))
    if err.value:
        raise ValueError(f"oracle error {err.value}")
    return r


def fit(ep, params_struct, X, Y, mode=1, nthreads=0):
    """Runs the restated fit! in place on Fortran-ordered X (k,m), Y (k,d).
    mode 0 = faithful dense-XY, 1 = sparse-evaluated.  Returns dict(objective, seconds, alpharow,
    alphacol, trials)."""
    assert X.flags.f_contiguous and Y.flags.f_contiguous
    cap = params_struct.max_iter + 1
    obj, sec = np.zeros(cap), np.zeros(cap)
    nrec = C.c_int32(0)
    ar, ac = np.zeros(int(ep.struct.m)), np.zeros(int(ep.struct.n))
    trials = (C.c_int64 * 2)()
    rc = lib().oracle_fit(C.addressof(ep.struct), C.addressof(params_struct), _d(X), _d(Y), _d(obj), _d(sec),
                          cap, C.byref(nrec), mode, nthreads, _d(ar), _d(ac), trials)
    if rc:
        raise ValueError(f"oracle_fit error {rc}")
    return dict(objective=obj[:nrec.value].copy(), seconds=sec[:nrec.value].copy(), alpharow=ar, alphacol=ac,
                trials=(trials[0], trials[1]))


def fit_units(ep, params_struct, X, Y, rows, cols, mode=0, nthreads=0):
    """The restated fit! over a sample of the units only: rows [rows[0], rows[1]) in the X sweeps, columns
    [cols[0], cols[1]) in the Y sweeps, everything else frozen (bench.py's bounded CPU step on the full-size problem).
    Returns dict(objective, seconds, trials); objective[t] is the sampled columns' share, objective[0] is 0."""
    assert X.flags.f_contiguous and Y.flags.f_contiguous
    cap = params_struct.max_iter + 1
    obj, sec = np.zeros(cap), np.zeros(cap)
    nrec = C.c_int32(0)
    trials = (C.c_int64 * 2)()
    rc = lib().oracle_fit_units(C.addressof(ep.struct), C.addressof(params_struct), _d(X), _d(Y), _d(obj), _d(sec), cap,
                                C.byref(nrec), mode, nthreads, int(rows[0]), int(rows[1]), int(cols[0]), int(cols[1]), trials)
    if rc:
        raise ValueError(f"oracle_fit_units error {rc}")
    return dict(objective=obj[:nrec.value].copy(), seconds=sec[:nrec.value].copy(), trials=(trials[0], trials[1]))


def half_sweep(ep, params_struct, X, Y, alpha, which, begin, end, obj_by_unit):
    """Update units [begin, end) of X (which=0) or Y (which=1) in place (sparse-evaluated form)."""
    rc = lib().oracle_half_sweep(C.addressof(ep.struct), C.addressof(params_struct), _d(X), _d(Y), _d(alpha),
                                 which, begin, end, _d(obj_by_unit))
    if rc:
        raise ValueError(f"oracle_half_sweep error {rc}")


def fit_sparse(ep, sparse_params_struct, X, Y):
    """Restated fit!(glrm, ::SparseProxGradParams); X, Y in/out (best model).  Returns dict(objective, seconds, alpha)."""
    assert X.flags.f_contiguous and Y.flags.f_contiguous
    cap = sparse_params_struct.max_iter + 2
    obj, sec = np.zeros(cap), np.zeros(cap)
    nrec = C.c_int32(0)
    alpha = np.zeros(1)
    rc = lib().oracle_fit_sparse(C.addressof(ep.struct), C.addressof(sparse_params_struct), _d(X), _d(Y), _d(obj), _d(sec),
                                 cap, C.byref(nrec), _d(alpha))
    if rc:
        raise ValueError(f"oracle_fit_sparse error {rc}")
    return dict(objective=obj[:nrec.value].copy(), seconds=sec[:nrec.value].copy(), alpha=float(alpha[0]))
