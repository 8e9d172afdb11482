"""fit! / fit — mirror of /root/reference/src/fit.jl:8-31 and the `fit!(glrm, params; ch, verbose)`
method contract of src/algorithms/proxgrad.jl:34-37,219, routed to the CUDA engine through the C ABI.

`Engine` is the residency handle the reference's re-fitting callers want (cross_validate.jl:141-240):
create once (uploads A / obs lists), fit many times.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .convergence import ConvergenceHistory, update_ch
from .encode import EncodedProblem, encode_params, encode_problem, encode_sparse_params
from .glrm import GLRM
from .params import ProxGradParams, SparseProxGradParams


class Engine:
    GATHER_ONLY = 1          # GLRMB200_CREATE_GATHER_ONLY

    def __init__(self, glrm_or_problem, device=0, rank=0, nranks=1, validate=True, gather_only=False):
        self.ep = (glrm_or_problem if isinstance(glrm_or_problem, EncodedProblem)
                   else encode_problem(glrm_or_problem, validate=validate))
        self.h = _abi.Handle()
        L = _abi.lib()
        _abi.check(L.glrmb200_create_ex(C.byref(self.h), C.byref(self.ep.struct), device, rank, nranks,
                                        self.GATHER_ONLY if gather_only else 0))
        s = self.ep.struct
        self.m, self.n, self.k, self.d = int(s.m), int(s.n), int(s.k), int(s.d)
        self.nranks = nranks
        self.last_profile = None

    # -- multi-GPU plumbing -----------------------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _abi.check(_abi.lib().glrmb200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid: bytes = None):
        """uid=None reuses the communicator this process created earlier (cached inside the library)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(uid) if uid is not None else None
        _abi.check(_abi.lib().glrmb200_comm_init(self.h, buf))

    def peer_init(self, dist):
        """Fused exchange over NVLink peer memory: all-gather the CUDA IPC blobs (any torch.distributed backend)
        and open the peers' factor replicas.  Call after comm_init on every rank."""
        if self.nranks == 1:
            return
        buf = (C.c_uint8 * 64)()
        _abi.check(_abi.lib().glrmb200_ipc_export(self.h, buf))
        blobs = [None] * self.nranks
        dist.all_gather_object(blobs, bytes(buf))
        allb = (C.c_uint8 * (64 * self.nranks)).from_buffer_copy(b"".join(blobs))
        _abi.check(_abi.lib().glrmb200_ipc_open(self.h, allb))
        self._collective_close = True

    def barrier(self):
        _abi.check(_abi.lib().glrmb200_comm_barrier(self.h))

    def shard(self):
        v = [C.c_int64() for _ in range(4)]
        _abi.check(_abi.lib().glrmb200_shard(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # -- the hot path ---------------------------------------------------------------------------------
    def _run(self, params, X, Y, resident):
        cap = params.max_iter + 1
        obj = np.zeros(cap)
        sec = np.zeros(cap)
        nrec = C.c_int32(0)
        prof = _abi.Profile()
        prm = encode_params(params)
        L = _abi.lib()
        if resident:
            rc = L.glrmb200_fit_resident(self.h, C.byref(prm), _abi.dptr(obj), _abi.dptr(sec), cap,
                                         C.byref(nrec), C.byref(prof))
        else:
            rc = L.glrmb200_fit(self.h, C.byref(prm), _abi.dptr(X), _abi.dptr(Y), _abi.dptr(obj),
                                _abi.dptr(sec), cap, C.byref(nrec), C.byref(prof))
        _abi.check(rc)
        self.last_profile = prof.as_dict()
        return obj[:nrec.value], sec[:nrec.value]

    def fit(self, params, X, Y):
        """X (k,m), Y (k,d) Fortran-ordered float64, updated in place.  Returns (objective, seconds)."""
        assert X.flags.f_contiguous and Y.flags.f_contiguous and X.dtype == np.float64 == Y.dtype
        assert X.shape == (self.k, self.m) and Y.shape == (self.k, self.d)
        return self._run(params, X, Y, resident=False)

    def fit_sparse(self, params, X, Y):
        """fit!(glrm, ::SparseProxGradParams) on the engine: X, Y in/out (best model found).  Returns (objective, seconds)
        exactly as the reference records them (initial, one per accepted iteration, final duplicate)."""
        assert X.flags.f_contiguous and Y.flags.f_contiguous and X.dtype == np.float64 == Y.dtype
        cap = params.max_iter + 2
        obj, sec = np.zeros(cap), np.zeros(cap)
        nrec = C.c_int32(0)
        prof = _abi.Profile()
        prm = encode_sparse_params(params)
        _abi.check(_abi.lib().glrmb200_fit_sparse(self.h, C.byref(prm), _abi.dptr(X), _abi.dptr(Y), _abi.dptr(obj),
                                                  _abi.dptr(sec), cap, C.byref(nrec), C.byref(prof)))
        self.last_profile = prof.as_dict()
        return obj[:nrec.value], sec[:nrec.value]

    def upload(self, X, Y):
        assert X.flags.f_contiguous and Y.flags.f_contiguous
        _abi.check(_abi.lib().glrmb200_upload_factors(self.h, _abi.dptr(X), _abi.dptr(Y)))

    def fit_resident(self, params):
        return self._run(params, None, None, resident=True)

    def download(self, X, Y):
        assert X.flags.f_contiguous and Y.flags.f_contiguous
        _abi.check(_abi.lib().glrmb200_download_factors(self.h, _abi.dptr(X), _abi.dptr(Y)))

    def objective(self, X, Y, include_regularization=True):
        out = C.c_double(0)
        X = np.asfortranarray(X, dtype=np.float64)
        Y = np.asfortranarray(Y, dtype=np.float64)
        _abi.check(_abi.lib().glrmb200_objective(self.h, _abi.dptr(X), _abi.dptr(Y),
                                                 int(bool(include_regularization)), C.byref(out)))
        return out.value

    def set_reg_scale(self, newscale):
        _abi.check(_abi.lib().glrmb200_set_reg_scale(self.h, float(newscale)))

    def set_obs(self, ep):
        """Swap in the observation lists of another EncodedProblem (same A shape / losses / regularizers): the
        training fold of cross_validate (src/cross_validate.jl:31-33) without re-creating the handle."""
        k_ = ep.keep
        _abi.check(_abi.lib().glrmb200_set_obs(self.h, _abi.i64ptr(k_["row_ptr"]), _abi.i32ptr(k_["row_idx"]),
                                               _abi.dptr(k_["row_val"]), _abi.i64ptr(k_["col_ptr"]),
                                               _abi.i32ptr(k_["col_idx"]), _abi.dptr(k_["col_val"])))
        self.ep = ep

    def stepsizes(self):
        ar, ac = np.zeros(self.m), np.zeros(self.n)
        _abi.check(_abi.lib().glrmb200_get_stepsizes(self.h, _abi.dptr(ar), _abi.dptr(ac)))
        return ar, ac

    def close(self):
        if self.h and getattr(self, "_collective_close", False):
            self._collective_close = False
            try:
                self.barrier()          # peers must be done storing into this rank's replicas
            except Exception:
                pass
        if self.h:
            _abi.lib().glrmb200_destroy(self.h)
            self.h = _abi.Handle()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fit_inplace(glrm: GLRM, params: ProxGradParams = None, *, ch: ConvergenceHistory = None,
                verbose=True, engine: Engine = None, **kwargs):
    """`fit!(glrm, params; ch, verbose)`: mutates glrm.X / glrm.Y in place and returns
    (glrm.X, glrm.Y, ch) (proxgrad.jl:34-37,43,219); appends to a caller-supplied `ch`
    (cross_validate.jl:174); warm-starts from the current factors."""
    if params is None:                                                          # fit.jl:13-19
        params = SparseProxGradParams() if glrm._sparse else ProxGradParams()
    if ch is None:
        ch = ConvergenceHistory("B200ProxGradGLRM")
    own = engine is None
    if own:
        engine = Engine(glrm, device=getattr(params, "device", 0), gather_only=isinstance(params, SparseProxGradParams))
    try:
        if verbose:
            print("Fitting GLRM")                                              # proxgrad.jl:75
        if isinstance(params, SparseProxGradParams):
            obj, sec = engine.fit_sparse(params, glrm.X, glrm.Y)
        else:
            obj, sec = engine.fit(params, glrm.X, glrm.Y)
        for i, (o, s) in enumerate(zip(obj, sec)):
            update_ch(ch, s, o)                                                # proxgrad.jl:76,207
            if verbose and i > 0 and i % 10 == 0:
                print(f"Iteration {i}: objective value = {o}")                 # proxgrad.jl:214-216
    finally:
        if own:
            engine.close()
    return glrm.X, glrm.Y, ch


def fit(glrm: GLRM, *args, **kwargs):
    """`fit(glrm, args...)` (fit.jl:24-31): fit without modifying the glrm; returns (X', Y, ch)."""
    X0, Y0 = glrm.X.copy(order="F"), glrm.Y.copy(order="F")
    X, Y, ch = fit_inplace(glrm, *args, **kwargs)
    Xo, Yo = X.copy(order="F"), Y.copy(order="F")
    glrm.X[...] = X0
    glrm.Y[...] = Y0
    return Xo.T, Yo, ch
