"""Encode a GLRM into the C ABI's `glrmb200_problem` (include/glrm_b200.h).

This is the work the Julia shim does before its `ccall`: flatten
glrm.observed_features / glrm.observed_examples (src/glrm.jl:17-18) into CSR / CSC with int32 0-based
indices — order and duplicates preserved bit-exactly — gather A's values next to each list so the
device never performs an `A[e,f]` lookup (proxgrad.jl:125,168), and turn losses / regularizers into
descriptor tables.  Label checks reproduce the reference's dispatch-time errors (myBool,
losses.jl:104-106; `u[a]` bounds for categorical losses).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .glrm import GLRM, Repeated
from .losses import encode_losses, get_yidxs
from .regularizers import encode_payloads, encode_regs


class EncodedProblem:
    """Owns the numpy buffers a `glrmb200_problem` points into (keeps them alive)."""

    def __init__(self):
        self.struct = _abi.Problem()
        self.keep = {}

    def set(self, name, arr, ptrfn):
        self.keep[name] = arr
        setattr(self.struct, name, ptrfn(arr))

    @property
    def nnz(self):
        s = self.struct
        return int(s.m * s.n) if s.obs_full else int(self.keep["row_ptr"][-1])


def _encode_reg_list(regs):
    """-> (codes, params, payload_ptr | None, payload | None)"""
    if isinstance(regs, Repeated):
        code, p = regs.item.encode()
        return (np.array([code], dtype=np.int32), p.reshape(1, -1).copy()) + encode_payloads([regs.item])
    codes, params = encode_regs(regs)
    if len(codes) == 1 and len(regs) > 1:
        return (codes, params, None, None)
    return (codes, params) + encode_payloads(regs)


def _check_labels(glrm: GLRM, cols, vals):
    """cols[t] = feature of entry t, vals[t] = A value.  Raise like the reference would."""
    kinds = np.array([{"real": 0, "bool": 1, "level": 2}[l.label_kind] for l in glrm.losses], dtype=np.int8)
    if not kinds.any():
        return
    kk = kinds[cols]
    b = kk == 1
    if b.any():
        v = vals[b]
        if not np.isin(v, (1.0, 0.0, -1.0)).all():
            raise ValueError("InexactError: Boolean losses take labels 1 (true) or 0/-1 (false) "
                             "(myBool, losses.jl:104)")
    lv = kk == 2
    if lv.any():
        v = vals[lv]
        mx = np.array([getattr(l, "max", 0) for l in glrm.losses], dtype=np.float64)[cols[lv]]
        if not ((v == np.floor(v)) & (v >= 1) & (v <= mx)).all():
            raise ValueError("BoundsError: categorical / ordinal levels must be integers in 1..max")


def encode_problem(glrm: GLRM, validate=True) -> EncodedProblem:
    m, n = glrm.shape
    ep = EncodedProblem()
    s = ep.struct
    ystart = get_yidxs(glrm.losses)
    s.m, s.n, s.k, s.d = m, n, glrm.k, int(ystart[-1])

    if isinstance(glrm.losses, Repeated):
        code, p = glrm.losses.item.encode()
        lcodes = np.full(n, code, dtype=np.int32)
        lparams = np.tile(p, (n, 1))
    else:
        lcodes, lparams = encode_losses(glrm.losses)
    ep.set("loss_code", np.ascontiguousarray(lcodes), _abi.i32ptr)
    ep.set("loss_param", np.ascontiguousarray(lparams), _abi.dptr)
    rxc, rxp, rxpp, rxpv = _encode_reg_list(glrm.rx)
    ryc, ryp, rypp, rypv = _encode_reg_list(glrm.ry)
    s.rx_count, s.ry_count = len(rxc), len(ryc)
    ep.set("rx_code", rxc, _abi.i32ptr)
    ep.set("rx_param", np.ascontiguousarray(rxp), _abi.dptr)
    ep.set("ry_code", ryc, _abi.i32ptr)
    ep.set("ry_param", np.ascontiguousarray(ryp), _abi.dptr)
    # vector payloads (fixed_latent_features.y, RemQuadReg.m): their lengths are checked like the reference's indexing would
    for side, ptr, vals, regs in (("rx", rxpp, rxpv, glrm.rx), ("ry", rypp, rypv, glrm.ry)):
        if ptr is None:
            continue
        for r in (regs.item,) if isinstance(regs, Repeated) else regs:
            pay = r.payload()
            if pay is not None and (len(pay) > glrm.k or (type(r).__name__ == "RemQuadReg" and len(pay) != glrm.k)):
                raise ValueError(f"DimensionMismatch: {type(r).__name__} payload of length {len(pay)} with k = {glrm.k}")
        ep.set(side + "_payload_ptr", ptr, _abi.i64ptr)
        ep.set(side + "_payload", vals if len(vals) else np.zeros(1), _abi.dptr)

    feats, exs = glrm.observed_features, glrm.observed_examples
    if feats.full is not None and exs.full is not None:
        s.obs_full = 1
        A = glrm.A.toarray() if glrm._sparse else glrm.A
        dense = np.asfortranarray(A, dtype=np.float64)
        if validate:
            _check_labels(glrm, np.repeat(np.arange(n), m), dense.ravel(order="F"))
        ep.set("dense_A", dense, _abi.dptr)
        return ep

    # materialise a UnitRange side if only one side was given explicitly
    def lists(o, count, length):
        if o.full is None:
            return o.ptr, o.idx
        ptr = np.arange(count + 1, dtype=np.int64) * length
        return ptr, np.tile(np.arange(length, dtype=np.int64), count)

    rptr, ridx = lists(feats, m, n)
    cptr, cidx = lists(exs, n, m)
    if max(m, n) >= 2**31:
        raise ValueError("index does not fit int32")
    rows_of = np.repeat(np.arange(m, dtype=np.int64), np.diff(rptr))
    cols_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(cptr))
    rval = glrm._row_val if glrm._row_val is not None else glrm.values_at(rows_of, ridx)
    cval = glrm._col_val if glrm._col_val is not None else glrm.values_at(cidx, cols_of)
    if validate:
        _check_labels(glrm, ridx, rval)
        _check_labels(glrm, cols_of, cval)
    s.obs_full = 0
    ep.set("row_ptr", np.ascontiguousarray(rptr, dtype=np.int64), _abi.i64ptr)
    ep.set("row_idx", np.ascontiguousarray(ridx, dtype=np.int32), _abi.i32ptr)
    ep.set("row_val", np.ascontiguousarray(rval, dtype=np.float64), _abi.dptr)
    ep.set("col_ptr", np.ascontiguousarray(cptr, dtype=np.int64), _abi.i64ptr)
    ep.set("col_idx", np.ascontiguousarray(cidx, dtype=np.int32), _abi.i32ptr)
    ep.set("col_val", np.ascontiguousarray(cval, dtype=np.float64), _abi.dptr)
    return ep


def encode_params(p) -> _abi.Params:
    return _abi.Params(p.stepsize, p.max_iter, p.inner_iter_X, p.inner_iter_Y, p.abs_tol, p.rel_tol,
                       p.min_stepsize)


def encode_sparse_params(p) -> _abi.SparseParams:
    return _abi.SparseParams(p.stepsize, p.max_iter, p.inner_iter, p.abs_tol, p.min_stepsize)
