"""Host-side plumbing for N > 1 (one process per GPU).  torch.distributed is used only for rendezvous:
broadcasting NCCL's unique id and host-side barriers / reductions of timings; the data-path exchange
(one all-gather of the freshly updated factor per half-iteration) runs inside the
engine over NCCL (csrc/glrm_engine.cu: allgather_units)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def plan_shards(ptr, count, nranks):
    """Contiguous unit ranges balanced by predicted sweep cost (32-entry chunks + a per-unit overhead) (C ABI glrmb200_plan_shards; host-only).
    ptr: int64 array [count+1] or None.  Returns int64 bounds [nranks+1]."""
    b = np.zeros(nranks + 1, dtype=np.int64)
    p = None if ptr is None else np.ascontiguousarray(ptr, dtype=np.int64)
    _abi.check(_abi.lib().glrmb200_plan_shards(_abi.i64ptr(p) if p is not None else None, int(count), int(nranks),
                                               _abi.i64ptr(b)))
    return b


def plan_dense_rows(m, nranks):
    """Row ranges of a fully observed problem on nranks GPUs (C ABI glrmb200_plan_dense_rows; host-only): whole groups of row
    blocks, so that the Y sweep's fixed-order sums over the 8 groups do not depend on the number of GPUs."""
    b = np.zeros(nranks + 1, dtype=np.int64)
    _abi.check(_abi.lib().glrmb200_plan_dense_rows(int(m), int(nranks), _abi.i64ptr(b)))
    return b


def shard_bounds(ep, nranks):
    """(row_bounds, col_bounds) for an EncodedProblem, exactly as glrmb200_create computes them."""
    s = ep.struct
    if s.obs_full:
        return plan_shards(None, s.m, nranks), plan_shards(None, s.n, nranks)
    return plan_shards(ep.keep["row_ptr"], s.m, nranks), plan_shards(ep.keep["col_ptr"], s.n, nranks)


def broadcast_unique_id(dist, rank, make_id):
    """Rank 0 creates NCCL's 128-byte unique id, every rank receives it (any torch.distributed backend)."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])
