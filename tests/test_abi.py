"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, refuses to compute without a device (no CPU fallback), and the host-only shard planner."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lowrankmodels_b200 as lrm
from helpers import glrm_from_config
from lowrankmodels_b200 import _abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "glrm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glrmb200_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _abi.lib()
    declared = header_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/glrm_b200.h but not exported"
    assert sorted(n for n, _, _ in _abi.SYMBOLS) == declared      # the ctypes table mirrors the header
    assert L.glrmb200_version() == 102


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.Params) == 48
    assert C.sizeof(_abi.Problem) == 8 * 4 + 8 * 2 + 3 * 8 * 2 + 8 + 8 + 6 * 8 + 4 * 8   # + the four payload pointers
    assert C.sizeof(_abi.Profile) == 6 * 8 + 5 * 8 + 8


def test_plan_shards_tiles_the_range_and_balances():
    L = _abi.lib()
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 200, size=1000)
    ptr = np.zeros(1001, dtype=np.int64)
    np.cumsum(deg, out=ptr[1:])
    for nranks in (1, 2, 3, 8):
        b = np.zeros(nranks + 1, dtype=np.int64)
        assert L.glrmb200_plan_shards(_abi.i64ptr(ptr), 1000, nranks, _abi.i64ptr(b)) == 0
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        share = np.diff(ptr[b])
        assert share.max() - share.min() <= 2 * deg.max()
    b = np.zeros(5, dtype=np.int64)
    assert L.glrmb200_plan_shards(None, 10, 4, _abi.i64ptr(b)) == 0
    assert list(b) == [0, 2, 5, 7, 10]


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_device_means_error_not_cpu_fallback():
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    with pytest.raises(_abi.GLRMB200Error) as ei:
        lrm.fit_inplace(g, lrm.ProxGradParams(max_iter=2), verbose=False)
    assert ei.value.code == -3      # GLRMB200_E_NO_DEVICE


def test_shape_errors_come_before_any_device_work_and_host_mirror_checks_labels():
    g = glrm_from_config(synth.config1(), lrm.QuadLoss(), lrm.QuadReg(0.1), lrm.QuadReg(0.1))
    ep = lrm.encode_problem(g)
    ep.struct.d = 99
    h = _abi.Handle()
    rc = _abi.lib().glrmb200_create(C.byref(h), C.byref(ep.struct), 0, 0, 1)
    assert rc == -1 and b"embedding" in _abi.lib().glrmb200_last_error()
    # NaN among the observations (glrm.jl:63-71): the host mirror's constructor refuses it
    A = synth.config1()["A"].copy()
    A[3, 4] = np.nan
    with pytest.raises(ValueError, match=r"\(4, 5\) is NaN"):
        lrm.GLRM(A, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2)
    # a label outside {1, 0, -1} for a Boolean loss (myBool, losses.jl:104): refused when encoding
    A = np.where(synth.normal_matrix(1, 1, 20, 10) > 0, 1.0, -1.0)
    A[2, 3] = 2.0
    g2 = lrm.GLRM(A, lrm.LogisticLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2)
    with pytest.raises(ValueError):
        lrm.encode_problem(g2)
    # (the C ABI repeats both checks on the device: tests/test_gpu_parity.py::test_errors_through_the_abi)


def test_host_mirror_constructor_checks():
    A = np.zeros((5, 4))
    with pytest.raises(ValueError):
        lrm.GLRM(A, [lrm.QuadLoss()] * 3, lrm.ZeroReg(), lrm.ZeroReg(), 2)       # glrm.jl:39
    with pytest.raises(ValueError):
        lrm.GLRM(A, lrm.QuadLoss(), [lrm.ZeroReg()] * 2, lrm.ZeroReg(), 2)       # glrm.jl:40
    with pytest.raises(ValueError):
        lrm.GLRM(A, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2, Y=np.zeros((2, 7)))   # glrm.jl:42
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2, X=np.ones((5, 2)))      # transposed X accepted
    assert g.X.shape == (2, 5)
    p = lrm.ProxGradParams(2.0, inner_iter=3)
    assert (p.inner_iter_X, p.inner_iter_Y, p.min_stepsize, p.max_iter) == (3, 3, 0.02, 100)
    with pytest.raises(ValueError, match="DimensionMismatch"):                    # RemQuadReg.m must have k entries
        lrm.encode_problem(lrm.GLRM(A, lrm.QuadLoss(), lrm.RemQuadReg(1.0, np.zeros(3)), lrm.ZeroReg(), 2))
    with pytest.raises(ValueError):                                              # nested wrappers: no device implementation
        lrm.encode_problem(lrm.GLRM(A, lrm.QuadLoss(), lrm.fixed_latent_features(lrm.lastentry1(lrm.QuadReg()), [1.0]),
                                    lrm.ZeroReg(), 2))


def test_plan_dense_rows_is_nested_and_aligned():
    """Host-only planner of the row-sharded fully observed mode: shards tile [0, m), start on 64-row tiles, and the shards of
    2 and 4 ranks are unions of the 8 groups (the reduction tree of the Y sweep does not depend on the number of GPUs)."""
    from lowrankmodels_b200 import distributed as D
    for m in (4096, 19531, 62500, 1_000_000, 10_000_000):
        b8 = D.plan_dense_rows(m, 8)
        assert b8[0] == 0 and b8[-1] == m and (np.diff(b8) > 0).all() and (b8[:-1] % 64 == 0).all()
        assert np.diff(b8).max() <= 1.35 * m / 8 + 64                       # the last groups are not left (nearly) empty
        for nranks in (1, 2, 4):
            b = D.plan_dense_rows(m, nranks)
            assert b[0] == 0 and b[-1] == m and set(b.tolist()) <= set(b8.tolist())
    with pytest.raises(_abi.GLRMB200Error) as ei:
        D.plan_dense_rows(100, 2)                                            # too small to shard by rows: observation-list mode
    assert ei.value.code == -2
    with pytest.raises(_abi.GLRMB200Error):
        D.plan_dense_rows(100000, 3)
