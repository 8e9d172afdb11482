"""Observation bookkeeping must be bit-exact: the CSR/CSC arrays the encoder hands to
the C ABI against a literal push!-loop restatement of sort_observations (modify_glrm.jl:5-18), and the
structural property the reference tests (test/sparse_test.jl:21-46)."""
import numpy as np
import scipy.sparse as sp

import lowrankmodels_b200 as lrm
from helpers import small_sparse
from lowrankmodels_b200 import synth


def push_loop(obs, m, n):
    feats = [[] for _ in range(m)]
    exs = [[] for _ in range(n)]
    for i, j in obs:
        feats[i].append(j)
        exs[j].append(i)
    return feats, exs


def test_sort_observations_matches_push_loop_with_duplicates_and_order():
    A, obs, _ = small_sparse(dup=True)
    m, n = A.shape
    feats, exs = push_loop(obs.tolist(), m, n)
    g = lrm.GLRM(A, lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 3, obs=obs)
    for i in range(m):
        assert list(g.observed_features[i]) == feats[i]
    for j in range(n):
        assert list(g.observed_examples[j]) == exs[j]
    ep = lrm.encode_problem(g)
    rp, ri, rv = ep.keep["row_ptr"], ep.keep["row_idx"], ep.keep["row_val"]
    cp, ci, cv = ep.keep["col_ptr"], ep.keep["col_idx"], ep.keep["col_val"]
    assert ri.dtype == np.int32 and ci.dtype == np.int32 and rp.dtype == np.int64
    assert rp[-1] == len(obs) == cp[-1]
    for i in range(m):
        assert ri[rp[i]:rp[i + 1]].tolist() == feats[i]
        assert (rv[rp[i]:rp[i + 1]] == A[i, feats[i]]).all()          # values bit-exact, co-located
    for j in range(n):
        assert ci[cp[j]:cp[j + 1]].tolist() == exs[j]
        assert (cv[cp[j]:cp[j + 1]] == A[exs[j], j]).all()


def test_sparse_matrix_nonzeros_are_the_observations():
    """test/sparse_test.jl:21-46: j in observed_features[i]  <=>  A[i,j] != 0 (and the transpose)."""
    m, n = 100, 100
    u = synth.uniform(4, 1, np.arange(m * n)).reshape(m, n)
    dense = np.where(u < 0.5, synth.uniform(4, 2, np.arange(m * n)).reshape(m, n), 0.0)
    g = lrm.GLRM(sp.csc_matrix(dense), lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 3)
    assert len(g.observed_features) == m and len(g.observed_examples) == n
    for i in range(m):
        assert set(g.observed_features[i].tolist()) == set(np.flatnonzero(dense[i]).tolist())
    for j in range(n):
        assert set(g.observed_examples[j].tolist()) == set(np.flatnonzero(dense[:, j]).tolist())
    ep = lrm.encode_problem(g)
    # CSC order as findall(!iszero, A) gives it (glrm.jl:46-48): rows ascending inside a column
    cp, ci = ep.keep["col_ptr"], ep.keep["col_idx"]
    for j in range(n):
        assert (np.diff(ci[cp[j]:cp[j + 1]]) > 0).all()
    assert (ep.keep["col_val"] == dense.T[dense.T != 0]).all()


def test_default_is_fully_observed_unit_ranges():
    g = lrm.GLRM(np.ones((7, 5)), lrm.QuadLoss(), lrm.ZeroReg(), lrm.ZeroReg(), 2)
    assert g.observed_features.full == (7, 5) and list(g.observed_features[3]) == [0, 1, 2, 3, 4]
    ep = lrm.encode_problem(g)
    assert ep.struct.obs_full == 1 and ep.nnz == 35
