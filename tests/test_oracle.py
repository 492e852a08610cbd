"""CPU tests that pin the oracle (oracle/mx_oracle.c) before any GPU parity claim rests on it:
  * against the committed golden vectors produced by the reference's own src/matmul.cpp
    (tests/golden/make_golden.py) — bit for bit;
  * against that reference library live, when oracle/_ref is present — bit for bit;
  * against the dense product the reference's testthat cases assert, and scipy for CSR->CSC."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import FP32_TOL, FP64_TOL, NA_INT, all_equal_style, powerlaw_csr, rel_err, rsparsematrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matmul_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


def test_port_reproduces_reference_golden_vectors_bit_exactly(port, golden):
    n = int(golden["n_cases"])
    assert n >= 12
    for c in range(n):
        pre = f"c{c}_"
        sfx = str(golden[pre + "kind"])
        p, j, x, X = golden[pre + "p"], golden[pre + "j"], golden[pre + "x"], golden[pre + "X"]
        got = getattr(port, "matmul_dense_csc_" + sfx)(X, p, j, x)
        assert np.array_equal(_bits(got), _bits(golden[pre + "matmul_dense_csc"])), (c, sfx)
        got = getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x)
        assert np.array_equal(_bits(got), _bits(golden[pre + "tcrossprod_dense_csr"])), (c, sfx)
        if pre + "tcrossprod_csr_dense" in golden.files:
            got = getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, X)
            want = golden[pre + "tcrossprod_csr_dense"]
            assert got.shape == want.shape and got.flags.f_contiguous
            assert np.array_equal(_bits(np.asfortranarray(got).ravel(order="K")), _bits(np.asfortranarray(want).ravel(order="K")))


def test_port_spmv_golden_vectors_including_na(port, golden):
    p, j, x = golden["v_p"], golden["v_j"], golden["v_x"]
    assert np.array_equal(_bits(port.matmul_csr_dvec_numeric(p, j, x, golden["v_y"])), _bits(golden["v_numeric"]))
    assert np.array_equal(_bits(port.matmul_csr_dvec_integer(p, j, x, golden["v_yi"])), _bits(golden["v_integer"]))
    assert np.array_equal(_bits(port.matmul_csr_dvec_logical(p, j, x, golden["v_yl"])), _bits(golden["v_logical"]))
    assert np.array_equal(_bits(port.matmul_csr_dvec_float32(p, j, x, golden["v_y"].astype(np.float32))),
                          _bits(golden["v_float32"]))
    for fn, y, want in (("integer", "v_yi_na", "v_integer_na"), ("logical", "v_yl_na", "v_logical_na")):
        got = getattr(port, "matmul_csr_dvec_" + fn)(p, j, x, golden[y])
        assert np.isnan(got).sum() > 0
        assert np.array_equal(_bits(got), _bits(golden[want]))  # NA payload (1954) included


def test_reference_literal_fixture(port, golden):
    # tests/testthat/test-utilities.R:33-37 — 4x5 CSR with an unsorted row
    p, j, x = golden["f_p"], golden["f_j"], golden["f_x"]
    dense = np.zeros((4, 5))
    for r in range(4):
        for e in range(p[r], p[r + 1]):
            dense[r, j[e]] += x[e]
    got = port.matmul_csr_dvec_numeric(p, j, x, np.arange(1.0, 6.0))
    assert np.array_equal(_bits(got), _bits(golden["f_spmv"]))
    assert rel_err(got, dense @ np.arange(1.0, 6.0)) <= FP64_TOL
    p2, i2, x2 = port.csr2csc(4, 5, p, j, x)
    assert np.array_equal(p2, golden["f_p2"]) and np.array_equal(i2, golden["f_i2"]) and np.array_equal(x2, golden["f_x2"])


def test_csr2csc_golden_and_scipy(port, golden):
    p2, i2, x2 = port.csr2csc(300, 257, golden["t_p"], golden["t_j"], golden["t_x"])
    assert np.array_equal(p2, golden["t_p2"]) and np.array_equal(i2, golden["t_i2"]) and np.array_equal(x2, golden["t_x2"])
    # duplicates and unsorted rows keep their stored order inside a column (stable)
    p = np.array([0, 4, 4, 9], dtype=np.int32)
    j = np.array([3, 0, 3, 1, 2, 2, 0, 2, 1], dtype=np.int32)
    x = np.arange(9.0)
    p2, i2, x2 = port.csr2csc(3, 4, p, j, x)
    assert p2.tolist() == [0, 2, 4, 7, 9]
    assert i2.tolist() == [0, 2, 0, 2, 2, 2, 2, 0, 0]
    assert x2.tolist() == [1.0, 6.0, 3.0, 8.0, 4.0, 5.0, 7.0, 0.0, 2.0]
    with pytest.raises(ValueError):
        port.csr2csc(1, 2, np.array([0, 1], np.int32), np.array([5], np.int32), np.ones(1))
    for seed in range(3):
        A = rsparsematrix(200 + seed, 1000, 0.02, seed)
        q2, k2, y2 = port.csr2csc(A.shape[0], 1000, A.indptr, A.indices, A.data)
        C = A.tocsc()
        assert np.array_equal(q2, C.indptr) and np.array_equal(k2, C.indices) and np.array_equal(y2, C.data)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_port_equals_reference_library_live(port, ref, seed):
    rng = np.random.default_rng(seed)
    m, K, n = int(rng.integers(40, 400)), int(rng.integers(20, 300)), int(rng.integers(1, 40))
    A = rsparsematrix(m, K, 0.2, seed)
    for dt, sfx in ((np.float64, "numeric"), (np.float32, "float32")):
        X = np.asfortranarray(rng.standard_normal((n, K)).astype(dt))
        a = getattr(port, "tcrossprod_dense_csr_" + sfx)(X, A.indptr, A.indices, A.data)
        b = getattr(ref, "tcrossprod_dense_csr_" + sfx)(X, A.indptr, A.indices, A.data, 2, K)
        assert np.array_equal(_bits(a), _bits(b))
        if m >= n:
            a = getattr(port, "tcrossprod_csr_dense_" + sfx)(A.indptr, A.indices, A.data, X)
            b = getattr(ref, "tcrossprod_csr_dense_" + sfx)(A.indptr, A.indices, A.data, X, 3)
            assert np.array_equal(a, b)
    y = rng.standard_normal(K)
    assert np.array_equal(port.matmul_csr_dvec_numeric(A.indptr, A.indices, A.data, y),
                          ref.matmul_csr_dvec_numeric(A.indptr, A.indices, A.data, y, 4))
    yi = rng.integers(-3, 4, K).astype(np.int32)
    yi[0] = NA_INT
    assert np.array_equal(_bits(port.matmul_csr_dvec_integer(A.indptr, A.indices, A.data, yi)),
                          _bits(ref.matmul_csr_dvec_integer(A.indptr, A.indices, A.data, yi, 1)))
    # multi-threaded port == single-threaded port (rows are independent)
    a1 = port.tcrossprod_dense_csr_numeric(X.astype(np.float64), A.indptr, A.indices, A.data, 1)
    a4 = port.tcrossprod_dense_csr_numeric(X.astype(np.float64), A.indptr, A.indices, A.data, 4)
    assert np.array_equal(a1, a4)


def test_algebraic_parity_spec_of_the_reference_tests(port):
    # every expect_equal in tests/testthat/test-matmul.R is "equals the dense product": tolerance 1.5e-8
    # mean-relative for fp64 (testthat default), 1e-5 for float32 (test-matmul.R:20-21, 60-64)
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.standard_normal((100, 50)))
    Yc = rsparsematrix(50, 20, 0.4, 1, "csc")
    assert all_equal_style(port.matmul_dense_csc_numeric(X, Yc.indptr, Yc.indices, Yc.data), X @ Yc.toarray()) < 1.5e-8
    got32 = port.matmul_dense_csc_float32(X.astype(np.float32), Yc.indptr, Yc.indices, Yc.data)
    assert got32.dtype == np.float32 and all_equal_style(got32, X @ Yc.toarray()) < 1e-5
    Yr = rsparsematrix(20, 50, 0.4, 2)
    assert all_equal_style(port.tcrossprod_dense_csr_numeric(X, Yr.indptr, Yr.indices, Yr.data), X @ Yr.toarray().T) < 1.5e-8
    A = rsparsematrix(100, 50, 0.4, 3)
    D = np.asfortranarray(rng.standard_normal((20, 50)))
    assert all_equal_style(port.tcrossprod_csr_dense_numeric(A.indptr, A.indices, A.data, D), A.toarray() @ D.T) < 1.5e-8
    # empty matrix: zero-filled output untouched; empty rows stay +0
    z = port.tcrossprod_csr_dense_numeric(np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0), D)
    assert z.shape == (5, 20) and not z.any()


def test_long_rows_and_powerlaw_inputs(port):
    p, j, x = powerlaw_csr(300, 2000, 30, 4, cap=1500)
    y = np.random.default_rng(4).standard_normal(2000)
    want = sp.csr_matrix((x, j, p), shape=(300, 2000)) @ y
    assert rel_err(port.matmul_csr_dvec_numeric(p, j, x, y), want) <= FP64_TOL
    got32 = port.matmul_csr_dvec_float32(p, j, x, y.astype(np.float32))
    assert rel_err(got32, want) <= 20 * FP32_TOL  # the reference narrows after every term (src/matmul.cpp:403, 476)


def test_rowvec_by_csc_restatement_equals_reference(port, ref):
    """src/matmul.cpp:643-684 (float32 row vector %*% CSC, with and without stored values): the plain-C restatement
    is the reference bit for bit, and both are the dense product within float32 rounding."""
    import scipy.sparse as sp
    rng = np.random.default_rng(12)
    for (K, ncols, dens) in ((50, 30, 0.3), (1, 7, 1.0), (200, 1, 0.5), (64, 40, 0.0)):
        Y = sp.random(K, ncols, dens, format="csc", random_state=rng, dtype=np.float64)
        Y.sort_indices()
        rv = rng.standard_normal(K).astype(np.float32)
        for x in (Y.data, None):
            a = port.matmul_rowvec_by_csc(rv, Y.indptr, Y.indices, x)
            b = ref.matmul_rowvec_by_csc(rv, Y.indptr, Y.indices, x)
            assert a.shape == (1, ncols) and np.array_equal(a.view(np.uint32), b.view(np.uint32))
            dense = Y.toarray() if x is not None else (Y.toarray() != 0).astype(np.float64)
            if x is None and dens > 0:
                dense = np.zeros((K, ncols))
                dense[Y.indices, np.repeat(np.arange(ncols), np.diff(Y.indptr))] = 1.0
            assert np.allclose(a.ravel(), rv.astype(np.float64) @ dense, rtol=1e-5, atol=1e-5)


def test_host_twin_of_the_synthetic_generator():
    """oracle/mx_synth.c (what bench.py's reference arm multiplies): exact entry counts, sorted unique in-range columns,
    the power-law tail and cap, values in [-1, 1), determinism."""
    from oracle.cpu_oracle import synth_csr_host
    p, j, x = synth_csr_host(10_000, 5_000, 500_000, row_model=0, col_model=0, seed=1001)  # BASELINE cfg1
    lens = np.diff(p)
    assert p[0] == 0 and p[-1] == 500_000 and lens.min() >= 40 and lens.max() <= 60
    m, K, nnz = 40_000, 30_000, 2_000_000
    for col_model in (0, 1):
        p, j, x = synth_csr_host(m, K, nnz, row_model=1, col_model=col_model, seed=1003)
        lens = np.diff(p)
        assert p[0] == 0 and abs(int(p[-1]) - nnz) <= nnz // 1000 and lens.min() >= 0 and lens.max() <= K
        assert lens.max() > 20 * lens.mean()  # heavy tail
        assert j.min() >= 0 and j.max() < K and np.abs(x).max() < 1.0
        inside = np.ones(j.size, dtype=bool)
        starts = p[1:-1]
        inside[starts[starts < j.size]] = False
        assert (np.diff(j.astype(np.int64), prepend=-1)[inside] > 0).all()  # sorted and unique inside every row
        if col_model == 1:  # popular low columns: the first fifth of the columns holds far more than a fifth of the entries
            assert (j < K // 5).mean() > 0.35
        p2, j2, x2 = synth_csr_host(m, K, nnz, row_model=1, col_model=col_model, seed=1003)
        assert np.array_equal(p, p2) and np.array_equal(j, j2) and np.array_equal(x, x2)
