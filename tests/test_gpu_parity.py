"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) with the layouts the reference's
Rcpp exports receive, against the CPU oracle (plain-C restatement; plus the reference's own
src/matmul.cpp from oracle/_ref when that prebuilt library is present).

Bars: bit-exact for index / transpose work; max|got-ref|/max|ref| <= 1e-12 (fp64), <= 1e-5 (fp32).
The shapes replay tests/testthat/test-matmul.R of the reference (100x50 . 50x20 at density .4,
1-row / 1-column operands, binary patterns, SpMV right-hand-side types) and add the edge cases the
GPU decomposition introduces (rows longer than a piece, ragged n, empty rows, unsorted duplicates).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import FP32_TOL, FP64_TOL, NA_INT, all_equal_style, powerlaw_csr, rel_err, rsparsematrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rx():
    from matrixextra_b200 import rcpp_exports
    return rcpp_exports


@pytest.fixture()
def small_piece():
    """Force the long-row piece path on small inputs."""
    from matrixextra_b200 import _lib
    old = _lib.get_option("piece")
    _lib.set_option("piece", 32)
    yield
    _lib.set_option("piece", old)


def _tol(dtype):
    return FP64_TOL if dtype == np.float64 else FP32_TOL


def _sfx(dtype):
    return "numeric" if dtype == np.float64 else "float32"


# ---------------------------------------------------------------------------------------------------
# the reference's own test grid (tests/testthat/test-matmul.R)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(100, 50, 20), (1, 50, 20), (100, 50, 1), (100, 1, 20), (37, 29, 3)])
def test_matmult_dense_csc(rx, port, dtype, shape):
    # test-matmul.R:12-31 — dense(a x K) %*% CSC(K x b)
    a, K, b = shape
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.standard_normal((a, K)).astype(dtype))
    Y = rsparsematrix(K, b, 0.4, 1, "csc")
    got = getattr(rx, "matmul_dense_csc_" + _sfx(dtype))(X, Y.indptr, Y.indices, Y.data, 1)
    want = getattr(port, "matmul_dense_csc_" + _sfx(dtype))(X, Y.indptr, Y.indices, Y.data, 1)
    assert got.flags.f_contiguous and got.dtype == dtype and got.shape == (a, b)
    assert rel_err(got, want) <= _tol(dtype)
    dense = X.astype(np.float64) @ Y.toarray()
    assert all_equal_style(got, dense) <= (1.5e-8 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(100, 50, 20), (1, 50, 20), (100, 50, 1), (64, 300, 77)])
def test_tcrossprod_dense_csr(rx, port, ref, dtype, shape):
    # test-matmul.R:53-96 — tcrossprod(dense(a x K), CSR(b x K))
    a, K, b = shape
    rng = np.random.default_rng(2)
    X = np.asfortranarray(rng.standard_normal((a, K)).astype(dtype))
    Y = rsparsematrix(b, K, 0.4, 2)
    name = "tcrossprod_dense_csr_" + _sfx(dtype)
    got = getattr(rx, name)(X, Y.indptr, Y.indices, Y.data, 1, K)
    assert rel_err(got, getattr(port, name)(X, Y.indptr, Y.indices, Y.data, 1, K)) <= _tol(dtype)
    assert rel_err(got, getattr(ref, name)(X, Y.indptr, Y.indices, Y.data, 1, K)) <= _tol(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(100, 50, 20), (100, 50, 1), (1, 50, 1), (300, 40, 33), (65, 31, 64)])
def test_tcrossprod_csr_dense(rx, port, ref, dtype, shape):
    # test-matmul.R:108-123 — CSR(m x K) %*% dense / tcrossprod(CSR, dense(n x K)); column-major output
    m, K, n = shape
    rng = np.random.default_rng(3)
    X = rsparsematrix(m, K, 0.4, 3)
    Y = np.asfortranarray(rng.standard_normal((n, K)).astype(dtype))
    name = "tcrossprod_csr_dense_" + _sfx(dtype)
    got = getattr(rx, name)(X.indptr, X.indices, X.data, Y, 1)
    assert got.flags.f_contiguous and got.shape == (m, n)
    assert rel_err(got, getattr(port, name)(X.indptr, X.indices, X.data, Y, 1)) <= _tol(dtype)
    if m >= n:  # the reference overflows its scratch row otherwise (src/matmul.cpp:176-182)
        assert rel_err(got, getattr(ref, name)(X.indptr, X.indices, X.data, Y, 1)) <= _tol(dtype)


def test_binary_pattern_csc(rx, port):
    # test-matmul.R:28-31 — binary / logical sparse inputs are coerced to values of 1
    rng = np.random.default_rng(4)
    X = np.asfortranarray(rng.standard_normal((100, 50)))
    Y = rsparsematrix(50, 20, 0.4, 4, "csc")
    ones = np.ones_like(Y.data)
    got = rx.matmul_dense_csc_numeric(X, Y.indptr, Y.indices, ones, 1)
    assert rel_err(got, port.matmul_dense_csc_numeric(X, Y.indptr, Y.indices, ones, 1)) <= FP64_TOL


def test_matmult_csr_vector_types(rx, port, ref):
    # test-matmul.R:134-165 — numeric, integer, logical right-hand sides (+ NA, exercised here)
    A = rsparsematrix(100, 50, 0.4, 5)
    rng = np.random.default_rng(5)
    y = rng.standard_normal(50)
    got = rx.matmul_csr_dvec_numeric(A.indptr, A.indices, A.data, y, 1)
    assert rel_err(got, port.matmul_csr_dvec_numeric(A.indptr, A.indices, A.data, y)) <= FP64_TOL
    assert rel_err(got, ref.matmul_csr_dvec_numeric(A.indptr, A.indices, A.data, y)) <= FP64_TOL

    yi = rng.integers(-5, 6, 50).astype(np.int32)
    got = rx.matmul_csr_dvec_integer(A.indptr, A.indices, A.data, yi, 1)
    assert rel_err(got, port.matmul_csr_dvec_integer(A.indptr, A.indices, A.data, yi)) <= FP64_TOL
    yl = (rng.random(50) < 0.5).astype(np.int32)
    yl[3] = 7  # any non-zero logical payload counts as TRUE (cast to bool, src/matmul.cpp:411)
    got = rx.matmul_csr_dvec_logical(A.indptr, A.indices, A.data, yl, 1)
    assert rel_err(got, port.matmul_csr_dvec_logical(A.indptr, A.indices, A.data, yl)) <= FP64_TOL

    for fn in ("matmul_csr_dvec_integer", "matmul_csr_dvec_logical"):
        yna = yi.copy()
        yna[[0, 17, 49]] = NA_INT
        got = getattr(rx, fn)(A.indptr, A.indices, A.data, yna, 1)
        want = getattr(ref, fn)(A.indptr, A.indices, A.data, yna)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert rel_err(got[ok], want[ok]) <= FP64_TOL
        # rows that met an NA carry R's NA_real_ payload (low word 1954), like the reference on x86
        assert np.array_equal(got[~ok].view(np.uint64), want[~ok].view(np.uint64))

    yf = y.astype(np.float32)
    got = rx.matmul_csr_dvec_float32(A.indptr, A.indices, A.data, yf, 1)
    assert got.dtype == np.float32
    assert rel_err(got, port.matmul_csr_dvec_float32(A.indptr, A.indices, A.data, yf)) <= FP32_TOL


# ---------------------------------------------------------------------------------------------------
# edge cases of the GPU decomposition
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 20, 32, 33, 64, 100, 128, 130, 300])
def test_ragged_column_counts(rx, port, dtype, n):
    A = rsparsematrix(211, 97, 0.15, 6)
    rng = np.random.default_rng(6)
    X = np.asfortranarray(rng.standard_normal((n, 97)).astype(dtype))
    sfx = _sfx(dtype)
    got = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, A.indptr, A.indices, A.data, 1, 97)
    assert rel_err(got, getattr(port, "tcrossprod_dense_csr_" + sfx)(X, A.indptr, A.indices, A.data)) <= _tol(dtype)
    got = getattr(rx, "tcrossprod_csr_dense_" + sfx)(A.indptr, A.indices, A.data, X, 1)
    assert rel_err(got, getattr(port, "tcrossprod_csr_dense_" + sfx)(A.indptr, A.indices, A.data, X)) <= _tol(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_long_rows_are_split_deterministically(rx, port, small_piece, dtype):
    # rows far longer than the piece size (32 here): pieces + fix-up, both output layouts, and SpMV
    p, j, x = powerlaw_csr(400, 3000, 40, 7, cap=2500)
    assert np.max(np.diff(p)) > 500
    rng = np.random.default_rng(7)
    sfx = _sfx(dtype)
    for n in (4, 64, 37):
        X = np.asfortranarray(rng.standard_normal((n, 3000)).astype(dtype))
        got1 = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 3000)
        got2 = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 3000)
        assert np.array_equal(got1, got2)  # run-to-run reproducible (no atomics in the sums)
        assert rel_err(got1, getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x)) <= _tol(dtype)
        got = getattr(rx, "tcrossprod_csr_dense_" + sfx)(p, j, x, X, 1)
        assert rel_err(got, getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, X)) <= _tol(dtype)
    y = rng.standard_normal(3000)
    got = rx.matmul_csr_dvec_numeric(p, j, x, y, 1)
    assert np.array_equal(got, rx.matmul_csr_dvec_numeric(p, j, x, y, 1))
    assert rel_err(got, port.matmul_csr_dvec_numeric(p, j, x, y)) <= FP64_TOL


def test_empty_and_degenerate(rx, port):
    # no stored entries: the reference returns its zero-filled matrix (src/matmul.cpp:128-129, 160-161)
    p = np.zeros(11, dtype=np.int32)
    j = np.zeros(0, dtype=np.int32)
    x = np.zeros(0)
    X = np.asfortranarray(np.random.default_rng(8).standard_normal((4, 6)))
    assert np.array_equal(rx.tcrossprod_dense_csr_numeric(X, p, j, x, 1, 6), np.zeros((4, 10)))
    assert np.array_equal(rx.tcrossprod_csr_dense_numeric(p, j, x, X, 1), np.zeros((10, 4)))
    assert np.array_equal(rx.matmul_csr_dvec_numeric(p, j, x, np.ones(6), 1), np.zeros(10))
    # empty rows in between, explicit zeros, +0.0 for empty rows
    A = sp.csr_matrix(np.array([[0, 0, 0], [1.5, 0, -2], [0, 0, 0], [0, 3, 0]], dtype=np.float64))
    Y = np.asfortranarray(np.array([[1.0, 2, 3], [-1, 0.5, 4]]))
    got = rx.tcrossprod_csr_dense_numeric(A.indptr, A.indices, A.data, Y, 1)
    want = port.tcrossprod_csr_dense_numeric(A.indptr, A.indices, A.data, Y)
    assert np.array_equal(got, want)
    assert not np.signbit(got[0]).any()
    # zero-row and zero-column operands
    assert rx.tcrossprod_csr_dense_numeric(np.zeros(1, np.int32), j, x, X, 1).shape == (0, 4)
    assert rx.tcrossprod_dense_csr_numeric(np.zeros((0, 6), order="F"), p, j, x, 1, 6).shape == (0, 10)


def test_unsorted_indices_and_duplicates(rx, port):
    # unsorted columns inside a row and duplicate entries are legal input
    p = np.array([0, 4, 4, 9], dtype=np.int32)
    j = np.array([3, 0, 3, 1, 2, 2, 0, 2, 1], dtype=np.int32)
    x = np.array([1.0, -2, 0.5, 0.0, 3, 3, -1, 1e-3, 7])
    Y = np.asfortranarray(np.random.default_rng(9).standard_normal((5, 4)))
    got = rx.tcrossprod_csr_dense_numeric(p, j, x, Y, 1)
    assert rel_err(got, port.tcrossprod_csr_dense_numeric(p, j, x, Y)) <= FP64_TOL
    p2, i2, x2 = rx.csr_to_csc(3, 4, p, j, x)
    q2, k2, y2 = port.csr2csc(3, 4, p, j, x)
    assert np.array_equal(p2, q2) and np.array_equal(i2, k2) and np.array_equal(x2, y2)


def test_out_of_range_index_is_an_error_not_a_fault(rx):
    from matrixextra_b200._lib import MXG_ERR_INDEX, MxgError
    p = np.array([0, 2], dtype=np.int32)
    j = np.array([0, 9], dtype=np.int32)
    with pytest.raises(MxgError) as ei:
        rx.matmul_csr_dvec_numeric(p, j, np.ones(2), np.ones(5), 1)
    assert ei.value.code == MXG_ERR_INDEX


def test_reference_fixture_csr(rx, port):
    # the only literal CSR fixture in the reference's tests (tests/testthat/test-utilities.R:33-37)
    p = np.array([0, 1, 4, 5, 6], dtype=np.int32)
    j = np.array([4, 2, 1, 4, 1, 0], dtype=np.int32)
    x = np.array([-0.91, 0.14, -0.12, -0.12, 1.1, 0.66])
    y = np.arange(1.0, 6.0)
    want = sp.csr_matrix((x, j, p), shape=(4, 5)) @ y
    assert rel_err(rx.matmul_csr_dvec_numeric(p, j, x, y, 1), want) <= FP64_TOL
    p2, i2, x2 = rx.csr_to_csc(4, 5, p, j, x)
    csc = sp.csr_matrix((x, j, p), shape=(4, 5)).tocsc()
    assert np.array_equal(p2, csc.indptr) and np.array_equal(i2, csc.indices) and np.array_equal(x2, csc.data)


# ---------------------------------------------------------------------------------------------------
# column panels of the SpMM kernels (one launch per L2-sized slab of the dense operand, csrc/spmm.cu)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(params=[7, 64, 333])
def narrow_panels(request):
    """Force column panels a few columns wide so that small inputs run 2 .. 60 read-modify-write launches."""
    from matrixextra_b200 import _lib
    old = (_lib.get_option("spmm_panel_cols"), _lib.get_option("piece"))
    _lib.set_option("spmm_panel_cols", request.param)
    _lib.set_option("piece", 48)
    yield request.param
    _lib.set_option("spmm_panel_cols", old[0])
    _lib.set_option("piece", old[1])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_column_panels_match_oracle(rx, port, narrow_panels, dtype):
    p, j, x = powerlaw_csr(1500, 400, 10, seed=31, cap=390)  # rows longer than a piece span many panels
    assert np.diff(p).max() > 48
    rng = np.random.default_rng(31)
    sfx = _sfx(dtype)
    for n in (64, 20, 5):
        X = np.asfortranarray(rng.standard_normal((n, 400)).astype(dtype))
        got_rm = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 400)
        got_cm = getattr(rx, "tcrossprod_csr_dense_" + sfx)(p, j, x, X, 1)
        assert rel_err(got_rm, getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 400)) <= _tol(dtype)
        assert rel_err(got_cm, getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, X, 1)) <= _tol(dtype)
        assert np.array_equal(got_rm.T, got_cm)


def test_column_panels_unsorted_rows_and_empty_rows(rx, port, narrow_panels):
    # unsorted columns with duplicates: the split points still partition every row, so only locality suffers
    rng = np.random.default_rng(32)
    m, K, n = 700, 300, 16
    lens = rng.integers(0, 40, m)
    lens[::9] = 0
    p = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    j = rng.integers(0, K, p[-1]).astype(np.int32)
    x = rng.standard_normal(p[-1])
    X = np.asfortranarray(rng.standard_normal((n, K)))
    got = rx.tcrossprod_dense_csr_numeric(X, p, j, x, 1, K)
    assert rel_err(got, port.tcrossprod_dense_csr_numeric(X, p, j, x, 1, K)) <= FP64_TOL
    assert np.all(got[:, lens == 0] == 0.0) and not np.signbit(got[:, lens == 0]).any()  # exact +0.0 rows
    gotc = rx.tcrossprod_csr_dense_numeric(p, j, x, X, 1)
    assert np.array_equal(got.T, gotc)


# ---------------------------------------------------------------------------------------------------
# the streamed level-1 path (row chunks over three streams, csrc/pipeline.cu)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture()
def tiny_chunks():
    """Force the streamed path to cut small inputs into many row chunks (odd entry offsets, chunks with and
    without long rows, more chunks than value staging buffers)."""
    from matrixextra_b200 import _lib
    old = (_lib.get_option("pipe_chunk_nnz"), _lib.get_option("piece"))
    _lib.set_option("pipe_chunk_nnz", 777)
    _lib.set_option("piece", 64)
    yield
    _lib.set_option("pipe_chunk_nnz", old[0])
    _lib.set_option("piece", old[1])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_streamed_chunks_match_oracle_and_whole_upload(rx, port, tiny_chunks, dtype):
    from matrixextra_b200 import _lib
    p, j, x = powerlaw_csr(3000, 900, 12, seed=21, cap=850)  # ~36k entries -> ~45 chunks, several long rows
    assert np.diff(p).max() > 64 and p[-1] > 30 * 777
    rng = np.random.default_rng(21)
    sfx = _sfx(dtype)
    for n in (24, 7):
        X = np.asfortranarray(rng.standard_normal((n, 900)).astype(dtype))
        got_rm = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 900)       # rows-contiguous output
        got_cm = getattr(rx, "tcrossprod_csr_dense_" + sfx)(p, j, x, X, 1)            # column-major output
        assert rel_err(got_rm, getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 900)) <= _tol(dtype)
        assert rel_err(got_cm, getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, X, 1)) <= _tol(dtype)
        assert np.array_equal(got_rm.T, got_cm)  # both layouts hold identical bits
        _lib.set_option("pipeline", 0)
        try:
            whole = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 900)
        finally:
            _lib.set_option("pipeline", 1)
        assert np.array_equal(whole, got_rm)  # chunking never changes a bit: every row is one warp's sum
        # caller-provided result buffer (bench.py hands in page-locked memory)
        buf = np.empty((n, 3000), dtype=dtype, order="F")
        res = getattr(rx, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, 900, out=buf)
        assert res is buf and np.array_equal(buf, got_rm)


def test_streamed_spmv_with_na_and_errors(rx, port, tiny_chunks):
    from matrixextra_b200._lib import MXG_ERR_INDEX, MxgError
    p, j, x = powerlaw_csr(4000, 700, 9, seed=22, cap=650)
    rng = np.random.default_rng(22)
    y = rng.standard_normal(700)
    assert rel_err(rx.matmul_csr_dvec_numeric(p, j, x, y, 1), port.matmul_csr_dvec_numeric(p, j, x, y, 1)) <= FP64_TOL
    yi = rng.integers(-5, 5, 700).astype(np.int32)
    yi[::37] = NA_INT
    got, want = rx.matmul_csr_dvec_integer(p, j, x, yi, 1), port.matmul_csr_dvec_integer(p, j, x, yi, 1)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert rel_err(got[ok], want[ok]) <= FP64_TOL
    yf = rng.standard_normal(700).astype(np.float32)
    assert rel_err(rx.matmul_csr_dvec_float32(p, j, x, yf, 1), port.matmul_csr_dvec_float32(p, j, x, yf, 1)) <= FP32_TOL
    # a bad column id in a late chunk: an error code, never a fault, and the library keeps working afterwards
    jb = j.copy()
    jb[-5] = 700
    with pytest.raises(MxgError) as ei:
        rx.matmul_csr_dvec_numeric(p, jb, x, y, 1)
    assert ei.value.code == MXG_ERR_INDEX
    X = np.asfortranarray(rng.standard_normal((16, 700)))
    jb[-5] = -1
    with pytest.raises(MxgError) as ei:
        rx.tcrossprod_dense_csr_numeric(X, p, jb, x, 1, 700)
    assert ei.value.code == MXG_ERR_INDEX
    pb = p.copy()
    pb[100] = pb[101] + 1  # decreasing indptr
    with pytest.raises(MxgError) as ei:
        rx.tcrossprod_dense_csr_numeric(X, pb, j, x, 1, 700)
    assert ei.value.code == MXG_ERR_INDEX
    assert rel_err(rx.matmul_csr_dvec_numeric(p, j, x, y, 1), port.matmul_csr_dvec_numeric(p, j, x, y, 1)) <= FP64_TOL


# ---------------------------------------------------------------------------------------------------
# CSR -> CSC (bit-exact) and the products built on it
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 1.0), (50, 1, 0.5), (1, 300, 0.5), (300, 255, 0.1), (300, 256, 0.1),
                                   (300, 257, 0.1), (2000, 70000, 0.002), (5000, 300, 0.05), (40, 17000000, 0.0)])
def test_csr2csc_bit_exact(rx, port, shape):
    m, K, density = shape
    if density > 0:
        A = rsparsematrix(m, K, density, 10)
        p, j, x = A.indptr, A.indices, A.data
    else:  # 4 radix passes: K > 2^24, a handful of entries
        rng = np.random.default_rng(10)
        lens = rng.integers(0, 50, m)
        p = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        j = np.concatenate([np.sort(rng.choice(K, l, replace=False)) for l in lens]).astype(np.int32)
        x = rng.standard_normal(p[-1])
    p2, i2, x2 = rx.csr_to_csc(m, K, p, j, x)
    q2, k2, y2 = port.csr2csc(m, K, p, j, x)
    assert np.array_equal(p2, q2)
    assert np.array_equal(i2, k2)
    assert np.array_equal(x2.view(np.uint64), y2.view(np.uint64))
    csc = sp.csr_matrix((x, j, p), shape=(m, K)).tocsc()  # independent second oracle
    assert np.array_equal(p2, csc.indptr) and np.array_equal(i2, csc.indices)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_crossprod_csr_dense(rx, port, dtype):
    # t(CSR) %*% dense: CPU route = stable CSR->CSC, then matmul_dense_csc on t(Y)
    from matrixextra_b200._lib import MXG_F32, MXG_F64
    m, K, n = 500, 120, 24
    A = rsparsematrix(m, K, 0.1, 11)
    rng = np.random.default_rng(11)
    Y = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
    got = rx.crossprod_csr_dense(A.indptr, A.indices, A.data, K, Y, MXG_F64 if dtype == np.float64 else MXG_F32)
    p2, i2, x2 = port.csr2csc(m, K, A.indptr, A.indices, A.data)
    # CSC(A) is CSR(t(A)) (K x m); t(A) %*% Y == t( t(Y) %*% A ) == t(matmul_dense_csc(t(Y), CSC(A)))
    want = getattr(port, "matmul_dense_csc_" + _sfx(dtype))(np.asfortranarray(Y.T), p2, i2, x2).T
    assert got.shape == (K, n) and got.flags.f_contiguous
    assert rel_err(got, want) <= _tol(dtype)


# ---------------------------------------------------------------------------------------------------
# device-resident handles, the synthetic generator and size-independent properties at larger sizes
# ---------------------------------------------------------------------------------------------------
def test_synth_generator_structure_and_device_products(port):
    import torch
    from matrixextra_b200._lib import MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_ROWS_CONTIGUOUS
    from matrixextra_b200.device import DeviceCSR
    for row_model, col_model in ((0, 0), (1, 0), (1, 1)):
        m, K, nnz = 20000, 9000, 600000
        A = DeviceCSR.synth(m, K, nnz, row_model=row_model, col_model=col_model, seed=1003)
        p, j, x = A.to_host()
        assert p[0] == 0 and p[-1] == A.nnz and abs(A.nnz - nnz) <= nnz // 1000
        lens = np.diff(p)
        assert lens.min() >= 0 and lens.max() <= min(K, 65536)
        if row_model == 1:
            assert lens.max() > 20 * lens.mean()  # heavy tail present
            assert A.n_long > 0 and A.n_pieces >= 2 * A.n_long
        assert j.min() >= 0 and j.max() < K
        inside = np.ones(j.size, dtype=bool)
        inside[p[1:-1][p[1:-1] < j.size]] = False  # row starts
        d = np.diff(j.astype(np.int64), prepend=-1)
        assert (d[inside] > 0).all()  # sorted and unique inside every row
        assert np.abs(x).max() < 1.0
        g = torch.Generator(device="cuda").manual_seed(5)
        for dtype, tdt, n in ((MXG_F64, torch.float64, 32), (MXG_F32, torch.float32, 64)):
            B = torch.randn(K, n, device="cuda", dtype=tdt, generator=g)
            out = torch.empty(m, n, device="cuda", dtype=tdt)
            A.spmm(B, out, n, dtype, MXG_ROWS_CONTIGUOUS)
            np_t = np.float64 if dtype == MXG_F64 else np.float32
            Xr = np.asfortranarray(B.cpu().numpy().T)  # (n x K) column-major == K x n row-major
            want = getattr(port, "tcrossprod_dense_csr_" + _sfx(np_t))(Xr, p, j, x)  # (n x m) F-order
            assert rel_err(out.cpu().numpy(), want.T) <= _tol(np_t)
            outc = torch.empty(n, m, device="cuda", dtype=tdt)  # column-major m x n
            A.spmm(B, outc, n, dtype, MXG_COLS_CONTIGUOUS)
            assert torch.equal(outc.T.contiguous(), out)  # both layouts hold identical bits
        y = torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
        o = torch.empty(m, device="cuda", dtype=torch.float64)
        A.spmv(y, o)
        assert rel_err(o.cpu().numpy(), port.matmul_csr_dvec_numeric(p, j, x, y.cpu().numpy())) <= FP64_TOL
        At = A.transpose()
        p2, i2, x2 = At.to_host()
        q2, k2, y2 = port.csr2csc(m, K, p, j, x)
        assert np.array_equal(p2, q2) and np.array_equal(i2, k2) and np.array_equal(x2, y2)
        At.free()
        A.free()


def test_properties_at_scale():
    """BASELINE-scale shapes (a 1/4-size cfg2/cfg3 matrix so it stays quick): properties that need no
    CPU pass — linearity, transpose round trip (idempotence), column sums through the transpose."""
    import torch
    from matrixextra_b200._lib import MXG_F32, MXG_F64, MXG_ROWS_CONTIGUOUS
    from matrixextra_b200.device import DeviceCSR
    m, K, nnz = 500_000, 1_000_000, 25_000_000
    A = DeviceCSR.synth(m, K, nnz, row_model=1, col_model=1, seed=1002)
    g = torch.Generator(device="cuda").manual_seed(7)
    y1 = torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
    y2 = torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
    o1, o2, o3 = (torch.empty(m, device="cuda", dtype=torch.float64) for _ in range(3))
    A.spmv(y1, o1)
    A.spmv(y2, o2)
    A.spmv(2.0 * y1 - 3.0 * y2, o3)
    scale = o3.abs().max().item()
    assert (o3 - (2.0 * o1 - 3.0 * o2)).abs().max().item() <= 1e-12 * scale * 50
    # SpMM column c equals SpMV with column c of B
    n = 64
    B = torch.randn(K, n, device="cuda", dtype=torch.float64, generator=g)
    out = torch.empty(m, n, device="cuda", dtype=torch.float64)
    A.spmm(B, out, n, MXG_F64, MXG_ROWS_CONTIGUOUS)
    for c in (0, 31, 63):
        A.spmv(B[:, c].contiguous(), o1)
        assert (out[:, c] - o1).abs().max().item() <= 1e-12 * out[:, c].abs().max().item()
    # fp32 product against the fp64 one
    out32 = torch.empty(m, n, device="cuda", dtype=torch.float32)
    A.spmm(B.float(), out32, n, MXG_F32, MXG_ROWS_CONTIGUOUS)
    assert (out32.double() - out).abs().max().item() <= 1e-4 * out.abs().max().item()
    # transpose twice == identity, bit for bit; t(A) . 1 == column sums == scatter-add of the values
    At = A.transpose()
    Att = At.transpose()
    pa, ja, xa = A.to_host()
    pb, jb, xb = Att.to_host()
    assert np.array_equal(pa, pb) and np.array_equal(ja, jb) and np.array_equal(xa, xb)
    ones = torch.ones(m, device="cuda", dtype=torch.float64)
    colsum = torch.empty(K, device="cuda", dtype=torch.float64)
    At.spmv(ones, colsum)
    want = np.zeros(K)
    np.add.at(want, ja, xa)
    assert np.max(np.abs(colsum.cpu().numpy() - want)) <= 1e-12 * max(1.0, np.abs(want).max())
    for h in (Att, At, A):
        h.free()


# ---------------------------------------------------------------------------------------------------
# float32 vector %*% CSC / tcrossprod(float32 vector, CSR) / crossprod(float32 vector, CSC)
# (matmul_rowvec_by_csc[bin], src/matmul.cpp:643-684; R/matmul.R:243-259, 350-366, 402-427)
# ---------------------------------------------------------------------------------------------------
def test_float32_row_vector_times_csc(rx, port):
    from matrixextra_b200 import crossprod, dgCMatrix, dgRMatrix, float32, matmul, tcrossprod
    rng = np.random.default_rng(77)
    for K, ncols, dens in ((300, 120, 0.2), (50, 1, 0.6), (2000, 900, 0.01)):
        Y = rsparsematrix(K, ncols, dens, 7, "csc")
        rv = rng.standard_normal(K).astype(np.float32)
        want = port.matmul_rowvec_by_csc(rv, Y.indptr, Y.indices, Y.data)
        got = rx.matmul_rowvec_by_csc(rv, Y.indptr, Y.indices, Y.data)
        assert got.shape == (1, ncols) and got.dtype == np.float32
        assert rel_err(got, want) <= FP32_TOL
        got_bin = rx.matmul_rowvec_by_cscbin(rv, Y.indptr, Y.indices)
        assert rel_err(got_bin, port.matmul_rowvec_by_csc(rv, Y.indptr, Y.indices, None)) <= FP32_TOL
        if K > 1:
            # the S4 routes that end in it
            Yc = dgCMatrix.from_scipy(Y)
            assert np.array_equal(matmul(float32(rv), Yc).Data, got)
            assert np.array_equal(crossprod(float32(rv), Yc).Data, got)
            if ncols > 1:
                Yr = dgRMatrix.from_scipy(sp.csr_matrix(Y.T))  # CSR of t(Y): tcrossprod(x, t(Y)) == x %*% Y
                assert np.array_equal(tcrossprod(float32(rv), Yr).Data, got)
    with pytest.raises(ValueError, match="vector-Matrix multiplication dimensions do not match"):
        matmul(float32(np.ones(5, dtype=np.float32)), dgCMatrix.from_scipy(rsparsematrix(7, 3, 0.5, 1, "csc")))
    with pytest.raises(ValueError, match="vector-Matrix crossprod dimensions do not match"):
        crossprod(float32(np.ones(5, dtype=np.float32)), dgCMatrix.from_scipy(rsparsematrix(7, 3, 0.5, 1, "csc")))
