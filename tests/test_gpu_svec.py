"""GPU parity for SURVEY.md §8 f2: CSR %*% sparseVector through the C ABI (mxg_spmv_csr_svec) against the CPU
oracle — the plain-C restatement and, when present, the reference's own matmul_csr_svec_* (src/matmul.cpp:486-641)
compiled in place.  Bar: max|got-ref|/max|ref| <= 1e-12 (sums are reassociated), NA pattern and payload identical.
Shapes follow tests/testthat/test-matmul.R:139-159 (sparse-vector right-hand sides of every class)."""
import numpy as np
import pytest

from helpers import FP64_TOL, NA_INT, powerlaw_csr, rel_err, rsparsematrix

pytestmark = pytest.mark.gpu

NA_QUIET = 0x7FF80000000007A2  # NA_real_ after one x86 arithmetic operation (payload 1954 kept)


@pytest.fixture(scope="module")
def rx():
    from matrixextra_b200 import rcpp_exports
    return rcpp_exports


def _svec(K, nnz, seed, kind):
    rng = np.random.default_rng(seed)
    i = np.sort(rng.choice(K, size=min(nnz, K), replace=False)).astype(np.int32) + 1
    if kind == "numeric":
        v = rng.standard_normal(i.size)
    elif kind == "integer":
        v = rng.integers(-9, 10, i.size).astype(np.int32)
    elif kind == "logical":
        v = rng.integers(0, 2, i.size).astype(np.int32)
    elif kind == "float32":
        v = rng.standard_normal(i.size).astype(np.float32)
    else:
        v = None
    return i, v


def _call(mod, kind, p, j, x, yi, yv, **kw):
    fn = getattr(mod, "matmul_csr_svec_" + kind)
    return fn(p, j, x, yi, **kw) if kind == "binary" else fn(p, j, x, yi, yv, **kw)


@pytest.mark.parametrize("kind", ["numeric", "integer", "logical", "binary", "float32"])
@pytest.mark.parametrize("shape", [(100, 50, 0.4, 20), (1, 50, 0.4, 10), (100, 1, 0.5, 1), (300, 2000, 0.05, 700)])
def test_svec_matches_oracle(rx, port, kind, shape):
    m, K, dens, ny = shape
    A = rsparsematrix(m, K, dens, 3)
    yi, yv = _svec(K, ny, 4, kind)
    got = _call(rx, kind, A.indptr, A.indices, A.data, yi, yv, ncols=K)
    want = _call(port, kind, A.indptr, A.indices, A.data, yi, yv)
    assert got.dtype == np.float64 and got.shape == (m,)
    assert rel_err(got, want) <= FP64_TOL
    # K unknown (what the reference's export receives): same answer
    got2 = _call(rx, kind, A.indptr, A.indices, A.data, yi, yv)
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("kind", ["numeric", "integer", "logical", "binary"])
def test_svec_matches_reference_build(rx, ref, kind):
    A = rsparsematrix(500, 800, 0.03, 11)
    yi, yv = _svec(800, 120, 12, kind)
    got = _call(rx, kind, A.indptr, A.indices, A.data, yi, yv, ncols=800)
    want = _call(ref, kind, A.indptr, A.indices, A.data, yi, yv)
    assert rel_err(got, want) <= FP64_TOL


@pytest.mark.parametrize("kind", ["integer", "logical", "numeric"])
def test_svec_na_rules(rx, port, kind):
    # src/matmul.cpp:523-528: an NA entry of y contributes NA_real_ to every row that stores its column
    A = rsparsematrix(400, 300, 0.1, 21)
    yi, yv = _svec(300, 60, 22, kind)
    if kind == "numeric":
        yv[[5, 17]] = np.frombuffer(np.uint64(0x7FF00000000007A2).tobytes(), dtype=np.float64)[0]
    else:
        yv[[5, 17]] = NA_INT
    got = _call(rx, kind, A.indptr, A.indices, A.data, yi, yv, ncols=300)
    want = _call(port, kind, A.indptr, A.indices, A.data, yi, yv)
    na_w = np.isnan(want)
    assert na_w.any() and np.array_equal(np.isnan(got), na_w)
    assert rel_err(got[~na_w], want[~na_w]) <= FP64_TOL
    assert np.all(got[na_w].view(np.uint64) == NA_QUIET)
    assert np.all(want[na_w].view(np.uint64) == NA_QUIET)


def test_svec_long_rows_and_pieces(rx, port):
    # rows longer than one piece go through the partial-sum + fix-up path
    from matrixextra_b200 import _lib
    p, j, x = powerlaw_csr(3000, 5000, 30, seed=5, cap=4000)
    yi, yv = _svec(5000, 1500, 6, "numeric")
    want = port.matmul_csr_svec_numeric(p, j, x, yi, yv)
    got = rx.matmul_csr_svec_numeric(p, j, x, yi, yv, ncols=5000)
    assert rel_err(got, want) <= FP64_TOL
    old = _lib.get_option("piece")
    _lib.set_option("piece", 32)
    try:
        got2 = rx.matmul_csr_svec_numeric(p, j, x, yi, yv, ncols=5000)
    finally:
        _lib.set_option("piece", old)
    assert rel_err(got2, want) <= FP64_TOL


def test_svec_global_bitmap_equals_shared(rx):
    from matrixextra_b200 import _lib
    p, j, x = powerlaw_csr(2000, 3000, 20, seed=7)
    yi, yv = _svec(3000, 900, 8, "numeric")
    a = rx.matmul_csr_svec_numeric(p, j, x, yi, yv, ncols=3000)
    _lib.set_option("svec_smem", 0)
    try:
        b = rx.matmul_csr_svec_numeric(p, j, x, yi, yv, ncols=3000)
    finally:
        _lib.set_option("svec_smem", 1)
    assert np.array_equal(a, b)  # same team shapes, same order: bit-identical


def test_svec_edge_cases(rx, port):
    A = rsparsematrix(50, 40, 0.3, 31)
    p, j, x = A.indptr, A.indices, A.data
    # empty vector -> zeros (src/matmul.cpp:495-496)
    got = rx.matmul_csr_svec_numeric(p, j, x, np.zeros(0, np.int32), np.zeros(0), ncols=40)
    assert np.array_equal(got, np.zeros(50))
    # indices beyond ncol(X) never match; first of a repeated index wins (the merge consumes it)
    yi = np.array([3, 3, 7, 41, 1000], dtype=np.int32)
    yv = np.array([1.5, 99.0, -2.0, 5.0, 6.0])
    got = rx.matmul_csr_svec_numeric(p, j, x, yi, yv, ncols=40)
    want = port.matmul_csr_svec_numeric(p, j, x, yi, yv)
    assert rel_err(got, want) <= FP64_TOL
    # empty matrix rows / zero-row matrix
    pe = np.zeros(6, dtype=np.int32)
    got = rx.matmul_csr_svec_numeric(pe, np.zeros(0, np.int32), np.zeros(0), yi, yv, ncols=40)
    assert np.array_equal(got, np.zeros(5))
    got = rx.matmul_csr_svec_numeric(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0), yi, yv, ncols=40)
    assert got.shape == (0,)
    # unsorted rows of A and unsorted y: membership test does not care
    rng = np.random.default_rng(0)
    jj, xx = j.copy(), x.copy()
    for r in range(50):
        q = rng.permutation(p[r + 1] - p[r])
        jj[p[r]:p[r + 1]] = j[p[r]:p[r + 1]][q]
        xx[p[r]:p[r + 1]] = x[p[r]:p[r + 1]][q]
    yi2, yv2 = _svec(40, 15, 9, "numeric")
    q = rng.permutation(15)
    got = rx.matmul_csr_svec_numeric(p, jj, xx, yi2[q], yv2[q], ncols=40)
    want = port.matmul_csr_svec_numeric(p, j, x, yi2, yv2)
    assert rel_err(got, want) <= FP64_TOL


def test_svec_unsorted_vector_with_distant_repeats_keeps_the_first(rx, port):
    # mxgpu.h: "a repeated index keeps its first entry; y need not be sorted" — also when the repeats are far apart
    # (resolved with an integer atomicMin over the positions: the same entry wins on every run)
    A = rsparsematrix(400, 300, 0.2, 61)
    p, j, x = A.indptr, A.indices, A.data
    rng = np.random.default_rng(61)
    base = rng.choice(300, 120, replace=False).astype(np.int32) + 1
    yi = np.concatenate([base, base[::-1][:80], base[:40]]).astype(np.int32)  # every index up to three times, far apart
    yv = rng.standard_normal(yi.size)
    first = {}
    for k, c in enumerate(yi):
        first.setdefault(int(c), k)
    keep = np.array(sorted(first.values()))
    order = np.argsort(yi[keep], kind="stable")
    want = port.matmul_csr_svec_numeric(p, j, x, yi[keep][order], yv[keep][order])  # the sorted, de-duplicated vector
    runs = [rx.matmul_csr_svec_numeric(p, j, x, yi, yv, ncols=300) for _ in range(5)]
    assert rel_err(runs[0], want) <= FP64_TOL
    for r in runs[1:]:
        assert np.array_equal(r, runs[0])


def test_svec_s4_dispatch(port):
    # R/matmul.R:595-646 through the S4 mirror
    from matrixextra_b200 import dgRMatrix, matmul, sparseVector
    A = rsparsematrix(120, 90, 0.2, 41)
    X = dgRMatrix(A.indptr, A.indices, A.data, A.shape)
    yi, yv = _svec(90, 30, 42, "numeric")
    got = matmul(X, sparseVector(yi, yv, 90, "d"))
    assert got.shape == (120, 1)
    assert rel_err(got[:, 0], port.matmul_csr_svec_numeric(A.indptr, A.indices, A.data, yi, yv)) <= FP64_TOL
    got = matmul(X, sparseVector(yi, None, 90, "n"))
    assert rel_err(got[:, 0], port.matmul_csr_svec_binary(A.indptr, A.indices, A.data, yi)) <= FP64_TOL
    with pytest.raises(ValueError, match="dimensions do not match"):
        matmul(X, sparseVector(yi, yv, 91, "d"))


def test_svec_device_handle(port):
    # level 2: device-resident handle, device vectors
    import torch
    from matrixextra_b200.device import DeviceCSR
    p, j, x = powerlaw_csr(5000, 4000, 25, seed=51)
    yi, yv = _svec(4000, 800, 52, "numeric")
    A = DeviceCSR.upload(5000, 4000, p, j, x)
    d_yi = torch.from_numpy(yi).cuda()
    d_yv = torch.from_numpy(yv).cuda()
    d_out = torch.empty(5000, dtype=torch.float64, device="cuda")
    A.spmv_svec(d_yi, d_yv, d_out)
    torch.cuda.synchronize()
    assert rel_err(d_out.cpu().numpy(), port.matmul_csr_svec_numeric(p, j, x, yi, yv)) <= FP64_TOL
