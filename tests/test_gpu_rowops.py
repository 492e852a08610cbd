"""GPU parity for SURVEY.md §8 f3 / f4 through the C ABI: per-row index sorting, CSR validity checks and the elementwise
CSR * dense products, against the CPU oracle (plain-C restatement, and the reference's own src/misc.cpp / src/operators.cpp
compiled in place when oracle/_ref/libmxref_ops.so is present).  Everything here is integer work or ONE multiply per
stored entry, so the bar is bit equality."""
import numpy as np
import pytest

from helpers import NA_INT, powerlaw_csr, rsparsematrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rx():
    from matrixextra_b200 import rcpp_exports
    return rcpp_exports


@pytest.fixture(scope="module")
def refops():
    from oracle.cpu_oracle import Ref
    if not (Ref.available() and Ref.ops_available()):
        pytest.skip("oracle/_ref/libmxref_ops.so not built")
    return Ref()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def scramble_rows(p, j, x, seed, frac=1.0):
    """Permute the entries inside a fraction of the rows (the rest stay sorted)."""
    rng = np.random.default_rng(seed)
    j2, x2 = j.copy(), x.copy()
    for r in range(p.size - 1):
        a, b = p[r], p[r + 1]
        if b - a > 1 and rng.random() < frac:
            q = rng.permutation(b - a)
            j2[a:b] = j[a:b][q]
            x2[a:b] = x[a:b][q]
    return j2, x2


# ---------------------------------------------------------------------------------------------------
# f4: elementwise products
# ---------------------------------------------------------------------------------------------------
DENSE_KINDS = [("double", np.float64), ("float32", np.float32), ("int", np.int32), ("bool", np.int32)]


def _dense(m, K, kind, seed, with_na=True):
    rng = np.random.default_rng(seed)
    if kind == "double":
        return np.asfortranarray(rng.standard_normal((m, K)))
    if kind == "float32":
        return np.asfortranarray(rng.standard_normal((m, K)).astype(np.float32))
    d = rng.integers(-4, 5, (m, K)) if kind == "int" else rng.integers(0, 2, (m, K))
    d = np.asfortranarray(d.astype(np.int32))
    if with_na:
        d[rng.integers(0, m, 25), rng.integers(0, K, 25)] = NA_INT
    return d


@pytest.mark.parametrize("kind,np_t", DENSE_KINDS)
@pytest.mark.parametrize("shape", [(100, 50, 0.4), (1, 50, 0.5), (100, 1, 0.5), (500, 300, 0.02)])
def test_mul_dense_bit_exact(rx, port, kind, np_t, shape):
    m, K, dens = shape
    A = rsparsematrix(m, K, dens, 5)
    D = _dense(m, K, kind, 6)
    fn = "multiply_csr_by_dense_elemwise_" + kind
    got = getattr(rx, fn)(A.indptr, A.indices, A.data, D)
    want = getattr(port, fn)(A.indptr, A.indices, A.data, D)
    assert got.shape == want.shape and np.array_equal(bits(got), bits(want))
    # the flat column-major vector form the reference's export receives
    got2 = getattr(rx, fn)(A.indptr, A.indices, A.data, D.reshape(-1, order="F"))
    assert np.array_equal(bits(got2), bits(got))


@pytest.mark.parametrize("kind,np_t", DENSE_KINDS)
def test_mul_dense_matches_reference_build(rx, refops, kind, np_t):
    A = rsparsematrix(300, 200, 0.05, 7)
    D = _dense(300, 200, kind, 8)
    fn = "multiply_csr_by_dense_elemwise_" + kind
    got = getattr(rx, fn)(A.indptr, A.indices, A.data, D)
    want = getattr(refops, fn)(A.indptr, A.indices, A.data, D)
    assert np.array_equal(bits(got), bits(want))


def test_mul_dense_long_rows(rx, port):
    from matrixextra_b200 import _lib
    p, j, x = powerlaw_csr(400, 3000, 40, seed=9, cap=2500)  # rows beyond one 1024-entry piece
    D = _dense(400, 3000, "double", 10)
    want = port.multiply_csr_by_dense_elemwise_double(p, j, x, D)
    assert np.array_equal(bits(rx.multiply_csr_by_dense_elemwise_double(p, j, x, D)), bits(want))
    old = _lib.get_option("piece")
    _lib.set_option("piece", 32)
    try:
        assert np.array_equal(bits(rx.multiply_csr_by_dense_elemwise_double(p, j, x, D)), bits(want))
    finally:
        _lib.set_option("piece", old)


@pytest.mark.parametrize("length", ["m", "mK", "m/4", "77", "mK+5", "3m+1", "1"])
def test_mul_dvec_recycling_bit_exact(rx, port, length):
    m, K = 200, 150
    A = rsparsematrix(m, K, 0.1, 11)
    n = {"m": m, "mK": m * K, "m/4": m // 4, "77": 77, "mK+5": m * K + 5, "3m+1": 3 * m + 1, "1": 1}[length]
    v = np.random.default_rng(12).standard_normal(n)
    got = rx.multiply_csr_by_dvec_no_NAs_numeric(A.indptr, A.indices, A.data, v, K)
    want = port.multiply_csr_by_dvec_no_NAs_numeric(A.indptr, A.indices, A.data, v, K)
    assert np.array_equal(bits(got), bits(want))
    # R semantics spelled out: the vector is recycled down the columns of the dense m x K shape
    full = np.resize(v, m * K).reshape((m, K), order="F") if n <= m * K else v[:m * K].reshape((m, K), order="F")
    rows = np.repeat(np.arange(m), np.diff(A.indptr))
    assert np.array_equal(bits(got), bits(A.data * full[rows, A.indices]))


def test_mul_dvec_matches_reference_build(rx, refops):
    A = rsparsematrix(240, 100, 0.08, 13)
    for n in (240, 60, 77, 24000, 1000):
        v = np.random.default_rng(n).standard_normal(n)
        got = rx.multiply_csr_by_dvec_no_NAs_numeric(A.indptr, A.indices, A.data, v, 100)
        want = refops.multiply_csr_by_dvec_no_NAs_numeric(A.indptr, A.indices, A.data, v, 100)
        assert np.array_equal(bits(got), bits(want))


def test_mul_edge_cases_and_s4(rx, port):
    from matrixextra_b200 import dgRMatrix
    from matrixextra_b200.operators import multiply, options
    # empty matrix / empty rows
    pe = np.zeros(6, dtype=np.int32)
    assert rx.multiply_csr_by_dense_elemwise_double(pe, np.zeros(0, np.int32), np.zeros(0), np.ones((5, 3))).size == 0
    A = rsparsematrix(60, 40, 0.2, 14)
    X = dgRMatrix(A.indptr, A.indices, A.data, A.shape)
    D = _dense(60, 40, "double", 15)
    out = multiply(X, D)
    assert isinstance(out, dgRMatrix) and np.array_equal(out.j, X.j) and np.array_equal(out.p, X.p)
    assert np.array_equal(bits(out.x), bits(port.multiply_csr_by_dense_elemwise_double(A.indptr, A.indices, A.data, D)))
    v = np.random.default_rng(16).standard_normal(60)
    out = multiply(v, X)  # the vignette's `X * as.numeric(pred - y)` with the operands swapped
    assert np.array_equal(bits(out.x), bits(A.data * np.repeat(v, np.diff(A.indptr))))
    with pytest.raises(ValueError, match="same dimensions"):
        multiply(X, np.ones((60, 41)))
    Dn = D.copy()
    Dn[0, 0] = np.nan
    with pytest.raises(NotImplementedError):
        multiply(X, Dn)
    options["MatrixExtra.ignore_na"] = True
    try:
        out = multiply(X, Dn)
        assert np.array_equal(bits(out.x), bits(port.multiply_csr_by_dense_elemwise_double(A.indptr, A.indices, A.data, Dn)))
    finally:
        options["MatrixExtra.ignore_na"] = False


# ---------------------------------------------------------------------------------------------------
# f3: validity
# ---------------------------------------------------------------------------------------------------
def test_check_valid_csr_codes(rx, port):
    A = rsparsematrix(200, 150, 0.1, 21)
    p, j = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    cases = [(p, j)]
    jb = j.copy(); jb[5] = -1; cases.append((p, jb))
    jb = j.copy(); jb[-1] = 150; cases.append((p, jb))
    jb = j.copy(); jb[7] = NA_INT; cases.append((p, jb))
    pb = p.copy(); pb[5] = NA_INT; cases.append((pb, j))
    pb = p.copy(); pb[100] = pb[101] + 1; cases.append((pb, j))
    pb = p.copy(); pb[5] = NA_INT; jb = j.copy(); jb[3] = 9999; cases.append((pb, jb))  # index check comes first
    seen = set()
    for pp, jj in cases:
        got = rx.check_valid_csr_matrix(pp, jj, 200, 150)
        want = port.check_valid_csr_matrix(pp, jj, 200, 150)
        assert got.get("err") == want
        seen.add(want)
    assert len(seen) == 5
    # no stored entries: only the pointer checks apply
    assert rx.check_valid_csr_matrix(np.zeros(4, np.int32), np.zeros(0, np.int32), 3, 10) == {}


def test_check_valid_matches_reference_build(rx, refops):
    A = rsparsematrix(80, 60, 0.2, 22)
    p, j = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    jb = j.copy(); jb[11] = 60
    pb = p.copy(); pb[40] = pb[41] + 3
    for pp, jj in ((p, j), (p, jb), (pb, j)):
        assert rx.check_valid_csr_matrix(pp, jj, 80, 60).get("err") == refops.check_valid_csr_matrix(pp, jj, 80, 60)


# ---------------------------------------------------------------------------------------------------
# f3: sortedness and sorting
# ---------------------------------------------------------------------------------------------------
def test_rows_sorted_flag(rx, port):
    A = rsparsematrix(500, 300, 0.05, 31)
    p, j, x = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    assert rx.check_indices_are_unsorted(p, j) is True and port.check_indices_are_unsorted(p, j)
    for r in (0, 123, 499):  # a single swapped pair anywhere is found
        a, b = p[r], p[r + 1]
        if b - a >= 2:
            jb = j.copy()
            jb[[a, a + 1]] = jb[[a + 1, a]]
            assert rx.check_indices_are_unsorted(p, jb) is False and not port.check_indices_are_unsorted(p, jb)
    # equal neighbours count as sorted (src/misc.cpp:124: strict <)
    jd = j.copy()
    a = p[np.argmax(np.diff(p) >= 2)]
    jd[a + 1] = jd[a]
    assert rx.check_indices_are_unsorted(p, jd) == port.check_indices_are_unsorted(p, jd) == True  # noqa: E712
    # descending ids across a row boundary are fine; empty rows in between too
    p2 = np.array([0, 2, 2, 2, 4, 4], dtype=np.int32)
    j2 = np.array([5, 9, 0, 3], dtype=np.int32)
    assert rx.check_indices_are_unsorted(p2, j2) is True
    assert rx.check_indices_are_unsorted(p2, np.array([5, 9, 3, 0], dtype=np.int32)) is False


@pytest.mark.parametrize("frac", [1.0, 0.3])
def test_sort_matches_oracle(rx, port, frac):
    p, j, x = powerlaw_csr(3000, 20000, 30, seed=33, cap=9000)  # warp-sized, CTA-sized and sorted rows
    js, xs = scramble_rows(p, j, x, 34, frac)
    wj, wx = port.sort_sparse_indices_numeric(p, js, xs)
    assert np.array_equal(wj, j) and np.array_equal(wx, x)  # distinct ids: the sorted form is unique
    gj, gx = js.copy(), xs.copy()
    rx.sort_sparse_indices_numeric(p, gj, gx)
    assert np.array_equal(gj, wj) and np.array_equal(bits(gx), bits(wx))
    rx.sort_sparse_indices_numeric(p, gj, gx)  # idempotent
    assert np.array_equal(gj, wj) and np.array_equal(bits(gx), bits(wx))
    assert rx.check_indices_are_unsorted(p, gj) is True


def test_sort_matches_reference_build(rx, refops):
    p, j, x = powerlaw_csr(800, 5000, 25, seed=35, cap=3000)
    js, xs = scramble_rows(p, j, x, 36)
    wj, wx = refops.sort_sparse_indices_numeric(p, js, xs)
    gj, gx = js.copy(), xs.copy()
    rx.sort_sparse_indices_numeric(p, gj, gx)
    assert np.array_equal(gj, wj) and np.array_equal(bits(gx), bits(wx))


@pytest.mark.parametrize("length", [1, 2, 31, 32, 33, 512, 513, 5000, 16384, 16385, 40000, 70000, 150000])
def test_sort_single_row_every_size_class(rx, length):
    # one row (= a sparse vector, R/utils.R:96-118) of each size class: warp network, CTA network, global hybrid
    rng = np.random.default_rng(length)
    j = rng.permutation(4 * length).astype(np.int32)[:length]
    x = rng.standard_normal(length)
    p = np.array([0, length], dtype=np.int32)
    gj, gx = j.copy(), x.copy()
    rx.sort_sparse_indices_numeric(p, gj, gx)
    o = np.argsort(j, kind="stable")
    assert np.array_equal(gj, j[o]) and np.array_equal(bits(gx), bits(x[o]))


def test_sort_duplicates_are_stable_and_pattern(rx):
    rng = np.random.default_rng(41)
    lens = np.array([0, 7, 40, 700, 0, 20000, 3], dtype=np.int32)
    p = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    j = rng.integers(0, 50, p[-1]).astype(np.int32)  # heavy repetition
    x = rng.standard_normal(p[-1])
    gj, gx = j.copy(), x.copy()
    rx.sort_sparse_indices_numeric(p, gj, gx)
    for r in range(lens.size):
        a, b = p[r], p[r + 1]
        o = np.argsort(j[a:b], kind="stable")
        assert np.array_equal(gj[a:b], j[a:b][o]) and np.array_equal(bits(gx[a:b]), bits(x[a:b][o]))
    gj2 = j.copy()
    rx.sort_sparse_indices_binary(p, gj2)
    assert np.array_equal(gj2, gj)


def test_sort_negative_ids_order_as_signed(rx):
    # invalid matrices still sort like the reference's signed comparison (src/misc.cpp:216)
    p = np.array([0, 6], dtype=np.int32)
    j = np.array([3, -1, 7, NA_INT, 0, -5], dtype=np.int32)
    x = np.arange(6, dtype=np.float64)
    rx.sort_sparse_indices_numeric(p, j, x)
    assert np.array_equal(j, np.array([NA_INT, -5, -1, 0, 3, 7], dtype=np.int32))
    assert np.array_equal(x, np.array([3.0, 5.0, 1.0, 4.0, 0.0, 2.0]))


def test_sort_s4_and_check_sparse_matrix(port):
    from matrixextra_b200 import dgRMatrix, sparseVector
    from matrixextra_b200.utils import check_sparse_matrix, sort_sparse_indices
    A = rsparsematrix(150, 120, 0.15, 43)
    p, j, x = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    js, xs = scramble_rows(p, j, x, 44)
    X = dgRMatrix(p, js, xs, A.shape)
    Y = sort_sparse_indices(X, copy=True)
    assert np.array_equal(Y.j, j) and np.array_equal(Y.x, x) and np.array_equal(X.j, js)  # copy left X alone
    Z = check_sparse_matrix(X, sort=True, copy=False)
    assert Z is X and np.array_equal(X.j, j) and np.array_equal(X.x, x)
    bad = dgRMatrix(p, np.where(np.arange(j.size) == 3, 500, j), x, A.shape)
    with pytest.raises(ValueError, match="invalid column indices"):
        check_sparse_matrix(bad)
    v = sparseVector(np.array([9, 2, 5], dtype=np.int32), np.array([1.0, 2.0, 3.0]), 10, "d")
    sort_sparse_indices(v)
    assert np.array_equal(v.i, [2, 5, 9]) and np.array_equal(v.x, [2.0, 3.0, 1.0])


def test_sort_then_svec_product_equals_reference_pipeline(rx, port):
    # the R pipeline of gemv_csr_vec for sparse vectors (R/matmul.R:598-613): sort both operands, then multiply
    A = rsparsematrix(300, 200, 0.1, 45)
    p, j, x = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    js, xs = scramble_rows(p, j, x, 46)
    rng = np.random.default_rng(47)
    yi = (rng.permutation(200)[:50] + 1).astype(np.int32)
    yv = rng.standard_normal(50)
    o = np.argsort(yi)
    want = port.matmul_csr_svec_numeric(p, j, x, yi[o], yv[o])
    gj, gx = js.copy(), xs.copy()
    rx.sort_sparse_indices_numeric(p, gj, gx)
    yi2, yv2 = yi.copy(), yv.copy()
    rx.sort_sparse_indices_numeric(np.array([0, 50], dtype=np.int32), yi2, yv2)
    got = rx.matmul_csr_svec_numeric(p, gj, gx, yi2, yv2, ncols=200)
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))


# ---------------------------------------------------------------------------------------------------
# device-resident forms at a size where every code path is busy (properties, no CPU oracle needed)
# ---------------------------------------------------------------------------------------------------
def test_device_sort_and_mul_at_scale():
    import ctypes as C
    import torch
    from matrixextra_b200 import _lib
    from matrixextra_b200.device import DeviceCSR
    A = DeviceCSR.synth(400_000, 200_000, 12_000_000, seed=1234)  # power-law rows up to 65 536 entries
    nnz, m = A.nnz, A.m

    p_h, j_h, x_h = A.to_host()
    p_t = torch.from_numpy(p_h).cuda()
    j_t = torch.from_numpy(j_h).cuda()
    x_t = torch.from_numpy(x_h).cuda()
    # reverse every row on the device: all rows with >= 2 entries become unsorted
    lens = (p_t[1:] - p_t[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(m, device="cuda"), lens)
    e = torch.arange(nnz, device="cuda")
    src = p_t[:-1].long()[rows] + p_t[1:].long()[rows] - 1 - e
    j_rev, x_rev = j_t[src].contiguous(), x_t[src].contiguous()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flag = C.c_int(-1)
    _lib.call("mxg_dev_rows_sorted", m, C.c_void_p(p_t.data_ptr()), C.c_void_p(j_rev.data_ptr()), C.byref(flag), st)
    assert flag.value == 0
    j_out, x_out = torch.empty_like(j_rev), torch.empty_like(x_rev)
    changed = C.c_int(0)
    _lib.call("mxg_dev_sort_csr_indices", m, C.c_void_p(p_t.data_ptr()), C.c_void_p(j_rev.data_ptr()),
              C.c_void_p(x_rev.data_ptr()), C.c_void_p(j_out.data_ptr()), C.c_void_p(x_out.data_ptr()), C.byref(changed), st)
    torch.cuda.synchronize()
    assert changed.value == int((lens >= 2).sum().item())
    assert torch.equal(j_out, j_t) and torch.equal(x_out.view(torch.int64), x_t.view(torch.int64))
    _lib.call("mxg_dev_rows_sorted", m, C.c_void_p(p_t.data_ptr()), C.c_void_p(j_out.data_ptr()), C.byref(flag), st)
    assert flag.value == 1
    code = C.c_int(-1)
    _lib.call("mxg_dev_check_valid_csr", m, A.K, C.c_void_p(p_t.data_ptr()), C.c_void_p(j_t.data_ptr()), nnz, C.byref(code), st)
    assert code.value == 0
    _lib.call("mxg_dev_check_valid_csr", m, A.K - 1000, C.c_void_p(p_t.data_ptr()), C.c_void_p(j_t.data_ptr()), nnz, C.byref(code), st)
    assert code.value == 2
    # elementwise product with a per-row vector: bit-equal to the torch expression of the same multiply
    v = torch.randn(m, dtype=torch.float64, device="cuda")
    out = torch.empty(nnz, dtype=torch.float64, device="cuda")
    _lib.call("mxg_dev_mul_csr_dvec", A._h, C.c_void_p(v.data_ptr()), m, C.c_void_p(out.data_ptr()), st)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int64), (x_t * v[rows]).view(torch.int64))
