"""Parity at the sizes of BASELINE.json's configs, against the reference's own src/matmul.cpp (oracle/_ref, compiled
in place) — not against the GPU itself:

* cfg1 at full size through the level-1 exports (10k x 5k, 1 % dense, n = 32 fp64);
* the cfg2 / cfg3 / cfg4-size synthetic matrices (the generator bench.py times) through the level-1 exports with
  pageable host arrays, every row of the result compared with the reference (rows are independent: the reference
  runs on all host cores in well under a second per pass);
* a result with MORE THAN 2^31 ELEMENTS (34 M rows x n = 64 fp32, both layouts): the reference's `int` strides
  overflow there (src/matmul.cpp:35-39, 156, 182); the device path uses size_t offsets throughout.  First, last and
  sampled row windows are compared with the reference run on exactly those rows.
Skipped when the GPU has less free memory than a test needs.
"""
import numpy as np
import pytest

from helpers import FP32_TOL, FP64_TOL, rel_err, rsparsematrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rx():
    from matrixextra_b200 import rcpp_exports
    return rcpp_exports


def _need_gpu_gb(gb):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free device memory")


def _synth_host(name):
    """The synthetic matrix bench.py times for this workload, as host arrays."""
    import bench
    from matrixextra_b200._lib import MXG_KEEP_F64
    from matrixextra_b200.device import DeviceCSR
    wl = bench.WORKLOADS[name]
    A = DeviceCSR.synth(wl["m"], wl["K"], wl["nnz"], wl["row_model"], wl["col_model"], seed=wl["seed"], keep=MXG_KEEP_F64)
    p, j, x = A.to_host()
    A.free()
    return wl, p, j, x


def test_cfg1_full_size_level1_vs_reference(rx, ref, port):
    S = rsparsematrix(10_000, 5_000, 0.01, 1001)
    assert S.nnz == 500_000
    rng = np.random.default_rng(1001)
    B = np.asfortranarray(rng.standard_normal((5_000, 32)))
    Yt = np.asfortranarray(B.T)  # `%*%`(Rsparse, matrix) = tcrossprod(x, t(y)), R/matmul.R:463-465
    got = rx.tcrossprod_csr_dense_numeric(S.indptr, S.indices, S.data, Yt)
    assert rel_err(got, ref.tcrossprod_csr_dense_numeric(S.indptr, S.indices, S.data, Yt)) <= FP64_TOL
    assert rel_err(got, port.tcrossprod_csr_dense_numeric(S.indptr, S.indices, S.data, Yt)) <= FP64_TOL
    assert rel_err(got, S @ B) <= 1e-12


def test_cfg2_size_spmv_level1_vs_reference(rx, ref):
    _need_gpu_gb(6)
    wl, p, j, x = _synth_host("cfg2")
    y = np.random.default_rng(2).standard_normal(wl["K"])
    got = rx.matmul_csr_dvec_numeric(p, j, x, y)
    want = ref.matmul_csr_dvec_numeric(p, j, x, y)
    assert got.shape == (wl["m"],) and rel_err(got, want) <= FP64_TOL
    # row by row as well (a global max-norm bound alone would hide a damaged short row)
    scale = np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
    assert np.max(np.abs(got - want) / scale) <= 1e-9


def test_cfg3_size_tcrossprod_level1_vs_reference(rx, ref):
    _need_gpu_gb(8)
    wl, p, j, x = _synth_host("cfg3")
    n, K, m = wl["n"], wl["K"], wl["m"]
    X = np.asfortranarray(np.random.default_rng(3).standard_normal((n, K)).astype(np.float32))
    got = rx.tcrossprod_dense_csr_float32(X, p, j, x, 0, K)  # (n x m) column-major: every output row contiguous
    want = ref.tcrossprod_dense_csr_float32(X, p, j, x, None, K)
    assert got.shape == (n, m) and rel_err(got, want) <= FP32_TOL
    # per output row (column of the R result): relative to that row's own magnitude
    err = np.max(np.abs(got.astype(np.float64) - want), axis=0)
    mag = np.maximum(np.max(np.abs(want), axis=0), 1e-3 * np.abs(want).max())
    assert np.max(err / mag) <= 1e-4
    del got, want
    # and the fp64 column-major product (k64f64 of bench.py) on 10 000 sampled rows of the same matrix
    rows = np.sort(np.random.default_rng(33).choice(m, 10_000, replace=False))
    Y = np.asfortranarray(np.random.default_rng(34).standard_normal((n, K)))
    got64 = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    lens = (p[rows + 1] - p[rows]).astype(np.int64)
    ps = np.zeros(rows.size + 1, dtype=np.int32)
    np.cumsum(lens, out=ps[1:])
    take = np.concatenate([np.arange(p[r], p[r + 1]) for r in rows])
    want64 = ref.tcrossprod_csr_dense_numeric(ps, j[take], x[take], Y)
    assert rel_err(got64[rows], want64) <= FP64_TOL


def test_cfg4_size_crossprod_level1_vs_reference(rx, ref, port):
    _need_gpu_gb(12)
    wl, p, j, x = _synth_host("cfg4")
    n, K, m = wl["n"], wl["K"], wl["m"]
    Y = np.asfortranarray(np.random.default_rng(4).standard_normal((m, n)))
    got = rx.crossprod_csr_dense(p, j, x, K, Y)  # (K x n) column-major
    # the all-MatrixExtra CPU route: stable CSR->CSC, then matmul_dense_csc(t(Y), CSC(A))
    p2, i2, x2 = port.csr2csc(m, K, p, j, x)
    p3, i3, x3 = rx.csr_to_csc(m, K, p, j, x)
    assert np.array_equal(p2, p3) and np.array_equal(i2, i3) and np.array_equal(x2, x3)  # bit-exact at 100 M entries
    want = ref.matmul_dense_csc_numeric(np.asfortranarray(Y.T), p2, i2, x2).T
    assert got.shape == (K, n) and rel_err(got, want) <= FP64_TOL


@pytest.mark.parametrize("layout", ["rows", "cols"])
def test_result_with_more_than_2_31_elements(ref, layout):
    """34 M x 64 float32 = 2.18e9 elements (8.7 GB): every offset into the result exceeds int32 from row 33.5 M on
    (rows-contiguous) or from column 63 on (column-major, stride 34 M)."""
    import torch
    from matrixextra_b200._lib import MXG_COLS_CONTIGUOUS, MXG_F32, MXG_KEEP_F32, MXG_KEEP_F64, MXG_ROWS_CONTIGUOUS
    from matrixextra_b200.device import DeviceCSR
    _need_gpu_gb(14)
    m, K, n = 34_000_000, 100_000, 64
    assert m * n > 2 ** 31
    A = DeviceCSR.synth(m, K, 51_000_000, row_model=0, col_model=0, seed=2031, keep=MXG_KEEP_F32 | MXG_KEEP_F64)
    p, j, x = A.to_host()
    g = torch.Generator(device="cuda").manual_seed(31)
    B = torch.randn(K, n, device="cuda", dtype=torch.float32, generator=g)
    Xh = np.asfortranarray(B.cpu().numpy().T)  # (n x K) column-major
    if layout == "rows":
        out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float32)
        A.spmm(B, out, n, MXG_F32, MXG_ROWS_CONTIGUOUS)
        rows_of = lambda r0, r1: out[r0:r1].cpu().numpy()  # noqa: E731
    else:
        out = torch.full((n, m), float("nan"), device="cuda", dtype=torch.float32)  # column-major m x n, ldc = m
        A.spmm(B, out, n, MXG_F32, MXG_COLS_CONTIGUOUS)
        rows_of = lambda r0, r1: out[:, r0:r1].T.cpu().numpy()  # noqa: E731
    torch.cuda.synchronize()
    rng = np.random.default_rng(7)
    windows = [0, m - 2000, (2 ** 31) // n - 1000, (2 ** 31) // n + 1000] + [int(v) for v in rng.integers(0, m - 2000, 8)]
    for r0 in windows:
        r1 = r0 + 2000
        ps = (p[r0:r1 + 1] - p[r0]).astype(np.int32)
        want = ref.tcrossprod_dense_csr_float32(Xh, ps, j[p[r0]:p[r1]], x[p[r0]:p[r1]], None, K)  # (n x 2000)
        assert rel_err(rows_of(r0, r1), want.T) <= FP32_TOL, r0
    # nothing was left unwritten anywhere (NaN-filled before the product), and empty rows are exact zeros
    assert not torch.isnan(out).any().item()
    empty = np.flatnonzero(np.diff(p[:200_001]) == 0)
    if empty.size:
        assert not rows_of(int(empty[0]), int(empty[0]) + 1).any()
    A.free()


def test_host_twin_generates_the_device_generators_matrix():
    """bench.py's reference arm multiplies the matrix of oracle/mx_synth.c; the GPU arm that of csrc/synth.cu.  Same
    recipe, same counters: identical row lengths except where pow() differs in the last ulp, identical columns and values
    wherever the row lengths agree."""
    from matrixextra_b200._lib import MXG_KEEP_F64
    from matrixextra_b200.device import DeviceCSR
    from oracle.cpu_oracle import synth_csr_host
    for row_model, col_model, m, K, nnz, seed in ((1, 1, 200_000, 100_000, 10_000_000, 1003), (0, 0, 10_000, 5_000, 500_000, 1001)):
        A = DeviceCSR.synth(m, K, nnz, row_model, col_model, seed=seed, keep=MXG_KEEP_F64)
        pd, jd, xd = A.to_host()
        A.free()
        ph, jh, xh = synth_csr_host(m, K, nnz, row_model, col_model, seed)
        assert abs(int(pd[-1]) - int(ph[-1])) <= 64
        same_len = np.diff(pd) == np.diff(ph)
        assert same_len.mean() > 0.999
        if np.array_equal(pd, ph):
            assert np.array_equal(jd, jh) and np.array_equal(xd, xh)
        else:  # compare the rows whose lengths agree
            rows = np.flatnonzero(same_len)[:20000]
            for r in rows[:: max(1, rows.size // 2000)]:
                assert np.array_equal(jd[pd[r]:pd[r + 1]], jh[ph[r]:ph[r + 1]])
                assert np.array_equal(xd[pd[r]:pd[r + 1]], xh[ph[r]:ph[r + 1]])
