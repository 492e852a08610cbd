"""The Rcpp glue (rglue/matmul_gpu_glue.cpp) without R: compiled unmodified against the Rcpp stand-in of the
test infrastructure (oracle/shim/Rcpp.h) together with tests/glue_driver.cpp, linked with libmxgpu.so, and
driven through ctypes with the argument layouts R would pass (column-major, float32 as int bits).

CPU: the glue builds, exports what the driver needs, and turns a failed C-ABI call into an R-style error.
GPU: every exported entry point of the glue against the CPU oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import FP32_TOL, FP64_TOL, NA_INT, powerlaw_csr, rel_err, rsparsematrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
LIB = os.path.join(BUILD, "libgluedrv.so")


def _build():
    from matrixextra_b200 import build_native
    build_native.build()
    srcs = [os.path.join(ROOT, "tests", "glue_driver.cpp"), os.path.join(ROOT, "rglue", "matmul_gpu_glue.cpp"),
            os.path.join(ROOT, "rglue", "rowops_gpu_glue.cpp"), os.path.join(ROOT, "rglue", "handle_gpu_glue.cpp"),
            os.path.join(ROOT, "rglue", "mxgpu_result_alloc.h"),
            os.path.join(ROOT, "include", "mxgpu.h"), os.path.join(ROOT, "oracle", "shim", "Rcpp.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DMXGPU_GLUE_SHIM",
           "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "rglue"),
           "-o", LIB, srcs[0],
           "-L" + os.path.join(ROOT, "matrixextra_b200", "csrc"), "-lmxgpu",
           "-Wl,-rpath," + os.path.join(ROOT, "matrixextra_b200", "csrc")]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env)
    return LIB


@pytest.fixture(scope="module")
def drv():
    lib = C.CDLL(_build())
    lib.gluedrv_last_error.restype = C.c_char_p
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def test_glue_builds_and_reports_errors_like_r(drv):
    for name in ("gluedrv_dense_sparse", "gluedrv_sparse_tdense", "gluedrv_csr_dvec", "gluedrv_crossprod",
                 "gluedrv_csr_to_csc"):
        assert hasattr(drv, name)
    # the ten export names of src/matmul.cpp:221-483 are all defined by the glue, with Rcpp::export tags
    text = open(os.path.join(ROOT, "rglue", "matmul_gpu_glue.cpp")).read()
    for name in ("matmul_dense_csc_numeric", "matmul_dense_csc_float32", "tcrossprod_dense_csr_numeric",
                 "tcrossprod_dense_csr_float32", "tcrossprod_csr_dense_numeric", "tcrossprod_csr_dense_float32",
                 "matmul_csr_dvec_numeric", "matmul_csr_dvec_integer", "matmul_csr_dvec_logical",
                 "matmul_csr_dvec_float32"):
        assert text.count(name + "(") >= 1
    assert text.count("// [[Rcpp::export(rng = false)]]") >= 10
    # an invalid matrix (decreasing indptr) or a missing device surfaces as an R error with the library's message
    p = _i32([0, 2, 1])
    j = _i32([0, 0])
    x = np.array([1.0, 2.0])
    y = np.array([3.0])
    out = np.zeros(2)
    rc = drv.gluedrv_csr_dvec(0, _p(p), 2, _p(j), _p(x), 2, _p(y), 1, _p(out))
    assert rc == 1
    msg = drv.gluedrv_last_error().decode()
    assert ("indptr" in msg) or ("cuda" in msg.lower()), msg


@pytest.mark.gpu
@pytest.mark.parametrize("f32", [0, 1])
def test_glue_products_match_oracle(drv, port, f32):
    dt = np.float32 if f32 else np.float64
    tol = FP32_TOL if f32 else FP64_TOL
    sfx = "float32" if f32 else "numeric"
    rng = np.random.default_rng(41)
    a, K, b = 37, 120, 90
    S = rsparsematrix(b, K, 0.2, 41)  # rows of S: CSR of S == CSC of t(S)
    p, j, x = _i32(S.indptr), _i32(S.indices), S.data
    X = np.asfortranarray(rng.standard_normal((a, K)).astype(dt))
    for which, name in ((0, "matmul_dense_csc_"), (1, "tcrossprod_dense_csr_")):
        out = np.empty((a, b), dtype=dt, order="F")
        assert drv.gluedrv_dense_sparse(which, f32, _p(X), a, K, _p(p), b, _p(j), _p(x), j.size, _p(out)) == 0
        want = getattr(port, name + sfx)(X, p, j, x, 1) if which == 0 else getattr(port, name + sfx)(X, p, j, x, 1, K)
        assert rel_err(out, want) <= tol
    Y = np.asfortranarray(rng.standard_normal((24, K)).astype(dt))
    out = np.empty((b, 24), dtype=dt, order="F")
    assert drv.gluedrv_sparse_tdense(f32, _p(p), b, _p(j), _p(x), j.size, _p(Y), 24, K, _p(out)) == 0
    assert rel_err(out, getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, Y, 1)) <= tol
    # crossprod(CSR(b x K), dense(b x n)) -> K x n  (new method on the device transpose)
    Z = np.asfortranarray(rng.standard_normal((b, 12)).astype(dt))
    out = np.empty((K, 12), dtype=dt, order="F")
    assert drv.gluedrv_crossprod(f32, _p(p), b, _p(j), _p(x), j.size, K, _p(Z), b, 12, _p(out)) == 0
    p2, i2, x2 = port.csr2csc(b, K, p, j, x)
    want = getattr(port, "matmul_dense_csc_" + sfx)(np.asfortranarray(Z.T), p2, i2, x2).T
    assert rel_err(out, want) <= tol
    # dimension mismatch -> the reference's message
    assert drv.gluedrv_crossprod(f32, _p(p), b, _p(j), _p(x), j.size, K, _p(Z), b - 1, 12, _p(out)) == 1
    assert drv.gluedrv_last_error().decode() == "Matrix dimensions do not match."


@pytest.mark.gpu
def test_glue_vectors_and_transpose(drv, port):
    p, j, x = powerlaw_csr(900, 300, 15, seed=42, cap=280)
    rng = np.random.default_rng(42)
    y = rng.standard_normal(300)
    out = np.empty(900)
    assert drv.gluedrv_csr_dvec(0, _p(p), 900, _p(j), _p(x), j.size, _p(y), 300, _p(out)) == 0
    assert rel_err(out, port.matmul_csr_dvec_numeric(p, j, x, y, 1)) <= FP64_TOL
    yi = rng.integers(-3, 4, 300).astype(np.int32)
    yi[::11] = NA_INT
    for ytype, fn in ((1, port.matmul_csr_dvec_integer), (2, port.matmul_csr_dvec_logical)):
        assert drv.gluedrv_csr_dvec(ytype, _p(p), 900, _p(j), _p(x), j.size, _p(yi), 300, _p(out)) == 0
        want = fn(p, j, x, yi, 1)
        assert np.array_equal(np.isnan(out), np.isnan(want))
        ok = ~np.isnan(want)
        assert rel_err(out[ok], want[ok]) <= FP64_TOL
    yf = rng.standard_normal(300).astype(np.float32)
    outf = np.empty(900, dtype=np.float32)
    assert drv.gluedrv_csr_dvec(3, _p(p), 900, _p(j), _p(x), j.size, _p(yf), 300, _p(outf)) == 0
    assert rel_err(outf, port.matmul_csr_dvec_float32(p, j, x, yf, 1)) <= FP32_TOL
    p2, i2, x2 = np.empty(301, np.int32), np.empty(j.size, np.int32), np.empty(j.size)
    assert drv.gluedrv_csr_to_csc(_p(p), 900, _p(j), _p(x), j.size, 300, _p(p2), _p(i2), _p(x2)) == 0
    q2, k2, y2 = port.csr2csc(900, 300, p, j, x)
    assert np.array_equal(p2, q2) and np.array_equal(i2, k2) and np.array_equal(x2, y2)
    # out-of-range column id: an R error, not a fault
    jb = j.copy()
    jb[7] = 300
    assert drv.gluedrv_csr_dvec(0, _p(p), 900, _p(jb), _p(x), j.size, _p(y), 300, _p(out)) == 1
    assert "column index" in drv.gluedrv_last_error().decode()


# ---- rglue/rowops_gpu_glue.cpp: the exports either side of the product ----------------
def test_rowops_glue_keeps_the_reference_export_names():
    text = open(os.path.join(ROOT, "rglue", "rowops_gpu_glue.cpp")).read()
    for name in ("matmul_csr_svec_numeric", "matmul_csr_svec_integer", "matmul_csr_svec_logical", "matmul_csr_svec_binary",
                 "matmul_csr_svec_float32", "matmul_rowvec_by_csc", "matmul_rowvec_by_cscbin", "check_indices_are_unsorted", "sort_sparse_indices_numeric",
                 "sort_sparse_indices_binary", "check_valid_csr_matrix", "multiply_csr_by_dense_elemwise_double",
                 "multiply_csr_by_dense_elemwise_float32", "multiply_csr_by_dense_elemwise_int",
                 "multiply_csr_by_dense_elemwise_bool", "multiply_csr_by_dvec_no_NAs_numeric"):
        assert text.count(name + "(") >= 1, name
    assert text.count("\n// [[Rcpp::export(rng = false)]]\n") == 16


@pytest.mark.gpu
def test_rowops_glue_sparse_vector_products(drv, port):
    p, j, x = powerlaw_csr(700, 400, 12, seed=43, cap=380)
    rng = np.random.default_rng(43)
    yi = (np.sort(rng.choice(400, size=90, replace=False)) + 1).astype(np.int32)
    out = np.empty(700)
    yv = rng.standard_normal(90)
    assert drv.gluedrv_csr_svec(0, _p(p), 700, _p(j), _p(x), j.size, _p(yi), 90, _p(yv), _p(out)) == 0
    assert rel_err(out, port.matmul_csr_svec_numeric(p, j, x, yi, yv)) <= FP64_TOL
    iv = rng.integers(-4, 5, 90).astype(np.int32)
    iv[::7] = NA_INT
    for ytype, fn in ((1, port.matmul_csr_svec_integer), (2, port.matmul_csr_svec_logical)):
        assert drv.gluedrv_csr_svec(ytype, _p(p), 700, _p(j), _p(x), j.size, _p(yi), 90, _p(iv), _p(out)) == 0
        want = fn(p, j, x, yi, iv)
        assert np.array_equal(np.isnan(out), np.isnan(want))
        ok = ~np.isnan(want)
        assert rel_err(out[ok], want[ok]) <= FP64_TOL
    fv = rng.standard_normal(90).astype(np.float32)
    assert drv.gluedrv_csr_svec(3, _p(p), 700, _p(j), _p(x), j.size, _p(yi), 90, _p(fv), _p(out)) == 0
    assert rel_err(out, port.matmul_csr_svec_float32(p, j, x, yi, fv)) <= FP64_TOL
    assert drv.gluedrv_csr_svec(4, _p(p), 700, _p(j), _p(x), j.size, _p(yi), 90, None, _p(out)) == 0
    assert rel_err(out, port.matmul_csr_svec_binary(p, j, x, yi)) <= FP64_TOL


@pytest.mark.gpu
def test_rowops_glue_sort_validity_and_elementwise(drv, port):
    p, j, x = powerlaw_csr(600, 500, 14, seed=44, cap=450)
    rng = np.random.default_rng(44)
    ju, xu = j.copy(), x.copy()
    for r in range(0, 600, 3):  # shuffle every third row
        a, b = p[r], p[r + 1]
        perm = rng.permutation(b - a)
        ju[a:b], xu[a:b] = ju[a:b][perm], xu[a:b][perm]
    s = C.c_int(-1)
    assert drv.gluedrv_rows_sorted(_p(p), 600, _p(j), j.size, C.byref(s)) == 0 and s.value == 1
    assert drv.gluedrv_rows_sorted(_p(p), 600, _p(ju), j.size, C.byref(s)) == 0 and s.value == 0
    js, xs = ju.copy(), xu.copy()
    assert drv.gluedrv_sort_indices(_p(p), 600, _p(js), _p(xs), j.size) == 0  # in place, values follow
    assert np.array_equal(js, j) and np.array_equal(xs, x)
    js = ju.copy()
    assert drv.gluedrv_sort_indices(_p(p), 600, _p(js), None, j.size) == 0    # pattern matrix
    assert np.array_equal(js, j)
    # validity: the reference's messages, first failing check first (src/misc.cpp:970-1016)
    msg = C.create_string_buffer(256)
    assert drv.gluedrv_check_valid(_p(p), p.size, _p(j), j.size, 600, 500, msg, 256) == 0 and msg.value == b""
    jb = j.copy()
    jb[5] = -1
    assert drv.gluedrv_check_valid(_p(p), p.size, _p(jb), j.size, 600, 500, msg, 256) == 0
    assert msg.value.decode() == "Matrix has negative indices."
    assert msg.value.decode() == port.check_valid_csr_matrix(p, jb, 600, 500)
    jb[5] = 500
    assert drv.gluedrv_check_valid(_p(p), p.size, _p(jb), j.size, 600, 500, msg, 256) == 0
    assert msg.value.decode() == "Matrix has invalid column indices."
    pb = p.copy()
    pb[10], pb[11] = pb[11], pb[10]
    if pb[10] > pb[11]:
        assert drv.gluedrv_check_valid(_p(pb), pb.size, _p(j), j.size, 600, 500, msg, 256) == 0
        assert msg.value.decode() == "Matrix index pointer is not monotonicaly increasing."
    # elementwise products: one multiply per entry, bit-exact
    D = np.asfortranarray(rng.standard_normal((600, 500)))
    out = np.empty(j.size)
    assert drv.gluedrv_mul_dense(0, _p(p), 600, _p(j), _p(x), j.size, _p(D), C.c_long(D.size), _p(out)) == 0
    assert np.array_equal(out, port.multiply_csr_by_dense_elemwise_double(p, j, x, D.ravel(order="F")))
    Df = np.asfortranarray(D.astype(np.float32))
    assert drv.gluedrv_mul_dense(3, _p(p), 600, _p(j), _p(x), j.size, _p(Df), C.c_long(Df.size), _p(out)) == 0
    assert np.array_equal(out, port.multiply_csr_by_dense_elemwise_float32(p, j, x, Df.ravel(order="F")))
    Di = np.asfortranarray(rng.integers(-3, 4, (600, 500)).astype(np.int32))
    Di[::17, ::13] = NA_INT
    for dtype, fn in ((1, port.multiply_csr_by_dense_elemwise_int), (2, port.multiply_csr_by_dense_elemwise_bool)):
        assert drv.gluedrv_mul_dense(dtype, _p(p), 600, _p(j), _p(x), j.size, _p(Di), C.c_long(Di.size), _p(out)) == 0
        want = fn(p, j, x, Di.ravel(order="F"))
        assert np.array_equal(np.isnan(out), np.isnan(want))
        assert np.array_equal(out[~np.isnan(want)], want[~np.isnan(want)])
    v = rng.standard_normal(600)  # one value per row, recycled along the columns (R/operators.R:236-397)
    assert drv.gluedrv_mul_dvec(_p(p), 600, _p(j), _p(x), j.size, _p(v), C.c_long(v.size), 500, 1, _p(out)) == 0
    assert np.array_equal(out, port.multiply_csr_by_dvec_no_NAs_numeric(p, j, x, v, 500))
    assert drv.gluedrv_mul_dvec(_p(p), 600, _p(j), _p(x), j.size, _p(v), C.c_long(v.size), 500, 0, _p(out)) == 1
    assert "only the multiplication" in drv.gluedrv_last_error().decode()


@pytest.mark.gpu
def test_rowops_glue_float32_row_vector_by_csc(drv, port):
    Y = rsparsematrix(400, 150, 0.1, 45, "csc")
    p, i, x = _i32(Y.indptr), _i32(Y.indices), Y.data
    rv = np.random.default_rng(45).standard_normal(400).astype(np.float32)
    out = np.empty(150, dtype=np.float32)
    assert drv.gluedrv_rowvec_by_csc(_p(rv), 400, _p(p), 150, _p(i), _p(x), i.size, _p(out)) == 0
    assert rel_err(out, port.matmul_rowvec_by_csc(rv, p, i, x).ravel()) <= FP32_TOL
    assert drv.gluedrv_rowvec_by_csc(_p(rv), 400, _p(p), 150, _p(i), None, i.size, _p(out)) == 0
    assert rel_err(out, port.matmul_rowvec_by_csc(rv, p, i, None).ravel()) <= FP32_TOL


# ---- rglue/handle_gpu_glue.cpp: device-resident matrices behind an external pointer ------------------
def test_handle_glue_exports_and_finalizer_wiring():
    text = open(os.path.join(ROOT, "rglue", "handle_gpu_glue.cpp")).read()
    for name in ("as_gpu_csr", "gpu_csr_free", "gpu_csr_dim", "gpu_csr_tcrossprod_dense_numeric", "gpu_csr_tcrossprod_dense_float32",
                 "gpu_csr_dense_tcrossprod_numeric", "gpu_csr_dense_tcrossprod_float32", "gpu_csr_crossprod_dense_numeric",
                 "gpu_csr_crossprod_dense_float32", "gpu_csr_dvec_numeric", "mxgpu_configure"):
        assert text.count(name + "(") >= 1, name
    assert text.count("\n// [[Rcpp::export(rng = false)]]\n") == 11
    assert "Rcpp::XPtr<MxGpuCsr, Rcpp::PreserveStorage, mxgpu_csr_finalizer, true>" in text  # mxg_csr_free runs on GC and at exit


@pytest.mark.gpu
@pytest.mark.parametrize("f32", [0, 1])
def test_handle_glue_products_match_the_level1_exports(drv, port, f32):
    dt = np.float32 if f32 else np.float64
    tol = FP32_TOL if f32 else FP64_TOL
    sfx = "float32" if f32 else "numeric"
    m, K = 900, 350
    p, j, x = powerlaw_csr(m, K, 14, seed=46, cap=340)
    rng = np.random.default_rng(46)
    drv.gluedrv_as_gpu_csr.restype = C.c_void_p
    box = drv.gluedrv_as_gpu_csr(_p(p), m, _p(j), _p(x), j.size, K, 1, 1)
    assert box, drv.gluedrv_last_error()
    box = C.c_void_p(box)
    try:
        Y = np.asfortranarray(rng.standard_normal((24, K)).astype(dt))  # A %*% t(Y)
        out = np.empty((m, 24), dtype=dt, order="F")
        assert drv.gluedrv_gpu_csr_product(box, 0, f32, _p(Y), 24, K, _p(out)) == 0
        assert rel_err(out, getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, Y, 1)) <= tol
        ref_out = np.empty((m, 24), dtype=dt, order="F")
        assert drv.gluedrv_sparse_tdense(f32, _p(p), m, _p(j), _p(x), j.size, _p(Y), 24, K, _p(ref_out)) == 0
        assert np.array_equal(out, ref_out)  # same bits as the level-1 export
        X = np.asfortranarray(rng.standard_normal((17, K)).astype(dt))  # X %*% t(A)
        out2 = np.empty((17, m), dtype=dt, order="F")
        assert drv.gluedrv_gpu_csr_product(box, 1, f32, _p(X), 17, K, _p(out2)) == 0
        assert rel_err(out2, getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, K)) <= tol
        Z = np.asfortranarray(rng.standard_normal((m, 9)).astype(dt))  # t(A) %*% Z, twice (the second uses the kept CSC)
        out3, out3b = np.empty((K, 9), dtype=dt, order="F"), np.empty((K, 9), dtype=dt, order="F")
        assert drv.gluedrv_gpu_csr_product(box, 2, f32, _p(Z), m, 9, _p(out3)) == 0
        assert drv.gluedrv_gpu_csr_product(box, 2, f32, _p(Z), m, 9, _p(out3b)) == 0
        p2, i2, x2 = port.csr2csc(m, K, p, j, x)
        want = getattr(port, "matmul_dense_csc_" + sfx)(np.asfortranarray(Z.T), p2, i2, x2).T
        assert rel_err(out3, want) <= tol and np.array_equal(out3, out3b)
        if not f32:
            y = rng.standard_normal(K)
            outv = np.empty(m)
            assert drv.gluedrv_gpu_csr_dvec(box, _p(y), K, _p(outv)) == 0
            assert rel_err(outv, port.matmul_csr_dvec_numeric(p, j, x, y, 1)) <= FP64_TOL
            assert drv.gluedrv_gpu_csr_dvec(box, _p(y), K - 1, _p(outv)) == 1
            assert drv.gluedrv_last_error().decode() == "Matrix dimensions do not match."
        # dimension errors carry the reference's message
        assert drv.gluedrv_gpu_csr_product(box, 0, f32, _p(Y), 24, K - 1, _p(out)) == 1
        assert drv.gluedrv_last_error().decode() == "Matrix dimensions do not match."
        # explicit free (gpu.free(x) in R): later products raise instead of touching freed device memory
        assert drv.gluedrv_gpu_csr_free(box) == 0
        assert drv.gluedrv_gpu_csr_product(box, 0, f32, _p(Y), 24, K, _p(out)) == 1
        assert "freed" in drv.gluedrv_last_error().decode()
    finally:
        drv.gluedrv_gpu_csr_drop(box)  # the R object is collected: the finalizer finds nothing left to release
    # an invalid matrix is rejected at construction with the library's message
    jb = j.copy()
    jb[3] = K
    assert not drv.gluedrv_as_gpu_csr(_p(p), m, _p(jb), _p(x), j.size, K, 1, 0)
    assert "column index" in drv.gluedrv_last_error().decode()
    n = C.c_int(0)
    assert drv.gluedrv_configure(1, 0, C.byref(n)) == 0 and n.value == 1
