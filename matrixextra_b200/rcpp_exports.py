"""Python mirror of the ten Rcpp exports on the multiplication path, with the reference's names,
argument order and array layouts (src/matmul.cpp:221-483; R wrappers R/RcppExports.R:132-170),
implemented on the CUDA library through the level-1 (host buffers) C ABI.

This file plays the role of rglue/matmul_gpu_glue.cpp (the real Rcpp glue, which cannot be executed
in an image without R): same dimension inference from the argument shapes, same zero-copy borrowing
of inputs, a freshly allocated column-major output, errors raised from mxg_last_error().
``nthreads`` (the reference's OpenMP team size) sets the number of HOST threads of the library's staging engine
(csrc/hoststage.cu: narrowing float64 values, bouncing pageable memory through page-locked slots); 0 = all
logical CPUs.  ``ncols_Y`` is accepted and ignored, as it already is in the reference (src/matmul.cpp:259).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MXG_COLS_CONTIGUOUS, MXG_F32, MXG_F64, MXG_ROWS_CONTIGUOUS, MXG_Y_BINARY, MXG_Y_FLOAT32,
                   MXG_Y_INTEGER, MXG_Y_LOGICAL, MXG_Y_NUMERIC)


def _vp(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _threads(nthreads) -> None:
    _lib.set_option("host_threads", max(0, int(nthreads)))


def _csr(p, j, x):
    return (np.ascontiguousarray(p, dtype=np.int32), np.ascontiguousarray(j, dtype=np.int32),
            np.ascontiguousarray(x, dtype=np.float64))


def _fmat(a, dtype):
    a = np.asarray(a, dtype=dtype)
    if a.ndim != 2:
        raise ValueError("expected a matrix")
    return np.asfortranarray(a)


#: results of at least this many bytes come from the library's page-locked pool, like the glue's (rglue/mxgpu_result_alloc.h)
PINNED_RESULT_MIN = 1 << 20


def _pooled_empty(shape, np_t):
    """What the glue does through Rf_allocVector3 + R_allocator_t: the result's memory is a block of the library's
    page-locked pool (``mxg_host_alloc``), handed back (``mxg_host_free``) when the array is garbage-collected.  Falls
    back to an ordinary array when the pool is full or there is no device."""
    import weakref
    nbytes = int(np.prod(shape)) * np.dtype(np_t).itemsize
    if nbytes < PINNED_RESULT_MIN:
        return None
    ptr = C.c_void_p()
    if _lib.load().mxg_host_alloc(nbytes, C.byref(ptr)) != _lib.MXG_OK or not ptr.value:
        return None
    buf = (C.c_char * nbytes).from_address(ptr.value)
    weakref.finalize(buf, _lib.load().mxg_host_free, C.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=np_t).reshape(shape, order="F")


def _result(shape, np_t, out):
    """The freshly allocated column-major result (what Rcpp returns), or the caller's buffer: ``out=`` is an
    addition of this mirror so that a caller can hand in memory of its own (a page-locked buffer it reuses, or an
    ordinary ``np.empty`` to see what a result on the caller's heap costs)."""
    if out is None:
        pooled = _pooled_empty(shape, np_t)
        return pooled if pooled is not None else np.empty(shape, dtype=np_t, order="F")
    if out.shape != tuple(shape) or out.dtype != np_t or not out.flags.f_contiguous or not out.flags.writeable:
        raise ValueError("out= must be a writeable column-major array of the result's shape and type")
    return out


def _dense_times_tcsr(X_colmajor, indptr, indices, values, dtype, out=None):
    """Out(a x rows, column-major) = X(a x K) . t(S) where S is given by rows (CSR) — the kernel call of
    matmul_dense_csc / tcrossprod_dense_csr (src/matmul.cpp:188-281): a rows-contiguous product with
    n = nrow(X), ldb = ldc = nrow(X)."""
    np_t = np.float64 if dtype == MXG_F64 else np.float32
    X = _fmat(X_colmajor, np_t)
    p, j, x = _csr(indptr, indices, values)
    a, K = X.shape
    rows = p.size - 1
    out = _result((a, rows), np_t, out)
    _lib.call("mxg_spmm_csr_dense", dtype, MXG_ROWS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, rows, K, a,
              _vp(p), _vp(j), _vp(x), _vp(X), max(a, 1), _vp(out), max(a, 1))
    return out


def matmul_dense_csc_numeric(X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values, nthreads=0, out=None):
    _threads(nthreads)
    return _dense_times_tcsr(X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values, MXG_F64, out)


def matmul_dense_csc_float32(X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values, nthreads=0, out=None):
    _threads(nthreads)
    return _dense_times_tcsr(X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values, MXG_F32, out)


def tcrossprod_dense_csr_numeric(X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values, nthreads=0, ncols_Y=0, out=None):
    _threads(nthreads)
    return _dense_times_tcsr(X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values, MXG_F64, out)


def tcrossprod_dense_csr_float32(X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values, nthreads=0, ncols_Y=0, out=None):
    _threads(nthreads)
    return _dense_times_tcsr(X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values, MXG_F32, out)


def _csr_times_tdense(indptr, indices, values, Y_colmajor, dtype, out=None):
    """Out(m x n, column-major) = A_csr(m x K) . t(Y), Y (n x K) column-major — tcrossprod_csr_dense
    (src/matmul.cpp:316-343): column-major output with ldb = nrow(Y), ldc = m."""
    np_t = np.float64 if dtype == MXG_F64 else np.float32
    Y = _fmat(Y_colmajor, np_t)
    p, j, x = _csr(indptr, indices, values)
    n, K = Y.shape
    m = p.size - 1
    out = _result((m, n), np_t, out)
    _lib.call("mxg_spmm_csr_dense", dtype, MXG_COLS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, m, K, n,
              _vp(p), _vp(j), _vp(x), _vp(Y), max(n, 1), _vp(out), max(m, 1))
    return out


def tcrossprod_csr_dense_numeric(X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor, nthreads=0, out=None):
    _threads(nthreads)
    return _csr_times_tdense(X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor, MXG_F64, out)


def tcrossprod_csr_dense_float32(X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor, nthreads=0, out=None):
    _threads(nthreads)
    return _csr_times_tdense(X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor, MXG_F32, out)


def _csr_dvec(indptr, indices, values, y, ytype, y_np, out_np, out=None):
    p, j, x = _csr(indptr, indices, values)
    y = np.ascontiguousarray(y, dtype=y_np)
    m = p.size - 1
    out = _result((m,), out_np, out)
    # the reference takes K from nowhere (it trusts the indices); here K = length(y) bounds them
    _lib.call("mxg_spmv_csr", ytype, m, int(y.size), _vp(p), _vp(j), _vp(x), _vp(y), _vp(out))
    return out


def matmul_csr_dvec_numeric(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, nthreads=0, out=None):
    _threads(nthreads)
    return _csr_dvec(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, MXG_Y_NUMERIC, np.float64, np.float64, out)


def matmul_csr_dvec_integer(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, nthreads=0, out=None):
    _threads(nthreads)
    return _csr_dvec(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, MXG_Y_INTEGER, np.int32, np.float64, out)


def matmul_csr_dvec_logical(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, nthreads=0, out=None):
    _threads(nthreads)
    return _csr_dvec(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, MXG_Y_LOGICAL, np.int32, np.float64, out)


def matmul_csr_dvec_float32(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, nthreads=0, out=None):
    _threads(nthreads)
    return _csr_dvec(X_csr_indptr, X_csr_indices, X_csr_values, y_dense, MXG_Y_FLOAT32, np.float32, np.float32, out)


# ---- float32 (row) vector %*% CSC (src/matmul.cpp:643-684; R/matmul.R:243-259, 350-366) ---------------------------

def matmul_rowvec_by_csc(rowvec_, indptr, indices, values):
    """out(1 x ncol, float32) = rowvec . Y for a CSC matrix: the float32 SpMV with the CSC arrays read as the CSR of
    t(Y).  ``rowvec_`` holds float32 values (the reference passes them as int bits, float32@Data)."""
    y = np.ascontiguousarray(rowvec_, dtype=np.float32)
    return _csr_dvec(indptr, indices, values, y, MXG_Y_FLOAT32, np.float32, np.float32).reshape(1, -1)


def matmul_rowvec_by_cscbin(rowvec_, indptr, indices):
    """Pattern matrix: every stored entry counts as 1 (src/matmul.cpp:664-684)."""
    return matmul_rowvec_by_csc(rowvec_, indptr, indices, np.ones(np.asarray(indices).size, dtype=np.float64))


# ---- CSR %*% sparseVector (src/matmul.cpp:553-641) ---------------------------------

def _csr_svec(indptr, indices, values, y_indices_base1, y_values, ytype, y_np, ncols=0, out=None):
    p, j, x = _csr(indptr, indices, values)
    yi = np.ascontiguousarray(y_indices_base1, dtype=np.int32)
    yv = None if y_values is None else np.ascontiguousarray(y_values, dtype=y_np)
    if yv is not None and yv.size != yi.size:
        raise ValueError("sparse vector: indices and values differ in length")
    m = p.size - 1
    out = _result((m,), np.float64, out)
    # the reference's exports do not receive ncol(X); ncols=0 lets the library bound the columns by max(y@i)
    _lib.call("mxg_spmv_csr_svec", ytype, m, int(ncols), _vp(p), _vp(j), _vp(x), int(yi.size), _vp(yi),
              _vp(yv) if yv is not None else None, _vp(out))
    return out


def matmul_csr_svec_numeric(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, nthreads=0, ncols=0, out=None):
    _threads(nthreads)
    return _csr_svec(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, MXG_Y_NUMERIC, np.float64, ncols, out)


def matmul_csr_svec_integer(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, nthreads=0, ncols=0, out=None):
    _threads(nthreads)
    return _csr_svec(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, MXG_Y_INTEGER, np.int32, ncols, out)


def matmul_csr_svec_logical(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, nthreads=0, ncols=0, out=None):
    _threads(nthreads)
    return _csr_svec(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, MXG_Y_LOGICAL, np.int32, ncols, out)


def matmul_csr_svec_binary(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, nthreads=0, ncols=0, out=None):
    _threads(nthreads)
    return _csr_svec(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, None, MXG_Y_BINARY, np.int32, ncols, out)


def matmul_csr_svec_float32(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, nthreads=0, ncols=0, out=None):
    _threads(nthreads)
    """Exported by the reference but never called from R (src/matmul.cpp:626-641); double result."""
    return _csr_svec(X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, y_values, MXG_Y_FLOAT32, np.float32, ncols, out)


# ---- index sorting and validity checks (src/misc.cpp:161-330, 970-1016) -----------

def sort_sparse_indices_numeric(indptr, indices, values):
    """Sorts ``indices`` / ``values`` IN PLACE row by row, like the reference (src/misc.cpp:300-313): the arrays
    must be writeable contiguous int32 / float64 (as R's are)."""
    p = np.ascontiguousarray(indptr, dtype=np.int32)
    for a, dt in ((indices, np.int32), (values, np.float64)):
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.flags.writeable):
            raise ValueError("sort_sparse_indices_numeric sorts in place: pass writeable contiguous int32 / float64 arrays")
    _lib.call("mxg_sort_csr_indices", p.size - 1, _vp(p), _vp(indices), _vp(values))


def sort_sparse_indices_binary(indptr, indices):
    """Pattern overload (src/misc.cpp:230-252, export 332-343)."""
    p = np.ascontiguousarray(indptr, dtype=np.int32)
    if not (isinstance(indices, np.ndarray) and indices.dtype == np.int32 and indices.flags.c_contiguous and indices.flags.writeable):
        raise ValueError("sort_sparse_indices_binary sorts in place: pass a writeable contiguous int32 array")
    _lib.call("mxg_sort_csr_indices", p.size - 1, _vp(p), _vp(indices), None)


def check_indices_are_unsorted(indptr, indices) -> bool:
    """True when every row is sorted — the reference's (misleading) name and meaning, src/misc.cpp:161-175."""
    p = np.ascontiguousarray(indptr, dtype=np.int32)
    j = np.ascontiguousarray(indices, dtype=np.int32)
    flag = C.c_int(1)
    _lib.call("mxg_rows_sorted", p.size - 1, _vp(p), _vp(j), C.byref(flag))
    return bool(flag.value)


def check_valid_csr_matrix(indptr, indices, nrows, ncols) -> dict:
    """``{}`` for a valid matrix, else ``{"err": <the reference's message>}`` (src/misc.cpp:970-1016)."""
    p = np.ascontiguousarray(indptr, dtype=np.int32)
    j = np.ascontiguousarray(indices, dtype=np.int32)
    if p.size != int(nrows) + 1:
        raise ValueError("indptr must have nrows + 1 entries")
    code = C.c_int(0)
    _lib.call("mxg_check_valid_csr", int(nrows), int(ncols), _vp(p), _vp(j), int(j.size), C.byref(code))
    if code.value == 0:
        return {}
    return {"err": _lib.load().mxg_csr_error_string(code.value).decode()}


# ---- elementwise CSR * dense (src/operators.cpp:239-314, 2147-2178) ------------------

def _mul_dense(indptr, indices, values, dense_mat, dtype, np_t):
    p, j, x = _csr(indptr, indices, values)
    m = p.size - 1
    d = np.asarray(dense_mat, dtype=np_t)
    if d.ndim == 2:
        if d.shape[0] != m:
            raise ValueError("dense matrix must have as many rows as the CSR matrix")
        K = d.shape[1]
        d = np.asfortranarray(d)
    else:  # the reference receives the matrix as a flat column-major vector
        if m == 0 or d.size % m:
            raise ValueError("dense_mat length is not a multiple of the number of rows")
        K = d.size // m
        d = np.ascontiguousarray(d)
    out = np.empty(j.size, dtype=np.float64)
    _lib.call("mxg_mul_csr_dense", dtype, m, int(K), _vp(p), _vp(j), _vp(x), _vp(d), _vp(out))
    return out


def multiply_csr_by_dense_elemwise_double(indptr, indices, values, dense_mat):
    return _mul_dense(indptr, indices, values, dense_mat, MXG_Y_NUMERIC, np.float64)


def multiply_csr_by_dense_elemwise_float32(indptr, indices, values, dense_mat):
    return _mul_dense(indptr, indices, values, dense_mat, MXG_Y_FLOAT32, np.float32)


def multiply_csr_by_dense_elemwise_int(indptr, indices, values, dense_mat):
    return _mul_dense(indptr, indices, values, dense_mat, MXG_Y_INTEGER, np.int32)


def multiply_csr_by_dense_elemwise_bool(indptr, indices, values, dense_mat):
    return _mul_dense(indptr, indices, values, dense_mat, MXG_Y_LOGICAL, np.int32)


def multiply_csr_by_dvec_no_NAs_numeric(indptr, indices, values, dvec, ncols, multiply=True, powerto=False, divide=False,
                                        divrest=False, intdiv=False, X_is_LHS=True):
    """Only the Multiply operation is on the scoped path; the other operators of the reference's template
    (src/operators.cpp:1501-2143) stay on its C++."""
    if not multiply or powerto or divide or divrest or intdiv:
        raise NotImplementedError("only multiply=TRUE is implemented on the device (SURVEY.md §8 f4)")
    p, j, x = _csr(indptr, indices, values)
    d = np.ascontiguousarray(dvec, dtype=np.float64)
    out = np.empty(j.size, dtype=np.float64)
    _lib.call("mxg_mul_csr_dvec", p.size - 1, int(ncols), _vp(p), _vp(j), _vp(x), _vp(d), int(d.size), _vp(out))
    return out


# ---- additions beyond the reference's exports -----------------------------

def csr_to_csc(m, K, indptr, indices, values):
    """Deep CSR -> CSC on device, bit-exact with Matrix's `as(x, "CsparseMatrix")` (R/conversions.R:390-392)."""
    p, j, x = _csr(indptr, indices, values)
    p2 = np.empty(K + 1, dtype=np.int32)
    i2 = np.empty(j.size, dtype=np.int32)
    x2 = np.empty(j.size, dtype=np.float64)
    _lib.call("mxg_csr2csc", int(m), int(K), _vp(p), _vp(j), _vp(x), _vp(p2), _vp(i2), _vp(x2))
    return p2, i2, x2


def crossprod_csr_dense(indptr, indices, values, ncols_X, Y_colmajor, dtype=MXG_F64):
    """Out(K x n, column-major) = t(A_csr(m x K)) . Y(m x n): device transpose + gather product."""
    np_t = np.float64 if dtype == MXG_F64 else np.float32
    Y = _fmat(Y_colmajor, np_t)
    p, j, x = _csr(indptr, indices, values)
    m, n = Y.shape
    if p.size - 1 != m:
        raise ValueError("Matrix dimensions do not match.")
    K = int(ncols_X)
    out = _result((K, n), np_t, None)
    _lib.call("mxg_spmm_csrT_dense", dtype, MXG_COLS_CONTIGUOUS, MXG_COLS_CONTIGUOUS, m, K, n,
              _vp(p), _vp(j), _vp(x), _vp(Y), max(m, 1), _vp(out), max(K, 1))
    return out


# ---- device-resident matrices (rglue/handle_gpu_glue.cpp) ----------------------------
# The callers of the reference multiply one sparse matrix many times (vignettes/Introducing_MatrixExtra.Rmd:454-476);
# these exports keep the CSR in HBM so that a product moves only the dense operand up and the result down.

class GpuCsrPtr:
    """What the Rcpp glue returns as an external pointer: the C handle plus its dimensions; released by ``gpu_csr_free``
    or when the object is collected (the glue registers mxg_csr_free as the pointer's finalizer)."""

    def __init__(self, handle: int, nrows: int, ncols: int, has_f64: bool, has_f32: bool):
        self.handle, self.nrows, self.ncols, self.has_f64, self.has_f32 = handle, nrows, ncols, has_f64, has_f32

    def _live(self):
        if not self.handle:
            raise RuntimeError("gpu matrix has been freed.")
        return C.c_void_p(self.handle)

    def __del__(self):
        try:
            gpu_csr_free(self)
        except Exception:
            pass


def as_gpu_csr(indptr, indices, values, ncols, keep_float64=True, keep_float32=False) -> GpuCsrPtr:
    p, j, x = _csr(indptr, indices, values)
    if not keep_float64 and not keep_float32:
        keep_float64 = True
    h = C.c_void_p()
    _lib.call("mxg_csr_upload", p.size - 1, int(ncols), _vp(p), _vp(j), _vp(x),
              (_lib.MXG_KEEP_F64 if keep_float64 else 0) | (_lib.MXG_KEEP_F32 if keep_float32 else 0), C.byref(h))
    return GpuCsrPtr(h.value, p.size - 1, int(ncols), bool(keep_float64), bool(keep_float32))


def gpu_csr_free(ptr: GpuCsrPtr) -> None:
    if ptr.handle:
        h, ptr.handle = ptr.handle, 0
        _lib.call("mxg_csr_free", C.c_void_p(h))


def gpu_csr_dim(ptr: GpuCsrPtr):
    ptr._live()
    return np.array([ptr.nrows, ptr.ncols], dtype=np.int32)


def _typed(ptr: GpuCsrPtr, dtype):
    if not (ptr.has_f64 if dtype == MXG_F64 else ptr.has_f32):
        raise RuntimeError("gpu matrix was created without values of this type.")
    return np.float64 if dtype == MXG_F64 else np.float32


def _gpu_csr_tcrossprod_dense(ptr, Y_colmajor, dtype, out=None):
    h = ptr._live()
    np_t = _typed(ptr, dtype)
    Y = _fmat(Y_colmajor, np_t)
    n, K = Y.shape
    if K != ptr.ncols:
        raise ValueError("Matrix dimensions do not match.")
    out = _result((ptr.nrows, n), np_t, out)
    _lib.call("mxg_csr_spmm_host", h, dtype, MXG_COLS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, n, _vp(Y), max(n, 1), _vp(out),
              max(ptr.nrows, 1))
    return out


def gpu_csr_tcrossprod_dense_numeric(X: GpuCsrPtr, Y_colmajor, nthreads=0, out=None):
    """A %*% t(Y): the handle form of tcrossprod_csr_dense_numeric (src/matmul.cpp:345-359)."""
    _threads(nthreads)
    return _gpu_csr_tcrossprod_dense(X, Y_colmajor, MXG_F64, out)


def gpu_csr_tcrossprod_dense_float32(X: GpuCsrPtr, Y_colmajor, nthreads=0, out=None):
    _threads(nthreads)
    return _gpu_csr_tcrossprod_dense(X, Y_colmajor, MXG_F32, out)


def _gpu_csr_dense_tcrossprod(X_colmajor, ptr, dtype, out=None):
    h = ptr._live()
    np_t = _typed(ptr, dtype)
    X = _fmat(X_colmajor, np_t)
    a, K = X.shape
    if K != ptr.ncols:
        raise ValueError("Matrix dimensions do not match.")
    out = _result((a, ptr.nrows), np_t, out)
    _lib.call("mxg_csr_spmm_host", h, dtype, MXG_ROWS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, a, _vp(X), max(a, 1), _vp(out), max(a, 1))
    return out


def gpu_csr_dense_tcrossprod_numeric(X_colmajor, Y: GpuCsrPtr, nthreads=0, out=None):
    """X %*% t(A): the handle form of tcrossprod_dense_csr_numeric (src/matmul.cpp:283-297)."""
    _threads(nthreads)
    return _gpu_csr_dense_tcrossprod(X_colmajor, Y, MXG_F64, out)


def gpu_csr_dense_tcrossprod_float32(X_colmajor, Y: GpuCsrPtr, nthreads=0, out=None):
    _threads(nthreads)
    return _gpu_csr_dense_tcrossprod(X_colmajor, Y, MXG_F32, out)


def _gpu_csr_crossprod_dense(ptr, Y_colmajor, dtype, out=None):
    h = ptr._live()
    np_t = _typed(ptr, dtype)
    Y = _fmat(Y_colmajor, np_t)
    m, n = Y.shape
    if m != ptr.nrows:
        raise ValueError("Matrix dimensions do not match.")
    out = _result((ptr.ncols, n), np_t, out)
    _lib.call("mxg_csr_spmm_t_host", h, dtype, MXG_COLS_CONTIGUOUS, MXG_COLS_CONTIGUOUS, n, _vp(Y), max(m, 1), _vp(out),
              max(ptr.ncols, 1))
    return out


def gpu_csr_crossprod_dense_numeric(X: GpuCsrPtr, Y_colmajor, nthreads=0, out=None):
    """t(A) %*% Y: the CSC is built on the device at the first call and kept with the handle."""
    _threads(nthreads)
    return _gpu_csr_crossprod_dense(X, Y_colmajor, MXG_F64, out)


def gpu_csr_crossprod_dense_float32(X: GpuCsrPtr, Y_colmajor, nthreads=0, out=None):
    _threads(nthreads)
    return _gpu_csr_crossprod_dense(X, Y_colmajor, MXG_F32, out)


def gpu_csr_dvec_numeric(X: GpuCsrPtr, y_dense, nthreads=0, out=None):
    """A %*% dense vector: the handle form of matmul_csr_dvec_numeric (src/matmul.cpp:421-435)."""
    _threads(nthreads)
    h = X._live()
    if not X.has_f64:
        raise RuntimeError("gpu matrix was created without values of this type.")
    y = np.ascontiguousarray(y_dense, dtype=np.float64)
    if y.size != X.ncols:
        raise ValueError("Matrix dimensions do not match.")
    out = _result((X.nrows,), np.float64, out)
    _lib.call("mxg_csr_spmv_host", h, MXG_Y_NUMERIC, _vp(y), _vp(out))
    return out


def mxgpu_configure(gpus=0, cache_mb=-1) -> int:
    """gpus > 0: spread every level-1 product over that many GPUs (MATRIXEXTRA_GPUS); cache_mb >= 0: size of the
    device-resident operand cache (MATRIXEXTRA_GPU_CACHE_MB).  Returns the number of devices in use."""
    if gpus > 0:
        _lib.call("mxg_set_devices", int(gpus))
    if cache_mb >= 0:
        _lib.set_option("cache_mb", int(cache_mb))
    n = C.c_int(1)
    _lib.call("mxg_get_devices", C.byref(n))
    return int(n.value)
