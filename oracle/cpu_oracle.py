"""CPU oracle bindings — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package (``matrixextra_b200``) never does.

Two back-ends with the same interface, both named after the reference's Rcpp exports
(src/matmul.cpp:221-483) and taking arguments with exactly the layouts R hands to them:

* :class:`Port`  — ``oracle/libmxoracle.so``: the plain-C restatement in ``oracle/mx_oracle.c``.
  The wrappers below restate the typed entry points (src/matmul.cpp:188-343: which kernel, which
  dimensions, zero-filled output).
* :class:`Ref`   — ``oracle/_ref/libmxref*.so``: the reference's own ``src/matmul.cpp`` compiled in
  place (see ``oracle/Makefile``, ``oracle/ref_driver.cpp``).  Present whenever it was built in the
  build container; it travels to the GPU box as a prebuilt file.

Conventions: dense matrices are Fortran-order (column-major, like R); float32 matrices are
``np.float32`` arrays (R stores the same bits in an integer matrix, src/matmul.cpp:213-214);
CSR/CSC arrays are int32 ``p``/``j`` and float64 ``x``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)


def build(quiet: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile.  Building the checker is not using it."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def _as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _fdense(a, dtype):
    a = np.asarray(a, dtype=dtype)
    if a.ndim != 2:
        raise ValueError("dense operand must be a matrix")
    return np.asfortranarray(a)


def cpu_has_avx2_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return {"avx2", "fma", "bmi2"} <= flags
    except OSError:
        pass
    return False


class Port:
    """Plain-C restatement (oracle/mx_oracle.c) behind the reference's entry-point names."""

    kind = "port"

    def __init__(self, nthreads: int | None = None):
        path = os.path.join(_HERE, "libmxoracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.mxo_max_threads.restype = C.c_int
        self.max_threads = int(self.lib.mxo_max_threads())
        self.nthreads = int(nthreads or 1)
        self.flags = "-O2 -ffp-contract=off -fopenmp (gcc), portable x86-64"
        for name in ("mxo_gemm_csr_drm_as_drm_f64", "mxo_gemm_csr_drm_as_dcm_f64"):
            getattr(self.lib, name).argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_size_t,
                                                _f64p, C.c_size_t, C.c_int]
            getattr(self.lib, name).restype = None
        for name in ("mxo_gemm_csr_drm_as_drm_f32", "mxo_gemm_csr_drm_as_dcm_f32"):
            getattr(self.lib, name).argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f32p, C.c_size_t,
                                                _f32p, C.c_size_t, C.c_int]
            getattr(self.lib, name).restype = None
        self.lib.mxo_spmv_numeric.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int]
        self.lib.mxo_spmv_integer.argtypes = [C.c_int, _i32p, _i32p, _f64p, _i32p, _f64p, C.c_int]
        self.lib.mxo_spmv_logical.argtypes = [C.c_int, _i32p, _i32p, _f64p, _i32p, _f64p, C.c_int]
        self.lib.mxo_spmv_float32.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f32p, _f32p, C.c_int]
        self.lib.mxo_csr2csc.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        self.lib.mxo_csr2csc.restype = C.c_int
        self.lib.mxo_spmv_svec.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, _i32p, C.c_void_p, _f64p, C.c_int]
        self.lib.mxo_spmv_svec.restype = None

    # -- kernels ------------------------------------------------------------------------------
    def _gather_rm(self, p, j, x, B_rows, nt):
        """Out(rows x n, row-major) = CSR . B  with B given as (K x n) row-major == (n x K) F-order."""
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        rows = p.size - 1
        n = B_rows.shape[0]  # B_rows is the F-order (n x K) matrix R passes; its memory is K x n row-major
        f32 = B_rows.dtype == np.float32
        out = np.zeros((n, rows), dtype=B_rows.dtype, order="F")  # zero-filled, src/matmul.cpp:197/261
        fn = self.lib.mxo_gemm_csr_drm_as_drm_f32 if f32 else self.lib.mxo_gemm_csr_drm_as_drm_f64
        tp = _f32p if f32 else _f64p
        fn(rows, n, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p), _ptr(B_rows, tp), n, _ptr(out, tp), n, nt)
        return out

    def _gather_cm(self, p, j, x, Y, nt):
        """Out(m x n, col-major) = CSR(m x K) . t(Y) with Y (n x K) F-order."""
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        m = p.size - 1
        n = Y.shape[0]
        f32 = Y.dtype == np.float32
        out = np.zeros((m, n), dtype=Y.dtype, order="F")  # src/matmul.cpp:323
        fn = self.lib.mxo_gemm_csr_drm_as_dcm_f32 if f32 else self.lib.mxo_gemm_csr_drm_as_dcm_f64
        tp = _f32p if f32 else _f64p
        fn(m, n, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p), _ptr(Y, tp), n, _ptr(out, tp), m, nt)
        return out

    # -- the reference's exports (src/matmul.cpp:221-483) ---------------------------------------
    def matmul_dense_csc_numeric(self, X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values, nthreads=None):
        return self._gather_rm(Y_csc_indptr, Y_csc_indices, Y_csc_values, _fdense(X_colmajor, np.float64),
                               nthreads or self.nthreads)

    def matmul_dense_csc_float32(self, X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values, nthreads=None):
        return self._gather_rm(Y_csc_indptr, Y_csc_indices, Y_csc_values, _fdense(X_colmajor, np.float32),
                               nthreads or self.nthreads)

    def tcrossprod_dense_csr_numeric(self, X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values,
                                     nthreads=None, ncols_Y=0):
        return self._gather_rm(Y_csr_indptr, Y_csr_indices, Y_csr_values, _fdense(X_colmajor, np.float64),
                               nthreads or self.nthreads)

    def tcrossprod_dense_csr_float32(self, X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values,
                                     nthreads=None, ncols_Y=0):
        return self._gather_rm(Y_csr_indptr, Y_csr_indices, Y_csr_values, _fdense(X_colmajor, np.float32),
                               nthreads or self.nthreads)

    def tcrossprod_csr_dense_numeric(self, X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor, nthreads=None):
        return self._gather_cm(X_csr_indptr, X_csr_indices, X_csr_values, _fdense(Y_colmajor, np.float64),
                               nthreads or self.nthreads)

    def tcrossprod_csr_dense_float32(self, X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor, nthreads=None):
        return self._gather_cm(X_csr_indptr, X_csr_indices, X_csr_values, _fdense(Y_colmajor, np.float32),
                               nthreads or self.nthreads)

    def _spmv(self, fn, p, j, x, y, ytype, otype, nt):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        y = np.ascontiguousarray(y, dtype=ytype)
        out = np.zeros(p.size - 1, dtype=otype)
        yp = {np.float64: _f64p, np.int32: _i32p, np.float32: _f32p}[ytype]
        op = _f32p if otype == np.float32 else _f64p
        fn(p.size - 1, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p), _ptr(y, yp), _ptr(out, op), nt)
        return out

    def matmul_csr_dvec_numeric(self, p, j, x, y_dense, nthreads=None):
        return self._spmv(self.lib.mxo_spmv_numeric, p, j, x, y_dense, np.float64, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_dvec_integer(self, p, j, x, y_dense, nthreads=None):
        return self._spmv(self.lib.mxo_spmv_integer, p, j, x, y_dense, np.int32, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_dvec_logical(self, p, j, x, y_dense, nthreads=None):
        return self._spmv(self.lib.mxo_spmv_logical, p, j, x, y_dense, np.int32, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_dvec_float32(self, p, j, x, y_dense, nthreads=None):
        return self._spmv(self.lib.mxo_spmv_float32, p, j, x, y_dense, np.float32, np.float32,
                          nthreads or self.nthreads)

    def matmul_rowvec_by_csc(self, rowvec, p, i, x=None):
        """float32 row vector %*% CSC (src/matmul.cpp:643-684; x=None: matmul_rowvec_by_cscbin): per column
        `out[col] += values[ix] * rowvec[indices[ix]]` with a float accumulator — the float32 SpMV loop over the CSC
        arrays, restated by mxo_spmv_float32."""
        x = np.ones(np.asarray(i).size) if x is None else x
        return self._spmv(self.lib.mxo_spmv_float32, p, i, x, rowvec, np.float32, np.float32, 1).reshape(1, -1)

    # -- CSR %*% sparse vector (src/matmul.cpp:486-641) ------------------------------
    def _svec(self, ytype, p, j, x, yi, yv, np_t, nt):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        yi = _as_i32(yi)
        yv = np.zeros(1, dtype=np.int32) if yv is None else np.ascontiguousarray(yv, dtype=np_t)
        out = np.zeros(p.size - 1, dtype=np.float64)
        self.lib.mxo_spmv_svec(ytype, p.size - 1, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p), yi.size,
                               _ptr(yi, _i32p), C.c_void_p(yv.ctypes.data), _ptr(out, _f64p), nt)
        return out

    def matmul_csr_svec_numeric(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(0, p, j, x, y_indices_base1, y_values, np.float64, nthreads or self.nthreads)

    def matmul_csr_svec_integer(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(1, p, j, x, y_indices_base1, y_values, np.int32, nthreads or self.nthreads)

    def matmul_csr_svec_logical(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(2, p, j, x, y_indices_base1, y_values, np.int32, nthreads or self.nthreads)

    def matmul_csr_svec_float32(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(3, p, j, x, y_indices_base1, y_values, np.float32, nthreads or self.nthreads)

    def matmul_csr_svec_binary(self, p, j, x, y_indices_base1, nthreads=None):
        return self._svec(4, p, j, x, y_indices_base1, None, np.int32, nthreads or self.nthreads)

    # -- index sorting and validity (src/misc.cpp:117-228, 970-1016) -----------------
    def sort_sparse_indices_numeric(self, p, j, x):
        """Returns sorted COPIES (the reference sorts in place)."""
        p = _as_i32(p)
        j, x = _as_i32(j).copy(), _as_f64(x).copy()
        self.lib.mxo_sort_sparse_indices.argtypes = [C.c_int, _i32p, _i32p, _f64p]
        if self.lib.mxo_sort_sparse_indices(p.size - 1, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p)) != 0:
            raise MemoryError
        return j, x

    def check_indices_are_unsorted(self, p, j):
        """True when every row is sorted (the reference's name is misleading, src/misc.cpp:161-175)."""
        p, j = _as_i32(p), _as_i32(j)
        self.lib.mxo_rows_sorted.argtypes = [C.c_int, _i32p, _i32p]
        return bool(self.lib.mxo_rows_sorted(p.size - 1, _ptr(p, _i32p), _ptr(j, _i32p)))

    CSR_ERRORS = (None, "Matrix has negative indices.", "Matrix has invalid column indices.",
                  "Matrix has indices with missing values.", "Matrix has missing values in the index pointer.",
                  "Matrix index pointer is not monotonicaly increasing.")

    def check_valid_csr_matrix(self, p, j, nrows, ncols):
        """None when valid, else the reference's error string (src/misc.cpp:970-1016)."""
        p, j = _as_i32(p), _as_i32(j)
        self.lib.mxo_check_valid_csr.argtypes = [C.c_int, C.c_int, _i32p, _i32p, C.c_size_t]
        return self.CSR_ERRORS[self.lib.mxo_check_valid_csr(int(nrows), int(ncols), _ptr(p, _i32p), _ptr(j, _i32p), j.size)]

    # -- elementwise CSR * dense (src/operators.cpp:239-330, 1501-2178) ---------------
    def _mult_dense(self, dtype, p, j, x, dense, np_t):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        d = np.asfortranarray(dense, dtype=np_t)
        out = np.zeros(j.size, dtype=np.float64)
        self.lib.mxo_multiply_csr_by_dense.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_void_p, _f64p]
        self.lib.mxo_multiply_csr_by_dense.restype = None
        self.lib.mxo_multiply_csr_by_dense(dtype, p.size - 1, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p),
                                           C.c_void_p(d.ctypes.data), _ptr(out, _f64p))
        return out

    def multiply_csr_by_dense_elemwise_double(self, p, j, x, dense_mat):
        return self._mult_dense(0, p, j, x, dense_mat, np.float64)

    def multiply_csr_by_dense_elemwise_float32(self, p, j, x, dense_mat):
        return self._mult_dense(1, p, j, x, dense_mat, np.float32)

    def multiply_csr_by_dense_elemwise_int(self, p, j, x, dense_mat):
        return self._mult_dense(2, p, j, x, dense_mat, np.int32)

    def multiply_csr_by_dense_elemwise_bool(self, p, j, x, dense_mat):
        return self._mult_dense(3, p, j, x, dense_mat, np.int32)

    def multiply_csr_by_dvec_no_NAs_numeric(self, p, j, x, dvec, ncols):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        d = _as_f64(dvec)
        out = np.zeros(j.size, dtype=np.float64)
        self.lib.mxo_multiply_csr_by_dvec.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_size_t, _f64p]
        self.lib.mxo_multiply_csr_by_dvec.restype = None
        self.lib.mxo_multiply_csr_by_dvec(p.size - 1, int(ncols), _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p),
                                          _ptr(d, _f64p), d.size, _ptr(out, _f64p))
        return out

    # -- CSR -> CSC (Matrix package; restated) ---------------------------------------------------
    def csr2csc(self, m, K, p, j, x):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        p2 = np.zeros(K + 1, dtype=np.int32)
        i2 = np.zeros(j.size, dtype=np.int32)
        x2 = np.zeros(j.size, dtype=np.float64)
        rc = self.lib.mxo_csr2csc(m, K, _ptr(p, _i32p), _ptr(j, _i32p), _ptr(x, _f64p),
                                  _ptr(p2, _i32p), _ptr(i2, _i32p), _ptr(x2, _f64p))
        if rc != 0:
            raise ValueError("column index out of range")
        return p2, i2, x2


class Ref:
    """The reference's own src/matmul.cpp (oracle/_ref), same interface as :class:`Port`."""

    kind = "reference"

    def __init__(self, nthreads: int | None = None, fast: bool = False):
        name = "libmxref_v3.so" if (fast and cpu_has_avx2_fma()) else "libmxref.so"
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.flags = ("-O3 -march=x86-64-v3 -fopenmp (g++)" if name.endswith("_v3.so")
                      else "-O2 -ffp-contract=off -fopenmp (g++), portable x86-64")
        self.lib = C.CDLL(path)
        self.lib.mxref_max_threads.restype = C.c_int
        self.max_threads = int(self.lib.mxref_max_threads())
        self.nthreads = int(nthreads or 1)
        self.lib.mxref_last_result.restype = C.c_void_p
        self.copy_result = True  # bench.py sets False: time the reference without an extra result copy
        self.lib.mxref_last_result.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libmxref.so"))

    def _result(self, dtype, copy=True):
        nr, nc = C.c_size_t(), C.c_size_t()
        addr = self.lib.mxref_last_result(C.byref(nr), C.byref(nc))
        n = nr.value * nc.value
        if n == 0:
            return np.zeros((nr.value, nc.value), dtype=dtype, order="F")
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
        arr = np.frombuffer(buf, dtype=dtype).reshape((nr.value, nc.value), order="F")
        return arr.copy(order="F") if (copy and self.copy_result) else arr

    def _dense_first(self, fn, X, dtype, p, j, x, nt, extra=()):
        X = _fdense(X, dtype)
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        rc = fn(C.c_void_p(X.ctypes.data), X.shape[0], X.shape[1], _ptr(p, _i32p), p.size - 1,
                _ptr(j, _i32p), _ptr(x, _f64p), j.size, nt, *extra, None)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")
        return self._result(dtype)

    def matmul_dense_csc_numeric(self, X, p, i, x, nthreads=None):
        return self._dense_first(self.lib.mxref_matmul_dense_csc_numeric, X, np.float64, p, i, x,
                                 nthreads or self.nthreads)

    def matmul_dense_csc_float32(self, X, p, i, x, nthreads=None):
        return self._dense_first(self.lib.mxref_matmul_dense_csc_float32, X, np.float32, p, i, x,
                                 nthreads or self.nthreads)

    def tcrossprod_dense_csr_numeric(self, X, p, j, x, nthreads=None, ncols_Y=0):
        return self._dense_first(self.lib.mxref_tcrossprod_dense_csr_numeric, X, np.float64, p, j, x,
                                 nthreads or self.nthreads, (int(ncols_Y),))

    def tcrossprod_dense_csr_float32(self, X, p, j, x, nthreads=None, ncols_Y=0):
        return self._dense_first(self.lib.mxref_tcrossprod_dense_csr_float32, X, np.float32, p, j, x,
                                 nthreads or self.nthreads, (int(ncols_Y),))

    def _csr_first(self, fn, p, j, x, Y, dtype, nt, copy=True):
        Y = _fdense(Y, dtype)
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        if p.size - 1 < Y.shape[0]:
            raise ValueError("reference limitation (src/matmul.cpp:176-182): needs CSR rows >= dense rows")
        rc = fn(_ptr(p, _i32p), p.size - 1, _ptr(j, _i32p), _ptr(x, _f64p), j.size,
                C.c_void_p(Y.ctypes.data), Y.shape[0], Y.shape[1], nt, None)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")
        return self._result(dtype, copy=copy)

    def tcrossprod_csr_dense_numeric(self, p, j, x, Y, nthreads=None):
        return self._csr_first(self.lib.mxref_tcrossprod_csr_dense_numeric, p, j, x, Y, np.float64,
                               nthreads or self.nthreads)

    def tcrossprod_csr_dense_float32(self, p, j, x, Y, nthreads=None):
        return self._csr_first(self.lib.mxref_tcrossprod_csr_dense_float32, p, j, x, Y, np.float32,
                               nthreads or self.nthreads)

    def _spmv(self, fn, p, j, x, y, ytype, otype, nt):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        y = np.ascontiguousarray(y, dtype=ytype)
        rc = fn(_ptr(p, _i32p), p.size - 1, _ptr(j, _i32p), _ptr(x, _f64p), j.size,
                C.c_void_p(y.ctypes.data), y.size, nt, None)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")
        return self._result(otype).reshape(-1)

    def matmul_csr_dvec_numeric(self, p, j, x, y, nthreads=None):
        return self._spmv(self.lib.mxref_matmul_csr_dvec_numeric, p, j, x, y, np.float64, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_dvec_integer(self, p, j, x, y, nthreads=None):
        return self._spmv(self.lib.mxref_matmul_csr_dvec_integer, p, j, x, y, np.int32, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_dvec_logical(self, p, j, x, y, nthreads=None):
        return self._spmv(self.lib.mxref_matmul_csr_dvec_logical, p, j, x, y, np.int32, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_dvec_float32(self, p, j, x, y, nthreads=None):
        return self._spmv(self.lib.mxref_matmul_csr_dvec_float32, p, j, x, y, np.float32, np.float32,
                          nthreads or self.nthreads)

    def matmul_rowvec_by_csc(self, rowvec, p, i, x=None):
        """The reference's own matmul_rowvec_by_csc / matmul_rowvec_by_cscbin (src/matmul.cpp:643-684)."""
        p, i = _as_i32(p), _as_i32(i)
        rv = np.ascontiguousarray(rowvec, dtype=np.float32)
        xx = None if x is None else _as_f64(x)
        out = np.zeros(p.size - 1, dtype=np.float32)
        rc = self.lib.mxref_matmul_rowvec_by_csc(C.c_void_p(rv.ctypes.data), rv.size, _ptr(p, _i32p), p.size - 1,
                                                 _ptr(i, _i32p), None if xx is None else _ptr(xx, _f64p), i.size,
                                                 C.c_void_p(out.ctypes.data))
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")
        return out.reshape(1, -1)

    # -- CSR %*% sparse vector, the reference's own exports (src/matmul.cpp:553-641) ---
    def _svec(self, fn, p, j, x, yi, yv, np_t, nt):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        yi = _as_i32(yi)
        yv = np.zeros(1, dtype=np.int32) if yv is None else np.ascontiguousarray(yv, dtype=np_t)
        rc = fn(_ptr(p, _i32p), p.size - 1, _ptr(j, _i32p), _ptr(x, _f64p), j.size, _ptr(yi, _i32p),
                C.c_void_p(yv.ctypes.data), yi.size, nt, None)
        if rc != 0:
            raise RuntimeError(f"reference driver returned {rc}")
        return self._result(np.float64).reshape(-1)

    def matmul_csr_svec_numeric(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(self.lib.mxref_matmul_csr_svec_numeric, p, j, x, y_indices_base1, y_values, np.float64,
                          nthreads or self.nthreads)

    def matmul_csr_svec_integer(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(self.lib.mxref_matmul_csr_svec_integer, p, j, x, y_indices_base1, y_values, np.int32,
                          nthreads or self.nthreads)

    def matmul_csr_svec_logical(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(self.lib.mxref_matmul_csr_svec_logical, p, j, x, y_indices_base1, y_values, np.int32,
                          nthreads or self.nthreads)

    def matmul_csr_svec_float32(self, p, j, x, y_indices_base1, y_values, nthreads=None):
        return self._svec(self.lib.mxref_matmul_csr_svec_float32, p, j, x, y_indices_base1, y_values, np.float32,
                          nthreads or self.nthreads)

    def matmul_csr_svec_binary(self, p, j, x, y_indices_base1, nthreads=None):
        return self._svec(self.lib.mxref_matmul_csr_svec_binary, p, j, x, y_indices_base1, None, np.int32,
                          nthreads or self.nthreads)

    # -- src/misc.cpp and src/operators.cpp compiled in place (libmxref_ops.so) ---
    @staticmethod
    def ops_available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libmxref_ops.so"))

    @property
    def ops(self):
        if getattr(self, "_ops", None) is None:
            self._ops = C.CDLL(os.path.join(_HERE, "_ref", "libmxref_ops.so"))
        return self._ops

    def sort_sparse_indices_numeric(self, p, j, x):
        """Returns sorted COPIES (the reference sorts in place, src/misc.cpp:300-313)."""
        p = _as_i32(p)
        j, x = _as_i32(j).copy(), _as_f64(x).copy()
        self.ops.mxref_sort_sparse_indices_numeric(_ptr(p, _i32p), C.c_int(p.size - 1), _ptr(j, _i32p), _ptr(x, _f64p),
                                                   C.c_int(j.size))
        return j, x

    def check_indices_are_unsorted(self, p, j):
        p, j = _as_i32(p), _as_i32(j)
        return bool(self.ops.mxref_check_indices_are_unsorted(_ptr(p, _i32p), C.c_int(p.size - 1), _ptr(j, _i32p)))

    def check_valid_csr_matrix(self, p, j, nrows, ncols):
        p, j = _as_i32(p), _as_i32(j)
        buf = C.create_string_buffer(256)
        rc = self.ops.mxref_check_valid_csr_matrix(_ptr(p, _i32p), C.c_int(int(nrows)), _ptr(j, _i32p), C.c_int(j.size),
                                                   C.c_int(int(ncols)), buf, C.c_int(256))
        return buf.value.decode() if rc else None

    def _mult_dense(self, fn, p, j, x, dense, np_t):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        d = np.asfortranarray(dense, dtype=np_t)
        out = np.zeros(j.size, dtype=np.float64)
        fn(_ptr(p, _i32p), C.c_int(p.size - 1), _ptr(j, _i32p), _ptr(x, _f64p), C.c_int(j.size),
           C.c_void_p(d.ctypes.data), C.c_size_t(d.size), _ptr(out, _f64p))
        return out

    def multiply_csr_by_dense_elemwise_double(self, p, j, x, dense_mat):
        return self._mult_dense(self.ops.mxref_multiply_csr_by_dense_elemwise_double, p, j, x, dense_mat, np.float64)

    def multiply_csr_by_dense_elemwise_float32(self, p, j, x, dense_mat):
        return self._mult_dense(self.ops.mxref_multiply_csr_by_dense_elemwise_float32, p, j, x, dense_mat, np.float32)

    def multiply_csr_by_dense_elemwise_int(self, p, j, x, dense_mat):
        return self._mult_dense(self.ops.mxref_multiply_csr_by_dense_elemwise_int, p, j, x, dense_mat, np.int32)

    def multiply_csr_by_dense_elemwise_bool(self, p, j, x, dense_mat):
        return self._mult_dense(self.ops.mxref_multiply_csr_by_dense_elemwise_bool, p, j, x, dense_mat, np.int32)

    def multiply_csr_by_dvec_no_NAs_numeric(self, p, j, x, dvec, ncols):
        p, j, x = _as_i32(p), _as_i32(j), _as_f64(x)
        d = _as_f64(dvec)
        out = np.zeros(j.size, dtype=np.float64)
        self.ops.mxref_multiply_csr_by_dvec_numeric(_ptr(p, _i32p), C.c_int(p.size - 1), _ptr(j, _i32p), _ptr(x, _f64p),
                                                    C.c_int(j.size), _ptr(d, _f64p), C.c_size_t(d.size), C.c_int(int(ncols)),
                                                    _ptr(out, _f64p))
        return out


def synth_csr_host(m, K, nnz, row_model=1, col_model=0, seed=1000):
    """The synthetic CSR of SURVEY.md 8d generated on the HOST (oracle/mx_synth.c, the twin of csrc/synth.cu): what
    bench.py's reference arm multiplies, so that it needs neither the product library nor a GPU."""
    lib = C.CDLL(os.path.join(_HERE, "libmxoracle.so"))
    lib.mxs_synth_indptr.restype = C.c_longlong
    lib.mxs_synth_indptr.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_uint64, _i32p]
    lib.mxs_synth_entries.restype = None
    lib.mxs_synth_entries.argtypes = [C.c_int, C.c_int, _i32p, C.c_int, C.c_uint64, _i32p, _f64p]
    p = np.empty(int(m) + 1, dtype=np.int32)
    got = lib.mxs_synth_indptr(int(m), int(K), int(nnz), int(row_model), int(seed), _ptr(p, _i32p))
    if got < 0:
        raise ValueError("synth_csr_host: bad arguments")
    j = np.empty(int(got), dtype=np.int32)
    x = np.empty(int(got), dtype=np.float64)
    lib.mxs_synth_entries(int(m), int(K), _ptr(p, _i32p), int(col_model), int(seed), _ptr(j, _i32p), _ptr(x, _f64p))
    return p, j, x


def best_cpu_baseline(nthreads: int | None = None):
    """The strongest available CPU implementation of the path: the reference itself when its
    prebuilt library is present (fast build if the CPU supports it), else the port."""
    if Ref.available():
        try:
            return Ref(nthreads=nthreads, fast=True)
        except OSError:
            pass
    return Port(nthreads=nthreads)
