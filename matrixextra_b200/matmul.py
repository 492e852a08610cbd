"""Python mirror of the S4 methods of R/matmul.R for the dense-operand products.

``matmul(x, y)`` is ``x %*% y``, ``crossprod(x, y)`` is ``t(x) %*% y``, ``tcrossprod(x, y)`` is
``x %*% t(y)``.  Dispatch is on the operand classes exactly like the ``setMethod`` table
(R/matmul.R:200, 281, 307, 385, 393, 434, 461, 469, 512, 538, 755-763): which operand is
transposed, which dimension check runs (and which branches skip it), which Rcpp export is called
and which class comes back.  The three signatures MatrixExtra leaves to the Matrix package
(``crossprod(Rsparse, matrix)``, ``Csparse %*% matrix``, ``matrix %*% Rsparse``)
are added on top of the device CSR->CSC transpose.

Sparse-vector right-hand sides (R/matmul.R:595-646) run on the device too; the
single-column outer products (R/matmul.R:659-744) are outside the scoped path and raise
``NotImplementedError``.
"""
from __future__ import annotations

import numpy as np

from . import rcpp_exports as rx
from ._lib import MXG_F32, MXG_F64
from .classes import check_valid_matrix, dgCMatrix, dgRMatrix, float32, gpuRsparse, sparseVector, t_shallow

#: options("MatrixExtra.*") read by the hot path (R/zzz.R:116-171).  ``nthreads`` defaults to all cores in the
#: reference (parallel::detectCores(), R/zzz.R:140-171); here it sizes the library's host staging threads, 0 = all.
options = {"MatrixExtra.nthreads": 0, "MatrixExtra.inplace_sort": False}


def _nthreads() -> int:
    return max(int(options.get("MatrixExtra.nthreads", 0)), 0)


def _is_dense(x) -> bool:
    return isinstance(x, np.ndarray) and x.ndim == 2


def _as_double(x):
    # `if (typeof(x) != "double") mode(x) <- "double"` (R/matmul.R:182, 290, 443)
    return np.asfortranarray(x, dtype=np.float64)


def check_dimensions_match(x, y, matmult=False, crossprod=False, tcrossprod=False):
    """R/matmul.R:130-146."""
    if matmult:
        inner_x, inner_y = x.shape[1], y.shape[0]
    elif crossprod:
        inner_x, inner_y = x.shape[0], y.shape[0]
    elif tcrossprod:
        inner_x, inner_y = x.shape[1], y.shape[1]
    else:
        raise RuntimeError("Internal error.")
    if inner_x != inner_y:
        raise ValueError("Matrix dimensions do not match.")


# ---- R/matmul.R:171-196 ---------------------------------------------------------------------------
def gemm_dense_csc(x, y: dgCMatrix):
    check_dimensions_match(x, y, matmult=True)
    x = _as_double(x)
    check_valid_matrix(y)
    return rx.matmul_dense_csc_numeric(x, y.p, y.i, y.x, _nthreads())


# ---- R/matmul.R:202-277 (matrix branch 264-276: no dimension check) ---------------------------------
def gemm_f32_csc(x: float32, y: dgCMatrix):
    if x.is_vector():
        check_valid_matrix(y)
        if y.Dim[0] == 1:  # R/matmul.R:221-241: [n,1] %*% [1,k] outer product with a sparse result
            raise NotImplementedError("single-row outer product (sparse result) is outside the scoped path (R/matmul.R:221-241)")
        if y.Dim[0] != x.Data.size:  # R/matmul.R:243-244
            raise ValueError("(row) vector-Matrix multiplication dimensions do not match.")
        return float32(rx.matmul_rowvec_by_csc(x.Data.ravel(), y.p, y.i, y.x))  # R/matmul.R:246-259
    check_valid_matrix(y)
    return float32(rx.matmul_dense_csc_float32(x.Data, y.p, y.i, y.x, _nthreads()))


# ---- R/matmul.R:283-303 ---------------------------------------------------------------------------
def tcrossprod_dense_csr(x, y: dgRMatrix):
    check_dimensions_match(x, y, tcrossprod=True)
    x = _as_double(x)
    check_valid_matrix(y)
    return rx.tcrossprod_dense_csr_numeric(x, y.p, y.j, y.x, _nthreads(), y.Dim[1])


# ---- R/matmul.R:309-381 (matrix branch 369-380: no dimension check) ---------------------------------
def tcrossprod_f32_csr(x: float32, y: dgRMatrix):
    if x.is_vector():
        check_valid_matrix(y)
        if y.Dim[1] == 1:  # R/matmul.R:326-348: outer product with a sparse result
            raise NotImplementedError("single-column outer product (sparse result) is outside the scoped path (R/matmul.R:326-348)")
        if y.Dim[1] != x.Data.size:
            raise ValueError("(row) vector-Matrix multiplication dimensions do not match.")
        return float32(rx.matmul_rowvec_by_csc(x.Data.ravel(), y.p, y.j, y.x))  # R/matmul.R:350-366: CSR(y) == CSC(t(y))
    check_valid_matrix(y)
    return float32(rx.tcrossprod_dense_csr_float32(x.Data, y.p, y.j, y.x, _nthreads(), y.Dim[1]))


# ---- R/matmul.R:387-389 ---------------------------------------------------------------------------
def crossprod_dense_csc(x, y: dgCMatrix):
    return gemm_dense_csc(np.asfortranarray(np.asarray(x).T), y)


# ---- R/matmul.R:395-430 ---------------------------------------------------------------------------
def crossprod_f32_csc(x: float32, y: dgCMatrix):
    if x.is_vector():  # R/matmul.R:402-427
        if x.Data.size != y.Dim[0]:
            raise ValueError("(column) vector-Matrix crossprod dimensions do not match.")
        check_valid_matrix(y)
        return float32(rx.matmul_rowvec_by_csc(x.Data.ravel(), y.p, y.i, y.x))
    return gemm_f32_csc(float32(x.Data.T), y)


# ---- R/matmul.R:436-457 ---------------------------------------------------------------------------
def tcrossprod_csr_dense(x: dgRMatrix, y):
    check_dimensions_match(x, y, tcrossprod=True)
    y = _as_double(y)
    check_valid_matrix(x)
    return rx.tcrossprod_csr_dense_numeric(x.p, x.j, x.x, y, _nthreads())


# ---- R/matmul.R:463-465 ---------------------------------------------------------------------------
def gemm_csr_dense(x: dgRMatrix, y):
    return tcrossprod_csr_dense(x, np.asfortranarray(np.asarray(y).T))


# ---- R/matmul.R:514-534 ---------------------------------------------------------------------------
def tcrossprod_csr_f32(x: dgRMatrix, y: float32):
    check_dimensions_match(x, y, tcrossprod=True)
    check_valid_matrix(x)
    return float32(rx.tcrossprod_csr_dense_float32(x.p, x.j, x.x, y.Data, _nthreads()))


# ---- R/matmul.R:471-508 ---------------------------------------------------------------------------
def gemm_csr_f32(x: dgRMatrix, y: float32):
    if y.is_vector():
        if x.Dim[1] == 1:
            raise NotImplementedError("single-column outer product is outside the scoped path (R/matmul.R:479-501)")
        return gemv_csr_vec(x, y)
    return tcrossprod_csr_f32(x, float32(y.Data.T))


# ---- R/matmul.R:545-657 (dense-vector branches) -----------------------------------------------------
def gemv_csr_vec(x: dgRMatrix, y):
    if isinstance(y, sparseVector):
        ylen = y.length
    else:
        ylen = y.Data.size if isinstance(y, float32) else np.asarray(y).size
    if x.Dim[1] != ylen:
        raise ValueError("Matrix-vector dimensions do not match.")
    check_valid_matrix(x)
    nt = _nthreads()
    if isinstance(y, sparseVector):
        # R/matmul.R:595-646.  The reference sorts x and y first (602-603) because its kernel merges two sorted
        # lists; the device kernel tests membership in a bitmap and needs neither, so no sort is done here.
        fn = {"d": rx.matmul_csr_svec_numeric, "i": rx.matmul_csr_svec_integer, "l": rx.matmul_csr_svec_logical}.get(y.kind)
        if fn is not None:
            res = fn(x.p, x.j, x.x, y.i, y.x, nt, ncols=x.Dim[1])
        else:
            res = rx.matmul_csr_svec_binary(x.p, x.j, x.x, y.i, nt, ncols=x.Dim[1])
        return res.reshape(-1, 1)
    if isinstance(y, float32):
        res = rx.matmul_csr_dvec_float32(x.p, x.j, x.x, y.Data, nt)
        return float32(res.reshape(-1, 1))
    y = np.asarray(y)
    if y.dtype == np.bool_:
        raise TypeError("logical vectors must be passed as int32 with NA_LOGICAL = INT_MIN (R's representation)")
    if y.dtype.kind == "f":
        res = rx.matmul_csr_dvec_numeric(x.p, x.j, x.x, y.astype(np.float64, copy=False), nt)
    elif y.dtype.kind in "iu":
        res = rx.matmul_csr_dvec_integer(x.p, x.j, x.x, y.astype(np.int32, copy=False), nt)
    else:
        raise TypeError("unsupported vector type")
    return res.reshape(-1, 1)  # matrix(res, ncol=1), R/matmul.R:652


def gemv_csr_logical(x: dgRMatrix, y_lgl_int32):
    """`%*%`(RsparseMatrix, logical): y as R stores it (int32, NA_LOGICAL = INT_MIN). R/matmul.R:572-579."""
    y = np.ascontiguousarray(y_lgl_int32, dtype=np.int32)
    if x.Dim[1] != y.size:
        raise ValueError("Matrix-vector dimensions do not match.")
    check_valid_matrix(x)
    return rx.matmul_csr_dvec_logical(x.p, x.j, x.x, y, 1).reshape(-1, 1)


# ---- R/matmul.R:746-751 ---------------------------------------------------------------------------
def matmul_csr_vec(x: dgRMatrix, y):
    if x.Dim[1] == 1:
        raise NotImplementedError("single-column outer product is outside the scoped path (R/matmul.R:659-744)")
    return gemv_csr_vec(x, y)


# ---- new methods: signatures the reference leaves to the Matrix package ------------
def crossprod_csr_dense(x: dgRMatrix, y):
    """crossprod(RsparseMatrix, matrix) = t(x) %*% y via the device CSR->CSC transpose."""
    check_dimensions_match(x, y, crossprod=True)
    if isinstance(y, float32):
        return float32(rx.crossprod_csr_dense(x.p, x.j, x.x, x.Dim[1], y.Data, MXG_F32))
    return rx.crossprod_csr_dense(x.p, x.j, x.x, x.Dim[1], _as_double(y), MXG_F64)


def gemm_csc_dense(x: dgCMatrix, y):
    """CsparseMatrix %*% matrix: the CSC of x is the CSR of t(x), so this is crossprod(t_shallow(x), y)."""
    check_dimensions_match(x, y, matmult=True)
    return crossprod_csr_dense(t_shallow(x), y)


def gemm_dense_csr(x, y: dgRMatrix):
    """matrix %*% RsparseMatrix = t(crossprod(y, t(x)))."""
    check_dimensions_match(x, y, matmult=True)
    xt = np.asfortranarray(np.asarray(x.Data if isinstance(x, float32) else x).T)
    res = crossprod_csr_dense(y, float32(xt) if isinstance(x, float32) else xt)
    if isinstance(res, float32):
        return float32(res.Data.T)
    return np.asfortranarray(res.T)


# ---- device-resident left/right operands (rglue/matmul_gpu_methods.R: class gpuRsparse) ----
def _gpu_product(kind, g: gpuRsparse, d):
    """kind: 'A.tD' = g %*% t(d), 'D.tA' = d %*% t(g), 'tA.D' = t(g) %*% d — same preparation as the methods above."""
    nt = _nthreads()
    f32 = isinstance(d, float32)
    dm = d.Data if f32 else _as_double(d)
    fn = {("A.tD", False): rx.gpu_csr_tcrossprod_dense_numeric, ("A.tD", True): rx.gpu_csr_tcrossprod_dense_float32,
          ("tA.D", False): rx.gpu_csr_crossprod_dense_numeric, ("tA.D", True): rx.gpu_csr_crossprod_dense_float32}
    if kind == "D.tA":
        res = (rx.gpu_csr_dense_tcrossprod_float32 if f32 else rx.gpu_csr_dense_tcrossprod_numeric)(dm, g.ptr, nt)
    else:
        res = fn[(kind, f32)](g.ptr, dm, nt)
    return float32(res) if f32 else res


def _t_dense(d):
    return float32(d.Data.T) if isinstance(d, float32) else np.asfortranarray(np.asarray(d).T)


# ---- dispatch ---------------------------------------------------------------------------------------
def matmul(x, y):
    """``x %*% y``."""
    if isinstance(x, gpuRsparse):
        if isinstance(y, np.ndarray) and y.ndim == 1:
            if x.Dim[1] != y.size:
                raise ValueError("Matrix-vector dimensions do not match.")
            return rx.gpu_csr_dvec_numeric(x.ptr, y.astype(np.float64, copy=False), _nthreads()).reshape(-1, 1)
        check_dimensions_match(x, y, matmult=True)
        return _gpu_product("A.tD", x, _t_dense(y))  # `%*%`(Rsparse, matrix) = tcrossprod(x, t(y)), R/matmul.R:463-465
    if isinstance(y, gpuRsparse):  # matrix %*% Rsparse = t(crossprod(y, t(x)))
        check_dimensions_match(x, y, matmult=True)
        return _t_dense(_gpu_product("tA.D", y, _t_dense(x)))
    if _is_dense(x) and isinstance(y, dgCMatrix):
        return gemm_dense_csc(x, y)
    if isinstance(x, float32) and isinstance(y, dgCMatrix):
        return gemm_f32_csc(x, y)
    if isinstance(x, dgRMatrix) and _is_dense(y):
        return gemm_csr_dense(x, y)
    if isinstance(x, dgRMatrix) and isinstance(y, float32):
        return gemm_csr_f32(x, y)
    if isinstance(x, dgRMatrix) and ((isinstance(y, np.ndarray) and y.ndim == 1) or isinstance(y, sparseVector)):
        return matmul_csr_vec(x, y)
    if isinstance(x, dgCMatrix) and (_is_dense(y) or isinstance(y, float32)):
        return gemm_csc_dense(x, y)
    if (_is_dense(x) or isinstance(x, float32)) and isinstance(y, dgRMatrix):
        return gemm_dense_csr(x, y)
    raise TypeError(f"no %*% method for ({type(x).__name__}, {type(y).__name__})")


def crossprod(x, y):
    """``t(x) %*% y``."""
    if isinstance(x, gpuRsparse):
        check_dimensions_match(x, y, crossprod=True)
        return _gpu_product("tA.D", x, y)
    if _is_dense(x) and isinstance(y, dgCMatrix):
        return crossprod_dense_csc(x, y)
    if isinstance(x, float32) and isinstance(y, dgCMatrix):
        return crossprod_f32_csc(x, y)
    if isinstance(x, dgRMatrix) and (_is_dense(y) or isinstance(y, float32)):
        return crossprod_csr_dense(x, y)
    raise TypeError(f"no crossprod method for ({type(x).__name__}, {type(y).__name__})")


def tcrossprod(x, y):
    """``x %*% t(y)``."""
    if isinstance(x, gpuRsparse):
        check_dimensions_match(x, y, tcrossprod=True)
        return _gpu_product("A.tD", x, y)
    if isinstance(y, gpuRsparse):
        check_dimensions_match(x, y, tcrossprod=True)
        return _gpu_product("D.tA", y, x)
    if _is_dense(x) and isinstance(y, dgRMatrix):
        return tcrossprod_dense_csr(x, y)
    if isinstance(x, float32) and isinstance(y, dgRMatrix):
        return tcrossprod_f32_csr(x, y)
    if isinstance(x, dgRMatrix) and _is_dense(y):
        return tcrossprod_csr_dense(x, y)
    if isinstance(x, dgRMatrix) and isinstance(y, float32):
        return tcrossprod_csr_f32(x, y)
    raise TypeError(f"no tcrossprod method for ({type(x).__name__}, {type(y).__name__})")
