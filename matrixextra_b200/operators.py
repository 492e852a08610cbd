"""Python mirror of the elementwise `*` methods between a CSR matrix and a dense operand (R/operators.R:236-397,
SURVEY.md §8 f4) — the operation the vignette's gradient uses, ``X * as.numeric(pred - y)``.

Scope: the products over the STORED entries of the sparse operand (``multiply_csr_by_dense_elemwise_*``,
``multiply_csr_by_dvec_no_NAs_numeric``), i.e. what the reference computes with
``options(MatrixExtra.ignore_na = TRUE)`` and for NA-free dense operands.  The default ``keep_NAs`` branch of
R/operators.R:252-330, which afterwards INSERTS new entries where the dense operand is NA/NaN/Inf and the sparse one is
empty (``add_NAs_from_dense_after_elemenwise_mult_*``), changes the sparsity pattern on the host and stays on the
reference's C++; ``multiply`` raises if the dense operand holds non-finite values and ``ignore_na`` is not set.
"""
from __future__ import annotations

import numpy as np

from . import rcpp_exports as rx
from .classes import check_valid_matrix, dgRMatrix, float32

options = {"MatrixExtra.ignore_na": False}


def _has_nonfinite(e2) -> bool:
    a = e2.Data if isinstance(e2, float32) else np.asarray(e2)
    if a.dtype.kind == "f":
        return not bool(np.isfinite(a).all())
    return bool((a == np.iinfo(np.int32).min).any())


def multiply_csr_by_dense(e1: dgRMatrix, e2) -> dgRMatrix:
    """``e1 * e2`` for a dense matrix of the same shape (R/operators.R:236-330)."""
    d = e2.Data if isinstance(e2, float32) else np.asarray(e2)
    if d.ndim != 2:
        return multiply_csr_by_dvec(e1, e2)
    if e1.Dim[0] != d.shape[0] or e1.Dim[1] != d.shape[1]:
        raise ValueError("Matrices must have the same dimensions in order to multiply them.")
    if not options.get("MatrixExtra.ignore_na", False) and _has_nonfinite(e2):
        raise NotImplementedError("dense operand holds NA/NaN/Inf: the pattern-changing keep_NAs branch is outside the scoped path")
    check_valid_matrix(e1)
    if isinstance(e2, float32):
        res = rx.multiply_csr_by_dense_elemwise_float32(e1.p, e1.j, e1.x, d)
    elif d.dtype.kind == "f":
        res = rx.multiply_csr_by_dense_elemwise_double(e1.p, e1.j, e1.x, d.astype(np.float64, copy=False))
    elif d.dtype == np.bool_:
        res = rx.multiply_csr_by_dense_elemwise_bool(e1.p, e1.j, e1.x, d.astype(np.int32))
    elif d.dtype.kind in "iu":
        res = rx.multiply_csr_by_dense_elemwise_int(e1.p, e1.j, e1.x, d.astype(np.int32, copy=False))
    else:
        raise TypeError("unsupported dense operand")
    return dgRMatrix(e1.p, e1.j, res, e1.Dim, e1.Dimnames)  # out <- e1; out@x <- res (R/operators.R:284-285)


def multiply_csr_by_dvec(e1: dgRMatrix, e2) -> dgRMatrix:
    """``e1 * v`` with R's recycling of the vector down the columns (R/operators.R:332-397)."""
    v = np.asarray(e2.Data if isinstance(e2, float32) else e2, dtype=np.float64).reshape(-1)
    if v.size == 0:
        raise ValueError("empty vector")
    if v.size > e1.Dim[0] * e1.Dim[1]:
        raise ValueError("Vector to multiply with has more entries than matrix dimensions.")
    if not options.get("MatrixExtra.ignore_na", False) and _has_nonfinite(v):
        raise NotImplementedError("vector holds NA/NaN/Inf: the pattern-changing keep_NAs branch is outside the scoped path")
    check_valid_matrix(e1)
    res = rx.multiply_csr_by_dvec_no_NAs_numeric(e1.p, e1.j, e1.x, v, e1.Dim[1], True, False, False, False, False, True)
    return dgRMatrix(e1.p, e1.j, res, e1.Dim, e1.Dimnames)


def multiply(e1, e2):
    """``e1 * e2`` dispatch for (RsparseMatrix, dense) in either order (multiplication commutes: X_is_LHS only
    matters for the other operators of the reference's template)."""
    if isinstance(e1, dgRMatrix):
        return multiply_csr_by_dense(e1, e2)
    if isinstance(e2, dgRMatrix):
        return multiply_csr_by_dense(e2, e1)
    raise TypeError("multiply: one operand must be a dgRMatrix")
