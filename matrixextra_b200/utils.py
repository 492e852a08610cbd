"""Python mirror of the index utilities that run before a product on hand-built inputs (R/utils.R:22-161, 439-489;
SURVEY.md §8 f3): ``sort_sparse_indices`` and ``check_sparse_matrix``, on the device library."""
from __future__ import annotations

import numpy as np

from . import rcpp_exports as rx
from .classes import check_valid_matrix, dgCMatrix, dgRMatrix, sparseVector


def sort_sparse_indices(X, copy: bool = False):
    """R/utils.R:22-120.  ``copy=False`` sorts the object's arrays in place (the reference's default)."""
    if isinstance(X, (dgRMatrix, dgCMatrix)):
        check_valid_matrix(X)
        idx_name = "j" if isinstance(X, dgRMatrix) else "i"
        if copy:
            X = type(X)(X.p, getattr(X, idx_name).copy(), X.x.copy(), X.Dim, X.Dimnames)
        rx.sort_sparse_indices_numeric(X.p, getattr(X, idx_name), X.x)
        return X
    if isinstance(X, sparseVector):
        # R/utils.R:96-118: a sparse vector is sorted as a one-row matrix
        if copy:
            X = sparseVector(X.i.copy(), None if X.x is None else X.x.copy(), X.length, X.kind)
        p = np.array([0, X.i.size], dtype=np.int32)
        if X.kind == "d":
            rx.sort_sparse_indices_numeric(p, X.i, X.x)
        elif X.x is None:
            rx.sort_sparse_indices_binary(p, X.i)
        else:  # integer / logical payloads ride along as float64 (exact for int32)
            v = X.x.astype(np.float64)
            rx.sort_sparse_indices_numeric(p, X.i, v)
            X.x[:] = v.astype(np.int32)
        return X
    raise TypeError("sort_sparse_indices: unsupported object")


def check_sparse_matrix(X, sort: bool = True, copy: bool = False):
    """R/utils.R:439-489 for CSR / CSC inputs: validity of the index arrays (device), then optional sorting."""
    if not isinstance(X, (dgRMatrix, dgCMatrix)):
        raise TypeError("check_sparse_matrix: unsupported object")
    check_valid_matrix(X)
    is_csr = isinstance(X, dgRMatrix)
    idx = X.j if is_csr else X.i
    res = rx.check_valid_csr_matrix(X.p, idx, X.Dim[0] if is_csr else X.Dim[1], X.Dim[1] if is_csr else X.Dim[0])
    if res:
        raise ValueError(res["err"])
    if sort and not rx.check_indices_are_unsorted(X.p, idx):
        X = sort_sparse_indices(X, copy=copy)
    return X
