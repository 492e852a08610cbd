"""Minimal Python stand-ins for the R classes on the multiplication path.

* ``dgRMatrix``  — Matrix::dgRMatrix: CSR with slots ``p`` (int32[m+1]), ``j`` (int32, 0-based), ``x`` (float64), ``Dim``.
* ``dgCMatrix``  — Matrix::dgCMatrix: CSC with slots ``p`` (int32[ncol+1]), ``i``, ``x``, ``Dim``.
* ``float32``    — float::float32: ``Data`` is the float32 payload (R keeps the same bits in an integer matrix).

Only what R/matmul.R touches is modelled: slots, ``dim``, shallow transposes (R/trans.R:1-29) and
the coercions ``as_csr_matrix`` / ``as_csc_matrix`` for objects that are already in that format
(R/conversions.R:187-191, 375-379).  Deep conversions go through the device transpose.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class dgRMatrix:
    p: np.ndarray
    j: np.ndarray
    x: np.ndarray
    Dim: Tuple[int, int]
    Dimnames: tuple = field(default_factory=lambda: (None, None))

    def __post_init__(self):
        self.p, self.j, self.x = _i32(self.p), _i32(self.j), _f64(self.x)
        self.Dim = (int(self.Dim[0]), int(self.Dim[1]))

    @property
    def shape(self):
        return self.Dim

    @classmethod
    def from_scipy(cls, a):
        a = a.tocsr()
        return cls(a.indptr, a.indices, a.data, a.shape)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.x, self.j, self.p), shape=self.Dim)


@dataclass
class dgCMatrix:
    p: np.ndarray
    i: np.ndarray
    x: np.ndarray
    Dim: Tuple[int, int]
    Dimnames: tuple = field(default_factory=lambda: (None, None))

    def __post_init__(self):
        self.p, self.i, self.x = _i32(self.p), _i32(self.i), _f64(self.x)
        self.Dim = (int(self.Dim[0]), int(self.Dim[1]))

    @property
    def shape(self):
        return self.Dim

    @classmethod
    def from_scipy(cls, a):
        a = a.tocsc()
        return cls(a.indptr, a.indices, a.data, a.shape)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.x, self.i, self.p), shape=self.Dim)


@dataclass
class float32:
    """float::float32 — ``Data`` holds binary32 values (a vector or a column-major matrix)."""
    Data: np.ndarray

    def __post_init__(self):
        d = np.asarray(self.Data, dtype=np.float32)
        self.Data = np.asfortranarray(d) if d.ndim == 2 else np.ascontiguousarray(d)

    @property
    def shape(self):
        return self.Data.shape

    def is_vector(self) -> bool:
        return self.Data.ndim == 1


@dataclass
class sparseVector:
    """Matrix::sparseVector family: ``i`` 1-based positions (R stores them as integer or double), ``x`` the stored
    values (absent for the pattern class), ``length``.  ``kind`` is the class letter: "d" dsparseVector (double),
    "i" isparseVector (int32, NA = INT_MIN), "l" lsparseVector (int32 0/1, NA = INT_MIN), "n" nsparseVector."""
    i: np.ndarray
    x: Optional[np.ndarray]
    length: int
    kind: str = "d"

    def __post_init__(self):
        self.i = _i32(self.i)  # as.integer(y@i), R/matmul.R:611
        if self.kind == "d":
            self.x = _f64(self.x)
        elif self.kind in ("i", "l"):
            self.x = _i32(self.x)
        elif self.kind == "n":
            self.x = None
        else:
            raise ValueError("sparseVector kind must be one of d, i, l, n")
        self.length = int(self.length)


def t_shallow(x):
    """CSR of X relabelled as the CSC of t(X) and vice versa: no data movement (R/trans.R:1-29, 135-149)."""
    if isinstance(x, dgRMatrix):
        return dgCMatrix(x.p, x.j, x.x, (x.Dim[1], x.Dim[0]), (x.Dimnames[1], x.Dimnames[0]))
    if isinstance(x, dgCMatrix):
        return dgRMatrix(x.p, x.i, x.x, (x.Dim[1], x.Dim[0]), (x.Dimnames[1], x.Dimnames[0]))
    raise TypeError("t_shallow: not a sparse matrix")


def check_valid_matrix(x) -> None:
    """R/utils.R:349-410 for Rsparse/Csparse inputs: slot lengths, p[1]==0, p[n+1]==nnz, same messages.
    No index-range check here (as in the reference); the device library validates indices at upload."""
    if x.Dim[0] < 0:
        raise ValueError("Matrix has invalid number of rows.")
    if x.Dim[1] < 0:
        raise ValueError("Matrix has invalid number of columns.")
    is_csr = isinstance(x, dgRMatrix)
    idx = x.j if is_csr else x.i
    outer = x.Dim[0] if is_csr else x.Dim[1]
    if idx.size != x.x.size:
        raise ValueError("Matrix is invalid (lengths of indices and values differ).")
    if x.p.size - 1 != outer:
        raise ValueError("Matrix is invalid ('p' doesn't match with dimension).")
    if x.p[0] != 0 or x.p[outer] != idx.size:
        raise ValueError("Matrix is invalid ('p' has bad start/end.)")


class gpuRsparse:
    """A dgRMatrix kept in HBM (the S4 class `gpuRsparse` of rglue/matmul_gpu_methods.R): made once
    with ``as_gpu(x)``, multiplied many times — every product moves only the dense operand and the result.
    ``float32=True`` also keeps float32 values so that products with ``float32`` operands are available."""

    def __init__(self, x: "dgRMatrix", float64: bool = True, float32: bool = False):
        from . import rcpp_exports as rx
        check_valid_matrix(x)
        self.Dim = tuple(x.Dim)
        self.Dimnames = x.Dimnames
        self.ptr = rx.as_gpu_csr(x.p, x.j, x.x, x.Dim[1], float64, float32)

    @property
    def shape(self):
        return self.Dim

    def free(self):
        from . import rcpp_exports as rx
        rx.gpu_csr_free(self.ptr)


def as_gpu(x: "dgRMatrix", float64: bool = True, float32: bool = False) -> gpuRsparse:
    return gpuRsparse(x, float64, float32)
