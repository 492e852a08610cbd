"""matrixextra_b200 — B200-native sparse x dense multiplication behind MatrixExtra's interface.

Layout: ``csrc/`` (CUDA kernels + the C ABI of include/mxgpu.h), ``rcpp_exports`` (the reference's
ten Rcpp entry points on top of the C ABI), ``matmul`` (the S4 methods of R/matmul.R), ``device``
(device-resident handles), ``sharded`` (row-block sharding across GPUs).
Importing the package never loads the oracle and never falls back to a CPU implementation.
"""
from .classes import as_gpu, dgCMatrix, dgRMatrix, float32, gpuRsparse, sparseVector, t_shallow  # noqa: F401
from .matmul import crossprod, matmul, tcrossprod  # noqa: F401

__all__ = ["dgRMatrix", "dgCMatrix", "float32", "sparseVector", "t_shallow", "matmul", "crossprod", "tcrossprod", "as_gpu",
           "gpuRsparse"]
