#!/usr/bin/env python
"""Generates tests/golden/matmul_golden.npz by running the REFERENCE's own src/matmul.cpp (compiled in
place into oracle/_ref/libmxref.so, -O2, no FMA) on seeded inputs.  Run in the build container only
(it needs oracle/_ref, which needs /root/reference); the .npz is committed and travels everywhere.

    python tests/golden/make_golden.py

Cases replay tests/testthat/test-matmul.R of the reference: (100x50).(50x20) at density .4, 1-row and
1-column operands, binary patterns, every dense-vector right-hand-side type incl. NA, plus the literal
4x5 CSR of tests/testthat/test-utilities.R:33-37.  CSR->CSC fixtures come from scipy (the `Matrix`
package is not part of the reference tree), see oracle/mx_oracle.c.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import NA_INT, rsparsematrix  # noqa: E402
from oracle.cpu_oracle import Ref  # noqa: E402


def main():
    ref = Ref(fast=False)
    out = {}
    case_id = 0
    shapes = [(100, 50, 20), (1, 50, 20), (100, 50, 1), (100, 1, 20), (37, 29, 3), (64, 300, 64)]
    for (a, K, b) in shapes:
        for dt, sfx in ((np.float64, "numeric"), (np.float32, "float32")):
            rng = np.random.default_rng(1000 + case_id)
            S = rsparsematrix(b, K, 0.4, 2000 + case_id)          # CSR b x K (== CSC of a K x b matrix)
            X = np.asfortranarray(rng.standard_normal((a, K)).astype(dt))
            pre = f"c{case_id}_"
            out[pre + "kind"] = np.array(sfx)
            out[pre + "p"], out[pre + "j"], out[pre + "x"] = S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data
            out[pre + "X"] = X
            # dense %*% CSC and tcrossprod(dense, CSR) run the same kernel on the same arrays (src/matmul.cpp:188-281)
            out[pre + "matmul_dense_csc"] = getattr(ref, "matmul_dense_csc_" + sfx)(X, S.indptr, S.indices, S.data, 1)
            out[pre + "tcrossprod_dense_csr"] = getattr(ref, "tcrossprod_dense_csr_" + sfx)(X, S.indptr, S.indices, S.data, 1, K)
            if b >= a:  # reference limitation: CSR rows >= dense rows (src/matmul.cpp:176-182)
                out[pre + "tcrossprod_csr_dense"] = getattr(ref, "tcrossprod_csr_dense_" + sfx)(S.indptr, S.indices, S.data, X, 1)
            case_id += 1
    out["n_cases"] = np.array(case_id)

    # SpMV right-hand-side types (tests/testthat/test-matmul.R:134-165)
    A = rsparsematrix(100, 50, 0.4, 77)
    rng = np.random.default_rng(77)
    y = rng.standard_normal(50)
    yi = rng.integers(-5, 6, 50).astype(np.int32)
    yi_na = yi.copy()
    yi_na[[0, 17, 49]] = NA_INT
    yl = (rng.random(50) < 0.5).astype(np.int32)
    yl[3] = 7
    yl_na = yl.copy()
    yl_na[[2, 30]] = NA_INT
    out["v_p"], out["v_j"], out["v_x"] = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    out["v_y"], out["v_yi"], out["v_yi_na"], out["v_yl"], out["v_yl_na"] = y, yi, yi_na, yl, yl_na
    out["v_numeric"] = ref.matmul_csr_dvec_numeric(A.indptr, A.indices, A.data, y, 1)
    out["v_integer"] = ref.matmul_csr_dvec_integer(A.indptr, A.indices, A.data, yi, 1)
    out["v_integer_na"] = ref.matmul_csr_dvec_integer(A.indptr, A.indices, A.data, yi_na, 1)
    out["v_logical"] = ref.matmul_csr_dvec_logical(A.indptr, A.indices, A.data, yl, 1)
    out["v_logical_na"] = ref.matmul_csr_dvec_logical(A.indptr, A.indices, A.data, yl_na, 1)
    out["v_float32"] = ref.matmul_csr_dvec_float32(A.indptr, A.indices, A.data, y.astype(np.float32), 1)

    # literal fixture of tests/testthat/test-utilities.R:33-37 (unsorted row 2) and its CSC
    p = np.array([0, 1, 4, 5, 6], dtype=np.int32)
    j = np.array([4, 2, 1, 4, 1, 0], dtype=np.int32)
    x = np.array([-0.91, 0.14, -0.12, -0.12, 1.1, 0.66])
    out["f_p"], out["f_j"], out["f_x"] = p, j, x
    out["f_spmv"] = ref.matmul_csr_dvec_numeric(p, j, x, np.arange(1.0, 6.0), 1)
    csc = sp.csr_matrix((x, j, p), shape=(4, 5)).tocsc()
    out["f_p2"], out["f_i2"], out["f_x2"] = csc.indptr.astype(np.int32), csc.indices.astype(np.int32), csc.data

    # CSR -> CSC fixtures (scipy's csr_tocsc == the stable counting sort of Matrix/CSparse)
    T = rsparsematrix(300, 257, 0.1, 5)
    Tc = T.tocsc()
    out["t_p"], out["t_j"], out["t_x"] = T.indptr.astype(np.int32), T.indices.astype(np.int32), T.data
    out["t_p2"], out["t_i2"], out["t_x2"] = Tc.indptr.astype(np.int32), Tc.indices.astype(np.int32), Tc.data

    path = os.path.join(HERE, "matmul_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", case_id, "matrix cases")


if __name__ == "__main__":
    main()
