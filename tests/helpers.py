"""Shared test helpers: seeded inputs in the layouts R hands to the Rcpp exports, error metrics."""
import numpy as np
import scipy.sparse as sp

NA_INT = np.iinfo(np.int32).min
FP64_TOL = 1e-12  # north_star: relative 1e-12 for fp64 (reassociated row sums)
FP32_TOL = 1e-5   # north_star: relative 1e-5 for fp32 (tests/testthat/test-matmul.R uses tolerance=1e-5 too)


def rsparsematrix(m, n, density, seed, fmt="csr"):
    """Like Matrix::rsparsematrix(m, n, density): random pattern, N(0,1) values."""
    rng = np.random.default_rng(seed)
    a = sp.random(m, n, density=density, format=fmt, random_state=rng, data_rvs=rng.standard_normal,
                  dtype=np.float64)
    a.sort_indices()
    return a


def powerlaw_csr(m, K, mean_len, seed, cap=None, alpha=1.5):
    """Rows with Pareto lengths (some far above the long-row piece size), sorted unique columns."""
    rng = np.random.default_rng(seed)
    cap = cap or K
    lmin = mean_len * (alpha - 1) / alpha
    lens = np.minimum(cap, np.floor(lmin * (1 - rng.random(m)) ** (-1 / alpha))).astype(np.int64)
    lens = np.minimum(lens, K)
    p = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(lens, out=p[1:])
    j = np.empty(p[-1], dtype=np.int32)
    for r in range(m):
        if lens[r]:
            j[p[r]:p[r + 1]] = np.sort(rng.choice(K, size=lens[r], replace=False))
    x = rng.uniform(-1, 1, size=p[-1])
    return p, j, x


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.size == 0:
        return 0.0
    denom = np.max(np.abs(want))
    if denom == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - want)) / denom)


def all_equal_style(got, want):
    """testthat::expect_equal's metric: mean relative difference."""
    got = np.asarray(got, dtype=np.float64).ravel()
    want = np.asarray(want, dtype=np.float64).ravel()
    s = np.sum(np.abs(want))
    return float(np.sum(np.abs(got - want)) / s) if s > 0 else float(np.sum(np.abs(got - want)))
