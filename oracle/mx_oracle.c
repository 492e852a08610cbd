/* oracle/mx_oracle.c — CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the algorithm on MatrixExtra's sparse x dense multiplication path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product library (matrixextra_b200/csrc) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 *   (1) bit-for-bit against the reference's own src/matmul.cpp compiled in place into
 *       oracle/_ref/libmxref.so (same -O2, no FMA contraction) whenever that library is present,
 *   (2) against the committed fixtures tests/golden/*.npz that were produced by that reference
 *       build (tests/golden/make_golden.py), and
 *   (3) against the only literal CSR fixture in the reference's test-suite
 *       (tests/testthat/test-utilities.R:33-37) and the algebraic dense product its testthat
 *       cases assert (tests/testthat/test-matmul.R:12-165).
 * mxo_csr2csc restates an algorithm that is NOT in /root/reference: MatrixExtra delegates
 * CSR->CSC to the `Matrix` package (DESCRIPTION:34 "Matrix (>= 1.3)"; call sites
 * R/conversions.R:248-250, 390-392).  Matrix/CSparse use the textbook stable counting sort
 * (count per column, exclusive scan, scatter rows in ascending order); it is pinned here against
 * scipy.sparse csr->csc, which implements the same algorithm, in tests/test_oracle.py.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off so that a*b+c is never fused: this matches
 * an R default build (-O2, generic x86-64) of the reference).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MXO_NA_INT INT_MIN

static double mxo_na_real(void)
{
    /* R's NA_real_: NaN with low word 1954 */
    const uint64_t bits = 0x7FF00000000007A2ULL;
    double out;
    memcpy(&out, &bits, sizeof(out));
    return out;
}

int mxo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * Row-major-output gather product.  Follows gemm_csr_drm_as_drm<real_t>, src/matmul.cpp:118-142:
 *   Out[row, 0:n] += sum over stored entries (in stored order) of (real_t)values[e] * B[indices[e], 0:n]
 * Out is NOT cleared here (the reference relies on the freshly zero-filled R matrix, 197/261).
 * fp32: the double CSR value is first narrowed to float (src/matmul.cpp:53-57), then a float
 * multiply and a float add per element (saxpy, 17-33).
 * ------------------------------------------------------------------------------------------- */
void mxo_gemm_csr_drm_as_drm_f64(int m, int n, const int *indptr, const int *indices,
                                 const double *values, const double *B, size_t ldb,
                                 double *Out, size_t ldc, int nthreads)
{
    if (m <= 0 || indptr[0] == indptr[m]) return; /* src/matmul.cpp:128-129 */
    (void)nthreads;
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        double *dst = Out + (size_t)row * ldc;
        for (int e = indptr[row]; e < indptr[row + 1]; e++) {
            const double a = values[e];
            const double *src = B + (size_t)indices[e] * ldb;
            for (int c = 0; c < n; c++) dst[c] += a * src[c];
        }
    }
}

void mxo_gemm_csr_drm_as_drm_f32(int m, int n, const int *indptr, const int *indices,
                                 const double *values, const float *B, size_t ldb,
                                 float *Out, size_t ldc, int nthreads)
{
    if (m <= 0 || indptr[0] == indptr[m]) return;
    (void)nthreads;
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        float *dst = Out + (size_t)row * ldc;
        for (int e = indptr[row]; e < indptr[row + 1]; e++) {
            const float a = (float)values[e]; /* narrowing first: src/matmul.cpp:55 */
            const float *src = B + (size_t)indices[e] * ldb;
            for (int c = 0; c < n; c++) dst[c] += a * src[c];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Column-major-output gather product.  Follows gemm_csr_drm_as_dcm<real_t>, src/matmul.cpp:150-185:
 * per non-empty row, a scratch row of n elements is cleared, accumulated exactly like above, then
 * copied with stride ldc into Out[row + c*ldc].  Empty rows are skipped (172), so they keep whatever
 * Out held (zeros from the caller).  The scratch here has n elements (the reference allocates ldc
 * elements, which is its m>=n limitation; results are identical whenever the reference is valid).
 * ------------------------------------------------------------------------------------------- */
void mxo_gemm_csr_drm_as_dcm_f64(int m, int n, const int *indptr, const int *indices,
                                 const double *values, const double *B, size_t ldb,
                                 double *Out, size_t ldc, int nthreads)
{
    if (m <= 0 || indptr[0] == indptr[m]) return; /* src/matmul.cpp:160-161 */
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        double *scratch = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        int row;
#ifdef _OPENMP
#pragma omp for schedule(dynamic)
#endif
        for (row = 0; row < m; row++) {
            if (indptr[row] >= indptr[row + 1]) continue;
            memset(scratch, 0, sizeof(double) * (size_t)n);
            for (int e = indptr[row]; e < indptr[row + 1]; e++) {
                const double a = values[e];
                const double *src = B + (size_t)indices[e] * ldb;
                for (int c = 0; c < n; c++) scratch[c] += a * src[c];
            }
            for (int c = 0; c < n; c++) Out[(size_t)row + (size_t)c * ldc] = scratch[c];
        }
        free(scratch);
    }
}

void mxo_gemm_csr_drm_as_dcm_f32(int m, int n, const int *indptr, const int *indices,
                                 const double *values, const float *B, size_t ldb,
                                 float *Out, size_t ldc, int nthreads)
{
    if (m <= 0 || indptr[0] == indptr[m]) return;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        float *scratch = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        int row;
#ifdef _OPENMP
#pragma omp for schedule(dynamic)
#endif
        for (row = 0; row < m; row++) {
            if (indptr[row] >= indptr[row + 1]) continue;
            memset(scratch, 0, sizeof(float) * (size_t)n);
            for (int e = indptr[row]; e < indptr[row + 1]; e++) {
                const float a = (float)values[e];
                const float *src = B + (size_t)indices[e] * ldb;
                for (int c = 0; c < n; c++) scratch[c] += a * src[c];
            }
            for (int c = 0; c < n; c++) Out[(size_t)row + (size_t)c * ldc] = scratch[c];
        }
        free(scratch);
    }
}

/* ---------------------------------------------------------------------------------------------
 * CSR x dense vector.  Follows matmul_csr_dvec<>, src/matmul.cpp:381-419.
 *   numeric : val(double) += x[e] * y[j[e]]
 *   integer : y == NA_INTEGER contributes NA_REAL (406-408), else x[e] * (double)y
 *   logical : y == NA_LOGICAL contributes NA_REAL (409-411), else x[e] * (bool)y
 *   float32 : accumulator is float; each term is the DOUBLE product x[e] * (double)y[j[e]] added
 *             in double to the float accumulator and narrowed back (usual arithmetic conversions
 *             on `float += double * float`, 403/413 with OutputDType=float, 476).
 * ------------------------------------------------------------------------------------------- */
void mxo_spmv_numeric(int m, const int *indptr, const int *indices, const double *values,
                      const double *y, double *out, int nthreads)
{
    (void)nthreads;
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        double val = 0;
        for (int e = indptr[row]; e < indptr[row + 1]; e++) val += values[e] * y[indices[e]];
        out[row] = val;
    }
}

void mxo_spmv_integer(int m, const int *indptr, const int *indices, const double *values,
                      const int *y, double *out, int nthreads)
{
    (void)nthreads;
    const double na = mxo_na_real();
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        double val = 0;
        for (int e = indptr[row]; e < indptr[row + 1]; e++) {
            const int yv = y[indices[e]];
            val += (yv == MXO_NA_INT) ? na : values[e] * yv;
        }
        out[row] = val;
    }
}

void mxo_spmv_logical(int m, const int *indptr, const int *indices, const double *values,
                      const int *y, double *out, int nthreads)
{
    (void)nthreads;
    const double na = mxo_na_real();
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        double val = 0;
        for (int e = indptr[row]; e < indptr[row + 1]; e++) {
            const int yv = y[indices[e]];
            val += (yv == MXO_NA_INT) ? na : values[e] * (yv != 0);
        }
        out[row] = val;
    }
}

void mxo_spmv_float32(int m, const int *indptr, const int *indices, const double *values,
                      const float *y, float *out, int nthreads)
{
    (void)nthreads;
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        float val = 0;
        for (int e = indptr[row]; e < indptr[row + 1]; e++)
            val = (float)((double)val + values[e] * (double)y[indices[e]]);
        out[row] = val;
    }
}

/* ---------------------------------------------------------------------------------------------
 * CSR(m x K) -> CSC(m x K): the deep conversion the reference delegates to the Matrix package
 * (R/conversions.R:390-392 `as(x, "CsparseMatrix")`).  Stable counting sort: within every column
 * the row indices come out ascending, duplicates keep their stored order.  Integer work only.
 * Returns 0, or 1 if an index is out of [0, K).
 * ------------------------------------------------------------------------------------------- */
int mxo_csr2csc(int m, int K, const int *indptr, const int *indices, const double *values,
                int *p2, int *i2, double *x2)
{
    const int nnz0 = indptr[0], nnz1 = indptr[m];
    memset(p2, 0, sizeof(int) * ((size_t)K + 1));
    for (int e = nnz0; e < nnz1; e++) {
        if (indices[e] < 0 || indices[e] >= K) return 1;
        p2[indices[e] + 1]++;
    }
    for (int c = 0; c < K; c++) p2[c + 1] += p2[c];
    int *cursor = (int *)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
    memcpy(cursor, p2, sizeof(int) * (size_t)K);
    for (int r = 0; r < m; r++) {
        for (int e = indptr[r]; e < indptr[r + 1]; e++) {
            const int dst = cursor[indices[e]]++;
            i2[dst] = r;
            if (values) x2[dst] = values[e];
        }
    }
    free(cursor);
    return 0;
}
