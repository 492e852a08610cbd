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

/* =============================================================================================
 * SURVEY.md §8 "next" rows f2 - f4.  Pinned like the rest: bit-for-bit against the reference's own
 * src/matmul.cpp / src/misc.cpp / src/operators.cpp compiled in place (oracle/_ref/libmxref.so,
 * libmxref_ops.so) in tests/test_oracle.py, and against fixtures generated from those builds.
 * ============================================================================================= */

/* first position in [lo, hi) whose value is >= key (std::lower_bound) */
static const int *mxo_lower_bound(const int *lo, const int *hi, int key)
{
    while (lo < hi) {
        const int *mid = lo + (hi - lo) / 2;
        if (*mid < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

/* ---------------------------------------------------------------------------------------------
 * CSR x sparse vector.  Follows matmul_csr_svec<RcppVector>, src/matmul.cpp:486-551: per row, a
 * merge of the row's (sorted) column ids with the vector's (sorted, 1-based) indices; on a match
 * out[row] += x * y (left to right), otherwise the lagging cursor jumps with lower_bound (537-543).
 * ytype: 0 numeric (double y), 1 integer (NA_INTEGER -> NA_REAL, 523-525), 2 logical (NA -> NA_REAL,
 * else (bool)y, 526-528), 3 float32 (float y widened, 626-641), 4 binary (out += x, 529-530).
 * An empty vector returns the zero-filled result (495-496).  out is cleared here.
 * ------------------------------------------------------------------------------------------- */
void mxo_spmv_svec(int ytype, int m, const int *indptr, const int *indices, const double *values,
                   int ny, const int *yidx_base1, const void *yvals, double *out, int nthreads)
{
    memset(out, 0, sizeof(double) * (size_t)(m > 0 ? m : 0));
    if (ny <= 0) return;
    const int *end_y = yidx_base1 + ny;
    const double na = mxo_na_real();
    int row;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (row = 0; row < m; row++) {
        const int *ptr1 = indices + indptr[row];
        const int *end1 = indices + indptr[row + 1];
        const int *ptr2 = yidx_base1;
        while (ptr1 < end1 && ptr2 < end_y) {
            if (*ptr1 == *ptr2 - 1) {
                const double xv = values[ptr1 - indices];
                const ptrdiff_t k = ptr2 - yidx_base1;
                double term;
                switch (ytype) {
                case 1: { const int v = ((const int *)yvals)[k]; term = v == MXO_NA_INT ? na : xv * v; break; }
                case 2: { const int v = ((const int *)yvals)[k]; term = v == MXO_NA_INT ? na : xv * (v != 0); break; }
                case 3: term = xv * ((const float *)yvals)[k]; break;
                case 4: term = xv; break;
                default: term = xv * ((const double *)yvals)[k]; break;
                }
                out[row] += term;
                ptr1++;
                ptr2++;
            } else if (*ptr2 - 1 > *ptr1) {
                ptr1 = mxo_lower_bound(ptr1, end1, *ptr2 - 1);
            } else {
                ptr2 = mxo_lower_bound(ptr2, end_y, *ptr1 + 1);
            }
        }
    }
    (void)nthreads;
}

/* ---------------------------------------------------------------------------------------------
 * Row sortedness.  check_is_sorted, src/misc.cpp:117-127: non-decreasing (equal neighbours pass).
 * mxo_rows_sorted follows check_indices_are_unsorted(int*, int*, int), src/misc.cpp:161-175, which
 * despite its name returns true when EVERY row is sorted.
 * ------------------------------------------------------------------------------------------- */
static int mxo_is_sorted(const int *v, size_t n)
{
    size_t i;
    if (n <= 1) return 1;
    if (v[n - 1] < v[0]) return 0;
    for (i = 1; i < n; i++)
        if (v[i] < v[i - 1]) return 0;
    return 1;
}

int mxo_rows_sorted(int m, const int *indptr, const int *indices)
{
    int row;
    for (row = 0; row < m; row++)
        if (!mxo_is_sorted(indices + indptr[row], (size_t)(indptr[row + 1] - indptr[row]))) return 0;
    return 1;
}

/* ---------------------------------------------------------------------------------------------
 * In-place per-row sort of (indices, values) by index.  Follows sort_sparse_indices<T>,
 * src/misc.cpp:192-228: rows that are already non-decreasing are left untouched; the others are
 * argsorted by index and permuted.  The reference uses std::sort (not stable): for rows with
 * DISTINCT column ids, the only kind a valid matrix has, the result is unique.  This restatement
 * uses a stable merge sort, i.e. repeated ids keep their stored order.  values may be NULL
 * (pattern overload, src/misc.cpp:230-252).
 * ------------------------------------------------------------------------------------------- */
static void mxo_merge_sort_perm(const int *keys, int *perm, int *tmp, int n)
{
    int width, i;
    for (width = 1; width < n; width *= 2) {
        for (i = 0; i < n; i += 2 * width) {
            int lo = i, mid = i + width < n ? i + width : n, hi = i + 2 * width < n ? i + 2 * width : n;
            int a = lo, b = mid, k = lo;
            while (a < mid && b < hi) tmp[k++] = keys[perm[b]] < keys[perm[a]] ? perm[b++] : perm[a++];
            while (a < mid) tmp[k++] = perm[a++];
            while (b < hi) tmp[k++] = perm[b++];
        }
        memcpy(perm, tmp, sizeof(int) * (size_t)n);
    }
}

int mxo_sort_sparse_indices(int m, const int *indptr, int *indices, double *values)
{
    int row, cap = 0, k;
    int *perm = NULL, *tmp = NULL, *ibuf = NULL;
    double *xbuf = NULL;
    for (row = 0; row < m; row++) {
        const int a = indptr[row], n = indptr[row + 1] - indptr[row];
        if (n <= 0 || mxo_is_sorted(indices + a, (size_t)n)) continue;
        if (n > cap) {
            cap = n;
            perm = (int *)realloc(perm, sizeof(int) * (size_t)cap);
            tmp = (int *)realloc(tmp, sizeof(int) * (size_t)cap);
            ibuf = (int *)realloc(ibuf, sizeof(int) * (size_t)cap);
            xbuf = (double *)realloc(xbuf, sizeof(double) * (size_t)cap);
            if (!perm || !tmp || !ibuf || !xbuf) return 1;
        }
        for (k = 0; k < n; k++) perm[k] = k;
        mxo_merge_sort_perm(indices + a, perm, tmp, n);
        for (k = 0; k < n; k++) ibuf[k] = indices[a + perm[k]];
        memcpy(indices + a, ibuf, sizeof(int) * (size_t)n);
        if (values) {
            for (k = 0; k < n; k++) xbuf[k] = values[a + perm[k]];
            memcpy(values + a, xbuf, sizeof(double) * (size_t)n);
        }
    }
    free(perm); free(tmp); free(ibuf); free(xbuf);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * CSR validity.  Follows check_valid_csr_matrix, src/misc.cpp:970-1016, returning the ordinal of the
 * FIRST failing check in the reference's order (0 = valid):
 *   1 "Matrix has negative indices."             (min index < 0; NA_INTEGER = INT_MIN lands here too)
 *   2 "Matrix has invalid column indices."       (max index >= ncols)
 *   3 "Matrix has indices with missing values."  (unreachable after 1, kept for the numbering)
 *   4 "Matrix has missing values in the index pointer."
 *   5 "Matrix index pointer is not monotonicaly increasing."
 * ------------------------------------------------------------------------------------------- */
int mxo_check_valid_csr(int m, int ncols, const int *indptr, const int *indices, size_t nnz)
{
    size_t e;
    int r;
    if (nnz > 0) {
        int imin = indices[0], imax = indices[0];
        for (e = 1; e < nnz; e++) {
            if (indices[e] < imin) imin = indices[e];
            if (indices[e] > imax) imax = indices[e];
        }
        if (imin < 0) return 1;
        if (imax >= ncols) return 2;
    }
    for (e = 0; e < nnz; e++)
        if (indices[e] == MXO_NA_INT) return 3;
    for (r = 0; r <= m; r++)
        if (indptr[r] == MXO_NA_INT) return 4;
    for (r = 0; r < m; r++)
        if (indptr[r] > indptr[r + 1]) return 5;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Elementwise CSR * dense matrix.  Follows multiply_csr_by_dense_elemwise<NumericVector, *>,
 * src/operators.cpp:239-289: values_out[e] = values[e] * dense[row + nrows * indices[e]] with the
 * dense matrix column-major; dtype 0 double, 1 float32 (widened), 2 integer (NA_INTEGER -> NA_REAL,
 * 263-265), 3 logical (NA_LOGICAL -> NA_REAL, else (bool), 260-262).  One IEEE multiply per entry.
 * ------------------------------------------------------------------------------------------- */
void mxo_multiply_csr_by_dense(int dtype, int m, const int *indptr, const int *indices, const double *values,
                               const void *dense, double *values_out)
{
    const size_t nrows = (size_t)m;
    const double na = mxo_na_real();
    size_t row;
    int el;
    for (row = 0; row < nrows; row++) {
        for (el = indptr[row]; el < indptr[row + 1]; el++) {
            const size_t pos = row + nrows * (size_t)indices[el];
            switch (dtype) {
            case 1: values_out[el] = values[el] * ((const float *)dense)[pos]; break;
            case 2: { const int v = ((const int *)dense)[pos]; values_out[el] = v == MXO_NA_INT ? na : values[el] * v; break; }
            case 3: { const int v = ((const int *)dense)[pos]; values_out[el] = v == MXO_NA_INT ? na : values[el] * (v != 0); break; }
            default: values_out[el] = values[el] * ((const double *)dense)[pos]; break;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Elementwise CSR * recycled dense vector (Multiply).  Follows multiply_csr_by_dvec_no_NAs,
 * src/operators.cpp:1501-2143 with op == Multiply: the vector is recycled along the column-major
 * position of the entry.  The reference's four cases are one formula:
 *   len == nrows              -> dvec[row]                         (1531-1545)
 *   len >= nrows * ncols      -> dvec[row + col * nrows]           (1771-1790)
 *   len <  nrows, nrows % len == 0 -> dvec[row % len]              (1871-1888)
 *   otherwise                 -> dvec[(row + col * nrows) % len]   (recyle_pos, 1478, 2031-2041)
 * ------------------------------------------------------------------------------------------- */
void mxo_multiply_csr_by_dvec(int m, int ncols, const int *indptr, const int *indices, const double *values,
                              const double *dvec, size_t len, double *values_out)
{
    const unsigned long long nrows = (unsigned long long)m;
    const int direct = (unsigned long long)len >= nrows * (unsigned long long)ncols;
    unsigned long long row;
    int el;
    for (row = 0; row < nrows; row++) {
        for (el = indptr[row]; el < indptr[row + 1]; el++) {
            unsigned long long pos = row + (unsigned long long)indices[el] * nrows;
            if (!direct) pos %= (unsigned long long)len;
            values_out[el] = values[el] * dvec[pos];
        }
    }
}
