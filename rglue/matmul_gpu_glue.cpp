/* rglue/matmul_gpu_glue.cpp — the Rcpp side of the drop-in: the ten exported entry points of MatrixExtra's
 * multiplication path, same names / argument order / return classes as the reference
 * (/root/reference/src/matmul.cpp:221-483; R callers R/matmul.R:186, 267, 294, 372, 447, 524, 557-587),
 * implemented on the C ABI of include/mxgpu.h instead of the OpenMP kernels.
 *
 * How a maintainer uses it (INTEGRATION.md has the full recipe): drop this file into MatrixExtra's src/,
 * compile the package with -DMATRIXEXTRA_USE_MXGPU (which must #if-out src/matmul.cpp:118-483), link
 * libmxgpu.so, re-run Rcpp::compileAttributes().  RcppExports.{R,cpp} come out identical because the
 * `// [[Rcpp::export(rng = false)]]` signatures below are identical, so R/matmul.R does not change.
 *
 * Contract kept from the reference:
 *   - inputs are borrowed (no copy, never written); the result is a freshly allocated R matrix / vector;
 *   - `nthreads` sets the library's HOST staging threads (narrowing, pageable bounce; 0 = auto); `ncols_Y` is accepted and ignored exactly as
 *     the reference ignores it (src/matmul.cpp:259);
 *   - float32 matrices arrive as IntegerMatrix holding IEEE-754 binary32 bits (src/matmul.cpp:213-214);
 *   - errors surface as R errors: a non-zero status becomes Rcpp::stop(mxg_last_error()) AFTER the C call has
 *     returned (the CUDA library never longjmps or throws across its frames).
 * Differences, all benign: column ids are validated (MXG_ERR_INDEX -> R error instead of the reference's
 * out-of-bounds read, R/utils.R:349-410 does not check them); the result is not zero-filled first.
 *
 * This image has no R / Rcpp, so the file is exercised through tests/glue_driver.cpp, which compiles it
 * against the small Rcpp stand-in of the test infrastructure (oracle/shim/Rcpp.h) and runs it on the GPU.
 */
#if defined(MXGPU_GLUE_SHIM)
#include <Rcpp.h> /* resolved to the stand-in by the test build (-Ioracle/shim) */
#include <stdexcept>
#include <string>
#define MXGPU_GLUE_STOP(msg) throw std::runtime_error(std::string(msg))
#else
#include <Rcpp.h>
#define MXGPU_GLUE_STOP(msg) Rcpp::stop("%s", (msg))
#endif

#include <cstdlib>

#include "mxgpu.h"
#include "mxgpu_result_alloc.h" /* results come from the library's page-locked pool (Rf_allocVector3) */

#define MXGPU_NEW_MATRIX(Type, nr, nc) mxgpu_new_matrix<Type>((nr), (nc))
#define MXGPU_NEW_VECTOR(Type, n) mxgpu_new_vector<Type>((size_t)(n))

namespace {

inline void mxgpu_check(int status)
{
    if (status != MXG_OK) MXGPU_GLUE_STOP(mxg_last_error());
}

/* Every export starts here.  `nthreads` (the reference's OpenMP team size, R/matmul.R:175-180) sizes the library's
 * host staging threads.  Once per session the environment is read:
 *   MATRIXEXTRA_GPUS=n          -> mxg_set_devices(n): one product is spread over n GPUs of the box, in this process
 *   MATRIXEXTRA_GPU_CACHE_MB=c  -> option "cache_mb": device-resident operand cache for repeated products */
inline void mxgpu_enter(int nthreads)
{
    static bool configured = false;
    if (!configured) {
        configured = true;
        const char *gpus = std::getenv("MATRIXEXTRA_GPUS");
        if (gpus && std::atoi(gpus) > 1) mxgpu_check(mxg_set_devices(std::atoi(gpus)));
        const char *cache = std::getenv("MATRIXEXTRA_GPU_CACHE_MB");
        if (cache && std::atol(cache) >= 0) mxgpu_check(mxg_set_option("cache_mb", std::atol(cache)));
    }
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
}

template <class RcppMatrix>
struct ElemType; /* which mxgpu dtype an R matrix class carries */
template <>
struct ElemType<Rcpp::NumericMatrix> {
    static const int dtype = MXG_F64;
    static void *ptr(const Rcpp::NumericMatrix &m) { return (void *)REAL(m); }
};
template <>
struct ElemType<Rcpp::IntegerMatrix> {
    static const int dtype = MXG_F32; /* float32@Data: binary32 bits in an integer matrix */
    static void *ptr(const Rcpp::IntegerMatrix &m) { return (void *)INTEGER(m); }
};

/* Out(a x rows, column-major) = X(a x K) . t(S), S given by rows (a CSR of S == a CSC of t(S)).
 * The reference's gemm_csr_drm_as_drm call of matmul_dense_csc / tcrossprod_dense_csr
 * (src/matmul.cpp:188-219, 254-281): m = rows of S, n = nrow(X), ldb = ldc = nrow(X). */
template <class RcppMatrix>
RcppMatrix dense_times_tsparse(const RcppMatrix &X_colmajor, const Rcpp::IntegerVector &indptr,
                               const Rcpp::IntegerVector &indices, const Rcpp::NumericVector &values)
{
    const int a = X_colmajor.nrow();
    const int K = X_colmajor.ncol();
    const int rows = (int)indptr.size() - 1;
    RcppMatrix out = MXGPU_NEW_MATRIX(RcppMatrix, a, rows);
    const size_t ld = (size_t)(a > 0 ? a : 1);
    mxgpu_check(mxg_spmm_csr_dense(ElemType<RcppMatrix>::dtype, MXG_ROWS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, rows, K, a,
                                   INTEGER(indptr), INTEGER(indices), REAL(values),
                                   ElemType<RcppMatrix>::ptr(X_colmajor), ld, ElemType<RcppMatrix>::ptr(out), ld));
    return out;
}

/* Out(m x n, column-major) = A_csr(m x K) . t(Y), Y (n x K) column-major: the reference's
 * gemm_csr_drm_as_dcm call of tcrossprod_csr_dense (src/matmul.cpp:316-343): ldb = nrow(Y), ldc = m. */
template <class RcppMatrix>
RcppMatrix sparse_times_tdense(const Rcpp::IntegerVector &indptr, const Rcpp::IntegerVector &indices,
                               const Rcpp::NumericVector &values, const RcppMatrix &Y_colmajor)
{
    const int m = (int)indptr.size() - 1;
    const int n = Y_colmajor.nrow();
    const int K = Y_colmajor.ncol();
    RcppMatrix out = MXGPU_NEW_MATRIX(RcppMatrix, m, n);
    mxgpu_check(mxg_spmm_csr_dense(ElemType<RcppMatrix>::dtype, MXG_COLS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, m, K, n,
                                   INTEGER(indptr), INTEGER(indices), REAL(values),
                                   ElemType<RcppMatrix>::ptr(Y_colmajor), (size_t)(n > 0 ? n : 1),
                                   ElemType<RcppMatrix>::ptr(out), (size_t)(m > 0 ? m : 1)));
    return out;
}

} /* namespace */

/* ---- dense %*% CSC : src/matmul.cpp:221-251 ---- */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix matmul_dense_csc_numeric(Rcpp::NumericMatrix X_colmajor, Rcpp::IntegerVector Y_csc_indptr,
                                             Rcpp::IntegerVector Y_csc_indices, Rcpp::NumericVector Y_csc_values,
                                             int nthreads)
{
    mxgpu_enter(nthreads);
    return dense_times_tsparse<Rcpp::NumericMatrix>(X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values);
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix matmul_dense_csc_float32(Rcpp::IntegerMatrix X_colmajor, Rcpp::IntegerVector Y_csc_indptr,
                                             Rcpp::IntegerVector Y_csc_indices, Rcpp::NumericVector Y_csc_values,
                                             int nthreads)
{
    mxgpu_enter(nthreads);
    return dense_times_tsparse<Rcpp::IntegerMatrix>(X_colmajor, Y_csc_indptr, Y_csc_indices, Y_csc_values);
}

/* ---- tcrossprod(dense, CSR) : src/matmul.cpp:283-313 ---- */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix tcrossprod_dense_csr_numeric(Rcpp::NumericMatrix X_colmajor, Rcpp::IntegerVector Y_csr_indptr,
                                                 Rcpp::IntegerVector Y_csr_indices, Rcpp::NumericVector Y_csr_values,
                                                 int nthreads, int ncols_Y)
{
    mxgpu_enter(nthreads);
    (void)ncols_Y;
    return dense_times_tsparse<Rcpp::NumericMatrix>(X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values);
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix tcrossprod_dense_csr_float32(Rcpp::IntegerMatrix X_colmajor, Rcpp::IntegerVector Y_csr_indptr,
                                                 Rcpp::IntegerVector Y_csr_indices, Rcpp::NumericVector Y_csr_values,
                                                 int nthreads, int ncols_Y)
{
    mxgpu_enter(nthreads);
    (void)ncols_Y;
    return dense_times_tsparse<Rcpp::IntegerMatrix>(X_colmajor, Y_csr_indptr, Y_csr_indices, Y_csr_values);
}

/* ---- tcrossprod(CSR, dense) : src/matmul.cpp:345-375 ---- */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix tcrossprod_csr_dense_numeric(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                                 Rcpp::NumericVector X_csr_values, Rcpp::NumericMatrix Y_colmajor,
                                                 int nthreads)
{
    mxgpu_enter(nthreads);
    return sparse_times_tdense<Rcpp::NumericMatrix>(X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor);
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix tcrossprod_csr_dense_float32(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                                 Rcpp::NumericVector X_csr_values, Rcpp::IntegerMatrix Y_colmajor,
                                                 int nthreads)
{
    mxgpu_enter(nthreads);
    return sparse_times_tdense<Rcpp::IntegerMatrix>(X_csr_indptr, X_csr_indices, X_csr_values, Y_colmajor);
}

/* ---- CSR %*% dense vector : src/matmul.cpp:421-483.  K = length(y) (the reference trusts the indices). ---- */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_dvec_numeric(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::NumericVector y_dense, int nthreads)
{
    mxgpu_enter(nthreads);
    const int m = (int)X_csr_indptr.size() - 1;
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, m);
    mxgpu_check(mxg_spmv_csr(MXG_Y_NUMERIC, m, (int)y_dense.size(), INTEGER(X_csr_indptr), INTEGER(X_csr_indices),
                             REAL(X_csr_values), REAL(y_dense), REAL(out)));
    return out;
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_dvec_integer(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_dense, int nthreads)
{
    mxgpu_enter(nthreads);
    const int m = (int)X_csr_indptr.size() - 1;
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, m);
    mxgpu_check(mxg_spmv_csr(MXG_Y_INTEGER, m, (int)y_dense.size(), INTEGER(X_csr_indptr), INTEGER(X_csr_indices),
                             REAL(X_csr_values), INTEGER(y_dense), REAL(out)));
    return out;
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_dvec_logical(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::LogicalVector y_dense, int nthreads)
{
    mxgpu_enter(nthreads);
    const int m = (int)X_csr_indptr.size() - 1;
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, m);
    mxgpu_check(mxg_spmv_csr(MXG_Y_LOGICAL, m, (int)y_dense.size(), INTEGER(X_csr_indptr), INTEGER(X_csr_indices),
                             REAL(X_csr_values), LOGICAL(y_dense), REAL(out)));
    return out;
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerVector matmul_csr_dvec_float32(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_dense, int nthreads)
{
    mxgpu_enter(nthreads);
    const int m = (int)X_csr_indptr.size() - 1;
    Rcpp::IntegerVector out = MXGPU_NEW_VECTOR(Rcpp::IntegerVector, m); /* binary32 bits, like y_dense */
    mxgpu_check(mxg_spmv_csr(MXG_Y_FLOAT32, m, (int)y_dense.size(), INTEGER(X_csr_indptr), INTEGER(X_csr_indices),
                             REAL(X_csr_values), INTEGER(y_dense), INTEGER(out)));
    return out;
}

/* ---- additions: the three signatures MatrixExtra leaves to the Matrix package ------------
 * R side: setMethod("crossprod", signature(x="RsparseMatrix", y="matrix"), ...) etc. */

/* Out(K x n) = t(A_csr(m x K)) . Y(m x n), all R-native column-major; ncols_X = K = ncol(A). */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix crossprod_csr_dense_numeric(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                                Rcpp::NumericVector X_csr_values, int ncols_X,
                                                Rcpp::NumericMatrix Y_colmajor, int nthreads)
{
    mxgpu_enter(nthreads);
    const int m = (int)X_csr_indptr.size() - 1;
    const int n = Y_colmajor.ncol();
    if (Y_colmajor.nrow() != m) MXGPU_GLUE_STOP("Matrix dimensions do not match.");
    Rcpp::NumericMatrix out = MXGPU_NEW_MATRIX(Rcpp::NumericMatrix, ncols_X, n);
    mxgpu_check(mxg_spmm_csrT_dense(MXG_F64, MXG_COLS_CONTIGUOUS, MXG_COLS_CONTIGUOUS, m, ncols_X, n,
                                    INTEGER(X_csr_indptr), INTEGER(X_csr_indices), REAL(X_csr_values),
                                    REAL(Y_colmajor), (size_t)(m > 0 ? m : 1), REAL(out),
                                    (size_t)(ncols_X > 0 ? ncols_X : 1)));
    return out;
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix crossprod_csr_dense_float32(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                                Rcpp::NumericVector X_csr_values, int ncols_X,
                                                Rcpp::IntegerMatrix Y_colmajor, int nthreads)
{
    mxgpu_enter(nthreads);
    const int m = (int)X_csr_indptr.size() - 1;
    const int n = Y_colmajor.ncol();
    if (Y_colmajor.nrow() != m) MXGPU_GLUE_STOP("Matrix dimensions do not match.");
    Rcpp::IntegerMatrix out = MXGPU_NEW_MATRIX(Rcpp::IntegerMatrix, ncols_X, n);
    mxgpu_check(mxg_spmm_csrT_dense(MXG_F32, MXG_COLS_CONTIGUOUS, MXG_COLS_CONTIGUOUS, m, ncols_X, n,
                                    INTEGER(X_csr_indptr), INTEGER(X_csr_indices), REAL(X_csr_values),
                                    INTEGER(Y_colmajor), (size_t)(m > 0 ? m : 1), INTEGER(out),
                                    (size_t)(ncols_X > 0 ? ncols_X : 1)));
    return out;
}

/* Deep CSR -> CSC (the `as(x, "CsparseMatrix")` of R/conversions.R:390-392) as list(p=, i=, x=). */
// [[Rcpp::export(rng = false)]]
Rcpp::List csr_to_csc_gpu(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices, Rcpp::NumericVector values, int ncols)
{
    const int m = (int)indptr.size() - 1;
    const size_t nnz = (size_t)indices.size();
    Rcpp::IntegerVector p2 = MXGPU_NEW_VECTOR(Rcpp::IntegerVector, (size_t)ncols + 1);
    Rcpp::IntegerVector i2 = MXGPU_NEW_VECTOR(Rcpp::IntegerVector, nnz);
    Rcpp::NumericVector x2 = MXGPU_NEW_VECTOR(Rcpp::NumericVector, nnz);
    mxgpu_check(mxg_csr2csc(m, ncols, INTEGER(indptr), INTEGER(indices), REAL(values), INTEGER(p2), INTEGER(i2), REAL(x2)));
    return Rcpp::List::create(Rcpp::_["p"] = p2, Rcpp::_["i"] = i2, Rcpp::_["x"] = x2);
}
