/* rglue/rowops_gpu_glue.cpp — Rcpp glue for the steps either side of the multiplication path:
 * CSR %*% sparseVector, float32 vector %*% CSC, per-row index sorting, CSR validity checks and elementwise CSR * dense products, on
 * libmxgpu.so.  Same rules as rglue/matmul_gpu_glue.cpp: the `// [[Rcpp::export(rng = false)]]` signatures are the
 * reference's (src/matmul.cpp:553-641, src/misc.cpp:177-330, 970-1016, src/operators.cpp:288-322, 2146-2178), so
 * Rcpp::compileAttributes() regenerates identical R wrappers and R/matmul.R, R/utils.R, R/operators.R do not
 * change; inputs are borrowed; a non-zero status becomes an R error after the C call has returned.
 *
 * In the MatrixExtra tree the reference's definitions of these exports are fenced with
 * `#ifndef MATRIXEXTRA_USE_MXGPU`.
 */
#if defined(MXGPU_GLUE_SHIM)
#include <Rcpp.h> /* the stand-in of the test build (-Ioracle/shim) */
#include <stdexcept>
#include <string>
#ifndef MXGPU_GLUE_STOP
#define MXGPU_GLUE_STOP(msg) throw std::runtime_error(std::string(msg))
#endif
#else
#include <Rcpp.h>
#ifndef MXGPU_GLUE_STOP
#define MXGPU_GLUE_STOP(msg) Rcpp::stop("%s", (msg))
#endif
#endif

#include "mxgpu.h"
#include "mxgpu_result_alloc.h" /* results come from the library's page-locked pool (Rf_allocVector3) */

#define MXGPU_NEW_MATRIX(Type, nr, nc) mxgpu_new_matrix<Type>((nr), (nc))
#define MXGPU_NEW_VECTOR(Type, n) mxgpu_new_vector<Type>((size_t)(n))

namespace {

inline void mxgpu_rowops_check(int status)
{
    if (status != MXG_OK) MXGPU_GLUE_STOP(mxg_last_error());
}

/* out = A_csr . y for a sparse vector (1-based indices); the reference's export does not receive ncol(A): K = 0 */
Rcpp::NumericVector csr_times_svec(int ytype, const Rcpp::IntegerVector &indptr, const Rcpp::IntegerVector &indices,
                                   const Rcpp::NumericVector &values, const Rcpp::IntegerVector &y_indices_base1,
                                   const void *y_values, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    const int m = (int)indptr.size() - 1;
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, m);
    mxgpu_rowops_check(mxg_spmv_csr_svec(ytype, m, 0, INTEGER(indptr), INTEGER(indices), REAL(values),
                                         (int)y_indices_base1.size(), INTEGER(y_indices_base1), y_values, REAL(out)));
    return out;
}

/* values * dense[row, col] for a column-major dense matrix passed as a plain vector (the reference infers
 * ncol = length / nrow the same way, src/operators.cpp:239-286) */
Rcpp::NumericVector csr_times_dense_elemwise(int dtype, const Rcpp::IntegerVector &indptr,
                                             const Rcpp::IntegerVector &indices, const Rcpp::NumericVector &values,
                                             const void *dense, size_t dense_len)
{
    const int m = (int)indptr.size() - 1;
    const int K = m > 0 ? (int)(dense_len / (size_t)m) : 0;
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, values.size());
    mxgpu_rowops_check(mxg_mul_csr_dense(dtype, m, K, INTEGER(indptr), INTEGER(indices), REAL(values), dense, REAL(out)));
    return out;
}

} /* namespace */

/* ---- CSR %*% sparseVector : src/matmul.cpp:553-641 ---- */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_svec_numeric(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_indices_base1,
                                            Rcpp::NumericVector y_values, int nthreads)
{
    return csr_times_svec(MXG_Y_NUMERIC, X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, REAL(y_values), nthreads);
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_svec_integer(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_indices_base1,
                                            Rcpp::IntegerVector y_values, int nthreads)
{
    return csr_times_svec(MXG_Y_INTEGER, X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, INTEGER(y_values), nthreads);
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_svec_logical(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_indices_base1,
                                            Rcpp::LogicalVector y_values, int nthreads)
{
    return csr_times_svec(MXG_Y_LOGICAL, X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, LOGICAL(y_values), nthreads);
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_svec_binary(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                           Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_indices_base1, int nthreads)
{
    return csr_times_svec(MXG_Y_BINARY, X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, nullptr, nthreads);
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector matmul_csr_svec_float32(Rcpp::IntegerVector X_csr_indptr, Rcpp::IntegerVector X_csr_indices,
                                            Rcpp::NumericVector X_csr_values, Rcpp::IntegerVector y_indices_base1,
                                            Rcpp::IntegerVector y_values, int nthreads)
{
    return csr_times_svec(MXG_Y_FLOAT32, X_csr_indptr, X_csr_indices, X_csr_values, y_indices_base1, INTEGER(y_values), nthreads);
}

/* ---- float32 (row) vector %*% CSC, tcrossprod(float32 vector, CSR) : src/matmul.cpp:643-684 ----
 * out[col] = sum over the column's entries of values * rowvec[index]: the float32 SpMV with the CSC arrays read as
 * the CSR of the transpose (K = length(rowvec)); result 1 x ncol, float32 bits in an integer matrix. */
// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix matmul_rowvec_by_csc(Rcpp::IntegerVector rowvec_, Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices,
                                         Rcpp::NumericVector values)
{
    const int ncols = (int)indptr.size() - 1;
    Rcpp::IntegerMatrix out(1, ncols);
    mxgpu_rowops_check(mxg_spmv_csr(MXG_Y_FLOAT32, ncols, (int)rowvec_.size(), INTEGER(indptr), INTEGER(indices), REAL(values),
                                    INTEGER(rowvec_), INTEGER(out)));
    return out;
}

/* pattern matrix (ngCMatrix / ngRMatrix): every stored entry counts as 1 */
// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix matmul_rowvec_by_cscbin(Rcpp::IntegerVector rowvec_, Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices)
{
    const int ncols = (int)indptr.size() - 1;
    Rcpp::NumericVector ones(indices.size());
    for (size_t e = 0; e < (size_t)indices.size(); e++) ones[e] = 1.0;
    Rcpp::IntegerMatrix out(1, ncols);
    mxgpu_rowops_check(mxg_spmv_csr(MXG_Y_FLOAT32, ncols, (int)rowvec_.size(), INTEGER(indptr), INTEGER(indices), REAL(ones),
                                    INTEGER(rowvec_), INTEGER(out)));
    return out;
}

/* ---- index sorting and validity : src/misc.cpp:177-189, 300-330, 970-1016 ---- */
/* true when every row is sorted (the reference's name says the opposite of what it returns, src/misc.cpp:161-175);
 * the reference declares `indices` as NumericVector (src/misc.cpp:181) although R passes integers — IntegerVector here */
// [[Rcpp::export(rng = false)]]
bool check_indices_are_unsorted(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices)
{
    int sorted = 0;
    mxgpu_rowops_check(mxg_rows_sorted((int)indptr.size() - 1, INTEGER(indptr), INTEGER(indices), &sorted));
    return sorted != 0;
}

/* in place, like the reference (R/utils.R:22-161 passes a deep copy unless MatrixExtra.inplace_sort is set) */
// [[Rcpp::export(rng = false)]]
void sort_sparse_indices_numeric(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices, Rcpp::NumericVector values)
{
    mxgpu_rowops_check(mxg_sort_csr_indices((int)indptr.size() - 1, INTEGER(indptr), INTEGER(indices), REAL(values)));
}

// [[Rcpp::export(rng = false)]]
void sort_sparse_indices_binary(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices)
{
    mxgpu_rowops_check(mxg_sort_csr_indices((int)indptr.size() - 1, INTEGER(indptr), INTEGER(indices), nullptr));
}

/* list(err = "<the reference's message>") for the first failing check in the reference's order, else an empty list */
// [[Rcpp::export(rng = false)]]
Rcpp::List check_valid_csr_matrix(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices, int nrows, int ncols)
{
    int code = 0;
    mxgpu_rowops_check(mxg_check_valid_csr(nrows, ncols, INTEGER(indptr), INTEGER(indices), (int64_t)indices.size(), &code));
    if (code != 0) return Rcpp::List::create(Rcpp::_["err"] = Rcpp::String(mxg_csr_error_string(code)));
    return Rcpp::List();
}

/* ---- elementwise CSR * dense : src/operators.cpp:288-322, 2146-2178 ---- */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector multiply_csr_by_dense_elemwise_double(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices,
                                                          Rcpp::NumericVector values, Rcpp::NumericVector dense_mat)
{
    return csr_times_dense_elemwise(MXG_Y_NUMERIC, indptr, indices, values, REAL(dense_mat), (size_t)dense_mat.size());
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector multiply_csr_by_dense_elemwise_float32(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices,
                                                           Rcpp::NumericVector values, Rcpp::IntegerVector dense_mat)
{
    return csr_times_dense_elemwise(MXG_Y_FLOAT32, indptr, indices, values, INTEGER(dense_mat), (size_t)dense_mat.size());
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector multiply_csr_by_dense_elemwise_int(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices,
                                                       Rcpp::NumericVector values, Rcpp::IntegerVector dense_mat)
{
    return csr_times_dense_elemwise(MXG_Y_INTEGER, indptr, indices, values, INTEGER(dense_mat), (size_t)dense_mat.size());
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector multiply_csr_by_dense_elemwise_bool(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices,
                                                        Rcpp::NumericVector values, Rcpp::LogicalVector dense_mat)
{
    return csr_times_dense_elemwise(MXG_Y_LOGICAL, indptr, indices, values, LOGICAL(dense_mat), (size_t)dense_mat.size());
}

/* The Multiply case of multiply_csr_by_dvec_no_NAs_numeric (R/operators.R:236-397 calls it for `X * vector` without
 * NAs); the other operator flags stay on the reference's C++ and are refused here. */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector multiply_csr_by_dvec_no_NAs_numeric(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices,
                                                        Rcpp::NumericVector values, Rcpp::NumericVector dvec,
                                                        const int ncols, const bool multiply, const bool powerto,
                                                        const bool divide, const bool divrest, const bool intdiv,
                                                        const bool X_is_LHS)
{
    (void)X_is_LHS;
    if (!multiply || powerto || divide || divrest || intdiv)
        MXGPU_GLUE_STOP("multiply_csr_by_dvec_no_NAs_numeric: only the multiplication is implemented on the GPU path");
    const int m = (int)indptr.size() - 1;
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, values.size());
    mxgpu_rowops_check(mxg_mul_csr_dvec(m, ncols, INTEGER(indptr), INTEGER(indices), REAL(values), REAL(dvec),
                                        (size_t)dvec.size(), REAL(out)));
    return out;
}
