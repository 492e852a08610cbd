/* oracle/ref_driver_ops.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Compiles the REFERENCE's src/operators.cpp in place (found through -I/root/reference/src, nothing
 * copied) into oracle/_ref/libmxref_ops.so and exposes the elementwise products SURVEY.md §8 f4 names:
 * CSR * dense matrix (src/operators.cpp:239-330) and CSR * recycled dense vector (1478, 1501-2143, export
 * 2147-2178; only the Multiply operation is driven here).
 */
#include "operators.cpp"

#include <cstdlib>

typedef Rcpp::IntegerVector IV;
typedef Rcpp::NumericVector NV;
typedef Rcpp::LogicalVector LV;

/* externals of operators.cpp that live in files outside the scoped path (src/slice.cpp) or in R's BLAS: none of
 * the functions driven below reaches them; they only have to exist for the library to load */
double extract_single_val_csr(int *, int *, double *, int, int, bool) { std::abort(); }
int extract_single_val_csr(int *, int *, int *, int, int, bool) { std::abort(); }
extern "C" void daxpy_(const int *n, const double *da, const double *dx, const int *incx, double *dy, const int *incy)
{
    for (ptrdiff_t k = 0; k < (ptrdiff_t)*n; k++) dy[k * *incy] += *da * dx[k * *incx];
}

extern "C" {

/* dense_mat is the column-major nrows x ncols matrix as a flat vector; out receives values_out[nnz] */
int mxref_multiply_csr_by_dense_elemwise_double(const int *p, int nrows, const int *j, const double *x, int nnz,
                                                const double *dense, size_t dense_len, double *out)
{
    NV r = multiply_csr_by_dense_elemwise_double(IV((int *)p, (size_t)nrows + 1), IV((int *)j, (size_t)nnz),
                                                 NV((double *)x, (size_t)nnz), NV((double *)dense, dense_len));
    std::memcpy(out, r.data_ptr(), sizeof(double) * (size_t)nnz);
    return 0;
}

int mxref_multiply_csr_by_dense_elemwise_float32(const int *p, int nrows, const int *j, const double *x, int nnz,
                                                 const int *dense_float_bits, size_t dense_len, double *out)
{
    NV r = multiply_csr_by_dense_elemwise_float32(IV((int *)p, (size_t)nrows + 1), IV((int *)j, (size_t)nnz),
                                                  NV((double *)x, (size_t)nnz), IV((int *)dense_float_bits, dense_len));
    std::memcpy(out, r.data_ptr(), sizeof(double) * (size_t)nnz);
    return 0;
}

int mxref_multiply_csr_by_dense_elemwise_int(const int *p, int nrows, const int *j, const double *x, int nnz,
                                             const int *dense, size_t dense_len, double *out)
{
    NV r = multiply_csr_by_dense_elemwise_int(IV((int *)p, (size_t)nrows + 1), IV((int *)j, (size_t)nnz),
                                              NV((double *)x, (size_t)nnz), IV((int *)dense, dense_len));
    std::memcpy(out, r.data_ptr(), sizeof(double) * (size_t)nnz);
    return 0;
}

int mxref_multiply_csr_by_dense_elemwise_bool(const int *p, int nrows, const int *j, const double *x, int nnz,
                                              const int *dense, size_t dense_len, double *out)
{
    NV r = multiply_csr_by_dense_elemwise_bool(IV((int *)p, (size_t)nrows + 1), IV((int *)j, (size_t)nnz),
                                               NV((double *)x, (size_t)nnz), LV((int *)dense, dense_len));
    std::memcpy(out, r.data_ptr(), sizeof(double) * (size_t)nnz);
    return 0;
}

/* X * dvec with R's recycling over the column-major position row + col*nrows (Multiply, X on the left) */
int mxref_multiply_csr_by_dvec_numeric(const int *p, int nrows, const int *j, const double *x, int nnz,
                                       const double *dvec, size_t dvec_len, int ncols, double *out)
{
    NV r = multiply_csr_by_dvec_no_NAs_numeric(IV((int *)p, (size_t)nrows + 1), IV((int *)j, (size_t)nnz),
                                               NV((double *)x, (size_t)nnz), NV((double *)dvec, dvec_len), ncols,
                                               true, false, false, false, false, true);
    std::memcpy(out, r.data_ptr(), sizeof(double) * (size_t)nnz);
    return 0;
}

} /* extern "C" */
