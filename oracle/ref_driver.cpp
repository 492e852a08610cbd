/* oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Builds the REFERENCE's own implementation of the hot path into oracle/_ref/libmxref*.so:
 * this translation unit #includes /root/reference/src/matmul.cpp where it lies (found through
 * -I/root/reference/src; no reference source is copied into this repository).  That file's
 * `#include "MatrixExtra.h"` resolves to the reference's real header next to it, whose
 * <Rcpp.h>/<R.h>/<R_ext/BLAS.h> includes are satisfied by the stand-ins under oracle/shim/.
 *
 * What is added here:
 *   - daxpy_/dcopy_ as plain reference-BLAS loops (R's BLAS is a third-party, un-pinned dependency:
 *     R_ext/BLAS.h, call sites src/matmul.cpp:45,50,72), with 64-bit index products;
 *   - SafeRcppVector (src/misc.cpp:3-62 in the reference) so the out-of-scope tail of matmul.cpp links;
 *   - extern "C" drivers with plain pointers for ctypes. Each driver wraps the caller's buffers in
 *     non-owning shim vectors (as Rcpp borrows a SEXP), calls the reference's exported function
 *     with the exact signature of src/matmul.cpp:221-483, and exposes the returned matrix/vector.
 *
 * Known reference limit the callers must respect: gemm_csr_drm_as_dcm allocates its scratch row
 * with new real_t[ldc] (ldc = CSR rows m) but uses n elements (src/matmul.cpp:176-182) => keep
 * m >= n when calling the tcrossprod_csr_dense_* drivers.
 */
#include "matmul.cpp" /* the reference's file, compiled in place (see Makefile: -I$(REF)/src) */

#include <cstdlib>

extern "C" {

void daxpy_(const int *n, const double *da, const double *dx, const int *incx,
            double *dy, const int *incy)
{
    const size_t n_ = (size_t)(*n > 0 ? *n : 0);
    const double a = *da;
    const ptrdiff_t ix = *incx, iy = *incy;
    if (ix == 1 && iy == 1) {
        for (size_t k = 0; k < n_; k++) dy[k] += a * dx[k];
    } else {
        for (size_t k = 0; k < n_; k++) dy[(ptrdiff_t)k * iy] += a * dx[(ptrdiff_t)k * ix];
    }
}

void dcopy_(const int *n, const double *dx, const int *incx, double *dy, const int *incy)
{
    const size_t n_ = (size_t)(*n > 0 ? *n : 0);
    const ptrdiff_t ix = *incx, iy = *incy;
    for (size_t k = 0; k < n_; k++) dy[(ptrdiff_t)k * iy] = dx[(ptrdiff_t)k * ix];
}

} /* extern "C" */

/* src/misc.cpp:3-62 builds an R vector from a std::vector; the shim SEXP just carries the data. */
SEXP SafeRcppVector(void *args_)
{
    VectorConstructorArgs *args = (VectorConstructorArgs *)args_;
    SEXP out = new mx_shim_sexprec();
    if (args->as_integer) {
        out->is_int = true;
        if (args->from_cpp_vec) out->ints = *(std::vector<int> *)args->int_vec_from;
        else out->ints.assign(args->size, 0);
    } else {
        if (args->from_cpp_vec) out->dbls = *(std::vector<double> *)args->num_vec_from;
        else out->dbls.assign(args->size, 0.0);
    }
    return out;
}

namespace {

/* keeps the last returned Rcpp object alive so Python can read it without a copy */
struct Holder {
    std::shared_ptr<void> keep;
    void *data = nullptr;
    size_t nrow = 0, ncol = 0;
};
Holder g_last;

template <class M>
void hold_matrix(const M &m)
{
    g_last.keep = std::static_pointer_cast<void>(m.keepalive());
    g_last.data = (void *)m.data_ptr();
    g_last.nrow = (size_t)m.nrow();
    g_last.ncol = (size_t)m.ncol();
}

template <class V>
void hold_vector(const V &v)
{
    g_last.keep = std::static_pointer_cast<void>(v.keepalive());
    g_last.data = (void *)v.data_ptr();
    g_last.nrow = (size_t)v.size();
    g_last.ncol = 1;
}

template <class T>
void maybe_copy(void *out)
{
    if (out) std::memcpy(out, g_last.data, g_last.nrow * g_last.ncol * sizeof(T));
}

typedef Rcpp::IntegerVector IV;
typedef Rcpp::NumericVector NV;
typedef Rcpp::LogicalVector LV;
typedef Rcpp::NumericMatrix NM;
typedef Rcpp::IntegerMatrix IM;

} /* namespace */

extern "C" {

int mxref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

const void *mxref_last_result(size_t *nrow, size_t *ncol)
{
    if (nrow) *nrow = g_last.nrow;
    if (ncol) *ncol = g_last.ncol;
    return g_last.data;
}

void mxref_release(void) { g_last = Holder(); }

/* ---- src/matmul.cpp:221-251 : dense %*% CSC ---- */
int mxref_matmul_dense_csc_numeric(const double *X, int nrowX, int ncolX,
                                   const int *p, int ncolsY, const int *i, const double *x, int nnz,
                                   int nthreads, double *out)
{
    hold_matrix(matmul_dense_csc_numeric(NM((double *)X, nrowX, ncolX), IV((int *)p, (size_t)ncolsY + 1),
                                         IV((int *)i, (size_t)nnz), NV((double *)x, (size_t)nnz), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_dense_csc_float32(const int *X, int nrowX, int ncolX,
                                   const int *p, int ncolsY, const int *i, const double *x, int nnz,
                                   int nthreads, int *out)
{
    hold_matrix(matmul_dense_csc_float32(IM((int *)X, nrowX, ncolX), IV((int *)p, (size_t)ncolsY + 1),
                                         IV((int *)i, (size_t)nnz), NV((double *)x, (size_t)nnz), nthreads));
    maybe_copy<int>(out);
    return 0;
}

/* ---- src/matmul.cpp:283-313 : tcrossprod(dense, CSR) ---- */
int mxref_tcrossprod_dense_csr_numeric(const double *X, int nrowX, int ncolX,
                                       const int *p, int nrowsY, const int *j, const double *x, int nnz,
                                       int nthreads, int ncols_Y, double *out)
{
    hold_matrix(tcrossprod_dense_csr_numeric(NM((double *)X, nrowX, ncolX), IV((int *)p, (size_t)nrowsY + 1),
                                             IV((int *)j, (size_t)nnz), NV((double *)x, (size_t)nnz),
                                             nthreads, ncols_Y));
    maybe_copy<double>(out);
    return 0;
}

int mxref_tcrossprod_dense_csr_float32(const int *X, int nrowX, int ncolX,
                                       const int *p, int nrowsY, const int *j, const double *x, int nnz,
                                       int nthreads, int ncols_Y, int *out)
{
    hold_matrix(tcrossprod_dense_csr_float32(IM((int *)X, nrowX, ncolX), IV((int *)p, (size_t)nrowsY + 1),
                                             IV((int *)j, (size_t)nnz), NV((double *)x, (size_t)nnz),
                                             nthreads, ncols_Y));
    maybe_copy<int>(out);
    return 0;
}

/* ---- src/matmul.cpp:345-375 : tcrossprod(CSR, dense).  Caller keeps nrowsX >= nrowY (see header). ---- */
int mxref_tcrossprod_csr_dense_numeric(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                       const double *Y, int nrowY, int ncolY, int nthreads, double *out)
{
    if (nrowsX < nrowY) return 2; /* would overflow the reference's scratch row */
    hold_matrix(tcrossprod_csr_dense_numeric(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                             NV((double *)x, (size_t)nnz), NM((double *)Y, nrowY, ncolY),
                                             nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_tcrossprod_csr_dense_float32(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                       const int *Y, int nrowY, int ncolY, int nthreads, int *out)
{
    if (nrowsX < nrowY) return 2;
    hold_matrix(tcrossprod_csr_dense_float32(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                             NV((double *)x, (size_t)nnz), IM((int *)Y, nrowY, ncolY),
                                             nthreads));
    maybe_copy<int>(out);
    return 0;
}

/* ---- src/matmul.cpp:421-483 : CSR %*% dense vector ---- */
int mxref_matmul_csr_dvec_numeric(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const double *y, int leny, int nthreads, double *out)
{
    hold_vector(matmul_csr_dvec_numeric(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), NV((double *)y, (size_t)leny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_dvec_integer(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *y, int leny, int nthreads, double *out)
{
    hold_vector(matmul_csr_dvec_integer(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), IV((int *)y, (size_t)leny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_dvec_logical(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *y, int leny, int nthreads, double *out)
{
    hold_vector(matmul_csr_dvec_logical(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), LV((int *)y, (size_t)leny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_dvec_float32(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *y, int leny, int nthreads, int *out)
{
    hold_vector(matmul_csr_dvec_float32(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), IV((int *)y, (size_t)leny), nthreads));
    maybe_copy<int>(out);
    return 0;
}

} /* extern "C" */

/* ---- src/matmul.cpp:643-684 : float32 row vector %*% CSC (values == NULL: the pattern-matrix export) ---- */
extern "C" {

int mxref_matmul_rowvec_by_csc(const int *rowvec_f32_bits, int K, const int *p, int ncols, const int *i, const double *x,
                               int nnz, int *out)
{
    IV RV((int *)rowvec_f32_bits, (size_t)K), P((int *)p, (size_t)ncols + 1), I((int *)i, (size_t)nnz);
    if (x) hold_matrix(matmul_rowvec_by_csc(RV, P, I, NV((double *)x, (size_t)nnz)));
    else hold_matrix(matmul_rowvec_by_cscbin(RV, P, I));
    maybe_copy<int>(out);
    return 0;
}

} /* extern "C" */

/* ---- src/matmul.cpp:553-641 : CSR %*% sparse vector.  y indices are 1-based. ---- */
extern "C" {

int mxref_matmul_csr_svec_numeric(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *yi, const double *yv, int ny, int nthreads, double *out)
{
    hold_vector(matmul_csr_svec_numeric(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), IV((int *)yi, (size_t)ny),
                                        NV((double *)yv, (size_t)ny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_svec_integer(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *yi, const int *yv, int ny, int nthreads, double *out)
{
    hold_vector(matmul_csr_svec_integer(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), IV((int *)yi, (size_t)ny),
                                        IV((int *)yv, (size_t)ny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_svec_logical(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *yi, const int *yv, int ny, int nthreads, double *out)
{
    hold_vector(matmul_csr_svec_logical(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), IV((int *)yi, (size_t)ny),
                                        LV((int *)yv, (size_t)ny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_svec_binary(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                 const int *yi, const void *unused, int ny, int nthreads, double *out)
{
    (void)unused;
    hold_vector(matmul_csr_svec_binary(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                       NV((double *)x, (size_t)nnz), IV((int *)yi, (size_t)ny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

int mxref_matmul_csr_svec_float32(const int *p, int nrowsX, const int *j, const double *x, int nnz,
                                  const int *yi, const int *yv_float_bits, int ny, int nthreads, double *out)
{
    hold_vector(matmul_csr_svec_float32(IV((int *)p, (size_t)nrowsX + 1), IV((int *)j, (size_t)nnz),
                                        NV((double *)x, (size_t)nnz), IV((int *)yi, (size_t)ny),
                                        IV((int *)yv_float_bits, (size_t)ny), nthreads));
    maybe_copy<double>(out);
    return 0;
}

} /* extern "C" */
