/* rglue/handle_gpu_glue.cpp — device-resident sparse matrices for MatrixExtra's multiplication path.
 *
 * The callers of the reference multiply ONE sparse matrix many times (/root/reference/vignettes/
 * Introducing_MatrixExtra.Rmd:454-476: `X %*% coefs` and `crossprod`-type gradients inside optim; rsparse-style
 * ALS alternates tcrossprod against a fixed ratings matrix).  Through the ten level-1 exports every call uploads the
 * CSR again — for BASELINE cfg3 that upload (1.06 GB) is the whole call.  These exports keep the CSR in HBM:
 *
 *     h <- as_gpu_csr(X@p, X@j, X@x, ncol(X), TRUE, FALSE)     # once: upload + validation + row statistics
 *     gpu_csr_tcrossprod_dense_numeric(h, t(B), nthreads)       # X %*% B      : moves B up, the result down
 *     gpu_csr_dense_tcrossprod_numeric(D, h, nthreads)          # D %*% t(X)
 *     gpu_csr_crossprod_dense_numeric(h, Y, nthreads)           # t(X) %*% Y   : CSC built on the device once
 *     gpu_csr_dvec_numeric(h, v, nthreads)                      # X %*% v
 *     gpu_csr_free(h)                                           # or let R's GC run the finalizer
 *
 * The R side (rglue/matmul_gpu_methods.R) wraps the pointer in an S4 class `gpuRsparse` whose `%*%` / crossprod /
 * tcrossprod methods call these; argument meaning and result classes are those of the level-1 exports they mirror
 * (src/matmul.cpp:283-375, 421-483).  The handle is an external pointer with a C finalizer (mxg_csr_free), run on
 * garbage collection and at exit.  float32 operands need keep_float32 = TRUE at construction (values narrowed once,
 * bit-identical to the reference's per-entry cast, src/matmul.cpp:53-57).
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

struct MxGpuCsr {
    mxg_csr_t handle;
    int nrows, ncols;
    bool has_f64, has_f32;
};

inline void mxgpu_csr_finalizer(MxGpuCsr *g)
{
    if (!g) return;
    if (g->handle) mxg_csr_free(g->handle);
    delete g;
}

typedef Rcpp::XPtr<MxGpuCsr, Rcpp::PreserveStorage, mxgpu_csr_finalizer, true> MxGpuCsrPtr;

namespace {

inline void mxgpu_handle_check(int status)
{
    if (status != MXG_OK) MXGPU_GLUE_STOP(mxg_last_error());
}

inline MxGpuCsr *mxgpu_live(const MxGpuCsrPtr &ptr)
{
    MxGpuCsr *g = ptr.get();
    if (!g || !g->handle) MXGPU_GLUE_STOP("gpu matrix has been freed.");
    return g;
}

template <class RcppMatrix> struct HandleElem;
template <> struct HandleElem<Rcpp::NumericMatrix> {
    static const int dtype = MXG_F64;
    static void *ptr(const Rcpp::NumericMatrix &m) { return (void *)REAL(m); }
    static bool ok(const MxGpuCsr *g) { return g->has_f64; }
};
template <> struct HandleElem<Rcpp::IntegerMatrix> { /* float32@Data */
    static const int dtype = MXG_F32;
    static void *ptr(const Rcpp::IntegerMatrix &m) { return (void *)INTEGER(m); }
    static bool ok(const MxGpuCsr *g) { return g->has_f32; }
};

/* A(m x K) %*% t(Y), Y (n x K) column-major: the handle form of tcrossprod_csr_dense (src/matmul.cpp:316-343) */
template <class RcppMatrix>
RcppMatrix handle_times_tdense(const MxGpuCsrPtr &ptr, const RcppMatrix &Y_colmajor)
{
    MxGpuCsr *g = mxgpu_live(ptr);
    if (!HandleElem<RcppMatrix>::ok(g)) MXGPU_GLUE_STOP("gpu matrix was created without values of this type.");
    if (Y_colmajor.ncol() != g->ncols) MXGPU_GLUE_STOP("Matrix dimensions do not match.");
    const int n = Y_colmajor.nrow();
    RcppMatrix out = MXGPU_NEW_MATRIX(RcppMatrix, g->nrows, n);
    mxgpu_handle_check(mxg_csr_spmm_host(g->handle, HandleElem<RcppMatrix>::dtype, MXG_COLS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, n,
                                         HandleElem<RcppMatrix>::ptr(Y_colmajor), (size_t)(n > 0 ? n : 1),
                                         HandleElem<RcppMatrix>::ptr(out), (size_t)(g->nrows > 0 ? g->nrows : 1)));
    return out;
}

/* X(a x K) %*% t(A): the handle form of tcrossprod_dense_csr (src/matmul.cpp:254-281) */
template <class RcppMatrix>
RcppMatrix dense_times_thandle(const RcppMatrix &X_colmajor, const MxGpuCsrPtr &ptr)
{
    MxGpuCsr *g = mxgpu_live(ptr);
    if (!HandleElem<RcppMatrix>::ok(g)) MXGPU_GLUE_STOP("gpu matrix was created without values of this type.");
    if (X_colmajor.ncol() != g->ncols) MXGPU_GLUE_STOP("Matrix dimensions do not match.");
    const int a = X_colmajor.nrow();
    RcppMatrix out = MXGPU_NEW_MATRIX(RcppMatrix, a, g->nrows);
    const size_t ld = (size_t)(a > 0 ? a : 1);
    mxgpu_handle_check(mxg_csr_spmm_host(g->handle, HandleElem<RcppMatrix>::dtype, MXG_ROWS_CONTIGUOUS, MXG_ROWS_CONTIGUOUS, a,
                                         HandleElem<RcppMatrix>::ptr(X_colmajor), ld, HandleElem<RcppMatrix>::ptr(out), ld));
    return out;
}

/* t(A) %*% Y, Y (m x n) column-major, result (K x n) column-major: the handle form of crossprod_csr_dense */
template <class RcppMatrix>
RcppMatrix thandle_times_dense(const MxGpuCsrPtr &ptr, const RcppMatrix &Y_colmajor)
{
    MxGpuCsr *g = mxgpu_live(ptr);
    if (!HandleElem<RcppMatrix>::ok(g)) MXGPU_GLUE_STOP("gpu matrix was created without values of this type.");
    if (Y_colmajor.nrow() != g->nrows) MXGPU_GLUE_STOP("Matrix dimensions do not match.");
    const int n = Y_colmajor.ncol();
    RcppMatrix out = MXGPU_NEW_MATRIX(RcppMatrix, g->ncols, n);
    mxgpu_handle_check(mxg_csr_spmm_t_host(g->handle, HandleElem<RcppMatrix>::dtype, MXG_COLS_CONTIGUOUS, MXG_COLS_CONTIGUOUS, n,
                                           HandleElem<RcppMatrix>::ptr(Y_colmajor), (size_t)(g->nrows > 0 ? g->nrows : 1),
                                           HandleElem<RcppMatrix>::ptr(out), (size_t)(g->ncols > 0 ? g->ncols : 1)));
    return out;
}

} /* namespace */

/* Upload once: validates column ids against ncols (R error instead of the reference's unchecked read), narrows the
 * values when float32 products are wanted, builds the long-row tables. */
// [[Rcpp::export(rng = false)]]
MxGpuCsrPtr as_gpu_csr(Rcpp::IntegerVector indptr, Rcpp::IntegerVector indices, Rcpp::NumericVector values, int ncols,
                       bool keep_float64, bool keep_float32)
{
    const int m = (int)indptr.size() - 1;
    if (m < 0 || ncols < 0) MXGPU_GLUE_STOP("Matrix has invalid dimensions.");
    if (!keep_float64 && !keep_float32) keep_float64 = true;
    mxg_csr_t h = 0;
    mxgpu_handle_check(mxg_csr_upload(m, ncols, INTEGER(indptr), INTEGER(indices), REAL(values),
                                      (keep_float64 ? MXG_KEEP_F64 : 0) | (keep_float32 ? MXG_KEEP_F32 : 0), &h));
    MxGpuCsr *g = new MxGpuCsr();
    g->handle = h;
    g->nrows = m;
    g->ncols = ncols;
    g->has_f64 = keep_float64;
    g->has_f32 = keep_float32;
    return MxGpuCsrPtr(g, true);
}

// [[Rcpp::export(rng = false)]]
void gpu_csr_free(MxGpuCsrPtr ptr)
{
    ptr.release(); /* runs the finalizer now; later products on this object raise "has been freed" */
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerVector gpu_csr_dim(MxGpuCsrPtr ptr)
{
    MxGpuCsr *g = mxgpu_live(ptr);
    Rcpp::IntegerVector out = MXGPU_NEW_VECTOR(Rcpp::IntegerVector, 2);
    out[0] = g->nrows;
    out[1] = g->ncols;
    return out;
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix gpu_csr_tcrossprod_dense_numeric(MxGpuCsrPtr X, Rcpp::NumericMatrix Y_colmajor, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    return handle_times_tdense<Rcpp::NumericMatrix>(X, Y_colmajor);
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix gpu_csr_tcrossprod_dense_float32(MxGpuCsrPtr X, Rcpp::IntegerMatrix Y_colmajor, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    return handle_times_tdense<Rcpp::IntegerMatrix>(X, Y_colmajor);
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix gpu_csr_dense_tcrossprod_numeric(Rcpp::NumericMatrix X_colmajor, MxGpuCsrPtr Y, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    return dense_times_thandle<Rcpp::NumericMatrix>(X_colmajor, Y);
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix gpu_csr_dense_tcrossprod_float32(Rcpp::IntegerMatrix X_colmajor, MxGpuCsrPtr Y, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    return dense_times_thandle<Rcpp::IntegerMatrix>(X_colmajor, Y);
}

// [[Rcpp::export(rng = false)]]
Rcpp::NumericMatrix gpu_csr_crossprod_dense_numeric(MxGpuCsrPtr X, Rcpp::NumericMatrix Y_colmajor, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    return thandle_times_dense<Rcpp::NumericMatrix>(X, Y_colmajor);
}

// [[Rcpp::export(rng = false)]]
Rcpp::IntegerMatrix gpu_csr_crossprod_dense_float32(MxGpuCsrPtr X, Rcpp::IntegerMatrix Y_colmajor, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    return thandle_times_dense<Rcpp::IntegerMatrix>(X, Y_colmajor);
}

/* X %*% dense vector: the handle form of matmul_csr_dvec_numeric (src/matmul.cpp:421-435) */
// [[Rcpp::export(rng = false)]]
Rcpp::NumericVector gpu_csr_dvec_numeric(MxGpuCsrPtr X, Rcpp::NumericVector y_dense, int nthreads)
{
    mxg_set_option("host_threads", nthreads > 0 ? nthreads : 0);
    MxGpuCsr *g = mxgpu_live(X);
    if (!g->has_f64) MXGPU_GLUE_STOP("gpu matrix was created without values of this type.");
    if ((int)y_dense.size() != g->ncols) MXGPU_GLUE_STOP("Matrix dimensions do not match.");
    Rcpp::NumericVector out = MXGPU_NEW_VECTOR(Rcpp::NumericVector, g->nrows);
    mxgpu_handle_check(mxg_csr_spmv_host(g->handle, MXG_Y_NUMERIC, REAL(y_dense), REAL(out)));
    return out;
}

/* The level-1 exports' own residency and device settings, for R's .onLoad / options():
 *   mxgpu_configure(gpus, cache_mb): gpus > 0 -> mxg_set_devices(gpus) (MATRIXEXTRA_GPUS), cache_mb >= 0 -> option
 *   "cache_mb" (MATRIXEXTRA_GPU_CACHE_MB).  Returns the number of devices now in use. */
// [[Rcpp::export(rng = false)]]
int mxgpu_configure(int gpus, int cache_mb)
{
    if (gpus > 0) mxgpu_handle_check(mxg_set_devices(gpus));
    if (cache_mb >= 0) mxgpu_handle_check(mxg_set_option("cache_mb", cache_mb));
    int n = 1;
    mxgpu_handle_check(mxg_get_devices(&n));
    return n;
}
