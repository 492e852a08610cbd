/* tests/glue_driver.cpp — TEST INFRASTRUCTURE: runs rglue/matmul_gpu_glue.cpp (the Rcpp glue a maintainer adds
 * to MatrixExtra) without R.  The glue is compiled unmodified against the Rcpp stand-in oracle/shim/Rcpp.h
 * (-DMXGPU_GLUE_SHIM) and linked with libmxgpu.so; the extern "C" functions below hand it caller buffers the
 * way R hands it SEXPs (borrowed, column-major, float32 as int bits) and copy the returned object out.
 * Return value: 0, or 1 with the R error message in gluedrv_last_error(). */
#include "../rglue/matmul_gpu_glue.cpp"
#include "../rglue/rowops_gpu_glue.cpp"
#include "../rglue/handle_gpu_glue.cpp"

#include <cstring>
#include <string>

namespace {
std::string g_err;
typedef Rcpp::IntegerVector IV;
typedef Rcpp::NumericVector NV;
typedef Rcpp::LogicalVector LV;
typedef Rcpp::NumericMatrix NM;
typedef Rcpp::IntegerMatrix IM;

template <class Obj, class T>
void copy_out(const Obj &o, T *out)
{
    if (o.size() > 0) std::memcpy(out, o.data_ptr(), sizeof(T) * (size_t)o.size());
}

template <class Fn>
int guarded(Fn fn)
{
    try {
        fn();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
} // namespace

extern "C" {

const char *gluedrv_last_error(void) { return g_err.c_str(); }

/* dense(a x K) . t(sparse rows) : which = 0 matmul_dense_csc, 1 tcrossprod_dense_csr ; f32 = float bits in int */
int gluedrv_dense_sparse(int which, int f32, const void *X, int a, int K, const int *p, int rows, const int *idx,
                         const double *x, int nnz, void *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)rows + 1), J((int *)idx, (size_t)nnz);
        NV V((double *)x, (size_t)nnz);
        if (f32) {
            IM Xm((int *)X, a, K);
            IM r = which == 0 ? matmul_dense_csc_float32(Xm, P, J, V, 1) : tcrossprod_dense_csr_float32(Xm, P, J, V, 1, K);
            copy_out(r, (int *)out);
        } else {
            NM Xm((double *)X, a, K);
            NM r = which == 0 ? matmul_dense_csc_numeric(Xm, P, J, V, 1) : tcrossprod_dense_csr_numeric(Xm, P, J, V, 1, K);
            copy_out(r, (double *)out);
        }
    });
}

/* tcrossprod(CSR(m x K), dense(n x K)) */
int gluedrv_sparse_tdense(int f32, const int *p, int m, const int *j, const double *x, int nnz, const void *Y, int n,
                          int K, void *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J((int *)j, (size_t)nnz);
        NV V((double *)x, (size_t)nnz);
        if (f32) copy_out(tcrossprod_csr_dense_float32(P, J, V, IM((int *)Y, n, K), 1), (int *)out);
        else copy_out(tcrossprod_csr_dense_numeric(P, J, V, NM((double *)Y, n, K), 1), (double *)out);
    });
}

/* CSR %*% dense vector; ytype as in mxgpu.h (0 numeric, 1 integer, 2 logical, 3 float32) */
int gluedrv_csr_dvec(int ytype, const int *p, int m, const int *j, const double *x, int nnz, const void *y, int K,
                     void *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J((int *)j, (size_t)nnz);
        NV V((double *)x, (size_t)nnz);
        switch (ytype) {
        case 0: copy_out(matmul_csr_dvec_numeric(P, J, V, NV((double *)y, (size_t)K), 1), (double *)out); break;
        case 1: copy_out(matmul_csr_dvec_integer(P, J, V, IV((int *)y, (size_t)K), 1), (double *)out); break;
        case 2: copy_out(matmul_csr_dvec_logical(P, J, V, LV((int *)y, (size_t)K), 1), (double *)out); break;
        default: copy_out(matmul_csr_dvec_float32(P, J, V, IV((int *)y, (size_t)K), 1), (int *)out); break;
        }
    });
}

/* crossprod(CSR(m x K), dense(m x n)) -> K x n */
int gluedrv_crossprod(int f32, const int *p, int m, const int *j, const double *x, int nnz, int K, const void *Y,
                      int yrows, int n, void *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J((int *)j, (size_t)nnz);
        NV V((double *)x, (size_t)nnz);
        if (f32) copy_out(crossprod_csr_dense_float32(P, J, V, K, IM((int *)Y, yrows, n), 1), (int *)out);
        else copy_out(crossprod_csr_dense_numeric(P, J, V, K, NM((double *)Y, yrows, n), 1), (double *)out);
    });
}

int gluedrv_csr_to_csc(const int *p, int m, const int *j, const double *x, int nnz, int K, int *p2, int *i2, double *x2)
{
    return guarded([&] {
        Rcpp::List r = csr_to_csc_gpu(IV((int *)p, (size_t)m + 1), IV((int *)j, (size_t)nnz), NV((double *)x, (size_t)nnz), K);
        std::memcpy(p2, r.entries[0].data, sizeof(int) * r.entries[0].size);
        if (nnz > 0) {
            std::memcpy(i2, r.entries[1].data, sizeof(int) * r.entries[1].size);
            std::memcpy(x2, r.entries[2].data, sizeof(double) * r.entries[2].size);
        }
    });
}

/* ---- rglue/rowops_gpu_glue.cpp ---- */

/* CSR %*% sparseVector; ytype as in mxgpu.h (4 = binary: yv ignored) */
int gluedrv_csr_svec(int ytype, const int *p, int m, const int *j, const double *x, int nnz, const int *yi, int ny,
                     const void *yv, double *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J((int *)j, (size_t)nnz), YI((int *)yi, (size_t)ny);
        NV V((double *)x, (size_t)nnz);
        switch (ytype) {
        case 0: copy_out(matmul_csr_svec_numeric(P, J, V, YI, NV((double *)yv, (size_t)ny), 1), out); break;
        case 1: copy_out(matmul_csr_svec_integer(P, J, V, YI, IV((int *)yv, (size_t)ny), 1), out); break;
        case 2: copy_out(matmul_csr_svec_logical(P, J, V, YI, LV((int *)yv, (size_t)ny), 1), out); break;
        case 3: copy_out(matmul_csr_svec_float32(P, J, V, YI, IV((int *)yv, (size_t)ny), 1), out); break;
        default: copy_out(matmul_csr_svec_binary(P, J, V, YI, 1), out); break;
        }
    });
}

/* float32 row vector (K) %*% CSC with ncols columns; x == NULL: pattern matrix */
int gluedrv_rowvec_by_csc(const float *rowvec, int K, const int *p, int ncols, const int *i, const double *x, int nnz, float *out)
{
    return guarded([&] {
        IV RV((int *)rowvec, (size_t)K), P((int *)p, (size_t)ncols + 1), I((int *)i, (size_t)nnz);
        if (x) copy_out(matmul_rowvec_by_csc(RV, P, I, NV((double *)x, (size_t)nnz)), (int *)out);
        else copy_out(matmul_rowvec_by_cscbin(RV, P, I), (int *)out);
    });
}

int gluedrv_rows_sorted(const int *p, int m, const int *j, int nnz, int *sorted)
{
    return guarded([&] { *sorted = check_indices_are_unsorted(IV((int *)p, (size_t)m + 1), IV((int *)j, (size_t)nnz)) ? 1 : 0; });
}

/* in place on the caller's arrays (the vectors borrow them); x == NULL: pattern matrix */
int gluedrv_sort_indices(const int *p, int m, int *j, double *x, int nnz)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J(j, (size_t)nnz);
        if (x) sort_sparse_indices_numeric(P, J, NV(x, (size_t)nnz));
        else sort_sparse_indices_binary(P, J);
    });
}

/* msg receives the list's "err" string ("" when the matrix is valid) */
int gluedrv_check_valid(const int *p, int plen, const int *j, int nnz, int nrows, int ncols, char *msg, int cap)
{
    return guarded([&] {
        Rcpp::List r = check_valid_csr_matrix(IV((int *)p, (size_t)plen), IV((int *)j, (size_t)nnz), nrows, ncols);
        std::string e = r.entries.empty() ? std::string() : r.entries[0].str;
        std::strncpy(msg, e.c_str(), (size_t)cap - 1);
        msg[cap - 1] = 0;
    });
}

/* elementwise CSR * dense matrix (flat column-major vector of length m * K); dtype 0 double, 1 int, 2 bool, 3 float32 */
int gluedrv_mul_dense(int dtype, const int *p, int m, const int *j, const double *x, int nnz, const void *dense, long len,
                      double *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J((int *)j, (size_t)nnz);
        NV V((double *)x, (size_t)nnz);
        switch (dtype) {
        case 0: copy_out(multiply_csr_by_dense_elemwise_double(P, J, V, NV((double *)dense, (size_t)len)), out); break;
        case 1: copy_out(multiply_csr_by_dense_elemwise_int(P, J, V, IV((int *)dense, (size_t)len)), out); break;
        case 2: copy_out(multiply_csr_by_dense_elemwise_bool(P, J, V, LV((int *)dense, (size_t)len)), out); break;
        default: copy_out(multiply_csr_by_dense_elemwise_float32(P, J, V, IV((int *)dense, (size_t)len)), out); break;
        }
    });
}

int gluedrv_mul_dvec(const int *p, int m, const int *j, const double *x, int nnz, const double *dvec, long len, int ncols,
                     int multiply, double *out)
{
    return guarded([&] {
        IV P((int *)p, (size_t)m + 1), J((int *)j, (size_t)nnz);
        NV V((double *)x, (size_t)nnz);
        copy_out(multiply_csr_by_dvec_no_NAs_numeric(P, J, V, NV((double *)dvec, (size_t)len), ncols, multiply != 0, false,
                                                     false, false, false, true),
                 out);
    });
}

/* ---- rglue/handle_gpu_glue.cpp: device-resident matrices ---- */

/* returns an opaque box holding the external pointer (NULL + gluedrv_last_error() on failure) */
void *gluedrv_as_gpu_csr(const int *p, int m, const int *j, const double *x, int nnz, int K, int keep64, int keep32)
{
    void *box = nullptr;
    guarded([&] {
        box = new MxGpuCsrPtr(as_gpu_csr(IV((int *)p, (size_t)m + 1), IV((int *)j, (size_t)nnz), NV((double *)x, (size_t)nnz), K,
                                         keep64 != 0, keep32 != 0));
    });
    return box;
}

/* explicit free (R: gpu_csr_free(h)); the box itself goes with gluedrv_gpu_csr_drop (R: the object is collected) */
int gluedrv_gpu_csr_free(void *box)
{
    return guarded([&] { gpu_csr_free(*static_cast<MxGpuCsrPtr *>(box)); });
}

void gluedrv_gpu_csr_drop(void *box) { delete static_cast<MxGpuCsrPtr *>(box); }

/* op 0: A %*% t(Y) (Y n x K)   1: X %*% t(A) (X a x K)   2: t(A) %*% Y (Y m x n); rows / cols = dims of the dense operand */
int gluedrv_gpu_csr_product(void *box, int op, int f32, const void *D, int rows, int cols, void *out)
{
    return guarded([&] {
        MxGpuCsrPtr &h = *static_cast<MxGpuCsrPtr *>(box);
        if (f32) {
            IM Dm((int *)D, rows, cols);
            IM r = op == 0 ? gpu_csr_tcrossprod_dense_float32(h, Dm, 1)
                           : (op == 1 ? gpu_csr_dense_tcrossprod_float32(Dm, h, 1) : gpu_csr_crossprod_dense_float32(h, Dm, 1));
            copy_out(r, (int *)out);
        } else {
            NM Dm((double *)D, rows, cols);
            NM r = op == 0 ? gpu_csr_tcrossprod_dense_numeric(h, Dm, 1)
                           : (op == 1 ? gpu_csr_dense_tcrossprod_numeric(Dm, h, 1) : gpu_csr_crossprod_dense_numeric(h, Dm, 1));
            copy_out(r, (double *)out);
        }
    });
}

int gluedrv_gpu_csr_dvec(void *box, const double *y, int K, double *out)
{
    return guarded([&] { copy_out(gpu_csr_dvec_numeric(*static_cast<MxGpuCsrPtr *>(box), NV((double *)y, (size_t)K), 1), out); });
}

int gluedrv_configure(int gpus, int cache_mb, int *devices)
{
    return guarded([&] { *devices = mxgpu_configure(gpus, cache_mb); });
}

} /* extern "C" */
