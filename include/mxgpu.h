/* mxgpu.h — C ABI of the B200-native sparse x dense multiplication library (libmxgpu.so).
 *
 * This is the drop-in boundary for MatrixExtra's multiplication path: the ten Rcpp exports of
 * /root/reference/src/matmul.cpp:221-483 (registered at src/RcppExports.cpp:2233-2242 and called
 * from the S4 methods in R/matmul.R) are re-implemented on top of the entry points below by the
 * glue in rglue/matmul_gpu_glue.cpp.  Plain pointers and sizes only; every function returns an
 * int status (MXG_OK == 0) and never throws / longjmps; mxg_last_error() describes the failure.
 *
 * Conventions shared with the reference:
 *   - CSR A is m x K: p[m+1] int32 (0-based offsets), j[nnz] int32 0-based column ids,
 *     x[nnz] ALWAYS float64 on the host (R's dgRMatrix@x), also on the float32 path
 *     (src/matmul.cpp:122, 154: `const double *restrict values`).
 *   - float32 dense data is IEEE-754 binary32; R's `float` package keeps those bits in an INTEGER
 *     matrix (src/matmul.cpp:213-214), so the glue passes INTEGER(x) as the float pointer.
 *   - Index validity is NOT checked by the reference on this path (R/utils.R:349-410 checks only
 *     lengths); this library validates column ids once per upload and returns MXG_ERR_INDEX
 *     instead of faulting.
 *   - No alpha/beta: Out is overwritten completely (the reference writes into a zero-filled
 *     R matrix, src/matmul.cpp:197/261/323/389).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * MXG_ERR_CUDA.
 * Threads: every entry point that takes HOST buffers (level 1, mxg_csr_upload, the mxg_csr_*_host handle products,
 * mxg_set_devices, mxg_cache_clear) runs under one process-wide lock — they share per-device streams, the page-locked
 * staging arena and the operand cache, and the reference's caller is a single R thread.  mxg_last_error() and
 * mxg_last_call_bytes() are per calling thread.  Device-buffer (mxg_dev_*) products are asynchronous on the caller's
 * stream and may run concurrently on one handle from different streams.
 */
#ifndef MXGPU_H
#define MXGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define MXG_OK 0
#define MXG_ERR_CUDA 1        /* a CUDA runtime call failed (no device, OOM, launch failure) */
#define MXG_ERR_ARG 2         /* inconsistent arguments (negative sizes, NULL where data is needed, bad enum) */
#define MXG_ERR_INDEX 3       /* CSR column index outside [0, K) or non-monotone / negative indptr */
#define MXG_ERR_UNSUPPORTED 4 /* valid request this build does not cover */

/* ---- enums (plain ints in the signatures) ---- */
#define MXG_F64 0 /* double: R numeric matrix */
#define MXG_F32 1 /* float : R float32@Data   */

#define MXG_ROWS_CONTIGUOUS 0 /* element (r, c) at r*ld + c : "row-major" */
#define MXG_COLS_CONTIGUOUS 1 /* element (r, c) at r + c*ld : R's native column-major */

#define MXG_Y_NUMERIC 0 /* double y             : matmul_csr_dvec_numeric  src/matmul.cpp:421 */
#define MXG_Y_INTEGER 1 /* int y, NA_INTEGER    : matmul_csr_dvec_integer  src/matmul.cpp:437 */
#define MXG_Y_LOGICAL 2 /* int y, NA_LOGICAL    : matmul_csr_dvec_logical  src/matmul.cpp:453 */
#define MXG_Y_FLOAT32 3 /* float y, float result: matmul_csr_dvec_float32  src/matmul.cpp:469 */
#define MXG_Y_BINARY 4  /* sparse vectors only: pattern vector, every stored entry is 1: matmul_csr_svec_binary src/matmul.cpp:608 */

/* most result buffers one product can write (the local one + the peer-mapped ones of the other GPUs of a box) */
#define MXG_MAX_DST 8

/* which copies of the CSR values a device-resident handle keeps */
#define MXG_KEEP_F64 1
#define MXG_KEEP_F32 2

typedef struct mxg_csr_s *mxg_csr_t; /* opaque device-resident CSR (or, by relabelling, CSC) */

/* ================================ library state ================================================ */

/* Human-readable description of the last failure on the calling thread ("" if none). */
const char *mxg_last_error(void);

/* Number of visible CUDA devices (0 and MXG_ERR_CUDA when there is none). */
int mxg_device_count(int *count);

/* Device used by subsequent calls from this thread (default: the current CUDA device). */
int mxg_set_device(int device);

/* Number of GPUs a host-buffer (level-1) product is spread over (default 1).  With n > 1, mxg_spmm_csr_dense and
 * mxg_spmv_csr cut the CSR into n nnz-balanced row blocks (mxg_row_partition) and run one streamed pipeline per
 * device from n host threads of the calling process: every device uploads only its block over its own PCIe link
 * and downloads its rows straight into the caller's result (no all-gather: the result is wanted on the host); the
 * dense operand crosses PCIe once, a slice per device, and is completed over NVLink peer copies.  This is the
 * in-process counterpart of the reference's one OpenMP team over the rows (src/matmul.cpp:132-136,
 * R/matmul.R:175-180) — an R session is one process.  Devices used: the calling thread's current device and the
 * n - 1 that follow it.  Calls with fewer than option "multi_min_nnz" stored entries, or with pageable CSR arrays
 * (option "multi_pageable"), stay on one device.
 * Results are bit-identical to n = 1.  The Rcpp glue reads MATRIXEXTRA_GPUS once and calls this. */
int mxg_set_devices(int n);
int mxg_get_devices(int *n);

/* Tuning knobs, all optional ("auto" when never set).  Unknown names return MXG_ERR_ARG.
 *   kernels : "piece" (stored entries per long-row piece, 1024), "spmm_lpr" (lanes per row of B), "spmm_cpl" (vectors
 *             per lane: 0 = auto, two when a row of B exceeds 128 bytes), "spmm_rpw" (rows per warp), "spmm_panel_mb" /
 *             "spmm_panel_cols" (column panels of the dense operand, off), "spmv_lpr", "spmv_tex" (numeric SpMV gathers y
 *             through the texture path, 1), "svec_smem" (sparse-vector product keeps the presence bitmap in shared memory
 *             when it fits, 1), "radix_bits" (largest digit of the CSR->CSC radix passes, 4..10, default 8);
 *   level-1 calls : "pipeline" (1 = streamed row chunks, 0 = whole-matrix upload first), "pipe_chunk_nnz" (stored
 *             entries per chunk, 0 = auto), "pipe_slots" (ring slots, 4), "h2d_chunk_mb";
 *   host staging (csrc/hoststage.cu) : "host_threads" (threads that narrow / bounce host memory; the Rcpp exports'
 *             `nthreads`; 0 = all logical CPUs up to 16), "host_narrow" (float32 products narrow the float64 values on
 *             the host before the copy: 1 = yes, except in multi-device calls fed from page-locked arrays, 2 = always), "host_stage" (pageable caller memory goes through the page-locked ring, 1),
 *             "host_pack" (streamed calls of >= 2^20 entries whose host threads are not narrowing values send, while
 *             the upload stream lags behind them, column ids as 2 / 2.5 / 3 bytes per entry when the matrix has <= 2^16 /
 *             2^20 / 2^24 columns: packed by the host threads, rebuilt on the device, 1; 2 = always, 3 = two chunks
 *             out of three: test modes), "host_pack_lag" (a chunk is packed while the upload of the chunk this many
 *             places before it is still pending, 2),
 *             "host_arena_max_mb" (largest page-locked arena the library may hold, 4096; beyond it the driver's own
 *             copies are used), "host_result_pool_mb" (page-locked result memory handed out by mxg_host_alloc, 4096),
 *             "host_thp" (madvise(MADV_HUGEPAGE) on a large pageable result before its first touch: a
 *             freshly allocated R matrix is otherwise filled at page-fault speed, 1),
 *             "host_pin_register" (a NEW page-locked block — staging arena, result pool — is huge-page backed anonymous
 *             memory touched by the host threads and registered with the driver, 5 - 10 x faster to create than a
 *             cudaHostAlloc block of that size, 1; 0 = cudaHostAlloc),
 *             "host_colsplit" (products on a device-resident CSR with host operands, rows-contiguous both: the dense
 *             operand goes up and the result comes down as two column halves, so the second upload and both kernels
 *             hide behind the first download; 1 = only for a page-locked dense operand and where each output element is
 *             summed in the same order as in the full-width product (bit-identical results: rows of 256 bytes),
 *             2 = wherever a half row is >= 128 bytes, 0 = off);
 *   several devices : "multi_min_nnz" (level-1 calls below this many stored entries stay on one device, 4 Mi),
 *             "multi_pageable" (0 = calls whose CSR arrays are pageable stay on one device: they are bound by the host
 *             threads' bounce copies, which more devices do not speed up; 1 = spread them too),
 *             "multi_dense_share" (1 = each device uploads one slice of the dense operand and pulls the rest over NVLink,
 *             0 = every device uploads all of it);
 *   residency : "cache_mb" (level-1 operand cache, MiB of device memory, 0 = off: see mxg_cache_clear). */
int mxg_set_option(const char *name, long value);
int mxg_get_option(const char *name, long *value);

/* Number of kernels this library has launched since load (bench.py's `gpu_launches` claim). */
unsigned long long mxg_launch_count(void);

/* Release cached device memory / pinned staging held between calls. */
int mxg_trim(void);

/* ============ level 1: host buffers in, host buffers out (what the Rcpp glue calls) ============
 * Each call uploads its operands (pinned staging ring, chunked, overlapped with compute), runs the
 * device kernels and writes the result straight into the caller's (R-allocated) output buffer.   */

/* Out = A_csr(m x K) . B, B with n columns.  Replaces gemm_csr_drm_as_drm (src/matmul.cpp:118-142,
 * out_layout = MXG_ROWS_CONTIGUOUS, ldc >= n) and gemm_csr_drm_as_dcm (src/matmul.cpp:150-185,
 * out_layout = MXG_COLS_CONTIGUOUS, ldc >= m) together with their typed wrappers
 * matmul_dense_csc / tcrossprod_dense_csr / tcrossprod_csr_dense (src/matmul.cpp:188-375).
 *   dtype     MXG_F64 | MXG_F32: element type of B and Out (x is narrowed to float first for F32,
 *             src/matmul.cpp:53-57)
 *   b_layout  MXG_ROWS_CONTIGUOUS: B[k, c] at k*ldb + c (what every reference entry point receives:
 *             an (n x K) column-major R matrix, ldb = n);  MXG_COLS_CONTIGUOUS: B[k, c] at k + c*ldb
 *             (an untransposed K x n R matrix; saves the R-side t(y) of R/matmul.R:464, 507)        */
int mxg_spmm_csr_dense(int dtype, int out_layout, int b_layout,
                       int m, int K, int n,
                       const int32_t *p, const int32_t *j, const double *x,
                       const void *B, size_t ldb,
                       void *Out, size_t ldc);

/* out[m] = A_csr(m x K) . y.  Replaces matmul_csr_dvec<> (src/matmul.cpp:381-483).
 * ytype selects the element type of y and the NA rules of src/matmul.cpp:406-411; the result is
 * double except for MXG_Y_FLOAT32 (float). */
int mxg_spmv_csr(int ytype, int m, int K,
                 const int32_t *p, const int32_t *j, const double *x,
                 const void *y, void *out);

/* out[m] (always double) = A_csr(m x K) . y for a SPARSE vector y given as n_y (index, value) pairs with 1-based
 * indices, as R's sparseVector@i / @x.  Replaces matmul_csr_svec<> (src/matmul.cpp:486-551) and its exports
 * matmul_csr_svec_{numeric,integer,logical,binary,float32} (553-641), called from gemv_csr_vec (R/matmul.R:595-646).
 * ytype: MXG_Y_NUMERIC (double values), MXG_Y_INTEGER / MXG_Y_LOGICAL (int values, INT_MIN = NA -> NA_real_,
 * src/matmul.cpp:523-528), MXG_Y_FLOAT32 (float values), MXG_Y_BINARY (y_vals ignored, may be NULL).
 * K: number of columns of A, or <= 0 when unknown (the reference's export does not receive it): column ids are
 * then only required to be non-negative.  Indices of y outside [1, K] never match and are ignored, a repeated
 * index keeps its first entry (what the reference's merge does); y and the rows of A need not be sorted. */
int mxg_spmv_csr_svec(int ytype, int m, int K,
                      const int32_t *p, const int32_t *j, const double *x,
                      int n_y, const int32_t *y_idx_base1, const void *y_vals, double *out);

/* ---- the steps either side of a product ---- */

/* Validity of hand-built CSR arrays.  Replaces check_valid_csr_matrix (src/misc.cpp:970-1016, called from
 * R/utils.R:439-489): *code = 0 when valid, else the ordinal of the FIRST failing check in the reference's order;
 * mxg_csr_error_string(code) is the reference's message ("Matrix has negative indices." ...).  nnz = length of j. */
int mxg_check_valid_csr(int m, int ncols, const int32_t *p, const int32_t *j, int64_t nnz, int *code);
const char *mxg_csr_error_string(int code);

/* *sorted = 1 when the column ids of EVERY row are non-decreasing.  Replaces check_indices_are_unsorted
 * (src/misc.cpp:161-175; the reference's name is misleading, it returns true for sorted input). */
int mxg_rows_sorted(int m, const int32_t *p, const int32_t *j, int *sorted);

/* Sort the column ids of every row (and the values with them; x may be NULL) IN PLACE in the caller's arrays.
 * Replaces sort_sparse_indices<T> (src/misc.cpp:192-252, exports 300-330): rows already non-decreasing are left
 * alone; repeated ids keep their stored order (the reference's std::sort leaves that unspecified).  A sparse
 * vector is the one-row case (m = 1, p = {0, n}). */
int mxg_sort_csr_indices(int m, const int32_t *p, int32_t *j, double *x);

/* values_out[nnz] = x[e] * dense[row(e), j[e]] for a column-major m x K dense matrix of element type `dtype`
 * (MXG_Y_NUMERIC double, MXG_Y_FLOAT32 float, MXG_Y_INTEGER / MXG_Y_LOGICAL int with INT_MIN = NA -> NA_real_).
 * Replaces multiply_csr_by_dense_elemwise_{double,float32,int,bool} (src/operators.cpp:239-314): the result keeps
 * the sparsity pattern of the CSR operand, only the values change.  Bit-exact (one multiply per entry). */
int mxg_mul_csr_dense(int dtype, int m, int K, const int32_t *p, const int32_t *j, const double *x,
                      const void *dense, double *values_out);

/* values_out[nnz] = x[e] * dvec[pos(e)] with R's recycling of a dense vector along the column-major position:
 * pos = row + col * m, taken modulo len when len < m * K.  Replaces the Multiply case of
 * multiply_csr_by_dvec_no_NAs_numeric (src/operators.cpp:1478, 1501-2178; R/operators.R:236-397). */
int mxg_mul_csr_dvec(int m, int K, const int32_t *p, const int32_t *j, const double *x,
                     const double *dvec, size_t len, double *values_out);

/* Deep CSR(m x K) -> CSC conversion, bit-exact stable counting order (rows ascending inside each
 * column, duplicates in stored order).  Replaces the `as(x, "CsparseMatrix")` that
 * R/conversions.R:390-392 delegates to the Matrix package.  p2[K+1], i2[nnz], x2[nnz] are caller
 * buffers; x/x2 may both be NULL for pattern matrices. */
int mxg_csr2csc(int m, int K,
                const int32_t *p, const int32_t *j, const double *x,
                int32_t *p2, int32_t *i2, double *x2);

/* Out(K x n) = t(A_csr(m x K)) . B(m x n): device CSR->CSC transpose followed by the gather
 * product.  Serves crossprod(CSR, dense), t(CSR) %*% dense and (by transposition of the result
 * layout) dense %*% CSR, which MatrixExtra leaves to the Matrix package. */
int mxg_spmm_csrT_dense(int dtype, int out_layout, int b_layout,
                        int m, int K, int n,
                        const int32_t *p, const int32_t *j, const double *x,
                        const void *B, size_t ldb,
                        void *Out, size_t ldc);

/* ---- level-1 operand cache (option "cache_mb" > 0, off by default) -------------------------------
 * The callers of the reference multiply ONE sparse matrix many times (vignettes/Introducing_MatrixExtra.Rmd:454-476:
 * `X %*% coefs` inside optim), and every .Call hands over the same R vectors.  With the cache on, the device copy
 * of the CSR that a level-1 product has streamed in is kept, keyed on the three host addresses, the sizes and a
 * sampled fingerprint of the contents (both ends + ~1000 positions of each array), LRU within cache_mb; the next
 * product on the same arrays moves only the dense operand and the result.  Dense operands are kept the same way,
 * crossprod keeps the device-built CSC with its matrix.  Opt-in because R objects can be modified in place
 * (MatrixExtra.inplace_sort): a change that touches none of the sampled positions is not noticed — call
 * mxg_cache_clear() after modifying a matrix in place, or use explicit handles (below), which have no such caveat. */
int mxg_cache_clear(void);
int mxg_cache_stats(unsigned long long *hits, unsigned long long *misses, size_t *bytes, int *entries);

/* ============ level 2: device-resident handles (repeated multiplies, benchmarks, sharding) ===== */

/* Upload a host CSR once (validates indices, converts values, computes row statistics). */
int mxg_csr_upload(int m, int K, const int32_t *p, const int32_t *j, const double *x,
                   int keep, mxg_csr_t *handle);

/* Wrap CSR arrays that already live in device memory (not copied, not freed by mxg_csr_free).
 * d_x64 / d_x32 may be NULL individually.  Runs validation + row statistics on `stream`. */
int mxg_csr_wrap_device(int m, int K, const int32_t *d_p, const int32_t *d_j,
                        const double *d_x64, const float *d_x32, int validate,
                        void *stream, mxg_csr_t *handle);

int mxg_csr_free(mxg_csr_t handle);

/* Products of a device-resident handle with HOST operands: the same semantics as mxg_spmm_csr_dense /
 * mxg_spmm_csrT_dense / mxg_spmv_csr, but the CSR does not cross PCIe again — only the dense operand goes up and
 * the result comes down, in row chunks that leave while later rows are computed (mxg_last_call_bytes counts them).
 * What the glue's as_gpu_csr() objects call (rglue/handle_gpu_glue.cpp).  _t_: Out(K x n) = t(A) . B(m x n); the
 * CSC is built on the device at the first such product and kept with the handle.  Synchronous. */
int mxg_csr_spmm_host(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n,
                      const void *B, size_t ldb, void *Out, size_t ldc);
int mxg_csr_spmm_t_host(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n,
                        const void *B, size_t ldb, void *Out, size_t ldc);
int mxg_csr_spmv_host(mxg_csr_t A, int ytype, const void *y, void *out);

/* m, K, nnz, number of long rows, number of long-row pieces, longest row */
int mxg_csr_info(mxg_csr_t handle, int64_t info[6]);

/* Device pointers of the handle's arrays (for sharding / tests); any out pointer may be NULL. */
int mxg_csr_device_arrays(mxg_csr_t handle, const int32_t **d_p, const int32_t **d_j,
                          const double **d_x64, const float **d_x32);

/* Copy the handle's arrays back to host buffers (any of them may be NULL): p[m+1] rebased to start at 0,
 * j[nnz], x[nnz] (float64; widened from the float32 copy when only that is kept). */
int mxg_csr_download(mxg_csr_t handle, int32_t *p, int32_t *j, double *x);

/* Same products as level 1 with every operand already in device memory; asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = legacy default stream). */
int mxg_dev_spmm(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n,
                 const void *d_B, size_t ldb, void *d_Out, size_t ldc, void *stream);

int mxg_dev_spmv(mxg_csr_t A, int ytype, const void *d_y, void *d_out, void *stream);

/* Sparse-vector product on a device-resident handle; d_yidx_base1 / d_yvals / d_out are device pointers. */
int mxg_dev_spmv_svec(mxg_csr_t A, int ytype, int n_y, const int32_t *d_yidx_base1, const void *d_yvals,
                      double *d_out, void *stream);

/* Device-array forms of the f3 / f4 entry points (raw device pointers or a handle; they synchronise `stream` where
 * a result is returned to the host).  The device sort is OUT of place: d_j_out / d_x_out must not alias the inputs;
 * *rows_sorted (may be NULL) receives how many rows needed sorting. */
int mxg_dev_check_valid_csr(int m, int ncols, const int32_t *d_p, const int32_t *d_j, int64_t nnz, int *code, void *stream);
int mxg_dev_rows_sorted(int m, const int32_t *d_p, const int32_t *d_j, int *sorted, void *stream);
int mxg_dev_sort_csr_indices(int m, const int32_t *d_p, const int32_t *d_j, const double *d_x,
                             int32_t *d_j_out, double *d_x_out, int *rows_sorted, void *stream);
int mxg_dev_mul_csr_dense(mxg_csr_t A, int dtype, const void *d_dense, double *d_values_out, void *stream);
int mxg_dev_mul_csr_dvec(mxg_csr_t A, const double *d_dvec, size_t len, double *d_values_out, void *stream);

/* Multi-GPU form of the two products (north_star subsystem 4; no counterpart in the reference, which is one
 * process on shared memory: the OpenMP row loop of src/matmul.cpp:132-136 is the decomposition kept here).
 * The CSR is split into row blocks, one per GPU; every finished output row is stored into ALL n_dst result
 * buffers — d_outs[0] the local one, the others the peers' buffers mapped with mxg_ipc_open (or plain
 * pointers after cudaDeviceEnablePeerAccess in a single process) — so the all-gather of the row blocks
 * travels over NVLink store by store while the product is still running.  Every d_outs[g] already points
 * at this block's first row inside rank g's full result and shares the leading dimension ldc. */
int mxg_dev_spmm_bcast(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *d_B, size_t ldb,
                       int n_dst, void *const *d_outs, size_t ldc, void *stream);
int mxg_dev_spmv_bcast(mxg_csr_t A, int ytype, const void *d_y, int n_dst, void *const *d_outs, void *stream);

/* One slice of a product: rows [r0, r1) of the handle without their long rows (pieces == 0), or only the long rows
 * (pieces != 0).  d_Out is the origin of the FULL result (ldc as for mxg_dev_spmm).  A product can so be issued as
 * "long rows, then row slices" with something else — a collective, a copy — started behind every slice. */
int mxg_dev_spmm_rows(mxg_csr_t A, int dtype, int out_layout, int n, const void *d_B, size_t ldb, void *d_Out, size_t ldc,
                      int r0, int r1, int pieces, void *stream);

/* Strided device-to-device copy (`height` lines of `width_bytes`) on a copy engine, asynchronous on `stream`. */
int mxg_dev_copy_2d(void *d_dst, size_t dpitch, const void *d_src, size_t spitch, size_t width_bytes, size_t height,
                    void *stream);

/* The same product + all-gather with the COPY ENGINES: the product runs in about 16 row slices of equal nnz into
 * d_outs[0] and every finished slice is pushed to the other n_dst - 1 destinations by DMA (peer copies over NVLink;
 * 2-D copies for column-major results, whose blocks are n strided column segments) on the library's copy streams
 * while the next slice is computed; `stream` continues once every push has landed locally-ordered (close the step
 * with mxg_dev_peer_barrier).  Bulk transfers instead of 128-byte SM stores: the way to gather column-major results
 * (BASELINE cfg5) and the faster one for large row-major blocks. */
int mxg_dev_spmm_push(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *d_B, size_t ldb,
                      int n_dst, void *const *d_outs, size_t ldc, void *stream);

/* The same fused product + all-gather through NVLS MULTICAST: mc_out is this block's first row inside a multicast
 * mapping of the full result (cuMulticast* / torch.distributed._symmetric_memory: one virtual address bound to the
 * result buffers of every GPU of the box).  Every finished row is written once with multimem.st and the NVSwitch
 * replicates it into all G buffers, the local one included, so a row block leaves its GPU once instead of G-1 times.
 * Rows-contiguous results only; the caller closes the step with a barrier across the GPUs. */
int mxg_dev_spmm_mcast(mxg_csr_t A, int dtype, int n, const void *d_B, size_t ldb, void *mc_out, size_t ldc, void *stream);

/* Device memory that can be shared with the other processes of the box, and its handles (cudaIpc*). */
int mxg_dev_alloc(size_t bytes, void **d_ptr);
int mxg_dev_free(void *d_ptr);
int mxg_ipc_export(const void *d_ptr, unsigned char handle[64]);
int mxg_ipc_open(const unsigned char handle[64], void **d_ptr);
int mxg_ipc_close(void *d_ptr);

/* Completion barrier of a bcast step, stream-ordered and device-side: stores `epoch` into slot `rank` of every
 * peer's flag array (peer_flags[g] = rank g's array of `world` ints, peer-mapped; peer_flags[rank] = the local
 * array) and waits until every local slot has reached `epoch`.  After it, all ranks' rows have landed in the
 * local result.  Gives up after ~10 s and raises the sticky flag readable with mxg_dev_barrier_failed(). */
int mxg_dev_peer_barrier(int rank, int world, int *const *peer_flags, int epoch, void *stream);
int mxg_dev_barrier_failed(int *failed);

/* New device-resident CSC of A (as a CSR handle of t(A): K rows, m columns). */
int mxg_dev_csr2csc(mxg_csr_t A, int keep, void *stream, mxg_csr_t *At);

/* Dense layout change on device: dst(c, r) = src(r, c); rows x cols elements of 4 or 8 bytes. */
int mxg_dev_transpose_dense(int elem_size, size_t rows, size_t cols,
                            const void *d_src, size_t ld_src, void *d_dst, size_t ld_dst, void *stream);

/* nnz-balanced contiguous row blocks for multi-GPU sharding (binary search of g*nnz/G in p):
 * row_starts[parts+1] is filled with 0 = r_0 <= r_1 <= ... <= r_parts = m.  Host indptr. */
int mxg_row_partition(int m, const int32_t *p, int parts, int32_t *row_starts);

/* ============ host-side staging helpers (no GPU involved) ============
 * The streamed level-1 calls run these on the library's pool of host threads (option "host_threads", what the
 * reference's `nthreads` argument of src/matmul.cpp:221-483 now controls; 0 = all logical CPUs up to 16).
 * mxg_host_narrow: dst[i] = (float)src[i], round-to-nearest-even — the reference's per-entry cast
 * (src/matmul.cpp:53-57) done once, before the values cross PCIe.  mxg_host_copy_2d: `height` lines of `width`
 * bytes between pitched host buffers (how pageable R memory enters and leaves the page-locked staging arena);
 * streaming_stores != 0 writes the destination with cache-bypassing stores (used when it is a ring slot that the
 * DMA engine reads next). */
int mxg_host_narrow(const double *src, float *dst, size_t n);
/* Page-locked memory for RESULTS the glue allocates (R: Rf_allocVector3 with an R_allocator_t whose hooks are these two;
 * rglue/mxgpu_result_alloc.h).  A freshly malloc'ed result consists of pages that do not exist yet: filling it costs a
 * bounce through a page-locked slot plus the kernel's page zeroing (cfg3's 512 MB: 27 of the call's 54 ms).  A block from
 * this pool is DMA'd into directly.  Blocks are recycled (2 MiB granularity, <= 25 % waste) when freed — by R's garbage
 * collector through the allocator's free hook; the pool holds at most option "host_result_pool_mb" (4096), beyond which
 * mxg_host_alloc fails and the glue falls back to the ordinary allocator.  mxg_host_free returns MXG_ERR_ARG, and does
 * nothing, for a pointer that is not a live block of the pool.  mxg_trim releases the free blocks. */
int mxg_host_alloc(size_t bytes, void **ptr);
int mxg_host_free(void *ptr);
int mxg_host_pool_stats(size_t *live_bytes, size_t *free_bytes, int *blocks);
int mxg_host_copy_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                     int streaming_stores);
/* mxg_host_pack_indices: the wire format of column ids in a streamed call (option "host_pack"): n uint16 low
 * halves, padded to 16 bytes, then the high parts — nothing for K <= 2^16, a nibble per entry (entry i in byte i/2,
 * even entries in the low nibble) for K <= 2^20, a byte per entry for K <= 2^24.  *packed_bytes receives the size
 * (0 when K > 2^24: such ids travel as int32); packed == NULL only queries it.  *in_range = 0 when an id lies
 * outside [0, K) (the streamed call then fails with MXG_ERR_INDEX before anything is computed from that chunk).
 * `packed` must be 16-byte aligned. */
int mxg_host_pack_indices(const int32_t *j, size_t n, int K, void *packed, size_t *packed_bytes, int *in_range);
/* mxg_last_call_bytes: bytes the calling thread's most recent streamed product (mxg_spmm_csr_dense / mxg_spmv_csr
 * with option "pipeline" = 1) copied host -> device and device -> host, counted copy by copy (after host narrowing
 * and id packing; bench.py's e2e.h2d_bytes_per_step / d2h_bytes_per_step). */
int mxg_last_call_bytes(size_t *h2d_bytes, size_t *d2h_bytes);
/* mxg_host_chunk_plan: how a streamed level-1 call would cut the CSR with this indptr into row chunks (first row of
 * every chunk in chunk_rows[0 .. *n_chunks], cap = slots available there; chunk_rows may be NULL to query the
 * counts): about 16 chunks of equal nnz (at most 16 Mi entries / 64 MiB of result rows each, result_row_bytes per
 * row) that shrink towards the end of the matrix, or chunks of option "pipe_chunk_nnz" entries when that is set.
 * Also reports the long rows (> option "piece" entries), their pieces and the longest row.  MXG_ERR_INDEX for a
 * negative or decreasing indptr — the same check the products make. */
int mxg_host_chunk_plan(int m, const int32_t *p, size_t result_row_bytes, int32_t *chunk_rows, int cap, int *n_chunks,
                        int *n_long, int *n_pieces, int *max_len);

/* ============ synthetic inputs for bench.py / tests (device-side, counter-based RNG) ============
 * Power-law row lengths, stratified sorted unique columns; see DESIGN.md "Synthetic inputs". */
int mxg_synth_csr(int m, int K, int64_t target_nnz, int row_model, int col_model, uint64_t seed,
                  int keep, void *stream, mxg_csr_t *handle);

/* Measurement probe: fetches `gathers` (rounded up; the exact count comes back in *gathers_done) random rows of
 * row_bytes (128 / 256 / 512) from a device table of `rows` rows and does nothing else — the random-row-gather
 * roof (DRAM when the table is far larger than L2, L2->SM when it fits) that bounds K1/K2.  d_sink: 1 float. */
int mxg_dev_gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed,
                         float *d_sink, long long *gathers_done, void *stream);

/* The same random-row gathers issued through the TMA unit: one cp.async.bulk.tensor.2d ... tile::gather4 per four rows
 * (one thread, destination shared memory, completion on an mbarrier), 8 operations in flight per warp — the Blackwell
 * alternative to LDG.128 gathers for the dense-operand rows of K1/K2, measured side by side with mxg_dev_gather_probe.
 * d_table 128-byte aligned, rows < 2^31; d_sink: 2 floats (d_sink[1] is set to -1 when a gather never completed). */
int mxg_dev_tma_gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed,
                             float *d_sink, long long *gathers_done, void *stream);

/* Measurement probe for K3: the handle's column ids streamed 16 bytes per thread with the row structure taken away.
 * mode 0: ids only; 1: + an 8-byte gather of d_y[j] per entry (plain loads); 2: the same through the texture path;
 * 3: ids + float64 values + texture gathers + FMA (12 streamed bytes and one gather per entry, what the SpMV must
 * do at the very least).  d_sink: 1 double.  The time of mode 3 is the floor the SpMV kernel is measured against. */
int mxg_dev_spmv_probe(mxg_csr_t A, int mode, const double *d_y, double *d_sink, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MXGPU_H */
