// capi.cu — the extern "C" surface declared in include/mxgpu.h: library state, device-resident
// handles (level 2) and the host-buffer entry points the Rcpp glue binds (level 1).
#include "mxg_internal.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <thread>
#include <vector>

namespace mxg {

int gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed, float *d_sink,
                 cudaStream_t stream);
int synth_csr_arrays(int m, int K, int64_t target_nnz, int row_model, int col_model, uint64_t seed, int keep,
                     cudaStream_t stream, int32_t **out_p, int32_t **out_j, double **out_x64, float **out_x32,
                     int64_t *out_nnz);

// ---- per-process state: internal streams for the level-1 calls, pool configured once per device ----
static DeviceState g_dev[64];

int current_state(DeviceState **out)
{
    int dev = 0;
    MXG_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(MXG_ERR_CUDA, "device ordinal %d out of range", dev);
    DeviceState &st = g_dev[dev];
    if (!st.ready) {
        MXG_CUDA_TRY(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        MXG_CUDA_TRY(cudaStreamCreateWithFlags(&st.h2d, cudaStreamNonBlocking));
        MXG_CUDA_TRY(cudaStreamCreateWithFlags(&st.d2h, cudaStreamNonBlocking));
        MXG_CUDA_TRY(cudaStreamCreateWithFlags(&st.p2p, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        MXG_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep_all = ~0ULL; // keep freed blocks cached between calls (mxg_trim releases them)
        MXG_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
        st.ready = true;
    }
    *out = &st;
    return MXG_OK;
}

// Level-2 entry points allocate from the device's default pool too: make sure it keeps freed blocks (otherwise
// every stream synchronisation hands the memory back to the driver and the next multi-GB cudaMallocAsync
// pays tens of milliseconds for fresh physical pages).
static int ensure_device_ready()
{
    DeviceState *st;
    return current_state(&st);
}

static int free_handle(mxg_csr_s *h)
{
    if (!h) return MXG_OK;
    // stream-ordered frees on the handle's stream: cheap (pool) and ordered after the handle's own work;
    // the caller guarantees nothing on OTHER streams still uses the handle
    cudaStream_t s = h->stream;
    auto rel = [s](const void *q) { if (q) cudaFreeAsync(const_cast<void *>(q), s); };
    if (h->owns) {
        rel(h->d_p);
        rel(h->d_j);
        rel(h->d_x64);
        rel(h->d_x32);
    }
    rel(h->d_long_rows);
    rel(h->d_long_first);
    rel(h->d_long_np);
    rel(h->d_piece_row);
    rel(h->d_piece_k);
    rel(h->d_partial);
    rel(h->d_seg);
    delete h->host_chunks;
    if (h->cached_t) free_handle(h->cached_t);
    delete h;
    return MXG_OK;
}

// Host CSR -> owned device handle.  All copies are issued on `stream`; returns after they completed.
static int upload_csr(int m, int K, const int32_t *p, const int32_t *j, const double *x, int keep,
                      cudaStream_t stream, mxg_csr_s **out)
{
    if (m < 0 || K < 0) return fail(MXG_ERR_ARG, "csr: negative dimension");
    if (!p) return fail(MXG_ERR_ARG, "csr: indptr is NULL");
    const int32_t base = p[0];
    const int64_t nnz = (int64_t)p[m] - (int64_t)base;
    if (base < 0 || nnz < 0) return fail(MXG_ERR_INDEX, "csr: indptr is negative or decreasing");
    if (nnz > 0 && !j) return fail(MXG_ERR_ARG, "csr: indices is NULL");
    const bool want64 = (keep & MXG_KEEP_F64) != 0, want32 = (keep & MXG_KEEP_F32) != 0;
    if (nnz > 0 && (want64 || want32) && !x) return fail(MXG_ERR_ARG, "csr: values is NULL");

    mxg_csr_s *h = new mxg_csr_s();
    h->m = m;
    h->K = K;
    h->nnz = nnz;
    h->base = 0;
    h->owns = true;
    cudaGetDevice(&h->device);
    int32_t *d_p = nullptr, *d_j = nullptr;
    double *d_x64 = nullptr;
    float *d_x32 = nullptr;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
#define MXG_UP_TRY(expr)                                                                             \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            free_handle(h);                                                                          \
            return fail(MXG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));               \
        }                                                                                            \
    } while (0)
    h->stream = stream;
    MXG_UP_TRY(cudaMallocAsync(&d_p, sizeof(int32_t) * ((size_t)m + 1), stream));
    h->d_p = d_p;
    MXG_UP_TRY(cudaMallocAsync(&d_j, sizeof(int32_t) * nz, stream));
    h->d_j = d_j;
    if (base == 0) {
        MXG_UP_TRY(cudaMemcpyAsync(d_p, p, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice, stream));
    } else {
        // R never produces p[0] != 0 (R/utils.R:349-410 rejects it); rebase so device offsets start at 0
        std::vector<int32_t> pr((size_t)m + 1);
        for (int r = 0; r <= m; r++) pr[(size_t)r] = p[r] - base;
        MXG_UP_TRY(cudaMemcpyAsync(d_p, pr.data(), sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice, stream));
        MXG_UP_TRY(cudaStreamSynchronize(stream));
    }
    if (nnz > 0) {
        // pageable (R) arrays are bounced through the page-locked arena by the host threads (hoststage.cu)
        DeviceState *st = nullptr;
        {
            int rc = current_state(&st);
            if (rc == MXG_OK) rc = staged_h2d(st, d_j, j + base, sizeof(int32_t) * (size_t)nnz, stream);
            if (rc != MXG_OK) { free_handle(h); return rc; }
        }
        if (want64) {
            MXG_UP_TRY(cudaMallocAsync(&d_x64, sizeof(double) * nz, stream));
            h->d_x64 = d_x64;
            int rc = staged_h2d(st, d_x64, x + base, sizeof(double) * (size_t)nnz, stream);
            if (rc != MXG_OK) { free_handle(h); return rc; }
        }
        if (want32) {
            MXG_UP_TRY(cudaMallocAsync(&d_x32, sizeof(float) * nz, stream));
            h->d_x32 = d_x32;
            if (d_x64) {
                int rc = convert_f64_to_f32(d_x64, d_x32, (size_t)nnz, stream);
                if (rc != MXG_OK) { free_handle(h); return rc; }
            } else if (options().host_narrow != 0 && staged_h2d_narrow(st, d_x32, x + base, (size_t)nnz, stream) == MXG_OK) {
                // float32 only: narrowed by the host threads, 4 instead of 8 bytes per value over PCIe
            } else {
                cudaGetLastError(); // (no page-locked arena: narrow on the device instead)
                // stage the float64 values through pool memory in chunks and narrow them on device (K6)
                const size_t chunk = (size_t)std::max<long>(1, options().h2d_chunk_mb) * (1u << 20) / sizeof(double);
                double *d_tmp = nullptr;
                const size_t c0 = std::min(chunk, (size_t)nnz);
                const size_t c = c0 + (c0 & 1);
                MXG_UP_TRY(cudaMallocAsync(&d_tmp, sizeof(double) * c, stream));
                for (size_t off = 0; off < (size_t)nnz; off += c) {
                    const size_t len = std::min(c, (size_t)nnz - off);
                    MXG_UP_TRY(cudaMemcpyAsync(d_tmp, x + base + off, sizeof(double) * len, cudaMemcpyHostToDevice, stream));
                    int rc = convert_f64_to_f32(d_tmp, d_x32 + off, len, stream);
                    if (rc != MXG_OK) { free_handle(h); return rc; }
                }
                MXG_UP_TRY(cudaFreeAsync(d_tmp, stream));
            }
        }
    }
#undef MXG_UP_TRY
    int rc = csr_build_stats(h, /*validate=*/1, stream); // synchronises the stream
    if (rc != MXG_OK) {
        free_handle(h);
        return rc;
    }
    *out = h;
    return MXG_OK;
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Dense operand -> tight rows-contiguous device copy [K][ld] with ld a multiple of the 16-byte vector.
static int upload_dense_rows(int dtype, int b_layout, size_t K, size_t n, const void *B, size_t ldb,
                             cudaStream_t stream, void **d_out, size_t *ld_out)
{
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    const size_t vec = 16 / s;
    const size_t ld = round_up(n, vec);
    if (K > 0 && n > 0) { // argument errors before anything is allocated
        if (b_layout != MXG_ROWS_CONTIGUOUS && b_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "dense operand: bad layout %d", b_layout);
        if (b_layout == MXG_ROWS_CONTIGUOUS && ldb < n) return fail(MXG_ERR_ARG, "dense operand: ldb < n");
        if (b_layout == MXG_COLS_CONTIGUOUS && ldb < K) return fail(MXG_ERR_ARG, "dense operand: ldb < K");
    }
    StreamBuf buf, tmp; // handed back on every failure below
    MXG_CUDA_TRY(buf.alloc(K * ld * s, stream));
    void *d_B = buf.ptr;
    if (K > 0 && n > 0) {
        if (b_layout == MXG_ROWS_CONTIGUOUS) {
            if (ld != n) MXG_CUDA_TRY(cudaMemsetAsync(d_B, 0, K * ld * s, stream));
            if (ld == n && ldb == n) {
                DeviceState *st = nullptr;
                MXG_TRY(current_state(&st));
                MXG_TRY(staged_h2d(st, d_B, B, K * n * s, stream));
            } else {
                MXG_CUDA_TRY(cudaMemcpy2DAsync(d_B, ld * s, B, ldb * s, n * s, K, cudaMemcpyHostToDevice, stream));
            }
        } else {
            MXG_CUDA_TRY(tmp.alloc(K * n * s, stream));
            void *d_tmp = tmp.ptr;
            if (ldb == K) {
                DeviceState *st = nullptr;
                MXG_TRY(current_state(&st));
                MXG_TRY(staged_h2d(st, d_tmp, B, K * n * s, stream));
            } else {
                MXG_CUDA_TRY(cudaMemcpy2DAsync(d_tmp, K * s, B, ldb * s, K * s, n, cudaMemcpyHostToDevice, stream));
            }
            if (ld != n) MXG_CUDA_TRY(cudaMemsetAsync(d_B, 0, K * ld * s, stream));
            // d_tmp is n rows of K contiguous -> d_B is K rows of n contiguous
            MXG_TRY(launch_transpose_dense((int)s, n, K, d_tmp, K, d_B, ld, stream));
        }
    }
    buf.release();
    *d_out = d_B;
    *ld_out = ld;
    return MXG_OK;
}

// shared tail of the two level-1 SpMM entry points: A is on the device, B/Out are host buffers
// d_B_early / b_ready: the dense operand was put on its way by the caller (another stream); b_ready marks its arrival.
// The operand's device copy is released here on every path, early or not.
static int spmm_host_io(mxg_csr_s *A, int dtype, int out_layout, int b_layout, int n, const void *B, size_t ldb,
                        void *Out, size_t ldc, cudaStream_t stream, void *d_B_early = nullptr, size_t ld_b_early = 0,
                        cudaEvent_t b_ready = nullptr)
{
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    const size_t vec = 16 / s;
    const size_t rows = (size_t)A->m, K = (size_t)A->K;
    void *d_B = d_B_early, *d_Out = nullptr;
    size_t ld_b = ld_b_early;
    auto body = [&]() -> int {
        if (rows == 0 || n == 0) return MXG_OK;
        if (!B && K > 0) return fail(MXG_ERR_ARG, "dense operand is NULL");
        if (!Out) return fail(MXG_ERR_ARG, "output is NULL");
        if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "bad out_layout %d", out_layout);
        const bool rm = out_layout == MXG_ROWS_CONTIGUOUS;
        if (rm && ldc < (size_t)n) return fail(MXG_ERR_ARG, "output: ldc < n");
        if (!rm && ldc < rows) return fail(MXG_ERR_ARG, "output: ldc < m");
        if (d_B) MXG_CUDA_TRY(cudaStreamWaitEvent(stream, b_ready, 0));
        else MXG_TRY(upload_dense_rows(dtype, b_layout, K, (size_t)n, B, ldb, stream, &d_B, &ld_b));
        const size_t ld_o = rm ? round_up((size_t)n, vec) : rows;
        MXG_CUDA_TRY(cudaMallocAsync(&d_Out, rm ? rows * ld_o * s : rows * (size_t)n * s, stream));
        MXG_TRY(launch_spmm(A, dtype, out_layout, n, d_B, ld_b, d_Out, ld_o, stream));
        DeviceState *st = nullptr;
        MXG_TRY(current_state(&st));
        const size_t width = rm ? (size_t)n * s : rows * s, height = rm ? rows : (size_t)n;
        if (ldc * s == width && ld_o * s == width) // tight on both sides: one staged block copy
            MXG_TRY(staged_d2h(st, Out, d_Out, width * height, stream));
        else
            MXG_CUDA_TRY(cudaMemcpy2DAsync(Out, ldc * s, d_Out, ld_o * s, width, height, cudaMemcpyDeviceToHost, stream));
        return MXG_OK;
    };
    int rc = body();
    if (d_B) cudaFreeAsync(d_B, stream);
    if (d_Out) cudaFreeAsync(d_Out, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess && rc == MXG_OK) rc = fail(MXG_ERR_CUDA, "spmm: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int csr_handle_free(mxg_csr_s *h) { return free_handle(h); }

// every host-buffer (level-1) entry point runs under this lock: they share the per-device streams, the page-locked
// arena and the operand cache, and the reference's callers are a single R thread anyway (SURVEY.md 8 b)
static std::recursive_mutex g_level1_mu;

static int transpose_handle(const mxg_csr_s *A, int keep, cudaStream_t stream, mxg_csr_s **out)
{
    const bool w64 = (keep & MXG_KEEP_F64) && A->d_x64, w32 = (keep & MXG_KEEP_F32) && A->d_x32;
    mxg_csr_s *t = new mxg_csr_s();
    t->m = A->K;
    t->K = A->m;
    t->nnz = A->nnz;
    t->owns = true;
    t->device = A->device;
    int32_t *d_p2 = nullptr, *d_i2 = nullptr;
    double *d_x64 = nullptr;
    float *d_x32 = nullptr;
    const size_t nz = (size_t)(A->nnz > 0 ? A->nnz : 1);
    t->stream = stream;
    cudaError_t e = cudaMallocAsync(&d_p2, sizeof(int32_t) * ((size_t)A->K + 1), stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_i2, sizeof(int32_t) * nz, stream);
    if (e == cudaSuccess && w64) e = cudaMallocAsync(&d_x64, sizeof(double) * nz, stream);
    if (e == cudaSuccess && w32) e = cudaMallocAsync(&d_x32, sizeof(float) * nz, stream);
    t->d_p = d_p2;
    t->d_j = d_i2;
    t->d_x64 = d_x64;
    t->d_x32 = d_x32;
    if (e != cudaSuccess) {
        free_handle(t);
        return fail(MXG_ERR_CUDA, "csr2csc: allocation failed: %s", cudaGetErrorString(e));
    }
    int rc = csr2csc_device(A->m, A->K, A->nnz, A->d_p, A->d_j, A->d_x64, A->d_x32, d_p2, d_i2, d_x64, d_x32, stream);
    if (rc == MXG_OK) rc = csr_build_stats(t, /*validate=*/0, stream);
    if (rc != MXG_OK) {
        free_handle(t);
        return rc;
    }
    *out = t;
    return MXG_OK;
}

// the CSC of a handle (as the CSR handle of t(A)), built on first use and kept with it
static int handle_transposed(mxg_csr_s *A, int keep, cudaStream_t stream)
{
    if (A->cached_t) {
        const bool ok = (!(keep & MXG_KEEP_F64) || A->cached_t->d_x64) && (!(keep & MXG_KEEP_F32) || A->cached_t->d_x32);
        if (ok || A->nnz == 0) return MXG_OK;
        free_handle(A->cached_t);
        A->cached_t = nullptr;
    }
    int have = 0;
    if (A->d_x64) have |= MXG_KEEP_F64;
    if (A->d_x32) have |= MXG_KEEP_F32;
    if ((keep & have) != keep && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "handle lacks the values of this type");
    MXG_TRY(transpose_handle(A, have, stream, &A->cached_t));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    return MXG_OK;
}

// warm product on a cached (or explicit) handle; with the cache on, the dense operand's device copy is kept too
static int cached_spmm(DeviceState *st, mxg_csr_s *A, int dtype, int out_layout, int b_layout, int n, const void *B, size_t ldb,
                       void *Out, size_t ldc)
{
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    const size_t K = (size_t)A->K, nz = (size_t)n;
    if (!cache_enabled() || K == 0 || n <= 0 || !B) return handle_spmm_host(st, A, dtype, out_layout, b_layout, n, B, ldb, Out, ldc);
    // host bytes the operand spans (a strided operand is fingerprinted over its whole span)
    const size_t lines = b_layout == MXG_ROWS_CONTIGUOUS ? K : nz, width = b_layout == MXG_ROWS_CONTIGUOUS ? nz : K;
    if (ldb < width) return fail(MXG_ERR_ARG, "dense operand: leading dimension too small");
    const size_t host_bytes = ((lines - 1) * ldb + width) * s;
    void *d_B = cache_find_dense(B, dtype, b_layout, K, nz, ldb, host_bytes);
    if (d_B) return handle_spmm_host(st, A, dtype, out_layout, b_layout, n, B, ldb, Out, ldc, d_B, nullptr);
    void *kept = nullptr;
    const int rc = handle_spmm_host(st, A, dtype, out_layout, b_layout, n, B, ldb, Out, ldc, nullptr, &kept);
    if (kept) {
        const size_t ld_b = (nz + 16 / s - 1) / (16 / s) * (16 / s);
        if (rc == MXG_OK) cache_insert_dense(B, dtype, b_layout, K, nz, ldb, host_bytes, kept, std::max<size_t>(K * ld_b * s, 16), st->stream);
        else cudaFreeAsync(kept, st->stream);
    }
    return rc;
}

} // namespace mxg

using namespace mxg;

extern "C" {

const char *mxg_last_error(void) { return last_error_ref().c_str(); }

int mxg_device_count(int *count)
{
    if (!count) return fail(MXG_ERR_ARG, "count is NULL");
    *count = 0;
    MXG_CUDA_TRY(cudaGetDeviceCount(count));
    return MXG_OK;
}

int mxg_set_device(int device)
{
    MXG_CUDA_TRY(cudaSetDevice(device));
    return MXG_OK;
}

int mxg_set_devices(int n)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    return set_devices(n);
}

int mxg_get_devices(int *n)
{
    if (!n) return fail(MXG_ERR_ARG, "get_devices: NULL argument");
    *n = multi_devices();
    return MXG_OK;
}

int mxg_cache_clear(void)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    return cache_clear();
}

int mxg_cache_stats(unsigned long long *hits, unsigned long long *misses, size_t *bytes, int *entries)
{
    cache_stats(hits, misses, bytes, entries);
    return MXG_OK;
}

/* ---- explicit handles with host operands: the CSR stays in HBM, a product moves B up and Out down ---- */

int mxg_csr_spmm_host(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *B, size_t ldb, void *Out, size_t ldc)
{
    if (!A) return fail(MXG_ERR_ARG, "csr_spmm_host: NULL handle");
    if (dtype != MXG_F64 && dtype != MXG_F32) return fail(MXG_ERR_ARG, "bad dtype %d", dtype);
    if (n < 0) return fail(MXG_ERR_ARG, "negative n");
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    DeviceState *st;
    MXG_TRY(current_state(&st));
    return handle_spmm_host(st, A, dtype, out_layout, b_layout, n, B, ldb, Out, ldc);
}

int mxg_csr_spmm_t_host(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *B, size_t ldb, void *Out, size_t ldc)
{
    if (!A) return fail(MXG_ERR_ARG, "csr_spmm_t_host: NULL handle");
    if (dtype != MXG_F64 && dtype != MXG_F32) return fail(MXG_ERR_ARG, "bad dtype %d", dtype);
    if (n < 0) return fail(MXG_ERR_ARG, "negative n");
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    DeviceState *st;
    MXG_TRY(current_state(&st));
    MXG_TRY(handle_transposed(A, dtype == MXG_F64 ? MXG_KEEP_F64 : MXG_KEEP_F32, st->stream));
    return handle_spmm_host(st, A->cached_t, dtype, out_layout, b_layout, n, B, ldb, Out, ldc);
}

int mxg_csr_spmv_host(mxg_csr_t A, int ytype, const void *y, void *out)
{
    if (!A) return fail(MXG_ERR_ARG, "csr_spmv_host: NULL handle");
    if (ytype < MXG_Y_NUMERIC || ytype > MXG_Y_FLOAT32) return fail(MXG_ERR_ARG, "bad ytype %d", ytype);
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    DeviceState *st;
    MXG_TRY(current_state(&st));
    return handle_spmv_host(st, A, ytype, y, out);
}

static long *option_slot(const char *name)
{
    Options &o = options();
    if (!name) return nullptr;
    if (!strcmp(name, "piece")) return &o.piece;
    if (!strcmp(name, "spmm_lpr")) return &o.spmm_lpr;
    if (!strcmp(name, "spmm_panel_mb")) return &o.spmm_panel_mb;
    if (!strcmp(name, "spmm_panel_cols")) return &o.spmm_panel_cols;
    if (!strcmp(name, "spmm_rpw")) return &o.spmm_rpw;
    if (!strcmp(name, "spmm_cpl")) return &o.spmm_cpl;
    if (!strcmp(name, "spmm_bulk")) return &o.spmm_bulk;
    if (!strcmp(name, "radix_bits")) return &o.radix_bits;
    if (!strcmp(name, "host_threads")) return &o.host_threads;
    if (!strcmp(name, "host_narrow")) return &o.host_narrow;
    if (!strcmp(name, "host_stage")) return &o.host_stage;
    if (!strcmp(name, "host_colsplit")) return &o.host_colsplit;
    if (!strcmp(name, "host_pin_register")) return &o.host_pin_register;
    if (!strcmp(name, "host_pack")) return &o.host_pack;
    if (!strcmp(name, "host_pack_lag")) return &o.host_pack_lag;
    if (!strcmp(name, "pipe_slots")) return &o.pipe_slots;
    if (!strcmp(name, "host_arena_max_mb")) return &o.host_arena_max_mb;
    if (!strcmp(name, "spmv_lpr")) return &o.spmv_lpr;
    if (!strcmp(name, "spmv_tex")) return &o.spmv_tex;
    if (!strcmp(name, "svec_smem")) return &o.svec_smem;
    if (!strcmp(name, "h2d_chunk_mb")) return &o.h2d_chunk_mb;
    if (!strcmp(name, "pipeline")) return &o.pipeline;
    if (!strcmp(name, "pipe_chunk_nnz")) return &o.pipe_chunk_nnz;
    if (!strcmp(name, "multi_min_nnz")) return &o.multi_min_nnz;
    if (!strcmp(name, "multi_dense_share")) return &o.multi_dense_share;
    if (!strcmp(name, "multi_pageable")) return &o.multi_pageable;
    if (!strcmp(name, "cache_mb")) return &o.cache_mb;
    if (!strcmp(name, "host_thp")) return &o.host_thp;
    if (!strcmp(name, "host_result_pool_mb")) return &o.host_result_pool_mb;
    return nullptr;
}

int mxg_set_option(const char *name, long value)
{
    long *slot = option_slot(name);
    if (!slot) return fail(MXG_ERR_ARG, "unknown option '%s'", name ? name : "(null)");
    *slot = value;
    return MXG_OK;
}

int mxg_get_option(const char *name, long *value)
{
    long *slot = option_slot(name);
    if (!slot || !value) return fail(MXG_ERR_ARG, "unknown option '%s'", name ? name : "(null)");
    *value = *slot;
    return MXG_OK;
}

unsigned long long mxg_launch_count(void) { return g_launches.load(); }

int mxg_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr) return fail(MXG_ERR_ARG, "host_alloc: NULL argument");
    return result_pool_alloc(bytes, ptr);
}

int mxg_host_free(void *ptr)
{
    if (!ptr) return MXG_OK;
    return result_pool_free(ptr) == MXG_OK ? MXG_OK : fail(MXG_ERR_ARG, "host_free: not a block of the result pool");
}

int mxg_host_pool_stats(size_t *live_bytes, size_t *free_bytes, int *blocks)
{
    result_pool_stats(live_bytes, free_bytes, blocks);
    return MXG_OK;
}

int mxg_host_narrow(const double *src, float *dst, size_t n)
{
    if (n > 0 && (!src || !dst)) return fail(MXG_ERR_ARG, "host_narrow: NULL buffer");
    host_narrow_f64_to_f32(src, dst, n);
    return MXG_OK;
}

int mxg_host_pack_indices(const int32_t *j, size_t n, int K, void *packed, size_t *packed_bytes, int *in_range)
{
    const int hb = index_pack_hi_bits(K);
    if (packed_bytes) *packed_bytes = hb < 0 ? 0 : packed_index_bytes(n, hb);
    if (!packed) return MXG_OK; // size query
    if (hb < 0) return fail(MXG_ERR_ARG, "host_pack_indices: K = %d does not pack (ids above 2^24 travel as int32)", K);
    if (n > 0 && !j) return fail(MXG_ERR_ARG, "host_pack_indices: NULL buffer");
    if (((uintptr_t)packed & 15) != 0) return fail(MXG_ERR_ARG, "host_pack_indices: destination must be 16-byte aligned");
    const bool ok = host_pack_indices(j, n, K, hb, packed);
    if (in_range) *in_range = ok ? 1 : 0;
    return MXG_OK;
}

int mxg_last_call_bytes(size_t *h2d_bytes, size_t *d2h_bytes)
{
    last_call_bytes(h2d_bytes, d2h_bytes);
    return MXG_OK;
}

int mxg_host_chunk_plan(int m, const int32_t *p, size_t result_row_bytes, int32_t *chunk_rows, int cap, int *n_chunks,
                        int *n_long, int *n_pieces, int *max_len)
{
    if (m < 0 || !p) return fail(MXG_ERR_ARG, "chunk_plan: bad arguments");
    return host_chunk_plan(m, p, result_row_bytes, chunk_rows, cap, n_chunks, n_long, n_pieces, max_len);
}

int mxg_host_copy_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                     int streaming_stores)
{
    if (width > 0 && height > 0 && (!src || !dst)) return fail(MXG_ERR_ARG, "host_copy_2d: NULL buffer");
    if (height > 1 && (dpitch < width || spitch < width)) return fail(MXG_ERR_ARG, "host_copy_2d: pitch < width");
    host_copy_2d(dst, dpitch, src, spitch, width, height, streaming_stores != 0);
    return MXG_OK;
}

int mxg_trim(void)
{
    int dev = 0;
    MXG_CUDA_TRY(cudaGetDevice(&dev));
    MXG_CUDA_TRY(cudaDeviceSynchronize());
    cache_clear();
    texture_cache_clear();
    result_pool_trim();
    cudaMemPool_t pool;
    MXG_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    MXG_CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
    if (dev >= 0 && dev < 64 && g_dev[dev].ready) {
        MXG_TRY(pinned_arena_release(&g_dev[dev]));
        if (g_dev[dev].share_buf) {
            cudaFree(g_dev[dev].share_buf);
            g_dev[dev].share_buf = nullptr;
            g_dev[dev].share_bytes = 0;
        }
    }
    return MXG_OK;
}

/* ------------------------------------ level 2: handles ---------------------------------------- */

int mxg_csr_upload(int m, int K, const int32_t *p, const int32_t *j, const double *x, int keep, mxg_csr_t *handle)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (!handle) return fail(MXG_ERR_ARG, "handle is NULL");
    *handle = nullptr;
    DeviceState *st;
    MXG_TRY(current_state(&st));
    mxg_csr_s *h = nullptr;
    MXG_TRY(upload_csr(m, K, p, j, x, keep, st->stream, &h));
    *handle = h;
    return MXG_OK;
}

int mxg_csr_wrap_device(int m, int K, const int32_t *d_p, const int32_t *d_j, const double *d_x64,
                        const float *d_x32, int validate, void *stream, mxg_csr_t *handle)
{
    MXG_TRY(ensure_device_ready());
    if (!handle) return fail(MXG_ERR_ARG, "handle is NULL");
    *handle = nullptr;
    if (m < 0 || K < 0 || !d_p) return fail(MXG_ERR_ARG, "csr_wrap_device: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int32_t ends[2] = {0, 0};
    MXG_CUDA_TRY(cudaMemcpyAsync(&ends[0], d_p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    MXG_CUDA_TRY(cudaMemcpyAsync(&ends[1], d_p + m, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    MXG_CUDA_TRY(cudaStreamSynchronize(s));
    if (ends[0] < 0 || ends[1] < ends[0]) return fail(MXG_ERR_INDEX, "csr_wrap_device: bad indptr");
    mxg_csr_s *h = new mxg_csr_s();
    h->m = m;
    h->K = K;
    h->base = ends[0];
    h->nnz = (int64_t)ends[1] - ends[0];
    h->d_p = d_p;
    h->d_j = d_j;
    h->d_x64 = d_x64;
    h->d_x32 = d_x32;
    h->owns = false;
    h->stream = s;
    cudaGetDevice(&h->device);
    int rc = csr_build_stats(h, validate, s);
    if (rc != MXG_OK) {
        free_handle(h);
        return rc;
    }
    *handle = h;
    return MXG_OK;
}

int mxg_csr_free(mxg_csr_t handle) { return free_handle(handle); }

int mxg_csr_info(mxg_csr_t h, int64_t info[6])
{
    if (!h || !info) return fail(MXG_ERR_ARG, "csr_info: NULL argument");
    info[0] = h->m;
    info[1] = h->K;
    info[2] = h->nnz;
    info[3] = h->n_long;
    info[4] = h->n_pieces;
    info[5] = h->max_len;
    return MXG_OK;
}

int mxg_csr_device_arrays(mxg_csr_t h, const int32_t **d_p, const int32_t **d_j, const double **d_x64,
                          const float **d_x32)
{
    if (!h) return fail(MXG_ERR_ARG, "csr_device_arrays: NULL handle");
    if (d_p) *d_p = h->d_p;
    if (d_j) *d_j = h->d_j;
    if (d_x64) *d_x64 = h->d_x64;
    if (d_x32) *d_x32 = h->d_x32;
    return MXG_OK;
}

int mxg_csr_download(mxg_csr_t h, int32_t *p, int32_t *j, double *x)
{
    if (!h) return fail(MXG_ERR_ARG, "csr_download: NULL handle");
    MXG_CUDA_TRY(cudaDeviceSynchronize());
    if (p) {
        MXG_CUDA_TRY(cudaMemcpy(p, h->d_p, sizeof(int32_t) * ((size_t)h->m + 1), cudaMemcpyDeviceToHost));
        if (h->base != 0)
            for (int r = 0; r <= h->m; r++) p[r] -= h->base;
    }
    if (h->nnz > 0) {
        if (j) MXG_CUDA_TRY(cudaMemcpy(j, h->d_j + h->base, sizeof(int32_t) * (size_t)h->nnz, cudaMemcpyDeviceToHost));
        if (x) {
            if (h->d_x64) {
                MXG_CUDA_TRY(cudaMemcpy(x, h->d_x64 + h->base, sizeof(double) * (size_t)h->nnz, cudaMemcpyDeviceToHost));
            } else if (h->d_x32) {
                std::vector<float> tmp((size_t)h->nnz);
                MXG_CUDA_TRY(cudaMemcpy(tmp.data(), h->d_x32 + h->base, sizeof(float) * (size_t)h->nnz, cudaMemcpyDeviceToHost));
                for (size_t e = 0; e < (size_t)h->nnz; e++) x[e] = (double)tmp[e];
            } else {
                for (size_t e = 0; e < (size_t)h->nnz; e++) x[e] = 1.0; // pattern matrix
            }
        }
    }
    return MXG_OK;
}

int mxg_dev_spmm(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *d_B, size_t ldb,
                 void *d_Out, size_t ldc, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmm: NULL handle");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (b_layout == MXG_ROWS_CONTIGUOUS) return launch_spmm(A, dtype, out_layout, n, d_B, ldb, d_Out, ldc, s);
    if (b_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "dev_spmm: bad b_layout %d", b_layout);
    if (A->K == 0 || n <= 0) return launch_spmm(A, dtype, out_layout, n, d_B, (size_t)(n > 0 ? n : 0), d_Out, ldc, s);
    // column-major dense operand: K5 into a temporary rows-contiguous copy first
    const size_t sz = dtype == MXG_F64 ? 8 : 4;
    const size_t ld = round_up((size_t)n, 16 / sz);
    StreamBuf tmp; // freed in stream order behind the product, on every path
    MXG_CUDA_TRY(tmp.alloc((size_t)A->K * ld * sz, s));
    if (ld != (size_t)n) MXG_CUDA_TRY(cudaMemsetAsync(tmp.ptr, 0, (size_t)A->K * ld * sz, s));
    MXG_TRY(launch_transpose_dense((int)sz, (size_t)n, (size_t)A->K, d_B, ldb, tmp.ptr, ld, s));
    MXG_TRY(launch_spmm(A, dtype, out_layout, n, tmp.ptr, ld, d_Out, ldc, s));
    return MXG_OK;
}

int mxg_dev_spmm_bcast(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *d_B, size_t ldb,
                       int n_dst, void *const *d_outs, size_t ldc, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmm_bcast: NULL handle");
    if (b_layout != MXG_ROWS_CONTIGUOUS) return fail(MXG_ERR_UNSUPPORTED, "dev_spmm_bcast: the dense operand must be rows-contiguous");
    return launch_spmm_multi(A, dtype, out_layout, n, d_B, ldb, n_dst, d_outs, ldc, static_cast<cudaStream_t>(stream));
}

/* One slice of a product on a handle: rows [r0, r1) without their long rows (pieces == 0), or only the long rows
 * (pieces != 0: piece kernels + fix-up, r0 / r1 ignored).  d_Out is the origin of the FULL result in both cases.
 * Lets a caller pipeline its own exchange of finished row slices (sharded.PipelinedColumnMajorGather). */
int mxg_dev_spmm_rows(mxg_csr_t A, int dtype, int out_layout, int n, const void *d_B, size_t ldb, void *d_Out, size_t ldc,
                      int r0, int r1, int pieces, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmm_rows: NULL handle");
    if (A->m == 0 || n <= 0 || A->nnz == 0) return fail(MXG_ERR_ARG, "dev_spmm_rows: empty product");
    if (out_layout == MXG_ROWS_CONTIGUOUS ? ldc < (size_t)n : ldc < (size_t)A->m) return fail(MXG_ERR_ARG, "dev_spmm_rows: ldc too small");
    if (ldb < (size_t)n) return fail(MXG_ERR_ARG, "dev_spmm_rows: ldb < n");
    return launch_spmm_rows(A, dtype, out_layout, n, d_B, ldb, d_Out, ldc, r0, r1, pieces, static_cast<cudaStream_t>(stream));
}

/* Strided device-to-device copy on `stream` (cudaMemcpy2DAsync): `height` lines of `width_bytes`.  Runs on a copy engine,
 * not on SMs: how sharded.PipelinedColumnMajorGather packs a row slice of a column-major block and unpacks the gathered
 * slices into their rows of the global result while the product kernel keeps the SMs. */
int mxg_dev_copy_2d(void *d_dst, size_t dpitch, const void *d_src, size_t spitch, size_t width_bytes, size_t height, void *stream)
{
    if (width_bytes == 0 || height == 0) return MXG_OK;
    if (!d_dst || !d_src || dpitch < width_bytes || spitch < width_bytes) return fail(MXG_ERR_ARG, "dev_copy_2d: bad arguments");
    MXG_CUDA_TRY(cudaMemcpy2DAsync(d_dst, dpitch, d_src, spitch, width_bytes, height, cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(stream)));
    return MXG_OK;
}

/* Product + all-gather with the copy engines: the product runs in row slices (about 16 of equal nnz) into d_outs[0];
 * every finished slice is pushed to the other destinations by DMA (peer copies over NVLink, 2-D for column-major
 * results) on the library's three copy streams while the next slice is being computed.  SM stores to peers carry
 * 128-byte requests (~550 GB/s into a B200); the copy engines move the same bytes in bulk. */
int mxg_dev_spmm_push(mxg_csr_t A, int dtype, int out_layout, int b_layout, int n, const void *d_B, size_t ldb, int n_dst,
                      void *const *d_outs, size_t ldc, void *stream)
{
    if (!A) return fail(MXG_ERR_ARG, "dev_spmm_push: NULL handle");
    if (b_layout != MXG_ROWS_CONTIGUOUS) return fail(MXG_ERR_UNSUPPORTED, "dev_spmm_push: the dense operand must be rows-contiguous");
    if (n_dst < 1 || n_dst > MXG_MAX_DST || !d_outs) return fail(MXG_ERR_ARG, "dev_spmm_push: 1 .. %d destinations", MXG_MAX_DST);
    if (dtype != MXG_F64 && dtype != MXG_F32) return fail(MXG_ERR_ARG, "bad dtype %d", dtype);
    if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "bad out_layout %d", out_layout);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (A->m == 0 || n <= 0) return MXG_OK;
    const bool rm = out_layout == MXG_ROWS_CONTIGUOUS;
    if (rm ? ldc < (size_t)n : ldc < (size_t)A->m) return fail(MXG_ERR_ARG, "dev_spmm_push: ldc too small");
    DeviceState *st;
    MXG_TRY(current_state(&st));
    if (A->nnz == 0 || n_dst == 1) { // nothing to overlap: one launch (a zero fill for an empty matrix) per destination list
        MXG_TRY(launch_spmm_multi(A, dtype, out_layout, n, d_B, ldb, n_dst, d_outs, ldc, s));
        return MXG_OK;
    }
    MXG_TRY(handle_chunks(A, s));
    const std::vector<int32_t> &cr = *A->host_chunks;
    const int C = (int)cr.size() - 1;
    const size_t es = dtype == MXG_F64 ? 8 : 4;
    // one copy stream per destination: the pushes to different peers run on different copy engines at the same time
    while ((int)st->push_streams.size() < MXG_MAX_DST - 1) {
        cudaStream_t q;
        MXG_CUDA_TRY(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
        st->push_streams.push_back(q);
    }
    const std::vector<cudaStream_t> &copy = st->push_streams;
    while ((int)st->ev_pool.size() < C + MXG_MAX_DST) {
        cudaEvent_t e;
        MXG_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st->ev_pool.push_back(e);
    }
    // long rows first: their rows are written by the fix-up launch, anywhere in the block, so slices wait for it too
    MXG_TRY(launch_spmm_rows(A, dtype, out_layout, n, d_B, ldb, d_outs[0], ldc, 0, 0, /*pieces=*/1, s));
    for (int c = 0; c < C; c++) {
        const size_t r0 = (size_t)cr[(size_t)c], nr = (size_t)cr[(size_t)c + 1] - r0;
        if (nr == 0) continue;
        MXG_TRY(launch_spmm_rows(A, dtype, out_layout, n, d_B, ldb, d_outs[0], ldc, (int)r0, (int)(r0 + nr), /*pieces=*/0, s));
        cudaEvent_t done = st->ev_pool[(size_t)c];
        MXG_CUDA_TRY(cudaEventRecord(done, s));
        const size_t off = rm ? r0 * ldc * es : r0 * es;
        for (int d = 1; d < n_dst; d++) {
            cudaStream_t q = copy[(size_t)(d - 1)];
            MXG_CUDA_TRY(cudaStreamWaitEvent(q, done, 0));
            char *dst = static_cast<char *>(d_outs[d]) + off;
            const char *src = static_cast<const char *>(d_outs[0]) + off;
            if (rm && ldc == (size_t)n) MXG_CUDA_TRY(cudaMemcpyAsync(dst, src, nr * ldc * es, cudaMemcpyDefault, q));
            else if (rm) MXG_CUDA_TRY(cudaMemcpy2DAsync(dst, ldc * es, src, ldc * es, (size_t)n * es, nr, cudaMemcpyDefault, q));
            else MXG_CUDA_TRY(cudaMemcpy2DAsync(dst, ldc * es, src, ldc * es, nr * es, (size_t)n, cudaMemcpyDefault, q));
        }
    }
    for (int d = 1; d < n_dst; d++) { // the caller's stream continues when every push has landed
        cudaEvent_t e = st->ev_pool[(size_t)C + (size_t)d];
        MXG_CUDA_TRY(cudaEventRecord(e, copy[(size_t)(d - 1)]));
        MXG_CUDA_TRY(cudaStreamWaitEvent(s, e, 0));
    }
    return MXG_OK;
}

int mxg_dev_spmm_mcast(mxg_csr_t A, int dtype, int n, const void *d_B, size_t ldb, void *mc_out, size_t ldc, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmm_mcast: NULL handle");
    if (!mc_out) return fail(MXG_ERR_ARG, "dev_spmm_mcast: NULL multicast address");
    void *outs[1] = {mc_out};
    return launch_spmm_multi(A, dtype, MXG_ROWS_CONTIGUOUS, n, d_B, ldb, 1, outs, ldc, static_cast<cudaStream_t>(stream), 1);
}

int mxg_dev_spmv_bcast(mxg_csr_t A, int ytype, const void *d_y, int n_dst, void *const *d_outs, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmv_bcast: NULL handle");
    return launch_spmv_multi(A, ytype, d_y, n_dst, d_outs, static_cast<cudaStream_t>(stream));
}

int mxg_dev_alloc(size_t bytes, void **d_ptr)
{
    if (!d_ptr) return fail(MXG_ERR_ARG, "dev_alloc: NULL argument");
    *d_ptr = nullptr;
    MXG_CUDA_TRY(cudaMalloc(d_ptr, bytes > 0 ? bytes : 16)); // not pool memory: cudaIpcGetMemHandle needs cudaMalloc
    return MXG_OK;
}

int mxg_dev_free(void *d_ptr)
{
    if (d_ptr) MXG_CUDA_TRY(cudaFree(d_ptr));
    return MXG_OK;
}

int mxg_ipc_export(const void *d_ptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!d_ptr || !handle) return fail(MXG_ERR_ARG, "ipc_export: NULL argument");
    cudaIpcMemHandle_t h;
    MXG_CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, 64);
    return MXG_OK;
}

int mxg_ipc_open(const unsigned char handle[64], void **d_ptr)
{
    if (!d_ptr || !handle) return fail(MXG_ERR_ARG, "ipc_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    *d_ptr = nullptr;
    MXG_CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MXG_OK;
}

int mxg_ipc_close(void *d_ptr)
{
    if (d_ptr) MXG_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return MXG_OK;
}

int mxg_dev_peer_barrier(int rank, int world, int *const *peer_flags, int epoch, void *stream)
{
    return launch_peer_barrier(rank, world, peer_flags, epoch, static_cast<cudaStream_t>(stream));
}

int mxg_dev_barrier_failed(int *failed)
{
    if (!failed) return fail(MXG_ERR_ARG, "barrier_failed: NULL argument");
    return peer_barrier_failed(failed);
}

int mxg_dev_spmv(mxg_csr_t A, int ytype, const void *d_y, void *d_out, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmv: NULL handle");
    return launch_spmv(A, ytype, d_y, d_out, static_cast<cudaStream_t>(stream));
}

int mxg_dev_csr2csc(mxg_csr_t A, int keep, void *stream, mxg_csr_t *At)
{
    MXG_TRY(ensure_device_ready());
    if (!A || !At) return fail(MXG_ERR_ARG, "dev_csr2csc: NULL argument");
    *At = nullptr;
    mxg_csr_s *t = nullptr;
    MXG_TRY(transpose_handle(A, keep, static_cast<cudaStream_t>(stream), &t));
    *At = t;
    return MXG_OK;
}

int mxg_dev_transpose_dense(int elem_size, size_t rows, size_t cols, const void *d_src, size_t ld_src, void *d_dst,
                            size_t ld_dst, void *stream)
{
    MXG_TRY(ensure_device_ready());
    return launch_transpose_dense(elem_size, rows, cols, d_src, ld_src, d_dst, ld_dst, static_cast<cudaStream_t>(stream));
}

int mxg_row_partition(int m, const int32_t *p, int parts, int32_t *row_starts)
{
    if (m < 0 || parts <= 0 || !p || !row_starts) return fail(MXG_ERR_ARG, "row_partition: bad arguments");
    return row_partition(m, p, parts, row_starts);
}

int mxg_dev_gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed,
                         float *d_sink, long long *gathers_done, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!d_table || !d_sink) return fail(MXG_ERR_ARG, "gather_probe: NULL buffer");
    MXG_TRY(gather_probe(row_bytes, d_table, rows, gathers, seed, d_sink, static_cast<cudaStream_t>(stream)));
    if (gathers_done) {
        const long long teams = 148LL * 8 * 256 / (row_bytes / 16);
        *gathers_done = teams * std::max<long long>(8, (gathers / teams + 7) / 8 * 8);
    }
    return MXG_OK;
}

int mxg_dev_tma_gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed, float *d_sink,
                             long long *gathers_done, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!d_table || !d_sink) return fail(MXG_ERR_ARG, "tma_gather_probe: NULL buffer");
    return tma_gather_probe(row_bytes, d_table, rows, gathers, seed, d_sink, gathers_done, static_cast<cudaStream_t>(stream));
}

int mxg_dev_spmv_probe(mxg_csr_t A, int mode, const double *d_y, double *d_sink, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A || !d_sink) return fail(MXG_ERR_ARG, "spmv_probe: NULL argument");
    return spmv_probe(A, mode, d_y, d_sink, static_cast<cudaStream_t>(stream));
}

int mxg_synth_csr(int m, int K, int64_t target_nnz, int row_model, int col_model, uint64_t seed, int keep,
                  void *stream, mxg_csr_t *handle)
{
    MXG_TRY(ensure_device_ready());
    if (!handle) return fail(MXG_ERR_ARG, "handle is NULL");
    *handle = nullptr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int32_t *d_p = nullptr, *d_j = nullptr;
    double *d_x64 = nullptr;
    float *d_x32 = nullptr;
    int64_t nnz = 0;
    MXG_TRY(synth_csr_arrays(m, K, target_nnz, row_model, col_model, seed, keep, s, &d_p, &d_j, &d_x64, &d_x32, &nnz));
    mxg_csr_s *h = new mxg_csr_s();
    h->m = m;
    h->K = K;
    h->nnz = nnz;
    h->base = 0;
    h->d_p = d_p;
    h->d_j = d_j;
    h->d_x64 = d_x64;
    h->d_x32 = d_x32;
    h->owns = true;
    cudaGetDevice(&h->device);
    int rc = csr_build_stats(h, /*validate=*/0, s);
    if (rc != MXG_OK) {
        free_handle(h);
        return rc;
    }
    *handle = h;
    return MXG_OK;
}

/* ------------------------------------ level 1: host buffers ----------------------------------- */

int mxg_spmm_csr_dense(int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p,
                       const int32_t *j, const double *x, const void *B, size_t ldb, void *Out, size_t ldc)
{
    if (dtype != MXG_F64 && dtype != MXG_F32) return fail(MXG_ERR_ARG, "bad dtype %d", dtype);
    if (n < 0) return fail(MXG_ERR_ARG, "negative n");
    if (m < 0 || K < 0) return fail(MXG_ERR_ARG, "csr: negative dimension");
    if (!p) return fail(MXG_ERR_ARG, "csr: indptr is NULL");
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    DeviceState *st;
    MXG_TRY(current_state(&st));
    // several devices (mxg_set_devices): nnz-balanced row blocks, one host thread and one streamed pipeline per device
    if (options().pipeline != 0 && multi_wanted(m, p, j, x) && (int64_t)p[m] >= (int64_t)p[0] && p[0] >= 0)
        return multi_spmm(dtype, out_layout, b_layout, m, K, n, p, j, x, B, ldb, Out, ldc);
    // operand cache (option cache_mb): the device CSR of these host arrays is kept between calls
    if (p[0] == 0 && options().pipeline != 0 && cache_enabled() && m > 0 && n > 0 && p[m] > 0 && j && x) {
        const int need = dtype == MXG_F64 ? MXG_KEEP_F64 : MXG_KEEP_F32;
        mxg_csr_s *hit = cache_find_csr(m, K, p, j, x, need);
        if (!hit) { // cold: the streamed call, which hands its device arrays over instead of releasing them
            mxg_csr_s *kept = nullptr;
            MXG_TRY(pipeline_spmm(st, dtype, out_layout, b_layout, m, K, n, p, j, x, B, ldb, Out, ldc, nullptr, 0, &kept));
            cache_insert_csr(m, K, p, j, x, kept);
            return MXG_OK;
        }
        return cached_spmm(st, hit, dtype, out_layout, b_layout, n, B, ldb, Out, ldc);
    }
    // every valid R matrix has p[0] == 0 (R/utils.R:349-410): streamed, chunk-overlapped path
    if (p[0] == 0 && options().pipeline != 0)
        return pipeline_spmm(st, dtype, out_layout, b_layout, m, K, n, p, j, x, B, ldb, Out, ldc);
    mxg_csr_s *A = nullptr;
    MXG_TRY(upload_csr(m, K, p, j, x, dtype == MXG_F64 ? MXG_KEEP_F64 : MXG_KEEP_F32, st->stream, &A));
    int rc = spmm_host_io(A, dtype, out_layout, b_layout, n, B, ldb, Out, ldc, st->stream);
    cudaStreamSynchronize(st->stream);
    free_handle(A);
    return rc;
}

int mxg_spmm_csrT_dense(int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p,
                        const int32_t *j, const double *x, const void *B, size_t ldb, void *Out, size_t ldc)
{
    if (dtype != MXG_F64 && dtype != MXG_F32) return fail(MXG_ERR_ARG, "bad dtype %d", dtype);
    if (n < 0) return fail(MXG_ERR_ARG, "negative n");
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    DeviceState *st;
    MXG_TRY(current_state(&st));
    const int keep = dtype == MXG_F64 ? MXG_KEEP_F64 : MXG_KEEP_F32;
    if (cache_enabled() && p && m > 0 && K > 0 && n > 0 && p[0] == 0 && p[m] > 0 && j && x) {
        // operand cache: the matrix AND its device-built CSC stay resident; a repeated crossprod only moves Y and the result
        mxg_csr_s *hit = cache_find_csr(m, K, p, j, x, keep);
        if (!hit) {
            MXG_TRY(upload_csr(m, K, p, j, x, keep, st->stream, &hit));
            cache_insert_csr(m, K, p, j, x, hit);
            hit = cache_find_csr(m, K, p, j, x, keep); // NULL when it did not fit the budget (and was released)
        }
        if (hit) {
            MXG_TRY(handle_transposed(hit, keep, st->stream));
            cache_account_csr(hit);
            return cached_spmm(st, hit->cached_t, dtype, out_layout, b_layout, n, B, ldb, Out, ldc);
        }
    }
    mxg_csr_s *A = nullptr, *At = nullptr;
    const char *tr = getenv("MXG_TRACE"); // development aid: host-side phase times on stderr
    const bool trace = tr && *tr && *tr != '0';
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    MXG_TRY(upload_csr(m, K, p, j, x, keep, st->stream, &A));
    const double t1 = now();
    // A page-locked dense operand (m rows: the columns of t(A)) starts to cross PCIe on the upload stream now, while
    // the CSR is transposed on the compute stream; a pageable one needs the host threads and waits its turn below.
    void *d_B = nullptr;
    size_t ld_b = 0;
    cudaEvent_t b_ready = nullptr;
    int rc = MXG_OK;
    if (m > 0 && K > 0 && n > 0 && B && Out && host_is_pinned(B)) {
        rc = upload_dense_rows(dtype, b_layout, (size_t)m, (size_t)n, B, ldb, st->h2d, &d_B, &ld_b);
        if (rc == MXG_OK && cudaEventCreateWithFlags(&b_ready, cudaEventDisableTiming) != cudaSuccess) rc = fail(MXG_ERR_CUDA, "csrT: event");
        if (rc == MXG_OK && cudaEventRecord(b_ready, st->h2d) != cudaSuccess) rc = fail(MXG_ERR_CUDA, "csrT: event record");
    }
    auto drop_early = [&]() {
        cudaStreamSynchronize(st->h2d);
        if (d_B) cudaFreeAsync(d_B, st->stream);
        if (b_ready) cudaEventDestroy(b_ready);
        d_B = nullptr;
        b_ready = nullptr;
    };
    if (rc == MXG_OK) rc = transpose_handle(A, keep, st->stream, &At);
    cudaStreamSynchronize(st->stream);
    free_handle(A);
    if (rc != MXG_OK) {
        drop_early();
        return rc;
    }
    const double t2 = now();
    rc = spmm_host_io(At, dtype, out_layout, b_layout, n, B, ldb, Out, ldc, st->stream, d_B, ld_b, b_ready);
    cudaStreamSynchronize(st->stream);
    if (b_ready) cudaEventDestroy(b_ready); // (spmm_host_io has released the operand)
    free_handle(At);
    if (trace)
        fprintf(stderr, "[mxg trace] csrT: upload+stats %.2f ms | transpose+stats %.2f ms | dense upload, product, download %.2f ms\n",
                t1 - t0, t2 - t1, now() - t2);
    return rc;
}

int mxg_spmv_csr(int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x, const void *y, void *out)
{
    if (ytype < MXG_Y_NUMERIC || ytype > MXG_Y_FLOAT32) return fail(MXG_ERR_ARG, "bad ytype %d", ytype);
    if (m < 0 || K < 0) return fail(MXG_ERR_ARG, "csr: negative dimension");
    if (!p) return fail(MXG_ERR_ARG, "csr: indptr is NULL");
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    DeviceState *st;
    MXG_TRY(current_state(&st));
    if (options().pipeline != 0 && multi_wanted(m, p, j, x) && (int64_t)p[m] >= (int64_t)p[0] && p[0] >= 0)
        return multi_spmv(ytype, m, K, p, j, x, y, out);
    if (p[0] == 0 && options().pipeline != 0 && cache_enabled() && m > 0 && p[m] > 0 && j && x) {
        mxg_csr_s *hit = cache_find_csr(m, K, p, j, x, MXG_KEEP_F64);
        if (hit) return handle_spmv_host(st, hit, ytype, y, out);
        mxg_csr_s *kept = nullptr;
        MXG_TRY(pipeline_spmv(st, ytype, m, K, p, j, x, y, out, &kept));
        cache_insert_csr(m, K, p, j, x, kept);
        return MXG_OK;
    }
    if (p[0] == 0 && options().pipeline != 0) return pipeline_spmv(st, ytype, m, K, p, j, x, y, out);
    mxg_csr_s *A = nullptr;
    MXG_TRY(upload_csr(m, K, p, j, x, MXG_KEEP_F64, st->stream, &A));
    int rc = MXG_OK;
    if (m > 0) {
        const size_t ys = ytype == MXG_Y_NUMERIC ? 8 : 4;
        const size_t os = ytype == MXG_Y_FLOAT32 ? 4 : 8;
        cudaStream_t s = st->stream;
        auto body = [&]() -> int {
            if (!out) return fail(MXG_ERR_ARG, "output is NULL");
            if (K > 0 && !y) return fail(MXG_ERR_ARG, "vector is NULL");
            StreamBuf d_y, d_out;
            MXG_CUDA_TRY(d_y.alloc((size_t)K * ys, s));
            MXG_CUDA_TRY(d_out.alloc((size_t)m * os, s));
            if (K > 0) MXG_CUDA_TRY(cudaMemcpyAsync(d_y.ptr, y, (size_t)K * ys, cudaMemcpyHostToDevice, s));
            MXG_TRY(launch_spmv(A, ytype, d_y.ptr, d_out.ptr, s));
            MXG_CUDA_TRY(cudaMemcpyAsync(out, d_out.ptr, (size_t)m * os, cudaMemcpyDeviceToHost, s));
            MXG_CUDA_TRY(cudaStreamSynchronize(s));
            return MXG_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st->stream);
    free_handle(A);
    return rc;
}

int mxg_dev_spmv_svec(mxg_csr_t A, int ytype, int n_y, const int32_t *d_yidx_base1, const void *d_yvals, double *d_out,
                      void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_spmv_svec: NULL handle");
    if (A->m > 0 && !d_out) return fail(MXG_ERR_ARG, "dev_spmv_svec: output is NULL");
    return launch_spmv_svec(A, ytype, A->K, n_y, d_yidx_base1, d_yvals, d_out, static_cast<cudaStream_t>(stream));
}

int mxg_spmv_csr_svec(int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x, int n_y,
                      const int32_t *y_idx_base1, const void *y_vals, double *out)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (ytype < MXG_Y_NUMERIC || ytype > MXG_Y_BINARY) return fail(MXG_ERR_ARG, "bad ytype %d", ytype);
    if (m < 0 || n_y < 0) return fail(MXG_ERR_ARG, "svec: negative size");
    if (!p) return fail(MXG_ERR_ARG, "csr: indptr is NULL");
    if (m > 0 && !out) return fail(MXG_ERR_ARG, "output is NULL");
    if (n_y > 0 && (!y_idx_base1 || (!y_vals && ytype != MXG_Y_BINARY))) return fail(MXG_ERR_ARG, "sparse vector is NULL");
    if (m == 0) return MXG_OK;
    // columns the presence bitmap has to cover: all of A's when K is known, else up to the largest index of y
    int kmask = K;
    if (K <= 0) {
        kmask = 0;
        for (int k = 0; k < n_y; k++) kmask = std::max(kmask, y_idx_base1[k]);
    }
    if (n_y == 0 || kmask <= 0 || p[m] - p[0] <= 0) { // src/matmul.cpp:495-496
        std::memset(out, 0, sizeof(double) * (size_t)m);
        return MXG_OK;
    }
    DeviceState *st;
    MXG_TRY(current_state(&st));
    mxg_csr_s *A = nullptr;
    MXG_TRY(upload_csr(m, K > 0 ? K : INT32_MAX, p, j, x, MXG_KEEP_F64, st->stream, &A));
    cudaStream_t s = st->stream;
    int32_t *d_yi = nullptr;
    void *d_yv = nullptr;
    double *d_out = nullptr;
    const size_t vs = ytype == MXG_Y_NUMERIC ? 8 : 4;
    auto body = [&]() -> int {
        MXG_CUDA_TRY(cudaMallocAsync(&d_yi, sizeof(int32_t) * (size_t)n_y, s));
        MXG_CUDA_TRY(cudaMemcpyAsync(d_yi, y_idx_base1, sizeof(int32_t) * (size_t)n_y, cudaMemcpyHostToDevice, s));
        if (ytype != MXG_Y_BINARY) {
            MXG_CUDA_TRY(cudaMallocAsync(&d_yv, vs * (size_t)n_y, s));
            MXG_CUDA_TRY(cudaMemcpyAsync(d_yv, y_vals, vs * (size_t)n_y, cudaMemcpyHostToDevice, s));
        }
        MXG_CUDA_TRY(cudaMallocAsync(&d_out, sizeof(double) * (size_t)m, s));
        MXG_TRY(launch_spmv_svec(A, ytype, kmask, n_y, d_yi, d_yv, d_out, s));
        MXG_CUDA_TRY(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s));
        MXG_CUDA_TRY(cudaStreamSynchronize(s));
        return MXG_OK;
    };
    const int rc = body();
    if (d_yi) cudaFreeAsync(d_yi, s);
    if (d_yv) cudaFreeAsync(d_yv, s);
    if (d_out) cudaFreeAsync(d_out, s);
    cudaStreamSynchronize(s);
    free_handle(A);
    return rc;
}

// ---- validity, sorting, elementwise products -------------------------------------------

const char *mxg_csr_error_string(int code)
{
    switch (code) { // src/misc.cpp:977-1013, verbatim messages of the reference
    case 0: return "";
    case 1: return "Matrix has negative indices.";
    case 2: return "Matrix has invalid column indices.";
    case 3: return "Matrix has indices with missing values.";
    case 4: return "Matrix has missing values in the index pointer.";
    case 5: return "Matrix index pointer is not monotonicaly increasing.";
    default: return "unknown";
    }
}

int mxg_dev_check_valid_csr(int m, int ncols, const int32_t *d_p, const int32_t *d_j, int64_t nnz, int *code, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!code || m < 0 || nnz < 0 || !d_p || (nnz > 0 && !d_j)) return fail(MXG_ERR_ARG, "check_valid_csr: bad argument");
    return dev_check_valid_csr(m, ncols, d_p, d_j, nnz, code, static_cast<cudaStream_t>(stream));
}

int mxg_dev_rows_sorted(int m, const int32_t *d_p, const int32_t *d_j, int *sorted, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!sorted || m < 0 || !d_p) return fail(MXG_ERR_ARG, "rows_sorted: bad argument");
    return dev_rows_sorted(m, d_p, d_j, sorted, static_cast<cudaStream_t>(stream));
}

int mxg_dev_sort_csr_indices(int m, const int32_t *d_p, const int32_t *d_j, const double *d_x, int32_t *d_j_out,
                             double *d_x_out, int *rows_sorted, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (m < 0) return fail(MXG_ERR_ARG, "sort_csr_indices: negative dimension");
    return dev_sort_csr_indices(m, d_p, d_j, d_x, d_j_out, d_x_out, rows_sorted, static_cast<cudaStream_t>(stream));
}

int mxg_dev_mul_csr_dense(mxg_csr_t A, int dtype, const void *d_dense, double *d_values_out, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_mul_csr_dense: NULL handle");
    return launch_mul_csr_dense(A, dtype, d_dense, d_values_out, static_cast<cudaStream_t>(stream));
}

int mxg_dev_mul_csr_dvec(mxg_csr_t A, const double *d_dvec, size_t len, double *d_values_out, void *stream)
{
    MXG_TRY(ensure_device_ready());
    if (!A) return fail(MXG_ERR_ARG, "dev_mul_csr_dvec: NULL handle");
    return launch_mul_csr_dvec(A, d_dvec, len, d_values_out, static_cast<cudaStream_t>(stream));
}

namespace {
// small RAII holder for the stream-ordered temporaries of the host-buffer entry points below
struct DevTemps {
    cudaStream_t s;
    std::vector<void *> ptrs;
    explicit DevTemps(cudaStream_t stream) : s(stream) {}
    int alloc(void **out, size_t bytes)
    {
        *out = nullptr;
        MXG_CUDA_TRY(cudaMallocAsync(out, std::max<size_t>(bytes, 16), s));
        ptrs.push_back(*out);
        return MXG_OK;
    }
    int upload(void **out, const void *src, size_t bytes)
    {
        MXG_TRY(alloc(out, bytes));
        if (bytes) MXG_CUDA_TRY(cudaMemcpyAsync(*out, src, bytes, cudaMemcpyHostToDevice, s));
        return MXG_OK;
    }
    ~DevTemps()
    {
        for (void *q : ptrs) cudaFreeAsync(q, s);
        cudaStreamSynchronize(s);
    }
};
} // namespace

int mxg_check_valid_csr(int m, int ncols, const int32_t *p, const int32_t *j, int64_t nnz, int *code)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (!code || m < 0 || nnz < 0 || !p || (nnz > 0 && !j)) return fail(MXG_ERR_ARG, "check_valid_csr: bad argument");
    DeviceState *st;
    MXG_TRY(current_state(&st));
    DevTemps t(st->stream);
    void *d_p, *d_j;
    MXG_TRY(t.upload(&d_p, p, sizeof(int32_t) * ((size_t)m + 1)));
    MXG_TRY(t.upload(&d_j, j, sizeof(int32_t) * (size_t)nnz));
    return dev_check_valid_csr(m, ncols, static_cast<int32_t *>(d_p), static_cast<int32_t *>(d_j), nnz, code, st->stream);
}

int mxg_rows_sorted(int m, const int32_t *p, const int32_t *j, int *sorted)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (!sorted || m < 0 || !p) return fail(MXG_ERR_ARG, "rows_sorted: bad argument");
    *sorted = 1;
    const int64_t nnz = (int64_t)p[m] - p[0];
    if (m == 0 || nnz <= 0) return MXG_OK;
    if (!j) return fail(MXG_ERR_ARG, "rows_sorted: indices is NULL");
    DeviceState *st;
    MXG_TRY(current_state(&st));
    DevTemps t(st->stream);
    void *d_p, *d_j;
    MXG_TRY(t.upload(&d_p, p, sizeof(int32_t) * ((size_t)m + 1)));
    MXG_TRY(t.upload(&d_j, j, sizeof(int32_t) * (size_t)p[m])); // offsets are absolute: keep j's origin
    return dev_rows_sorted(m, static_cast<int32_t *>(d_p), static_cast<int32_t *>(d_j), sorted, st->stream);
}

int mxg_sort_csr_indices(int m, const int32_t *p, int32_t *j, double *x)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (m < 0 || !p) return fail(MXG_ERR_ARG, "sort_csr_indices: bad argument");
    if (m == 0 || p[m] <= 0) return MXG_OK;
    if (!j) return fail(MXG_ERR_ARG, "sort_csr_indices: indices is NULL");
    for (int r = 0; r < m; r++)
        if (p[r] < 0 || p[r] > p[r + 1]) return fail(MXG_ERR_INDEX, "sort_csr_indices: indptr is negative or decreasing");
    const size_t n = (size_t)p[m];
    DeviceState *st;
    MXG_TRY(current_state(&st));
    cudaStream_t s = st->stream;
    DevTemps t(s);
    void *d_p, *d_j, *d_x = nullptr, *d_j2, *d_x2 = nullptr;
    MXG_TRY(t.upload(&d_p, p, sizeof(int32_t) * ((size_t)m + 1)));
    MXG_TRY(t.upload(&d_j, j, sizeof(int32_t) * n));
    MXG_TRY(t.alloc(&d_j2, sizeof(int32_t) * n));
    if (x) {
        MXG_TRY(t.upload(&d_x, x, sizeof(double) * n));
        MXG_TRY(t.alloc(&d_x2, sizeof(double) * n));
    }
    int changed = 0;
    MXG_TRY(dev_sort_csr_indices(m, static_cast<int32_t *>(d_p), static_cast<int32_t *>(d_j), static_cast<double *>(d_x),
                                 static_cast<int32_t *>(d_j2), static_cast<double *>(d_x2), &changed, s));
    if (changed > 0) { // rows before p[0] (none for an R matrix) are not touched
        const size_t off = (size_t)p[0];
        MXG_CUDA_TRY(cudaMemcpyAsync(j + off, static_cast<int32_t *>(d_j2) + off, sizeof(int32_t) * (n - off), cudaMemcpyDeviceToHost, s));
        if (x) MXG_CUDA_TRY(cudaMemcpyAsync(x + off, static_cast<double *>(d_x2) + off, sizeof(double) * (n - off), cudaMemcpyDeviceToHost, s));
        MXG_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return MXG_OK;
}

int mxg_mul_csr_dense(int dtype, int m, int K, const int32_t *p, const int32_t *j, const double *x, const void *dense,
                      double *values_out)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (dtype < MXG_Y_NUMERIC || dtype > MXG_Y_FLOAT32) return fail(MXG_ERR_ARG, "mul_csr_dense: bad element type %d", dtype);
    if (m < 0 || K < 0 || !p) return fail(MXG_ERR_ARG, "mul_csr_dense: bad argument");
    const int64_t nnz = (int64_t)p[m] - p[0];
    if (m == 0 || nnz <= 0) return MXG_OK;
    if (!dense || !values_out) return fail(MXG_ERR_ARG, "mul_csr_dense: NULL operand");
    DeviceState *st;
    MXG_TRY(current_state(&st));
    mxg_csr_s *A = nullptr;
    MXG_TRY(upload_csr(m, K, p, j, x, MXG_KEEP_F64, st->stream, &A));
    int rc;
    {
        DevTemps t(st->stream);
        void *d_dense, *d_out;
        const size_t es = dtype == MXG_Y_NUMERIC ? 8 : 4;
        auto body = [&]() -> int {
            MXG_TRY(t.upload(&d_dense, dense, es * (size_t)m * (size_t)K));
            MXG_TRY(t.alloc(&d_out, sizeof(double) * (size_t)nnz));
            MXG_TRY(launch_mul_csr_dense(A, dtype, d_dense, static_cast<double *>(d_out), st->stream));
            MXG_CUDA_TRY(cudaMemcpyAsync(values_out, d_out, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st->stream));
            MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
            return MXG_OK;
        };
        rc = body();
    }
    free_handle(A);
    return rc;
}

int mxg_mul_csr_dvec(int m, int K, const int32_t *p, const int32_t *j, const double *x, const double *dvec, size_t len,
                     double *values_out)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (m < 0 || K < 0 || !p) return fail(MXG_ERR_ARG, "mul_csr_dvec: bad argument");
    const int64_t nnz = (int64_t)p[m] - p[0];
    if (m == 0 || nnz <= 0) return MXG_OK;
    if (!dvec || !values_out || len == 0) return fail(MXG_ERR_ARG, "mul_csr_dvec: NULL or empty operand");
    DeviceState *st;
    MXG_TRY(current_state(&st));
    mxg_csr_s *A = nullptr;
    MXG_TRY(upload_csr(m, K, p, j, x, MXG_KEEP_F64, st->stream, &A));
    int rc;
    {
        DevTemps t(st->stream);
        void *d_vec, *d_out;
        auto body = [&]() -> int {
            MXG_TRY(t.upload(&d_vec, dvec, sizeof(double) * len));
            MXG_TRY(t.alloc(&d_out, sizeof(double) * (size_t)nnz));
            MXG_TRY(launch_mul_csr_dvec(A, static_cast<double *>(d_vec), len, static_cast<double *>(d_out), st->stream));
            MXG_CUDA_TRY(cudaMemcpyAsync(values_out, d_out, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st->stream));
            MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
            return MXG_OK;
        };
        rc = body();
    }
    free_handle(A);
    return rc;
}

int mxg_csr2csc(int m, int K, const int32_t *p, const int32_t *j, const double *x, int32_t *p2, int32_t *i2, double *x2)
{
    std::lock_guard<std::recursive_mutex> level1(g_level1_mu);
    if (!p2) return fail(MXG_ERR_ARG, "p2 is NULL");
    DeviceState *st;
    MXG_TRY(current_state(&st));
    const bool vals = x && x2;
    mxg_csr_s *A = nullptr, *At = nullptr;
    MXG_TRY(upload_csr(m, K, p, j, vals ? x : nullptr, vals ? MXG_KEEP_F64 : 0, st->stream, &A));
    int rc = transpose_handle(A, MXG_KEEP_F64, st->stream, &At);
    if (rc == MXG_OK) {
        cudaStream_t s = st->stream;
        auto body = [&]() -> int {
            MXG_TRY(staged_d2h(st, p2, At->d_p, sizeof(int32_t) * ((size_t)K + 1), s));
            if (At->nnz > 0) {
                if (!i2) return fail(MXG_ERR_ARG, "i2 is NULL");
                MXG_TRY(staged_d2h(st, i2, At->d_j, sizeof(int32_t) * (size_t)At->nnz, s));
                if (vals) MXG_TRY(staged_d2h(st, x2, At->d_x64, sizeof(double) * (size_t)At->nnz, s));
            }
            MXG_CUDA_TRY(cudaStreamSynchronize(s));
            return MXG_OK;
        };
        rc = body();
    }
    cudaStreamSynchronize(st->stream);
    free_handle(A);
    free_handle(At);
    return rc;
}

} /* extern "C" */
