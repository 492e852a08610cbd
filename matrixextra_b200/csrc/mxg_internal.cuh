// mxg_internal.cuh — shared declarations of libmxgpu.so (not part of the public C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <vector>

#include "../../include/mxgpu.h"

namespace mxg {

// ------------------------------------------------------------------------------------------------
// error plumbing: every public entry point returns a status; the message is kept per thread
// ------------------------------------------------------------------------------------------------
std::string &last_error_ref();
int fail(int code, const char *fmt, ...);

#define MXG_CUDA_TRY(expr)                                                                           \
    do {                                                                                             \
        cudaError_t mxg_e_ = (expr);                                                                 \
        if (mxg_e_ != cudaSuccess)                                                                   \
            return ::mxg::fail(MXG_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,    \
                               cudaGetErrorString(mxg_e_));                                          \
    } while (0)

#define MXG_TRY(expr)                       \
    do {                                    \
        int mxg_rc_ = (expr);               \
        if (mxg_rc_ != MXG_OK) return mxg_rc_; \
    } while (0)

extern std::atomic<unsigned long long> g_launches;

// launch + count + cheap launch-error check (no sync)
#define MXG_LAUNCH(kernel, grid, block, smem, stream, ...)                                           \
    do {                                                                                             \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                  \
        ::mxg::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        MXG_CUDA_TRY(cudaGetLastError());                                                            \
    } while (0)

// ------------------------------------------------------------------------------------------------
// options (mxg_set_option)
// ------------------------------------------------------------------------------------------------
struct Options {
    long piece = 1024;        // nnz per long-row piece; rows longer than this are split
    long spmm_lpr = 0;        // lanes per row of B; 0 = auto
    long spmm_panel_mb = 0;   // column-panel size of the dense operand in MiB; 0 = no panels (default)
    long spmm_panel_cols = 0; // force a panel width in columns of A (tests / sweeps); 0 = by size
    long spmm_rpw = 0;        // consecutive rows per warp (row-major output); 0 = auto
    long spmm_bulk = 1;       // multi-destination rows-contiguous products ship finished rows as bulk copies from shared memory (TMA)
    long spmm_cpl = 0;        // vectors per lane of the SpMM teams: 0 = auto (2 when a row of B exceeds 128 bytes), 1, 2
    long radix_bits = 8;      // CSR->CSC: largest digit of the radix passes (4 .. 10 bits)
    long spmv_lpr = 0;        // 0 = auto
    long svec_smem = 1;       // sparse-vector product: keep the presence bitmap in shared memory when it fits (<= 200 KB)
    long spmv_tex = 1;        // gather a numeric y through the texture path (7 % faster than LDG on cfg2); 0 = plain loads
    long h2d_chunk_mb = 64;   // staging chunk of the value narrowing in mxg_csr_upload
    long pipe_chunk_nnz = 0;  // stored entries (and rows) per chunk of the streamed path; 0 = auto (nnz/16, >= 1 Mi)
    long pipeline = 1;        // level-1 products: 1 = streamed row chunks (pipeline.cu), 0 = upload-all-then-compute
    long host_threads = 0;    // host threads of the staging engine (hoststage.cu); 0 = auto (all logical CPUs, <= 16)
    long host_narrow = 1;     // float32 products: narrow the float64 values on the host (8 instead of 12 PCIe bytes per entry); 1 = automatic, 2 = always
    long host_stage = 1;      // bounce pageable caller memory through the page-locked arena with the host threads
    long host_pack = 1;       // streamed calls: column ids cross PCIe as 2 / 2.5 / 3 bytes (K <= 2^16 / 2^20 / 2^24), packed by the host threads
    long host_pack_lag = 2;   // a chunk's ids are packed while the upload of the chunk this many places before it is pending
    long pipe_slots = 4;      // ring slots of the staging arena (chunks in flight between host and device)
    long host_arena_max_mb = 4096; // largest page-locked arena the library may hold; beyond it copies take the driver's path
    long multi_min_nnz = 4 << 20;  // mxg_set_devices(n > 1): level-1 calls with fewer stored entries stay on one device
    long multi_pageable = 0;       // ... and calls whose CSR arrays are pageable stay on one device too (they are bound by the host threads)
    long multi_dense_share = 1;    // ... the dense operand crosses PCIe once (a slice per device) and is completed over NVLink
    long host_result_pool_mb = 4096; // page-locked result memory the glue's allocator hook may hold (mxg_host_alloc)
    long host_pin_register = 1;    // new page-locked blocks (arena, result pool): huge-page backed anonymous memory + cudaHostRegister (5 - 10 x faster
                                   // to create than cudaHostAlloc); 0 = cudaHostAlloc
    long host_thp = 1;             // ask for transparent huge pages on large pageable result buffers before their first touch
    long host_colsplit = 1;        // warm products (device-resident CSR, host operands): the dense operand and the result cross PCIe as two column
                                   // halves so that uploads and downloads overlap; 1 = page-locked operand and unchanged summation order only,
                                   // 2 = wherever the halves are >= 128 bytes wide, 0 = off
    long cache_mb = 0;             // level-1 operand cache (device-resident CSR + dense operands keyed on the host arrays); 0 = off
};
Options &options();

// ------------------------------------------------------------------------------------------------
// device-resident CSR
// ------------------------------------------------------------------------------------------------
} // namespace mxg

struct mxg_csr_s {
    int device = 0;
    int m = 0, K = 0;
    int64_t nnz = 0;
    int32_t base = 0; // p[0] (0 for every valid R matrix; kept so offsets stay exact otherwise)
    const int32_t *d_p = nullptr;
    const int32_t *d_j = nullptr;
    const double *d_x64 = nullptr;
    const float *d_x32 = nullptr;
    bool owns = false; // arrays were allocated by the library
    cudaStream_t stream = nullptr; // every allocation of this handle is stream-ordered on this stream (async pool)

    // row statistics (K7)
    int piece = 1024;
    int max_len = 0;
    int n_long = 0;
    int n_pieces = 0;
    int32_t *d_long_rows = nullptr;  // [n_long] row id
    int32_t *d_long_first = nullptr; // [n_long] first piece slot
    int32_t *d_long_np = nullptr;    // [n_long] number of pieces
    int32_t *d_piece_row = nullptr;  // [n_pieces]
    int32_t *d_piece_k = nullptr;    // [n_pieces] piece number inside its row

    // partial-sum workspace for long rows: only set on the chunk views of the streamed level-1 path (one workspace per
    // call); products on real handles lease theirs per call (PartialLease)
    void *d_partial = nullptr;
    size_t partial_bytes = 0;

    // column-panel split table of the SpMM kernels (spmm.cu: plan_panels), cached per panel width
    int32_t *d_seg = nullptr; // [(seg_panels - 1)][m]
    int seg_panels = 0, seg_width = 0;

    // device flag set by the index validation of the streamed (level-1) path: kernels that see it non-zero
    // return at once instead of gathering through an out-of-range column id
    const int *d_abort = nullptr;

    // host-side extras of handles that serve host-buffer products (pipeline.cu: handle_spmm_host): the row chunks
    // the result leaves the device in, and the cached CSC (as the CSR handle of t(A)) for crossprod-type products
    std::vector<int32_t> *host_chunks = nullptr;
    mxg_csr_s *cached_t = nullptr;
};

namespace mxg {

// per-device state of the host-buffer (level-1) entry points: three streams so that uploads, kernels and
// downloads of consecutive row chunks overlap (PCIe is full duplex)
// a page-locked host block: huge-page backed anonymous memory registered with the driver, or a cudaHostAlloc block
struct PinnedBlock {
    void *ptr = nullptr;
    size_t bytes = 0;
    void *map_base = nullptr; // non-NULL: mmap'ed + cudaHostRegister'ed (ptr is the 2 MiB-aligned interior)
    size_t map_len = 0;
};
int pinned_block_alloc(size_t bytes, PinnedBlock *blk);
void pinned_block_free(PinnedBlock *blk);

struct DeviceState {
    bool ready = false;
    cudaStream_t stream = nullptr; // kernels (and everything of the non-pipelined calls)
    cudaStream_t h2d = nullptr;
    cudaStream_t d2h = nullptr;
    cudaStream_t p2p = nullptr; // pulls of the other devices' dense-operand slices over NVLink (multi-device calls)
    std::vector<cudaEvent_t> ev_pool; // timing-disabled events reused by mxg_dev_spmm_push
    std::vector<cudaStream_t> push_streams; // one per destination of mxg_dev_spmm_push
    // this device's copy of the dense operand of a multi-device call (cudaMalloc: visible to the peers that pull slices
    // out of it), grow-only, released by mxg_trim
    void *share_buf = nullptr;
    size_t share_bytes = 0;
    // page-locked staging arena of the streamed path (hoststage.cu), grow-only, released by mxg_trim
    PinnedBlock pin; // the page-locked staging arena (grow-only; hoststage.cu)
};
int current_state(DeviceState **out);

// Multi-device level-1 products (mxg_set_devices(n > 1); capi.cu: one host thread per device, each running the
// streamed pipeline on its nnz-balanced row block).  The dense operand is replicated: every device uploads ONE
// slice of its rows over its own PCIe link and pulls the other slices from its peers over NVLink.
struct DenseShare {
    int G = 1;
    int device[MXG_MAX_DST] = {};
    void *d_B[MXG_MAX_DST] = {};
    cudaEvent_t slice_ready[MXG_MAX_DST] = {};
    // host barrier between the device threads; a thread that has returned counts as arrived at every phase
    std::mutex mu;
    std::condition_variable cv;
    int phase_of[MXG_MAX_DST] = {}; // barriers rank g has entered
    bool gone[MXG_MAX_DST] = {};
    bool failed = false;
    bool wait(int phase, int g) // false: a participant has failed, the slices are not all there
    {
        std::unique_lock<std::mutex> lk(mu);
        phase_of[g] = phase + 1;
        cv.notify_all();
        cv.wait(lk, [&] {
            for (int q = 0; q < G; q++)
                if (!gone[q] && phase_of[q] <= phase) return false;
            return true;
        });
        return !failed;
    }
    void leave(int g, bool ok)
    {
        std::lock_guard<std::mutex> lk(mu);
        gone[g] = true;
        if (!ok) failed = true;
        cv.notify_all();
    }
};

// pipeline.cu: streamed level-1 products (host CSR with p[0] == 0)
// share / share_rank: this call is one device of a multi-device product (NULL otherwise)
// keep: when non-NULL and the call succeeds, the device CSR is not released but returned as an owned handle
int pipeline_spmm(DeviceState *st, int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p,
                  const int32_t *j, const double *x, const void *B, size_t ldb, void *Out, size_t ldc,
                  DenseShare *share = nullptr, int share_rank = 0, mxg_csr_s **keep = nullptr);
// host dense operand in, host result out, CSR already device-resident (the warm path of SURVEY.md 8 f1)
int handle_spmm_host(DeviceState *st, mxg_csr_s *A, int dtype, int out_layout, int b_layout, int n, const void *B, size_t ldb,
                     void *Out, size_t ldc, const void *d_B_resident = nullptr, void **d_B_keep = nullptr);
int handle_spmv_host(DeviceState *st, mxg_csr_s *A, int ytype, const void *y, void *out);
int handle_chunks(mxg_csr_s *A, cudaStream_t stream); // fills A->host_chunks (row chunks of about equal nnz) once
void set_last_call_bytes(size_t h2d, size_t d2h);

// capi.cu
int csr_handle_free(mxg_csr_s *h);

// residency.cu: devices of a level-1 call, level-1 operand cache
int multi_devices();
int set_devices(int n);
int row_partition(int m, const int32_t *p, int parts, int32_t *row_starts);
bool multi_wanted(int m, const int32_t *p, const int32_t *j, const double *x);
bool multi_wanted_now(); // the calling thread is one device pipeline of a multi-device call
int multi_spmm(int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p, const int32_t *j,
               const double *x, const void *B, size_t ldb, void *Out, size_t ldc);
int multi_spmv(int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x, const void *y, void *out);
bool cache_enabled();
mxg_csr_s *cache_find_csr(int m, int K, const int32_t *p, const int32_t *j, const double *x, int need);
void cache_insert_csr(int m, int K, const int32_t *p, const int32_t *j, const double *x, mxg_csr_s *h);
void cache_account_csr(mxg_csr_s *h);
void *cache_find_dense(const void *ptr, int dtype, int layout, size_t K, size_t n, size_t ldb, size_t host_bytes);
void cache_insert_dense(const void *ptr, int dtype, int layout, size_t K, size_t n, size_t ldb, size_t host_bytes, void *d_B,
                        size_t dev_bytes, cudaStream_t stream);
int cache_clear();
void cache_stats(unsigned long long *hits, unsigned long long *misses, size_t *bytes, int *entries);
int pipeline_spmv(DeviceState *st, int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x,
                  const void *y, void *out, mxg_csr_s **keep = nullptr);
void last_call_bytes(size_t *h2d, size_t *d2h);
int host_chunk_plan(int m, const int32_t *p, size_t result_row_bytes, int32_t *chunk_rows, int cap, int *n_chunks, int *n_long,
                    int *n_pieces, int *max_len);
int check_indices_flag(size_t nnz, const int32_t *d_j, int K, int *d_flag, cudaStream_t stream);
int unpack_indices_flag(size_t n, const void *d_packed, int hi_bits, int K, int32_t *d_j, int *d_flag, cudaStream_t stream);

// hoststage.cu: worker threads + page-locked arena for pageable caller memory and host-side narrowing
int host_threads();
// fn(0 .. ntasks-1) on the calling thread plus pool workers (one per MiB of bytes_touched, up to host_threads())
void host_parallel_for(size_t ntasks, size_t bytes_touched, const std::function<void(size_t)> &fn);
void host_narrow_f64_to_f32(const double *src, float *dst, size_t n);
// column ids as [uint16 low halves][high nibbles / bytes] (2, 2.5 or 3 bytes per entry on the wire)
int index_pack_hi_bits(int K);                      // 0, 4, 8, or -1 when K > 2^24 (ids travel as int32)
size_t packed_index_lo_bytes(size_t n);             // offset of the high parts
size_t packed_index_bytes(size_t n, int hi_bits);
bool host_pack_indices(const int32_t *j, size_t n, int K, int hi_bits, void *dst); // false: an id outside [0, K)
// nt_dst: the destination is a ring slot the DMA engine reads next -> cache-bypassing stores
void host_copy(void *dst, const void *src, size_t bytes, bool nt_dst = false);
void host_copy_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, bool nt_dst = false);
bool host_is_pinned(const void *ptr);
void host_prepare_result(void *ptr, size_t bytes);
// page-locked result blocks handed to the glue's allocator hook (recycled; hoststage.cu)
int result_pool_alloc(size_t bytes, void **out);
int result_pool_free(void *ptr);
void result_pool_stats(size_t *live_bytes, size_t *free_bytes, int *blocks);
void result_pool_trim(); // madvise(MADV_HUGEPAGE) on a fresh pageable result
int pinned_arena(DeviceState *st, size_t bytes, char **base);
int pinned_arena_release(DeviceState *st);
// one-shot staged copies (non-streamed entry points); every slot is idle again when they return
int staged_h2d(DeviceState *st, void *d_dst, const void *src, size_t bytes, cudaStream_t stream);
int staged_h2d_narrow(DeviceState *st, float *d_dst, const double *src, size_t n, cudaStream_t stream);
int staged_d2h(DeviceState *st, void *dst, const void *d_src, size_t bytes, cudaStream_t stream);

// layout.cu: device-side completion barrier between the GPUs of a box (bcast products)
int launch_peer_barrier(int rank, int world, int *const *peer_flags, int epoch, cudaStream_t stream);
int peer_barrier_failed(int *failed);

// layout.cu
int csr_build_stats(mxg_csr_s *h, int validate, cudaStream_t stream);
int convert_f64_to_f32(const double *d_src, float *d_dst, size_t n, cudaStream_t stream);
int exclusive_scan_i32(const int32_t *d_in, int32_t *d_out, size_t n, cudaStream_t stream); // out[i] = sum in[0..i)
// one stream-ordered device buffer that is handed back on every exit path unless release()d to the caller
struct StreamBuf {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    StreamBuf() = default;
    StreamBuf(const StreamBuf &) = delete;
    StreamBuf &operator=(const StreamBuf &) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t s)
    {
        stream = s;
        return cudaMallocAsync(&ptr, bytes > 16 ? bytes : 16, s);
    }
    void *release()
    {
        void *q = ptr;
        ptr = nullptr;
        return q;
    }
    ~StreamBuf()
    {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
};

// workspace of the long-row partial sums for ONE product: leased on the call's stream, released when the lease dies
// (stream-ordered, i.e. behind the fix-up launch that reads it)
struct PartialLease {
    void *ptr = nullptr;
    bool owned = false;
    cudaStream_t stream = nullptr;
    int acquire(const mxg_csr_s *A, size_t bytes, cudaStream_t s);
    ~PartialLease();
};

// spmm.cu
int launch_spmm(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                void *d_Out, size_t ldc, cudaStream_t stream);
// the same product written to n_dst result buffers (local and / or peer-mapped), all with leading dimension ldc
// mcast != 0: d_outs[0] is an NVLS multicast address, rows are written with multimem.st (row-major, one destination)
int launch_spmm_multi(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                      int n_dst, void *const *d_outs, size_t ldc, cudaStream_t stream, int mcast = 0);
// one slice of the product: rows [r0, r1) of a handle without their long rows (pieces == 0), or only the long rows
// (pieces != 0: piece kernels + fix-up, whatever r0 / r1).  d_Out is the FULL result's origin in both cases.
int launch_spmm_rows(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb, void *d_Out,
                     size_t ldc, int r0, int r1, int pieces, cudaStream_t stream);
// spmv.cu
int launch_spmv(const mxg_csr_s *A, int ytype, const void *d_y, void *d_out, cudaStream_t stream);
int launch_spmv_multi(const mxg_csr_s *A, int ytype, const void *d_y, int n_dst, void *const *d_outs, cudaStream_t stream);
int spmv_probe(const mxg_csr_s *A, int mode, const double *d_y, double *d_sink, cudaStream_t stream);
void texture_cache_clear(); // cached texture objects over dense vectors (mxg_trim)
// CSR x sparse vector (indices base 1, K = columns covered by the presence bitmap), double result
int launch_spmv_svec(const mxg_csr_s *A, int ytype, int K, int n_y, const int32_t *d_yidx_base1, const void *d_yvals,
                     double *d_out, cudaStream_t stream);
// tmaprobe.cu: random-row gathers through the TMA unit (tile::gather4), measurement only
int tma_gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed, float *d_sink,
                     long long *gathers_done, cudaStream_t stream);
// rowops.cu
int launch_mul_csr_dense(const mxg_csr_s *A, int dtype, const void *d_dense, double *d_out, cudaStream_t stream);
int launch_mul_csr_dvec(const mxg_csr_s *A, const double *d_dvec, size_t len, double *d_out, cudaStream_t stream);
int dev_check_valid_csr(int m, int ncols, const int32_t *d_p, const int32_t *d_j, int64_t nnz, int *code, cudaStream_t stream);
int dev_rows_sorted(int m, const int32_t *d_p, const int32_t *d_j, int *sorted, cudaStream_t stream);
int dev_sort_csr_indices(int m, const int32_t *d_p, const int32_t *d_j, const double *d_x, int32_t *d_j_out, double *d_x_out,
                         int *rows_sorted, cudaStream_t stream);
// transpose.cu
int launch_transpose_dense(int elem_size, size_t rows, size_t cols, const void *d_src, size_t ld_src,
                           void *d_dst, size_t ld_dst, cudaStream_t stream);
int csr2csc_device(int m, int K, int64_t nnz, const int32_t *d_p, const int32_t *d_j,
                   const double *d_x64, const float *d_x32,
                   int32_t *d_p2, int32_t *d_i2, double *d_x64o, float *d_x32o, cudaStream_t stream);

inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

} // namespace mxg
