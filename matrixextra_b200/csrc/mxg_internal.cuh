// mxg_internal.cuh — shared declarations of libmxgpu.so (not part of the public C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <atomic>
#include <functional>

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
    long spmm_cpl = 0;        // vectors per lane of the SpMM teams: 0 = auto (2 when a row of B exceeds 128 bytes), 1, 2
    long radix_bits = 8;      // CSR->CSC: largest digit of the radix passes (4 .. 10 bits)
    long spmv_lpr = 0;        // 0 = auto
    long svec_smem = 1;       // sparse-vector product: keep the presence bitmap in shared memory when it fits (<= 200 KB)
    long spmv_tex = 1;        // gather a numeric y through the texture path (7 % faster than LDG on cfg2); 0 = plain loads
    long h2d_chunk_mb = 64;   // staging chunk of the value narrowing in mxg_csr_upload
    long pipe_chunk_nnz = 0;  // stored entries (and rows) per chunk of the streamed path; 0 = auto (nnz/16, >= 1 Mi)
    long pipeline = 1;        // level-1 products: 1 = streamed row chunks (pipeline.cu), 0 = upload-all-then-compute
    long host_threads = 0;    // host threads of the staging engine (hoststage.cu); 0 = auto (all logical CPUs, <= 16)
    long host_narrow = 1;     // float32 products: narrow the float64 values on the host (8 instead of 12 PCIe bytes per entry)
    long host_stage = 1;      // bounce pageable caller memory through the page-locked arena with the host threads
    long host_pack = 1;       // streamed calls: column ids cross PCIe as 2 / 2.5 / 3 bytes (K <= 2^16 / 2^20 / 2^24), packed by the host threads
    long host_pack_lag = 2;   // a chunk's ids are packed while the upload of the chunk this many places before it is pending
    long pipe_slots = 4;      // ring slots of the staging arena (chunks in flight between host and device)
    long host_arena_max_mb = 4096; // largest page-locked arena the library may hold; beyond it copies take the driver's path
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

    // partial-sum workspace for long rows (grow-only)
    void *d_partial = nullptr;
    size_t partial_bytes = 0;

    // column-panel split table of the SpMM kernels (spmm.cu: plan_panels), cached per panel width
    int32_t *d_seg = nullptr; // [(seg_panels - 1)][m]
    int seg_panels = 0, seg_width = 0;

    // device flag set by the index validation of the streamed (level-1) path: kernels that see it non-zero
    // return at once instead of gathering through an out-of-range column id
    const int *d_abort = nullptr;
};

namespace mxg {

// per-device state of the host-buffer (level-1) entry points: three streams so that uploads, kernels and
// downloads of consecutive row chunks overlap (PCIe is full duplex)
struct DeviceState {
    bool ready = false;
    cudaStream_t stream = nullptr; // kernels (and everything of the non-pipelined calls)
    cudaStream_t h2d = nullptr;
    cudaStream_t d2h = nullptr;
    // page-locked staging arena of the streamed path (hoststage.cu), grow-only, released by mxg_trim
    void *pin_base = nullptr;
    size_t pin_bytes = 0;
};
int current_state(DeviceState **out);

// pipeline.cu: streamed level-1 products (host CSR with p[0] == 0)
int pipeline_spmm(DeviceState *st, int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p,
                  const int32_t *j, const double *x, const void *B, size_t ldb, void *Out, size_t ldc);
int pipeline_spmv(DeviceState *st, int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x,
                  const void *y, void *out);
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
int ensure_partial(mxg_csr_s *h, size_t bytes);

// spmm.cu
int launch_spmm(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                void *d_Out, size_t ldc, cudaStream_t stream);
// the same product written to n_dst result buffers (local and / or peer-mapped), all with leading dimension ldc
// mcast != 0: d_outs[0] is an NVLS multicast address, rows are written with multimem.st (row-major, one destination)
int launch_spmm_multi(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                      int n_dst, void *const *d_outs, size_t ldc, cudaStream_t stream, int mcast = 0);
// spmv.cu
int launch_spmv(const mxg_csr_s *A, int ytype, const void *d_y, void *d_out, cudaStream_t stream);
int launch_spmv_multi(const mxg_csr_s *A, int ytype, const void *d_y, int n_dst, void *const *d_outs, cudaStream_t stream);
// CSR x sparse vector (indices base 1, K = columns covered by the presence bitmap), double result
int launch_spmv_svec(const mxg_csr_s *A, int ytype, int K, int n_y, const int32_t *d_yidx_base1, const void *d_yvals,
                     double *d_out, cudaStream_t stream);
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
