// layout.cu — device-resident CSR layout: validation, value narrowing (K6), row statistics and the
// long-row piece tables (K7), plus the int32 exclusive scan shared by the transpose and the
// synthetic generator.  sm_100a only.
#include "mxg_internal.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace mxg {

std::atomic<unsigned long long> g_launches{0};

std::string &last_error_ref()
{
    static thread_local std::string err;
    return err;
}

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

Options &options()
{
    static Options opt;
    return opt;
}

// ------------------------------------------------------------------------------------------------
// K6: narrow CSR values once at upload; bit-identical to the reference's per-nnz `(float)values[e]`
// (src/matmul.cpp:53-57) because both are a single round-to-nearest-even conversion.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_f64_to_f32(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // two elements per thread per step: 16-byte loads, 8-byte stores
    const size_t n2 = n / 2;
    const double2 *src2 = reinterpret_cast<const double2 *>(src);
    float2 *dst2 = reinterpret_cast<float2 *>(dst);
    for (size_t k = i; k < n2; k += stride) {
        const double2 v = __ldg(src2 + k);
        dst2[k] = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
    }
    if (i == 0 && (n & 1)) dst[n - 1] = __double2float_rn(src[n - 1]);
}

// element-wise variant for destinations that start at an odd element (row chunks of the streamed path)
__global__ void __launch_bounds__(256) k_f64_to_f32_scalar(const double *__restrict__ src, float *__restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) dst[k] = __double2float_rn(__ldg(src + k));
}

int convert_f64_to_f32(const double *d_src, float *d_dst, size_t n, cudaStream_t stream)
{
    if (n == 0) return MXG_OK;
    const bool aligned = (((uintptr_t)d_src & 15) == 0) && (((uintptr_t)d_dst & 7) == 0);
    if (aligned) {
        int grid = ceil_div_i((long long)(n / 2 + 1), 256);
        if (grid > 148 * 16) grid = 148 * 16;
        MXG_LAUNCH(k_f64_to_f32, grid, 256, 0, stream, d_src, d_dst, n);
    } else {
        int grid = ceil_div_i((long long)n, 256 * 4);
        if (grid > 148 * 16) grid = 148 * 16;
        MXG_LAUNCH(k_f64_to_f32_scalar, grid, 256, 0, stream, d_src, d_dst, n);
    }
    return MXG_OK;
}

// ------------------------------------------------------------------------------------------------
// int32 exclusive scan: out[i] = sum_{k<i} in[k], i in [0, n).  Three small kernels (per-block
// sums, scan of the block sums by one block, per-block scan + offset).  Sums must fit in int32
// (they are nnz counts, < 2^31 by construction of R's int indptr).
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int per thread (256 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int *total)
{
    __shared__ int warp_sums[SCAN_THREADS / 32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0;
        const int wi = warp_incl_scan(w, lane);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) block_total = wi;
    }
    __syncthreads();
    const int out = incl - v + warp_sums[warp];
    *total = block_total;
    __syncthreads(); // shared arrays may be reused by the caller's next call
    return out;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const int32_t *__restrict__ in, size_t n,
                                                                  int32_t *__restrict__ block_sums)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    int total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_of_sums(int32_t *__restrict__ block_sums, int nblocks)
{
    // single block: sequential over chunks of 256 with a running carry
    int carry = 0;
    for (int base = 0; base < nblocks; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const int v = (i < nblocks) ? block_sums[i] : 0;
        int total;
        const int ex = block_excl_scan(v, &total);
        if (i < nblocks) block_sums[i] = ex + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int32_t *in, int32_t *out, // may alias (in-place scan)
                                                             size_t n, const int32_t *__restrict__ block_offsets)
{
    // thread t owns SCAN_ITEMS consecutive elements so the scan order is the array order
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int run = block_excl_scan(s, &total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

int exclusive_scan_i32(const int32_t *d_in, int32_t *d_out, size_t n, cudaStream_t stream)
{
    if (n == 0) return MXG_OK;
    const int nblocks = ceil_div_i((long long)n, SCAN_TILE);
    int32_t *d_sums = nullptr;
    MXG_CUDA_TRY(cudaMallocAsync(&d_sums, sizeof(int32_t) * (size_t)nblocks, stream));
    MXG_LAUNCH(k_scan_block_sums, nblocks, SCAN_THREADS, 0, stream, d_in, n, d_sums);
    MXG_LAUNCH(k_scan_of_sums, 1, SCAN_THREADS, 0, stream, d_sums, nblocks);
    MXG_LAUNCH(k_scan_apply, nblocks, SCAN_THREADS, 0, stream, d_in, d_out, n, d_sums);
    MXG_CUDA_TRY(cudaFreeAsync(d_sums, stream));
    return MXG_OK;
}

// ------------------------------------------------------------------------------------------------
// validation + K7 row statistics
//   stats[0] = error flags (1: indptr decreasing/negative, 2: column id out of range)
//   stats[1] = longest row, stats[2] = #rows longer than `piece`, stats[3] = #pieces of those rows
//   stats[4], stats[5] = fill cursors
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_stats(int m, const int32_t *__restrict__ p, int piece, int *__restrict__ stats)
{
    int flags = 0, mx = 0, nl = 0, np = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const int a = p[r], b = p[r + 1];
        if (a < 0 || b < a) flags |= 1;
        const int len = b - a;
        mx = max(mx, len);
        if (len > piece) {
            nl += 1;
            np += (len + piece - 1) / piece;
        }
    }
    // warp-aggregate before touching the global counters
    flags = __reduce_or_sync(0xffffffffu, flags);
    mx = __reduce_max_sync(0xffffffffu, mx);
    nl = __reduce_add_sync(0xffffffffu, nl);
    np = __reduce_add_sync(0xffffffffu, np);
    if ((threadIdx.x & 31) == 0) {
        if (flags) atomicOr(&stats[0], flags);
        atomicMax(&stats[1], mx);
        if (nl) {
            atomicAdd(&stats[2], nl);
            atomicAdd(&stats[3], np);
        }
    }
}

__global__ void __launch_bounds__(256) k_check_indices(size_t nnz, const int32_t *__restrict__ j, int K, int *__restrict__ stats)
{
    int bad = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        const int c = __ldg(j + e);
        bad |= (c < 0 || c >= K);
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(&stats[0], 2);
}

// Streamed path: the flag is a lone device int that later kernels of the same stream test (mxg_csr_s::d_abort).
__global__ void __launch_bounds__(256) k_check_indices_flag(size_t nnz, const int32_t *__restrict__ j, int K, int *__restrict__ flag)
{
    int bad = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        const int c = __ldg(j + e);
        bad |= (c < 0 || c >= K);
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(flag, 1);
}

// Streamed path with packed column ids (hoststage.cu: host_pack_indices): rebuild the int32 ids of a chunk and
// validate them in the same pass (packed ids cannot be negative; the host has already rejected those).
template <int HI_BITS>
__global__ void __launch_bounds__(256) k_unpack_indices(size_t n, const uint16_t *__restrict__ lo, const unsigned char *__restrict__ hi,
                                                        int K, int32_t *__restrict__ out, int *__restrict__ flag)
{
    int bad = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        int c = (int)__ldg(lo + e);
        if (HI_BITS == 8) c |= (int)__ldg(hi + e) << 16;
        if (HI_BITS == 4) c |= (int)((__ldg(hi + (e >> 1)) >> ((e & 1) * 4)) & 15u) << 16;
        out[e] = c;
        bad |= (c >= K);
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(flag, 1);
}

int unpack_indices_flag(size_t n, const void *d_packed, int hi_bits, int K, int32_t *d_j, int *d_flag, cudaStream_t stream)
{
    if (n == 0) return MXG_OK;
    const uint16_t *lo = static_cast<const uint16_t *>(d_packed);
    const unsigned char *hi = static_cast<const unsigned char *>(d_packed) + packed_index_lo_bytes(n);
    int g = ceil_div_i((long long)n, 256 * 8);
    if (g > 148 * 16) g = 148 * 16;
    if (hi_bits == 0) MXG_LAUNCH(k_unpack_indices<0>, g, 256, 0, stream, n, lo, hi, K, d_j, d_flag);
    else if (hi_bits == 4) MXG_LAUNCH(k_unpack_indices<4>, g, 256, 0, stream, n, lo, hi, K, d_j, d_flag);
    else if (hi_bits == 8) MXG_LAUNCH(k_unpack_indices<8>, g, 256, 0, stream, n, lo, hi, K, d_j, d_flag);
    else return fail(MXG_ERR_ARG, "unpack_indices: hi_bits %d", hi_bits);
    return MXG_OK;
}

int check_indices_flag(size_t nnz, const int32_t *d_j, int K, int *d_flag, cudaStream_t stream)
{
    if (nnz == 0) return MXG_OK;
    int g = ceil_div_i((long long)nnz, 256 * 8);
    if (g > 148 * 16) g = 148 * 16;
    MXG_LAUNCH(k_check_indices_flag, g, 256, 0, stream, nnz, d_j, K, d_flag);
    return MXG_OK;
}

// Which slot a long row gets is decided by atomics (arbitrary), but a row's pieces are contiguous and
// ordered, so every result computed from these tables is run-to-run deterministic.
__global__ void __launch_bounds__(256) k_fill_long_tables(int m, const int32_t *__restrict__ p, int piece, int *__restrict__ stats,
                                                          int32_t *__restrict__ long_rows, int32_t *__restrict__ long_first,
                                                          int32_t *__restrict__ long_np, int32_t *__restrict__ piece_row,
                                                          int32_t *__restrict__ piece_k)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const int len = p[r + 1] - p[r];
        if (len > piece) {
            const int np = (len + piece - 1) / piece;
            const int slot = atomicAdd(&stats[4], 1);
            const int first = atomicAdd(&stats[5], np);
            long_rows[slot] = r;
            long_first[slot] = first;
            long_np[slot] = np;
            for (int k = 0; k < np; k++) {
                piece_row[first + k] = r;
                piece_k[first + k] = k;
            }
        }
    }
}

int csr_build_stats(mxg_csr_s *h, int validate, cudaStream_t stream)
{
    h->piece = (int)options().piece;
    if (h->piece < 32) h->piece = 32;
    h->stream = stream;
    h->max_len = h->n_long = h->n_pieces = 0;
    if (h->m == 0) return MXG_OK;
    int *d_stats = nullptr;
    MXG_CUDA_TRY(cudaMallocAsync(&d_stats, sizeof(int) * 8, stream));
    struct Release { // on every path, stream-ordered
        int *q;
        cudaStream_t s;
        ~Release() { cudaFreeAsync(q, s); }
    } release{d_stats, stream};
    MXG_CUDA_TRY(cudaMemsetAsync(d_stats, 0, sizeof(int) * 8, stream));
    int grid = ceil_div_i(h->m, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    MXG_LAUNCH(k_row_stats, grid, 256, 0, stream, h->m, h->d_p, h->piece, d_stats);
    if (validate && h->nnz > 0) {
        int g2 = ceil_div_i(h->nnz, 256 * 8);
        if (g2 > 148 * 16) g2 = 148 * 16;
        MXG_LAUNCH(k_check_indices, g2, 256, 0, stream, (size_t)h->nnz, h->d_j + h->base, h->K, d_stats);
    }
    int stats[8];
    MXG_CUDA_TRY(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    if (stats[0] & 1) return fail(MXG_ERR_INDEX, "CSR indptr is negative or decreasing");
    if (stats[0] & 2) return fail(MXG_ERR_INDEX, "CSR column index outside [0, %d)", h->K);
    h->max_len = stats[1];
    h->n_long = stats[2];
    h->n_pieces = stats[3];
    if (h->n_long > 0) {
        MXG_CUDA_TRY(cudaMallocAsync(&h->d_long_rows, sizeof(int32_t) * (size_t)h->n_long, stream));
        MXG_CUDA_TRY(cudaMallocAsync(&h->d_long_first, sizeof(int32_t) * (size_t)h->n_long, stream));
        MXG_CUDA_TRY(cudaMallocAsync(&h->d_long_np, sizeof(int32_t) * (size_t)h->n_long, stream));
        MXG_CUDA_TRY(cudaMallocAsync(&h->d_piece_row, sizeof(int32_t) * (size_t)h->n_pieces, stream));
        MXG_CUDA_TRY(cudaMallocAsync(&h->d_piece_k, sizeof(int32_t) * (size_t)h->n_pieces, stream));
        MXG_LAUNCH(k_fill_long_tables, grid, 256, 0, stream, h->m, h->d_p, h->piece, d_stats, h->d_long_rows,
                   h->d_long_first, h->d_long_np, h->d_piece_row, h->d_piece_k);
    }
    return MXG_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross-GPU completion barrier of the bcast products.  Launched behind the product kernel on the same
// stream, so the product's peer stores are complete (kernel boundary) before the flag stores are issued;
// one lane per peer publishes `epoch` in the peer's flag array, then polls its own slot of the local array.
// ------------------------------------------------------------------------------------------------
struct PeerFlags {
    int *flags[MXG_MAX_DST];
};
__device__ int g_barrier_failed = 0;

__global__ void __launch_bounds__(32) k_peer_barrier(int rank, int world, const PeerFlags pf, int epoch)
{
    const int g = threadIdx.x;
    if (g >= world || g == rank) return;
    __threadfence_system();
    *reinterpret_cast<volatile int *>(pf.flags[g] + rank) = epoch;
    __threadfence_system();
    const volatile int *mine = reinterpret_cast<volatile int *>(pf.flags[rank] + g);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*mine < epoch) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ULL) { // 10 s: a peer died; do not hang the device
            g_barrier_failed = 1;
            return;
        }
        __nanosleep(200);
    }
}

int launch_peer_barrier(int rank, int world, int *const *peer_flags, int epoch, cudaStream_t stream)
{
    if (world < 1 || world > MXG_MAX_DST || rank < 0 || rank >= world || !peer_flags)
        return fail(MXG_ERR_ARG, "peer_barrier: bad arguments");
    if (world == 1) return MXG_OK;
    PeerFlags pf;
    for (int g = 0; g < MXG_MAX_DST; g++) pf.flags[g] = g < world ? peer_flags[g] : nullptr;
    MXG_LAUNCH(k_peer_barrier, 1, 32, 0, stream, rank, world, pf, epoch);
    return MXG_OK;
}

int peer_barrier_failed(int *failed)
{
    MXG_CUDA_TRY(cudaMemcpyFromSymbol(failed, g_barrier_failed, sizeof(int)));
    return MXG_OK;
}

// Long-row partial sums live in a workspace that belongs to the CALL, not to the handle: stream-ordered allocation on
// the call's stream, released behind the fix-up launch.  Two products on the same handle on different streams never
// share it, and nothing synchronises the device.  The streamed level-1 path brings its own (one per call, sized for its
// largest chunk) in mxg_csr_s::d_partial.
int PartialLease::acquire(const mxg_csr_s *A, size_t bytes, cudaStream_t s)
{
    stream = s;
    if (A->d_partial && A->partial_bytes >= bytes) {
        ptr = A->d_partial;
        owned = false;
        return MXG_OK;
    }
    MXG_CUDA_TRY(cudaMallocAsync(&ptr, bytes > 0 ? bytes : 16, s));
    owned = true;
    return MXG_OK;
}

PartialLease::~PartialLease()
{
    if (owned && ptr) cudaFreeAsync(ptr, stream);
}

} // namespace mxg
