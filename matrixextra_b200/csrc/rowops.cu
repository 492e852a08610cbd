// rowops.cu — the steps either side of the multiplication path, sm_100a:
//
//   f3  per-row index sorting and CSR validity checks — replaces sort_sparse_indices<T> (src/misc.cpp:192-228,
//       exports 300-330), check_indices_are_unsorted (161-175) and check_valid_csr_matrix (970-1016), which
//       R/utils.R:22-161, 439-489 runs before a product on hand-built inputs;
//   f4  elementwise CSR * dense matrix and CSR * recycled dense vector — replaces
//       multiply_csr_by_dense_elemwise (src/operators.cpp:239-330) and the Multiply case of
//       multiply_csr_by_dvec_no_NAs (1478, 1501-2178).
//
// All of it is integer / one-multiply-per-entry work bound by HBM: coalesced streaming of the CSR arrays, one
// team of lanes per row (one warp per 1024-entry piece for long rows), no atomics on data.
// Sorting: rows that are already non-decreasing are copied (the reference leaves them alone, src/misc.cpp:213);
// an unsorted row is sorted as 64-bit (column id, stored position) keys by a bitonic network — in a warp's
// shared-memory slice up to 512 entries, in a CTA's shared memory up to 16384, and through an L2-resident
// global scratch with shared-memory sub-merges beyond that.  The stored position in the key makes the sort
// stable; for rows with distinct column ids (every valid matrix) the result is the reference's bit for bit.
#include "mxg_internal.cuh"

#include <algorithm>
#include <limits.h>

namespace mxg {

// ================================================================================================
// f4: elementwise products
// ================================================================================================

__device__ __forceinline__ double rowops_na_real() { return __longlong_as_double(0x7FF00000000007A2LL); }

template <int DT>
struct DenseElem;
template <>
struct DenseElem<MXG_Y_NUMERIC> {
    typedef double type;
    static __device__ __forceinline__ double mul(double x, double d) { return x * d; }
};
template <>
struct DenseElem<MXG_Y_FLOAT32> {
    typedef float type;
    static __device__ __forceinline__ double mul(double x, float d) { return x * (double)d; }
};
template <>
struct DenseElem<MXG_Y_INTEGER> {
    typedef int type;
    static __device__ __forceinline__ double mul(double x, int d) { return d == INT_MIN ? rowops_na_real() : x * (double)d; }
};
template <>
struct DenseElem<MXG_Y_LOGICAL> {
    typedef int type;
    static __device__ __forceinline__ double mul(double x, int d) { return d == INT_MIN ? rowops_na_real() : x * (d != 0 ? 1.0 : 0.0); }
};

struct RowWalk {
    int m;
    const int32_t *p;
    const int32_t *j;
    const double *x;
    double *out;
    int piece, n_pieces, piece_blocks;
    const int32_t *piece_row;
    const int32_t *piece_k;
    int rows_per_team;
};

// out[e] = x[e] * dense[row + m * j[e]]  (dense column-major m x K, src/operators.cpp:256-268)
template <int DT>
struct MulDenseOp {
    const typename DenseElem<DT>::type *dense;
    size_t m;
    __device__ __forceinline__ double operator()(int row, int col, double xv) const
    {
        return DenseElem<DT>::mul(xv, __ldg(dense + (size_t)row + m * (size_t)col));
    }
};

// out[e] = x[e] * dvec[recycled position]; MODE 0: position depends on the row only (len == m or len | m),
// 1: len >= m * K (direct), 2: generic (row + col * m) % len  (src/operators.cpp:1478)
template <int MODE>
struct MulDvecOp {
    const double *dvec;
    unsigned long long m, len;
    __device__ __forceinline__ double operator()(int row, int col, double xv) const
    {
        unsigned long long pos;
        if (MODE == 0) pos = (unsigned long long)row % len;
        else {
            pos = (unsigned long long)row + (unsigned long long)col * m;
            if (MODE == 2) pos %= len;
        }
        return xv * __ldg(dvec + pos);
    }
};

template <int LPR, class Op>
__global__ void __launch_bounds__(256) k_rowwise_mul(const RowWalk w, const Op op)
{
    const int32_t *__restrict__ p = w.p;
    const int32_t *__restrict__ j = w.j;
    const double *__restrict__ x = w.x;
    double *__restrict__ out = w.out;
    if ((int)blockIdx.x < w.piece_blocks) {
        const int lane = threadIdx.x & 31;
        const int pc = blockIdx.x * 8 + (threadIdx.x >> 5);
        if (pc >= w.n_pieces) return;
        const int row = w.piece_row[pc];
        const int a = p[row] + w.piece_k[pc] * w.piece;
        const int b = min(a + w.piece, p[row + 1]);
#pragma unroll 4
        for (int e = a + lane; e < b; e += 32) __stcs(out + e, op(row, __ldcs(j + e), __ldcs(x + e)));
        return;
    }
    constexpr int TEAMS = 256 / LPR;
    const int team = threadIdx.x / LPR;
    const int l = threadIdx.x % LPR;
    const int row0 = (blockIdx.x - w.piece_blocks) * TEAMS * w.rows_per_team;
    for (int k = 0; k < w.rows_per_team; k++) {
        const int row = row0 + k * TEAMS + team;
        if (row >= w.m) break;
        const int a = p[row], b = p[row + 1];
        if (b - a > w.piece) continue; // done by the piece warps
#pragma unroll 4
        for (int e = a + l; e < b; e += LPR) __stcs(out + e, op(row, __ldcs(j + e), __ldcs(x + e)));
    }
}

template <class Op>
static int launch_rowwise(const mxg_csr_s *A, double *d_out, const Op &op, cudaStream_t stream)
{
    if (A->m == 0 || A->nnz == 0) return MXG_OK;
    if (!A->d_x64) return fail(MXG_ERR_UNSUPPORTED, "elementwise product: handle holds no float64 values");
    RowWalk w;
    w.m = A->m;
    w.p = A->d_p;
    w.j = A->d_j;
    w.x = A->d_x64;
    w.out = d_out;
    w.piece = A->piece;
    w.n_pieces = A->n_pieces;
    w.piece_blocks = ceil_div_i(A->n_pieces, 8);
    w.piece_row = A->d_piece_row;
    w.piece_k = A->d_piece_k;
    w.rows_per_team = 4;
    const double mean = (double)A->nnz / (double)A->m;
    const int lpr = mean <= 6 ? 4 : mean <= 24 ? 8 : mean <= 96 ? 16 : 32;
#define MXG_ROWWISE(L)                                                                       \
    if (lpr == L) {                                                                          \
        const int grid = w.piece_blocks + ceil_div_i(A->m, (256 / L) * w.rows_per_team);     \
        MXG_LAUNCH((k_rowwise_mul<L, Op>), grid, 256, 0, stream, w, op);                     \
    }
    MXG_ROWWISE(4) MXG_ROWWISE(8) MXG_ROWWISE(16) MXG_ROWWISE(32)
#undef MXG_ROWWISE
    return MXG_OK;
}

int launch_mul_csr_dense(const mxg_csr_s *A, int dtype, const void *d_dense, double *d_out, cudaStream_t stream)
{
    if (A->nnz > 0 && (!d_dense || !d_out)) return fail(MXG_ERR_ARG, "mul_csr_dense: NULL operand");
    switch (dtype) {
    case MXG_Y_NUMERIC: return launch_rowwise(A, d_out, MulDenseOp<MXG_Y_NUMERIC>{static_cast<const double *>(d_dense), (size_t)A->m}, stream);
    case MXG_Y_FLOAT32: return launch_rowwise(A, d_out, MulDenseOp<MXG_Y_FLOAT32>{static_cast<const float *>(d_dense), (size_t)A->m}, stream);
    case MXG_Y_INTEGER: return launch_rowwise(A, d_out, MulDenseOp<MXG_Y_INTEGER>{static_cast<const int *>(d_dense), (size_t)A->m}, stream);
    case MXG_Y_LOGICAL: return launch_rowwise(A, d_out, MulDenseOp<MXG_Y_LOGICAL>{static_cast<const int *>(d_dense), (size_t)A->m}, stream);
    default: return fail(MXG_ERR_ARG, "mul_csr_dense: bad element type %d", dtype);
    }
}

int launch_mul_csr_dvec(const mxg_csr_s *A, const double *d_dvec, size_t len, double *d_out, cudaStream_t stream)
{
    if (A->nnz == 0 || A->m == 0) return MXG_OK;
    if (len == 0) return fail(MXG_ERR_ARG, "mul_csr_dvec: empty vector");
    if (!d_dvec || !d_out) return fail(MXG_ERR_ARG, "mul_csr_dvec: NULL operand");
    const unsigned long long m = (unsigned long long)A->m, K = (unsigned long long)A->K;
    if (len >= m * K) return launch_rowwise(A, d_out, MulDvecOp<1>{d_dvec, m, (unsigned long long)len}, stream);
    if (len == m || (len < m && m % len == 0)) return launch_rowwise(A, d_out, MulDvecOp<0>{d_dvec, m, (unsigned long long)len}, stream);
    return launch_rowwise(A, d_out, MulDvecOp<2>{d_dvec, m, (unsigned long long)len}, stream);
}

// ================================================================================================
// f3: validity checks
// ================================================================================================

// stats[0] = min column id, stats[1] = max column id, stats[2] = indptr holds NA_INTEGER, stats[3] = indptr decreases
__global__ void __launch_bounds__(256) k_check_valid(int m, size_t nnz, const int32_t *__restrict__ p, const int32_t *__restrict__ j,
                                                    int *__restrict__ stats)
{
    int lo = INT_MAX, hi = INT_MIN, pna = 0, pdec = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t e = t0; e < nnz; e += stride) {
        const int c = __ldcs(j + e);
        lo = min(lo, c);
        hi = max(hi, c);
    }
    for (size_t r = t0; r <= (size_t)m; r += stride) {
        const int v = p[r];
        pna |= (v == INT_MIN);
        if (r < (size_t)m) pdec |= (v > p[r + 1]);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    pna = __reduce_or_sync(0xffffffffu, pna);
    pdec = __reduce_or_sync(0xffffffffu, pdec);
    if ((threadIdx.x & 31) == 0) {
        if (lo != INT_MAX) atomicMin(&stats[0], lo);
        if (hi != INT_MIN) atomicMax(&stats[1], hi);
        if (pna) atomicOr(&stats[2], 1);
        if (pdec) atomicOr(&stats[3], 1);
    }
}

int dev_check_valid_csr(int m, int ncols, const int32_t *d_p, const int32_t *d_j, int64_t nnz, int *code, cudaStream_t stream)
{
    int *d_stats = nullptr;
    const int init[4] = {INT_MAX, INT_MIN, 0, 0};
    int h[4];
    MXG_CUDA_TRY(cudaMallocAsync(&d_stats, sizeof(init), stream));
    auto body = [&]() -> int {
        MXG_CUDA_TRY(cudaMemcpyAsync(d_stats, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        const long long work = std::max<long long>(nnz, (long long)m + 1);
        const int grid = (int)std::min<long long>((work + 255) / 256, 148 * 16);
        MXG_LAUNCH(k_check_valid, std::max(grid, 1), 256, 0, stream, m, (size_t)nnz, d_p, d_j, d_stats);
        MXG_CUDA_TRY(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(stream));
        return MXG_OK;
    };
    const int rc = body();
    cudaFreeAsync(d_stats, stream);
    if (rc != MXG_OK) return rc;
    // the reference's order of checks (src/misc.cpp:977-1013); a NA column id is INT_MIN and is caught as negative
    if (nnz > 0 && h[0] < 0) *code = 1;
    else if (nnz > 0 && h[1] >= ncols) *code = 2;
    else if (h[2]) *code = 4;
    else if (h[3]) *code = 5;
    else *code = 0;
    return MXG_OK;
}

// ================================================================================================
// f3: sortedness + per-row sort
// ================================================================================================

// one team per row: flag[0] |= 1 when some row has j[e] < j[e-1]  (check_is_sorted, src/misc.cpp:117-127)
__global__ void __launch_bounds__(256) k_rows_sorted(int m, const int32_t *__restrict__ p, const int32_t *__restrict__ j, int *__restrict__ flag)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    // warps walk contiguous row ranges so that the index stream stays coalesced across short rows
    const int rows_per_warp = (m + n_warps - 1) / n_warps;
    const int r0 = warp * rows_per_warp, r1 = min(m, r0 + rows_per_warp);
    if (r0 >= r1) return;
    const int a = p[r0], b = p[r1];
    int bad = 0;
    // entry e starts a row when p[row] == e: walk the row boundaries alongside the entries
    int row = r0;
    for (int e0 = a; e0 < b; e0 += 32) {
        const int e = e0 + lane;
        if (e < b && e > a) {
            const int cur = __ldcs(j + e), prev = __ldg(j + e - 1);
            if (cur < prev) {
                // a drop is fine only across a row boundary: find whether some row starts at e
                int lo = row, hi = r1; // first row with p[row] >= e
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (p[mid] < e) lo = mid + 1;
                    else hi = mid;
                }
                if (!(lo < r1 && p[lo] == e)) bad = 1;
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 1);
}

__device__ __forceinline__ unsigned long long sort_key(int col, unsigned pos)
{
    return ((unsigned long long)((unsigned)col ^ 0x80000000u) << 32) | pos; // signed order, stored position breaks ties
}
__device__ __forceinline__ int key_col(unsigned long long k) { return (int)((unsigned)(k >> 32) ^ 0x80000000u); }

struct WarpScope {
    int tid, n;
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
struct BlockScope {
    int tid, n;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// steps j = j_start, j_start/2, ..., 1 of bitonic stage k on a local array of n (power of two) keys whose first
// element has index goff in the whole network (the direction of a pair depends on that global index)
template <class Scope>
__device__ __forceinline__ void bitonic_local(unsigned long long *s, int n, unsigned goff, unsigned k, int j_start, const Scope &sc)
{
    for (int jj = j_start; jj > 0; jj >>= 1) {
        for (int i = sc.tid; i < n / 2; i += sc.n) {
            const int lo = 2 * i - (i & (jj - 1));
            const int hi = lo + jj;
            const bool up = ((goff + (unsigned)lo) & k) == 0;
            const unsigned long long a = s[lo], b = s[hi];
            if ((a > b) == up) {
                s[lo] = b;
                s[hi] = a;
            }
        }
        sc.sync();
    }
}

template <class Scope>
__device__ __forceinline__ void bitonic_sort_local(unsigned long long *s, int n, const Scope &sc)
{
    for (unsigned k = 2; k <= (unsigned)n; k <<= 1) bitonic_local(s, n, 0u, k, (int)(k >> 1), sc);
}

__device__ __forceinline__ int next_pow2(int v)
{
    return v <= 1 ? 1 : 1 << (32 - __clz(v - 1));
}

constexpr int SORT_WARP_CAP = 512;   // entries a warp sorts in its shared-memory slice
constexpr int SORT_BLOCK_CAP = 16384; // entries a CTA sorts in shared memory (128 KB)

struct SortArgs {
    int m;
    const int32_t *p;
    const int32_t *j;
    const double *x;
    int32_t *j_out;
    double *x_out;
    int *counters;      // [0] rows queued for the CTA kernel, [1] longest queued row, [2] rows sorted so far
    int32_t *big_rows;  // [m] queue
    unsigned long long *scratch; // CTA kernel: per-CTA global scratch for rows beyond SORT_BLOCK_CAP
    size_t scratch_per_cta;      // in keys
};

// warp per row: copy sorted rows, sort short unsorted rows, queue the long unsorted ones
__global__ void __launch_bounds__(256) k_sort_rows_warp(const SortArgs g)
{
    __shared__ unsigned long long s_keys[8][SORT_WARP_CAP];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * 8 + wib;
    const int n_warps = gridDim.x * 8;
    const bool copy = g.j_out != g.j;
    const WarpScope sc = {lane, 32};
    for (int row = warp; row < g.m; row += n_warps) {
        const int a = g.p[row], b = g.p[row + 1];
        const int len = b - a;
        if (len <= 0) continue;
        int bad = 0;
        for (int e = a + 1 + lane; e < b; e += 32) bad |= (g.j[e] < g.j[e - 1]);
        bad = __any_sync(0xffffffffu, bad);
        if (!bad) {
            if (copy) {
                for (int e = a + lane; e < b; e += 32) {
                    g.j_out[e] = g.j[e];
                    if (g.x) g.x_out[e] = g.x[e];
                }
            }
            continue;
        }
        if (len > SORT_WARP_CAP) {
            if (lane == 0) {
                const int slot = atomicAdd(&g.counters[0], 1);
                g.big_rows[slot] = row;
                atomicMax(&g.counters[1], len);
            }
            continue;
        }
        const int n2 = next_pow2(len);
        unsigned long long *s = s_keys[wib];
        for (int i = lane; i < n2; i += 32) s[i] = i < len ? sort_key(g.j[a + i], (unsigned)i) : ~0ULL;
        __syncwarp();
        bitonic_sort_local(s, n2, sc);
        for (int i = lane; i < len; i += 32) {
            const unsigned long long k = s[i];
            g.j_out[a + i] = key_col(k);
            if (g.x) g.x_out[a + i] = g.x[a + (unsigned)k];
        }
        __syncwarp();
        if (lane == 0) atomicAdd(&g.counters[2], 1);
    }
}

// CTA per queued row
__global__ void __launch_bounds__(1024) k_sort_rows_block(const SortArgs g)
{
    extern __shared__ unsigned long long s_big[];
    const BlockScope sc = {(int)threadIdx.x, (int)blockDim.x};
    const int n_big = g.counters[0];
    for (int q = blockIdx.x; q < n_big; q += gridDim.x) {
        const int row = g.big_rows[q];
        const int a = g.p[row], len = g.p[row + 1] - a;
        const int n2 = next_pow2(len);
        if (n2 <= SORT_BLOCK_CAP) {
            for (int i = sc.tid; i < n2; i += sc.n) s_big[i] = i < len ? sort_key(g.j[a + i], (unsigned)i) : ~0ULL;
            __syncthreads();
            bitonic_sort_local(s_big, n2, sc);
            for (int i = sc.tid; i < len; i += sc.n) {
                const unsigned long long k = s_big[i];
                g.j_out[a + i] = key_col(k);
                if (g.x) g.x_out[a + i] = g.x[a + (unsigned)k];
            }
            __syncthreads();
        } else {
            // network over n2 keys in global scratch; every step whose partner distance fits a shared-memory chunk
            // runs on chunks staged in shared memory
            unsigned long long *gk = g.scratch + (size_t)blockIdx.x * g.scratch_per_cta;
            const int chunks = n2 / SORT_BLOCK_CAP;
            for (int c = 0; c < chunks; c++) {
                const unsigned goff = (unsigned)c * SORT_BLOCK_CAP;
                for (int i = sc.tid; i < SORT_BLOCK_CAP; i += sc.n) {
                    const unsigned gi = goff + (unsigned)i;
                    s_big[i] = gi < (unsigned)len ? sort_key(g.j[a + gi], gi) : ~0ULL;
                }
                __syncthreads();
                for (unsigned k = 2; k <= (unsigned)SORT_BLOCK_CAP; k <<= 1) bitonic_local(s_big, SORT_BLOCK_CAP, goff, k, (int)(k >> 1), sc);
                for (int i = sc.tid; i < SORT_BLOCK_CAP; i += sc.n) gk[goff + i] = s_big[i];
                __syncthreads();
            }
            for (unsigned k = 2u * SORT_BLOCK_CAP; k <= (unsigned)n2; k <<= 1) {
                for (unsigned jj = k >> 1; jj >= (unsigned)SORT_BLOCK_CAP; jj >>= 1) {
                    for (unsigned i = sc.tid; i < (unsigned)n2 / 2; i += sc.n) {
                        const unsigned lo = 2 * i - (i & (jj - 1));
                        const unsigned hi = lo + jj;
                        const bool up = (lo & k) == 0;
                        const unsigned long long va = gk[lo], vb = gk[hi];
                        if ((va > vb) == up) {
                            gk[lo] = vb;
                            gk[hi] = va;
                        }
                    }
                    __syncthreads();
                }
                for (int c = 0; c < chunks; c++) {
                    const unsigned goff = (unsigned)c * SORT_BLOCK_CAP;
                    for (int i = sc.tid; i < SORT_BLOCK_CAP; i += sc.n) s_big[i] = gk[goff + i];
                    __syncthreads();
                    bitonic_local(s_big, SORT_BLOCK_CAP, goff, k, SORT_BLOCK_CAP / 2, sc);
                    for (int i = sc.tid; i < SORT_BLOCK_CAP; i += sc.n) gk[goff + i] = s_big[i];
                    __syncthreads();
                }
            }
            for (int i = sc.tid; i < len; i += sc.n) {
                const unsigned long long k = gk[i];
                g.j_out[a + i] = key_col(k);
                if (g.x) g.x_out[a + i] = g.x[a + (unsigned)k];
            }
            __syncthreads();
        }
        if (sc.tid == 0) atomicAdd(&g.counters[2], 1);
    }
}

int dev_rows_sorted(int m, const int32_t *d_p, const int32_t *d_j, int *sorted, cudaStream_t stream)
{
    *sorted = 1;
    if (m <= 0) return MXG_OK;
    int *d_flag = nullptr;
    int h = 0;
    MXG_CUDA_TRY(cudaMallocAsync(&d_flag, sizeof(int), stream));
    auto body = [&]() -> int {
        MXG_CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
        // ~256 entries per warp at the usual densities; never more warps than rows
        const int grid = std::max(1, std::min(148 * 32, ceil_div_i(m, 8)));
        MXG_LAUNCH(k_rows_sorted, grid, 256, 0, stream, m, d_p, d_j, d_flag);
        MXG_CUDA_TRY(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(stream));
        return MXG_OK;
    };
    const int rc = body();
    cudaFreeAsync(d_flag, stream);
    if (rc == MXG_OK) *sorted = h ? 0 : 1;
    return rc;
}

// d_j_out / d_x_out may equal d_j / d_x only when no row needs sorting is NOT knowable up front: the caller passes
// distinct output arrays (the level-1 entry point does); d_x / d_x_out may be NULL (pattern matrices).
int dev_sort_csr_indices(int m, const int32_t *d_p, const int32_t *d_j, const double *d_x, int32_t *d_j_out, double *d_x_out,
                         int *rows_sorted, cudaStream_t stream)
{
    if (rows_sorted) *rows_sorted = 0;
    if (m <= 0) return MXG_OK;
    if (!d_p || !d_j || !d_j_out) return fail(MXG_ERR_ARG, "sort_csr_indices: NULL array");
    if ((d_x == nullptr) != (d_x_out == nullptr)) return fail(MXG_ERR_ARG, "sort_csr_indices: values in and out must both be given or both be NULL");
    if (d_j_out == d_j || (d_x && d_x_out == d_x)) return fail(MXG_ERR_ARG, "sort_csr_indices: device sort is out of place");
    SortArgs g;
    g.m = m;
    g.p = d_p;
    g.j = d_j;
    g.x = d_x;
    g.j_out = d_j_out;
    g.x_out = d_x_out;
    g.counters = nullptr;
    g.big_rows = nullptr;
    g.scratch = nullptr;
    g.scratch_per_cta = 0;
    int h[3] = {0, 0, 0};
    auto body = [&]() -> int {
        MXG_CUDA_TRY(cudaMallocAsync(&g.counters, 3 * sizeof(int), stream));
        MXG_CUDA_TRY(cudaMallocAsync(&g.big_rows, sizeof(int32_t) * (size_t)m, stream));
        MXG_CUDA_TRY(cudaMemsetAsync(g.counters, 0, 3 * sizeof(int), stream));
        const int grid = std::max(1, std::min(148 * 8, ceil_div_i(m, 8)));
        MXG_LAUNCH(k_sort_rows_warp, grid, 256, 0, stream, g);
        MXG_CUDA_TRY(cudaMemcpyAsync(h, g.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(stream));
        if (h[0] > 0) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int grid2 = std::min(h[0], sms);
            if (h[1] > SORT_BLOCK_CAP) {
                size_t n2 = 1;
                while (n2 < (size_t)h[1]) n2 <<= 1;
                g.scratch_per_cta = n2;
                const size_t budget = (size_t)1 << 30; // 1 GiB of scratch at most
                grid2 = (int)std::max<size_t>(1, std::min<size_t>((size_t)grid2, budget / (n2 * 8)));
                MXG_CUDA_TRY(cudaMallocAsync(&g.scratch, n2 * 8 * (size_t)grid2, stream));
            }
            const size_t smem = (size_t)SORT_BLOCK_CAP * sizeof(unsigned long long);
            MXG_CUDA_TRY(cudaFuncSetAttribute(k_sort_rows_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MXG_LAUNCH(k_sort_rows_block, grid2, 1024, smem, stream, g);
            MXG_CUDA_TRY(cudaMemcpyAsync(h, g.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
            MXG_CUDA_TRY(cudaStreamSynchronize(stream));
        }
        return MXG_OK;
    };
    const int rc = body();
    if (g.counters) cudaFreeAsync(g.counters, stream);
    if (g.big_rows) cudaFreeAsync(g.big_rows, stream);
    if (g.scratch) cudaFreeAsync(g.scratch, stream);
    if (rc == MXG_OK && rows_sorted) *rows_sorted = h[2];
    return rc;
}

} // namespace mxg
