// transpose.cu — K4: bit-exact CSR -> CSC on device, and K5: dense layout change.
//
// K4 replaces the deep conversion MatrixExtra delegates to the Matrix package
// (`as(x, "CsparseMatrix")`, R/conversions.R:390-392): a STABLE partition of the stored entries by
// column, so that inside every column the row ids ascend and duplicates keep their stored order —
// exactly what the CPU counting sort produces.  Integer work, no floating point, bit-exact.
//
//   1. histogram  : count[c] += 1 for every entry (integer atomics commute => deterministic),
//                   exclusive scan -> p2[K+1];
//   2. row expand : rowid[e] = r for e in [p[r], p[r+1]);
//   3. stable LSD radix sort of (column key, entry id e) by 8-bit digits, least significant first:
//        per pass  a) per-tile digit histogram, b) scan over (digit, tile), c) scatter where each
//        warp ranks its 32 consecutive entries with __match_any_sync — ranks follow entry order, so
//        every pass is stable; ceil(log2(K)/8) passes;
//   4. gather     : i2[q] = rowid[e_q], x2[q] = x[e_q] for the sorted entry ids.
// All of it is HBM-bound integer traffic; nothing here belongs on tensor cores.
#include "mxg_internal.cuh"

namespace mxg {

// ================================ K5: dense transpose ==============================================
template <typename T>
__global__ void __launch_bounds__(256) k_transpose_dense(const T *__restrict__ src, size_t ld_src, T *__restrict__ dst,
                                                         size_t ld_dst, size_t rows, size_t cols)
{
    // src(r, c) at r*ld_src + c  ->  dst(c, r) at c*ld_dst + r ; 32x32 tile, 32x8 threads
    __shared__ T tile[32][33];
    const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const size_t r = r0 + ty + k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + k][tx] = src[r * ld_src + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const size_t c = c0 + ty + k, r = r0 + tx;
        if (r < rows && c < cols) dst[c * ld_dst + r] = tile[tx][ty + k];
    }
}

int launch_transpose_dense(int elem_size, size_t rows, size_t cols, const void *d_src, size_t ld_src,
                           void *d_dst, size_t ld_dst, cudaStream_t stream)
{
    if (rows == 0 || cols == 0) return MXG_OK;
    if (ld_src < cols || ld_dst < rows) return fail(MXG_ERR_ARG, "transpose_dense: leading dimension too small");
    const size_t gx = (cols + 31) / 32, gy = (rows + 31) / 32;
    if (gy > 65535) {
        // grid.y is limited to 65535: walk the rows in slabs
        const size_t slab = (size_t)65535 * 32;
        for (size_t r = 0; r < rows; r += slab) {
            const size_t nr = rows - r < slab ? rows - r : slab;
            const char *s = static_cast<const char *>(d_src) + r * ld_src * (size_t)elem_size;
            char *d = static_cast<char *>(d_dst) + r * (size_t)elem_size;
            MXG_TRY(launch_transpose_dense(elem_size, nr, cols, s, ld_src, d, ld_dst, stream));
        }
        return MXG_OK;
    }
    dim3 grid((unsigned)gx, (unsigned)gy, 1);
    if (elem_size == 4)
        MXG_LAUNCH(k_transpose_dense<float>, grid, 256, 0, stream, static_cast<const float *>(d_src), ld_src,
                   static_cast<float *>(d_dst), ld_dst, rows, cols);
    else if (elem_size == 8)
        MXG_LAUNCH(k_transpose_dense<double>, grid, 256, 0, stream, static_cast<const double *>(d_src), ld_src,
                   static_cast<double *>(d_dst), ld_dst, rows, cols);
    else
        return fail(MXG_ERR_ARG, "transpose_dense: element size %d", elem_size);
    return MXG_OK;
}

// ================================ K4: CSR -> CSC ===================================================
__global__ void __launch_bounds__(256) k_col_histogram(size_t nnz, const int32_t *__restrict__ j, int32_t *__restrict__ count)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) atomicAdd(&count[__ldg(j + e)], 1);
}

// one warp per row writes the row id over the row's entries
__global__ void __launch_bounds__(256) k_expand_rows(int m, const int32_t *__restrict__ p, int32_t base, int32_t *__restrict__ rowid)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < m; r += warps) {
        const int a = p[r] - base, b = p[r + 1] - base;
        for (int e = a + lane; e < b; e += 32) rowid[e] = r;
    }
}

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_STEPS = 16;                     // 32-entry steps per warp
constexpr int RS_WARP_ITEMS = 32 * RS_STEPS;     // 512 consecutive entries per warp
constexpr int RS_TILE = RS_WARPS * RS_WARP_ITEMS; // 4096 entries per CTA
constexpr int RS_BINS = 256;

// a) digit histogram of every tile, written digit-major: hist[d * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(size_t n, const int32_t *__restrict__ keys, int shift,
                                                           int32_t *__restrict__ hist, int ntiles)
{
    __shared__ int bins[RS_BINS];
    bins[threadIdx.x] = 0;
    __syncthreads();
    const size_t t0 = (size_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_TILE / RS_THREADS; k++) {
        const size_t e = t0 + (size_t)k * RS_THREADS + threadIdx.x;
        if (e < n) atomicAdd(&bins[(__ldg(keys + e) >> shift) & (RS_BINS - 1)], 1);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = bins[threadIdx.x];
}

// c) stable scatter.  offs = exclusive scan of hist (digit-major), i.e. the first destination of
//    (digit d, tile t).  src_ids == nullptr means "entry id = position" (first pass).
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(size_t n, const int32_t *__restrict__ keys,
                                                              const int32_t *__restrict__ src_ids, int shift,
                                                              const int32_t *__restrict__ offs, int ntiles,
                                                              int32_t *__restrict__ keys_out, int32_t *__restrict__ ids_out)
{
    __shared__ int wcnt[RS_WARPS][RS_BINS]; // per-warp running digit counts, then per-warp bases
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();

    const size_t w0 = (size_t)blockIdx.x * RS_TILE + (size_t)warp * RS_WARP_ITEMS;
    int key[RS_STEPS];
    int rank[RS_STEPS]; // rank of the entry among same-digit entries of this warp, in entry order
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < RS_STEPS; s++) {
        const size_t e = w0 + (size_t)s * 32 + lane;
        const bool valid = e < n;
        key[s] = valid ? __ldg(keys + e) : 0;
        // invalid lanes get a digit no real lane can have inside the match (bit 8 set)
        const int d = valid ? ((key[s] >> shift) & (RS_BINS - 1)) : RS_BINS;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        int before = 0;
        if (valid) before = wcnt[warp][d];
        __syncwarp();
        rank[s] = before + __popc(peers & lt_mask);
        if (valid && (peers & lt_mask) == 0) wcnt[warp][d] = before + __popc(peers); // lowest peer updates
        __syncwarp();
    }
    __syncthreads();
    // turn per-warp counts into per-warp destination bases: global base of (digit, tile) + counts of lower warps
    {
        const int d = threadIdx.x; // RS_THREADS == RS_BINS
        int run = offs[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const int c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_STEPS; s++) {
        const size_t e = w0 + (size_t)s * 32 + lane;
        if (e < n) {
            const int d = (key[s] >> shift) & (RS_BINS - 1);
            const int dst = wcnt[warp][d] + rank[s];
            if (keys_out) keys_out[dst] = key[s];
            ids_out[dst] = src_ids ? __ldg(src_ids + e) : (int32_t)e;
        }
    }
}

template <bool HAS64, bool HAS32>
__global__ void __launch_bounds__(256) k_csc_gather(size_t nnz, const int32_t *__restrict__ ids, const int32_t *__restrict__ rowid,
                                                    const double *__restrict__ x64, const float *__restrict__ x32,
                                                    int32_t *__restrict__ i2, double *__restrict__ x64o, float *__restrict__ x32o)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
        const int e = __ldg(ids + q);
        i2[q] = __ldg(rowid + e);
        if (HAS64) x64o[q] = __ldg(x64 + e);
        if (HAS32) x32o[q] = __ldg(x32 + e);
    }
}

int csr2csc_device(int m, int K, int64_t nnz, const int32_t *d_p, const int32_t *d_j, const double *d_x64,
                   const float *d_x32, int32_t *d_p2, int32_t *d_i2, double *d_x64o, float *d_x32o,
                   cudaStream_t stream)
{
    // p2: histogram + scan (count has K+1 slots so the scan output is the full pointer array)
    MXG_CUDA_TRY(cudaMemsetAsync(d_p2, 0, sizeof(int32_t) * ((size_t)K + 1), stream));
    if (nnz == 0) return MXG_OK;
    if (m <= 0 || K <= 0) return fail(MXG_ERR_ARG, "csr2csc: entries in an empty matrix");

    int32_t base = 0; // arrays passed here start at entry p[0]
    MXG_CUDA_TRY(cudaMemcpyAsync(&base, d_p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    const int32_t *j = d_j + base;

    int32_t *d_count = nullptr;
    MXG_CUDA_TRY(cudaMallocAsync(&d_count, sizeof(int32_t) * ((size_t)K + 1), stream));
    MXG_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int32_t) * ((size_t)K + 1), stream));
    int g = ceil_div_i(nnz, 256 * 4);
    if (g > 148 * 32) g = 148 * 32;
    MXG_LAUNCH(k_col_histogram, g, 256, 0, stream, (size_t)nnz, j, d_count);
    MXG_TRY(exclusive_scan_i32(d_count, d_p2, (size_t)K + 1, stream));
    MXG_CUDA_TRY(cudaFreeAsync(d_count, stream));

    // row ids per entry
    int32_t *d_rowid = nullptr;
    MXG_CUDA_TRY(cudaMallocAsync(&d_rowid, sizeof(int32_t) * (size_t)nnz, stream));
    int gr = ceil_div_i(m, 8);
    if (gr > 148 * 32) gr = 148 * 32;
    MXG_LAUNCH(k_expand_rows, gr, 256, 0, stream, m, d_p, base, d_rowid);

    // LSD radix passes over the column key
    int bits = 0;
    while (bits < 31 && ((int64_t)1 << bits) < (int64_t)K) bits++;
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    const int ntiles = ceil_div_i(nnz, RS_TILE);
    int32_t *d_hist = nullptr, *d_keyA = nullptr, *d_keyB = nullptr, *d_idA = nullptr, *d_idB = nullptr;
    const size_t hist_n = (size_t)RS_BINS * (size_t)ntiles;
    MXG_CUDA_TRY(cudaMallocAsync(&d_hist, sizeof(int32_t) * hist_n, stream));
    MXG_CUDA_TRY(cudaMallocAsync(&d_idA, sizeof(int32_t) * (size_t)nnz, stream));
    if (passes > 1) {
        MXG_CUDA_TRY(cudaMallocAsync(&d_keyA, sizeof(int32_t) * (size_t)nnz, stream));
        MXG_CUDA_TRY(cudaMallocAsync(&d_idB, sizeof(int32_t) * (size_t)nnz, stream));
    }
    if (passes > 2) MXG_CUDA_TRY(cudaMallocAsync(&d_keyB, sizeof(int32_t) * (size_t)nnz, stream));

    const int32_t *keys_in = j;
    const int32_t *ids_in = nullptr;
    int32_t *key_bufs[2] = {d_keyA, d_keyB};
    int32_t *id_bufs[2] = {d_idA, d_idB};
    for (int pass = 0; pass < passes; pass++) {
        const int shift = pass * 8;
        const bool last = pass == passes - 1;
        int32_t *keys_out = last ? nullptr : key_bufs[pass & 1];
        int32_t *ids_out = id_bufs[pass & 1];
        MXG_LAUNCH(k_radix_hist, ntiles, RS_THREADS, 0, stream, (size_t)nnz, keys_in, shift, d_hist, ntiles);
        MXG_TRY(exclusive_scan_i32(d_hist, d_hist, hist_n, stream));
        MXG_LAUNCH(k_radix_scatter, ntiles, RS_THREADS, 0, stream, (size_t)nnz, keys_in, ids_in, shift, d_hist, ntiles,
                   keys_out, ids_out);
        keys_in = keys_out;
        ids_in = ids_out;
    }

    const int32_t *d_sorted_ids = ids_in;
    int gg = ceil_div_i(nnz, 256 * 4);
    if (gg > 148 * 32) gg = 148 * 32;
    const double *x64 = d_x64 ? d_x64 + base : nullptr;
    const float *x32 = d_x32 ? d_x32 + base : nullptr;
    const bool h64 = x64 && d_x64o, h32 = x32 && d_x32o;
    if (h64 && h32)
        MXG_LAUNCH((k_csc_gather<true, true>), gg, 256, 0, stream, (size_t)nnz, d_sorted_ids, d_rowid, x64, x32, d_i2, d_x64o, d_x32o);
    else if (h64)
        MXG_LAUNCH((k_csc_gather<true, false>), gg, 256, 0, stream, (size_t)nnz, d_sorted_ids, d_rowid, x64, x32, d_i2, d_x64o, d_x32o);
    else if (h32)
        MXG_LAUNCH((k_csc_gather<false, true>), gg, 256, 0, stream, (size_t)nnz, d_sorted_ids, d_rowid, x64, x32, d_i2, d_x64o, d_x32o);
    else
        MXG_LAUNCH((k_csc_gather<false, false>), gg, 256, 0, stream, (size_t)nnz, d_sorted_ids, d_rowid, x64, x32, d_i2, d_x64o, d_x32o);

    MXG_CUDA_TRY(cudaFreeAsync(d_hist, stream));
    MXG_CUDA_TRY(cudaFreeAsync(d_idA, stream));
    if (d_keyA) MXG_CUDA_TRY(cudaFreeAsync(d_keyA, stream));
    if (d_idB) MXG_CUDA_TRY(cudaFreeAsync(d_idB, stream));
    if (d_keyB) MXG_CUDA_TRY(cudaFreeAsync(d_keyB, stream));
    MXG_CUDA_TRY(cudaFreeAsync(d_rowid, stream));
    return MXG_OK;
}

} // namespace mxg
