// transpose.cu — K4: bit-exact CSR -> CSC on device, and K5: dense layout change.
//
// K4 replaces the deep conversion MatrixExtra delegates to the Matrix package
// (`as(x, "CsparseMatrix")`, R/conversions.R:390-392): a STABLE partition of the stored entries by
// column, so that inside every column the row ids ascend and duplicates keep their stored order —
// exactly what the CPU counting sort produces.  Integer work, no floating point, bit-exact.
//
//   1. histogram  : count[c] = entries in column c (integer atomics commute => deterministic; accumulated by the
//                   last radix pass, one atomic per run of equal ids in a tile), exclusive scan -> p2[K+1];
//   2. row expand : rowid[e] = r for e in [p[r], p[r+1]);
//   3. stable LSD radix sort of the records (column key, row id, value) by digits of the key, least significant
//      first: ceil(log2(K) / radix_bits) passes of equal width (option "radix_bits", default 8: 19 bits -> 7+7+5).
//      The kernels take digits of up to 10 bits (two passes for K <= 2^20), but measured on B200 a 1024-bin pass
//      costs 1.9 ms against 1.12 ms for a 128/256-bin pass (4-entry runs per digit and tile: twice the store
//      sectors, more shared-memory conflicts), so 2 x 10 bits (4.75 ms on cfg4) loses to 3 x 7 bits (4.48 ms).
//      Per pass:
//        a) per-tile digit histogram (4096-entry tiles), b) exclusive scan over (digit, tile),
//        c) scatter: every warp ranks its 32 consecutive entries with __match_any_sync against
//           per-warp digit counters, so ranks follow entry order (=> each pass is stable); the tile is
//           then re-ordered by digit in shared memory and written out as contiguous runs, i.e. with
//           coalesced stores.  The last pass writes straight into i2 / x2.
// All of it is HBM-bound integer/byte traffic; nothing here belongs on tensor cores.
#include "mxg_internal.cuh"

#include <algorithm>
#include <vector>

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
// a team of LPR lanes per row writes the row id over the row's entries (LPR by the mean row length: a whole warp
// on 20-entry rows leaves a third of the lanes idle and issues one store instruction per row)
template <int LPR>
__global__ void __launch_bounds__(256) k_expand_rows(int m, const int32_t *__restrict__ p, int32_t base, int32_t *__restrict__ rowid)
{
    const int l = threadIdx.x % LPR;
    const int teams = (int)(((size_t)gridDim.x * blockDim.x) / LPR);
    for (int r = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR); r < m; r += teams) {
        const int a = __ldg(p + r) - base, b = __ldg(p + r + 1) - base;
        for (int e = a + l; e < b; e += LPR) __stcs(rowid + e, r);
    }
}

constexpr int RS_THREADS = 512;                   // scatter CTA: 16 warps keep enough loads in flight at 2 CTAs / SM
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_STEPS = 8;                       // 32-entry steps per warp
constexpr int RS_WARP_ITEMS = 32 * RS_STEPS;      // 256 consecutive entries per warp
constexpr int RS_TILE = RS_WARPS * RS_WARP_ITEMS; // 4096 entries per CTA
constexpr int RS_MAX_BITS = 10;                   // widest digit the kernels support (see the header: 7-8 bits is faster)
constexpr int RS_BINS = 1 << RS_MAX_BITS;
constexpr int RS_HIST_THREADS = 256;
constexpr int RS_DPT = RS_BINS / RS_THREADS;      // digits per thread in the tile-local scan (2)
constexpr int RS_MIN_CTAS = 3;                   // register budget: 65536 / (3 * 512) = 42 per thread
typedef unsigned short rs_cnt_t;                  // per-warp digit counters (<= RS_WARP_ITEMS) live in 16 bits

// a) digit histogram of every tile, written digit-major: hist[d * ntiles + tile], d < nbins = 1 << nbits
__global__ void __launch_bounds__(RS_HIST_THREADS) k_radix_hist(size_t n, const int32_t *__restrict__ keys, int shift, int nbins,
                                                                int32_t *__restrict__ hist, int ntiles)
{
    __shared__ int bins[RS_BINS];
    for (int d = threadIdx.x; d < nbins; d += RS_HIST_THREADS) bins[d] = 0;
    __syncthreads();
    const size_t t0 = (size_t)blockIdx.x * RS_TILE;
    constexpr int PER = RS_TILE / RS_HIST_THREADS; // 16 keys per thread: all loads in flight before the first atomic
    int key[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const size_t e = t0 + (size_t)k * RS_HIST_THREADS + threadIdx.x;
        key[k] = e < n ? __ldg(keys + e) : 0;
    }
#pragma unroll
    for (int k = 0; k < PER; k++)
        if (t0 + (size_t)k * RS_HIST_THREADS + threadIdx.x < n) atomicAdd(&bins[(key[k] >> shift) & (nbins - 1)], 1);
    __syncthreads();
    for (int d = threadIdx.x; d < nbins; d += RS_HIST_THREADS) hist[(size_t)d * ntiles + blockIdx.x] = bins[d];
}

struct RadixIO {
    const int32_t *keys_in;
    const int32_t *rows_in;
    const double *x64_in;
    const float *x32_in;
    int32_t *keys_out; // nullptr on the last pass
    int32_t *rows_out;
    double *x64_out;
    float *x32_out;
    int32_t *col_count; // last pass only: entries per column (-> p2); nullptr otherwise
};

// shared memory of a scatter CTA: the tile's records + tables sized by the digit of the pass (128 bins: 69 KB with
// float64 values, three CTAs per SM; 1024 bins: 104 KB, two)
__host__ __device__ constexpr int radix_stride(int nbins) { return nbins < 2 ? 2 : nbins; }
// the tile is re-ordered as whole RECORDS: (key, row, value) in one 16-byte slot (8 bytes without values, 16 + 4 when both
// value types travel), so that an entry costs one shared-memory store and one load instead of three of each
__host__ __device__ constexpr size_t radix_rec_bytes(bool h64, bool h32) { return (h64 || h32) ? 16 : 8; }
constexpr size_t radix_smem_bytes(bool h64, bool h32, int nbins)
{
    return (size_t)RS_TILE * (radix_rec_bytes(h64, h32) + ((h64 && h32) ? 4 : 0)) + sizeof(rs_cnt_t) * RS_WARPS * radix_stride(nbins) +
           sizeof(int) * 2 * radix_stride(nbins);
}

// c) stable scatter.  offs = exclusive scan of hist (digit-major): first destination of (digit, tile).
template <bool H64, bool H32>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_CTAS) k_radix_scatter(size_t n, const RadixIO io, int shift, int nbins,
                                                                 const int32_t *__restrict__ offs, int ntiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int REC = (int)radix_rec_bytes(H64, H32);                                      // bytes per record slot
    constexpr int RW = REC / 4;                                                              // ints per record
    int *s_rec = reinterpret_cast<int *>(smem_raw);                                          // [RS_TILE][RW]: key, row, value bits
    float *s_x32 = reinterpret_cast<float *>(smem_raw + (size_t)RS_TILE * REC);              // [RS_TILE] only if H64 && H32
    const int stride = radix_stride(nbins);                                                  // table rows are nbins wide
    int *dig_off = reinterpret_cast<int *>(smem_raw + (size_t)RS_TILE * (REC + ((H64 && H32) ? 4 : 0))); // [stride] tile-local digit starts
    int *gdelta = dig_off + stride;                                                          // [stride] global - local
    rs_cnt_t *wcnt = reinterpret_cast<rs_cnt_t *>(gdelta + stride);                          // [RS_WARPS][stride]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mask = nbins - 1;
    {
        // counters are 16-bit and handled two at a time as one 32-bit word (digits 2k, 2k+1); only the digits this
        // pass can produce are cleared
        unsigned *z = reinterpret_cast<unsigned *>(wcnt);
        const int words = nbins > 1 ? nbins / 2 : 1;
        for (int i = threadIdx.x; i < RS_WARPS * words; i += RS_THREADS) z[i] = 0u; // words == stride / 2: rows are contiguous
    }
    __syncthreads();

    const size_t t0 = (size_t)blockIdx.x * RS_TILE;
    const size_t w0 = t0 + (size_t)warp * RS_WARP_ITEMS;
    int key[RS_STEPS];
    int rank[RS_STEPS]; // rank of the entry among same-digit entries of this warp, in entry order
    const unsigned lt_mask = (1u << lane) - 1u;
    rs_cnt_t *my_cnt = wcnt + warp * stride;
#pragma unroll
    for (int s = 0; s < RS_STEPS; s++) {
        const size_t e = w0 + (size_t)s * 32 + lane;
        const bool valid = e < n;
        key[s] = valid ? __ldg(io.keys_in + e) : 0;
        if (valid) { // the payload is read after two barriers: have it on its way to L2 meanwhile (no registers held)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(io.rows_in + e));
            if (H64) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.x64_in + e));
            if (H32) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.x32_in + e));
        }
        // invalid lanes get a digit no real lane can have (bit RS_MAX_BITS set) so they never match a real one
        const int d = valid ? ((key[s] >> shift) & mask) : RS_BINS;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        int before = 0;
        if (valid) before = my_cnt[d];
        __syncwarp();
        rank[s] = before + __popc(peers & lt_mask);
        if (valid && (peers & lt_mask) == 0) my_cnt[d] = (rs_cnt_t)(before + __popc(peers)); // lowest peer updates
        __syncwarp();
    }
    __syncthreads();
    // per-warp counts -> per-warp starts inside the digit; digit totals -> tile-local digit starts.
    // Thread t owns the digit pair (2t, 2t+1) = one 32-bit word per warp row: both 16-bit prefix sums advance with
    // a single 32-bit add (a digit's total is at most RS_TILE = 4096, so the low half never carries into the high).
    __shared__ int warp_tot[RS_WARPS];
    static_assert(RS_DPT == 2, "the packed scan handles two digits per thread");
    int run[RS_DPT] = {0, 0};
    {
        unsigned *wc32 = reinterpret_cast<unsigned *>(wcnt);
        const int words = nbins > 1 ? nbins / 2 : 1;
        if ((int)threadIdx.x < words) {
            unsigned acc = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                const unsigned c = wc32[w * (stride / 2) + threadIdx.x];
                wc32[w * (stride / 2) + threadIdx.x] = acc;
                acc += c;
            }
            run[0] = (int)(acc & 0xffffu);
            run[1] = (int)(acc >> 16);
        }
    }
    const int mine = run[0] + run[1];
    int incl = mine; // warp-wide inclusive scan of the per-thread digit totals
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    {
        int excl = incl - mine;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++)
            if (w < warp) excl += warp_tot[w];
#pragma unroll
        for (int q = 0; q < RS_DPT; q++) {
            const int d = threadIdx.x * RS_DPT + q;
            if (d < nbins) {
                dig_off[d] = excl;
                gdelta[d] = offs[(size_t)d * ntiles + blockIdx.x] - excl;
                excl += run[q];
            }
        }
    }
    __syncthreads();
    // re-order the tile by digit in shared memory (payload read coalesced from global)
#pragma unroll
    for (int s = 0; s < RS_STEPS; s++) {
        const size_t e = w0 + (size_t)s * 32 + lane;
        if (e < n) {
            const int d = (key[s] >> shift) & mask;
            const int pos = dig_off[d] + my_cnt[d] + rank[s];
            const int row = __ldg(io.rows_in + e);
            if (H64) {
                const double xv = __ldg(io.x64_in + e);
                *reinterpret_cast<int4 *>(s_rec + (size_t)pos * RW) = make_int4(key[s], row, __double2loint(xv), __double2hiint(xv));
                if (H32) s_x32[pos] = __ldg(io.x32_in + e);
            } else if (H32) {
                *reinterpret_cast<int4 *>(s_rec + (size_t)pos * RW) = make_int4(key[s], row, __float_as_int(__ldg(io.x32_in + e)), 0);
            } else {
                *reinterpret_cast<int2 *>(s_rec + (size_t)pos * RW) = make_int2(key[s], row);
            }
        }
    }
    __syncthreads();
    const int tile_n = (int)((n - t0) < (size_t)RS_TILE ? (n - t0) : (size_t)RS_TILE);
    for (int i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        int k, row;
        if (H64 || H32) {
            const int4 r = *reinterpret_cast<const int4 *>(s_rec + (size_t)i * RW);
            k = r.x;
            row = r.y;
            const int dst = i + gdelta[(k >> shift) & mask];
            if (H64) io.x64_out[dst] = __hiloint2double(r.w, r.z);
            else io.x32_out[dst] = __int_as_float(r.z);
            if (H64 && H32) io.x32_out[dst] = s_x32[i];
        } else {
            const int2 r = *reinterpret_cast<const int2 *>(s_rec + (size_t)i * RW);
            k = r.x;
            row = r.y;
        }
        const int dst = i + gdelta[(k >> shift) & mask];
        if (io.keys_out) io.keys_out[dst] = k;
        io.rows_out[dst] = row;
        // Last pass: the input was sorted by the lower digits and the re-order above is stable, so the tile is now
        // sorted by the whole column id.  The first entry of every run of equal ids adds the run length to the
        // column's count: one global atomic per (tile, column) instead of one per stored entry.
        if (io.col_count != nullptr && (i == 0 || s_rec[(size_t)(i - 1) * RW] != k)) {
            int a = i + 1, b = tile_n; // first position after i whose key differs
            while (a < b) {
                const int mid = (a + b) >> 1;
                if (s_rec[(size_t)mid * RW] == k) a = mid + 1;
                else b = mid;
            }
            atomicAdd(&io.col_count[k], a - i);
        }
    }
}

template <bool H64, bool H32>
static int launch_radix_scatter(size_t n, const RadixIO &io, int shift, int nbins, const int32_t *offs, int ntiles,
                                cudaStream_t stream)
{
    const size_t smem = radix_smem_bytes(H64, H32, nbins);
    static bool configured = false;
    if (!configured) {
        MXG_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<H64, H32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)radix_smem_bytes(H64, H32, RS_BINS)));
        configured = true;
    }
    MXG_LAUNCH((k_radix_scatter<H64, H32>), ntiles, RS_THREADS, smem, stream, n, io, shift, nbins, offs, ntiles);
    return MXG_OK;
}

int csr2csc_device(int m, int K, int64_t nnz, const int32_t *d_p, const int32_t *d_j, const double *d_x64,
                   const float *d_x32, int32_t *d_p2, int32_t *d_i2, double *d_x64o, float *d_x32o,
                   cudaStream_t stream)
{
    MXG_CUDA_TRY(cudaMemsetAsync(d_p2, 0, sizeof(int32_t) * ((size_t)K + 1), stream));
    if (nnz == 0) return MXG_OK;
    if (m <= 0 || K <= 0) return fail(MXG_ERR_ARG, "csr2csc: entries in an empty matrix");

    int32_t base = 0; // arrays passed here start at entry p[0]
    MXG_CUDA_TRY(cudaMemcpyAsync(&base, d_p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    const int32_t *j = d_j + base;
    const size_t n = (size_t)nnz;
    const bool h64 = d_x64 && d_x64o, h32 = d_x32 && d_x32o;
    const double *x64 = h64 ? d_x64 + base : nullptr;
    const float *x32 = h32 ? d_x32 + base : nullptr;

    // stream-ordered temporaries, released on every path (an out-of-memory error half way must not strand the rest)
    struct Temps {
        cudaStream_t s;
        std::vector<void *> q;
        int alloc(void **out, size_t bytes)
        {
            *out = nullptr;
            MXG_CUDA_TRY(cudaMallocAsync(out, bytes > 0 ? bytes : 16, s));
            q.push_back(*out);
            return MXG_OK;
        }
        ~Temps()
        {
            for (void *v : q) cudaFreeAsync(v, s);
        }
    } temps{stream, {}};

    // p2 = exclusive scan of the per-column counts (K+1 slots so the scan output is the full pointer array); the
    // counts are accumulated by the LAST radix pass, where equal column ids sit next to each other in a tile
    int32_t *d_count = nullptr;
    MXG_TRY(temps.alloc((void **)&d_count, sizeof(int32_t) * ((size_t)K + 1)));
    MXG_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int32_t) * ((size_t)K + 1), stream));

    // row ids per entry
    int32_t *d_rowid = nullptr;
    MXG_TRY(temps.alloc((void **)&d_rowid, sizeof(int32_t) * n));
    {
        const double mean = (double)nnz / (double)m;
        // power-law rows: the median is about half the mean, so teams of (mean / 2) lanes rounded down to a power of two
        const int lpr = mean >= 64.0 ? 32 : (mean >= 32.0 ? 16 : (mean >= 16.0 ? 8 : 4));
        int gr = ceil_div_i((long long)m * lpr, 256);
        if (gr > 148 * 32) gr = 148 * 32;
        if (lpr == 32) MXG_LAUNCH(k_expand_rows<32>, gr, 256, 0, stream, m, d_p, base, d_rowid);
        else if (lpr == 16) MXG_LAUNCH(k_expand_rows<16>, gr, 256, 0, stream, m, d_p, base, d_rowid);
        else if (lpr == 8) MXG_LAUNCH(k_expand_rows<8>, gr, 256, 0, stream, m, d_p, base, d_rowid);
        else MXG_LAUNCH(k_expand_rows<4>, gr, 256, 0, stream, m, d_p, base, d_rowid);
    }

    // LSD radix passes over the column key, records = (key, row, values)
    int bits = 0;
    while (bits < 31 && ((int64_t)1 << bits) < (int64_t)K) bits++;
    if (bits < 1) bits = 1;
    // as few passes as the digit width allows, split evenly
    int max_bits = (int)options().radix_bits;
    if (max_bits < 4 || max_bits > RS_MAX_BITS) max_bits = 8;
    const int passes = (bits + max_bits - 1) / max_bits;
    const int digit_bits = (bits + passes - 1) / passes;
    const int ntiles = ceil_div_i(nnz, RS_TILE);
    const size_t hist_n = ((size_t)1 << digit_bits) * (size_t)ntiles;
    int32_t *d_hist = nullptr;
    MXG_TRY(temps.alloc((void **)&d_hist, sizeof(int32_t) * hist_n));
    // ping-pong record buffers (only as many as the pass count needs)
    struct Rec { int32_t *key = nullptr, *row = nullptr; double *x64 = nullptr; float *x32 = nullptr; } buf[2];
    const int nbuf = passes >= 3 ? 2 : (passes == 2 ? 1 : 0);
    for (int b = 0; b < nbuf; b++) {
        MXG_TRY(temps.alloc((void **)&buf[b].key, sizeof(int32_t) * n));
        MXG_TRY(temps.alloc((void **)&buf[b].row, sizeof(int32_t) * n));
        if (h64) MXG_TRY(temps.alloc((void **)&buf[b].x64, sizeof(double) * n));
        if (h32) MXG_TRY(temps.alloc((void **)&buf[b].x32, sizeof(float) * n));
    }

    RadixIO io;
    io.keys_in = j;
    io.rows_in = d_rowid;
    io.x64_in = x64;
    io.x32_in = x32;
    for (int pass = 0; pass < passes; pass++) {
        const int shift = pass * digit_bits;
        const int nbits = std::min(digit_bits, bits - shift);
        const int nbins = 1 << nbits;
        const size_t hist_used = (size_t)nbins * (size_t)ntiles;
        const bool last = pass == passes - 1;
        const Rec &o = buf[pass & 1];
        io.keys_out = last ? nullptr : o.key;
        io.rows_out = last ? d_i2 : o.row;
        io.x64_out = last ? d_x64o : o.x64;
        io.x32_out = last ? d_x32o : o.x32;
        io.col_count = last ? d_count : nullptr;
        MXG_LAUNCH(k_radix_hist, ntiles, RS_HIST_THREADS, 0, stream, n, io.keys_in, shift, nbins, d_hist, ntiles);
        MXG_TRY(exclusive_scan_i32(d_hist, d_hist, hist_used, stream));
        int rc;
        if (h64 && h32) rc = launch_radix_scatter<true, true>(n, io, shift, nbins, d_hist, ntiles, stream);
        else if (h64) rc = launch_radix_scatter<true, false>(n, io, shift, nbins, d_hist, ntiles, stream);
        else if (h32) rc = launch_radix_scatter<false, true>(n, io, shift, nbins, d_hist, ntiles, stream);
        else rc = launch_radix_scatter<false, false>(n, io, shift, nbins, d_hist, ntiles, stream);
        MXG_TRY(rc);
        io.keys_in = io.keys_out;
        io.rows_in = io.rows_out;
        io.x64_in = io.x64_out;
        io.x32_in = io.x32_out;
    }

    MXG_TRY(exclusive_scan_i32(d_count, d_p2, (size_t)K + 1, stream));
    return MXG_OK; // (temporaries released by ~Temps, stream-ordered behind the last kernel)
}

} // namespace mxg
