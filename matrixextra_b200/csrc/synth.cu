// synth.cu — deterministic synthetic CSR inputs generated directly in device memory (bench.py and
// tests only; nothing on the multiply path depends on this file).  SURVEY.md §8(d).
//
//   row lengths  row_model 0: uniform  — target/m per row with +-20% jitter that cancels in pairs,
//                              so nnz == target exactly (BASELINE cfg1: 10k x 5k, 500 000 entries);
//                row_model 1: power law — len_r = min(cap, floor(L * (1-u_r)^(-1/alpha))), alpha = 1.5,
//                              cap = min(K, 65536), L found by bisection so that sum(len) == target
//                              (remainder spread one entry per row); rows are NOT sorted by length.
//   columns      col_model 0: stratified uniform — entry t of a row of length len is drawn inside
//                              [floor(t*K/len), floor((t+1)*K/len)): sorted and unique by construction;
//                col_model 1: recommender-style popularity — strata [b_t, b_{t+1}) with
//                              b_t = t + floor((K-len) * (t/len)^2): still sorted and unique, but the
//                              density of column c falls like c^(-1/2) (a few hot columns).
//   values       uniform in [-1, 1) (float64; the float32 copy is the same narrowing as K6).
// RNG: Philox4x32-10, counter = (row, entry, purpose, 0), key = seed.
#include "mxg_internal.cuh"

#include <math.h>

namespace mxg {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo)
{
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11; // 53 bits
    return (double)v * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ uint4 rng(uint64_t seed, uint32_t row, uint32_t t, uint32_t purpose)
{
    return philox4x32_10(make_uint4(row, t, purpose, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__global__ void __launch_bounds__(256) k_row_weights(int m, uint64_t seed, double inv_alpha, double *__restrict__ w)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const uint4 z = rng(seed, (uint32_t)r, 0u, 1u);
        const double u = u01(z.x, z.y);
        w[r] = pow(1.0 - u, -inv_alpha); // >= 1
    }
}

__global__ void __launch_bounds__(256) k_sum_lengths(int m, const double *__restrict__ w, double L, int cap,
                                                     unsigned long long *__restrict__ total)
{
    unsigned long long s = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const double v = floor(L * w[r]);
        s += (unsigned long long)(v < (double)cap ? v : (double)cap);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

__global__ void __launch_bounds__(256) k_lengths_powerlaw(int m, const double *__restrict__ w, double L, int cap,
                                                          long long remainder, int32_t *__restrict__ len)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const double v = floor(L * w[r]);
        int l = (int)(v < (double)cap ? v : (double)cap);
        if ((long long)r < remainder && l < cap) l += 1;
        len[r] = l;
    }
}

__global__ void __launch_bounds__(256) k_lengths_uniform(int m, int K, long long target, uint64_t seed, int32_t *__restrict__ len)
{
    const long long base = target / m;
    const long long rem = target - base * m;
    long long jit = base / 5;
    if (jit > (long long)K - base - 1) jit = (long long)K - base - 1;
    if (jit < 0) jit = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const int pair = r >> 1;
        long long d = 0;
        if ((pair * 2 + 1) < m && jit > 0) {
            const uint4 z = rng(seed, (uint32_t)pair, 0u, 2u);
            d = (long long)(z.x % (uint32_t)(jit + 1));
        }
        long long l = base + ((r & 1) ? -d : d);
        if ((long long)r < rem) l += 1;
        if (l > K) l = K;
        len[r] = (int)l;
    }
}

template <bool HAS64, bool HAS32>
__global__ void __launch_bounds__(256) k_fill_entries(int m, int K, const int32_t *__restrict__ p, int col_model, uint64_t seed,
                                                      int32_t *__restrict__ j, double *__restrict__ x64, float *__restrict__ x32)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < m; r += warps) {
        const int a = p[r], len = p[r + 1] - a;
        for (int t = lane; t < len; t += 32) {
            long long lo, hi;
            if (col_model == 0) {
                lo = ((long long)t * K) / len;
                hi = ((long long)(t + 1) * K) / len;
            } else {
                const double f0 = (double)t / (double)len, f1 = (double)(t + 1) / (double)len;
                lo = t + (long long)floor((double)(K - len) * f0 * f0);
                hi = (t + 1 == len) ? (long long)K : (t + 1) + (long long)floor((double)(K - len) * f1 * f1);
            }
            const uint4 z = rng(seed, (uint32_t)r, (uint32_t)t, 3u);
            long long c = lo + (long long)(u01(z.x, z.y) * (double)(hi - lo));
            if (c >= hi) c = hi - 1;
            if (c < lo) c = lo;
            j[a + t] = (int32_t)c;
            const double v = 2.0 * u01(z.z, z.w) - 1.0;
            if (HAS64) x64[a + t] = v;
            if (HAS32) x32[a + t] = (float)v;
        }
    }
}

static int sum_lengths(int m, const double *d_w, double L, int cap, unsigned long long *d_total, long long *out,
                       cudaStream_t stream)
{
    MXG_CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream));
    int grid = ceil_div_i(m, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    MXG_LAUNCH(k_sum_lengths, grid, 256, 0, stream, m, d_w, L, cap, d_total);
    unsigned long long h = 0;
    MXG_CUDA_TRY(cudaMemcpyAsync(&h, d_total, sizeof(h), cudaMemcpyDeviceToHost, stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    *out = (long long)h;
    return MXG_OK;
}

int synth_csr_arrays(int m, int K, int64_t target_nnz, int row_model, int col_model, uint64_t seed, int keep,
                     cudaStream_t stream, int32_t **out_p, int32_t **out_j, double **out_x64, float **out_x32,
                     int64_t *out_nnz)
{
    if (m <= 0 || K <= 0 || target_nnz < 0) return fail(MXG_ERR_ARG, "synth: bad shape");
    if (target_nnz > 2147483647LL) return fail(MXG_ERR_ARG, "synth: nnz must fit R's int32 indptr");
    if (target_nnz > (int64_t)m * (int64_t)K) return fail(MXG_ERR_ARG, "synth: more entries than cells");
    int grid = ceil_div_i(m, 256);
    if (grid > 148 * 8) grid = 148 * 8;

    int32_t *d_len = nullptr, *d_p = nullptr;
    MXG_CUDA_TRY(cudaMallocAsync(&d_p, sizeof(int32_t) * ((size_t)m + 1), stream));
    MXG_CUDA_TRY(cudaMallocAsync(&d_len, sizeof(int32_t) * ((size_t)m + 1), stream));
    MXG_CUDA_TRY(cudaMemsetAsync(d_len, 0, sizeof(int32_t) * ((size_t)m + 1), stream));

    if (row_model == 0) {
        MXG_LAUNCH(k_lengths_uniform, grid, 256, 0, stream, m, K, (long long)target_nnz, seed, d_len);
    } else if (row_model == 1) {
        const int cap = K < 65536 ? K : 65536;
        if ((int64_t)m * cap < target_nnz) return fail(MXG_ERR_ARG, "synth: target nnz unreachable with the row cap");
        double *d_w = nullptr;
        unsigned long long *d_total = nullptr;
        MXG_CUDA_TRY(cudaMallocAsync(&d_w, sizeof(double) * (size_t)m, stream));
        MXG_CUDA_TRY(cudaMallocAsync(&d_total, sizeof(unsigned long long), stream));
        MXG_LAUNCH(k_row_weights, grid, 256, 0, stream, m, seed, 1.0 / 1.5, d_w);
        // bisection on the scale L: sum_lengths is monotone in L
        double lo = 0.0, hi = (double)target_nnz / (double)m + 1.0;
        long long s = 0;
        for (int it = 0; it < 64; it++) {
            MXG_TRY(sum_lengths(m, d_w, hi, cap, d_total, &s, stream));
            if (s >= target_nnz) break;
            hi *= 2.0;
        }
        for (int it = 0; it < 60; it++) {
            const double mid = 0.5 * (lo + hi);
            MXG_TRY(sum_lengths(m, d_w, mid, cap, d_total, &s, stream));
            if (s <= target_nnz) lo = mid;
            else hi = mid;
        }
        MXG_TRY(sum_lengths(m, d_w, lo, cap, d_total, &s, stream));
        long long remainder = target_nnz - s;
        if (remainder < 0) remainder = 0;
        if (remainder > m) remainder = m;
        MXG_LAUNCH(k_lengths_powerlaw, grid, 256, 0, stream, m, d_w, lo, cap, remainder, d_len);
        MXG_CUDA_TRY(cudaFreeAsync(d_w, stream));
        MXG_CUDA_TRY(cudaFreeAsync(d_total, stream));
    } else {
        return fail(MXG_ERR_ARG, "synth: unknown row_model %d", row_model);
    }

    MXG_TRY(exclusive_scan_i32(d_len, d_p, (size_t)m + 1, stream));
    int32_t nnz32 = 0;
    MXG_CUDA_TRY(cudaMemcpyAsync(&nnz32, d_p + m, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    MXG_CUDA_TRY(cudaFreeAsync(d_len, stream));
    const size_t nnz = (size_t)nnz32;

    int32_t *d_j = nullptr;
    double *d_x64 = nullptr;
    float *d_x32 = nullptr;
    MXG_CUDA_TRY(cudaMallocAsync(&d_j, sizeof(int32_t) * (nnz ? nnz : 1), stream));
    if (keep & MXG_KEEP_F64) MXG_CUDA_TRY(cudaMallocAsync(&d_x64, sizeof(double) * (nnz ? nnz : 1), stream));
    if (keep & MXG_KEEP_F32) MXG_CUDA_TRY(cudaMallocAsync(&d_x32, sizeof(float) * (nnz ? nnz : 1), stream));
    int gf = ceil_div_i(m, 8);
    if (gf > 148 * 32) gf = 148 * 32;
    if (d_x64 && d_x32)
        MXG_LAUNCH((k_fill_entries<true, true>), gf, 256, 0, stream, m, K, d_p, col_model, seed, d_j, d_x64, d_x32);
    else if (d_x64)
        MXG_LAUNCH((k_fill_entries<true, false>), gf, 256, 0, stream, m, K, d_p, col_model, seed, d_j, d_x64, d_x32);
    else if (d_x32)
        MXG_LAUNCH((k_fill_entries<false, true>), gf, 256, 0, stream, m, K, d_p, col_model, seed, d_j, d_x64, d_x32);
    else
        MXG_LAUNCH((k_fill_entries<false, false>), gf, 256, 0, stream, m, K, d_p, col_model, seed, d_j, d_x64, d_x32);
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    *out_p = d_p;
    *out_j = d_j;
    *out_x64 = d_x64;
    *out_x32 = d_x32;
    *out_nnz = (int64_t)nnz;
    return MXG_OK;
}


// ------------------------------------------------------------------------------------------------
// Measurement probe (tools/sweep.py --what gatherroof): the fastest this GPU can fetch RANDOM rows of
// `row_bytes` (128 / 256 / 512) from a table of `rows` rows — the access pattern of the dense-operand
// gathers of K1/K2 with nothing else attached (no CSR stream, no FMAs, no output): one team of
// row_bytes/16 lanes per row, 8 independent 16-byte loads in flight per lane, row ids from a hash of a
// counter.  With a table far larger than L2 this is the DRAM random-access roof for that row size, with a
// table inside L2 the L2->SM gather roof.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t probe_hash(uint64_t v)
{
    v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
    return (uint32_t)v;
}

template <int LPR>
__global__ void __launch_bounds__(256) k_gather_probe(const float4 *__restrict__ table, uint32_t rows, size_t row_vec,
                                                      long long gathers_per_team, uint64_t seed, float *sink)
{
    constexpr int U = 8;
    const int l = threadIdx.x % LPR;
    const uint64_t team = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long g = 0; g < gathers_per_team; g += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t r = (uint32_t)(((uint64_t)probe_hash(seed + team * 0x9E3779B97F4A7C15ULL + (uint64_t)(g + u)) * rows) >> 32);
            v[u] = __ldg(table + (size_t)r * row_vec + l);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x; // never true: keeps the loads alive
}

int gather_probe(int row_bytes, const void *d_table, size_t rows, long long gathers, uint64_t seed, float *d_sink,
                 cudaStream_t stream)
{
    if (rows == 0 || rows > 0xffffffffULL || gathers <= 0) return fail(MXG_ERR_ARG, "gather_probe: bad size");
    const int lpr = row_bytes / 16;
    const int grid = 148 * 8;
    const long long teams = (long long)grid * 256 / lpr;
    const long long per_team = std::max<long long>(8, (gathers / teams + 7) / 8 * 8);
    const float4 *t = static_cast<const float4 *>(d_table);
    switch (lpr) {
    case 8: MXG_LAUNCH(k_gather_probe<8>, grid, 256, 0, stream, t, (uint32_t)rows, (size_t)lpr, per_team, seed, d_sink); break;
    case 16: MXG_LAUNCH(k_gather_probe<16>, grid, 256, 0, stream, t, (uint32_t)rows, (size_t)lpr, per_team, seed, d_sink); break;
    case 32: MXG_LAUNCH(k_gather_probe<32>, grid, 256, 0, stream, t, (uint32_t)rows, (size_t)lpr, per_team, seed, d_sink); break;
    default: return fail(MXG_ERR_ARG, "gather_probe: row_bytes must be 128, 256 or 512");
    }
    return MXG_OK;
}

} // namespace mxg
