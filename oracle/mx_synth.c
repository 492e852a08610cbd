/* oracle/mx_synth.c — HOST twin of the device generator of synthetic CSR inputs (matrixextra_b200/csrc/synth.cu),
 * TEST INFRASTRUCTURE ONLY: lets bench.py's reference arm time the reference's CPU path on the SAME workload
 * (same shape, same row-length law, same column law, same RNG and counters: SURVEY.md 8d) without loading the
 * product library or touching a GPU.  The recipe, restated:
 *   row lengths  row_model 0: target/m per row, +-20 % jitter that cancels in pairs (nnz == target exactly);
 *                row_model 1: len_r = min(cap, floor(L * (1-u_r)^(-1/1.5))), cap = min(K, 65536), L by bisection so
 *                that sum(len) <= target, remainder spread one entry per row over the first rows;
 *   columns      col_model 0: entry t of a row of length len in [floor(t*K/len), floor((t+1)*K/len));
 *                col_model 1: strata b_t = t + floor((K-len) * (t/len)^2) (popular low columns);
 *   values       uniform in [-1, 1).
 *   RNG: Philox4x32-10, counter = (row, entry, purpose, 0), key = seed.
 * Integer-for-integer the same as the device code; the row weights go through libm's pow() here and CUDA's there,
 * so a row length can differ by one entry where L*w lands within an ulp of an integer (tests/test_oracle.py
 * bounds the difference; the reference arm needs the same workload, not the same bits).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { uint32_t x, y, z, w; } u4;

static inline u4 philox4x32_10(u4 c, uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        u4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

static inline double u01(uint32_t hi, uint32_t lo)
{
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

static inline u4 rng(uint64_t seed, uint32_t row, uint32_t t, uint32_t purpose)
{
    u4 c = {row, t, purpose, 0u};
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

static long long sum_lengths(int m, const double *w, double L, int cap)
{
    long long s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int r = 0; r < m; r++) {
        const double v = floor(L * w[r]);
        s += (long long)(v < (double)cap ? v : (double)cap);
    }
    return s;
}

/* Step 1: p[m+1] (caller-allocated); returns nnz = p[m], or -1 on bad arguments. */
long long mxs_synth_indptr(int m, int K, long long target, int row_model, uint64_t seed, int32_t *p)
{
    if (m <= 0 || K <= 0 || target < 0 || target > 2147483647LL || target > (long long)m * K) return -1;
    int32_t *len = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    if (!len) return -1;
    if (row_model == 0) {
        const long long base = target / m, rem = target - base * m;
        long long jit = base / 5;
        if (jit > (long long)K - base - 1) jit = (long long)K - base - 1;
        if (jit < 0) jit = 0;
#pragma omp parallel for schedule(static)
        for (int r = 0; r < m; r++) {
            const int pair = r >> 1;
            long long d = 0;
            if ((pair * 2 + 1) < m && jit > 0) d = (long long)(rng(seed, (uint32_t)pair, 0u, 2u).x % (uint32_t)(jit + 1));
            long long l = base + ((r & 1) ? -d : d);
            if ((long long)r < rem) l += 1;
            if (l > K) l = K;
            len[r] = (int32_t)l;
        }
    } else if (row_model == 1) {
        const int cap = K < 65536 ? K : 65536;
        if ((long long)m * cap < target) { free(len); return -1; }
        double *w = (double *)malloc(sizeof(double) * (size_t)m);
        if (!w) { free(len); return -1; }
#pragma omp parallel for schedule(static)
        for (int r = 0; r < m; r++) {
            const u4 z = rng(seed, (uint32_t)r, 0u, 1u);
            w[r] = pow(1.0 - u01(z.x, z.y), -1.0 / 1.5);
        }
        double lo = 0.0, hi = (double)target / (double)m + 1.0;
        for (int it = 0; it < 64 && sum_lengths(m, w, hi, cap) < target; it++) hi *= 2.0;
        for (int it = 0; it < 60; it++) {
            const double mid = 0.5 * (lo + hi);
            if (sum_lengths(m, w, mid, cap) <= target) lo = mid;
            else hi = mid;
        }
        long long remainder = target - sum_lengths(m, w, lo, cap);
        if (remainder < 0) remainder = 0;
        if (remainder > m) remainder = m;
#pragma omp parallel for schedule(static)
        for (int r = 0; r < m; r++) {
            const double v = floor(lo * w[r]);
            int l = (int)(v < (double)cap ? v : (double)cap);
            if ((long long)r < remainder && l < cap) l += 1;
            len[r] = l;
        }
        free(w);
    } else {
        free(len);
        return -1;
    }
    long long run = 0;
    for (int r = 0; r < m; r++) {
        p[r] = (int32_t)run;
        run += len[r];
    }
    p[m] = (int32_t)run;
    free(len);
    return run;
}

/* Step 2: j[nnz], x[nnz] for the indptr of step 1. */
void mxs_synth_entries(int m, int K, const int32_t *p, int col_model, uint64_t seed, int32_t *j, double *x)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (int r = 0; r < m; r++) {
        const int a = p[r], len = p[r + 1] - a;
        for (int t = 0; t < len; t++) {
            long long lo, hi;
            if (col_model == 0) {
                lo = ((long long)t * K) / len;
                hi = ((long long)(t + 1) * K) / len;
            } else {
                const double f0 = (double)t / (double)len, f1 = (double)(t + 1) / (double)len;
                lo = t + (long long)floor((double)(K - len) * f0 * f0);
                hi = (t + 1 == len) ? (long long)K : (t + 1) + (long long)floor((double)(K - len) * f1 * f1);
            }
            const u4 z = rng(seed, (uint32_t)r, (uint32_t)t, 3u);
            long long c = lo + (long long)(u01(z.x, z.y) * (double)(hi - lo));
            if (c >= hi) c = hi - 1;
            if (c < lo) c = lo;
            j[a + t] = (int32_t)c;
            x[a + t] = 2.0 * u01(z.z, z.w) - 1.0;
        }
    }
}
