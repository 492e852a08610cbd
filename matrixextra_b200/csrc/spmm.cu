// spmm.cu — K1/K2: CSR x dense gather products for sm_100a.
//
// Replaces gemm_csr_drm_as_drm (src/matmul.cpp:118-142, row-major output) and gemm_csr_drm_as_dcm
// (src/matmul.cpp:150-185, column-major output) of the reference.
//
// Decomposition ("row split, column lanes"):
//   * a TEAM of LPR lanes owns one CSR row at a time; lane l of the team owns CPL vectors of V
//     consecutive output columns (V*sizeof(T) = 16 bytes on the vector path), so one gather of a
//     dense row B[j,:] is LPR coalesced 128-bit loads and the row sum stays in registers in the
//     reference's order (sequential over the stored entries, starting from 0);
//   * the team's lanes load LPR (index, value) pairs coalesced, then broadcast them one at a time
//     with width-LPR shuffles; U=4 gathers are issued back to back before their FMAs so every lane
//     keeps 4*CPL 16-byte loads in flight;
//   * teams of one warp run in lock-step over max(row length) with predicates, never divergent;
//   * rows longer than `piece` entries are cut into pieces at upload time (K7, layout.cu); pieces
//     are scheduled first in the grid, write partial sums to a workspace and a tiny fix-up kernel
//     adds each row's pieces in piece order => deterministic, no atomics;
//   * column-major output (K2): a CTA owns BR consecutive rows, parks the finished rows in a
//     shared-memory tile [columns][BR+1] and writes every column as BR consecutive elements.
// Bound: L2->SM gather bandwidth for n >= 32 (each stored entry pulls n*sizeof(T) bytes), HBM for
// the streamed CSR arrays and the output; no tensor cores (unstructured, not a dense contraction).
#include "mxg_internal.cuh"

#include <algorithm>

namespace mxg {

template <typename T, int V>
struct alignas(sizeof(T) * V) Pack {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Pack<T, V> ld_ro(const T *p);

template <>
__device__ __forceinline__ Pack<float, 4> ld_ro<float, 4>(const float *p)
{
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    Pack<float, 4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
template <>
__device__ __forceinline__ Pack<double, 2> ld_ro<double, 2>(const double *p)
{
    const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    Pack<double, 2> r;
    r.v[0] = t.x; r.v[1] = t.y;
    return r;
}
template <>
__device__ __forceinline__ Pack<float, 1> ld_ro<float, 1>(const float *p)
{
    Pack<float, 1> r;
    r.v[0] = __ldg(p);
    return r;
}
template <>
__device__ __forceinline__ Pack<double, 1> ld_ro<double, 1>(const double *p)
{
    Pack<double, 1> r;
    r.v[0] = __ldg(p);
    return r;
}

template <typename T, int V>
__device__ __forceinline__ void st_pack(T *p, const Pack<T, V> &v)
{
    *reinterpret_cast<Pack<T, V> *>(p) = v;
}

// streaming (evict-first) accesses for data that is touched once per launch: the CSR arrays and the output
// rows must not push rows of the dense operand out of L2
__device__ __forceinline__ Pack<float, 4> ld_stream(const float *p, Pack<float, 4> *)
{
    const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
    Pack<float, 4> r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ Pack<double, 2> ld_stream(const double *p, Pack<double, 2> *)
{
    const double2 t = __ldcs(reinterpret_cast<const double2 *>(p));
    Pack<double, 2> r;
    r.v[0] = t.x; r.v[1] = t.y;
    return r;
}
__device__ __forceinline__ Pack<float, 1> ld_stream(const float *p, Pack<float, 1> *)
{
    Pack<float, 1> r;
    r.v[0] = __ldcs(p);
    return r;
}
__device__ __forceinline__ Pack<double, 1> ld_stream(const double *p, Pack<double, 1> *)
{
    Pack<double, 1> r;
    r.v[0] = __ldcs(p);
    return r;
}
__device__ __forceinline__ void st_stream(float *p, const Pack<float, 4> &v)
{
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
}
__device__ __forceinline__ void st_stream(double *p, const Pack<double, 2> &v)
{
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v.v[0], v.v[1]));
}
__device__ __forceinline__ void st_stream(float *p, const Pack<float, 1> &v) { __stcs(p, v.v[0]); }
__device__ __forceinline__ void st_stream(double *p, const Pack<double, 1> &v) { __stcs(p, v.v[0]); }

// NVLS multicast stores: one store to a multicast address (cuMulticast / torch symmetric memory) is replicated by
// the NVSwitch into the buffers of ALL GPUs bound to it, so a row block leaves its GPU once instead of G-1 times
__device__ __forceinline__ void st_mcast(float *p, const Pack<float, 4> &v)
{
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.v[0]), "f"(v.v[1]), "f"(v.v[2]),
                 "f"(v.v[3])
                 : "memory");
}
__device__ __forceinline__ void st_mcast(double *p, const Pack<double, 2> &v)
{
    // 16 bytes are 16 bytes: the switch replicates bits, so a double pair travels as four 32-bit lanes
    const float a = __int_as_float(__double2loint(v.v[0])), b = __int_as_float(__double2hiint(v.v[0]));
    const float c = __int_as_float(__double2loint(v.v[1])), d = __int_as_float(__double2hiint(v.v[1]));
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_mcast(float *p, const Pack<float, 1> &v)
{
    asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v.v[0]) : "memory");
}
__device__ __forceinline__ void st_mcast(double *p, const Pack<double, 1> &v)
{
    asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v.v[0]) : "memory");
}
__device__ __forceinline__ void st_mcast_scalar(float *p, float v) { asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void st_mcast_scalar(double *p, double v) { asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// ------------------------------------------------------------------------------------------------
// One WARP owns one row at a time.  Its 32 lanes form SPLIT = 32 / LPR sub-teams of LPR lanes; lane l
// of a sub-team owns CPL vectors of V output columns, sub-team s takes the row's entries
// s, s + SPLIT, s + 2 SPLIT, ... (so a gather instruction of the warp fetches SPLIT whole rows of B),
// and the sub-team sums are added by a butterfly when the row ends.  With SPLIT == 1 (n*sizeof(T)
// >= 512 bytes) the sum is the reference's left-to-right chain; otherwise it is SPLIT interleaved
// chains plus a log2(SPLIT) tree — the reassociation the fp tolerances (1e-12 / 1e-5) allow for.
// ------------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;
constexpr int SPMM_THREADS = 128;
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
#ifndef SPMM_MINB
#define SPMM_MINB 8 // resident CTAs per SM the register allocation must allow (8 => 64 registers)
#endif
constexpr int SPMM_CM_RPW = 8; // rows per warp of a column-major CTA tile (tile = 32 rows)
// rows per warp of a BULK CTA: as many as keep the CTA's staging tile within 32 KB, 8 .. 16 (fp32 n = 64: 16 rows per
// warp, 64 rows = 16 KB per bulk copy and destination)
__host__ __device__ constexpr int spmm_bulk_rpw(int nb, int elem_bytes)
{
    return (32768 / (SPMM_WARPS * nb * elem_bytes)) >= 16 ? 16 : ((32768 / (SPMM_WARPS * nb * elem_bytes)) <= 8 ? 8 : (32768 / (SPMM_WARPS * nb * elem_bytes)));
}

// Entries pos .. pos + cnt - 1 (cnt <= 32, one per lane in jj / xx) of the current row.
template <typename T, int V, int LPR, int CPL, int U>
__device__ __forceinline__ void batch_gather(Pack<T, V> (&acc)[CPL], const int jj, const T xx, const int cnt,
                                             const int sub, const T *__restrict__ B, const size_t ldb,
                                             const int (&col)[CPL], const bool (&cok)[CPL])
{
    constexpr int SPLIT = 32 / LPR;
    for (int k = 0; k < cnt; k += SPLIT * U) { // warp-uniform trip count
        const T *src[U];
        T xk[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int idx = k + u * SPLIT + sub;
            const int jk = __shfl_sync(FULL, jj, idx & 31);
            xk[u] = __shfl_sync(FULL, xx, idx & 31);
            ok[u] = idx < cnt;
            src[u] = B + (size_t)jk * ldb;
        }
        Pack<T, V> bv[U][CPL];
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int c = 0; c < CPL; c++)
                if (ok[u] && cok[c]) bv[u][c] = ld_ro<T, V>(src[u] + col[c]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int c = 0; c < CPL; c++)
                if (ok[u] && cok[c]) {
#pragma unroll
                    for (int i = 0; i < V; i++) acc[c].v[i] = fma_t(xk[u], bv[u][c].v[i], acc[c].v[i]);
                }
        }
    }
}

template <typename T, int V, int LPR, int CPL>
__device__ __forceinline__ void subteam_reduce(Pack<T, V> (&acc)[CPL])
{
#pragma unroll
    for (int d = LPR; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < CPL; c++)
#pragma unroll
            for (int i = 0; i < V; i++) acc[c].v[i] += __shfl_xor_sync(FULL, acc[c].v[i], d);
    }
}

template <typename T, int V, int CPL>
__device__ __forceinline__ void zero_acc(Pack<T, V> (&acc)[CPL])
{
#pragma unroll
    for (int c = 0; c < CPL; c++)
#pragma unroll
        for (int i = 0; i < V; i++) acc[c].v[i] = T(0);
}

struct SpmmArgs {
    int m, n;
    const int32_t *p;
    const int32_t *j;
    const void *x;
    const void *B;
    size_t ldb;
    void *Out;  // the local result
    size_t ldc;
    // additional copies of every finished row (peer-mapped result buffers of the other GPUs of the box, same
    // leading dimension): the all-gather of the row blocks happens store by store, inside the product
    void *extra[MXG_MAX_DST - 1];
    int n_extra;
    int mcast; // Out is a multicast address: rows are written with multimem.st and land on every GPU (row-major only)
    int bulk;  // MULTI launches: rows leave as bulk copies from shared memory (BULK instantiation; ldc == n == one column block)
    int rpw; // consecutive rows per warp (<= 31; column-major output uses SPMM_CM_RPW)
    int piece;
    int n_pieces;
    int piece_blocks;
    const int32_t *piece_row;
    const int32_t *piece_k;
    void *partial; // [n_pieces][n]
    const int *abort; // optional device flag: non-zero => the column ids failed validation, do nothing
    // column panels (PANELS kernels): this launch handles the entries with column id in
    // [panel * panel_width, (panel + 1) * panel_width); seg[(q - 1) * m + r] = first entry of row r in panel q
    const int32_t *seg;
    int panel, n_panels, panel_width;
};

// Grid: x = [piece CTAs | row CTAs], y = column blocks of NB = LPR*V*CPL output columns.
// A row CTA owns SPMM_WARPS * rpw consecutive rows, rpw consecutive rows per warp.  A warp walks its rows
// batch by batch (32 stored entries per batch, loaded coalesced); the (index, value) batch that follows
// the one being gathered — in the same row or at the start of the next — is already in flight, so the
// only exposed latencies are the gathers themselves, U whole-row gathers deep per sub-team.
//
// PANELS: the dense operand is wider than L2, so the product runs as one launch per COLUMN PANEL of A (a slab
// of rows of B small enough to stay L2-resident): launch q adds, for every row, the entries whose column lies
// in panel q to the output row (read-modify-write from the second panel on; rows without entries in the
// panel are not touched).  Entries are consumed in stored order panel after panel, so for sorted rows the sum
// order is unchanged; for unsorted rows the split points still partition the row (see k_panel_segments).
// MULTI (row-major results only): the launch writes more than the local result — peer copies (g.extra) or an NVLS
// multicast address (g.mcast).  A separate instantiation so that the single-destination kernel keeps the exact
// instruction schedule it was tuned with (the shared version cost the fp64 variant 17 %).
// BULK (a MULTI variant for results whose rows are exactly one column block wide, ldc == n == NB): a warp's SPMM_BULK_RPW
// finished rows are parked in its slice of shared memory — they are contiguous in the rows-contiguous result — and
// shipped to every destination (the local result and the peer-mapped results of the other GPUs) as ONE bulk
// asynchronous copy each (cp.async.bulk.global.shared::cta: the TMA unit streams 2 - 4 KB per destination over NVLink)
// instead of rpw x n/V 16-byte stores per destination issued by the SM's load/store path.
template <typename T, int V, int LPR, int CPL, int U, bool COLMAJOR, bool PANELS, int MB = (CPL == 1 ? SPMM_MINB : 4),
          bool MULTI = false, bool BULK = false>
__global__ void __launch_bounds__(SPMM_THREADS, MB) k_spmm(const SpmmArgs g)
{
    constexpr int NB = LPR * V * CPL; // output columns per CTA column block
    constexpr int BR = SPMM_WARPS * SPMM_CM_RPW;
    constexpr int TLD = BR + 1;
    __shared__ T tile[COLMAJOR ? NB * TLD : 1];
    __shared__ unsigned char s_write[COLMAJOR ? BR : 1]; // tile rows this launch has to write
    constexpr int BRPW = spmm_bulk_rpw(NB, (int)sizeof(T));
    __shared__ __align__(128) T stage[BULK ? SPMM_WARPS * BRPW * NB : 1]; // [warp][row of the warp][column]

    const int32_t *__restrict__ p = g.p;
    const int32_t *__restrict__ j = g.j;
    const T *__restrict__ x = static_cast<const T *>(g.x);
    const T *__restrict__ B = static_cast<const T *>(g.B);
    T *__restrict__ Out = static_cast<T *>(g.Out);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane / LPR;
    const int l = lane % LPR;
    const int col0 = blockIdx.y * NB;

    int col[CPL];
    bool cok[CPL];
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        col[c] = col0 + (c * LPR + l) * V;
        cok[c] = col[c] < g.n; // n % V == 0 on the vector path, so a vector is entirely in or out
    }

    if (g.abort != nullptr && *g.abort != 0) return; // whole grid, before any barrier

    Pack<T, V> acc[CPL];
    zero_acc<T, V, CPL>(acc);

    if ((int)blockIdx.x < g.piece_blocks) {
        // ---- long-row pieces: one warp per piece, partial sums to the workspace -----------------------
        const int pc = blockIdx.x * SPMM_WARPS + warp;
        if (pc >= g.n_pieces) return; // warp-uniform, no barrier in this section
        const int row = g.piece_row[pc];
        const int a = p[row] + g.piece_k[pc] * g.piece;
        const int b = min(a + g.piece, p[row + 1]);
        if (PANELS) {
            // a piece (<= `piece` consecutive entries of one row) runs in the launch of its first entry's panel
            const int pq = min(__ldg(j + a) / g.panel_width, g.n_panels - 1);
            if (pq != g.panel) return;
        }
        int cj = 0;
        T cx = T(0);
        if (a + lane < b) {
            cj = __ldcs(j + a + lane);
            cx = __ldcs(x + a + lane);
        }
        for (int pos = a; pos < b; pos += 32) {
            int nj = 0;
            T nx = T(0);
            if (pos + 32 + lane < b) {
                nj = __ldcs(j + pos + 32 + lane);
                nx = __ldcs(x + pos + 32 + lane);
            }
            batch_gather<T, V, LPR, CPL, U>(acc, cj, cx, min(32, b - pos), sub, B, g.ldb, col, cok);
            cj = nj;
            cx = nx;
        }
        subteam_reduce<T, V, LPR, CPL>(acc);
        if (sub == 0) {
            T *dst = static_cast<T *>(g.partial) + (size_t)pc * g.n;
#pragma unroll
            for (int c = 0; c < CPL; c++)
                if (cok[c]) st_pack<T, V>(dst + col[c], acc[c]);
        }
        return;
    }

    const int rb = blockIdx.x - g.piece_blocks;
    const int rpw = COLMAJOR ? SPMM_CM_RPW : (BULK ? BRPW : g.rpw);
    const int row0 = (rb * SPMM_WARPS + warp) * rpw;
    const int nr = min(rpw, g.m - row0); // rows this warp owns (<= 0: none)

    const bool first_panel = !PANELS || g.panel == 0;
    if (COLMAJOR) {
        if (lane < SPMM_CM_RPW) s_write[warp * SPMM_CM_RPW + lane] = 0;
        __syncwarp();
    }

    if (nr > 0) {
        int pv = 0, sa = 0, se = 0;
        if (lane <= nr) pv = __ldg(p + row0 + lane);
        if (PANELS) {
            const int pnext = __shfl_down_sync(FULL, pv, 1);
            if (lane < nr) {
                const size_t r_ = (size_t)row0 + lane;
                sa = g.panel > 0 ? __ldg(g.seg + (size_t)(g.panel - 1) * g.m + r_) : pv;
                se = g.panel < g.n_panels - 1 ? __ldg(g.seg + (size_t)g.panel * g.m + r_) : pnext;
            }
        }
        // row rr: entries [ra, re); rows longer than a piece are left to the piece section + fix-up kernel
        auto bounds = [&](const int rr, int &ra, int &re, bool &rskip) {
            const int rp0 = __shfl_sync(FULL, pv, rr);
            const int rp1 = __shfl_sync(FULL, pv, rr + 1);
            rskip = (rp1 - rp0) > g.piece;
            if (PANELS) {
                ra = __shfl_sync(FULL, sa, rr);
                re = rskip ? ra : __shfl_sync(FULL, se, rr);
            } else {
                ra = rp0;
                re = rskip ? ra : rp1;
            }
        };
        int r = 0, a, e;
        bool skip;
        bounds(0, a, e, skip);
        int pos = a;
        int cj = 0;
        T cx = T(0);
        if (pos + lane < e) {
            cj = __ldcs(j + pos + lane);
            cx = __ldcs(x + pos + lane);
        }
        Pack<T, V> prev[CPL]; // later panels: the output row so far, fetched while the row's gathers are in flight
        while (true) {
            const int cnt = min(32, e - pos); // <= 0 for a row without (eligible) entries
            if (PANELS && !COLMAJOR && !first_panel && pos == a && e > a && sub == 0) {
                const T *src = Out + (size_t)(row0 + r) * g.ldc;
#pragma unroll
                for (int c = 0; c < CPL; c++)
                    if (cok[c]) prev[c] = ld_stream(src + col[c], (Pack<T, V> *)nullptr);
            }
            // where the next batch starts: further along this row, or at the start of the next one
            int npos = pos + 32, nrow = r, na = a, ne = e;
            bool nskip = skip;
            const bool last = npos >= e;
            if (last) {
                nrow = r + 1;
                if (nrow < nr) {
                    bounds(nrow, na, ne, nskip);
                    npos = na;
                }
            }
            int nj = 0;
            T nx = T(0);
            if (nrow < nr && npos + lane < ne) {
                nj = __ldcs(j + npos + lane);
                nx = __ldcs(x + npos + lane);
            }
            batch_gather<T, V, LPR, CPL, U>(acc, cj, cx, cnt, sub, B, g.ldb, col, cok);
            if (last) {
                // the first panel writes every (short) row, later panels only the rows they have entries for
                if (!skip && (first_panel || e > a)) {
                    subteam_reduce<T, V, LPR, CPL>(acc);
                    if (sub == 0) {
                        if (COLMAJOR) {
                            const int rl = warp * SPMM_CM_RPW + r;
                            if (l == 0) s_write[rl] = 1;
#pragma unroll
                            for (int c = 0; c < CPL; c++)
                                if (cok[c]) {
#pragma unroll
                                    for (int i = 0; i < V; i++) tile[(col[c] - col0 + i) * TLD + rl] = acc[c].v[i];
                                }
                        } else if (BULK) {
                            T *srow = stage + (size_t)(warp * BRPW + r) * NB;
#pragma unroll
                            for (int c = 0; c < CPL; c++)
                                if (cok[c]) st_pack<T, V>(srow + (col[c] - col0), acc[c]);
                        } else {
                            T *dst = Out + (size_t)(row0 + r) * g.ldc;
#pragma unroll
                            for (int c = 0; c < CPL; c++)
                                if (cok[c]) {
                                    if (!first_panel) {
#pragma unroll
                                        for (int i = 0; i < V; i++) acc[c].v[i] = prev[c].v[i] + acc[c].v[i];
                                    }
                                    if (MULTI && g.mcast) st_mcast(dst + col[c], acc[c]);
                                    else st_stream(dst + col[c], acc[c]);
                                }
                        }
                    }
                    if (MULTI && !COLMAJOR && !PANELS && !BULK) {
                        // copies for the other GPUs: after the butterfly every sub-team holds the same bits, so the
                        // destinations are dealt out over the sub-teams and one store instruction of the warp writes
                        // to 32 / LPR peers at once
                        for (int d = sub; d < g.n_extra; d += 32 / LPR) {
                            T *dst = static_cast<T *>(g.extra[d]) + (size_t)(row0 + r) * g.ldc;
#pragma unroll
                            for (int c = 0; c < CPL; c++)
                                if (cok[c]) st_stream(dst + col[c], acc[c]);
                        }
                    }
                }
                if (BULK && skip && sub == 0) {
                    // a long row: written by the fix-up launch afterwards; its slot must not ship stale shared memory
                    T *srow = stage + (size_t)(warp * BRPW + r) * NB;
#pragma unroll
                    for (int c = 0; c < CPL; c++)
                        if (cok[c]) st_pack<T, V>(srow + (col[c] - col0), acc[c]); // acc is zero: nothing was gathered
                }
                zero_acc<T, V, CPL>(acc);
                if (nrow >= nr) break;
            }
            r = nrow;
            a = na;
            e = ne;
            skip = nskip;
            pos = npos;
            cj = nj;
            cx = nx;
        }
    }

    if (BULK && g.bulk == 2) {
        // CTA-wide variant (option spmm_bulk = 2): the four warps' rows are one contiguous block of every destination —
        // one bulk copy of up to 32 rows (8 KB for fp32 n = 64) per destination, issued by warp 0 once every warp is done
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const int cta_row0 = rb * SPMM_WARPS * BRPW;
        const int cta_rows = min(SPMM_WARPS * BRPW, g.m - cta_row0);
        if (warp == 0 && lane <= g.n_extra && cta_rows > 0) {
            T *base = Out;
#pragma unroll
            for (int d = 0; d < MXG_MAX_DST - 1; d++)
                if (lane == d + 1) base = static_cast<T *>(g.extra[d]);
            T *dst = base + (size_t)cta_row0 * g.ldc;
            const unsigned bytes = (unsigned)(cta_rows * NB * (int)sizeof(T));
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(stage);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else if (BULK && nr > 0) {
        // the warp's nr rows are nr * NB contiguous elements of every destination: one bulk copy per destination,
        // issued by one lane each (generic proxy writes -> async proxy reads need the proxy fence)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane <= g.n_extra) {
            T *base = Out;
#pragma unroll
            for (int d = 0; d < MXG_MAX_DST - 1; d++)
                if (lane == d + 1) base = static_cast<T *>(g.extra[d]);
            T *dst = base + (size_t)row0 * g.ldc;
            const unsigned bytes = (unsigned)(nr * NB * (int)sizeof(T));
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(stage + (size_t)warp * BRPW * NB);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory stays valid until it has been read
        }
    }
    if (MULTI && !COLMAJOR && !PANELS && g.mcast) __threadfence_system(); // multicast rows are on their way before the step's barrier
    if (COLMAJOR) {
        __syncthreads();
        const int tile_row0 = rb * BR;
        const int row = tile_row0 + lane; // BR == 32: one tile row per lane
        const int ncols = min(NB, g.n - col0);
        // long rows are written by the fix-up kernel, rows without entries in a later panel stay as they are
        if (row < g.m && s_write[lane]) {
            T *dst = Out + (size_t)row;
            for (int c = warp; c < ncols; c += SPMM_WARPS) {
                T v = tile[c * TLD + lane];
                T *q = dst + (size_t)(col0 + c) * g.ldc;
                if (!first_panel) v = __ldcs(q) + v;
                __stcs(q, v);
                for (int d = 0; d < g.n_extra; d++) __stcs(static_cast<T *>(g.extra[d]) + (size_t)row + (size_t)(col0 + c) * g.ldc, v);
            }
        }
    }
}

// seg[(q - 1) * m + r] = first entry of row r whose column id is >= q * width, q = 1 .. P-1, searched from the
// previous split point on: the P segments always partition the row's entries, sorted or not (for unsorted rows
// only the L2 locality of the panels suffers, never the result).
__global__ void __launch_bounds__(256) k_panel_segments(int m, const int32_t *__restrict__ p, const int32_t *__restrict__ j,
                                                        int width, int n_panels, int32_t *__restrict__ seg)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    int lo = p[r];
    const int hi = p[r + 1];
    for (int q = 1; q < n_panels; q++) {
        const int target = q * width;
        int a = lo, b = hi; // first e in [lo, hi) with j[e] >= target
        while (a < b) {
            const int mid = a + ((b - a) >> 1);
            if (__ldg(j + mid) < target) a = mid + 1;
            else b = mid;
        }
        lo = a;
        seg[(size_t)(q - 1) * m + r] = lo;
    }
}

// Adds the pieces of every long row in piece order and writes the row (to every destination).
struct DstList {
    void *dst[MXG_MAX_DST];
    int n;
};

template <typename T, bool COLMAJOR>
__global__ void __launch_bounds__(128) k_spmm_fixup(int n, const int32_t *__restrict__ long_rows,
                                                    const int32_t *__restrict__ long_first,
                                                    const int32_t *__restrict__ long_np,
                                                    const T *__restrict__ partial, const DstList out, size_t ldc,
                                                    const int *__restrict__ abort, int mcast)
{
    if (abort != nullptr && *abort != 0) return;
    const int row = long_rows[blockIdx.x];
    const int first = long_first[blockIdx.x];
    const int np = long_np[blockIdx.x];
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        T s = T(0);
        for (int k = 0; k < np; k++) s += partial[(size_t)(first + k) * n + c];
        const size_t at = COLMAJOR ? (size_t)row + (size_t)c * ldc : (size_t)row * ldc + c;
        if (mcast) st_mcast_scalar(static_cast<T *>(out.dst[0]) + at, s);
        else
            for (int d = 0; d < out.n; d++) static_cast<T *>(out.dst[d])[at] = s;
    }
    if (mcast) __threadfence_system();
}

// Output rows of a matrix without stored entries (or with m rows but n == 0) still have to be zero.
template <typename T>
__global__ void __launch_bounds__(256) k_fill_zero_2d(T *__restrict__ Out, size_t rows, size_t cols, size_t ld)
{
    const size_t total = rows * cols;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        Out[(i / cols) * ld + (i % cols)] = T(0);
}

// Two vectors per lane on a half-width team (LPR = half the row's vectors): one gather instruction of the warp
// fetches 2 * SPLIT half rows, so a lane keeps 8 independent 16-byte loads in flight for the shuffle / address
// work of 4.  Costs registers: MB resident CTAs per SM (80 / 96 registers) instead of 8 (64).  Measured on B200
// (profiles/README.md, r01 v5 sweep): -12 % on 256-byte rows (fp32 n = 64), -11 % on 512-byte fp32 rows, -4 % fp64.
template <typename T, int V, int LPR, bool COLMAJOR>
static int launch_two(SpmmArgs &args, int row_blocks, cudaStream_t stream)
{
    constexpr int MB = sizeof(T) == 4 ? 6 : 5;
    constexpr int NB = LPR * V * 2;
    dim3 grid((unsigned)(args.piece_blocks + row_blocks), (unsigned)ceil_div_i(args.n, NB), 1);
    if (!COLMAJOR && args.bulk && args.n == NB)
        MXG_LAUNCH((k_spmm<T, V, LPR, 2, 4, COLMAJOR, false, MB, !COLMAJOR, !COLMAJOR>), grid, SPMM_THREADS, 0, stream, args);
    else if (!COLMAJOR && (args.n_extra > 0 || args.mcast))
        MXG_LAUNCH((k_spmm<T, V, LPR, 2, 4, COLMAJOR, false, MB, !COLMAJOR>), grid, SPMM_THREADS, 0, stream, args);
    else
        MXG_LAUNCH((k_spmm<T, V, LPR, 2, 4, COLMAJOR, false, MB>), grid, SPMM_THREADS, 0, stream, args);
    return MXG_OK;
}

template <typename T, int V, int LPR, int CPL, int U, bool COLMAJOR>
static int launch_variant(SpmmArgs &args, int row_blocks, cudaStream_t stream)
{
    constexpr int NB = LPR * V * CPL;
    dim3 grid((unsigned)(args.piece_blocks + row_blocks), (unsigned)ceil_div_i(args.n, NB), 1);
    if (args.n_panels <= 1) {
        if (!COLMAJOR && args.bulk && args.n == NB)
            MXG_LAUNCH((k_spmm<T, V, LPR, CPL, U, COLMAJOR, false, (CPL == 1 ? SPMM_MINB : 4), !COLMAJOR, !COLMAJOR>), grid,
                       SPMM_THREADS, 0, stream, args);
        else if (!COLMAJOR && (args.n_extra > 0 || args.mcast))
            MXG_LAUNCH((k_spmm<T, V, LPR, CPL, U, COLMAJOR, false, (CPL == 1 ? SPMM_MINB : 4), !COLMAJOR>), grid, SPMM_THREADS, 0,
                       stream, args);
        else
            MXG_LAUNCH((k_spmm<T, V, LPR, CPL, U, COLMAJOR, false>), grid, SPMM_THREADS, 0, stream, args);
        return MXG_OK;
    }
    for (int q = 0; q < args.n_panels; q++) {
        args.panel = q;
        MXG_LAUNCH((k_spmm<T, V, LPR, CPL, U, COLMAJOR, true>), grid, SPMM_THREADS, 0, stream, args);
    }
    return MXG_OK;
}

template <typename T, int V, bool COLMAJOR>
static int dispatch_geom(int lpr, int cpl, SpmmArgs &args, cudaStream_t stream)
{
    int rpw = COLMAJOR ? SPMM_CM_RPW : (int)options().spmm_rpw;
    if (rpw <= 0) rpw = 8;
    if (rpw > 31) rpw = 31;
    // bulk copies ship exactly one column block per row: the row must be one block wide (and whole vectors)
    if (args.bulk && (COLMAJOR || V == 1 || args.n != lpr * V * cpl)) args.bulk = 0;
    if (args.bulk) rpw = spmm_bulk_rpw(lpr * V * cpl, (int)sizeof(T));
    args.rpw = rpw;
    args.piece_blocks = ceil_div_i(args.n_pieces, SPMM_WARPS);
    const int row_blocks = ceil_div_i(args.m, SPMM_WARPS * rpw);
    if (cpl == 2 && lpr < 32) {
        if constexpr (V > 1) {
            if (args.n_panels <= 1) {
                if (lpr == 4) return launch_two<T, V, 4, COLMAJOR>(args, row_blocks, stream);
                if (lpr == 8) return launch_two<T, V, 8, COLMAJOR>(args, row_blocks, stream);
                if (lpr == 16) return launch_two<T, V, 16, COLMAJOR>(args, row_blocks, stream);
            }
        }
        lpr *= 2; // column panels / scalar path: the one-vector geometry
        cpl = 1;
    }
#define MXG_GEOM(L, C)                                                                                   \
    if (lpr == L && cpl == C) return launch_variant<T, V, L, C, 4, COLMAJOR>(args, row_blocks, stream);
    MXG_GEOM(4, 1) MXG_GEOM(8, 1) MXG_GEOM(16, 1) MXG_GEOM(32, 1) MXG_GEOM(32, 2)
#undef MXG_GEOM
    return fail(MXG_ERR_ARG, "spmm: unsupported team geometry lpr=%d cpl=%d", lpr, cpl);
}

static int pow2_at_least(int v)
{
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

// Column panels pay when the dense operand overflows L2 (every stored entry then re-fetches its row of B from
// HBM) and the extra read-modify-write passes over the output cost less than those re-fetches:
//   saved  ~ nnz * row_bytes * (1 - L2 / B_bytes) / 2      extra ~ m * row_bytes * (2 P - 2)
// The split table is cached in the handle (one table per panel width).
static int plan_panels(mxg_csr_s *A, size_t b_bytes, cudaStream_t stream, SpmmArgs &args)
{
    const long forced_cols = options().spmm_panel_cols;
    long panel_mb = options().spmm_panel_mb;
    // Off unless asked for: measured on B200 (profiles/README.md, r01 v3 sweep) the read-modify-write passes and
    // the short per-panel row segments cost more than the saved HBM re-fetches of B at the BASELINE shapes
    // (cfg3 fp32 k=64: 3.24 ms without panels, 3.56 / 4.05 / 5.9 ms with 128 / 64 / 32 MiB panels).
    if (panel_mb <= 0 && forced_cols <= 0) return MXG_OK;
    if (A->K <= 1 || A->m <= 0 || A->nnz <= 0) return MXG_OK;
    int width;
    if (forced_cols > 0) {
        width = (int)std::min<long>(forced_cols, A->K);
    } else {
        const size_t panel_bytes = (size_t)panel_mb << 20;
        if (b_bytes <= panel_bytes + panel_bytes / 2) return MXG_OK; // fits L2 well enough
        int P = (int)((b_bytes + panel_bytes - 1) / panel_bytes);
        if (P > 32) return MXG_OK;
        const double saved = 0.5 * (double)A->nnz * (1.0 - (double)panel_bytes / (double)b_bytes);
        const double extra = (double)A->m * (2.0 * P - 2.0);
        if (extra >= saved) return MXG_OK;
        width = (A->K + P - 1) / P;
    }
    if (width < 1) width = 1;
    const int P = (A->K + width - 1) / width;
    if (P <= 1) return MXG_OK;
    if (P > 64) return fail(MXG_ERR_ARG, "spmm: more than 64 column panels requested");
    if (A->d_seg == nullptr || A->seg_width != width || A->seg_panels != P) {
        if (A->d_seg) MXG_CUDA_TRY(cudaFreeAsync(A->d_seg, stream));
        A->d_seg = nullptr;
        MXG_CUDA_TRY(cudaMallocAsync(&A->d_seg, sizeof(int32_t) * (size_t)(P - 1) * (size_t)A->m, stream));
        A->seg_width = width;
        A->seg_panels = P;
        MXG_LAUNCH(k_panel_segments, ceil_div_i(A->m, 256), 256, 0, stream, A->m, A->d_p, A->d_j, width, P, A->d_seg);
    }
    args.seg = A->d_seg;
    args.n_panels = P;
    args.panel_width = width;
    return MXG_OK;
}

template <typename T>
static int spmm_typed(const mxg_csr_s *A, const T *d_x, int out_layout, int n, const T *d_B, size_t ldb,
                      int n_dst, void *const *d_outs, size_t ldc, cudaStream_t stream, int mcast, int part = 0)
{
    // part 0: the whole product; 1: short rows only (row CTAs); 2: long rows only (piece CTAs + fix-up)
    constexpr int VEC = 16 / (int)sizeof(T);
    const bool colmajor = out_layout == MXG_COLS_CONTIGUOUS;
    const size_t out_rows = (size_t)A->m;
    T *d_Out = static_cast<T *>(d_outs[0]);

    if (A->nnz == 0) {
        // the reference returns its zero-filled matrix untouched (src/matmul.cpp:128-129, 160-161)
        for (int d = 0; d < n_dst; d++) {
            T *o = static_cast<T *>(d_outs[d]);
            if (colmajor) MXG_LAUNCH(k_fill_zero_2d<T>, 148 * 4, 256, 0, stream, o, (size_t)n, out_rows, ldc);
            else MXG_LAUNCH(k_fill_zero_2d<T>, 148 * 4, 256, 0, stream, o, out_rows, (size_t)n, ldc);
        }
        return MXG_OK;
    }

    // 128-bit path needs whole vectors and 16-byte aligned rows of B (and of every Out when row-major)
    bool vec = (n % VEC == 0) && (ldb % VEC == 0) && (((uintptr_t)d_B & 15) == 0);
    if (!colmajor) {
        vec = vec && (ldc % VEC == 0);
        for (int d = 0; d < n_dst; d++) vec = vec && (((uintptr_t)d_outs[d] & 15) == 0);
    }
    const int V = vec ? VEC : 1;
    const int nvec = n / V;

    // lanes per row: the smallest power of two covering the row of B (>= 4), two vectors per lane when a
    // row is wider than one warp; wider still => several column blocks (grid.y), re-reading the CSR
    int lpr = (int)options().spmm_lpr;
    if (lpr <= 0) {
        lpr = pow2_at_least(nvec);
        if (lpr < 4) lpr = 4;
        if (lpr > 32) lpr = 32;
    }
    int cpl = (lpr == 32 && nvec > 32) ? 2 : 1;
    // rows of B of more than 128 bytes: half-width teams with two vectors per lane (launch_two); option spmm_cpl
    // forces one (1) or two (2) vectors per lane for sweeps and tests
    const long want_cpl = options().spmm_cpl;
    if (vec && cpl == 1 && lpr >= 8 && want_cpl != 1 && (nvec > 8 || want_cpl == 2) && options().spmm_lpr <= 0) {
        lpr /= 2;
        cpl = 2;
    }

    SpmmArgs args;
    args.m = A->m;
    args.n = n;
    args.p = A->d_p;
    args.j = A->d_j;
    args.x = d_x;
    args.B = d_B;
    args.ldb = ldb;
    args.Out = d_Out;
    args.ldc = ldc;
    args.n_extra = n_dst - 1;
    args.mcast = mcast;
    // several destinations of a rows-contiguous result whose rows are stored back to back: bulk copies from shared memory
    args.bulk = (n_dst > 1 && !mcast && !colmajor && vec && ldc == (size_t)n && options().spmm_bulk != 0)
                    ? (options().spmm_bulk == 2 ? 2 : 1) : 0;
    for (int d = 0; d < MXG_MAX_DST - 1; d++) args.extra[d] = d + 1 < n_dst ? d_outs[d + 1] : nullptr;
    args.piece = A->piece;
    args.n_pieces = part == 1 ? 0 : A->n_pieces;
    if (part == 2) args.m = 0; // no row CTAs
    args.piece_row = A->d_piece_row;
    args.piece_k = A->d_piece_k;
    args.partial = nullptr;
    args.abort = A->d_abort;
    args.seg = nullptr;
    args.panel = 0;
    args.n_panels = 1;
    args.panel_width = A->K > 0 ? A->K : 1;
    if (n_dst == 1 && part == 0) MXG_TRY(plan_panels(const_cast<mxg_csr_s *>(A), (size_t)A->K * ldb * sizeof(T), stream, args));
    PartialLease partial; // lives until the fix-up launch below has been enqueued
    if (args.n_pieces > 0) {
        MXG_TRY(partial.acquire(A, (size_t)A->n_pieces * (size_t)n * sizeof(T), stream));
        args.partial = partial.ptr;
    }

    int rc;
    if (vec) {
        rc = colmajor ? dispatch_geom<T, VEC, true>(lpr, cpl, args, stream)
                      : dispatch_geom<T, VEC, false>(lpr, cpl, args, stream);
    } else {
        rc = colmajor ? dispatch_geom<T, 1, true>(lpr, cpl, args, stream)
                      : dispatch_geom<T, 1, false>(lpr, cpl, args, stream);
    }
    MXG_TRY(rc);

    if (A->n_long > 0 && part != 1) {
        DstList out;
        out.n = n_dst;
        for (int d = 0; d < MXG_MAX_DST; d++) out.dst[d] = d < n_dst ? d_outs[d] : nullptr;
        if (colmajor)
            MXG_LAUNCH((k_spmm_fixup<T, true>), A->n_long, 128, 0, stream, n, A->d_long_rows, A->d_long_first,
                       A->d_long_np, static_cast<const T *>(partial.ptr), out, ldc, A->d_abort, mcast);
        else
            MXG_LAUNCH((k_spmm_fixup<T, false>), A->n_long, 128, 0, stream, n, A->d_long_rows, A->d_long_first,
                       A->d_long_np, static_cast<const T *>(partial.ptr), out, ldc, A->d_abort, mcast);
    }
    return MXG_OK;
}

int launch_spmm(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                void *d_Out, size_t ldc, cudaStream_t stream)
{
    void *outs[1] = {d_Out};
    return launch_spmm_multi(A, dtype, out_layout, n, d_B, ldb, 1, outs, ldc, stream, 0);
}

// Rows [r0, r1) of the product without their long rows, or the long rows alone: how the warm host-buffer path
// (pipeline.cu: handle_spmm_host) lets result chunks leave the device while later rows are still being computed.
int launch_spmm_rows(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb, void *d_Out,
                     size_t ldc, int r0, int r1, int pieces, cudaStream_t stream)
{
    if (n <= 0 || A->nnz == 0) return fail(MXG_ERR_ARG, "spmm_rows: empty product");
    if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "spmm: bad out_layout %d", out_layout);
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    if (dtype != MXG_F64 && dtype != MXG_F32) return fail(MXG_ERR_ARG, "spmm: bad dtype %d", dtype);
    if ((dtype == MXG_F64 && !A->d_x64) || (dtype == MXG_F32 && !A->d_x32)) return fail(MXG_ERR_UNSUPPORTED, "spmm: handle lacks the values of this type");
    if (pieces) { // the handle itself: piece tables and the partial-sum workspace belong to it
        if (A->n_pieces == 0) return MXG_OK;
        void *outs[1] = {d_Out};
        if (dtype == MXG_F64)
            return spmm_typed<double>(A, A->d_x64, out_layout, n, static_cast<const double *>(d_B), ldb, 1, outs, ldc, stream, 0, 2);
        return spmm_typed<float>(A, A->d_x32, out_layout, n, static_cast<const float *>(d_B), ldb, 1, outs, ldc, stream, 0, 2);
    }
    if (r0 < 0 || r1 > A->m || r0 > r1) return fail(MXG_ERR_ARG, "spmm_rows: bad row range");
    if (r0 == r1) return MXG_OK;
    mxg_csr_s v = *A; // a view: same arrays, a window of the rows (offsets in d_p are absolute)
    v.m = r1 - r0;
    v.d_p = A->d_p + r0;
    void *outs[1] = {static_cast<char *>(d_Out) + (out_layout == MXG_ROWS_CONTIGUOUS ? (size_t)r0 * ldc * s : (size_t)r0 * s)};
    if (dtype == MXG_F64)
        return spmm_typed<double>(&v, A->d_x64, out_layout, n, static_cast<const double *>(d_B), ldb, 1, outs, ldc, stream, 0, 1);
    return spmm_typed<float>(&v, A->d_x32, out_layout, n, static_cast<const float *>(d_B), ldb, 1, outs, ldc, stream, 0, 1);
}

int launch_spmm_multi(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                      int n_dst, void *const *d_outs, size_t ldc, cudaStream_t stream, int mcast)
{
    if (mcast && (n_dst != 1 || out_layout != MXG_ROWS_CONTIGUOUS))
        return fail(MXG_ERR_UNSUPPORTED, "spmm: multicast stores take one rows-contiguous destination");
    if (mcast && A->nnz == 0) return fail(MXG_ERR_UNSUPPORTED, "spmm: multicast product of a matrix without stored entries");
    if (n_dst < 1 || n_dst > MXG_MAX_DST || !d_outs) return fail(MXG_ERR_ARG, "spmm: 1 .. %d destinations", MXG_MAX_DST);
    if (n < 0) return fail(MXG_ERR_ARG, "spmm: negative n");
    if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS)
        return fail(MXG_ERR_ARG, "spmm: bad out_layout %d", out_layout);
    if (A->m == 0 || n == 0) return MXG_OK;
    if (ldb < (size_t)n) return fail(MXG_ERR_ARG, "spmm: ldb (%zu) < n (%d)", ldb, n);
    if (out_layout == MXG_ROWS_CONTIGUOUS && ldc < (size_t)n) return fail(MXG_ERR_ARG, "spmm: ldc < n");
    if (out_layout == MXG_COLS_CONTIGUOUS && ldc < (size_t)A->m) return fail(MXG_ERR_ARG, "spmm: ldc < m");
    if (dtype == MXG_F64) {
        if (!A->d_x64 && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "spmm: handle holds no float64 values");
        return spmm_typed<double>(A, A->d_x64, out_layout, n, static_cast<const double *>(d_B), ldb, n_dst, d_outs, ldc, stream, mcast);
    }
    if (dtype == MXG_F32) {
        if (!A->d_x32 && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "spmm: handle holds no float32 values");
        return spmm_typed<float>(A, A->d_x32, out_layout, n, static_cast<const float *>(d_B), ldb, n_dst, d_outs, ldc, stream, mcast);
    }
    return fail(MXG_ERR_ARG, "spmm: bad dtype %d", dtype);
}

} // namespace mxg
