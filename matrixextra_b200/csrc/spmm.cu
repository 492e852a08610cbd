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

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// One team accumulates entries [a, b) of its row.  maxlen is the warp-wide maximum of (b - a).
template <typename T, int V, int LPR, int CPL>
__device__ __forceinline__ void team_gather(Pack<T, V> (&acc)[CPL], const int a, const int b, const int maxlen,
                                            const int l, const int32_t *__restrict__ j, const T *__restrict__ x,
                                            const T *__restrict__ B, const size_t ldb, const int (&col)[CPL],
                                            const bool (&cok)[CPL])
{
    constexpr int U = 4;
    static_assert(LPR % U == 0, "LPR must be a multiple of the gather unroll");
    for (int e0 = 0; e0 < maxlen; e0 += LPR) {
        const int e = a + e0 + l;
        int jj = 0;
        T xx = T(0);
        if (e < b) {
            jj = __ldg(j + e);
            xx = __ldg(x + e);
        }
        const int cnt = b - (a + e0);           // entries this team still has (may be <= 0)
        const int kmax = min(LPR, maxlen - e0); // warp-uniform
        for (int k = 0; k < kmax; k += U) {
            int jk[U];
            T xk[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                jk[u] = __shfl_sync(0xffffffffu, jj, k + u, LPR);
                xk[u] = __shfl_sync(0xffffffffu, xx, k + u, LPR);
                ok[u] = (k + u) < cnt;
            }
            Pack<T, V> bv[U][CPL];
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    if (ok[u] && cok[c]) bv[u][c] = ld_ro<T, V>(B + (size_t)jk[u] * ldb + col[c]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    if (ok[u] && cok[c]) {
#pragma unroll
                        for (int i = 0; i < V; i++) acc[c].v[i] = fma_t(xk[u], bv[u][c].v[i], acc[c].v[i]);
                    }
                }
            }
        }
    }
}

template <int LPR>
struct TeamGeom {
    static constexpr int THREADS = 256;
    static constexpr int TEAMS = THREADS / LPR;
    static constexpr int BR = TEAMS > 32 ? TEAMS : 32; // rows of a column-major CTA tile
};

struct SpmmArgs {
    int m, n;
    const int32_t *p;
    const int32_t *j;
    const void *x;
    const void *B;
    size_t ldb;
    void *Out;
    size_t ldc;
    int block_rows; // rows per CTA (row-major output); column-major uses TeamGeom::BR
    int piece;
    int n_pieces;
    int piece_blocks;
    const int32_t *piece_row;
    const int32_t *piece_k;
    void *partial; // [n_pieces][n]
};

template <typename T, int V, int LPR, int CPL, bool COLMAJOR>
__global__ void __launch_bounds__(256) k_spmm(const SpmmArgs g)
{
    constexpr int TEAMS = TeamGeom<LPR>::TEAMS;
    constexpr int BR = TeamGeom<LPR>::BR;
    constexpr int NB = LPR * V * CPL; // output columns per CTA column block
    constexpr int TLD = BR + 1;
    __shared__ T tile[COLMAJOR ? NB * TLD : 1];

    const int32_t *__restrict__ p = g.p;
    const int32_t *__restrict__ j = g.j;
    const T *__restrict__ x = static_cast<const T *>(g.x);
    const T *__restrict__ B = static_cast<const T *>(g.B);
    T *__restrict__ Out = static_cast<T *>(g.Out);

    const int lane = threadIdx.x & 31;
    const int team = threadIdx.x / LPR;
    const int l = threadIdx.x % LPR;
    const int col0 = blockIdx.y * NB;

    int col[CPL];
    bool cok[CPL];
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        col[c] = col0 + (c * LPR + l) * V;
        cok[c] = col[c] < g.n; // n % V == 0 on the vector path, so a vector is entirely in or out
    }

    if ((int)blockIdx.x < g.piece_blocks) {
        // ---- long-row pieces: partial sums to the workspace -------------------------------------
        const int pc = blockIdx.x * TEAMS + team;
        int a = 0, b = 0;
        if (pc < g.n_pieces) {
            const int row = g.piece_row[pc];
            const int r0 = p[row], r1 = p[row + 1];
            a = r0 + g.piece_k[pc] * g.piece;
            b = min(a + g.piece, r1);
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, b - a);
        Pack<T, V> acc[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++)
#pragma unroll
            for (int i = 0; i < V; i++) acc[c].v[i] = T(0);
        team_gather<T, V, LPR, CPL>(acc, a, b, maxlen, l, j, x, B, g.ldb, col, cok);
        if (pc < g.n_pieces) {
            T *dst = static_cast<T *>(g.partial) + (size_t)pc * g.n;
#pragma unroll
            for (int c = 0; c < CPL; c++)
                if (cok[c]) st_pack<T, V>(dst + col[c], acc[c]);
        }
        return;
    }

    const int rb = blockIdx.x - g.piece_blocks;
    const int block_rows = COLMAJOR ? BR : g.block_rows;
    const int row0 = rb * block_rows;

    for (int base = 0; base < block_rows; base += TEAMS) {
        const int rl = base + team;
        const int row = row0 + rl;
        int a = 0, b = 0;
        bool store = false;
        if (rl < block_rows && row < g.m) {
            a = p[row];
            b = p[row + 1];
            store = true;
            if (b - a > g.piece) { // long row: handled by the piece section + fix-up
                b = a;
                store = false;
            }
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, b - a);
        Pack<T, V> acc[CPL];
#pragma unroll
        for (int c = 0; c < CPL; c++)
#pragma unroll
            for (int i = 0; i < V; i++) acc[c].v[i] = T(0);
        team_gather<T, V, LPR, CPL>(acc, a, b, maxlen, l, j, x, B, g.ldb, col, cok);
        if (store) {
            if (COLMAJOR) {
#pragma unroll
                for (int c = 0; c < CPL; c++)
                    if (cok[c]) {
#pragma unroll
                        for (int i = 0; i < V; i++) tile[(col[c] - col0 + i) * TLD + rl] = acc[c].v[i];
                    }
            } else {
                T *dst = Out + (size_t)row * g.ldc;
#pragma unroll
                for (int c = 0; c < CPL; c++)
                    if (cok[c]) st_pack<T, V>(dst + col[c], acc[c]);
            }
        }
    }

    if (COLMAJOR) {
        __syncthreads();
        const int warp = threadIdx.x >> 5;
        const int ncols = min(NB, g.n - col0);
#pragma unroll
        for (int rr = 0; rr < BR; rr += 32) {
            const int rl = rr + lane;
            const int row = row0 + rl;
            bool w = row < g.m;
            if (w) w = (p[row + 1] - p[row]) <= g.piece; // long rows are written by the fix-up kernel
            if (w) {
                T *dst = Out + (size_t)row;
                for (int c = warp; c < ncols; c += 8) dst[(size_t)(col0 + c) * g.ldc] = tile[c * TLD + rl];
            }
        }
    }
}

// Adds the pieces of every long row in piece order and writes the row.
template <typename T, bool COLMAJOR>
__global__ void __launch_bounds__(128) k_spmm_fixup(int n, const int32_t *__restrict__ long_rows,
                                                    const int32_t *__restrict__ long_first,
                                                    const int32_t *__restrict__ long_np,
                                                    const T *__restrict__ partial, T *__restrict__ Out, size_t ldc)
{
    const int row = long_rows[blockIdx.x];
    const int first = long_first[blockIdx.x];
    const int np = long_np[blockIdx.x];
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        T s = T(0);
        for (int k = 0; k < np; k++) s += partial[(size_t)(first + k) * n + c];
        if (COLMAJOR) Out[(size_t)row + (size_t)c * ldc] = s;
        else Out[(size_t)row * ldc + c] = s;
    }
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

template <typename T, int V, int LPR, int CPL, bool COLMAJOR>
static int launch_variant(const SpmmArgs &args, int row_blocks, cudaStream_t stream)
{
    constexpr int NB = LPR * V * CPL;
    dim3 grid((unsigned)(args.piece_blocks + row_blocks), (unsigned)ceil_div_i(args.n, NB), 1);
    MXG_LAUNCH((k_spmm<T, V, LPR, CPL, COLMAJOR>), grid, 256, 0, stream, args);
    return MXG_OK;
}

template <typename T, int V, bool COLMAJOR>
static int dispatch_geom(int lpr, int cpl, SpmmArgs &args, cudaStream_t stream)
{
    const int teams = 256 / lpr;
    int block_rows;
    if (COLMAJOR) {
        block_rows = teams > 32 ? teams : 32;
    } else {
        block_rows = (int)options().spmm_block_rows;
        if (block_rows <= 0) block_rows = teams * 4;
        block_rows = ((block_rows + teams - 1) / teams) * teams;
    }
    args.block_rows = block_rows;
    args.piece_blocks = ceil_div_i(args.n_pieces, teams);
    const int row_blocks = ceil_div_i(args.m, block_rows);
#define MXG_GEOM(L, C)                                                              \
    if (lpr == L && cpl == C) return launch_variant<T, V, L, C, COLMAJOR>(args, row_blocks, stream);
    MXG_GEOM(4, 1) MXG_GEOM(4, 2) MXG_GEOM(8, 1) MXG_GEOM(8, 2)
    MXG_GEOM(16, 1) MXG_GEOM(16, 2) MXG_GEOM(32, 1) MXG_GEOM(32, 2)
#undef MXG_GEOM
    return fail(MXG_ERR_ARG, "spmm: unsupported team geometry lpr=%d cpl=%d", lpr, cpl);
}

static int pow2_at_least(int v)
{
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

template <typename T>
static int spmm_typed(const mxg_csr_s *A, const T *d_x, int out_layout, int n, const T *d_B, size_t ldb,
                      T *d_Out, size_t ldc, cudaStream_t stream)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    const bool colmajor = out_layout == MXG_COLS_CONTIGUOUS;
    const size_t out_rows = (size_t)A->m;

    if (A->nnz == 0) {
        // the reference returns its zero-filled matrix untouched (src/matmul.cpp:128-129, 160-161)
        if (colmajor) MXG_LAUNCH(k_fill_zero_2d<T>, 148 * 4, 256, 0, stream, d_Out, (size_t)n, out_rows, ldc);
        else MXG_LAUNCH(k_fill_zero_2d<T>, 148 * 4, 256, 0, stream, d_Out, out_rows, (size_t)n, ldc);
        return MXG_OK;
    }

    // 128-bit path needs whole vectors and 16-byte aligned rows of B (and of Out when row-major)
    bool vec = (n % VEC == 0) && (ldb % VEC == 0) && (((uintptr_t)d_B & 15) == 0);
    if (!colmajor) vec = vec && (ldc % VEC == 0) && (((uintptr_t)d_Out & 15) == 0);
    const int V = vec ? VEC : 1;
    const int nvec = n / V;

    int lpr = (int)options().spmm_lpr;
    int cpl = (int)options().spmm_cpl;
    if (lpr <= 0) {
        lpr = pow2_at_least(nvec);
        if (lpr < 4) lpr = 4;
        if (lpr > 32) lpr = 32;
    }
    if (cpl <= 0) cpl = (nvec > lpr) ? 2 : 1;
    if (cpl > 2) cpl = 2;

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
    args.piece = A->piece;
    args.n_pieces = A->n_pieces;
    args.piece_row = A->d_piece_row;
    args.piece_k = A->d_piece_k;
    args.partial = nullptr;
    if (A->n_pieces > 0) {
        MXG_TRY(ensure_partial(const_cast<mxg_csr_s *>(A), (size_t)A->n_pieces * (size_t)n * sizeof(T)));
        args.partial = A->d_partial;
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

    if (A->n_long > 0) {
        if (colmajor)
            MXG_LAUNCH((k_spmm_fixup<T, true>), A->n_long, 128, 0, stream, n, A->d_long_rows, A->d_long_first,
                       A->d_long_np, static_cast<const T *>(A->d_partial), d_Out, ldc);
        else
            MXG_LAUNCH((k_spmm_fixup<T, false>), A->n_long, 128, 0, stream, n, A->d_long_rows, A->d_long_first,
                       A->d_long_np, static_cast<const T *>(A->d_partial), d_Out, ldc);
    }
    return MXG_OK;
}

int launch_spmm(const mxg_csr_s *A, int dtype, int out_layout, int n, const void *d_B, size_t ldb,
                void *d_Out, size_t ldc, cudaStream_t stream)
{
    if (n < 0) return fail(MXG_ERR_ARG, "spmm: negative n");
    if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS)
        return fail(MXG_ERR_ARG, "spmm: bad out_layout %d", out_layout);
    if (A->m == 0 || n == 0) return MXG_OK;
    if (ldb < (size_t)n) return fail(MXG_ERR_ARG, "spmm: ldb (%zu) < n (%d)", ldb, n);
    if (out_layout == MXG_ROWS_CONTIGUOUS && ldc < (size_t)n) return fail(MXG_ERR_ARG, "spmm: ldc < n");
    if (out_layout == MXG_COLS_CONTIGUOUS && ldc < (size_t)A->m) return fail(MXG_ERR_ARG, "spmm: ldc < m");
    if (dtype == MXG_F64) {
        if (!A->d_x64 && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "spmm: handle holds no float64 values");
        return spmm_typed<double>(A, A->d_x64, out_layout, n, static_cast<const double *>(d_B), ldb,
                                  static_cast<double *>(d_Out), ldc, stream);
    }
    if (dtype == MXG_F32) {
        if (!A->d_x32 && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "spmm: handle holds no float32 values");
        return spmm_typed<float>(A, A->d_x32, out_layout, n, static_cast<const float *>(d_B), ldb,
                                 static_cast<float *>(d_Out), ldc, stream);
    }
    return fail(MXG_ERR_ARG, "spmm: bad dtype %d", dtype);
}

} // namespace mxg
