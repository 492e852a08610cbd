// spmv.cu — K3: CSR x dense vector for sm_100a.  Replaces matmul_csr_dvec<> (src/matmul.cpp:381-419)
// and its four exports (421-483), including the NA rules for integer / logical vectors (406-411).
//
// A team of LPR lanes owns a row; its lanes stride over the row's stored entries (coalesced index and
// value loads, U=4 independent gathers of y in flight per lane), then a width-LPR butterfly adds the
// lane sums.  Teams of a warp run in lock-step over the longest row.  Rows longer than `piece`
// entries use the same piece tables as the SpMM kernels: one warp per piece writes a partial sum and
// a fix-up pass adds a row's pieces in piece order (deterministic, no atomics).
// The sum is a tree over lanes instead of the reference's left-to-right chain: identical up to
// reassociation (<= 1e-12 relative for fp64, see tests).  Accumulation is always in double; the
// float32 variant narrows once at the end (the reference narrows after every term, 403/476).
// Bound: HBM for the streamed CSR (12 B per entry) plus one 32-byte L2 sector per gathered y element.
#include "mxg_internal.cuh"

#include <limits.h>
#include <algorithm>
#include <mutex>
#include <vector>

namespace mxg {

__device__ __forceinline__ double na_real()
{
    return __longlong_as_double(0x7FF00000000007A2LL); // R's NA_real_ (payload 1954)
}

template <int YTYPE>
struct YTraits;
template <>
struct YTraits<MXG_Y_NUMERIC> {
    typedef double elem;
    typedef double out;
    static __device__ __forceinline__ double value(double v, bool &na) { return v; }
};
template <>
struct YTraits<MXG_Y_INTEGER> {
    typedef int elem;
    typedef double out;
    static __device__ __forceinline__ double value(int v, bool &na)
    {
        if (v == INT_MIN) { na = true; return na_real(); }
        return (double)v;
    }
};
template <>
struct YTraits<MXG_Y_LOGICAL> {
    typedef int elem;
    typedef double out;
    static __device__ __forceinline__ double value(int v, bool &na)
    {
        if (v == INT_MIN) { na = true; return na_real(); }
        return v != 0 ? 1.0 : 0.0;
    }
};
template <>
struct YTraits<MXG_Y_FLOAT32> {
    typedef float elem;
    typedef float out;
    static __device__ __forceinline__ double value(float v, bool &na) { return (double)v; }
};

// lane-strided partial dot product of entries [a, b); maxlen = warp-wide max of (b - a)
template <int YTYPE, typename XT, int LPR>
__device__ __forceinline__ double team_dot(const int a, const int b, const int maxlen, const int l,
                                           const int32_t *__restrict__ j, const XT *__restrict__ x,
                                           const typename YTraits<YTYPE>::elem *__restrict__ y, bool &na,
                                           const cudaTextureObject_t tex = 0)
{
    constexpr int U = 4;
    double acc = 0.0;
    for (int e0 = 0; e0 < maxlen; e0 += LPR * U) {
        int jj[U];
        double xx[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = a + e0 + u * LPR + l;
            ok[u] = e < b;
            jj[u] = 0;
            xx[u] = 0.0;
            if (ok[u]) {
                jj[u] = __ldg(j + e);
                xx[u] = (double)__ldg(x + e);
            }
        }
        double yy[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            yy[u] = 0.0;
            if (ok[u]) {
                if (YTYPE == MXG_Y_NUMERIC && tex != 0) {
                    const int2 t = tex1Dfetch<int2>(tex, jj[u]);
                    yy[u] = __hiloint2double(t.y, t.x);
                } else {
                    yy[u] = YTraits<YTYPE>::value(__ldg(y + jj[u]), na);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) acc = fma(xx[u], yy[u], acc);
    }
    return acc;
}

template <int LPR>
__device__ __forceinline__ double team_reduce(double v)
{
#pragma unroll
    for (int d = LPR / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d, LPR);
    return v;
}

template <int LPR>
__device__ __forceinline__ bool team_any(bool f)
{
    int v = f ? 1 : 0;
#pragma unroll
    for (int d = LPR / 2; d >= 1; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d, LPR);
    return v != 0;
}

// result copies beyond the local one (peer-mapped vectors of the other GPUs of the box)
struct SpmvExtra {
    void *dst[MXG_MAX_DST - 1];
    int n;
};

template <int YTYPE>
__device__ __forceinline__ void store_result(typename YTraits<YTYPE>::out *out, const SpmvExtra &extra, int row, double v, bool na)
{
    // a row that met an NA keeps R's NA payload: x86 propagates NA_real_ through the reference's sum as
    // the quieted NaN 0x7FF80000000007A2 (low word still 1954, so R's is.na() holds)
    if (na && v != v) v = __longlong_as_double(0x7FF80000000007A2LL);
    typedef typename YTraits<YTYPE>::out OE;
    out[row] = (OE)v;
    for (int d = 0; d < extra.n; d++) static_cast<OE *>(extra.dst[d])[row] = (OE)v;
}

struct SpmvArgs {
    int m;
    const int32_t *p;
    const int32_t *j;
    const void *x;
    const void *y;
    void *out;
    int piece;
    int n_pieces;
    int piece_blocks;
    const int32_t *piece_row;
    const int32_t *piece_k;
    double *partial;   // [n_pieces]
    int *partial_na;   // [n_pieces]
    cudaTextureObject_t tex; // numeric y: gather through the texture path (fewer L1 wavefronts per warp gather); 0 = plain loads
    int rows_per_team; // rows each team walks inside its CTA
    const int *abort;  // optional device flag: non-zero => the column ids failed validation, do nothing
    SpmvExtra extra;
};

template <int YTYPE, typename XT, int LPR>
__global__ void __launch_bounds__(256) k_spmv(const SpmvArgs g)
{
    typedef typename YTraits<YTYPE>::elem YE;
    typedef typename YTraits<YTYPE>::out OE;
    constexpr int TEAMS = 256 / LPR;
    const int32_t *__restrict__ p = g.p;
    const int32_t *__restrict__ j = g.j;
    const XT *__restrict__ x = static_cast<const XT *>(g.x);
    const YE *__restrict__ y = static_cast<const YE *>(g.y);
    OE *__restrict__ out = static_cast<OE *>(g.out);

    if (g.abort != nullptr && *g.abort != 0) return;

    if ((int)blockIdx.x < g.piece_blocks) {
        // one full warp per long-row piece
        const int lane = threadIdx.x & 31;
        const int pc = blockIdx.x * 8 + (threadIdx.x >> 5);
        if (pc >= g.n_pieces) return; // warp-uniform
        const int row = g.piece_row[pc];
        const int a = p[row] + g.piece_k[pc] * g.piece;
        const int b = min(a + g.piece, p[row + 1]);
        bool na = false;
        double acc = team_dot<YTYPE, XT, 32>(a, b, b - a, lane, j, x, y, na, g.tex);
        acc = team_reduce<32>(acc);
        na = team_any<32>(na);
        if (lane == 0) {
            g.partial[pc] = acc;
            g.partial_na[pc] = na ? 1 : 0;
        }
        return;
    }

    const int team = threadIdx.x / LPR;
    const int l = threadIdx.x % LPR;
    const int rb = blockIdx.x - g.piece_blocks;
    const int block_rows = TEAMS * g.rows_per_team;
    const int row0 = rb * block_rows;
    for (int base = 0; base < block_rows; base += TEAMS) {
        const int row = row0 + base + team;
        int a = 0, b = 0;
        bool store = false;
        if (row < g.m) {
            a = p[row];
            b = p[row + 1];
            store = true;
            if (b - a > g.piece) {
                b = a;
                store = false;
            }
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, b - a);
        bool na = false;
        double acc = team_dot<YTYPE, XT, LPR>(a, b, maxlen, l, j, x, y, na, g.tex);
        acc = team_reduce<LPR>(acc);
        na = team_any<LPR>(na);
        if (store && l == 0) store_result<YTYPE>(out, g.extra, row, acc, na);
    }
}

template <int YTYPE>
__global__ void __launch_bounds__(128) k_spmv_fixup(int n_long, const int32_t *__restrict__ long_rows,
                                                    const int32_t *__restrict__ long_first,
                                                    const int32_t *__restrict__ long_np,
                                                    const double *__restrict__ partial,
                                                    const int *__restrict__ partial_na,
                                                    typename YTraits<YTYPE>::out *__restrict__ out, const SpmvExtra extra,
                                                    const int *__restrict__ abort)
{
    if (abort != nullptr && *abort != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    const int first = long_first[i], np = long_np[i];
    double s = 0.0;
    bool na = false;
    for (int k = 0; k < np; k++) {
        s += partial[first + k];
        na = na || (partial_na[first + k] != 0);
    }
    store_result<YTYPE>(out, extra, long_rows[i], s, na);
}

// Texture objects over the dense vector y.  A kernel argument carries only the 64-bit handle of the object, so the
// object has to outlive every launch that fetches through it: objects are cached per (device, address, size) and
// each carries an event recorded behind its most recent launch.  A repeated y (power iteration, L-BFGS) re-uses its
// object; an entry is destroyed only when it is evicted AND its event has completed (the oldest one is waited for
// when all 16 are in flight).  mxg_trim() empties the cache.
namespace {
struct TexEntry {
    int device = -1;
    const void *ptr = nullptr;
    size_t bytes = 0;
    cudaTextureObject_t tex = 0;
    cudaEvent_t last_use = nullptr;
    bool used = false; // last_use has been recorded
    int leases = 0;    // launches being enqueued through it right now
    unsigned long long stamp = 0;
};
std::mutex g_tex_mu;
std::vector<TexEntry> g_tex;
unsigned long long g_tex_clock = 0;
constexpr size_t TEX_CACHE = 16;

void tex_destroy(TexEntry &e)
{
    if (e.used) cudaEventSynchronize(e.last_use);
    if (e.tex) cudaDestroyTextureObject(e.tex);
    if (e.last_use) cudaEventDestroy(e.last_use);
    e = TexEntry();
}
} // namespace

void texture_cache_clear()
{
    std::lock_guard<std::mutex> lk(g_tex_mu);
    for (TexEntry &e : g_tex) tex_destroy(e);
    g_tex.clear();
}

struct TexLease {
    cudaTextureObject_t tex = 0;
    int slot = -1;
    int acquire(const void *d_y, size_t bytes)
    {
        int dev = 0;
        MXG_CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(g_tex_mu);
        for (size_t i = 0; i < g_tex.size(); i++) {
            TexEntry &e = g_tex[i];
            if (e.tex && e.device == dev && e.ptr == d_y && e.bytes == bytes) {
                e.leases++;
                e.stamp = ++g_tex_clock;
                tex = e.tex;
                slot = (int)i;
                return MXG_OK;
            }
        }
        // a free slot, else the least recently used entry nobody is launching through (finished ones first)
        int pick = -1;
        for (size_t i = 0; i < g_tex.size() && pick < 0; i++)
            if (!g_tex[i].tex) pick = (int)i;
        if (pick < 0 && g_tex.size() < TEX_CACHE) {
            g_tex.emplace_back();
            pick = (int)g_tex.size() - 1;
        }
        if (pick < 0) {
            int oldest = -1, oldest_done = -1;
            for (size_t i = 0; i < g_tex.size(); i++) {
                TexEntry &e = g_tex[i];
                if (e.leases > 0) continue;
                if (oldest < 0 || e.stamp < g_tex[(size_t)oldest].stamp) oldest = (int)i;
                const bool done = !e.used || cudaEventQuery(e.last_use) == cudaSuccess;
                if (done && (oldest_done < 0 || e.stamp < g_tex[(size_t)oldest_done].stamp)) oldest_done = (int)i;
            }
            cudaGetLastError(); // cudaErrorNotReady is an answer, not an error
            pick = oldest_done >= 0 ? oldest_done : oldest;
            if (pick < 0) return MXG_OK; // every entry is mid-launch on other threads: this product uses plain loads
            tex_destroy(g_tex[(size_t)pick]); // waits for its last launch when that is still running
        }
        TexEntry &e = g_tex[(size_t)pick];
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = const_cast<void *>(d_y);
        rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindSigned);
        rd.res.linear.sizeInBytes = bytes;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        MXG_CUDA_TRY(cudaCreateTextureObject(&e.tex, &rd, &td, nullptr));
        if (cudaEventCreateWithFlags(&e.last_use, cudaEventDisableTiming) != cudaSuccess) {
            cudaDestroyTextureObject(e.tex);
            e = TexEntry();
            return fail(MXG_ERR_CUDA, "spmv: cudaEventCreate failed");
        }
        e.device = dev;
        e.ptr = d_y;
        e.bytes = bytes;
        e.leases = 1;
        e.stamp = ++g_tex_clock;
        tex = e.tex;
        slot = pick;
        return MXG_OK;
    }
    // after the launch (or instead of it, on an error path): mark the object busy until `stream` gets here
    int release(cudaStream_t stream)
    {
        if (slot < 0) return MXG_OK;
        std::lock_guard<std::mutex> lk(g_tex_mu);
        TexEntry &e = g_tex[(size_t)slot];
        slot = -1;
        e.leases--;
        MXG_CUDA_TRY(cudaEventRecord(e.last_use, stream));
        e.used = true;
        return MXG_OK;
    }
    ~TexLease()
    {
        if (slot < 0) return; // released
        std::lock_guard<std::mutex> lk(g_tex_mu);
        g_tex[(size_t)slot].leases--; // error path before any launch: nothing is in flight through this lease
    }
};

template <int YTYPE, typename XT>
static int spmv_dispatch(const mxg_csr_s *A, const XT *d_x, const void *d_y, int n_dst, void *const *d_outs, cudaStream_t stream)
{
    void *d_out = d_outs[0];
    SpmvArgs args;
    args.extra.n = n_dst - 1;
    for (int d = 0; d < MXG_MAX_DST - 1; d++) args.extra.dst[d] = d + 1 < n_dst ? d_outs[d + 1] : nullptr;
    args.m = A->m;
    args.p = A->d_p;
    args.j = A->d_j;
    args.x = d_x;
    args.y = d_y;
    args.out = d_out;
    args.piece = A->piece;
    args.n_pieces = A->n_pieces;
    args.piece_blocks = ceil_div_i(A->n_pieces, 8);
    args.piece_row = A->d_piece_row;
    args.piece_k = A->d_piece_k;
    args.partial = nullptr;
    args.partial_na = nullptr;
    args.abort = A->d_abort;
    PartialLease partial; // lives until the fix-up launch below has been enqueued
    if (A->n_pieces > 0) {
        // doubles first, flags after (16 bytes per piece reserved)
        MXG_TRY(partial.acquire(A, (size_t)A->n_pieces * 16, stream));
        args.partial = static_cast<double *>(partial.ptr);
        args.partial_na = reinterpret_cast<int *>(static_cast<double *>(partial.ptr) + A->n_pieces);
    }

    int lpr = (int)options().spmv_lpr;
    if (lpr <= 0) {
        // team size from the mean row length: enough entries per lane to amortise the butterfly
        const double mean = A->m > 0 ? (double)A->nnz / (double)A->m : 0.0;
        if (mean <= 6) lpr = 2;
        else if (mean <= 12) lpr = 4;
        else if (mean <= 48) lpr = 8;
        else if (mean <= 192) lpr = 16;
        else lpr = 32;
    }
    args.rows_per_team = 4;
    args.tex = 0;
    TexLease lease;
    if (YTYPE == MXG_Y_NUMERIC && options().spmv_tex != 0 && A->K > 0 && A->K <= (1 << 27) && (((uintptr_t)d_y & 255) == 0)) {
        MXG_TRY(lease.acquire(d_y, (size_t)A->K * 8));
        args.tex = lease.tex;
    }
#define MXG_SPMV(L)                                                                              \
    if (lpr == L) {                                                                              \
        const int block_rows = (256 / L) * args.rows_per_team;                                   \
        const int grid = args.piece_blocks + ceil_div_i(A->m, block_rows);                       \
        MXG_LAUNCH((k_spmv<YTYPE, XT, L>), grid, 256, 0, stream, args);                          \
    } else
    MXG_SPMV(2) MXG_SPMV(4) MXG_SPMV(8) MXG_SPMV(16) MXG_SPMV(32)
    {
        return fail(MXG_ERR_ARG, "spmv: unsupported team size %d", lpr);
    }
#undef MXG_SPMV
    MXG_TRY(lease.release(stream)); // the object stays alive (cached) until the kernel that fetches through it has run
    if (A->n_long > 0) {
        MXG_LAUNCH((k_spmv_fixup<YTYPE>), ceil_div_i(A->n_long, 128), 128, 0, stream, A->n_long, A->d_long_rows,
                   A->d_long_first, A->d_long_np, args.partial, args.partial_na,
                   static_cast<typename YTraits<YTYPE>::out *>(d_out), args.extra, A->d_abort);
    }
    return MXG_OK;
}

template <int YTYPE>
static int spmv_values(const mxg_csr_s *A, const void *d_y, int n_dst, void *const *d_outs, cudaStream_t stream)
{
    if (A->d_x64) return spmv_dispatch<YTYPE, double>(A, A->d_x64, d_y, n_dst, d_outs, stream);
    if (A->d_x32 && YTYPE == MXG_Y_FLOAT32) return spmv_dispatch<YTYPE, float>(A, A->d_x32, d_y, n_dst, d_outs, stream);
    if (A->nnz == 0) return spmv_dispatch<YTYPE, double>(A, nullptr, d_y, n_dst, d_outs, stream);
    return fail(MXG_ERR_UNSUPPORTED, "spmv: handle holds no float64 values");
}

int launch_spmv(const mxg_csr_s *A, int ytype, const void *d_y, void *d_out, cudaStream_t stream)
{
    void *outs[1] = {d_out};
    return launch_spmv_multi(A, ytype, d_y, 1, outs, stream);
}

int launch_spmv_multi(const mxg_csr_s *A, int ytype, const void *d_y, int n_dst, void *const *d_outs, cudaStream_t stream)
{
    if (n_dst < 1 || n_dst > MXG_MAX_DST || !d_outs) return fail(MXG_ERR_ARG, "spmv: 1 .. %d destinations", MXG_MAX_DST);
    if (A->m == 0) return MXG_OK;
    switch (ytype) {
    case MXG_Y_NUMERIC: return spmv_values<MXG_Y_NUMERIC>(A, d_y, n_dst, d_outs, stream);
    case MXG_Y_INTEGER: return spmv_values<MXG_Y_INTEGER>(A, d_y, n_dst, d_outs, stream);
    case MXG_Y_LOGICAL: return spmv_values<MXG_Y_LOGICAL>(A, d_y, n_dst, d_outs, stream);
    case MXG_Y_FLOAT32: return spmv_values<MXG_Y_FLOAT32>(A, d_y, n_dst, d_outs, stream);
    default: return fail(MXG_ERR_ARG, "spmv: bad ytype %d", ytype);
    }
}


// ------------------------------------------------------------------------------------------------
// Measurement probe (mxg_dev_spmv_probe): the access pattern of K3 with the row structure taken away.  Every
// thread streams 4 consecutive column ids of the handle per step (one 16-byte load, U steps in flight) and,
// depending on `mode`, gathers y[j] for each of them — no indptr, no per-row reduction, no result vector.
//   mode 0: ids only (the streaming floor of the 4-byte index array)
//   mode 1: ids + 8-byte gathers of y through plain loads          mode 2: ... through the texture path
//   mode 3: ids + values (12 streamed bytes per entry) + gathers through the texture path + FMA
// The time of mode 2 / 3 is what one L1 wavefront + one 32-byte L2 sector per stored entry costs on this matrix:
// the floor K3 is measured against.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_spmv_probe(size_t nnz4, const int4 *__restrict__ j4, const double2 *__restrict__ x2,
                                                    const double *__restrict__ y, cudaTextureObject_t tex, double *sink)
{
    constexpr int U = 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz4; e += stride * U) {
        int4 jj[U];
        double2 xa[U], xb[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t q = e + (size_t)u * stride;
            jj[u] = q < nnz4 ? __ldcs(j4 + q) : make_int4(0, 0, 0, 0);
            if (MODE == 3) {
                xa[u] = q < nnz4 ? __ldcs(x2 + 2 * q) : make_double2(0.0, 0.0);
                xb[u] = q < nnz4 ? __ldcs(x2 + 2 * q + 1) : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (MODE == 0) {
                acc += (double)(jj[u].x ^ jj[u].y ^ jj[u].z ^ jj[u].w);
            } else if (MODE == 1) {
                acc += __ldg(y + jj[u].x) + __ldg(y + jj[u].y) + __ldg(y + jj[u].z) + __ldg(y + jj[u].w);
            } else {
                const int2 a = tex1Dfetch<int2>(tex, jj[u].x), b = tex1Dfetch<int2>(tex, jj[u].y);
                const int2 c = tex1Dfetch<int2>(tex, jj[u].z), d = tex1Dfetch<int2>(tex, jj[u].w);
                const double ya = __hiloint2double(a.y, a.x), yb = __hiloint2double(b.y, b.x);
                const double yc = __hiloint2double(c.y, c.x), yd = __hiloint2double(d.y, d.x);
                if (MODE == 3) acc = fma(xa[u].x, ya, fma(xa[u].y, yb, fma(xb[u].x, yc, fma(xb[u].y, yd, acc))));
                else acc += ya + yb + yc + yd;
            }
        }
    }
    if (acc == 123.456) sink[0] = acc; // never true: keeps the loads alive
}

int spmv_probe(const mxg_csr_s *A, int mode, const double *d_y, double *d_sink, cudaStream_t stream)
{
    if (mode < 0 || mode > 3) return fail(MXG_ERR_ARG, "spmv_probe: mode 0 .. 3");
    if (A->nnz < 4 || A->base != 0) return fail(MXG_ERR_ARG, "spmv_probe: needs an owned handle with stored entries");
    if (mode == 3 && !A->d_x64) return fail(MXG_ERR_UNSUPPORTED, "spmv_probe: handle holds no float64 values");
    if (mode >= 1 && (!d_y || A->K <= 0)) return fail(MXG_ERR_ARG, "spmv_probe: NULL vector");
    const size_t nnz4 = (size_t)A->nnz / 4;
    const int4 *j4 = reinterpret_cast<const int4 *>(A->d_j);
    const double2 *x2 = reinterpret_cast<const double2 *>(A->d_x64);
    TexLease lease;
    if (mode >= 2) {
        MXG_TRY(lease.acquire(d_y, (size_t)A->K * 8));
        if (!lease.tex) return fail(MXG_ERR_CUDA, "spmv_probe: no texture object");
    }
    const int grid = 148 * 16;
    switch (mode) {
    case 0: MXG_LAUNCH(k_spmv_probe<0>, grid, 256, 0, stream, nnz4, j4, x2, d_y, lease.tex, d_sink); break;
    case 1: MXG_LAUNCH(k_spmv_probe<1>, grid, 256, 0, stream, nnz4, j4, x2, d_y, lease.tex, d_sink); break;
    case 2: MXG_LAUNCH(k_spmv_probe<2>, grid, 256, 0, stream, nnz4, j4, x2, d_y, lease.tex, d_sink); break;
    default: MXG_LAUNCH(k_spmv_probe<3>, grid, 256, 0, stream, nnz4, j4, x2, d_y, lease.tex, d_sink); break;
    }
    MXG_TRY(lease.release(stream));
    return MXG_OK;
}


// ================================================================================================
// CSR x SPARSE vector.  Replaces matmul_csr_svec<> (src/matmul.cpp:486-551) and its
// exports _numeric/_integer/_logical/_binary/_float32 (553-641): out[r] = sum over the columns that
// row r and the sparse vector share.  The reference walks both sorted index lists per row (merge with
// std::lower_bound skips); here the sparse vector is scattered once into a dense double image plus a
// presence bitmap (K/8 bytes: 125 KB for 1 M columns, kept in SHARED memory when it fits), and a
// row-split kernel tests each stored column against the bitmap, touching x[] and the dense image
// only on a hit.  The sums visit a row's matches in the same set as the reference (tree order instead
// of left-to-right, 1e-12); rows of A need not be sorted here.  NA rules of src/matmul.cpp:523-531:
// integer / logical NA contributes NA_real_ (also a stored NA_real_ of a numeric vector keeps R's payload).
// Bound: HBM on the 4-byte column ids (x is read only where the vector has an entry).
// ================================================================================================

template <int YTYPE>
struct SvecValue;
template <>
struct SvecValue<MXG_Y_NUMERIC> {
    typedef double elem;
    static __device__ __forceinline__ double get(const double *v, int k) { return v[k]; }
};
template <>
struct SvecValue<MXG_Y_INTEGER> {
    typedef int elem;
    static __device__ __forceinline__ double get(const int *v, int k) { return v[k] == INT_MIN ? na_real() : (double)v[k]; }
};
template <>
struct SvecValue<MXG_Y_LOGICAL> {
    typedef int elem;
    static __device__ __forceinline__ double get(const int *v, int k) { return v[k] == INT_MIN ? na_real() : (v[k] != 0 ? 1.0 : 0.0); }
};
template <>
struct SvecValue<MXG_Y_FLOAT32> {
    typedef float elem;
    static __device__ __forceinline__ double get(const float *v, int k) { return (double)v[k]; }
};
template <>
struct SvecValue<MXG_Y_BINARY> {
    typedef int elem;
    static __device__ __forceinline__ double get(const int *, int) { return 1.0; }
};

// dense image + bitmap of the sparse vector.  A repeated index keeps its FIRST occurrence, as the
// reference's merge does for a sorted list (src/matmul.cpp:520-535: both cursors advance on a match) — also when the
// vector is not sorted and the repeats are far apart: pass 1 records, per column, the smallest position k that names
// it (integer atomicMin: commutative, so the winner does not depend on scheduling), pass 2 lets exactly that entry
// write the value.  `first` is a K-entry int scratch initialised to INT_MAX.
__global__ void __launch_bounds__(256) k_svec_first(int n_y, const int32_t *__restrict__ yidx_base1, int K, int *__restrict__ first,
                                                    unsigned *__restrict__ mask)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_y) return;
    const int c = yidx_base1[k] - 1;
    if (c < 0 || c >= K) return; // can never equal a column id of A
    atomicMin(first + c, k);
    atomicOr(mask + (c >> 5), 1u << (c & 31));
}

template <int YTYPE>
__global__ void __launch_bounds__(256) k_svec_scatter(int n_y, const int32_t *__restrict__ yidx_base1,
                                                      const typename SvecValue<YTYPE>::elem *__restrict__ yvals, int K,
                                                      double *__restrict__ yd, const int *__restrict__ first)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_y) return;
    const int c = yidx_base1[k] - 1;
    if (c < 0 || c >= K) return;
    if (first[c] == k) yd[c] = SvecValue<YTYPE>::get(yvals, k);
}

struct SvecArgs {
    int m;
    const int32_t *p;
    const int32_t *j;
    const double *x;
    const double *yd;
    const unsigned *mask;
    int K;          // columns covered by the bitmap
    int mask_words; // ceil(K / 32)
    double *out;
    int piece, n_pieces;
    const int32_t *piece_row;
    const int32_t *piece_k;
    double *partial;
    int *partial_na;
    const int *abort;
};

template <int LPR, bool SMASK>
__device__ __forceinline__ double team_dot_masked(const int a, const int b, const int maxlen, const int l,
                                                  const int32_t *__restrict__ j, const double *__restrict__ x,
                                                  const double *__restrict__ yd, const unsigned *mask, const int K, bool &na)
{
    constexpr int U = 8; // column ids in flight per lane: the ids are the only streamed operand
    double acc = 0.0;
    for (int e0 = 0; e0 < maxlen; e0 += LPR * U) {
        int jj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = a + e0 + u * LPR + l;
            jj[u] = e < b ? __ldcs(j + e) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned c = (unsigned)jj[u];
            if (c < (unsigned)K) {
                const unsigned w = SMASK ? mask[c >> 5] : __ldg(mask + (c >> 5));
                if ((w >> (c & 31)) & 1u) {
                    const double yy = __ldg(yd + c);
                    const double xx = __ldg(x + (a + e0 + u * LPR + l));
                    if (__double_as_longlong(yy) == 0x7FF00000000007A2LL) na = true;
                    acc = fma(xx, yy, acc);
                }
            }
        }
    }
    return acc;
}

template <int LPR, bool SMASK>
__global__ void __launch_bounds__(SMASK ? 1024 : 256) k_spmv_svec(const SvecArgs g)
{
    extern __shared__ unsigned s_mask[];
    const unsigned *mask = g.mask;
    if (g.abort != nullptr && *g.abort != 0) return;
    if (SMASK) {
        for (int i = threadIdx.x; i < g.mask_words; i += blockDim.x) s_mask[i] = g.mask[i];
        __syncthreads();
        mask = s_mask;
    }
    const int32_t *__restrict__ p = g.p;
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    // long-row pieces first: one warp each, partial sums combined by k_spmv_fixup in piece order
    for (int pc = blockIdx.x * warps + (threadIdx.x >> 5); pc < g.n_pieces; pc += gridDim.x * warps) {
        const int row = g.piece_row[pc];
        const int a = p[row] + g.piece_k[pc] * g.piece;
        const int b = min(a + g.piece, p[row + 1]);
        bool na = false;
        double acc = team_dot_masked<32, SMASK>(a, b, b - a, lane, g.j, g.x, g.yd, mask, g.K, na);
        acc = team_reduce<32>(acc);
        na = team_any<32>(na);
        if (lane == 0) {
            g.partial[pc] = acc;
            g.partial_na[pc] = na ? 1 : 0;
        }
    }
    const int teams = blockDim.x / LPR;
    const int team = threadIdx.x / LPR;
    const int l = threadIdx.x % LPR;
    const int n_blocks = (g.m + teams - 1) / teams;
    const SpmvExtra none = {{nullptr}, 0};
    for (int rb = blockIdx.x; rb < n_blocks; rb += gridDim.x) {
        const int row = rb * teams + team;
        int a = 0, b = 0;
        bool store = false;
        if (row < g.m) {
            a = p[row];
            b = p[row + 1];
            store = true;
            if (b - a > g.piece) {
                b = a;
                store = false;
            }
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, b - a);
        bool na = false;
        double acc = team_dot_masked<LPR, SMASK>(a, b, maxlen, l, g.j, g.x, g.yd, mask, g.K, na);
        acc = team_reduce<LPR>(acc);
        na = team_any<LPR>(na);
        if (store && l == 0) store_result<MXG_Y_NUMERIC>(g.out, none, row, acc, na);
    }
}

template <int YTYPE>
static int svec_scatter(int n_y, const int32_t *d_yidx, const void *d_yvals, int K, double *yd, const int *first, cudaStream_t stream)
{
    MXG_LAUNCH((k_svec_scatter<YTYPE>), ceil_div_i(n_y, 256), 256, 0, stream, n_y, d_yidx,
               static_cast<const typename SvecValue<YTYPE>::elem *>(d_yvals), K, yd, first);
    return MXG_OK;
}

int launch_spmv_svec(const mxg_csr_s *A, int ytype, int K, int n_y, const int32_t *d_yidx_base1, const void *d_yvals,
                     double *d_out, cudaStream_t stream)
{
    if (A->m == 0) return MXG_OK;
    if (n_y < 0 || K < 0) return fail(MXG_ERR_ARG, "svec: negative size");
    if (n_y == 0 || K == 0 || A->nnz == 0) { // src/matmul.cpp:495-496: an empty vector gives the zero-filled result
        MXG_CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)A->m, stream));
        return MXG_OK;
    }
    if (!A->d_x64) return fail(MXG_ERR_UNSUPPORTED, "svec: handle holds no float64 values");
    if (!d_yidx_base1 || (!d_yvals && ytype != MXG_Y_BINARY)) return fail(MXG_ERR_ARG, "svec: NULL vector");
    const int words = ceil_div_i(K, 32);
    double *yd = nullptr;
    unsigned *mask = nullptr;
    int *first = nullptr;
    // stream-ordered temporaries, released on every path
    struct Temps {
        cudaStream_t s;
        void *q[3] = {nullptr, nullptr, nullptr};
        ~Temps()
        {
            for (void *v : q)
                if (v) cudaFreeAsync(v, s);
        }
    } temps{stream};
    MXG_CUDA_TRY(cudaMallocAsync(&yd, sizeof(double) * (size_t)K, stream));
    temps.q[0] = yd;
    MXG_CUDA_TRY(cudaMallocAsync(&mask, sizeof(unsigned) * (size_t)words, stream));
    temps.q[1] = mask;
    MXG_CUDA_TRY(cudaMallocAsync(&first, sizeof(int) * (size_t)K, stream));
    temps.q[2] = first;
    MXG_CUDA_TRY(cudaMemsetAsync(mask, 0, sizeof(unsigned) * (size_t)words, stream));
    MXG_CUDA_TRY(cudaMemsetAsync(first, 0x7f, sizeof(int) * (size_t)K, stream)); // 0x7f7f7f7f: above every position
    MXG_LAUNCH(k_svec_first, ceil_div_i(n_y, 256), 256, 0, stream, n_y, d_yidx_base1, K, first, mask);
    int rc = MXG_OK;
    switch (ytype) {
    case MXG_Y_NUMERIC: rc = svec_scatter<MXG_Y_NUMERIC>(n_y, d_yidx_base1, d_yvals, K, yd, first, stream); break;
    case MXG_Y_INTEGER: rc = svec_scatter<MXG_Y_INTEGER>(n_y, d_yidx_base1, d_yvals, K, yd, first, stream); break;
    case MXG_Y_LOGICAL: rc = svec_scatter<MXG_Y_LOGICAL>(n_y, d_yidx_base1, d_yvals, K, yd, first, stream); break;
    case MXG_Y_FLOAT32: rc = svec_scatter<MXG_Y_FLOAT32>(n_y, d_yidx_base1, d_yvals, K, yd, first, stream); break;
    case MXG_Y_BINARY: rc = svec_scatter<MXG_Y_BINARY>(n_y, d_yidx_base1, d_yvals, K, yd, first, stream); break;
    default: rc = fail(MXG_ERR_ARG, "svec: bad ytype %d", ytype);
    }
    if (rc == MXG_OK) {
        SvecArgs args;
        args.m = A->m;
        args.p = A->d_p;
        args.j = A->d_j;
        args.x = A->d_x64;
        args.yd = yd;
        args.mask = mask;
        args.K = K;
        args.mask_words = words;
        args.out = d_out;
        args.piece = A->piece;
        args.n_pieces = A->n_pieces;
        args.piece_row = A->d_piece_row;
        args.piece_k = A->d_piece_k;
        args.partial = nullptr;
        args.partial_na = nullptr;
        args.abort = A->d_abort;
        PartialLease partial;
        if (A->n_pieces > 0) {
            rc = partial.acquire(A, (size_t)A->n_pieces * 16, stream);
            args.partial = static_cast<double *>(partial.ptr);
            args.partial_na = reinterpret_cast<int *>(static_cast<double *>(partial.ptr) + A->n_pieces);
        }
        int lpr = (int)options().spmv_lpr;
        if (lpr != 4 && lpr != 8 && lpr != 16 && lpr != 32) {
            const double mean = (double)A->nnz / (double)A->m;
            lpr = mean <= 12 ? 4 : mean <= 96 ? 8 : mean <= 384 ? 16 : 32; // 8 ids per lane per step
        }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t smem = sizeof(unsigned) * (size_t)words;
        const bool smask = options().svec_smem != 0 && smem <= 200 * 1024;
#define MXG_SVEC(L)                                                                                                    \
    if (rc == MXG_OK && lpr == L) {                                                                                    \
        auto body = [&]() -> int {                                                                                     \
            if (smask) {                                                                                               \
                MXG_CUDA_TRY(cudaFuncSetAttribute(k_spmv_svec<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                MXG_LAUNCH((k_spmv_svec<L, true>), sms, 1024, smem, stream, args);                                     \
            } else {                                                                                                   \
                const int want = std::max(ceil_div_i(A->m, 256 / L), ceil_div_i(A->n_pieces, 8));                      \
                MXG_LAUNCH((k_spmv_svec<L, false>), std::min(want, sms * 64), 256, 0, stream, args);                   \
            }                                                                                                          \
            return MXG_OK;                                                                                             \
        };                                                                                                             \
        rc = body();                                                                                                   \
    }
        MXG_SVEC(4) MXG_SVEC(8) MXG_SVEC(16) MXG_SVEC(32)
#undef MXG_SVEC
        if (rc == MXG_OK && A->n_long > 0) {
            auto body = [&]() -> int {
                const SpmvExtra none = {{nullptr}, 0};
                MXG_LAUNCH((k_spmv_fixup<MXG_Y_NUMERIC>), ceil_div_i(A->n_long, 128), 128, 0, stream, A->n_long, A->d_long_rows,
                           A->d_long_first, A->d_long_np, args.partial, args.partial_na, d_out, none, A->d_abort);
                return MXG_OK;
            };
            rc = body();
        }
    }
    return rc; // (temps released by ~Temps)
}

} // namespace mxg
