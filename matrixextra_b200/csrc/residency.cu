// residency.cu — where the operands of the host-buffer (level-1) products live between and during calls.
//
//  1. Several devices for one call (mxg_set_devices(n), SURVEY.md 8 b/e).  The reference parallelises INSIDE one
//     process (one OpenMP team over the rows, src/matmul.cpp:132-136; R/matmul.R:175-180), and an R session is one
//     process: so the row blocks of a level-1 product are dealt out to n GPUs by n host threads of the calling
//     process — nnz-balanced blocks (row_partition), each thread running the streamed pipeline (pipeline.cu) on its
//     block: it uploads only its block of the CSR over its OWN PCIe link and downloads its rows straight into the
//     caller's result, so no all-gather exists on this path (the result is wanted on the host).  The dense
//     operand is needed whole by every device: each uploads one slice and pulls the others over NVLink (DenseShare).
//
//  2. The level-1 operand cache (option "cache_mb").  The callers of the reference multiply one
//     matrix hundreds of times (vignettes/Introducing_MatrixExtra.Rmd:454-476: `X %*% coefs` inside optim), and
//     every .Call hands the glue the same R vectors again.  With the cache on, the device CSR of a streamed call is
//     kept (pipeline_spmm(keep)) under a key made of the host addresses, the sizes and a sampled fingerprint of
//     the three arrays; the next call with the same key moves only the dense operand and the result
//     (handle_spmm_host).  Dense operands are kept the same way.  LRU by bytes.  It is opt-in because R objects
//     can be modified in place (MatrixExtra.inplace_sort, R/utils.R:22-161): a change that misses all sampled
//     positions would go unnoticed.  Explicit handles (mxg_csr_upload + mxg_csr_spmm_host) have no such caveat.
#include "mxg_internal.cuh"

#include <algorithm>
#include <cstring>
#include <list>
#include <thread>

namespace mxg {

// ------------------------------------------------------------------------------------------------
// devices of a level-1 call
// ------------------------------------------------------------------------------------------------
namespace {
int g_ndev = 1;
int g_devs[MXG_MAX_DST] = {0};
bool g_peer_ok = false;
} // namespace

int multi_devices() { return g_ndev; }

int set_devices(int n)
{
    int count = 0;
    MXG_CUDA_TRY(cudaGetDeviceCount(&count));
    if (n < 1 || n > MXG_MAX_DST) return fail(MXG_ERR_ARG, "set_devices: 1 .. %d devices", MXG_MAX_DST);
    if (n > count) return fail(MXG_ERR_ARG, "set_devices: %d devices requested, %d visible", n, count);
    int cur = 0;
    MXG_CUDA_TRY(cudaGetDevice(&cur));
    g_ndev = 1;
    if (n == 1) return MXG_OK;
    // devices cur, cur+1, ... (mod count): the calling thread's device keeps block 0
    for (int g = 0; g < n; g++) g_devs[g] = (cur + g) % count;
    bool peer = true;
    for (int g = 0; g < n; g++) {
        MXG_CUDA_TRY(cudaSetDevice(g_devs[g]));
        DeviceState *st;
        MXG_TRY(current_state(&st)); // streams + pool settings
        for (int q = 0; q < n; q++) {
            if (q == g) continue;
            int can = 0;
            MXG_CUDA_TRY(cudaDeviceCanAccessPeer(&can, g_devs[g], g_devs[q]));
            if (!can) {
                peer = false;
                continue;
            }
            // peer access covers cudaMalloc'ed memory: the devices' copies of a shared dense operand live in such a
            // buffer (DeviceState::share_buf).  The default memory pool is left alone on purpose: making IT
            // peer-accessible (cudaMemPoolSetAccess) made later multi-GB cudaMallocAsync calls fail with "out of
            // memory" on a box with 176 GB free.
            const cudaError_t e = cudaDeviceEnablePeerAccess(g_devs[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false;
            cudaGetLastError();
        }
    }
    MXG_CUDA_TRY(cudaSetDevice(cur));
    g_peer_ok = peer;
    g_ndev = n;
    return MXG_OK;
}

int row_partition(int m, const int32_t *p, int parts, int32_t *row_starts)
{
    const int64_t base = p[0], nnz = (int64_t)p[m] - base;
    row_starts[0] = 0;
    for (int g = 1; g < parts; g++) {
        // first row whose start offset reaches g/parts of the entries
        const int64_t target = base + (nnz * g) / parts;
        const int32_t *it = std::lower_bound(p, p + m + 1, (int32_t)std::min<int64_t>(target, INT32_MAX));
        int r = (int)(it - p);
        if (r > m) r = m;
        if (nnz == 0) r = (int)(((int64_t)m * g) / parts);
        if (r < row_starts[g - 1]) r = row_starts[g - 1];
        row_starts[g] = r;
    }
    row_starts[parts] = m;
    return MXG_OK;
}

// Should this call be spread over the devices?  Small products stay on one (a second device costs a thread start,
// a second staging arena and — for SpMM — the NVLink exchange of the dense operand).
// Nor do calls whose CSR arrays are ordinary pageable memory (option multi_pageable = 0): those are bound by the host
// threads that bounce the arrays into page-locked slots — the memory bandwidth of the host cores, not a PCIe link — and
// more pipelines only cut the same work into smaller pieces (measured on the 16-core B200 box, cfg3: 55 ms on one
// device, 64 ms on two, 97 ms on eight; page-locked arrays: 26.4 -> 14.6 ms on eight, the host's DMA roof).
bool multi_wanted(int m, const int32_t *p, const int32_t *j, const double *x)
{
    if (g_ndev <= 1 || m < g_ndev) return false;
    if ((int64_t)p[m] - (int64_t)p[0] < (int64_t)std::max<long>(options().multi_min_nnz, 1)) return false;
    if (options().multi_pageable == 0 && !(host_is_pinned(j) && host_is_pinned(x))) return false;
    return true;
}

namespace {
thread_local bool g_in_multi_call = false;
}
bool multi_wanted_now() { return g_in_multi_call; }

namespace {

// body(g, st) runs on a thread bound to device g_devs[g]; block 0 runs on the calling thread
template <class Body>
int run_on_devices(int G, Body body)
{
    int cur = 0;
    MXG_CUDA_TRY(cudaGetDevice(&cur));
    std::vector<int> rc((size_t)G, MXG_OK);
    std::vector<std::string> err((size_t)G);
    std::vector<size_t> up((size_t)G, 0), down((size_t)G, 0);
    auto work = [&](int g) {
        int r = MXG_OK;
        DeviceState *st = nullptr;
        if (cudaSetDevice(g_devs[g]) != cudaSuccess) r = fail(MXG_ERR_CUDA, "cudaSetDevice(%d) failed", g_devs[g]);
        if (r == MXG_OK) r = current_state(&st);
        g_in_multi_call = true;
        r = body(g, st, r);
        g_in_multi_call = false;
        rc[(size_t)g] = r;
        if (r != MXG_OK) err[(size_t)g] = last_error_ref();
        last_call_bytes(&up[(size_t)g], &down[(size_t)g]);
    };
    std::vector<std::thread> threads;
    threads.reserve((size_t)G);
    int started = 1;
    try {
        for (; started < G; started++) threads.emplace_back(work, started);
    } catch (...) { // no thread to be had: the blocks without one fail here, so that nobody waits for them (DenseShare)
        for (int g = started; g < G; g++) {
            rc[(size_t)g] = body(g, nullptr, fail(MXG_ERR_CUDA, "multi-device call: could not start a host thread for device %d", g_devs[g]));
            err[(size_t)g] = last_error_ref();
        }
    }
    work(0);
    for (std::thread &t : threads) t.join();
    cudaSetDevice(cur);
    size_t h2d = 0, d2h = 0;
    for (int g = 0; g < G; g++) {
        h2d += up[(size_t)g];
        d2h += down[(size_t)g];
    }
    set_last_call_bytes(h2d, d2h);
    // report the root cause: a device that gave up because a peer had failed says so, the peer says why
    int first = -1;
    for (int g = 0; g < G; g++)
        if (rc[(size_t)g] != MXG_OK && (first < 0 || (err[(size_t)first].find("another device failed") != std::string::npos &&
                                                      err[(size_t)g].find("another device failed") == std::string::npos)))
            first = g;
    if (first >= 0) return fail(rc[(size_t)first], "device %d: %s", g_devs[first], err[(size_t)first].c_str());
    return MXG_OK;
}

} // namespace

int multi_spmm(int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p, const int32_t *j,
               const double *x, const void *B, size_t ldb, void *Out, size_t ldc)
{
    const int G = g_ndev;
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    int32_t bounds[MXG_MAX_DST + 1];
    MXG_TRY(row_partition(m, p, G, bounds));
    // the dense operand crosses PCIe once, a slice per device, and is completed over NVLink (worth it from a few MiB)
    DenseShare share;
    share.G = G;
    for (int g = 0; g < G; g++) share.device[g] = g_devs[g];
    bool use_share = g_peer_ok && options().multi_dense_share != 0 && b_layout == MXG_ROWS_CONTIGUOUS && K >= G && n > 0 &&
                     (size_t)K * (size_t)n * s >= ((size_t)4 << 20);
    for (int g = 0; g < G; g++)
        if (bounds[g + 1] == bounds[g]) use_share = false; // a block without rows returns at once: nobody to pull its slice from
    const bool rm = out_layout == MXG_ROWS_CONTIGUOUS;
    const int rc = run_on_devices(G, [&](int g, DeviceState *st, int r) -> int {
        const int r0 = bounds[g], r1 = bounds[g + 1];
        if (r == MXG_OK && use_share && cudaEventCreateWithFlags(&share.slice_ready[g], cudaEventDisableTiming) != cudaSuccess)
            r = fail(MXG_ERR_CUDA, "multi-device product: cudaEventCreate failed");
        if (r == MXG_OK) {
            char *out_g = static_cast<char *>(Out) + (rm ? (size_t)r0 * ldc * s : (size_t)r0 * s);
            r = pipeline_spmm(st, dtype, out_layout, b_layout, r1 - r0, K, n, p + r0, j, x, B, ldb, out_g, ldc,
                              use_share ? &share : nullptr, g);
        }
        if (use_share) share.leave(g, r == MXG_OK);
        return r;
    });
    if (use_share) {
        int cur = 0;
        cudaGetDevice(&cur);
        for (int g = 0; g < G; g++)
            if (share.slice_ready[g]) {
                cudaSetDevice(g_devs[g]);
                cudaEventDestroy(share.slice_ready[g]);
            }
        cudaSetDevice(cur);
    }
    return rc;
}

int multi_spmv(int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x, const void *y, void *out)
{
    const int G = g_ndev;
    const size_t os = ytype == MXG_Y_FLOAT32 ? 4 : 8;
    int32_t bounds[MXG_MAX_DST + 1];
    MXG_TRY(row_partition(m, p, G, bounds));
    return run_on_devices(G, [&](int g, DeviceState *st, int r) -> int {
        if (r != MXG_OK) return r;
        const int r0 = bounds[g], r1 = bounds[g + 1];
        return pipeline_spmv(st, ytype, r1 - r0, K, p + r0, j, x, y, static_cast<char *>(out) + (size_t)r0 * os);
    });
}

// ------------------------------------------------------------------------------------------------
// level-1 operand cache
// ------------------------------------------------------------------------------------------------
namespace {

// 64-bit mix of up to ~1000 sampled 8-byte words plus both ends of the array (splitmix64 finaliser per word)
uint64_t mix64(uint64_t h, uint64_t v)
{
    v += 0x9E3779B97F4A7C15ULL + h;
    v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ULL;
    v = (v ^ (v >> 27)) * 0x94D049BB133111EBULL;
    return v ^ (v >> 31);
}

uint64_t fingerprint(const void *ptr, size_t bytes)
{
    uint64_t h = mix64(0x6d7867ULL, bytes);
    if (!ptr || bytes == 0) return h;
    const unsigned char *b = static_cast<const unsigned char *>(ptr);
    auto word = [&](size_t off) {
        uint64_t w = 0;
        memcpy(&w, b + off, std::min<size_t>(8, bytes - off));
        return w;
    };
    const size_t ends = std::min<size_t>(bytes, 512);
    for (size_t o = 0; o < ends; o += 8) h = mix64(h, word(o));
    for (size_t o = bytes - ends; o < bytes; o += 8) h = mix64(h, word(o));
    const size_t samples = 1000, words = bytes / 8;
    if (words > 2 * samples) {
        const size_t stride = words / samples;
        for (size_t k = 0; k < samples; k++) h = mix64(h, word((k * stride + (k * 7) % stride) * 8));
    } else {
        for (size_t o = 0; o + 8 <= bytes; o += 8) h = mix64(h, word(o));
    }
    return h;
}

struct CsrEntry {
    int device;
    const void *p, *j, *x;
    int m, K;
    int64_t nnz;
    uint64_t fp;
    mxg_csr_s *h;
    size_t bytes;
};
struct DenseEntry {
    int device;
    const void *ptr;
    int dtype, layout;
    size_t K, n, ldb;
    uint64_t fp;
    void *d_B;
    cudaStream_t stream; // the stream its allocation is ordered on
    size_t bytes;
};
std::mutex g_cache_mu;
std::list<CsrEntry> g_csr_cache;     // most recently used first
std::list<DenseEntry> g_dense_cache; // most recently used first
size_t g_cache_bytes = 0;
unsigned long long g_cache_hits = 0, g_cache_misses = 0;

size_t handle_bytes(const mxg_csr_s *h)
{
    size_t b = 4 * ((size_t)h->m + 1) + 4 * (size_t)h->nnz;
    if (h->d_x64) b += 8 * (size_t)h->nnz;
    if (h->d_x32) b += 4 * (size_t)h->nnz;
    if (h->cached_t) b += handle_bytes(h->cached_t);
    return b;
}

// entries are released on the device that holds them, whichever device the caller has made current since
struct OnDevice {
    int prev = -1;
    explicit OnDevice(int dev)
    {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
    }
    ~OnDevice()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

void drop_csr(std::list<CsrEntry>::iterator it)
{
    OnDevice on(it->device);
    g_cache_bytes -= it->bytes;
    csr_handle_free(it->h);
    g_csr_cache.erase(it);
}

void drop_dense(std::list<DenseEntry>::iterator it)
{
    OnDevice on(it->device);
    g_cache_bytes -= it->bytes;
    cudaFreeAsync(it->d_B, it->stream);
    g_dense_cache.erase(it);
}

// make room for `incoming` bytes: least recently used entries go first, dense operands before matrices
void evict_for(size_t incoming)
{
    const size_t cap = (size_t)std::max<long>(options().cache_mb, 0) << 20;
    while (g_cache_bytes + incoming > cap && !g_dense_cache.empty()) drop_dense(std::prev(g_dense_cache.end()));
    while (g_cache_bytes + incoming > cap && !g_csr_cache.empty()) drop_csr(std::prev(g_csr_cache.end()));
}

} // namespace

bool cache_enabled() { return options().cache_mb > 0; }

uint64_t csr_fingerprint(int m, const int32_t *p, const int32_t *j, const double *x)
{
    const size_t nnz = (size_t)((int64_t)p[m] - (int64_t)p[0]);
    uint64_t h = fingerprint(p, 4 * ((size_t)m + 1));
    h = mix64(h, fingerprint(j, 4 * nnz));
    return mix64(h, fingerprint(x, 8 * nnz));
}

// the cached handle of these host arrays on the current device, or NULL; need = MXG_KEEP_* bits the product reads
mxg_csr_s *cache_find_csr(int m, int K, const int32_t *p, const int32_t *j, const double *x, int need)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t nnz = (int64_t)p[m] - (int64_t)p[0];
    const uint64_t fp = csr_fingerprint(m, p, j, x);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto it = g_csr_cache.begin(); it != g_csr_cache.end(); ++it) {
        if (it->device != dev || it->p != p || it->j != j || it->x != x || it->m != m || it->K != K || it->nnz != nnz) continue;
        if (it->fp != fp) { // same arrays, other contents: modified in place since
            drop_csr(it);
            break;
        }
        const bool ok = (!(need & MXG_KEEP_F64) || it->h->d_x64 || nnz == 0) && (!(need & MXG_KEEP_F32) || it->h->d_x32 || nnz == 0);
        if (!ok) { // holds the other value type: the caller streams the matrix again and replaces the entry
            drop_csr(it);
            break;
        }
        g_csr_cache.splice(g_csr_cache.begin(), g_csr_cache, it);
        g_cache_hits++;
        return it->h;
    }
    g_cache_misses++;
    return nullptr;
}

// takes ownership of h (it is released at once when it does not fit the budget)
void cache_insert_csr(int m, int K, const int32_t *p, const int32_t *j, const double *x, mxg_csr_s *h)
{
    if (!h) return;
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t bytes = handle_bytes(h);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const size_t cap = (size_t)std::max<long>(options().cache_mb, 0) << 20;
    if (bytes > cap) {
        csr_handle_free(h);
        return;
    }
    evict_for(bytes);
    CsrEntry e;
    e.device = dev;
    e.p = p;
    e.j = j;
    e.x = x;
    e.m = m;
    e.K = K;
    e.nnz = (int64_t)p[m] - (int64_t)p[0];
    e.fp = csr_fingerprint(m, p, j, x);
    e.h = h;
    e.bytes = bytes;
    g_csr_cache.push_front(e);
    g_cache_bytes += bytes;
}

// a cached entry grew (its CSC was built): keep the byte count honest
void cache_account_csr(mxg_csr_s *h)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (CsrEntry &e : g_csr_cache)
        if (e.h == h) {
            const size_t now = handle_bytes(h);
            g_cache_bytes += now - e.bytes;
            e.bytes = now;
        }
}

void *cache_find_dense(const void *ptr, int dtype, int layout, size_t K, size_t n, size_t ldb, size_t host_bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t fp = fingerprint(ptr, host_bytes);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto it = g_dense_cache.begin(); it != g_dense_cache.end(); ++it) {
        if (it->device != dev || it->ptr != ptr || it->dtype != dtype || it->layout != layout || it->K != K || it->n != n || it->ldb != ldb)
            continue;
        if (it->fp != fp) {
            drop_dense(it);
            break;
        }
        g_dense_cache.splice(g_dense_cache.begin(), g_dense_cache, it);
        return it->d_B;
    }
    return nullptr;
}

void cache_insert_dense(const void *ptr, int dtype, int layout, size_t K, size_t n, size_t ldb, size_t host_bytes, void *d_B,
                        size_t dev_bytes, cudaStream_t stream)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const size_t cap = (size_t)std::max<long>(options().cache_mb, 0) << 20;
    if (dev_bytes > cap / 2) { // a dense operand never pushes the matrices out
        cudaFreeAsync(d_B, stream);
        return;
    }
    evict_for(dev_bytes);
    DenseEntry e;
    e.device = dev;
    e.ptr = ptr;
    e.dtype = dtype;
    e.layout = layout;
    e.K = K;
    e.n = n;
    e.ldb = ldb;
    e.fp = fingerprint(ptr, host_bytes);
    e.d_B = d_B;
    e.stream = stream;
    e.bytes = dev_bytes;
    g_dense_cache.push_front(e);
    g_cache_bytes += dev_bytes;
}

int cache_clear()
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    while (!g_dense_cache.empty()) drop_dense(g_dense_cache.begin());
    while (!g_csr_cache.empty()) drop_csr(g_csr_cache.begin());
    return MXG_OK;
}

void cache_stats(unsigned long long *hits, unsigned long long *misses, size_t *bytes, int *entries)
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (hits) *hits = g_cache_hits;
    if (misses) *misses = g_cache_misses;
    if (bytes) *bytes = g_cache_bytes;
    if (entries) *entries = (int)(g_csr_cache.size() + g_dense_cache.size());
}

} // namespace mxg
