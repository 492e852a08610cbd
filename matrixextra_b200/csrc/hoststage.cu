// hoststage.cu — host side of the streamed (level-1) products: worker threads + a page-locked staging arena.
//
// An Rcpp export (src/matmul.cpp:221-483) is handed R vectors: PAGEABLE memory, float64 values even when the
// product is float32 (src/matmul.cpp:213-214 narrows per entry), and it returns a freshly allocated R matrix
// whose pages have never been touched.  cudaMemcpyAsync on such memory is a single-threaded bounce through the
// driver's own staging buffer (~10 GB/s, and ~4 GB/s into untouched pages), five times slower than the link.
// So the library brings its own bounce: a grow-only page-locked arena per device, cut into ring slots, and a
// small pool of host threads (what the reference's `nthreads` argument now controls) that
//     * narrows float64 values to float32 straight into a slot (bit-identical to the device narrowing and to
//       the reference's per-entry `(float)values[ix]`: round-to-nearest-even) — a float32 product then moves
//       8 instead of 12 bytes per stored entry over PCIe,
//     * copies pageable indices / dense operands into slots,
//     * copies finished output rows from slots into the caller's matrix, first-touching its pages in parallel.
// Page-locked caller memory (cudaHostAlloc / cudaHostRegister) is detected and DMA'd directly.
#include "mxg_internal.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <emmintrin.h>
#include <sys/mman.h>

namespace mxg {

// ------------------------------------------------------------------------------------------------
// worker pool: run(ntasks, fn) executes fn(0..ntasks-1) on the calling thread plus the workers
// ------------------------------------------------------------------------------------------------
namespace {

class HostPool {
public:
    static HostPool &get()
    {
        static HostPool *pool = new HostPool; // leaked on purpose: workers may still be parked at process exit
        return *pool;
    }

    void run(size_t ntasks, int threads, const std::function<void(size_t)> &fn)
    {
        if (ntasks == 0) return;
        threads = (int)std::min<size_t>((size_t)std::max(threads, 1), ntasks);
        if (threads <= 1) {
            for (size_t i = 0; i < ntasks; i++) fn(i);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mu_); // one parallel region at a time
        ensure(threads - 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &fn;
            ntasks_ = ntasks;
            next_.store(0, std::memory_order_relaxed);
            invited_ = threads - 1;
            pending_ = threads - 1;
            generation_++;
        }
        cv_work_.notify_all();
        for (size_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < ntasks;) fn(i);
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void ensure(int workers)
    {
        while ((int)workers_.size() < workers) {
            const int id = (int)workers_.size();
            workers_.emplace_back([this, id] { loop(id); });
            workers_.back().detach();
        }
    }

    void loop(int id)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(size_t)> *job = nullptr;
            size_t n = 0;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (id >= invited_) continue; // this region runs on fewer threads
                job = job_;
                n = ntasks_;
            }
            for (size_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < n;) (*job)(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                pending_--;
            }
            cv_done_.notify_one();
        }
    }

    std::mutex run_mu_, mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    const std::function<void(size_t)> *job_ = nullptr;
    size_t ntasks_ = 0;
    std::atomic<size_t> next_{0};
    int invited_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
};

// memcpy whose stores bypass the caches (destination = a ring slot the DMA engine reads next: no dirty lines to
// snoop out of the cores, and the caller's data does not get evicted by a copy of itself)
void copy_block_nt(char *d, const char *s, size_t n)
{
    size_t i = 0;
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head && head <= n) {
        memcpy(d, s, head);
        i = head;
    }
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + i + 48), e);
    }
    if (i < n) memcpy(d + i, s + i, n - i);
    _mm_sfence();
}

inline void copy_block(char *d, const char *s, size_t n, bool nt)
{
    if (nt && n >= 256) copy_block_nt(d, s, n);
    else memcpy(d, s, n);
}

void narrow_block(const double *s, float *d, size_t n)
{
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 15)) {
        d[i] = (float)s[i];
        i++;
    }
    for (; i + 4 <= n; i += 4) { // cvtpd2ps rounds by MXCSR: nearest-even, like the cast
        const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(s + i));
        const __m128 hi = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
        _mm_stream_ps(d + i, _mm_movelh_ps(lo, hi));
    }
    for (; i < n; i++) d[i] = (float)s[i];
    _mm_sfence();
}


// 32-bit column ids -> low halves (uint16 per entry) + high parts (a nibble per entry when hi_bits == 4, a byte
// when 8, nothing when 0); entry i's nibble is the low one of byte i/2 for even i.  [a, b) with a % 32 == 0;
// lo and hi 16-byte aligned.  Returns non-zero when an id lies outside [0, K).
int pack_block(const int32_t *j, size_t a, size_t b, int K, int hi_bits, uint16_t *lo, unsigned char *hi)
{
    const __m128i zero = _mm_setzero_si128(), kmax = _mm_set1_epi32(K - 1), low8 = _mm_set1_epi16(0x00ff);
    __m128i bad = zero;
    size_t i = a;
    for (; i + 32 <= b; i += 32) {
        __m128i v[8], h[8];
#pragma GCC unroll 8
        for (int q = 0; q < 8; q++) {
            v[q] = _mm_loadu_si128(reinterpret_cast<const __m128i *>(j + i) + q);
            bad = _mm_or_si128(bad, _mm_or_si128(_mm_cmplt_epi32(v[q], zero), _mm_cmpgt_epi32(v[q], kmax)));
            h[q] = _mm_srli_epi32(v[q], 16);
            v[q] = _mm_srai_epi32(_mm_slli_epi32(v[q], 16), 16); // sign-extended low half: the signed pack keeps its bits
        }
#pragma GCC unroll 4
        for (int q = 0; q < 4; q++)
            _mm_stream_si128(reinterpret_cast<__m128i *>(lo + i) + q, _mm_packs_epi32(v[2 * q], v[2 * q + 1]));
        if (hi_bits == 0) continue;
        // valid ids have high parts < 256 (< 16): the saturating packs are exact
        const __m128i b0 = _mm_packus_epi16(_mm_packs_epi32(h[0], h[1]), _mm_packs_epi32(h[2], h[3]));
        const __m128i b1 = _mm_packus_epi16(_mm_packs_epi32(h[4], h[5]), _mm_packs_epi32(h[6], h[7]));
        if (hi_bits == 8) {
            _mm_stream_si128(reinterpret_cast<__m128i *>(hi + i), b0);
            _mm_stream_si128(reinterpret_cast<__m128i *>(hi + i) + 1, b1);
        } else {
            // 16-bit lane = byte[2k] | byte[2k+1] << 8  ->  byte[2k] | byte[2k+1] << 4 in its low byte
            const __m128i n0 = _mm_and_si128(_mm_or_si128(b0, _mm_srli_epi16(b0, 4)), low8);
            const __m128i n1 = _mm_and_si128(_mm_or_si128(b1, _mm_srli_epi16(b1, 4)), low8);
            _mm_stream_si128(reinterpret_cast<__m128i *>(hi + i / 2), _mm_packus_epi16(n0, n1));
        }
    }
    int tail_bad = 0;
    for (; i < b; i++) {
        const int32_t c = j[i];
        tail_bad |= (c < 0 || c >= K);
        lo[i] = (uint16_t)c;
        const unsigned h = (unsigned)c >> 16;
        if (hi_bits == 8) hi[i] = (unsigned char)h;
        else if (hi_bits == 4) hi[i / 2] = (i & 1) ? (unsigned char)(hi[i / 2] | ((h & 15u) << 4)) : (unsigned char)(h & 15u);
    }
    _mm_sfence();
    return tail_bad | (_mm_movemask_epi8(bad) != 0);
}

} // namespace

// Threads for a region that moves `bytes`: one per MiB up to the pool size.  Small regions stay on one or two
// cores on purpose: lines of a slot that sit in many cores' caches make the DMA that follows snoop all of them
// (measured on the B200 box: a 2.5 MB download into a slot last read by 16 threads takes 0.41 ms, by one thread
// 0.055 ms), which costs a small call more than the parallel copy saves.
static int threads_for(size_t bytes)
{
    const size_t want = std::max<size_t>(1, bytes >> 20);
    return (int)std::min<size_t>(want, (size_t)host_threads());
}

int host_threads()
{
    long t = options().host_threads;
    if (t <= 0) {
        t = (long)std::thread::hardware_concurrency();
        if (t <= 0) t = 4;
        t = std::min<long>(t, 16);
    }
    return (int)std::min<long>(t, 64);
}

void host_parallel_for(size_t ntasks, size_t bytes_touched, const std::function<void(size_t)> &fn)
{
    HostPool::get().run(ntasks, threads_for(bytes_touched), fn);
}

void host_narrow_f64_to_f32(const double *src, float *dst, size_t n)
{
    const size_t grain = (size_t)1 << 15; // 256 KiB of doubles per task: small calls still spread over the pool
    HostPool::get().run((n + grain - 1) / grain, threads_for(n * sizeof(double)), [&](size_t t) {
        const size_t a = t * grain, b = std::min(n, a + grain);
        narrow_block(src + a, dst + a, b - a);
    });
}

// Column ids on the wire (streamed level-1 calls): ids below 2^16 / 2^20 / 2^24 travel as 2 / 2.5 / 3 bytes per
// entry instead of 4 — [uint16 low halves][high nibbles or bytes], unpacked by k_unpack_indices on the device.
int index_pack_hi_bits(int K) { return K <= (1 << 16) ? 0 : (K <= (1 << 20) ? 4 : (K <= (1 << 24) ? 8 : -1)); }

size_t packed_index_lo_bytes(size_t n) { return (2 * n + 15) & ~(size_t)15; }

size_t packed_index_bytes(size_t n, int hi_bits)
{
    const size_t hi = hi_bits == 0 ? 0 : (hi_bits == 4 ? (n + 1) / 2 : n);
    return packed_index_lo_bytes(n) + ((hi + 15) & ~(size_t)15);
}

bool host_pack_indices(const int32_t *j, size_t n, int K, int hi_bits, void *dst)
{
    uint16_t *lo = static_cast<uint16_t *>(dst);
    unsigned char *hi = static_cast<unsigned char *>(dst) + packed_index_lo_bytes(n);
    const size_t grain = (size_t)1 << 15; // entries per task (a multiple of 32: tasks never share a byte)
    std::atomic<int> bad{0};
    HostPool::get().run((n + grain - 1) / grain, threads_for(n * 8), [&](size_t t) {
        const size_t a = t * grain, b = std::min(n, a + grain);
        if (pack_block(j, a, b, K, hi_bits, lo, hi)) bad.store(1, std::memory_order_relaxed);
    });
    return bad.load() == 0;
}

void host_copy(void *dst, const void *src, size_t bytes, bool nt_dst)
{
    const size_t grain = (size_t)1 << 18;
    HostPool::get().run((bytes + grain - 1) / grain, threads_for(bytes), [&](size_t t) {
        const size_t a = t * grain, b = std::min(bytes, a + grain);
        copy_block(static_cast<char *>(dst) + a, static_cast<const char *>(src) + a, b - a, nt_dst);
    });
}

void host_copy_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, bool nt_dst)
{
    if (width == 0 || height == 0) return;
    if (dpitch == width && spitch == width) {
        host_copy(dst, src, width * height, nt_dst);
        return;
    }
    // tasks of about 256 KiB: several short lines, or a slice of a long one
    const size_t target = (size_t)1 << 18;
    if (width >= target) {
        const size_t per_line = (width + target - 1) / target;
        HostPool::get().run(height * per_line, threads_for(width * height), [&](size_t t) {
            const size_t line = t / per_line, a = (t % per_line) * target, b = std::min(width, a + target);
            copy_block(static_cast<char *>(dst) + line * dpitch + a, static_cast<const char *>(src) + line * spitch + a, b - a, nt_dst);
        });
    } else {
        const size_t lines = std::max<size_t>(1, target / width);
        HostPool::get().run((height + lines - 1) / lines, threads_for(width * height), [&](size_t t) {
            const size_t l0 = t * lines, l1 = std::min(height, l0 + lines);
            for (size_t l = l0; l < l1; l++)
                copy_block(static_cast<char *>(dst) + l * dpitch, static_cast<const char *>(src) + l * spitch, width, nt_dst);
        });
    }
}

// A pageable result the caller has just allocated (R: a fresh matrix; numpy: np.empty) consists of pages that do not
// exist yet: the copy out of the page-locked slots is then bound by page faults (one per 4 KiB, serialised on the
// process's address-space lock: measured 17 GB/s with 16 threads), not by memory bandwidth.  Asking for transparent huge
// pages on the 2 MiB-aligned interior turns 512 faults into one.  Advisory: ignored where THP is off, never touches data.
void host_prepare_result(void *ptr, size_t bytes)
{
    if (!ptr || bytes < ((size_t)8 << 20) || options().host_thp == 0) return;
    const uintptr_t huge = (uintptr_t)1 << 21;
    const uintptr_t a = (reinterpret_cast<uintptr_t>(ptr) + huge - 1) & ~(huge - 1);
    const uintptr_t b = (reinterpret_cast<uintptr_t>(ptr) + bytes) & ~(huge - 1);
    if (b > a) madvise(reinterpret_cast<void *>(a), b - a, MADV_HUGEPAGE);
}

// true when the range can be DMA'd directly: page-locked by CUDA (cudaHostAlloc / cudaHostRegister) or managed
bool host_is_pinned(const void *ptr)
{
    if (!ptr) return true;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------------------------------------------
// A new page-locked block.  cudaHostAlloc creates and pins its pages 4 KiB at a time, on one thread: 120 - 360 ms for
// 256 MiB on the B200 host — ten times what a product of that size takes.  Cheaper (tools/pinalloc_probe.cu,
// profiles/r02_pinned_block_alloc_probe.jsonl): map anonymous memory, ask for transparent huge pages, touch it with the
// host threads (one fault per 2 MiB) and register it with the driver — 20 ms for 256 MiB, 42 - 95 ms for 512 MiB, the
// same 57 GB/s as a DMA target.  Where the mapping or the registration is refused, cudaHostAlloc it is.
// ------------------------------------------------------------------------------------------------
int pinned_block_alloc(size_t bytes, PinnedBlock *blk)
{
    const size_t huge = (size_t)2 << 20;
    const size_t want = (std::max<size_t>(bytes, 1) + huge - 1) & ~(huge - 1);
    *blk = PinnedBlock();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { // nothing to map and touch memory for
        cudaGetLastError();
        return fail(MXG_ERR_CUDA, "no CUDA device: no page-locked memory to be had");
    }
    if (options().host_pin_register != 0) {
        const size_t map_len = want + huge;
        void *q = mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q != MAP_FAILED) {
            char *al = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(q) + huge - 1) & ~(uintptr_t)(huge - 1));
            madvise(al, want, MADV_HUGEPAGE); // advisory: 4 KiB pages otherwise
            HostPool::get().run(want / huge, host_threads(), [&](size_t t) {
                for (size_t o = 0; o < huge; o += 4096) al[t * huge + o] = 0;
            });
            if (cudaHostRegister(al, want, cudaHostRegisterPortable) == cudaSuccess) {
                blk->ptr = al;
                blk->bytes = want;
                blk->map_base = q;
                blk->map_len = map_len;
                return MXG_OK;
            }
            cudaGetLastError();
            munmap(q, map_len);
        }
    }
    void *h = nullptr;
    if (cudaHostAlloc(&h, want, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return fail(MXG_ERR_CUDA, "cudaHostAlloc of %zu MiB failed", want >> 20);
    }
    blk->ptr = h;
    blk->bytes = want;
    return MXG_OK;
}

void pinned_block_free(PinnedBlock *blk)
{
    if (!blk->ptr) return;
    if (blk->map_base) {
        cudaHostUnregister(blk->ptr);
        munmap(blk->map_base, blk->map_len);
    } else {
        cudaFreeHost(blk->ptr);
    }
    cudaGetLastError();
    *blk = PinnedBlock();
}

int pinned_arena(DeviceState *st, size_t bytes, char **base)
{
    const long cap_mb = options().host_arena_max_mb;
    if (cap_mb > 0 && bytes > ((size_t)cap_mb << 20))
        return fail(MXG_ERR_CUDA, "staging arena of %zu MiB exceeds host_arena_max_mb = %ld", bytes >> 20, cap_mb);
    if (bytes > st->pin.bytes) {
        pinned_block_free(&st->pin);
        const size_t want = (bytes + ((size_t)1 << 22) - 1) & ~(((size_t)1 << 22) - 1);
        MXG_TRY(pinned_block_alloc(want, &st->pin));
        // every page is touched already (by the host threads, or by cudaHostAlloc): the first DMA does not pay for it
    }
    *base = static_cast<char *>(st->pin.ptr);
    return MXG_OK;
}

int pinned_arena_release(DeviceState *st)
{
    pinned_block_free(&st->pin);
    return MXG_OK;
}


// ------------------------------------------------------------------------------------------------
// Page-locked RESULT memory for the glue (mxg_host_alloc / mxg_host_free).  The result of a product is allocated by the
// callee: the Rcpp glue creates the R matrix it returns.  A freshly malloc'ed 512 MB matrix consists of pages that do not
// exist yet, and filling it is bound by the kernel's page zeroing (19 - 26 GB/s on the 16-core B200 host, against a
// 52 GB/s link; profiles/r02_host_first_touch_probe.jsonl) on top of a bounce through a page-locked slot.  R lets a
// package supply the allocator of a vector (Rf_allocVector3 + R_allocator_t), so the glue allocates large results HERE:
// a pool of page-locked blocks (pinned_block_alloc above) that are recycled when R's garbage collector frees the matrix.  The device then
// writes the result straight into the R object — no slot, no host copy, no first touch.
// Blocks are rounded up to 2 MiB and reused for requests they fit with at most 25 % waste; the pool holds at most
// option "host_result_pool_mb" (4096) of free + live blocks, beyond which mxg_host_alloc fails and the glue falls back to
// R's own allocator (the bounce path).
// ------------------------------------------------------------------------------------------------
namespace {
struct PoolBlock {
    PinnedBlock mem;
    void *ptr; // = mem.ptr
    size_t bytes;
    unsigned long long stamp;
};
std::mutex g_rp_mu;
std::vector<PoolBlock> g_rp_free, g_rp_live;
size_t g_rp_bytes = 0; // free + live
unsigned long long g_rp_clock = 0;
} // namespace

int result_pool_alloc(size_t bytes, void **out)
{
    *out = nullptr;
    const size_t cap = (size_t)std::max<long>(options().host_result_pool_mb, 0) << 20;
    const size_t want = (std::max<size_t>(bytes, 1) + ((size_t)1 << 21) - 1) & ~(((size_t)1 << 21) - 1);
    if (want > cap) return fail(MXG_ERR_CUDA, "host_alloc: %zu MiB exceeds host_result_pool_mb", want >> 20);
    std::lock_guard<std::mutex> lk(g_rp_mu);
    int best = -1;
    for (size_t i = 0; i < g_rp_free.size(); i++)
        if (g_rp_free[i].bytes >= want && g_rp_free[i].bytes <= want + want / 4 &&
            (best < 0 || g_rp_free[i].bytes < g_rp_free[(size_t)best].bytes))
            best = (int)i;
    if (best >= 0) {
        PoolBlock b = g_rp_free[(size_t)best];
        g_rp_free.erase(g_rp_free.begin() + best);
        g_rp_live.push_back(b);
        *out = b.ptr;
        return MXG_OK;
    }
    // make room: the least recently freed blocks go back to the driver
    while (g_rp_bytes + want > cap && !g_rp_free.empty()) {
        size_t oldest = 0;
        for (size_t i = 1; i < g_rp_free.size(); i++)
            if (g_rp_free[i].stamp < g_rp_free[oldest].stamp) oldest = i;
        pinned_block_free(&g_rp_free[oldest].mem);
        g_rp_bytes -= g_rp_free[oldest].bytes;
        g_rp_free.erase(g_rp_free.begin() + (long)oldest);
    }
    if (g_rp_bytes + want > cap) return fail(MXG_ERR_CUDA, "host_alloc: result pool is full (%zu MiB live)", g_rp_bytes >> 20);
    PinnedBlock mem;
    MXG_TRY(pinned_block_alloc(want, &mem));
    g_rp_live.push_back(PoolBlock{mem, mem.ptr, want, 0});
    g_rp_bytes += want;
    *out = mem.ptr;
    return MXG_OK;
}

// MXG_ERR_ARG (and nothing else happens) when `ptr` is not a live block of the pool: the glue's free hook then knows the
// block came from its fallback allocator
int result_pool_free(void *ptr)
{
    std::lock_guard<std::mutex> lk(g_rp_mu);
    for (size_t i = 0; i < g_rp_live.size(); i++)
        if (g_rp_live[i].ptr == ptr) {
            PoolBlock b = g_rp_live[i];
            g_rp_live.erase(g_rp_live.begin() + (long)i);
            b.stamp = ++g_rp_clock;
            g_rp_free.push_back(b);
            return MXG_OK;
        }
    return MXG_ERR_ARG;
}

void result_pool_stats(size_t *live_bytes, size_t *free_bytes, int *blocks)
{
    std::lock_guard<std::mutex> lk(g_rp_mu);
    size_t l = 0, f = 0;
    for (const PoolBlock &b : g_rp_live) l += b.bytes;
    for (const PoolBlock &b : g_rp_free) f += b.bytes;
    if (live_bytes) *live_bytes = l;
    if (free_bytes) *free_bytes = f;
    if (blocks) *blocks = (int)(g_rp_live.size() + g_rp_free.size());
}

// mxg_trim: free blocks go back to the driver (live ones belong to the caller)
void result_pool_trim()
{
    std::lock_guard<std::mutex> lk(g_rp_mu);
    for (PoolBlock &b : g_rp_free) {
        pinned_block_free(&b.mem);
        g_rp_bytes -= b.bytes;
    }
    g_rp_free.clear();
}

// ------------------------------------------------------------------------------------------------
// Staged one-shot copies for the entry points that are not streamed chunk by chunk (handle uploads, crossprod,
// CSR->CSC): the same bounce through the arena, 16 MiB blocks over a 4-slot ring on `stream`.  They return with
// every slot idle again (the arena is shared with the streamed calls), i.e. after the last block has been copied.
// Page-locked caller memory, or staging switched off, takes the plain cudaMemcpyAsync.
// ------------------------------------------------------------------------------------------------
namespace {

constexpr size_t ST_BLOCK = (size_t)16 << 20;
constexpr int ST_SLOTS = 4;

struct StagedRing {
    char *base = nullptr;
    cudaEvent_t ev[ST_SLOTS] = {};
    bool used[ST_SLOTS] = {};
    int made = 0;
    int init(DeviceState *st)
    {
        MXG_TRY(pinned_arena(st, ST_BLOCK * ST_SLOTS, &base));
        for (; made < ST_SLOTS; made++) MXG_CUDA_TRY(cudaEventCreateWithFlags(&ev[made], cudaEventDisableTiming));
        return MXG_OK;
    }
    ~StagedRing()
    {
        for (int i = 0; i < made; i++) {
            if (used[i]) cudaEventSynchronize(ev[i]);
            cudaEventDestroy(ev[i]);
        }
    }
};

bool use_staging(const void *host_ptr, size_t bytes)
{
    return options().host_stage != 0 && bytes >= ((size_t)1 << 20) && !host_is_pinned(host_ptr);
}

} // namespace

int staged_h2d(DeviceState *st, void *d_dst, const void *src, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return MXG_OK;
    if (!use_staging(src, bytes)) {
        MXG_CUDA_TRY(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return MXG_OK;
    }
    StagedRing ring;
    if (ring.init(st) != MXG_OK) { // no page-locked memory: the driver's own bounce
        cudaGetLastError();
        MXG_CUDA_TRY(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return MXG_OK;
    }
    int k = 0;
    for (size_t off = 0; off < bytes; off += ST_BLOCK, k = (k + 1) % ST_SLOTS) {
        const size_t len = std::min(ST_BLOCK, bytes - off);
        if (ring.used[k]) MXG_CUDA_TRY(cudaEventSynchronize(ring.ev[k]));
        char *slot = ring.base + (size_t)k * ST_BLOCK;
        host_copy(slot, static_cast<const char *>(src) + off, len, /*nt_dst=*/true);
        MXG_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(d_dst) + off, slot, len, cudaMemcpyHostToDevice, stream));
        MXG_CUDA_TRY(cudaEventRecord(ring.ev[k], stream));
        ring.used[k] = true;
    }
    return MXG_OK; // ~StagedRing waits for the copies still reading the slots
}

// float32 device array from float64 host values: narrowed by the host threads, half the bytes on the link
int staged_h2d_narrow(DeviceState *st, float *d_dst, const double *src, size_t n, cudaStream_t stream)
{
    if (n == 0) return MXG_OK;
    StagedRing ring;
    MXG_TRY(ring.init(st));
    const size_t block = ST_BLOCK / sizeof(float);
    int k = 0;
    for (size_t off = 0; off < n; off += block, k = (k + 1) % ST_SLOTS) {
        const size_t len = std::min(block, n - off);
        if (ring.used[k]) MXG_CUDA_TRY(cudaEventSynchronize(ring.ev[k]));
        float *slot = reinterpret_cast<float *>(ring.base + (size_t)k * ST_BLOCK);
        host_narrow_f64_to_f32(src + off, slot, len);
        MXG_CUDA_TRY(cudaMemcpyAsync(d_dst + off, slot, sizeof(float) * len, cudaMemcpyHostToDevice, stream));
        MXG_CUDA_TRY(cudaEventRecord(ring.ev[k], stream));
        ring.used[k] = true;
    }
    return MXG_OK;
}

// device -> pageable host: blocks land in the ring and the host threads move them out (first touch in parallel)
int staged_d2h(DeviceState *st, void *dst, const void *d_src, size_t bytes, cudaStream_t stream)
{
    if (bytes == 0) return MXG_OK;
    if (!use_staging(dst, bytes)) {
        MXG_CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, stream));
        return MXG_OK;
    }
    StagedRing ring;
    if (ring.init(st) != MXG_OK) {
        cudaGetLastError();
        MXG_CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, stream));
        return MXG_OK;
    }
    host_prepare_result(dst, bytes);
    const size_t nblocks = (bytes + ST_BLOCK - 1) / ST_BLOCK;
    auto drain = [&](size_t b) -> int {
        const int k = (int)(b % ST_SLOTS);
        const size_t off = b * ST_BLOCK, len = std::min(ST_BLOCK, bytes - off);
        MXG_CUDA_TRY(cudaEventSynchronize(ring.ev[k]));
        ring.used[k] = false;
        host_copy(static_cast<char *>(dst) + off, ring.base + (size_t)k * ST_BLOCK, len);
        return MXG_OK;
    };
    for (size_t b = 0; b < nblocks; b++) {
        const int k = (int)(b % ST_SLOTS);
        if (b >= ST_SLOTS) MXG_TRY(drain(b - ST_SLOTS)); // the slot's previous block leaves first
        const size_t off = b * ST_BLOCK, len = std::min(ST_BLOCK, bytes - off);
        MXG_CUDA_TRY(cudaMemcpyAsync(ring.base + (size_t)k * ST_BLOCK, static_cast<const char *>(d_src) + off, len,
                                     cudaMemcpyDeviceToHost, stream));
        MXG_CUDA_TRY(cudaEventRecord(ring.ev[k], stream));
        ring.used[k] = true;
    }
    for (size_t b = nblocks > ST_SLOTS ? nblocks - ST_SLOTS : 0; b < nblocks; b++) MXG_TRY(drain(b));
    return MXG_OK;
}

} // namespace mxg
