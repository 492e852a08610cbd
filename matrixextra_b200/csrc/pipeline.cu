// pipeline.cu — the streamed host-buffer (level-1) products: what one call of an Rcpp export costs end to end.
//
// The reference's entry points (src/matmul.cpp:221-483) receive host arrays and return a host matrix, so a
// GPU drop-in pays PCIe for every operand.  The CSR is by far the largest of them (12 bytes per stored entry
// on the host: int32 index + float64 value), and every output row depends on one CSR row only, so the call
// is cut into contiguous ROW CHUNKS of roughly equal nnz that flow through three streams:
//
//     h2d  : dense operand, indptr, then per chunk  indices + float64 values   (pinned or pageable source)
//     comp : per chunk  [narrow values to float32] -> validate column ids -> SpMM / SpMV on the chunk's rows
//     d2h  : per chunk  output rows -> the caller's (R-allocated) result buffer
//
// PCIe is full duplex, so uploads of chunk c+1, kernels of chunk c and downloads of chunk c-1 overlap and the
// call costs about max(H2D bytes, D2H bytes) / link rate instead of their sum plus the kernel time.
// Row statistics (long-row piece tables, K7) are computed on the HOST from the host indptr while the first
// copies are in flight, which removes every mid-pipeline device->host synchronisation; column ids are
// validated on the device per chunk and a device flag makes the product kernels of that and all later
// chunks return immediately (no out-of-range gather ever executes); the flag is read back once at the end.
//
// Host side (hoststage.cu): R hands the glue PAGEABLE vectors and float64 values.  Pageable operands are
// bounced through ring slots of a page-locked arena by a pool of host threads (parallel memcpy in, parallel
// first-touch copy out) instead of the driver's single-threaded bounce, and for float32 products the values
// are narrowed by those threads straight into the slot, so 8 instead of 12 bytes per entry cross PCIe.
// Caller memory that is already page-locked is DMA'd in place.
#include "mxg_internal.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mxg {
namespace {

// MXG_TRACE=1: host-side timeline of a streamed call on stderr (development aid, off by default)
struct Trace {
    bool on = false;
    std::chrono::steady_clock::time_point t0;
    double wait_ms = 0, fill_ms = 0, drain_wait_ms = 0, drain_ms = 0, plan_ms = 0, finish_ms = 0;
    int packed_chunks = 0;
    Trace()
    {
        const char *e = getenv("MXG_TRACE");
        on = e && *e && *e != '0';
        t0 = std::chrono::steady_clock::now();
    }
    cudaEvent_t g[4] = {nullptr, nullptr, nullptr, nullptr}; // GPU-side marks: start, last upload, last kernel, last download
    void mark(int i, cudaStream_t s)
    {
        if (!on) return;
        if (!g[i]) cudaEventCreate(&g[i]);
        cudaEventRecord(g[i], s);
    }
    ~Trace()
    {
        for (cudaEvent_t e : g)
            if (e) cudaEventDestroy(e);
    }
    double now() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
    void report(const char *what, int chunks) const
    {
        if (!on) return;
        float up = -1, comp = -1, down = -1;
        if (g[0] && g[1] && g[2] && g[3]) {
            cudaEventSynchronize(g[3]);
            cudaEventElapsedTime(&up, g[0], g[1]);
            cudaEventElapsedTime(&comp, g[1], g[2]);
            cudaEventElapsedTime(&down, g[2], g[3]);
        }
        fprintf(stderr, "[mxg trace] %s: %d chunks (%d with packed ids), total %.2f ms | plan %.2f | slot waits %.2f | host fill %.2f | "
                        "out waits %.2f | out copies %.2f | finish %.2f | gpu: uploads %.3f, last kernel after last upload %.3f, "
                        "last download %.3f\n",
                what, chunks, packed_chunks, now(), plan_ms, wait_ms, fill_ms, drain_wait_ms, drain_ms, finish_ms, up, comp, down);
    }
};

struct HostPlan {
    int piece = 1024;
    int max_len = 0;
    std::vector<int> chunk_row;       // [C+1] first row of every chunk
    std::vector<int> chunk_long_off;  // [C+1] offsets into long_*
    std::vector<int> chunk_piece_off; // [C+1] offsets into piece_*
    // chunk-relative row ids / piece slots, all chunks concatenated
    std::vector<int32_t> long_rows, long_first, long_np, piece_row, piece_k;
    int max_chunk_pieces = 0;
    size_t max_chunk_nnz = 0;
    size_t max_chunk_rows = 0;
};

// The host indptr -> chunk boundaries, monotonicity check, long-row tables (the host twin of k_row_stats /
// k_fill_long_tables in layout.cu; same table format, rows numbered from the chunk start).  The scan over the rows
// (validity, longest row, long rows) runs on the host threads; boundaries are binary searches in the indptr.
int build_plan(int m, const int32_t *p, HostPlan &plan, size_t out_row_bytes)
{
    plan.piece = (int)std::max<long>(32, options().piece);
    const int64_t nnz = (int64_t)p[m] - (int64_t)p[0]; // a row block of a larger matrix keeps its absolute offsets
    // ~16 chunks, but never more than 16 Mi entries or 64 MiB of result rows per chunk: the ring slots of the
    // page-locked arena are sized by the largest chunk (a 2-billion-entry call must not pin gigabytes)
    int64_t target_nnz = std::min<int64_t>(std::max<int64_t>(nnz / 16, (int64_t)1 << 20), (int64_t)16 << 20);
    int target_rows = std::max(m / 16, 1 << 16);
    if (out_row_bytes > 0) {
        const int64_t cap = std::max<int64_t>(1 << 16, ((int64_t)64 << 20) / (int64_t)out_row_bytes);
        target_rows = (int)std::min<int64_t>(target_rows, cap);
    }
    const bool forced = options().pipe_chunk_nnz > 0; // tests: many small chunks of exactly this size
    if (forced) {
        target_nnz = options().pipe_chunk_nnz;
        target_rows = (int)std::min<int64_t>(target_nnz, INT32_MAX);
    }

    // 1. scan: every task owns a range of rows and reports its long rows in row order
    struct Part {
        int bad = 0, max_len = 0;
        std::vector<int> long_rows;
    };
    const size_t grain = (size_t)1 << 16;
    const size_t ntasks = ((size_t)m + grain - 1) / grain;
    std::vector<Part> parts(ntasks);
    const int piece = plan.piece;
    host_parallel_for(ntasks, (size_t)m * sizeof(int32_t), [&](size_t t) {
        Part &q = parts[t];
        const int r0 = (int)(t * grain), r1 = (int)std::min<size_t>((size_t)m, (t + 1) * grain);
        int bad = 0, mx = 0;
        for (int r = r0; r < r1; r++) {
            const int a = p[r], b = p[r + 1];
            bad |= (a < 0) | (b < a);
            const int len = b - a;
            mx = std::max(mx, len);
            if (len > piece && !bad) q.long_rows.push_back(r);
        }
        q.bad = bad;
        q.max_len = mx;
    });
    for (const Part &q : parts) {
        if (q.bad) return fail(MXG_ERR_INDEX, "CSR indptr is negative or decreasing");
        plan.max_len = std::max(plan.max_len, q.max_len);
    }

    // 2. boundaries: a chunk ends at the first row that brings it to the target (entries or rows).  Towards the end
    // of the matrix the chunks shrink (a third of what is left, down to an eighth of the target): what follows the
    // last upload — that chunk's kernel and its download — is exposed, so the last chunk should be a small one.
    plan.chunk_row.push_back(0);
    for (int start = 0; start < m;) {
        const int64_t first = p[start];
        int64_t tn = target_nnz;
        int tr = target_rows;
        if (!forced) {
            tn = std::min(tn, std::max<int64_t>(target_nnz / 8, ((int64_t)p[m] - first) / 3));
            tr = (int)std::min<int64_t>(tr, std::max<int64_t>(target_rows / 8, (int64_t)(m - start) / 3));
        }
        const int32_t *hit = std::lower_bound(p + start + 1, p + m + 1, first + tn,
                                              [](int32_t v, int64_t want) { return (int64_t)v < want; });
        int end = (int)std::min<int64_t>(hit - p, (int64_t)start + tr);
        end = std::min(std::max(end, start + 1), m);
        plan.chunk_row.push_back(end);
        plan.max_chunk_nnz = std::max(plan.max_chunk_nnz, (size_t)((int64_t)p[end] - first));
        plan.max_chunk_rows = std::max(plan.max_chunk_rows, (size_t)(end - start));
        start = end;
    }

    // 3. long-row tables per chunk
    plan.chunk_long_off.push_back(0);
    plan.chunk_piece_off.push_back(0);
    size_t c = 0; // chunk of the current long row
    int pieces_in_chunk = 0;
    auto close_chunk = [&]() {
        plan.chunk_long_off.push_back((int)plan.long_rows.size());
        plan.chunk_piece_off.push_back((int)plan.piece_row.size());
        plan.max_chunk_pieces = std::max(plan.max_chunk_pieces, pieces_in_chunk);
        pieces_in_chunk = 0;
        c++;
    };
    const size_t C = plan.chunk_row.size() - 1;
    for (const Part &q : parts)
        for (int r : q.long_rows) {
            while (r >= plan.chunk_row[c + 1]) close_chunk();
            const int chunk_start = plan.chunk_row[c];
            const int len = p[r + 1] - p[r];
            const int np = (len + plan.piece - 1) / plan.piece;
            plan.long_rows.push_back(r - chunk_start);
            plan.long_first.push_back(pieces_in_chunk);
            plan.long_np.push_back(np);
            for (int k = 0; k < np; k++) {
                plan.piece_row.push_back(r - chunk_start);
                plan.piece_k.push_back(k);
            }
            pieces_in_chunk += np;
        }
    while (c < C) close_chunk();
    return MXG_OK;
}

// everything a call allocates, released on every exit path
struct Scratch {
    DeviceState *st;
    std::vector<void *> bufs;
    std::vector<cudaEvent_t> events;
    explicit Scratch(DeviceState *s) : st(s) {}
    int alloc(void **ptr, size_t bytes)
    {
        MXG_CUDA_TRY(cudaMallocAsync(ptr, std::max<size_t>(bytes, 16), st->stream));
        bufs.push_back(*ptr);
        return MXG_OK;
    }
    int event(cudaEvent_t *ev)
    {
        MXG_CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        events.push_back(*ev);
        return MXG_OK;
    }
    ~Scratch()
    {
        cudaStreamSynchronize(st->h2d);
        cudaStreamSynchronize(st->d2h);
        for (void *q : bufs) cudaFreeAsync(q, st->stream);
        cudaStreamSynchronize(st->stream);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
};

// `later` waits for everything enqueued on `earlier` so far
int chain(Scratch &sc, cudaStream_t earlier, cudaStream_t later)
{
    cudaEvent_t ev;
    MXG_TRY(sc.event(&ev));
    MXG_CUDA_TRY(cudaEventRecord(ev, earlier));
    MXG_CUDA_TRY(cudaStreamWaitEvent(later, ev, 0));
    return MXG_OK;
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// bytes the most recent streamed call of this thread moved over PCIe (mxg_last_call_bytes; bench.py reports them)
thread_local size_t g_h2d_bytes = 0, g_d2h_bytes = 0;

cudaError_t copy_async(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream)
{
    (kind == cudaMemcpyHostToDevice ? g_h2d_bytes : g_d2h_bytes) += bytes;
    return cudaMemcpyAsync(dst, src, bytes, kind, stream);
}

int copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
              cudaStream_t stream)
{
    if (width == 0 || height == 0) return MXG_OK;
    if (dpitch == width && spitch == width) MXG_CUDA_TRY(copy_async(dst, src, width * height, kind, stream));
    else {
        (kind == cudaMemcpyHostToDevice ? g_h2d_bytes : g_d2h_bytes) += width * height;
        MXG_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, stream));
    }
    return MXG_OK;
}

// Ring of page-locked slots for host -> device traffic.  acquire() hands out the next slot once the copy that
// last read it has finished; release(stream) marks it busy until everything enqueued on `stream` so far is done.
thread_local Trace *g_trace = nullptr;

struct InRing {
    char *base = nullptr;
    size_t slot_bytes = 0;
    int slots = 0, cur = -1, next = 0;
    std::vector<cudaEvent_t> busy;
    std::vector<char> used;
    bool enabled() const { return slots > 0 && slot_bytes > 0; }
    int init(Scratch &sc, char *mem, size_t bytes_per_slot, int n)
    {
        base = mem;
        slot_bytes = bytes_per_slot;
        slots = n;
        busy.resize((size_t)n);
        used.assign((size_t)n, 0);
        for (int i = 0; i < n; i++) MXG_TRY(sc.event(&busy[(size_t)i]));
        return MXG_OK;
    }
    int acquire(char **slot)
    {
        cur = next;
        next = (next + 1) % slots;
        const double t = g_trace ? g_trace->now() : 0;
        if (used[(size_t)cur]) MXG_CUDA_TRY(cudaEventSynchronize(busy[(size_t)cur]));
        if (g_trace) g_trace->wait_ms += g_trace->now() - t;
        *slot = base + (size_t)cur * slot_bytes;
        return MXG_OK;
    }
    int release(cudaStream_t stream)
    {
        MXG_CUDA_TRY(cudaEventRecord(busy[(size_t)cur], stream));
        used[(size_t)cur] = 1;
        return MXG_OK;
    }
};

// host -> device copy of `height` lines of `width` bytes; a pageable source goes through the ring block by block
int upload_lines(InRing *ring, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                 cudaStream_t stream)
{
    if (width == 0 || height == 0) return MXG_OK;
    if (!ring || !ring->enabled() || width > ring->slot_bytes)
        return copy_rows(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, stream);
    const size_t lines = ring->slot_bytes / width;
    for (size_t l0 = 0; l0 < height; l0 += lines) {
        const size_t nl = std::min(lines, height - l0);
        char *slot = nullptr;
        MXG_TRY(ring->acquire(&slot));
        host_copy_2d(slot, width, static_cast<const char *>(src) + l0 * spitch, spitch, width, nl, /*nt_dst=*/true);
        MXG_TRY(copy_rows(static_cast<char *>(dst) + l0 * dpitch, dpitch, slot, width, width, nl, cudaMemcpyHostToDevice, stream));
        MXG_TRY(ring->release(stream));
    }
    return MXG_OK;
}

// state shared by the SpMM and SpMV pipelines: device CSR arrays, chunk plan, long-row tables
struct CsrStream {
    Scratch &sc;
    HostPlan plan;
    int m, K;
    int64_t nnz;
    int64_t base; // p[0]: host arrays are indexed with the absolute offsets of p, device arrays start at this entry
    const int32_t *p, *j;
    const double *x;
    bool narrow; // the product runs on float32 values
    bool multi_call = false; // one of several device pipelines of a multi-device call
    // host staging (hoststage.cu)
    bool staging_unavailable = false; // the arena could not be allocated: every copy takes the driver's path
    bool narrow_on_host = false; // float32 values are produced by the host threads, no device narrowing
    bool stage_x = false, stage_j = false;
    // column ids packed by the host threads (2 / 2.5 / 3 bytes per entry) and unpacked on the device: decided per
    // chunk (pack_mode 1: while the link is the bottleneck; 2: always; 3: two chunks out of three)
    bool pack_j = false, pack_pageable = false;
    int pack_mode = 0, hi_bits = -1;
    std::vector<char> chunk_packed; // [C]
    int pack_buf_last[3] = {-1, -1, -1}; // the last packed chunk that used d_pack[b]
    size_t x_part = 0; // bytes of a slot reserved for the values (indices follow)
    InRing ring;       // slots for chunk uploads (also used for the dense operand before the first chunk)
    char *out_base = nullptr; // ring of output slots (one per chunk, c % out_slots)
    size_t out_slot_bytes = 0;
    int out_slots = 0;
    int32_t *d_p = nullptr, *d_j = nullptr;
    double *d_x64 = nullptr;
    float *d_x32 = nullptr;
    static constexpr int NSTAGE = 3;
    double *d_stage[NSTAGE] = {nullptr, nullptr, nullptr};
    char *d_pack[NSTAGE] = {nullptr, nullptr, nullptr};
    int32_t *d_tables = nullptr;
    int *d_flag = nullptr;
    void *d_partial = nullptr;
    size_t partial_bytes = 0;
    std::vector<cudaEvent_t> ev_h2d, ev_conv, ev_unpack;

    CsrStream(Scratch &s, int m_, int K_, const int32_t *p_, const int32_t *j_, const double *x_, bool narrow_)
        : sc(s), m(m_), K(K_), nnz((int64_t)p_[m_] - (int64_t)p_[0]), base(p_[0]), p(p_), j(j_), x(x_), narrow(narrow_)
    {
    }
    int chunks() const { return (int)plan.chunk_row.size() - 1; }

    // Plan + page-locked arena.  dense_pageable: the dense operand wants ring slots too; out_row_bytes > 0:
    // the result is pageable and every chunk's rows (out_row_bytes each) are bounced through an output slot.
    int plan_host(bool dense_pageable, size_t out_row_bytes, size_t result_row_bytes)
    {
        MXG_TRY(build_plan(m, p, plan, result_row_bytes));
        const bool stage = options().host_stage != 0;
        // float32 products: the host threads narrow the float64 values on their way into the slot (8 instead of 12
        // PCIe bytes per entry).  Not when this call is one of several device pipelines fed from page-locked arrays
        // (host_narrow = 1, automatic): the devices then share the host's memory system rather than one link, and
        // narrowing costs it 12 bytes of traffic per entry that a plain DMA of the float64 values does not
        // (measured, 8 devices: 127 ms with host narrowing, 100 ms without).  host_narrow = 2: always, 0: never.
        const long hn = options().host_narrow;
        narrow_on_host = narrow && nnz > 0 && (hn == 2 || (hn == 1 && !(multi_call && host_is_pinned(x))));
        stage_x = nnz > 0 && (narrow_on_host || (stage && !host_is_pinned(x)));
        // Packed column ids pay while the call is PCIe-bound and the host threads have time to spare; a packed chunk
        // costs them more memory traffic (read 4, write 2-3 bytes per entry), so when they are also narrowing values
        // or bouncing a result they can become the bottleneck instead.  upload_chunk() therefore decides chunk by
        // chunk from the state of the upload stream (profiles/r01_v7_e2e_packed_ids.jsonl).  Only for long calls and
        // a full pool, and not when the same threads narrow the values of a float32 product: there the fills are
        // already within a third of the link time, packing makes them the critical path, and the outcome depends on
        // the box (measured: -3 % on hosts that narrow 100 M values in 11.8 ms, +3..6 % on one that takes 16.5 ms).
        // host_pack = 2 packs every chunk whatever the size (tests).
        hi_bits = index_pack_hi_bits(K);
        pack_mode = (int)options().host_pack;
        stage_j = nnz > 0 && stage && !host_is_pinned(j); // unpacked ids are bounced through the slot
        // pageable ids have to be read and rewritten by the host threads anyway (the bounce): packing them instead
        // costs no extra pass, writes 2 - 3 bytes per entry instead of 4 and shortens the upload — always on for them
        pack_pageable = pack_mode == 1 && stage_j && nnz >= ((int64_t)1 << 20);
        pack_j = stage && hi_bits >= 0 && nnz > 0 &&
                 (pack_mode == 2 || pack_mode == 3 || pack_pageable ||
                  (pack_mode == 1 && !narrow_on_host && nnz >= ((int64_t)1 << 20) && host_threads() >= 8));
        auto up = [](size_t v) { return (v + 4095) & ~(size_t)4095; };
        x_part = stage_x ? up(plan.max_chunk_nnz * (narrow_on_host ? sizeof(float) : sizeof(double))) : 0;
        size_t in_slot = x_part + std::max(pack_j ? up(packed_index_bytes(plan.max_chunk_nnz, hi_bits)) : 0,
                                           stage_j ? up(plan.max_chunk_nnz * sizeof(int32_t)) : 0);
        if (dense_pageable && stage) in_slot = std::max(in_slot, (size_t)16 << 20);
        const int S = (int)std::min<long>(std::max<long>(options().pipe_slots, 3), 8);
        out_slots = out_row_bytes > 0 && stage ? S : 0;
        out_slot_bytes = out_slots ? up(plan.max_chunk_rows * out_row_bytes) : 0;
        const size_t total = (size_t)S * in_slot + (size_t)out_slots * out_slot_bytes;
        if (total > 0) {
            char *base = nullptr;
            if (pinned_arena(sc.st, total, &base) != MXG_OK) {
                // no page-locked memory to be had (ulimit -l, fragmentation): the plain driver copies still work
                cudaGetLastError();
                last_error_ref().clear();
                narrow_on_host = stage_x = stage_j = pack_j = false;
                x_part = 0;
                out_slots = 0;
                out_slot_bytes = 0;
                staging_unavailable = true;
                return MXG_OK;
            }
            if (in_slot > 0) MXG_TRY(ring.init(sc, base, in_slot, S));
            out_base = base + (size_t)S * in_slot;
        }
        return MXG_OK;
    }
    char *out_slot(int c) const { return out_base + (size_t)(c % out_slots) * out_slot_bytes; }

    // device allocations + indptr upload + table upload.  partial_per_piece = workspace bytes per piece.
    int begin(size_t partial_per_piece)
    {
        DeviceState *st = sc.st;
        const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
        MXG_TRY(sc.alloc((void **)&d_p, sizeof(int32_t) * ((size_t)m + 1)));
        MXG_TRY(sc.alloc((void **)&d_j, sizeof(int32_t) * nz));
        if (narrow) MXG_TRY(sc.alloc((void **)&d_x32, sizeof(float) * nz));
        else MXG_TRY(sc.alloc((void **)&d_x64, sizeof(double) * nz));
        MXG_TRY(sc.alloc((void **)&d_flag, sizeof(int)));
        MXG_CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), st->stream));
        MXG_TRY(chain(sc, st->stream, st->h2d));
        MXG_CUDA_TRY(copy_async(d_p, p, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice, st->h2d));
        const size_t nl = plan.long_rows.size(), np = plan.piece_row.size();
        if (nl > 0) {
            MXG_TRY(sc.alloc((void **)&d_tables, sizeof(int32_t) * (3 * nl + 2 * np)));
            MXG_TRY(sc.alloc(&d_partial, (size_t)plan.max_chunk_pieces * partial_per_piece));
            partial_bytes = (size_t)plan.max_chunk_pieces * partial_per_piece;
            MXG_TRY(chain(sc, st->stream, st->h2d));
            const std::vector<int32_t> *src[5] = {&plan.long_rows, &plan.long_first, &plan.long_np, &plan.piece_row, &plan.piece_k};
            size_t off = 0;
            for (int t = 0; t < 5; t++) {
                MXG_CUDA_TRY(copy_async(d_tables + off, src[t]->data(), sizeof(int32_t) * src[t]->size(),
                                             cudaMemcpyHostToDevice, st->h2d));
                off += src[t]->size();
            }
        }
        if (narrow && !narrow_on_host) {
            const size_t stage_n = plan.max_chunk_nnz + (plan.max_chunk_nnz & 1);
            for (int b = 0; b < NSTAGE && b < chunks(); b++) MXG_TRY(sc.alloc((void **)&d_stage[b], sizeof(double) * stage_n));
            MXG_TRY(chain(sc, st->stream, st->h2d));
        }
        if (pack_j) {
            for (int b = 0; b < NSTAGE && b < chunks(); b++)
                MXG_TRY(sc.alloc((void **)&d_pack[b], packed_index_bytes(plan.max_chunk_nnz, hi_bits)));
            MXG_TRY(chain(sc, st->stream, st->h2d));
        }
        ev_h2d.resize((size_t)chunks());
        ev_conv.resize((size_t)chunks());
        ev_unpack.resize((size_t)chunks());
        chunk_packed.assign((size_t)chunks(), 0);
        for (int c = 0; c < chunks(); c++) {
            MXG_TRY(sc.event(&ev_h2d[(size_t)c]));
            if (narrow && !narrow_on_host) MXG_TRY(sc.event(&ev_conv[(size_t)c]));
            if (pack_j) MXG_TRY(sc.event(&ev_unpack[(size_t)c]));
        }
        return MXG_OK;
    }

    // h2d stream: indices and values of chunk c (through a ring slot where the host threads are involved)
    int upload_chunk(int c)
    {
        DeviceState *st = sc.st;
        const int64_t e0 = p[plan.chunk_row[(size_t)c]], e1 = p[plan.chunk_row[(size_t)c + 1]];
        const size_t len = (size_t)(e1 - e0);
        if (len > 0) {
            char *slot = nullptr;
            // Pack this chunk's ids?  Yes while uploads are queueing up (the chunk before the previous one has not
            // arrived yet: the link is behind the host); no when the link is about to run dry (the host is behind).
            // The first two chunks queue behind the dense operand.
            // (pack_mode 3, tests: a fixed mix — two chunks out of three)
            const int lag = (int)std::min<long>(std::max<long>(options().host_pack_lag, 1), 8);
            const bool pk = pack_j && (pack_mode == 2 || pack_pageable || (pack_mode == 3 ? c % 3 != 1
                                                         : (c < lag || cudaEventQuery(ev_h2d[(size_t)(c - lag)]) == cudaErrorNotReady)));
            cudaGetLastError(); // cudaErrorNotReady is an answer, not an error
            chunk_packed[(size_t)c] = pk;
            if (stage_x || stage_j || pk) MXG_TRY(ring.acquire(&slot));
            // values: float32 made on the host -> d_x32; float64 -> d_x64, or -> device staging + narrowing kernel
            const void *xsrc = x + e0;
            const double tf = g_trace ? g_trace->now() : 0;
            if (narrow_on_host) {
                host_narrow_f64_to_f32(x + e0, reinterpret_cast<float *>(slot), len);
                if (g_trace) g_trace->fill_ms += g_trace->now() - tf;
                MXG_CUDA_TRY(copy_async(d_x32 + (e0 - base), slot, sizeof(float) * len, cudaMemcpyHostToDevice, st->h2d));
            } else {
                if (stage_x) {
                    host_copy(slot, x + e0, sizeof(double) * len, /*nt_dst=*/true);
                    xsrc = slot;
                    if (g_trace) g_trace->fill_ms += g_trace->now() - tf;
                }
                if (narrow) {
                    // the device staging buffer is free again once the narrowing of chunk c - NSTAGE has run
                    if (c >= NSTAGE) MXG_CUDA_TRY(cudaStreamWaitEvent(st->h2d, ev_conv[(size_t)(c - NSTAGE)], 0));
                    MXG_CUDA_TRY(copy_async(d_stage[c % NSTAGE], xsrc, sizeof(double) * len, cudaMemcpyHostToDevice, st->h2d));
                } else {
                    MXG_CUDA_TRY(copy_async(d_x64 + (e0 - base), xsrc, sizeof(double) * len, cudaMemcpyHostToDevice, st->h2d));
                }
            }
            const void *jsrc = j + e0;
            if (pk) {
                const double tj = g_trace ? g_trace->now() : 0;
                if (!host_pack_indices(j + e0, len, K, hi_bits, slot + x_part))
                    return fail(MXG_ERR_INDEX, "CSR column index outside [0, %d)", K);
                if (g_trace) {
                    g_trace->fill_ms += g_trace->now() - tj;
                    g_trace->packed_chunks++;
                }
                // the device staging buffer is free again once the chunk that used it last has been unpacked
                const int b = c % NSTAGE;
                if (pack_buf_last[b] >= 0) MXG_CUDA_TRY(cudaStreamWaitEvent(st->h2d, ev_unpack[(size_t)pack_buf_last[b]], 0));
                pack_buf_last[b] = c;
                MXG_CUDA_TRY(copy_async(d_pack[b], slot + x_part, packed_index_bytes(len, hi_bits), cudaMemcpyHostToDevice, st->h2d));
            } else if (stage_j) {
                const double tj = g_trace ? g_trace->now() : 0;
                host_copy(slot + x_part, j + e0, sizeof(int32_t) * len, /*nt_dst=*/true);
                jsrc = slot + x_part;
                if (g_trace) g_trace->fill_ms += g_trace->now() - tj;
            }
            if (!pk) MXG_CUDA_TRY(copy_async(d_j + (e0 - base), jsrc, sizeof(int32_t) * len, cudaMemcpyHostToDevice, st->h2d));
            if (slot) MXG_TRY(ring.release(st->h2d));
        }
        MXG_CUDA_TRY(cudaEventRecord(ev_h2d[(size_t)c], st->h2d));
        return MXG_OK;
    }

    // compute stream: wait for chunk c, narrow its values, validate its column ids, describe it as a handle
    int prepare_chunk(int c, mxg_csr_s &h)
    {
        DeviceState *st = sc.st;
        const int r0 = plan.chunk_row[(size_t)c], r1 = plan.chunk_row[(size_t)c + 1];
        const int64_t e0 = p[r0], e1 = p[r1];
        const size_t len = (size_t)(e1 - e0);
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->stream, ev_h2d[(size_t)c], 0));
        if (narrow && !narrow_on_host) {
            if (len > 0) {
                // chunk starts are not always even: narrow element-wise from the staging buffer's start
                MXG_TRY(convert_f64_to_f32(d_stage[c % NSTAGE], d_x32 + (e0 - base), len, st->stream));
            }
            MXG_CUDA_TRY(cudaEventRecord(ev_conv[(size_t)c], st->stream));
        }
        if (chunk_packed[(size_t)c]) {
            MXG_TRY(unpack_indices_flag(len, d_pack[c % NSTAGE], hi_bits, K, d_j + (e0 - base), d_flag, st->stream));
            MXG_CUDA_TRY(cudaEventRecord(ev_unpack[(size_t)c], st->stream));
        } else {
            MXG_TRY(check_indices_flag(len, d_j + (e0 - base), K, d_flag, st->stream));
        }
        const size_t nl = plan.long_rows.size(), np = plan.piece_row.size();
        const int l0 = plan.chunk_long_off[(size_t)c], l1 = plan.chunk_long_off[(size_t)c + 1];
        const int q0 = plan.chunk_piece_off[(size_t)c], q1 = plan.chunk_piece_off[(size_t)c + 1];
        h = mxg_csr_s();
        cudaGetDevice(&h.device);
        h.m = r1 - r0;
        h.K = K;
        h.nnz = e1 - e0;
        h.base = (int32_t)e0;
        h.d_p = d_p + r0; // offsets stay absolute: the kernels index d_j / d_x with them directly, so the array
        h.d_j = d_j - base; // origins handed to them are those of entry 0 (never dereferenced below entry `base`)
        h.d_x64 = d_x64 ? d_x64 - base : nullptr;
        h.d_x32 = d_x32 ? d_x32 - base : nullptr;
        h.owns = false;
        h.stream = st->stream;
        h.piece = plan.piece;
        h.max_len = plan.max_len;
        h.n_long = l1 - l0;
        h.n_pieces = q1 - q0;
        if (h.n_long > 0) {
            h.d_long_rows = d_tables + l0;
            h.d_long_first = d_tables + nl + l0;
            h.d_long_np = d_tables + 2 * nl + l0;
            h.d_piece_row = d_tables + 3 * nl + q0;
            h.d_piece_k = d_tables + 3 * nl + np + q0;
        }
        h.d_partial = d_partial;
        h.partial_bytes = partial_bytes;
        h.d_abort = d_flag;
        return MXG_OK;
    }

    // The device CSR of a finished, valid call becomes an owned handle instead of being released (level-1 cache).
    int detach_handle(mxg_csr_s **out)
    {
        mxg_csr_s *h = new mxg_csr_s();
        h->m = m;
        h->K = K;
        h->nnz = nnz;
        h->base = 0;
        h->d_p = d_p;
        h->d_j = d_j;
        h->d_x64 = d_x64;
        h->d_x32 = d_x32;
        h->owns = true;
        h->stream = sc.st->stream;
        cudaGetDevice(&h->device);
        for (const void *q : {(const void *)d_p, (const void *)d_j, (const void *)d_x64, (const void *)d_x32})
            if (q) sc.bufs.erase(std::remove(sc.bufs.begin(), sc.bufs.end(), const_cast<void *>(q)), sc.bufs.end());
        const int rc = csr_build_stats(h, /*validate=*/0, sc.st->stream);
        if (rc != MXG_OK) {
            for (const void *q : {(const void *)d_p, (const void *)d_j, (const void *)d_x64, (const void *)d_x32})
                if (q) cudaFreeAsync(const_cast<void *>(q), sc.st->stream);
            delete h;
            return rc;
        }
        *out = h;
        return MXG_OK;
    }

    // after the last chunk: read the validation flag back and wait for every stream
    int finish()
    {
        DeviceState *st = sc.st;
        int flag = 0;
        MXG_CUDA_TRY(copy_async(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st->stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(st->d2h));
        MXG_CUDA_TRY(cudaStreamSynchronize(st->h2d));
        if (flag) return fail(MXG_ERR_INDEX, "CSR column index outside [0, %d)", K);
        return MXG_OK;
    }
};

} // namespace

void last_call_bytes(size_t *h2d, size_t *d2h)
{
    if (h2d) *h2d = g_h2d_bytes;
    if (d2h) *d2h = g_d2h_bytes;
}

void set_last_call_bytes(size_t h2d, size_t d2h)
{
    g_h2d_bytes = h2d;
    g_d2h_bytes = d2h;
}

// the chunk plan of a streamed call, for inspection (mxg_host_chunk_plan; no device involved)
int host_chunk_plan(int m, const int32_t *p, size_t result_row_bytes, int32_t *chunk_rows, int cap, int *n_chunks, int *n_long,
                    int *n_pieces, int *max_len)
{
    HostPlan plan;
    MXG_TRY(build_plan(m, p, plan, result_row_bytes));
    const int C = (int)plan.chunk_row.size() - 1;
    if (n_chunks) *n_chunks = C;
    if (n_long) *n_long = (int)plan.long_rows.size();
    if (n_pieces) *n_pieces = (int)plan.piece_row.size();
    if (max_len) *max_len = plan.max_len;
    if (chunk_rows) {
        if (cap < C + 1) return fail(MXG_ERR_ARG, "chunk_plan: %d boundaries do not fit %d slots", C + 1, cap);
        for (int c = 0; c <= C; c++) chunk_rows[c] = plan.chunk_row[(size_t)c];
    }
    return MXG_OK;
}

int pipeline_spmm(DeviceState *st, int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p,
                  const int32_t *j, const double *x, const void *B, size_t ldb, void *Out, size_t ldc, DenseShare *share,
                  int share_rank, mxg_csr_s **keep)
{
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    const size_t vec = 16 / s;
    const size_t rows = (size_t)m, Kz = (size_t)K, nz = (size_t)n;
    if (m == 0 || n == 0) return MXG_OK;
    if (!Out) return fail(MXG_ERR_ARG, "output is NULL");
    if (!B && K > 0) return fail(MXG_ERR_ARG, "dense operand is NULL");
    if (b_layout != MXG_ROWS_CONTIGUOUS && b_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "dense operand: bad layout %d", b_layout);
    if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "bad out_layout %d", out_layout);
    if (b_layout == MXG_ROWS_CONTIGUOUS && ldb < nz) return fail(MXG_ERR_ARG, "dense operand: ldb < n");
    if (b_layout == MXG_COLS_CONTIGUOUS && ldb < Kz) return fail(MXG_ERR_ARG, "dense operand: ldb < K");
    if (out_layout == MXG_ROWS_CONTIGUOUS && ldc < nz) return fail(MXG_ERR_ARG, "output: ldc < n");
    if (out_layout == MXG_COLS_CONTIGUOUS && ldc < rows) return fail(MXG_ERR_ARG, "output: ldc < m");
    const int64_t nnz = (int64_t)p[m] - (int64_t)p[0];
    if (nnz > 0 && (!j || !x)) return fail(MXG_ERR_ARG, "csr: indices / values is NULL");
    if (keep && p[0] != 0) return fail(MXG_ERR_ARG, "pipeline: a kept handle needs p[0] == 0");
    // one device of a multi-device call: only a rows-contiguous operand is cut into slices
    if (share && (b_layout != MXG_ROWS_CONTIGUOUS || K < share->G)) return fail(MXG_ERR_ARG, "pipeline: dense operand cannot be shared");

    g_h2d_bytes = g_d2h_bytes = 0;
    Trace trace;
    struct TraceScope {
        explicit TraceScope(Trace *t) { g_trace = t->on ? t : nullptr; }
        ~TraceScope() { g_trace = nullptr; }
    } trace_scope(&trace);
    Scratch sc(st);
    CsrStream cs(sc, m, K, p, j, x, /*narrow=*/dtype == MXG_F32);
    cs.multi_call = multi_devices() > 1 && multi_wanted_now();
    const bool stage = options().host_stage != 0;
    const bool stage_B = stage && K > 0 && !host_is_pinned(B);
    bool stage_out = stage && !host_is_pinned(Out);
    if (stage_out) host_prepare_result(Out, (out_layout == MXG_ROWS_CONTIGUOUS ? (rows - 1) * ldc + nz : (nz - 1) * ldc + rows) * s);
    // dense operand first: every chunk needs all of it.  Device copy is rows-contiguous [K][ld_b].
    const size_t ld_b = round_up(nz, vec);
    char *d_B = nullptr, *d_Out = nullptr, *d_tmp = nullptr;
    if (share) {
        // one device of a multi-device call: the operand's copy lives in the device's peer-visible buffer
        const size_t need = std::max<size_t>(Kz * ld_b * s, 16);
        if (st->share_bytes < need) {
            MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
            if (st->share_buf) MXG_CUDA_TRY(cudaFree(st->share_buf));
            st->share_buf = nullptr;
            st->share_bytes = 0;
            MXG_CUDA_TRY(cudaMalloc(&st->share_buf, need));
            st->share_bytes = need;
        }
        d_B = static_cast<char *>(st->share_buf);
    } else {
        MXG_TRY(sc.alloc((void **)&d_B, Kz * ld_b * s));
    }
    const size_t ld_o = out_layout == MXG_ROWS_CONTIGUOUS ? round_up(nz, vec) : rows;
    MXG_TRY(sc.alloc((void **)&d_Out, out_layout == MXG_ROWS_CONTIGUOUS ? rows * ld_o * s : rows * nz * s));
    if (K > 0 && b_layout == MXG_COLS_CONTIGUOUS) MXG_TRY(sc.alloc((void **)&d_tmp, Kz * nz * s));
    if (K > 0 && ld_b != nz) MXG_CUDA_TRY(cudaMemsetAsync(d_B, 0, Kz * ld_b * s, st->stream));
    MXG_TRY(chain(sc, st->stream, st->h2d));
    trace.mark(0, st->h2d);
    // rows [k0, k1) of the dense operand cross this device's PCIe link (all of them unless the call is shared)
    auto slice = [&](int q, size_t &k0, size_t &k1) {
        k0 = Kz * (size_t)q / (size_t)share->G;
        k1 = Kz * (size_t)(q + 1) / (size_t)share->G;
    };
    auto upload_dense = [&](InRing *ring) -> int {
        if (K == 0) return MXG_OK;
        if (share) {
            size_t k0, k1;
            slice(share_rank, k0, k1);
            return upload_lines(ring, d_B + k0 * ld_b * s, ld_b * s, static_cast<const char *>(B) + k0 * ldb * s, ldb * s, nz * s,
                                k1 - k0, st->h2d);
        }
        if (b_layout == MXG_ROWS_CONTIGUOUS) return upload_lines(ring, d_B, ld_b * s, B, ldb * s, nz * s, Kz, st->h2d);
        return upload_lines(ring, d_tmp, Kz * s, B, ldb * s, Kz * s, nz, st->h2d);
    };
    // a page-locked operand starts to fly before the host pass over the indptr, a pageable one needs the ring first
    if (!stage_B) MXG_TRY(upload_dense(nullptr));
    const double t_plan = trace.now();
    MXG_TRY(cs.plan_host(stage_B, stage_out ? nz * s : 0, nz * s));
    trace.plan_ms = trace.now() - t_plan;
    if (cs.staging_unavailable) stage_out = false;
    if (stage_B) MXG_TRY(upload_dense(&cs.ring));
    if (share) {
        // publish this device's copy and the event behind its slice, then pull the peers' slices over NVLink on a
        // stream of their own (the CSR chunks keep the PCIe upload stream busy meanwhile)
        MXG_CUDA_TRY(cudaEventRecord(share->slice_ready[share_rank], st->h2d));
        share->d_B[share_rank] = d_B;
        if (!share->wait(0, share_rank)) return fail(MXG_ERR_CUDA, "multi-device product: another device failed");
        for (int dq = 1; dq < share->G; dq++) {
            const int q = (share_rank + dq) % share->G; // every device starts with a different peer
            size_t k0, k1;
            slice(q, k0, k1);
            MXG_CUDA_TRY(cudaStreamWaitEvent(st->p2p, share->slice_ready[q], 0));
            MXG_CUDA_TRY(cudaMemcpyPeerAsync(d_B + k0 * ld_b * s, share->device[share_rank],
                                             static_cast<char *>(share->d_B[q]) + k0 * ld_b * s, share->device[q],
                                             (k1 - k0) * ld_b * s, st->p2p));
        }
        MXG_TRY(chain(sc, st->p2p, st->stream));
    }
    MXG_TRY(chain(sc, st->h2d, st->stream));
    if (K > 0 && b_layout == MXG_COLS_CONTIGUOUS)
        MXG_TRY(launch_transpose_dense((int)s, nz, Kz, d_tmp, Kz, d_B, ld_b, st->stream)); // [n][K] -> [K][ld_b]

    MXG_TRY(cs.begin(nz * s));
    const int C = cs.chunks();
    std::vector<cudaEvent_t> ev_done((size_t)C), ev_out((size_t)(stage_out ? C : 0));
    for (int c = 0; c < C; c++) MXG_TRY(sc.event(&ev_done[(size_t)c]));
    for (size_t c = 0; c < ev_out.size(); c++) MXG_TRY(sc.event(&ev_out[c]));

    // geometry of chunk c's block of the result: `height` lines of `width` bytes
    const bool rm = out_layout == MXG_ROWS_CONTIGUOUS;
    auto process = [&](int c) -> int {
        mxg_csr_s h;
        MXG_TRY(cs.prepare_chunk(c, h));
        const size_t r0 = (size_t)cs.plan.chunk_row[(size_t)c], nr = (size_t)h.m;
        char *d_o = d_Out + (rm ? r0 * ld_o * s : r0 * s);
        if (c == C - 1) trace.mark(1, st->h2d);
        MXG_TRY(launch_spmm(&h, dtype, out_layout, n, d_B, ld_b, d_o, ld_o, st->stream));
        if (h.d_seg) MXG_CUDA_TRY(cudaFreeAsync(h.d_seg, st->stream)); // the chunk's column-panel table
        if (c == C - 1) trace.mark(2, st->stream);
        MXG_CUDA_TRY(cudaEventRecord(ev_done[(size_t)c], st->stream));
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->d2h, ev_done[(size_t)c], 0));
        const size_t width = rm ? nz * s : nr * s, height = rm ? nr : nz;
        if (stage_out) { // packed into the chunk's output slot; drain() moves it into the caller's matrix
            MXG_TRY(copy_rows(cs.out_slot(c), width, d_o, ld_o * s, width, height, cudaMemcpyDeviceToHost, st->d2h));
            MXG_CUDA_TRY(cudaEventRecord(ev_out[(size_t)c], st->d2h));
        } else {
            char *h_o = static_cast<char *>(Out) + (rm ? r0 * ldc * s : r0 * s);
            MXG_TRY(copy_rows(h_o, ldc * s, d_o, ld_o * s, width, height, cudaMemcpyDeviceToHost, st->d2h));
        }
        if (c == C - 1) trace.mark(3, st->d2h);
        return MXG_OK;
    };
    auto drain = [&](int c) -> int {
        if (!stage_out) return MXG_OK;
        const size_t r0 = (size_t)cs.plan.chunk_row[(size_t)c], nr = (size_t)cs.plan.chunk_row[(size_t)c + 1] - r0;
        const size_t width = rm ? nz * s : nr * s, height = rm ? nr : nz;
        const double t0 = trace.now();
        MXG_CUDA_TRY(cudaEventSynchronize(ev_out[(size_t)c]));
        const double t1 = trace.now();
        char *h_o = static_cast<char *>(Out) + (rm ? r0 * ldc * s : r0 * s);
        host_copy_2d(h_o, ldc * s, cs.out_slot(c), width, width, height);
        trace.drain_wait_ms += t1 - t0;
        trace.drain_ms += trace.now() - t1;
        return MXG_OK;
    };
    // chunk c is uploaded while c - 1 is computed and c - 2 leaves its output slot
    for (int c = 0; c < C + 2; c++) {
        if (c < C) MXG_TRY(cs.upload_chunk(c));
        if (c >= 1 && c <= C) MXG_TRY(process(c - 1));
        if (c >= 2) MXG_TRY(drain(c - 2));
    }
    const double t_fin = trace.now();
    int rc = cs.finish();
    trace.finish_ms = trace.now() - t_fin;
    trace.report("spmm", C);
    // nobody may release its copy of the dense operand while a peer is still pulling slices out of it
    if (share && !share->wait(1, share_rank) && rc == MXG_OK) rc = fail(MXG_ERR_CUDA, "multi-device product: another device failed");
    if (rc == MXG_OK && keep) rc = cs.detach_handle(keep);
    return rc;
}

int pipeline_spmv(DeviceState *st, int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x,
                  const void *y, void *out, mxg_csr_s **keep)
{
    if (m == 0) return MXG_OK;
    if (!out) return fail(MXG_ERR_ARG, "output is NULL");
    if (K > 0 && !y) return fail(MXG_ERR_ARG, "vector is NULL");
    const int64_t nnz = (int64_t)p[m] - (int64_t)p[0];
    if (nnz > 0 && (!j || !x)) return fail(MXG_ERR_ARG, "csr: indices / values is NULL");
    const size_t ys = ytype == MXG_Y_NUMERIC ? 8 : 4;
    const size_t os = ytype == MXG_Y_FLOAT32 ? 4 : 8;

    g_h2d_bytes = g_d2h_bytes = 0;
    Scratch sc(st);
    CsrStream cs(sc, m, K, p, j, x, /*narrow=*/false);
    const bool stage = options().host_stage != 0;
    const bool stage_y = stage && K > 0 && !host_is_pinned(y);
    bool stage_out = stage && !host_is_pinned(out);
    if (stage_out) host_prepare_result(out, (size_t)m * os);
    char *d_y = nullptr, *d_out = nullptr;
    MXG_TRY(sc.alloc((void **)&d_y, (size_t)K * ys));
    MXG_TRY(sc.alloc((void **)&d_out, (size_t)m * os));
    MXG_TRY(chain(sc, st->stream, st->h2d));
    if (!stage_y) MXG_TRY(upload_lines(nullptr, d_y, (size_t)K * ys, y, (size_t)K * ys, (size_t)K * ys, 1, st->h2d));
    MXG_TRY(cs.plan_host(stage_y, stage_out ? os : 0, os));
    if (cs.staging_unavailable) stage_out = false;
    // the vector as lines of <= 4 MiB so that it fits the ring slots whatever K is
    if (stage_y) {
        const size_t line = (size_t)4 << 20, total = (size_t)K * ys;
        MXG_TRY(upload_lines(&cs.ring, d_y, line, y, line, line, total / line, st->h2d));
        const size_t done = total / line * line;
        MXG_TRY(upload_lines(&cs.ring, d_y + done, total - done, static_cast<const char *>(y) + done, total - done, total - done, 1, st->h2d));
    }
    MXG_TRY(chain(sc, st->h2d, st->stream));

    MXG_TRY(cs.begin(16)); // a double and a flag per piece
    const int C = cs.chunks();
    std::vector<cudaEvent_t> ev_done((size_t)C), ev_out((size_t)(stage_out ? C : 0));
    for (int c = 0; c < C; c++) MXG_TRY(sc.event(&ev_done[(size_t)c]));
    for (size_t c = 0; c < ev_out.size(); c++) MXG_TRY(sc.event(&ev_out[c]));
    auto process = [&](int c) -> int {
        mxg_csr_s h;
        MXG_TRY(cs.prepare_chunk(c, h));
        const size_t r0 = (size_t)cs.plan.chunk_row[(size_t)c], nr = (size_t)h.m;
        MXG_TRY(launch_spmv(&h, ytype, d_y, d_out + r0 * os, st->stream));
        MXG_CUDA_TRY(cudaEventRecord(ev_done[(size_t)c], st->stream));
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->d2h, ev_done[(size_t)c], 0));
        if (nr == 0) return MXG_OK;
        char *dst = stage_out ? cs.out_slot(c) : static_cast<char *>(out) + r0 * os;
        MXG_CUDA_TRY(copy_async(dst, d_out + r0 * os, nr * os, cudaMemcpyDeviceToHost, st->d2h));
        if (stage_out) MXG_CUDA_TRY(cudaEventRecord(ev_out[(size_t)c], st->d2h));
        return MXG_OK;
    };
    auto drain = [&](int c) -> int {
        if (!stage_out) return MXG_OK;
        const size_t r0 = (size_t)cs.plan.chunk_row[(size_t)c], nr = (size_t)cs.plan.chunk_row[(size_t)c + 1] - r0;
        if (nr == 0) return MXG_OK;
        MXG_CUDA_TRY(cudaEventSynchronize(ev_out[(size_t)c]));
        host_copy(static_cast<char *>(out) + r0 * os, cs.out_slot(c), nr * os);
        return MXG_OK;
    };
    for (int c = 0; c < C + 2; c++) {
        if (c < C) MXG_TRY(cs.upload_chunk(c));
        if (c >= 1 && c <= C) MXG_TRY(process(c - 1));
        if (c >= 2) MXG_TRY(drain(c - 2));
    }
    int rc = cs.finish();
    if (rc == MXG_OK && keep && p[0] == 0) rc = cs.detach_handle(keep);
    return rc;
}

// ================================================================================================
// Warm path: the CSR is already device-resident (an explicit handle from mxg_csr_upload, or an
// entry of the level-1 operand cache), so a product only moves the dense operand up and the result down.
// The result leaves in row chunks: chunk c is downloaded (and, for a pageable result, copied out of its
// page-locked slot by the host threads) while chunk c + 1 is computed.  Long rows (pieces + fix-up) run first.
// ================================================================================================
// row chunks of a handle (about 16 of equal nnz, tapered like the streamed plan), computed once per handle
int handle_chunks(mxg_csr_s *A, cudaStream_t stream)
{
    if (A->host_chunks) return MXG_OK;
    std::vector<int32_t> hp((size_t)A->m + 1);
    MXG_CUDA_TRY(cudaMemcpyAsync(hp.data(), A->d_p, sizeof(int32_t) * hp.size(), cudaMemcpyDeviceToHost, stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(stream));
    HostPlan plan;
    MXG_TRY(build_plan(A->m, hp.data(), plan, 0));
    A->host_chunks = new std::vector<int32_t>(plan.chunk_row.begin(), plan.chunk_row.end());
    return MXG_OK;
}

// Sub-teams per warp the SpMM kernel runs an n-column product of packed, 16-byte aligned operands with (the team
// geometry of spmm.cu: spmm_typed / dispatch_geom).  A row's entries are dealt to the sub-teams round-robin and the
// sub-team sums are combined by a tree, so two products whose sub-team counts agree sum every output element in the
// same order.  0 = not predictable from here (scalar path, forced geometry).
static int spmm_subteams(size_t n, size_t s)
{
    const size_t V = 16 / s;
    if (n == 0 || n % V != 0 || options().spmm_lpr > 0 || options().spmm_cpl != 0) return 0;
    const size_t nvec = n / V;
    size_t lpr = 4;
    while (lpr < nvec && lpr < 32) lpr *= 2;
    if (lpr >= 8 && nvec > 8 && !(lpr == 32 && nvec > 32)) lpr /= 2; // two vectors per lane on a half-width team
    return (int)(32 / lpr);
}

// The warm product in two column halves (host_colsplit).  PCIe is full duplex but the plain warm call uses one direction
// at a time: the result needs ALL of the dense operand.  Cut by columns, half 0 of the result can leave while half 1 of
// the operand arrives: 2-D copies of >= 128-byte lines at the caller's pitch run at link speed in both directions
// (tools/dma2d_probe.cu: 51.4 / 52.2 GB/s for 128-byte lines at a 256-byte pitch, against 55.6 / 56.5 GB/s for whole
// rows: the copies alone take 13.1 ms in this schedule and 13.6 ms in one piece).  Timeline for cfg3 fp32 n = 64:
//   h2d    | B[:, :32] 2.5 ms | B[:, 32:] 2.5 ms |
//   kernel |                  | half 0, chunk by chunk | half 1 ...
//   d2h    |                    | Out[:, :32] 4.9 ms ............| Out[:, 32:] 4.9 ms |
// Both halves are packed on the device ([K][n/2] and [m][n/2]): the kernels see ordinary n/2-column products.
static int handle_spmm_host_split(DeviceState *st, mxg_csr_s *A, int dtype, int n, const void *B, size_t ldb, void *Out, size_t ldc,
                                  const std::vector<int32_t> &chunk_row, size_t max_rows)
{
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    const size_t rows = (size_t)A->m, Kz = (size_t)A->K, nh = (size_t)n / 2;
    const int C = (int)chunk_row.size() - 1, U = 2 * C;
    Scratch sc(st);
    const bool stage = options().host_stage != 0;
    const bool stage_B = stage && !host_is_pinned(B);
    bool stage_out = stage && !host_is_pinned(Out);
    if (stage_out) host_prepare_result(Out, ((rows - 1) * ldc + (size_t)n) * s);
    char *d_Bh[2] = {nullptr, nullptr}, *d_Oh[2] = {nullptr, nullptr};
    for (int h = 0; h < 2; h++) {
        MXG_TRY(sc.alloc((void **)&d_Bh[h], Kz * nh * s));
        MXG_TRY(sc.alloc((void **)&d_Oh[h], rows * nh * s));
    }
    auto up = [](size_t v) { return (v + 4095) & ~(size_t)4095; };
    const int S = (int)std::min<long>(std::max<long>(options().pipe_slots, 3), 8);
    const size_t in_slot = stage_B ? (size_t)16 << 20 : 0;
    const size_t out_slot_bytes = stage_out ? up(max_rows * nh * s) : 0;
    InRing ring;
    char *out_base = nullptr;
    if (in_slot + out_slot_bytes > 0) {
        char *base = nullptr;
        if (pinned_arena(st, (size_t)S * (in_slot + out_slot_bytes), &base) != MXG_OK) {
            cudaGetLastError();
            last_error_ref().clear();
            stage_out = false; // the driver's own copies still work
        } else {
            if (in_slot) MXG_TRY(ring.init(sc, base, in_slot, S));
            out_base = base + (size_t)S * in_slot;
        }
    }
    std::vector<cudaEvent_t> ev_done((size_t)U), ev_out((size_t)(stage_out ? U : 0));
    for (int u = 0; u < U; u++) MXG_TRY(sc.event(&ev_done[(size_t)u]));
    for (size_t u = 0; u < ev_out.size(); u++) MXG_TRY(sc.event(&ev_out[u]));
    auto out_slot = [&](int u) { return out_base + (size_t)(u % S) * out_slot_bytes; };
    MXG_TRY(chain(sc, st->stream, st->h2d)); // the buffers exist in stream order of st->stream

    // unit u = (half u / C, chunk u % C): the order in which result blocks leave the device
    auto upload = [&](int h) -> int {
        const char *src = static_cast<const char *>(B) + (size_t)h * nh * s;
        MXG_TRY(upload_lines(ring.enabled() ? &ring : nullptr, d_Bh[h], nh * s, src, ldb * s, nh * s, Kz, st->h2d));
        return chain(sc, st->h2d, st->stream);
    };
    auto compute = [&](int h) -> int { // long rows first, then the row chunks in download order
        if (A->n_pieces > 0)
            MXG_TRY(launch_spmm_rows(A, dtype, MXG_ROWS_CONTIGUOUS, (int)nh, d_Bh[h], nh, d_Oh[h], nh, 0, 0, /*pieces=*/1, st->stream));
        for (int c = 0; c < C; c++) {
            MXG_TRY(launch_spmm_rows(A, dtype, MXG_ROWS_CONTIGUOUS, (int)nh, d_Bh[h], nh, d_Oh[h], nh, chunk_row[(size_t)c],
                                     chunk_row[(size_t)c + 1], /*pieces=*/0, st->stream));
            MXG_CUDA_TRY(cudaEventRecord(ev_done[(size_t)(h * C + c)], st->stream));
        }
        return MXG_OK;
    };
    auto host_block = [&](int u) { // where unit u lives in the caller's result
        const size_t r0 = (size_t)chunk_row[(size_t)(u % C)];
        return static_cast<char *>(Out) + (r0 * ldc + (size_t)(u / C) * nh) * s;
    };
    auto download = [&](int u) -> int {
        const int h = u / C, c = u % C;
        const size_t r0 = (size_t)chunk_row[(size_t)c], nr = (size_t)chunk_row[(size_t)c + 1] - r0;
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->d2h, ev_done[(size_t)u], 0));
        const char *d_o = d_Oh[h] + r0 * nh * s;
        if (stage_out) {
            MXG_TRY(copy_rows(out_slot(u), nh * s, d_o, nh * s, nh * s, nr, cudaMemcpyDeviceToHost, st->d2h));
            MXG_CUDA_TRY(cudaEventRecord(ev_out[(size_t)u], st->d2h));
        } else {
            MXG_TRY(copy_rows(host_block(u), ldc * s, d_o, nh * s, nh * s, nr, cudaMemcpyDeviceToHost, st->d2h));
        }
        return MXG_OK;
    };
    auto drain = [&](int u) -> int {
        if (!stage_out) return MXG_OK;
        const size_t nr = (size_t)chunk_row[(size_t)(u % C) + 1] - (size_t)chunk_row[(size_t)(u % C)];
        MXG_CUDA_TRY(cudaEventSynchronize(ev_out[(size_t)u]));
        host_copy_2d(host_block(u), ldc * s, out_slot(u), nh * s, nh * s, nr);
        return MXG_OK;
    };
    // A pageable operand is packed by THIS thread (upload blocks while the host threads fill the ring), so everything of
    // half 0 that needs no draining is put on the streams before half 1 is touched: the device then computes and
    // downloads half 0 while the host packs half 1.
    MXG_TRY(upload(0));
    MXG_TRY(compute(0));
    const int pre = stage_out ? std::min(S - 1, C) : C; // downloads that fit the free output slots
    for (int u = 0; u < pre; u++) MXG_TRY(download(u));
    MXG_TRY(upload(1));
    MXG_TRY(compute(1));
    for (int u = pre; u < U + S - 1; u++) {
        if (stage_out && u >= S - 1 && u - (S - 1) < U) MXG_TRY(drain(u - (S - 1)));
        if (u < U) MXG_TRY(download(u));
    }
    MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(st->d2h));
    return MXG_OK;
}

int handle_spmm_host(DeviceState *st, mxg_csr_s *A, int dtype, int out_layout, int b_layout, int n, const void *B, size_t ldb,
                     void *Out, size_t ldc, const void *d_B_resident, void **d_B_keep)
{
    const size_t s = dtype == MXG_F64 ? 8 : 4;
    const size_t vec = 16 / s;
    const int m = A->m, K = A->K;
    const size_t rows = (size_t)m, Kz = (size_t)K, nz = (size_t)n;
    if (d_B_keep) *d_B_keep = nullptr;
    if (m == 0 || n == 0) return MXG_OK;
    if (!Out) return fail(MXG_ERR_ARG, "output is NULL");
    if (!B && K > 0 && !d_B_resident) return fail(MXG_ERR_ARG, "dense operand is NULL");
    if (b_layout != MXG_ROWS_CONTIGUOUS && b_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "dense operand: bad layout %d", b_layout);
    if (out_layout != MXG_ROWS_CONTIGUOUS && out_layout != MXG_COLS_CONTIGUOUS) return fail(MXG_ERR_ARG, "bad out_layout %d", out_layout);
    if (b_layout == MXG_ROWS_CONTIGUOUS && ldb < nz) return fail(MXG_ERR_ARG, "dense operand: ldb < n");
    if (b_layout == MXG_COLS_CONTIGUOUS && ldb < Kz) return fail(MXG_ERR_ARG, "dense operand: ldb < K");
    if (out_layout == MXG_ROWS_CONTIGUOUS && ldc < nz) return fail(MXG_ERR_ARG, "output: ldc < n");
    if (out_layout == MXG_COLS_CONTIGUOUS && ldc < rows) return fail(MXG_ERR_ARG, "output: ldc < m");
    if (dtype == MXG_F64 && !A->d_x64 && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "handle holds no float64 values");
    if (dtype == MXG_F32 && !A->d_x32 && A->nnz > 0) return fail(MXG_ERR_UNSUPPORTED, "handle holds no float32 values");

    g_h2d_bytes = g_d2h_bytes = 0;
    MXG_TRY(handle_chunks(A, st->stream));
    const std::vector<int32_t> &cr = *A->host_chunks;
    // chunks of at most 64 MiB of result rows (the output slots of the page-locked arena are sized by the largest)
    std::vector<int32_t> chunk_row;
    const size_t cap_rows = std::max<size_t>(1 << 12, ((size_t)64 << 20) / (nz * s));
    chunk_row.push_back(0);
    for (size_t c = 0; c + 1 < cr.size(); c++) {
        const size_t r0 = (size_t)cr[c], r1 = (size_t)cr[c + 1];
        const size_t parts = (r1 - r0 + cap_rows - 1) / cap_rows;
        for (size_t q = 1; q <= parts; q++) chunk_row.push_back((int32_t)(r0 + (r1 - r0) * q / parts));
    }
    const int C = (int)chunk_row.size() - 1;
    size_t max_rows = 0;
    for (int c = 0; c < C; c++) max_rows = std::max(max_rows, (size_t)(chunk_row[(size_t)c + 1] - chunk_row[(size_t)c]));

    // Two column halves where that lets the directions of the link overlap.  By default (1) only where it changes no bit of
    // the result and the dense operand is page-locked: a pageable operand is packed into the ring by the host threads,
    // and packing half rows reads every row of it twice (measured, cfg3: 13.6 instead of 14.7 ms page-locked, 15.7
    // instead of 15.1 ms pageable).
    const long colsplit = options().host_colsplit;
    if (colsplit != 0 && out_layout == MXG_ROWS_CONTIGUOUS && b_layout == MXG_ROWS_CONTIGUOUS && !d_B_resident && !d_B_keep && K > 0 &&
        A->nnz > 0 && nz % (2 * vec) == 0 && nz / 2 * s >= 128 && rows * nz * s >= ((size_t)32 << 20) &&
        (colsplit == 2 ||
         (host_is_pinned(B) && spmm_subteams(nz, s) != 0 && spmm_subteams(nz, s) == spmm_subteams(nz / 2, s))))
        return handle_spmm_host_split(st, A, dtype, n, B, ldb, Out, ldc, chunk_row, max_rows);

    Scratch sc(st);
    const bool stage = options().host_stage != 0;
    const bool have_B = d_B_resident != nullptr;
    const bool stage_B = stage && K > 0 && !have_B && !host_is_pinned(B);
    bool stage_out = stage && !host_is_pinned(Out);
    const size_t ld_b = round_up(nz, vec);
    const size_t ld_o = out_layout == MXG_ROWS_CONTIGUOUS ? round_up(nz, vec) : rows;
    const bool rm = out_layout == MXG_ROWS_CONTIGUOUS;
    if (stage_out) host_prepare_result(Out, (rm ? (rows - 1) * ldc + nz : (nz - 1) * ldc + rows) * s);
    char *d_B = const_cast<char *>(static_cast<const char *>(d_B_resident)), *d_Out = nullptr, *d_tmp = nullptr;
    if (!have_B) {
        if (d_B_keep) { // the caller keeps the device copy (operand cache): not a temporary of this call
            MXG_CUDA_TRY(cudaMallocAsync((void **)&d_B, std::max<size_t>(Kz * ld_b * s, 16), st->stream));
            *d_B_keep = d_B;
        } else {
            MXG_TRY(sc.alloc((void **)&d_B, Kz * ld_b * s));
        }
    }
    MXG_TRY(sc.alloc((void **)&d_Out, rm ? rows * ld_o * s : rows * nz * s));
    // page-locked arena: input slots for a pageable dense operand, output slots for a pageable result
    auto up = [](size_t v) { return (v + 4095) & ~(size_t)4095; };
    const int S = (int)std::min<long>(std::max<long>(options().pipe_slots, 3), 8);
    const size_t in_slot = stage_B ? (size_t)16 << 20 : 0;
    size_t out_slot_bytes = stage_out ? up(max_rows * nz * s) : 0;
    InRing ring;
    char *out_base = nullptr;
    if (in_slot + out_slot_bytes > 0) {
        char *base = nullptr;
        if (pinned_arena(st, (size_t)S * (in_slot + out_slot_bytes), &base) != MXG_OK) {
            cudaGetLastError();
            last_error_ref().clear();
            stage_out = false; // the driver's own copies still work
        } else {
            if (in_slot) MXG_TRY(ring.init(sc, base, in_slot, S));
            out_base = base + (size_t)S * in_slot;
        }
    }
    if (!have_B && K > 0) {
        if (b_layout == MXG_COLS_CONTIGUOUS) MXG_TRY(sc.alloc((void **)&d_tmp, Kz * nz * s));
        if (ld_b != nz) MXG_CUDA_TRY(cudaMemsetAsync(d_B, 0, Kz * ld_b * s, st->stream));
        MXG_TRY(chain(sc, st->stream, st->h2d));
        InRing *rg = ring.enabled() ? &ring : nullptr;
        if (b_layout == MXG_ROWS_CONTIGUOUS) MXG_TRY(upload_lines(rg, d_B, ld_b * s, B, ldb * s, nz * s, Kz, st->h2d));
        else MXG_TRY(upload_lines(rg, d_tmp, Kz * s, B, ldb * s, Kz * s, nz, st->h2d));
        MXG_TRY(chain(sc, st->h2d, st->stream));
        if (b_layout == MXG_COLS_CONTIGUOUS) MXG_TRY(launch_transpose_dense((int)s, nz, Kz, d_tmp, Kz, d_B, ld_b, st->stream));
    }
    std::vector<cudaEvent_t> ev_done((size_t)C), ev_out((size_t)(stage_out ? C : 0));
    for (int c = 0; c < C; c++) MXG_TRY(sc.event(&ev_done[(size_t)c]));
    for (size_t c = 0; c < ev_out.size(); c++) MXG_TRY(sc.event(&ev_out[c]));
    auto out_slot = [&](int c) { return out_base + (size_t)(c % S) * out_slot_bytes; };

    // a matrix without stored entries: one zero fill of the whole result (src/matmul.cpp:128-129), chunks only copy
    if (A->nnz == 0) MXG_TRY(launch_spmm(A, dtype, out_layout, n, d_B, ld_b, d_Out, ld_o, st->stream));
    else if (A->n_pieces > 0) MXG_TRY(launch_spmm_rows(A, dtype, out_layout, n, d_B, ld_b, d_Out, ld_o, 0, 0, /*pieces=*/1, st->stream));
    auto process = [&](int c) -> int {
        const size_t r0 = (size_t)chunk_row[(size_t)c], nr = (size_t)chunk_row[(size_t)c + 1] - r0;
        if (A->nnz > 0)
            MXG_TRY(launch_spmm_rows(A, dtype, out_layout, n, d_B, ld_b, d_Out, ld_o, (int)r0, (int)(r0 + nr), /*pieces=*/0, st->stream));
        MXG_CUDA_TRY(cudaEventRecord(ev_done[(size_t)c], st->stream));
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->d2h, ev_done[(size_t)c], 0));
        char *d_o = d_Out + (rm ? r0 * ld_o * s : r0 * s);
        const size_t width = rm ? nz * s : nr * s, height = rm ? nr : nz;
        if (stage_out) {
            MXG_TRY(copy_rows(out_slot(c), width, d_o, ld_o * s, width, height, cudaMemcpyDeviceToHost, st->d2h));
            MXG_CUDA_TRY(cudaEventRecord(ev_out[(size_t)c], st->d2h));
        } else {
            char *h_o = static_cast<char *>(Out) + (rm ? r0 * ldc * s : r0 * s);
            MXG_TRY(copy_rows(h_o, ldc * s, d_o, ld_o * s, width, height, cudaMemcpyDeviceToHost, st->d2h));
        }
        return MXG_OK;
    };
    auto drain = [&](int c) -> int {
        if (!stage_out) return MXG_OK;
        const size_t r0 = (size_t)chunk_row[(size_t)c], nr = (size_t)chunk_row[(size_t)c + 1] - r0;
        const size_t width = rm ? nz * s : nr * s, height = rm ? nr : nz;
        MXG_CUDA_TRY(cudaEventSynchronize(ev_out[(size_t)c]));
        char *h_o = static_cast<char *>(Out) + (rm ? r0 * ldc * s : r0 * s);
        host_copy_2d(h_o, ldc * s, out_slot(c), width, width, height);
        return MXG_OK;
    };
    // output slot c % S is free again once chunk c - S has been drained: keep S - 1 downloads in flight
    for (int c = 0; c < C + S - 1; c++) {
        if (c >= S - 1) MXG_TRY(drain(c - (S - 1)));
        if (c < C) MXG_TRY(process(c));
    }
    MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
    MXG_CUDA_TRY(cudaStreamSynchronize(st->d2h));
    return MXG_OK;
}

int handle_spmv_host(DeviceState *st, mxg_csr_s *A, int ytype, const void *y, void *out)
{
    const int m = A->m, K = A->K;
    if (m == 0) return MXG_OK;
    if (!out) return fail(MXG_ERR_ARG, "output is NULL");
    if (K > 0 && !y) return fail(MXG_ERR_ARG, "vector is NULL");
    const size_t ys = ytype == MXG_Y_NUMERIC ? 8 : 4;
    const size_t os = ytype == MXG_Y_FLOAT32 ? 4 : 8;
    g_h2d_bytes = g_d2h_bytes = 0;
    Scratch sc(st);
    char *d_y = nullptr, *d_out = nullptr;
    MXG_TRY(sc.alloc((void **)&d_y, (size_t)K * ys));
    MXG_TRY(sc.alloc((void **)&d_out, (size_t)m * os));
    cudaStream_t q = st->stream;
    g_h2d_bytes += (size_t)K * ys;
    g_d2h_bytes += (size_t)m * os;
    MXG_TRY(staged_h2d(st, d_y, y, (size_t)K * ys, q));
    MXG_TRY(launch_spmv(A, ytype, d_y, d_out, q));
    MXG_TRY(staged_d2h(st, out, d_out, (size_t)m * os, q));
    MXG_CUDA_TRY(cudaStreamSynchronize(q));
    return MXG_OK;
}

} // namespace mxg
