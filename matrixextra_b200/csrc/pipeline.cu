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
#include "mxg_internal.cuh"

#include <algorithm>
#include <climits>
#include <vector>

namespace mxg {
namespace {

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
};

// One pass over the host indptr: chunk boundaries, monotonicity check, long-row tables (the host twin of
// k_row_stats / k_fill_long_tables in layout.cu; same table format, rows numbered from the chunk start).
int build_plan(int m, const int32_t *p, HostPlan &plan)
{
    plan.piece = (int)std::max<long>(32, options().piece);
    const int64_t nnz = p[m];
    int64_t target_nnz = std::max<int64_t>(nnz / 16, (int64_t)1 << 20);
    int target_rows = std::max(m / 16, 1 << 16);
    if (options().pipe_chunk_nnz > 0) { // tests: force many small chunks
        target_nnz = options().pipe_chunk_nnz;
        target_rows = (int)std::min<int64_t>(target_nnz, INT32_MAX);
    }
    plan.chunk_row.push_back(0);
    plan.chunk_long_off.push_back(0);
    plan.chunk_piece_off.push_back(0);
    int chunk_start = 0;
    int64_t chunk_first = 0;
    int pieces_in_chunk = 0;
    auto close_chunk = [&](int row_end) {
        plan.chunk_row.push_back(row_end);
        plan.chunk_long_off.push_back((int)plan.long_rows.size());
        plan.chunk_piece_off.push_back((int)plan.piece_row.size());
        plan.max_chunk_pieces = std::max(plan.max_chunk_pieces, pieces_in_chunk);
        plan.max_chunk_nnz = std::max(plan.max_chunk_nnz, (size_t)((int64_t)p[row_end] - chunk_first));
        chunk_start = row_end;
        chunk_first = p[row_end];
        pieces_in_chunk = 0;
    };
    for (int r = 0; r < m; r++) {
        const int a = p[r], b = p[r + 1];
        if (a < 0 || b < a) return fail(MXG_ERR_INDEX, "CSR indptr is negative or decreasing");
        const int len = b - a;
        plan.max_len = std::max(plan.max_len, len);
        if (len > plan.piece) {
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
        if ((int64_t)b - chunk_first >= target_nnz || r + 1 - chunk_start >= target_rows) close_chunk(r + 1);
    }
    if (chunk_start < m) close_chunk(m);
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

int copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
              cudaStream_t stream)
{
    if (width == 0 || height == 0) return MXG_OK;
    if (dpitch == width && spitch == width) MXG_CUDA_TRY(cudaMemcpyAsync(dst, src, width * height, kind, stream));
    else MXG_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, stream));
    return MXG_OK;
}

// state shared by the SpMM and SpMV pipelines: device CSR arrays, chunk plan, long-row tables
struct CsrStream {
    Scratch &sc;
    HostPlan plan;
    int m, K;
    int64_t nnz;
    const int32_t *p, *j;
    const double *x;
    bool narrow; // values are narrowed to float32 on the device
    int32_t *d_p = nullptr, *d_j = nullptr;
    double *d_x64 = nullptr;
    float *d_x32 = nullptr;
    static constexpr int NSTAGE = 3;
    double *d_stage[NSTAGE] = {nullptr, nullptr, nullptr};
    int32_t *d_tables = nullptr;
    int *d_flag = nullptr;
    void *d_partial = nullptr;
    size_t partial_bytes = 0;
    std::vector<cudaEvent_t> ev_h2d, ev_conv;

    CsrStream(Scratch &s, int m_, int K_, const int32_t *p_, const int32_t *j_, const double *x_, bool narrow_)
        : sc(s), m(m_), K(K_), nnz(p_[m_]), p(p_), j(j_), x(x_), narrow(narrow_)
    {
    }
    int chunks() const { return (int)plan.chunk_row.size() - 1; }

    // allocations + indptr upload + host plan + table upload.  partial_per_piece = workspace bytes per piece.
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
        MXG_CUDA_TRY(cudaMemcpyAsync(d_p, p, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice, st->h2d));
        MXG_TRY(build_plan(m, p, plan)); // host work, overlaps the copies already in flight
        const size_t nl = plan.long_rows.size(), np = plan.piece_row.size();
        if (nl > 0) {
            MXG_TRY(sc.alloc((void **)&d_tables, sizeof(int32_t) * (3 * nl + 2 * np)));
            MXG_TRY(sc.alloc(&d_partial, (size_t)plan.max_chunk_pieces * partial_per_piece));
            partial_bytes = (size_t)plan.max_chunk_pieces * partial_per_piece;
            MXG_TRY(chain(sc, st->stream, st->h2d));
            const std::vector<int32_t> *src[5] = {&plan.long_rows, &plan.long_first, &plan.long_np, &plan.piece_row, &plan.piece_k};
            size_t off = 0;
            for (int t = 0; t < 5; t++) {
                MXG_CUDA_TRY(cudaMemcpyAsync(d_tables + off, src[t]->data(), sizeof(int32_t) * src[t]->size(),
                                             cudaMemcpyHostToDevice, st->h2d));
                off += src[t]->size();
            }
        }
        if (narrow) {
            const size_t stage_n = plan.max_chunk_nnz + (plan.max_chunk_nnz & 1);
            for (int b = 0; b < NSTAGE && b < chunks(); b++) MXG_TRY(sc.alloc((void **)&d_stage[b], sizeof(double) * stage_n));
            MXG_TRY(chain(sc, st->stream, st->h2d));
        }
        ev_h2d.resize((size_t)chunks());
        ev_conv.resize((size_t)chunks());
        for (int c = 0; c < chunks(); c++) {
            MXG_TRY(sc.event(&ev_h2d[(size_t)c]));
            if (narrow) MXG_TRY(sc.event(&ev_conv[(size_t)c]));
        }
        return MXG_OK;
    }

    // h2d stream: indices and values of chunk c
    int upload_chunk(int c)
    {
        DeviceState *st = sc.st;
        const int64_t e0 = p[plan.chunk_row[(size_t)c]], e1 = p[plan.chunk_row[(size_t)c + 1]];
        const size_t len = (size_t)(e1 - e0);
        if (len > 0) {
            MXG_CUDA_TRY(cudaMemcpyAsync(d_j + e0, j + e0, sizeof(int32_t) * len, cudaMemcpyHostToDevice, st->h2d));
            if (narrow) {
                // the staging buffer is free again once the narrowing of chunk c - NSTAGE has run
                if (c >= NSTAGE) MXG_CUDA_TRY(cudaStreamWaitEvent(st->h2d, ev_conv[(size_t)(c - NSTAGE)], 0));
                MXG_CUDA_TRY(cudaMemcpyAsync(d_stage[c % NSTAGE], x + e0, sizeof(double) * len, cudaMemcpyHostToDevice, st->h2d));
            } else {
                MXG_CUDA_TRY(cudaMemcpyAsync(d_x64 + e0, x + e0, sizeof(double) * len, cudaMemcpyHostToDevice, st->h2d));
            }
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
        if (narrow) {
            if (len > 0) {
                // chunk starts are not always even: narrow element-wise from the staging buffer's start
                MXG_TRY(convert_f64_to_f32(d_stage[c % NSTAGE], d_x32 + e0, len, st->stream));
            }
            MXG_CUDA_TRY(cudaEventRecord(ev_conv[(size_t)c], st->stream));
        }
        MXG_TRY(check_indices_flag(len, d_j + e0, K, d_flag, st->stream));
        const size_t nl = plan.long_rows.size(), np = plan.piece_row.size();
        const int l0 = plan.chunk_long_off[(size_t)c], l1 = plan.chunk_long_off[(size_t)c + 1];
        const int q0 = plan.chunk_piece_off[(size_t)c], q1 = plan.chunk_piece_off[(size_t)c + 1];
        h = mxg_csr_s();
        cudaGetDevice(&h.device);
        h.m = r1 - r0;
        h.K = K;
        h.nnz = e1 - e0;
        h.base = (int32_t)e0;
        h.d_p = d_p + r0; // offsets stay absolute: the kernels index d_j / d_x with them directly
        h.d_j = d_j;
        h.d_x64 = d_x64;
        h.d_x32 = d_x32;
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

    // after the last chunk: read the validation flag back and wait for every stream
    int finish()
    {
        DeviceState *st = sc.st;
        int flag = 0;
        MXG_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st->stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(st->stream));
        MXG_CUDA_TRY(cudaStreamSynchronize(st->d2h));
        MXG_CUDA_TRY(cudaStreamSynchronize(st->h2d));
        if (flag) return fail(MXG_ERR_INDEX, "CSR column index outside [0, %d)", K);
        return MXG_OK;
    }
};

} // namespace

int pipeline_spmm(DeviceState *st, int dtype, int out_layout, int b_layout, int m, int K, int n, const int32_t *p,
                  const int32_t *j, const double *x, const void *B, size_t ldb, void *Out, size_t ldc)
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
    const int64_t nnz = p[m];
    if (nnz > 0 && (!j || !x)) return fail(MXG_ERR_ARG, "csr: indices / values is NULL");

    Scratch sc(st);
    // dense operand first: every chunk needs all of it.  Device copy is rows-contiguous [K][ld_b].
    const size_t ld_b = round_up(nz, vec);
    char *d_B = nullptr, *d_Out = nullptr;
    MXG_TRY(sc.alloc((void **)&d_B, Kz * ld_b * s));
    const size_t ld_o = out_layout == MXG_ROWS_CONTIGUOUS ? round_up(nz, vec) : rows;
    MXG_TRY(sc.alloc((void **)&d_Out, out_layout == MXG_ROWS_CONTIGUOUS ? rows * ld_o * s : rows * nz * s));
    if (K > 0) {
        if (ld_b != nz) MXG_CUDA_TRY(cudaMemsetAsync(d_B, 0, Kz * ld_b * s, st->stream));
        if (b_layout == MXG_ROWS_CONTIGUOUS) {
            MXG_TRY(chain(sc, st->stream, st->h2d));
            MXG_TRY(copy_rows(d_B, ld_b * s, B, ldb * s, nz * s, Kz, cudaMemcpyHostToDevice, st->h2d));
            MXG_TRY(chain(sc, st->h2d, st->stream));
        } else {
            char *d_tmp = nullptr;
            MXG_TRY(sc.alloc((void **)&d_tmp, Kz * nz * s));
            MXG_TRY(chain(sc, st->stream, st->h2d));
            MXG_TRY(copy_rows(d_tmp, Kz * s, B, ldb * s, Kz * s, nz, cudaMemcpyHostToDevice, st->h2d));
            MXG_TRY(chain(sc, st->h2d, st->stream));
            MXG_TRY(launch_transpose_dense((int)s, nz, Kz, d_tmp, Kz, d_B, ld_b, st->stream)); // [n][K] -> [K][ld_b]
        }
    }

    CsrStream cs(sc, m, K, p, j, x, /*narrow=*/dtype == MXG_F32);
    MXG_TRY(cs.begin(nz * s));
    const int C = cs.chunks();
    std::vector<cudaEvent_t> ev_done((size_t)C);
    for (int c = 0; c < C; c++) MXG_TRY(sc.event(&ev_done[(size_t)c]));

    auto process = [&](int c) -> int {
        mxg_csr_s h;
        MXG_TRY(cs.prepare_chunk(c, h));
        const size_t r0 = (size_t)cs.plan.chunk_row[(size_t)c], nr = (size_t)h.m;
        char *d_o = d_Out + (out_layout == MXG_ROWS_CONTIGUOUS ? r0 * ld_o * s : r0 * s);
        MXG_TRY(launch_spmm(&h, dtype, out_layout, n, d_B, ld_b, d_o, ld_o, st->stream));
        if (h.d_seg) MXG_CUDA_TRY(cudaFreeAsync(h.d_seg, st->stream)); // the chunk's column-panel table
        MXG_CUDA_TRY(cudaEventRecord(ev_done[(size_t)c], st->stream));
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->d2h, ev_done[(size_t)c], 0));
        char *h_o = static_cast<char *>(Out) + (out_layout == MXG_ROWS_CONTIGUOUS ? r0 * ldc * s : r0 * s);
        if (out_layout == MXG_ROWS_CONTIGUOUS)
            MXG_TRY(copy_rows(h_o, ldc * s, d_o, ld_o * s, nz * s, nr, cudaMemcpyDeviceToHost, st->d2h));
        else
            MXG_TRY(copy_rows(h_o, ldc * s, d_o, ld_o * s, nr * s, nz, cudaMemcpyDeviceToHost, st->d2h));
        return MXG_OK;
    };
    for (int c = 0; c < C; c++) {
        MXG_TRY(cs.upload_chunk(c));
        if (c >= 1) MXG_TRY(process(c - 1));
    }
    if (C >= 1) MXG_TRY(process(C - 1));
    return cs.finish();
}

int pipeline_spmv(DeviceState *st, int ytype, int m, int K, const int32_t *p, const int32_t *j, const double *x,
                  const void *y, void *out)
{
    if (m == 0) return MXG_OK;
    if (!out) return fail(MXG_ERR_ARG, "output is NULL");
    if (K > 0 && !y) return fail(MXG_ERR_ARG, "vector is NULL");
    const int64_t nnz = p[m];
    if (nnz > 0 && (!j || !x)) return fail(MXG_ERR_ARG, "csr: indices / values is NULL");
    const size_t ys = ytype == MXG_Y_NUMERIC ? 8 : 4;
    const size_t os = ytype == MXG_Y_FLOAT32 ? 4 : 8;

    Scratch sc(st);
    char *d_y = nullptr, *d_out = nullptr;
    MXG_TRY(sc.alloc((void **)&d_y, (size_t)K * ys));
    MXG_TRY(sc.alloc((void **)&d_out, (size_t)m * os));
    MXG_TRY(chain(sc, st->stream, st->h2d));
    if (K > 0) MXG_CUDA_TRY(cudaMemcpyAsync(d_y, y, (size_t)K * ys, cudaMemcpyHostToDevice, st->h2d));
    MXG_TRY(chain(sc, st->h2d, st->stream));

    CsrStream cs(sc, m, K, p, j, x, /*narrow=*/false);
    MXG_TRY(cs.begin(16)); // a double and a flag per piece
    const int C = cs.chunks();
    std::vector<cudaEvent_t> ev_done((size_t)C);
    for (int c = 0; c < C; c++) MXG_TRY(sc.event(&ev_done[(size_t)c]));
    auto process = [&](int c) -> int {
        mxg_csr_s h;
        MXG_TRY(cs.prepare_chunk(c, h));
        const size_t r0 = (size_t)cs.plan.chunk_row[(size_t)c], nr = (size_t)h.m;
        MXG_TRY(launch_spmv(&h, ytype, d_y, d_out + r0 * os, st->stream));
        MXG_CUDA_TRY(cudaEventRecord(ev_done[(size_t)c], st->stream));
        MXG_CUDA_TRY(cudaStreamWaitEvent(st->d2h, ev_done[(size_t)c], 0));
        MXG_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(out) + r0 * os, d_out + r0 * os, nr * os, cudaMemcpyDeviceToHost, st->d2h));
        return MXG_OK;
    };
    for (int c = 0; c < C; c++) {
        MXG_TRY(cs.upload_chunk(c));
        if (c >= 1) MXG_TRY(process(c - 1));
    }
    if (C >= 1) MXG_TRY(process(C - 1));
    return cs.finish();
}

} // namespace mxg
