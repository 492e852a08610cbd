// tools/dma2d_probe.cu — how fast are PCIe copies of NARROW lines?  (development probe: would a column-split of the dense
// operand / the result — 128-byte segments at a 256-byte pitch — let the warm path's upload and download overlap?)
//   nvcc -O2 -o /tmp/dma2d tools/dma2d_probe.cu && /tmp/dma2d
// Part 1: one 2-D copy of 2 M lines of 256 / 128 / 64 bytes at a 256-byte host pitch, each direction alone.
// Part 2: the schedule of handle_spmm_host_split without its kernels (cfg3 fp32 n = 64: K = 1 M rows of B up, m = 2 M rows
//         of the result down, two halves): B half 0 up | B half 1 up while result half 0 comes down in C chunks | half 1
//         down — against the unsplit schedule (all of B up, then the result down as C contiguous chunks).
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t rows = 2000000, full = 256, K = 1000000;
    char *h = nullptr, *d = nullptr, *hb = nullptr, *db = nullptr;
    cudaHostAlloc(&h, rows * full, cudaHostAllocDefault);
    cudaHostAlloc(&hb, K * full, cudaHostAllocDefault);
    cudaMalloc(&d, rows * full);
    cudaMalloc(&db, K * full);
    cudaStream_t s, s2;
    cudaStreamCreate(&s);
    cudaStreamCreate(&s2);
    cudaEvent_t a, b, e0;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventCreate(&e0);
    for (size_t width : {(size_t)256, (size_t)128, (size_t)64}) {
        for (int dir = 0; dir < 2; dir++) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(a, s);
                if (dir == 0) cudaMemcpy2DAsync(d, width, h, full, width, rows, cudaMemcpyHostToDevice, s);   // strided host -> packed device
                else cudaMemcpy2DAsync(h, full, d, width, width, rows, cudaMemcpyDeviceToHost, s);             // packed device -> strided host
                cudaEventRecord(b, s);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            printf("{\"width_bytes\": %zu, \"host_pitch\": %zu, \"dir\": \"%s\", \"rows\": %zu, \"ms\": %.3f, \"GBps\": %.1f}\n", width, full,
                   dir == 0 ? "h2d" : "d2h", rows, best, rows * width / best / 1e6);
        }
    }
    const size_t half = full / 2;
    for (int C : {1, 16, 64}) {
        for (int mode = 0; mode < 3; mode++) { // 0: unsplit, 1: split (both directions overlap), 2: split, downloads only after both uploads
            float best = 1e9f;
            for (int rep = 0; rep < 4; rep++) {
                cudaDeviceSynchronize();
                cudaEventRecord(a, s);
                if (mode == 0) {
                    cudaMemcpyAsync(db, hb, K * full, cudaMemcpyHostToDevice, s);
                    for (int c = 0; c < C; c++) {
                        const size_t r0 = rows * c / C, r1 = rows * (c + 1) / C;
                        cudaMemcpyAsync(h + r0 * full, d + r0 * full, (r1 - r0) * full, cudaMemcpyDeviceToHost, s);
                    }
                } else {
                    cudaMemcpy2DAsync(db, half, hb, full, half, K, cudaMemcpyHostToDevice, s);
                    cudaEventRecord(e0, s);
                    cudaMemcpy2DAsync(db + K * half, half, hb + half, full, half, K, cudaMemcpyHostToDevice, s);
                    if (mode == 2) cudaEventRecord(e0, s);
                    cudaStreamWaitEvent(s2, e0, 0);
                    for (int hh = 0; hh < 2; hh++)
                        for (int c = 0; c < C; c++) {
                            const size_t r0 = rows * c / C, r1 = rows * (c + 1) / C;
                            cudaMemcpy2DAsync(h + r0 * full + hh * half, full, d + hh * rows * half + r0 * half, half, half, r1 - r0,
                                              cudaMemcpyDeviceToHost, s2);
                        }
                    cudaEventRecord(b, s2);
                    cudaStreamWaitEvent(s, b, 0);
                }
                cudaEventRecord(b, s);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            printf("{\"schedule\": \"%s\", \"chunks_per_half\": %d, \"ms\": %.3f}\n",
                   mode == 0 ? "unsplit: B up, result down" : mode == 1 ? "split: half 1 up while half 0 down" : "split copies, no overlap", C, best);
        }
    }
    return 0;
}
