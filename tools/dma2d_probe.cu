// tools/dma2d_probe.cu — how fast are PCIe copies of NARROW lines?  (development probe: would a column-split of the dense
// operand / the result — 128-byte segments at a 256-byte pitch — let the warm path's upload and download overlap?)
//   nvcc -O2 -o /tmp/dma2d tools/dma2d_probe.cu && /tmp/dma2d
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t rows = 2000000, full = 256;
    char *h = nullptr, *d = nullptr;
    cudaHostAlloc(&h, rows * full, cudaHostAllocDefault);
    cudaMalloc(&d, rows * full);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
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
    return 0;
}
