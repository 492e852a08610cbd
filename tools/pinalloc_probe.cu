// tools/pinalloc_probe.cu — what does a NEW page-locked result block cost, and is there a cheaper way to get one?
//   nvcc -O2 -o /tmp/pinalloc tools/pinalloc_probe.cu -lpthread && /tmp/pinalloc
// (a) cudaHostAlloc(bytes)                                         — what the result pool does today
// (b) mmap + madvise(MADV_HUGEPAGE) + touch by T threads + cudaHostRegister — huge pages: 512 x fewer faults, fewer pins
// (c) mmap (4 KiB pages) + touch by T threads + cudaHostRegister
// and, for each, the speed of a device -> host copy into the block (the block must be as good a DMA target).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <sys/mman.h>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void touch(char *p, size_t bytes, int T)
{
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([=] {
            const size_t a = bytes * t / T, b = bytes * (t + 1) / T;
            for (size_t o = a; o < b; o += 4096) p[o] = 0;
        });
    for (auto &x : th) x.join();
}

int main()
{
    cudaFree(0);
    char *d = nullptr;
    const size_t maxb = (size_t)512 << 20;
    cudaMalloc(&d, maxb);
    cudaMemset(d, 1, maxb);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    auto d2h = [&](void *h, size_t bytes) {
        float best = 1e9f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(a);
            cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        return bytes / best / 1e6;
    };
    for (size_t mb : {(size_t)64, (size_t)256, (size_t)512}) {
        const size_t bytes = mb << 20;
        for (int rep = 0; rep < 2; rep++) {
            double t0 = now();
            void *h = nullptr;
            cudaError_t e = cudaHostAlloc(&h, bytes, cudaHostAllocPortable);
            double t1 = now();
            printf("{\"how\": \"cudaHostAlloc\", \"MiB\": %zu, \"alloc_ms\": %.1f, \"ok\": %d, \"d2h_GBps\": %.1f}\n", mb, (t1 - t0) * 1e3, e == cudaSuccess,
                   e == cudaSuccess ? d2h(h, bytes) : 0.0);
            t0 = now();
            if (e == cudaSuccess) cudaFreeHost(h);
            printf("{\"how\": \"cudaFreeHost\", \"MiB\": %zu, \"free_ms\": %.1f}\n", mb, (now() - t0) * 1e3);
            for (int huge = 1; huge >= 0; huge--)
                for (int T : {1, 16}) {
                    t0 = now();
                    char *q = (char *)mmap(nullptr, bytes + ((size_t)2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                    char *al = (char *)(((uintptr_t)q + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1));
                    if (huge) madvise(al, bytes, MADV_HUGEPAGE);
                    touch(al, bytes, T);
                    t1 = now();
                    e = cudaHostRegister(al, bytes, cudaHostRegisterPortable);
                    double t2 = now();
                    printf("{\"how\": \"mmap%s + touch(%d threads) + cudaHostRegister\", \"MiB\": %zu, \"touch_ms\": %.1f, \"register_ms\": %.1f, "
                           "\"total_ms\": %.1f, \"ok\": %d, \"d2h_GBps\": %.1f}\n",
                           huge ? " + MADV_HUGEPAGE" : "", T, mb, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, e == cudaSuccess,
                           e == cudaSuccess ? d2h(al, bytes) : 0.0);
                    t0 = now();
                    if (e == cudaSuccess) cudaHostUnregister(al);
                    munmap(q, bytes + ((size_t)2 << 20));
                    printf("{\"how\": \"unregister + munmap\", \"MiB\": %zu, \"free_ms\": %.1f}\n", mb, (now() - t0) * 1e3);
                }
        }
    }
    return 0;
}
