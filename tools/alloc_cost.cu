// alloc_cost.cu — cost of device allocations behind the DevBuf policy: cudaMalloc vs cudaMallocAsync
// from the default pool, cold (pool has to grow) and warm (block cached in the pool).
#include <cuda_runtime.h>
#include <chrono>
#include <stdio.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaFree(0);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaMemPool_t pool;
    cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long keep = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    for (size_t mb : {8, 128, 1280}) {
        const size_t bytes = mb << 20;
        void* p;
        double t0 = now();
        cudaMalloc(&p, bytes);
        double t1 = now();
        cudaMemsetAsync(p, 0, bytes, st);
        cudaStreamSynchronize(st);
        double t2 = now();
        cudaFree(p);
        double t3 = now();
        printf("%5zu MB cudaMalloc %.3f ms, first touch (memset) %.3f ms, cudaFree %.3f ms\n", mb, t1 - t0, t2 - t1, t3 - t2);
        for (int rep = 0; rep < 3; ++rep) {
            t0 = now();
            cudaMallocAsync(&p, bytes, st);
            cudaStreamSynchronize(st);
            t1 = now();
            cudaMemsetAsync(p, 0, bytes, st);
            cudaStreamSynchronize(st);
            t2 = now();
            cudaFreeAsync(p, st);
            cudaStreamSynchronize(st);
            t3 = now();
            printf("%5zu MB cudaMallocAsync[%d] %.3f ms, memset %.3f ms, cudaFreeAsync %.3f ms\n", mb, rep, t1 - t0, t2 - t1, t3 - t2);
        }
    }
    // host -> device copy rate from pinned memory, 1 GB
    void *h, *d;
    cudaMallocHost(&h, 1u << 30);
    cudaMalloc(&d, 1u << 30);
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now();
        cudaMemcpyAsync(d, h, 1u << 30, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        printf("H2D 1 GiB pinned: %.2f ms\n", now() - t0);
    }
    return 0;
}
