// gather_patterns.cu — micro-benchmark behind the bucket-major re-rank design: HBM rate of gathering
// random 1 KB rows (d = 128 FP64) under the access patterns a DMMA consumer can use.
//   A: MMA-fragment pattern — lane (g,t) loads 16 B at row g, column chunk 4w+t: 8 rows x 64 B per instruction
//   B: row pattern          — a warp instruction covers 512 contiguous bytes of one row
//   C: TMA bulk             — one cp.async.bulk of the whole row into shared memory per lane, mbarrier completion
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_patterns gather_patterns.cu && ./gather_patterns
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int D = 128;

template <int BLOCKS>   // 8-row blocks in flight per warp
__global__ void __launch_bounds__(256) k_frag(const double* __restrict__ X, const int* __restrict__ ids, int64_t nrows,
                                              double* __restrict__ out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    double acc = 0;
    for (int64_t r0 = warp * 8 * BLOCKS; r0 < nrows; r0 += nwarps * 8 * BLOCKS) {
        double2 a[BLOCKS][16];
#pragma unroll
        for (int b = 0; b < BLOCKS; ++b) {
            const int64_t r = min(r0 + b * 8 + g, nrows - 1);
            const double* xr = X + (int64_t)__ldg(ids + r) * D + 2 * t;
#pragma unroll
            for (int w = 0; w < 16; ++w) a[b][w] = __ldg(reinterpret_cast<const double2*>(xr + 8 * w));
        }
#pragma unroll
        for (int b = 0; b < BLOCKS; ++b)
#pragma unroll
            for (int w = 0; w < 16; ++w) acc += a[b][w].x + a[b][w].y;
    }
    if (acc == 123.456) out[0] = acc;
}

template <int ROWS>   // rows in flight per warp
__global__ void __launch_bounds__(256) k_row(const double* __restrict__ X, const int* __restrict__ ids, int64_t nrows,
                                             double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    double acc = 0;
    for (int64_t r0 = warp * ROWS; r0 < nrows; r0 += nwarps * ROWS) {
        double2 a[ROWS][2];
#pragma unroll
        for (int b = 0; b < ROWS; ++b) {
            const int64_t r = min(r0 + b, nrows - 1);
            const double2* xr = reinterpret_cast<const double2*>(X + (int64_t)__ldg(ids + r) * D);
            a[b][0] = __ldg(xr + lane);
            a[b][1] = __ldg(xr + 32 + lane);
        }
#pragma unroll
        for (int b = 0; b < ROWS; ++b) acc += a[b][0].x + a[b][0].y + a[b][1].x + a[b][1].y;
    }
    if (acc == 123.456) out[0] = acc;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(bar)),
        "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// per warp: STAGES stages of RPS rows (pitch 1088 B); lane r < RPS issues the copy of row r of the stage
template <int STAGES, int RPS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tma(const double* __restrict__ X, const int* __restrict__ ids, int64_t nrows,
                                                    double* __restrict__ out) {
    constexpr int PITCH = 136;
    extern __shared__ __align__(128) double sm[];
    __shared__ uint64_t bars[WARPS][STAGES];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double* ring = sm + (size_t)wl * STAGES * RPS * PITCH;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[wl][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int64_t warp = (int64_t)blockIdx.x * WARPS + wl;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    // this warp's rows: chunks of RPS at r0 = (warp + i*nwarps)*RPS
    const int64_t nchunks = (nrows / RPS - warp + nwarps - 1) / nwarps;
    auto issue = [&](int64_t i) {
        const int s = (int)(i % STAGES);
        const int64_t r0 = (warp + i * nwarps) * RPS;
        if (lane == 0) mbar_expect_tx(&bars[wl][s], RPS * D * 8);
        __syncwarp();
        if (lane < RPS) bulk_g2s(ring + (s * RPS + lane) * PITCH, X + (int64_t)__ldg(ids + r0 + lane) * D, D * 8, &bars[wl][s]);
    };
    for (int64_t i = 0; i < STAGES - 1 && i < nchunks; ++i) issue(i);
    double acc = 0;
    for (int64_t i = 0; i < nchunks; ++i) {
        if (i + STAGES - 1 < nchunks) issue(i + STAGES - 1);
        const int s = (int)(i % STAGES);
        mbar_wait(&bars[wl][s], (unsigned)((i / STAGES) & 1));
        for (int rb = 0; rb < RPS; rb += 8) {
            const double* xr = ring + (s * RPS + rb + g) * PITCH + 2 * t;
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                const double2 v = *reinterpret_cast<const double2*>(xr + 8 * w);
                acc += v.x + v.y;
            }
        }
        __syncwarp();
    }
    if (acc == 123.456) out[0] = acc;
}

int main() {
    const int64_t N = 1000000, nrows = 8 << 20;   // 8M row fetches = 8.6 GB
    double* X;
    int* ids;
    double* out;
    CK(cudaMalloc(&X, N * D * 8));
    CK(cudaMemset(X, 0, N * D * 8));
    CK(cudaMalloc(&out, 8));
    std::vector<int> h(nrows);
    uint64_t s = 88172645463325252ULL;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int)(s % N); }
    CK(cudaMalloc(&ids, nrows * 4));
    CK(cudaMemcpy(ids, h.data(), nrows * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        launch();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("%-44s %8.3f ms  %7.1f GB/s\n", name, best, nrows * D * 8.0 / best / 1e6);
    };
    for (int cps : {2, 4, 8}) {
        const int grid = 148 * cps;
        char nm[128];
        snprintf(nm, 128, "A frag 8x64B, 1 block/warp, %d warps/SM", cps * 8);
        run(nm, [&] { k_frag<1><<<grid, 256>>>(X, ids, nrows, out); });
        snprintf(nm, 128, "A frag 8x64B, 2 blocks/warp, %d warps/SM", cps * 8);
        run(nm, [&] { k_frag<2><<<grid, 256>>>(X, ids, nrows, out); });
        snprintf(nm, 128, "B row 512B, 4 rows/warp, %d warps/SM", cps * 8);
        run(nm, [&] { k_row<4><<<grid, 256>>>(X, ids, nrows, out); });
        snprintf(nm, 128, "B row 512B, 8 rows/warp, %d warps/SM", cps * 8);
        run(nm, [&] { k_row<8><<<grid, 256>>>(X, ids, nrows, out); });
    }
    {
        constexpr int W = 4;
        auto k = k_tma<4, 8, W>;
        const size_t smem = (size_t)W * 4 * 8 * 136 * 8;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int cps : {1, 2}) {
            char nm[128];
            snprintf(nm, 128, "C TMA 1KB rows, 4 stages x 8 rows, %d warps/SM", cps * W);
            run(nm, [&] { k<<<148 * cps, W * 32, smem>>>(X, ids, nrows, out); });
        }
    }
    {
        constexpr int W = 4;
        auto k = k_tma<3, 16, W>;
        const size_t smem = (size_t)W * 3 * 16 * 136 * 8;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("C TMA 1KB rows, 3 stages x 16 rows, 4 warps/SM", [&] { k<<<148, W * 32, smem>>>(X, ids, nrows, out); });
    }
    {
        constexpr int W = 8;
        auto k = k_tma<3, 8, W>;
        const size_t smem = (size_t)W * 3 * 8 * 136 * 8;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        run("C TMA 1KB rows, 3 stages x 8 rows, 8 warps/SM", [&] { k<<<148, W * 32, smem>>>(X, ids, nrows, out); });
    }
    return 0;
}
