// dmma_peak.cu — measures the FP64 tensor-pipe (DMMA.8x8x4) peak of the GPU it runs on: the roofline denominator of
// the kernels that multiply on that pipe (k_project_dmma, k_score_stream).  MEASURED_PEAKS.json has no FP64 entry.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_peak tools/dmma_peak.cu && tools/dmma_peak
// Each warp runs CH independent accumulator chains of mma.sync.m8n8k4.f64; warps/SM and chains are swept, the best
// rate is the peak.  Also prints the DFMA (CUDA-core FP64) rate for comparison.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.6789) out[0] = s;
}

template <int CH>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i];
    if (s == 12345.6789) out[0] = s;
}

template <class K>
double time_ms(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000;
    double best = 0;
    printf("%s, %d SMs\n", p.name, sms);
    for (int warps : {4, 8, 16, 32}) {
        auto run = [&](auto kern, int ch) {
            const double ms = time_ms([&] { kern<<<sms, warps * 32, 0>>>(out, iters, 1.0, 1.0); });
            const double tf = 2.0 * 256 * ch * (double)iters * warps * sms / (ms * 1e-3) / 1e12;
            printf("DMMA.8x8x4  %2d warps/SM  %d chains/warp  %8.3f ms  %6.2f TFLOP/s\n", warps, ch, ms, tf);
            if (tf > best) best = tf;
        };
        run(k_dmma<1>, 1); run(k_dmma<2>, 2); run(k_dmma<4>, 4); run(k_dmma<8>, 8);
    }
    double bestf = 0;
    for (int warps : {8, 16, 32}) {
        const double ms = time_ms([&] { k_dfma<8><<<sms, warps * 32, 0>>>(out, iters, 1.0, 1.0); });
        const double tf = 2.0 * 32 * 8 * (double)iters * warps * sms / (ms * 1e-3) / 1e12;
        printf("DFMA        %2d warps/SM  8 chains/thread  %8.3f ms  %6.2f TFLOP/s\n", warps, ms, tf);
        if (tf > bestf) bestf = tf;
    }
    printf("{\"fp64_dmma_tflops\": %.2f, \"fp64_dfma_tflops\": %.2f, \"sms\": %d}\n", best, bestf, sms);
    return 0;
}
