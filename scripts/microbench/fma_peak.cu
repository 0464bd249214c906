// FP32 FMA / MUFU issue-rate micro-benchmark for the roofline denominators ("measure with a FMA
// micro-benchmark on the box; do not quote vendor numbers unmeasured").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/fma_peak scripts/microbench/fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) peak_kernel(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                                   // FFMA
            else if (MODE == 1) x[i] = x[i] * a;                                      // FMUL
            else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));             // MUFU.EX2
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, double ops_per_inst, int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 2, threads = 1024;
    float* out;
    cudaMalloc(&out, size_t(blocks) * threads * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<MODE><<<blocks, threads>>>(out, iters / 10, 1.0001f, 1e-7f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    peak_kernel<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 1e-7f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = double(blocks) * threads * double(iters) * 8;
    const double rate = inst / (ms * 1e-3);
    printf("{\"op\": \"%s\", \"sms\": %d, \"thread_inst_per_s\": %.4e, \"warp_inst_per_clk_per_sm_at_1965MHz\": %.3f, \"tflops\": %.2f}\n",
           name, sms, rate, rate / 32 / sms / 1.965e9, rate * ops_per_inst / 1e12);
    cudaFree(out);
    return rate;
}

int main() {
    run<0>("FFMA", 2, 200000);
    run<1>("FMUL", 1, 200000);
    run<2>("MUFU.EX2", 1, 50000);
    return 0;
}
