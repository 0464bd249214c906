// Issue-rate micro-benchmark of the sm_100 packed Float32 instructions (FFMA2 / FADD2 / FMUL2 from
// fma.rn.f32x2, add.f32x2, mul.f32x2) beside scalar FFMA, FMNMX (ALU pipe) and mixed streams: does a packed
// instruction cost one issue slot for two results, and does it co-issue with the ALU / MUFU streams of an
// issue-bound kernel?  (k1d_tile_kernel is issue-bound with the FMA pipe a third busy: DESIGN.md §3.4.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/f32x2_peak scripts/microbench/f32x2_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 r) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); return a + b; }

// MODE: 0 FFMA, 1 FFMA2, 2 FADD2, 3 FMUL2, 4 FMNMX, 5 FFMA + FMNMX (1:1), 6 FFMA2 + FMNMX (1:1), 7 FFMA2 + MUFU (4:1),
//       8 FFMA + MUFU (4:1), 9 FFMA2 + FFMA (1:1)
template <int MODE>
__global__ void __launch_bounds__(1024) peak_kernel(float* out, int iters, float a, float b) {
    float x[8];
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; p[i] = pk(x[i], x[i] + 0.5f); }
    const u64 pa = pk(a, a), pb = pk(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            if (MODE == 4) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (MODE == 5) { if (i & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); }
            if (MODE == 6) { if (i & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b)); else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); }
            if (MODE == 7) { if (i == 3 || i == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); }
            if (MODE == 8) { if (i == 3 || i == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); }
            if (MODE == 9) { if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + lo(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flops_per_loop_body, int iters) {
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
    const double inst = double(blocks) * threads * double(iters) * 8;  // thread-instructions (8 per loop body)
    const double rate = inst / (ms * 1e-3);
    printf("{\"stream\": \"%s\", \"sms\": %d, \"warp_inst_per_clk_per_sm_at_1965MHz\": %.3f, \"tflops\": %.2f, \"ms\": %.3f}\n", name, sms,
           rate / 32 / sms / 1.965e9, double(blocks) * threads * double(iters) * flops_per_loop_body / (ms * 1e-3) / 1e12, ms);
    cudaFree(out);
}

// dependent-issue latency: ONE warp per SM runs a single dependent chain; cycles per instruction from clock64()
template <int MODE>
__global__ void latency_kernel(float* out, long long* cyc, int iters, float a, float b) {
    float x = threadIdx.x * 1e-3f;
    u64 p = pk(x, x + 0.5f);
    const u64 pa = pk(a, a), pb = pk(b, b);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(a), "f"(b));
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(pa), "l"(pb));
            if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pb));
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + lo(p);
}
template <int MODE>
void latency(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 32 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 20000;
    latency_kernel<MODE><<<148, 32>>>(out, cyc, iters, 1.0001f, 1e-7f);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"dependent_chain\": \"%s\", \"cycles_per_instruction\": %.2f}\n", name, double(h[0]) / (double(iters) * 16));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    latency<0>("FFMA"); latency<1>("FFMA2"); latency<2>("FADD2"); latency<3>("MUFU.EX2");
    const int N = 100000;
    run<0>("FFMA", 16, N);
    run<1>("FFMA2", 32, N);
    run<2>("FADD2", 16, N);
    run<3>("FMUL2", 16, N);
    run<4>("FMNMX", 8, N);
    run<5>("FFMA+FMNMX 1:1", 4 * 2 + 4, N);
    run<6>("FFMA2+FMNMX 1:1", 4 * 4 + 4, N);
    run<7>("FFMA2+MUFU 6:2", 6 * 4 + 2, N / 2);
    run<8>("FFMA+MUFU 6:2", 6 * 2 + 2, N / 2);
    run<9>("FFMA2+FFMA 1:1", 4 * 4 + 4 * 2, N);
    return 0;
}
