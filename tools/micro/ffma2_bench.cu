// Micro-benchmark: issue rate of FFMA vs FFMA2 (packed fp32x2, with and without a broadcast scalar operand) on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float a, float b) {
    float x[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    unsigned long long w, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(w), "l"(w));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(w), "l"(bb));
        }
    }
    float s = 0;
    if (MODE == 0) { for (int i = 0; i < 16; ++i) s += x[i]; }
    else { for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; } }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * threads * ITERS * 16;
    printf("%-28s threads/SM=%4d  %.3f ms  %.1f TFMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, threads, ms, fma / ms / 1e9, fma / ms / 1e3 / 148 / 1.9e6);
    cudaFree(out);
}
int main() {
    for (int t : {128, 256, 512, 1024}) { run<0>("FFMA", t); run<1>("FFMA2 packed*packed", t); run<2>("FFMA2 packed*broadcast", t); }
    return 0;
}
