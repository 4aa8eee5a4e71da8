// Micro-benchmark: issue rate of the fp32 -> fp16x2 conversion (F2FP.PACK_AB), the half -> float unpack (HADD2.F32), FMNMX and
// the whole relu + hi/lo split sequence of the decoders' A producers on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cvt_bench cvt_bench.cu && ./cvt_bench
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float a, float b) {
    float x[16];
    unsigned h[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a * (i + 1) + threadIdx.x * b;
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0x3c003c00u + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {            // 8 F2FP per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) { unsigned r; asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[2 * i + 1]), "f"(x[2 * i])); x[2 * i] = __uint_as_float(r | 0x3f000000u); }
        } else if (MODE == 1) {     // 16 HADD2.F32 (half -> float) per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[i])); h[i] = __float_as_uint(f.x) ^ (__float_as_uint(f.y) >> 3); }
        } else if (MODE == 2) {     // 16 FMNMX per iteration
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i] * 1.0001f, b);
        } else {                    // the producers' sequence per channel pair: cvt.rz.relu, 2 x unpack, sub2, cvt.rn.relu (8 pairs)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned hi, lo;
                asm volatile("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                const float r0 = x[2 * i] - hf.x, r1 = x[2 * i + 1] - hf.y;
                asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
                x[2 * i] = x[2 * i] * 1.0001f + __uint_as_float((hi ^ lo) & 0xffu);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += x[i];
    for (int i = 0; i < 8; ++i) s += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads, double per_iter) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = 148.0 * threads * ITERS * per_iter;
    printf("%-34s threads/SM=%4d  %.3f ms  %.1f thread-ops/clk/SM at 1.9 GHz\n", name, threads, ms, n / ms / 1e3 / 148 / 1.9e6);
    cudaFree(out);
}
int main() {
    for (int t : {256, 1024}) {
        run<0>("F2FP.PACK_AB (cvt f16x2.f32)", t, 8);
        run<1>("HADD2.F32 (half -> float)", t, 16);
        run<2>("FMNMX (+FMUL)", t, 16);
        run<3>("relu + hi/lo split per pair", t, 8);
    }
    return 0;
}
