// Shared helpers for the garmentnets_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/garmentnets_b200.h"

namespace gnb {

void set_error(const char* fmt, ...);

inline int32_t check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GNB_ERR_CUDA;
    }
    return GNB_OK;
}

#define GNB_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            ::gnb::set_error(__VA_ARGS__);          \
            return GNB_ERR_INVALID;                 \
        }                                           \
    } while (0)

#define GNB_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            ::gnb::set_error("%s: %s", #call, cudaGetErrorString(e__));         \
            return GNB_ERR_CUDA;                                                \
        }                                                                       \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

int sm_count();  // cached SM count of the current device (148 on B200)

// fp32 squared distance with every product and sum individually rounded (no FMA contraction), so the CPU
// oracle (numpy float32) reproduces it bit for bit:  ((dx*dx + dy*dy) + dz*dz).
__device__ __forceinline__ float sqdist_nofma(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace gnb
