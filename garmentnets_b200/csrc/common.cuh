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

// ---- trilinear sampling set-up (grid_sample: bilinear, padding border, align_corners), shared by gridding.cu and
// decode_tc.cu so that both paths blend with identical weights
struct TriW { int64_t off[8]; float w[8]; };

// coordinates in [-1,1] along (W,H,D) -> 8 corner offsets (in voxels*C units) and weights, ATen order
// tnw,tne,tsw,tse,bnw,bne,bsw,bse.  Out-of-range corners (index == size) get weight 0 and a clamped offset.
__device__ __forceinline__ void trilinear_setup(float gx, float gy, float gz, int D, int H, int W, int C, TriW& t) {
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    float iz = ((gz + 1.f) / 2.f) * (float)(D - 1);
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    iz = fminf((float)(D - 1), fmaxf(iz, 0.f));
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const float x1w = ix - fx, y1w = iy - fy, z1w = iz - fz;          // (ix - ix_tnw)
    const float x0w = (fx + 1.f) - ix, y0w = (fy + 1.f) - iy, z0w = (fz + 1.f) - iz;  // (ix_bse - ix)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
        const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        const bool ok = xx < W && yy < H && zz < D;
        const float w = (dx ? x1w : x0w) * (dy ? y1w : y0w) * (dz ? z1w : z0w);
        t.w[k] = ok ? w : 0.f;
        const int xc = xx < W ? xx : W - 1, yc = yy < H ? yy : H - 1, zc = zz < D ? zz : D - 1;
        t.off[k] = (((int64_t)zc * H + yc) * W + xc) * C;
    }
}

}  // namespace gnb
