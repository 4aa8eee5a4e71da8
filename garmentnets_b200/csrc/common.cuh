// Shared helpers for the garmentnets_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
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
uint32_t* f16_flag_ptr();  // per-device flag: an fp32 value outside the fp16 range reached a saturating operand split (capi.cu)

// Serialises the users of one __constant__ operand bank.  Three launch families (lattice decode, query / row decode,
// Linear blocks) refresh a per-device constant array with a stream-ordered copy in front of every launch; on ONE stream
// that is ordered by the stream itself.  A launch of the same family on ANOTHER stream (or from another host thread)
// would overwrite the bank under a running kernel, so the guard (a) holds a host mutex from the copy to the launch and
// (b) makes the new stream wait, on the device, for the event the previous user recorded after its kernel.  Same-stream
// sequences pay one cudaEventRecord per launch; nothing is serialised on the host.
enum ConstBank { BANK_LATTICE = 0, BANK_DECODE_TC = 1, BANK_LINEAR_TC = 2, BANK_DECODE_QUERY = 3, BANK_SA_MLP = 4, BANK_COUNT = 5 };
class ConstBankGuard {
public:
    ConstBankGuard(ConstBank bank, cudaStream_t st);
    ~ConstBankGuard();
    ConstBankGuard(const ConstBankGuard&) = delete;
    ConstBankGuard& operator=(const ConstBankGuard&) = delete;
private:
    int bank_, dev_;
    cudaStream_t st_;
};

// Debug / profiling knock-out switches of the tensor-core decoders (GNB_DL2_DBG, GNB_TC_DBG) exist only in builds with
// -DGNB_PROFILE_KNOBS; the shipped library never reads the environment on the launch path.
inline int profile_knob(const char* name) {
#ifdef GNB_PROFILE_KNOBS
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
#else
    (void)name;
    return 0;
#endif
}

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
