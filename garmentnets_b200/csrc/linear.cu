// Per-point Linear -> ReLU -> BatchNorm(eval) (ref components/mlp.py:9-20) as an fp32 register-blocked GEMM with
// the bias / ReLU / BN-affine fused into the epilogue.  This is the general-shape fp32 path (exact fp32 FMA
// accumulation, any K, any row stride); the dominant dense contractions of the pipeline (implicit decoder, 3x3x3
// convolutions) have their own tensor-core kernels.
#include "common.cuh"

namespace gnb {

constexpr int LBM = 128, LBK = 16, LTHREADS = 256;

// Y[R,N] = post(X[R,K] * W[N,K]^T + bias);  tile BM=128 x BN (64 or 32), BK=16, 256 threads, 8 x (BN/16) per thread.
template <int BN, bool VEC>
__global__ void __launch_bounds__(LTHREADS)
linear_kernel(const float* __restrict__ X, int64_t R, int K, int64_t ldx, const float* __restrict__ Wt,
              const float* __restrict__ bias, int N, int relu, const float* __restrict__ bn_scale,
              const float* __restrict__ bn_shift, float* __restrict__ Y, int64_t ldy,
              const int64_t* __restrict__ rows_dev) {
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[LBK][LBM + 4];
    __shared__ __align__(16) float Bs[LBK][BN + 4];
    int64_t rows = R;
    if (rows_dev != nullptr) { const int64_t rd = *rows_dev; rows = rd < rows ? rd : rows; }
    const int64_t m0 = (int64_t)blockIdx.x * LBM;
    if (m0 >= rows) return;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // loader mapping: A tile 128 x 16 -> 2 float4 per thread (VEC) or 8 scalars; B tile BN x 16.
    float4 ra[2];
    float rb[BN * LBK / LTHREADS];
    float rs[8];

    auto load_tiles = [&](int k0) {
        if (VEC) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = (tid >> 2) + 64 * i, kk = (tid & 3) * 4;
                const int64_t gm = m0 + m;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gm < rows && k0 + kk < K) v = *reinterpret_cast<const float4*>(X + gm * ldx + k0 + kk);
                ra[i] = v;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = (tid >> 4) + 16 * i, kk = tid & 15;
                const int64_t gm = m0 + m;
                rs[i] = (gm < rows && k0 + kk < K) ? X[gm * ldx + k0 + kk] : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < BN * LBK / LTHREADS; ++i) {
            const int n = (tid >> 4) + 16 * i, kk = tid & 15;
            rb[i] = (n0 + n < N && k0 + kk < K) ? Wt[(int64_t)(n0 + n) * K + k0 + kk] : 0.f;
        }
    };
    auto store_tiles = [&]() {
        if (VEC) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = (tid >> 2) + 64 * i, kk = (tid & 3) * 4;
                As[kk + 0][m] = ra[i].x; As[kk + 1][m] = ra[i].y; As[kk + 2][m] = ra[i].z; As[kk + 3][m] = ra[i].w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) As[tid & 15][(tid >> 4) + 16 * i] = rs[i];
        }
#pragma unroll
        for (int i = 0; i < BN * LBK / LTHREADS; ++i) Bs[tid & 15][(tid >> 4) + 16 * i] = rb[i];
    };

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += LBK) {
        store_tiles();
        __syncthreads();
        if (k0 + LBK < K) load_tiles(k0 + LBK);
#pragma unroll
        for (int kk = 0; kk < LBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n >= N) continue;
        const float bv = bias ? bias[n] : 0.f;
        const float sc = bn_scale ? bn_scale[n] : 1.f;
        const float sh = bn_shift ? bn_shift[n] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t gm = m0 + ty * 8 + i;
            if (gm >= rows) continue;
            float v = acc[i][j] + bv;
            if (relu) v = fmaxf(v, 0.f);
            if (bn_scale) v = fmaf(v, sc, sh);
            Y[gm * ldy + n] = v;
        }
    }
}

// N <= 8 outputs (decoder heads 256->1 / 256->3): one warp per row, lanes stride K, shuffle reduction.
__global__ void __launch_bounds__(256)
linear_smalln_kernel(const float* __restrict__ X, int64_t R, int K, int64_t ldx, const float* __restrict__ Wt,
                     const float* __restrict__ bias, int N, int relu, const float* __restrict__ bn_scale,
                     const float* __restrict__ bn_shift, float* __restrict__ Y, int64_t ldy,
                     const int64_t* __restrict__ rows_dev) {
    int64_t rows = R;
    if (rows_dev != nullptr) { const int64_t rd = *rows_dev; rows = rd < rows ? rd : rows; }
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float xv = X[r * ldx + k];
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (n < N) acc[n] = fmaf(xv, __ldg(Wt + (int64_t)n * K + k), acc[n]);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (n < N) {
            float v = acc[n];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
                v += bias ? bias[n] : 0.f;
                if (relu) v = fmaxf(v, 0.f);
                if (bn_scale) v = fmaf(v, bn_scale[n], bn_shift[n]);
                Y[r * ldy + n] = v;
            }
        }
    }
}

// A handful of rows (the global head of PointNet++: 32 rows x K = 1024): the tiled kernel above would run as N/64 CTAs that
// each walk K sequentially (65 us for 2 MB of weights).  One CTA per output column instead: every warp takes rows
// warp, warp + 8, ... and reduces its 32 partial dot products with shuffles (weights of the column stay in L1).
__global__ void __launch_bounds__(256)
linear_smallr_kernel(const float* __restrict__ X, int64_t R, int K, int64_t ldx, const float* __restrict__ Wt,
                     const float* __restrict__ bias, int N, int relu, const float* __restrict__ bn_scale,
                     const float* __restrict__ bn_shift, float* __restrict__ Y, int64_t ldy,
                     const int64_t* __restrict__ rows_dev) {
    int64_t rows = R;
    if (rows_dev != nullptr) { const int64_t rd = *rows_dev; rows = rd < rows ? rd : rows; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x;
    const float* __restrict__ w = Wt + (int64_t)n * K;
    for (int64_t r = warp; r < rows; r += 8) {
        const float* __restrict__ x = X + r * ldx;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = lane;
        for (; k + 96 < K; k += 128) {
            a0 = fmaf(x[k], __ldg(w + k), a0);
            a1 = fmaf(x[k + 32], __ldg(w + k + 32), a1);
            a2 = fmaf(x[k + 64], __ldg(w + k + 64), a2);
            a3 = fmaf(x[k + 96], __ldg(w + k + 96), a3);
        }
        for (; k < K; k += 32) a0 = fmaf(x[k], __ldg(w + k), a0);
        float v = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            v += bias ? bias[n] : 0.f;
            if (relu) v = fmaxf(v, 0.f);
            if (bn_scale) v = fmaf(v, bn_scale[n], bn_shift[n]);
            Y[r * ldy + n] = v;
        }
    }
}

}  // namespace gnb

using namespace gnb;

extern "C" int32_t gnb_linear(const float* X, int64_t R, int32_t K, int64_t ldx, const float* W, const float* bias,
                              int32_t N, int32_t relu, const float* bn_scale, const float* bn_shift, float* Y,
                              int64_t ldy, const int64_t* rows_dev, void* stream) {
    GNB_REQUIRE(X && W && Y, "gnb_linear: null pointer");
    GNB_REQUIRE(R >= 0 && K > 0 && N > 0 && ldx >= K && ldy >= N, "gnb_linear: bad shape R=%lld K=%d N=%d ldx=%lld ldy=%lld",
                (long long)R, K, N, (long long)ldx, (long long)ldy);
    GNB_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), "gnb_linear: bn_scale/bn_shift must come together");
    if (R == 0) return GNB_OK;
    cudaStream_t st = as_stream(stream);
    if (N <= 8) {
        linear_smalln_kernel<<<(unsigned)ceil_div<int64_t>(R, 8), 256, 0, st>>>(X, R, K, ldx, W, bias, N, relu, bn_scale,
                                                                             bn_shift, Y, ldy, rows_dev);
        return check_launch("gnb_linear(small N)");
    }
    if (R <= 64) {
        linear_smallr_kernel<<<(unsigned)N, 256, 0, st>>>(X, R, K, ldx, W, bias, N, relu, bn_scale, bn_shift, Y, ldy, rows_dev);
        return check_launch("gnb_linear(few rows)");
    }
    const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const unsigned gx = (unsigned)ceil_div<int64_t>(R, LBM);
    if (N <= 32) {
        dim3 grid(gx, 1);
        if (vec) linear_kernel<32, true><<<grid, LTHREADS, 0, st>>>(X, R, K, ldx, W, bias, N, relu, bn_scale, bn_shift, Y, ldy, rows_dev);
        else linear_kernel<32, false><<<grid, LTHREADS, 0, st>>>(X, R, K, ldx, W, bias, N, relu, bn_scale, bn_shift, Y, ldy, rows_dev);
    } else {
        dim3 grid(gx, (unsigned)ceil_div(N, 64));
        if (vec) linear_kernel<64, true><<<grid, LTHREADS, 0, st>>>(X, R, K, ldx, W, bias, N, relu, bn_scale, bn_shift, Y, ldy, rows_dev);
        else linear_kernel<64, false><<<grid, LTHREADS, 0, st>>>(X, R, K, ldx, W, bias, N, relu, bn_scale, bn_shift, Y, ldy, rows_dev);
    }
    return check_launch("gnb_linear");
}
