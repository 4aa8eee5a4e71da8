// Gridding kernels: scatter-reduce into the voxel grid, aggregator feature assembly, NOCS bin head,
// trilinear sampling of the feature volume and the Gaussian-gradient-magnitude stencil.  All are HBM/L2-bound
// elementwise, gather or scatter kernels: channels are the fastest axis so every warp access is coalesced.
//
// Reference semantics restated:
//   networks/conv_implicit_wnf.py:92-94, components/gridding.py:32-35   torch_scatter.scatter  -> gnb_scatter_reduce
//   networks/conv_implicit_wnf.py:62-85 + components/gridding.py:161-256                        -> gnb_aggregator_features
//   networks/conv_implicit_wnf.py:222-231                                                       -> gnb_nocs_head
//   networks/conv_implicit_wnf.py:135-142, components/gridding.py:45-98  F.grid_sample          -> gnb_trilinear_sample*
//   predict.py:162-163  scipy.ndimage.gaussian_gradient_magnitude                               -> gnb_gaussian_gradient_magnitude
#include "common.cuh"
#include <math.h>

namespace gnb {

// ---- scatter -------------------------------------------------------------------------------------
// Order-preserving float <-> uint encoding: enc(a) < enc(b)  <=>  a < b, and every real value encodes to > 0,
// so a zero-filled output doubles as "nothing received yet" AND as the final 0.0f of empty slots.
__device__ __forceinline__ unsigned enc_f32(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned e) {
    unsigned b = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    return __uint_as_float(b);
}

template <int REDUCE>
__global__ void __launch_bounds__(256)
scatter_accum_kernel(const float* __restrict__ src, int64_t src_sc, int64_t src_sn, const int64_t* __restrict__ index,
                     int64_t N, int C, int64_t dim_size, float* __restrict__ out, int64_t out_sc, int64_t out_sm,
                     int32_t* __restrict__ scratch) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t n = t / C;
    const int c = (int)(t - n * C);
    const int64_t m = index[n];
    if (m < 0 || m >= dim_size) return;
    const float v = src[c * src_sc + n * src_sn];
    float* o = out + c * out_sc + m * out_sm;
    if (REDUCE == GNB_REDUCE_MAX) {
        atomicMax(reinterpret_cast<unsigned*>(o), enc_f32(v));
        if (c == 0) atomicMin(scratch + m, (int32_t)n);
    } else {
        atomicAdd(o, v);
        if (c == 0) atomicAdd(scratch + m, 1);
    }
}

// decode only the slots that received something; the lowest contributing row of each slot does it.
__global__ void __launch_bounds__(256)
scatter_decode_kernel(const int64_t* __restrict__ index, int64_t N, int C, int64_t dim_size, float* __restrict__ out,
                      int64_t out_sc, int64_t out_sm, const int32_t* __restrict__ scratch) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t n = t / C;
    const int c = (int)(t - n * C);
    const int64_t m = index[n];
    if (m < 0 || m >= dim_size) return;
    if (scratch[m] != (int32_t)n) return;
    float* o = out + c * out_sc + m * out_sm;
    const unsigned e = *reinterpret_cast<unsigned*>(o);
    *o = dec_f32(e);
}

__global__ void scatter_mean_div_kernel(float* __restrict__ out, int64_t out_sc, int64_t out_sm, int C,
                                        int64_t dim_size, const int32_t* __restrict__ cnt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= dim_size * C) return;
    const int64_t m = t / C;
    const int c = (int)(t - m * C);
    const int k = cnt[m];
    if (k > 1) out[c * out_sc + m * out_sm] = __fdiv_rn(out[c * out_sc + m * out_sm], (float)k);
}

// MIN is computed as -MAX(-x) so the zero sentinel keeps working.
__global__ void __launch_bounds__(256)
scatter_min_accum_kernel(const float* __restrict__ src, int64_t src_sc, int64_t src_sn,
                         const int64_t* __restrict__ index, int64_t N, int C, int64_t dim_size,
                         float* __restrict__ out, int64_t out_sc, int64_t out_sm, int32_t* __restrict__ scratch) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t n = t / C;
    const int c = (int)(t - n * C);
    const int64_t m = index[n];
    if (m < 0 || m >= dim_size) return;
    const float v = -src[c * src_sc + n * src_sn];
    atomicMax(reinterpret_cast<unsigned*>(out + c * out_sc + m * out_sm), enc_f32(v));
    if (c == 0) atomicMin(scratch + m, (int32_t)n);
}
__global__ void __launch_bounds__(256)
scatter_min_decode_kernel(const int64_t* __restrict__ index, int64_t N, int C, int64_t dim_size,
                          float* __restrict__ out, int64_t out_sc, int64_t out_sm,
                          const int32_t* __restrict__ scratch) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t n = t / C;
    const int c = (int)(t - n * C);
    const int64_t m = index[n];
    if (m < 0 || m >= dim_size) return;
    if (scratch[m] != (int32_t)n) return;
    float* o = out + c * out_sc + m * out_sm;
    *o = -dec_f32(*reinterpret_cast<unsigned*>(o));
}

// ---- aggregator features ---------------------------------------------------------------------------
struct Corners { float lc[3], uc[3]; };

__global__ void __launch_bounds__(256)
aggregator_features_kernel(const float* __restrict__ feat, int64_t ldf, int Cf, const float* __restrict__ nocs,
                           const float* __restrict__ sim, const float* __restrict__ conf,
                           const int64_t* __restrict__ batch, int64_t N, int G, Corners cr, int with_point, int with_conf,
                           int64_t* __restrict__ flat_idx, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    float* dst = out + n * ldo;
    for (int c = lane; c < Cf; c += 32) dst[c] = feat[n * ldf + c];
    if (lane < 3) {
        const float gm1 = (float)(G - 1);
        const float ext = __fsub_rn(cr.uc[lane], cr.lc[lane]);
        const float p = nocs[n * 3 + lane];
        // components/gridding.py:173-181: ((p + (-lc)) * ((shape-1)/(uc-lc))).to(int64), clamp
        const float f = __fmul_rn(__fadd_rn(p, -cr.lc[lane]), __fdiv_rn(gm1, ext));
        long long i = (long long)f;  // trunc toward zero, like Tensor.to(int64)
        i = i < 0 ? 0 : (i > G - 1 ? G - 1 : i);
        int col = Cf;
        if (with_point) {   // include_point_feature: offset inside the voxel + the simulation points
            // components/gridding.py:249-255: idx * ((uc-lc)/(shape-1)) + lc
            const float origin = __fadd_rn(__fmul_rn((float)i, __fdiv_rn(ext, gm1)), cr.lc[lane]);
            dst[col + lane] = __fsub_rn(p, origin);
            dst[col + 3 + lane] = sim[n * 3 + lane];
            col += 6;
        }
        if (with_conf) dst[col + lane] = conf[n * 3 + lane];   // include_confidence_feature
        const long long i0 = __shfl_sync(0x7u, i, 0), i1 = __shfl_sync(0x7u, i, 1), i2 = __shfl_sync(0x7u, i, 2);
        if (lane == 0) flat_idx[n] = ((batch[n] * G + i0) * G + i1) * G + i2;
    }
}

// ---- NOCS bin head ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nocs_head_kernel(const float* __restrict__ logits, int64_t R, int bins, int64_t* __restrict__ bin,
                 float* __restrict__ conf, float* __restrict__ nocs) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const float* row = logits + r * (int64_t)bins * 3;
    const float inv = __fdiv_rn(1.0f, (float)(bins - 1));
    // the three axes run side by side (three independent load / arg-max / exp-sum chains per lane instead of three serial
    // passes over the row); per axis the operation order is unchanged
    float best[3] = {-INFINITY, -INFINITY, -INFINITY};
    int bi[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int k = lane; k < bins; k += 32) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = row[k * 3 + a];
            if (v > best[a]) { best[a] = v; bi[a] = k; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float ov = __shfl_xor_sync(0xffffffffu, best[a], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi[a], o);
            if (ov > best[a] || (ov == best[a] && oi < bi[a])) { best[a] = ov; bi[a] = oi; }
        }
    }
    float s[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < bins; k += 32) {
#pragma unroll
        for (int a = 0; a < 3; ++a) s[a] += expf(row[k * 3 + a] - best[a]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) s[a] += __shfl_xor_sync(0xffffffffu, s[a], o);
    }
    if (lane < 3) {   // every lane holds all three results: lane a writes axis a
        const int b_ = lane == 0 ? bi[0] : (lane == 1 ? bi[1] : bi[2]);
        const float s_ = lane == 0 ? s[0] : (lane == 1 ? s[1] : s[2]);
        bin[r * 3 + lane] = b_;
        conf[r * 3 + lane] = __fdiv_rn(1.0f, s_);
        nocs[r * 3 + lane] = __fmul_rn((float)b_, inv);
    }
}

// ---- trilinear sampling (grid_sample: bilinear, padding border, align_corners) ------------------------
template <bool GRID>
__global__ void __launch_bounds__(256)
trilinear_kernel(const float* __restrict__ vol, int b_fixed, int D, int H, int W, int C, const float* __restrict__ q,
                 int Q, int64_t m0, int64_t M, int64_t rows, int flip, int post, const float* __restrict__ bn_scale,
                 const float* __restrict__ bn_shift, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float q0, q1, q2;
    int b;
    if (GRID) {
        b = b_fixed;
        const int64_t m = m0 + r;
        const int k = (int)(m % Q);
        const int j = (int)((m / Q) % Q);
        const int i = (int)(m / ((int64_t)Q * Q));
        const float s = __fdiv_rn(1.0f, (float)(Q - 1));  // components/gridding.py:150-158
        q0 = __fmul_rn((float)i, s); q1 = __fmul_rn((float)j, s); q2 = __fmul_rn((float)k, s);
    } else {
        b = (int)(r / M);
        q0 = q[r * 3]; q1 = q[r * 3 + 1]; q2 = q[r * 3 + 2];
    }
    // 2q-1 (networks/conv_implicit_wnf.py:135); un-flipped: coordinate 0 -> W axis
    const float g0 = __fsub_rn(__fmul_rn(2.0f, q0), 1.0f), g1 = __fsub_rn(__fmul_rn(2.0f, q1), 1.0f),
                g2 = __fsub_rn(__fmul_rn(2.0f, q2), 1.0f);
    TriW t;
    if (flip) trilinear_setup(g2, g1, g0, D, H, W, C, t);
    else trilinear_setup(g0, g1, g2, D, H, W, C, t);
    const float* vb = vol + (int64_t)b * D * H * W * C;
    float* dst = out + r * ldo;
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += vb[t.off[k] + c] * t.w[k];
        if (post) acc = fmaxf(acc, 0.f) * bn_scale[c] + bn_shift[c];
        dst[c] = acc;
    }
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_scatter_reduce(const float* src, int64_t src_sc, int64_t src_sn, const int64_t* index, int64_t N,
                           int32_t C, int64_t dim_size, int32_t reduce, float* out, int64_t out_sc, int64_t out_sm,
                           int32_t* scratch, void* stream) {
    GNB_REQUIRE(out && scratch && (N == 0 || (src && index)), "gnb_scatter_reduce: null pointer");
    GNB_REQUIRE(C > 0 && dim_size >= 0 && N >= 0 && N < (1ll << 31), "gnb_scatter_reduce: bad size");
    GNB_REQUIRE((out_sc == 1 && out_sm == C) || (out_sc == dim_size && out_sm == 1) || dim_size <= 1 || C == 1,
                "gnb_scatter_reduce: out must be a dense [dim_size,C] or [C,dim_size] block");
    GNB_REQUIRE(reduce >= 0 && reduce <= 3, "gnb_scatter_reduce: unknown reduce %d", reduce);
    cudaStream_t st = as_stream(stream);
    GNB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C * (size_t)dim_size, st));
    const bool arg = reduce == GNB_REDUCE_MAX || reduce == GNB_REDUCE_MIN;
    GNB_CUDA(cudaMemsetAsync(scratch, arg ? 0x7f : 0, sizeof(int32_t) * (size_t)dim_size, st));
    if (N == 0 || dim_size == 0) return GNB_OK;
    const unsigned grid = (unsigned)ceil_div<int64_t>(N * C, 256);
    switch (reduce) {
        case GNB_REDUCE_MAX:
            scatter_accum_kernel<GNB_REDUCE_MAX><<<grid, 256, 0, st>>>(src, src_sc, src_sn, index, N, C, dim_size, out,
                                                                      out_sc, out_sm, scratch);
            scatter_decode_kernel<<<grid, 256, 0, st>>>(index, N, C, dim_size, out, out_sc, out_sm, scratch);
            break;
        case GNB_REDUCE_MIN:
            scatter_min_accum_kernel<<<grid, 256, 0, st>>>(src, src_sc, src_sn, index, N, C, dim_size, out, out_sc,
                                                           out_sm, scratch);
            scatter_min_decode_kernel<<<grid, 256, 0, st>>>(index, N, C, dim_size, out, out_sc, out_sm, scratch);
            break;
        default:
            scatter_accum_kernel<GNB_REDUCE_SUM><<<grid, 256, 0, st>>>(src, src_sc, src_sn, index, N, C, dim_size, out,
                                                                      out_sc, out_sm, scratch);
            if (reduce == GNB_REDUCE_MEAN)
                scatter_mean_div_kernel<<<(unsigned)ceil_div<int64_t>(dim_size * C, 256), 256, 0, st>>>(
                    out, out_sc, out_sm, C, dim_size, scratch);
            break;
    }
    return check_launch("gnb_scatter_reduce");
}

int32_t gnb_aggregator_features(const float* feat, int64_t ldf, int32_t Cf, const float* nocs,
                                const float* sim_points, const float* conf, const int64_t* batch, int64_t N,
                                int32_t G, const float* lower_corner, const float* upper_corner,
                                int32_t include_point_feature, int32_t include_confidence_feature,
                                int64_t* flat_idx, float* out, int64_t ldo, void* stream) {
    GNB_REQUIRE(feat && nocs && batch && flat_idx && out, "gnb_aggregator_features: null pointer");
    GNB_REQUIRE(!include_point_feature || sim_points, "gnb_aggregator_features: include_point_feature needs sim_points");
    GNB_REQUIRE(!include_confidence_feature || conf, "gnb_aggregator_features: include_confidence_feature needs conf");
    const int cols = Cf + (include_point_feature ? 6 : 0) + (include_confidence_feature ? 3 : 0);
    GNB_REQUIRE(G >= 2 && ldo >= cols, "gnb_aggregator_features: bad G/ldo");
    if (N == 0) return GNB_OK;
    Corners cr = {{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
    for (int a = 0; a < 3; ++a) {
        if (lower_corner) cr.lc[a] = lower_corner[a];
        if (upper_corner) cr.uc[a] = upper_corner[a];
        GNB_REQUIRE(cr.uc[a] > cr.lc[a], "gnb_aggregator_features: upper corner must lie above the lower corner");
    }
    aggregator_features_kernel<<<(unsigned)ceil_div<int64_t>(N, 8), 256, 0, as_stream(stream)>>>(
        feat, ldf, Cf, nocs, sim_points, conf, batch, N, G, cr, include_point_feature ? 1 : 0,
        include_confidence_feature ? 1 : 0, flat_idx, out, ldo);
    return check_launch("gnb_aggregator_features");
}

int32_t gnb_nocs_head(const float* logits, int64_t R, int32_t bins, int64_t* bin, float* conf, float* nocs,
                      void* stream) {
    GNB_REQUIRE(logits && bin && conf && nocs, "gnb_nocs_head: null pointer");
    GNB_REQUIRE(bins >= 2, "gnb_nocs_head: bins < 2");
    if (R == 0) return GNB_OK;
    nocs_head_kernel<<<(unsigned)ceil_div<int64_t>(R, 8), 256, 0, as_stream(stream)>>>(logits, R, bins, bin, conf, nocs);
    return check_launch("gnb_nocs_head");
}

int32_t gnb_trilinear_sample(const float* vol, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, const float* q,
                             int64_t M, int32_t flip, int32_t post, const float* bn_scale, const float* bn_shift,
                             float* out, int64_t ldo, void* stream) {
    GNB_REQUIRE(vol && q && out, "gnb_trilinear_sample: null pointer");
    GNB_REQUIRE(!post || (bn_scale && bn_shift), "gnb_trilinear_sample: post needs bn_scale/bn_shift");
    const int64_t rows = (int64_t)B * M;
    if (rows == 0) return GNB_OK;
    trilinear_kernel<false><<<(unsigned)ceil_div<int64_t>(rows, 8), 256, 0, as_stream(stream)>>>(
        vol, 0, D, H, W, C, q, 0, 0, M, rows, flip, post, bn_scale, bn_shift, out, ldo);
    return check_launch("gnb_trilinear_sample");
}

int32_t gnb_trilinear_sample_grid(const float* vol, int32_t b, int32_t D, int32_t H, int32_t W, int32_t C, int32_t Q,
                                  int64_t m0, int64_t M, int32_t post, const float* bn_scale, const float* bn_shift,
                                  float* out, int64_t ldo, void* stream) {
    GNB_REQUIRE(vol && out, "gnb_trilinear_sample_grid: null pointer");
    GNB_REQUIRE(Q >= 2 && m0 >= 0 && m0 + M <= (int64_t)Q * Q * Q, "gnb_trilinear_sample_grid: bad lattice range");
    GNB_REQUIRE(!post || (bn_scale && bn_shift), "gnb_trilinear_sample_grid: post needs bn_scale/bn_shift");
    if (M == 0) return GNB_OK;
    trilinear_kernel<true><<<(unsigned)ceil_div<int64_t>(M, 8), 256, 0, as_stream(stream)>>>(
        vol, b, D, H, W, C, nullptr, Q, m0, M, M, 0, post, bn_scale, bn_shift, out, ldo);
    return check_launch("gnb_trilinear_sample_grid");
}

}  // extern "C"
