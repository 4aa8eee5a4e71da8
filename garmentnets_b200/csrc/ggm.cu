// Gaussian gradient magnitude (ref predict.py:162-163: scipy.ndimage.gaussian_gradient_magnitude(wnf, sigma,
// mode='nearest')).
//
// scipy evaluates, for every axis a, the separable filter  G'(axis a) x G(other axes)  as three 1-D correlations run
// in axis order 0 -> 1 -> 2, each with double accumulation (NI_Correlate1D, symmetric / anti-symmetric pairing,
// farthest pair first) and float32 storage after every pass, then  sqrt((d0^2 + d1^2) + d2^2)  in float32.
//
// ggm_fused_kernel<R> does all of that in ONE pass over the volume: an 8 x 8 x 32 output tile per CTA, the
// (8+2R)(8+2R)(32+2R) input halo read once (neighbouring tiles share it through L2), the two axis-0 and three axis-1
// intermediates kept in shared memory as float32 (the same rounding points as scipy), so HBM sees one read and one
// write of the volume instead of nine.  Edge replication ('nearest') is index clamping of the halo: the intermediate at
// a clamped coordinate is exactly what the next pass of the multi-pass form reads.  HBM-bound by design
// (8 B/voxel algorithmic); the double-precision pairing arithmetic is the actual limiter (~50 DFMA-class ops/voxel).
// filter1d_kernel is the general multi-pass form, kept for radii above GGM_RMAX.
#include "common.cuh"

namespace gnb {

struct Taps { double w[33]; int radius; };

constexpr int GGM_TD = 8, GGM_TH = 8, GGM_TW = 32, GGM_THREADS = 256, GGM_RMAX = 4;

// correlation of the 2R+1 window `c` (already double) with symmetric / anti-symmetric weights, scipy's operation order
template <int R, bool ANTI>
__device__ __forceinline__ float corr1d(const double* c, const double* w) {
    double acc = ANTI ? 0.0 : c[R] * w[R];
#pragma unroll
    for (int k = R; k >= 1; --k) {
        if (ANTI) acc += (c[R - k] - c[R + k]) * w[R - k];
        else acc += (c[R - k] + c[R + k]) * w[R - k];
    }
    return (float)acc;
}

template <int R>
__global__ void __launch_bounds__(GGM_THREADS)
ggm_fused_kernel(const float* __restrict__ in, int D, int H, int W, int tiles_d, int tiles_h, int tiles_w, Taps g0,
                 Taps g1, float* __restrict__ out) {
    constexpr int HH = GGM_TH + 2 * R, WW = GGM_TW + 2 * R, DD = GGM_TD + 2 * R;
    extern __shared__ float sm[];
    float* s0 = sm;                                  // [2][TD][HH][WW]  axis-0 pass: 0 = derivative, 1 = smoothing
    float* s1 = s0 + 2 * GGM_TD * HH * WW;           // [3][TD][TH][WW]  axis-1 pass of gradient component a
    int64_t t = blockIdx.x;
    const int tw = (int)(t % tiles_w); t /= tiles_w;
    const int th = (int)(t % tiles_h); t /= tiles_h;
    const int td = (int)(t % tiles_d);
    const int64_t vol = t / tiles_d;
    const float* v = in + vol * (int64_t)D * H * W;
    const int d0 = td * GGM_TD, h0 = th * GGM_TH, w0 = tw * GGM_TW;

    // ---- axis 0: one (hh, ww) column per thread, DD inputs in registers -> TD outputs of both filters
    for (int col = threadIdx.x; col < HH * WW; col += GGM_THREADS) {
        const int ww = col % WW, hh = col / WW;
        int gh = h0 - R + hh; gh = gh < 0 ? 0 : (gh > H - 1 ? H - 1 : gh);
        int gw = w0 - R + ww; gw = gw < 0 ? 0 : (gw > W - 1 ? W - 1 : gw);
        const float* colp = v + (int64_t)gh * W + gw;
        double c[DD];
#pragma unroll
        for (int dd = 0; dd < DD; ++dd) {
            int gd = d0 - R + dd; gd = gd < 0 ? 0 : (gd > D - 1 ? D - 1 : gd);
            c[dd] = (double)__ldg(colp + (int64_t)gd * H * W);
        }
#pragma unroll
        for (int d = 0; d < GGM_TD; ++d) {
            s0[((0 * GGM_TD + d) * HH + hh) * WW + ww] = corr1d<R, true>(c + d, g1.w);
            s0[((1 * GGM_TD + d) * HH + hh) * WW + ww] = corr1d<R, false>(c + d, g0.w);
        }
    }
    __syncthreads();
    // ---- axis 1: one (d, ww) column per thread and source array
    for (int col = threadIdx.x; col < 2 * GGM_TD * WW; col += GGM_THREADS) {
        const int ww = col % WW, d = (col / WW) % GGM_TD, src = col / (WW * GGM_TD);
        double c[HH];
#pragma unroll
        for (int hh = 0; hh < HH; ++hh) c[hh] = (double)s0[((src * GGM_TD + d) * HH + hh) * WW + ww];
#pragma unroll
        for (int h = 0; h < GGM_TH; ++h) {
            if (src == 0) {
                s1[((0 * GGM_TD + d) * GGM_TH + h) * WW + ww] = corr1d<R, false>(c + h, g0.w);
            } else {
                s1[((1 * GGM_TD + d) * GGM_TH + h) * WW + ww] = corr1d<R, true>(c + h, g1.w);
                s1[((2 * GGM_TD + d) * GGM_TH + h) * WW + ww] = corr1d<R, false>(c + h, g0.w);
            }
        }
    }
    __syncthreads();
    // ---- axis 2 + magnitude: one output voxel per thread and step
    for (int o = threadIdx.x; o < GGM_TD * GGM_TH * GGM_TW; o += GGM_THREADS) {
        const int w = o % GGM_TW, h = (o / GGM_TW) % GGM_TH, d = o / (GGM_TW * GGM_TH);
        const int gd = d0 + d, gh = h0 + h, gw = w0 + w;
        if (gd >= D || gh >= H || gw >= W) continue;
        float comp[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float* row = s1 + ((a * GGM_TD + d) * GGM_TH + h) * WW + w;
            double c[2 * R + 1];
#pragma unroll
            for (int k = 0; k < 2 * R + 1; ++k) c[k] = (double)row[k];
            comp[a] = a == 2 ? corr1d<R, true>(c, g1.w) : corr1d<R, false>(c, g0.w);
        }
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(comp[0], comp[0]), __fmul_rn(comp[1], comp[1])),
                                  __fmul_rn(comp[2], comp[2]));
        out[vol * (int64_t)D * H * W + ((int64_t)gd * H + gh) * W + gw] = __fsqrt_rn(s);
    }
}

// one 1-D correlation pass along `axis` with edge replication ('nearest'), double accumulation like scipy's
// NI_Correlate1D (symmetric / anti-symmetric pairing), float32 storage after every pass.
// mode: 0 store v | 1 store v*v | 2 out += v*v | 3 out = sqrt(out + v*v)
template <bool ANTI>
__global__ void __launch_bounds__(256)
filter1d_kernel(const float* __restrict__ in, int nvol, int D, int H, int W, int axis, Taps tp, int mode,
                float* __restrict__ out) {
    const int64_t total = (int64_t)nvol * D * H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)((i / ((int64_t)W * H)) % D);
    const int len = axis == 0 ? D : (axis == 1 ? H : W);
    const int pos = axis == 0 ? d : (axis == 1 ? h : w);
    const int64_t stride = axis == 0 ? (int64_t)H * W : (axis == 1 ? W : 1);
    const float* line = in + i - (int64_t)pos * stride;
    const int R = tp.radius;
    double acc = ANTI ? 0.0 : (double)line[(int64_t)pos * stride] * tp.w[R];
    for (int k = R; k >= 1; --k) {  // farthest pair first, like NI_Correlate1D
        int lo = pos - k; lo = lo < 0 ? 0 : lo;
        int hi = pos + k; hi = hi > len - 1 ? len - 1 : hi;
        const double a = (double)line[(int64_t)lo * stride], b = (double)line[(int64_t)hi * stride];
        // correlation weights fw[R - k] multiplies in[pos - k]; symmetric: fw[R-k]==fw[R+k]; anti: fw[R-k]==-fw[R+k]
        if (ANTI) acc += (a - b) * tp.w[R - k];
        else acc += (a + b) * tp.w[R - k];
    }
    const float v = (float)acc;
    if (mode == 0) out[i] = v;
    else if (mode == 1) out[i] = __fmul_rn(v, v);
    else if (mode == 2) out[i] = __fadd_rn(out[i], __fmul_rn(v, v));
    else out[i] = __fsqrt_rn(__fadd_rn(out[i], __fmul_rn(v, v)));
}

static void gaussian_taps(double sigma, int order, Taps& tp) {
    // scipy.ndimage._filters._gaussian_kernel1d (order 0 / 1), reversed for correlate1d
    const int R = (int)(4.0 * sigma + 0.5);
    tp.radius = R;
    double sum = 0.0;
    double phi[33];
    for (int x = -R; x <= R; ++x) { phi[x + R] = exp(-0.5 / (sigma * sigma) * (double)x * (double)x); sum += phi[x + R]; }
    for (int x = -R; x <= R; ++x) phi[x + R] /= sum;
    for (int x = -R; x <= R; ++x) {
        double v = phi[x + R];
        if (order == 1) v = ((double)x * (-1.0 / (sigma * sigma))) * v;
        tp.w[R - x] = v;  // [::-1]
    }
}

template <int R>
static int32_t launch_fused(const float* v, int nvol, int D, int H, int W, const Taps& g0, const Taps& g1, float* out,
                            cudaStream_t st) {
    constexpr int HH = GGM_TH + 2 * R, WW = GGM_TW + 2 * R;
    const int smem = (2 * GGM_TD * HH * WW + 3 * GGM_TD * GGM_TH * WW) * (int)sizeof(float);
    GNB_CUDA(cudaFuncSetAttribute(ggm_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int td = ceil_div(D, GGM_TD), th = ceil_div(H, GGM_TH), tw = ceil_div(W, GGM_TW);
    const int64_t tiles = (int64_t)nvol * td * th * tw;
    GNB_REQUIRE(tiles < (1ll << 31), "gnb_gaussian_gradient_magnitude: too many tiles");
    ggm_fused_kernel<R><<<(unsigned)tiles, GGM_THREADS, smem, st>>>(v, D, H, W, td, th, tw, g0, g1, out);
    return check_launch("gnb_gaussian_gradient_magnitude");
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_gaussian_gradient_magnitude_batched(const float* v, int32_t nvol, int32_t D, int32_t H, int32_t W,
                                                double sigma, float* out, float* tmp, void* stream) {
    GNB_REQUIRE(v && out, "gnb_gaussian_gradient_magnitude: null pointer");
    GNB_REQUIRE(sigma > 0 && (int)(4.0 * sigma + 0.5) <= 16, "gnb_gaussian_gradient_magnitude: sigma out of range");
    const int64_t total = (int64_t)nvol * D * H * W;
    if (total == 0) return GNB_OK;
    Taps g0, g1;
    gaussian_taps(sigma, 0, g0);
    gaussian_taps(sigma, 1, g1);
    cudaStream_t st = as_stream(stream);
    switch (g0.radius) {  // sigma 0.5 (the shipped value, predict_default.yaml:45) -> radius 2
        case 1: return launch_fused<1>(v, nvol, D, H, W, g0, g1, out, st);
        case 2: return launch_fused<2>(v, nvol, D, H, W, g0, g1, out, st);
        case 3: return launch_fused<3>(v, nvol, D, H, W, g0, g1, out, st);
        case 4: return launch_fused<4>(v, nvol, D, H, W, g0, g1, out, st);
        default: break;
    }
    GNB_REQUIRE(tmp, "gnb_gaussian_gradient_magnitude: radius > %d needs the tmp workspace", GGM_RMAX);
    const unsigned grid = (unsigned)ceil_div<int64_t>(total, 256);
    float* t1 = tmp;
    float* t2 = tmp + total;
    for (int a = 0; a < 3; ++a) {
        const float* src = v;
        for (int ax = 0; ax < 3; ++ax) {
            const bool last = ax == 2;
            float* dst = last ? out : (ax == 0 ? t1 : t2);
            const int mode = !last ? 0 : (a == 0 ? 1 : (a == 1 ? 2 : 3));
            if (ax == a) filter1d_kernel<true><<<grid, 256, 0, st>>>(src, nvol, D, H, W, ax, g1, mode, dst);
            else filter1d_kernel<false><<<grid, 256, 0, st>>>(src, nvol, D, H, W, ax, g0, mode, dst);
            src = dst;
        }
    }
    return check_launch("gnb_gaussian_gradient_magnitude");
}

int32_t gnb_gaussian_gradient_magnitude(const float* v, int32_t D, int32_t H, int32_t W, double sigma, float* out,
                                        float* tmp, void* stream) {
    return gnb_gaussian_gradient_magnitude_batched(v, 1, D, H, W, sigma, out, tmp, stream);
}

}  // extern "C"
