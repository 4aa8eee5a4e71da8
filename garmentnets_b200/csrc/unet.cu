// 3D-UNet layers on channels-last (NDHWC) activations: GroupNorm statistics, 3x3x3 convolution as an implicit
// GEMM with the GroupNorm affine fused into the operand loader and ReLU into the epilogue, 2x2x2 max-pool,
// nearest-upsample + concat, and the NCDHW <-> NDHWC boundary transpose.  fp32 path (exact fp32 accumulation).
//
// Reference semantics restated: components/unet3d.py:43-72 (SingleConv 'gcr'), :222 (MaxPool3d), :291 (concat,
// encoder features first), :325-330 (nearest upsampling to the skip's size).
#include "common.cuh"

namespace gnb {

// ---- GroupNorm statistics ---------------------------------------------------------------------------
// pass 1: per-CTA partial sums of x and x^2 per (b, group) accumulated in double; pass 2: scale/shift per (b,c).
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, int64_t voxels, int C, int groups, double* __restrict__ ws) {
    // grid: (chunks, B); each CTA reduces a slab of voxels for all channels.
    extern __shared__ double sh[];  // [2*groups]
    const int b = blockIdx.y;
    const int cg = C / groups;
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int64_t per = ceil_div<int64_t>(voxels, gridDim.x);
    const int64_t v0 = (int64_t)blockIdx.x * per;
    const int64_t v1 = v0 + per < voxels ? v0 + per : voxels;
    const float* xb = x + (int64_t)b * voxels * C;
    // thread -> fixed channel (coalesced along C), strided over voxels
    const int64_t total = (v1 > v0 ? (v1 - v0) : 0) * C;
    // each thread keeps the running sums of ONE group at a time: element index e = v*C + c, stride blockDim.
    // blockDim (256) is a multiple of or divides C for every layer (C in {32..384}); handle the general case anyway.
    double s = 0.0, ss = 0.0;
    int cur_g = -1;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int c = (int)(e % C);
        const int g = c / cg;
        if (g != cur_g) {
            if (cur_g >= 0) { atomicAdd(&sh[2 * cur_g], s); atomicAdd(&sh[2 * cur_g + 1], ss); }
            s = 0.0; ss = 0.0; cur_g = g;
        }
        const double v = (double)xb[v0 * C + e];
        s += v; ss += v * v;
    }
    if (cur_g >= 0) { atomicAdd(&sh[2 * cur_g], s); atomicAdd(&sh[2 * cur_g + 1], ss); }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) atomicAdd(&ws[(int64_t)b * groups * 2 + i], sh[i]);
}

// Fast path (C % 4 == 0 and (C/groups) % 4 == 0, every layer of the shipped UNet): each thread owns one fixed
// float4 channel quad -- always inside one group -- and strides over the voxels of its CTA's slab with 16-byte
// coalesced loads; double accumulators; one shared + one global double atomic per thread / per CTA and group.
__global__ void __launch_bounds__(256)
gn_partial_vec4_kernel(const float* __restrict__ x, int64_t voxels, int C, int groups, double* __restrict__ ws,
                       int c_off = 0, int c_total = 0, double mult = 1.0) {
    // c_off / c_total / mult: x is a channel slice [c_off, c_off + C) of a virtual tensor with c_total channels whose statistics
    // are accumulated in ws (gnb_groupnorm_stats_cat); mult = 8 for a nearest-upsampled source (every value appears 8 times)
    if (c_total == 0) c_total = C;
    extern __shared__ double sh[];  // [2*groups]
    const int b = blockIdx.y;
    const int quads = C >> 2;
    const int rows = blockDim.x / quads;            // voxels handled per iteration by this CTA
    const int q = threadIdx.x % quads, r = threadIdx.x / quads;
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int64_t per = ceil_div<int64_t>(voxels, gridDim.x);
    const int64_t v0 = (int64_t)blockIdx.x * per;
    const int64_t v1 = v0 + per < voxels ? v0 + per : voxels;
    if (r < rows) {
        const float4* xb = reinterpret_cast<const float4*>(x + (int64_t)b * voxels * C) + q;
        double s = 0.0, ss = 0.0;
        int64_t v = v0 + r;
        for (; v + 3 * rows < v1; v += 4 * rows) {   // four independent 16-byte loads in flight
            const float4 a0 = __ldg(xb + v * quads), a1 = __ldg(xb + (v + rows) * quads);
            const float4 a2 = __ldg(xb + (v + 2 * rows) * quads), a3 = __ldg(xb + (v + 3 * rows) * quads);
            const float4 aa[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double e0 = aa[k].x, e1 = aa[k].y, e2 = aa[k].z, e3 = aa[k].w;
                s += (e0 + e1) + (e2 + e3);
                ss += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
            }
        }
        for (; v < v1; v += rows) {
            const float4 a = __ldg(xb + v * quads);
            const double e0 = a.x, e1 = a.y, e2 = a.z, e3 = a.w;
            s += (e0 + e1) + (e2 + e3);
            ss += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
        }
        const int g = (c_off + (q << 2)) / (c_total / groups);
        atomicAdd(&sh[2 * g], s * mult);
        atomicAdd(&sh[2 * g + 1], ss * mult);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) atomicAdd(&ws[(int64_t)b * groups * 2 + i], sh[i]);
}

__global__ void gn_finalize_kernel(const double* __restrict__ ws, int B, int64_t voxels, int C, int groups, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ scale, float* __restrict__ shift) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * C) return;
    const int b = t / C, c = t % C;
    const int cg = C / groups, g = c / cg;
    const double n = (double)voxels * cg;
    const double mean = ws[((int64_t)b * groups + g) * 2] / n;
    double var = ws[((int64_t)b * groups + g) * 2 + 1] / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
    scale[t] = (float)(rstd * ga);
    shift[t] = (float)(be - mean * rstd * ga);
}

// ---- 3x3x3 convolution, implicit GEMM ------------------------------------------------------------------
// rows = voxels (b,d,h,w), cols = Cout, K = 27 taps x Cin.  Tile 128 x BN, K chunk 16 channels of one tap.
constexpr int CBM = 128, CBK = 16, CTHREADS = 256;

template <int BN>
__global__ void __launch_bounds__(CTHREADS)
conv3d_k3_kernel(const float* __restrict__ x, int B, int D, int H, int W, int Cin, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ Wt, int Cout, int relu,
                 float* __restrict__ y) {
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[CBK][CBM + 4];
    __shared__ __align__(16) float Bs[CBK][BN + 4];
    const int64_t M = (int64_t)B * D * H * W;
    const int64_t m0 = (int64_t)blockIdx.x * CBM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    // loader rows: thread loads float4 (4 channels) for rows (tid>>2) and (tid>>2)+64
    int rb[2], rd[2], rh[2], rw[2];
    bool rok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t gm = m0 + (tid >> 2) + 64 * i;
        rok[i] = gm < M;
        const int64_t g = rok[i] ? gm : 0;
        rw[i] = (int)(g % W);
        rh[i] = (int)((g / W) % H);
        rd[i] = (int)((g / ((int64_t)W * H)) % D);
        rb[i] = (int)(g / ((int64_t)W * H * D));
    }
    const int kk4 = (tid & 3) * 4;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int chunks_per_tap = ceil_div(Cin, CBK);
    const int nchunks = 27 * chunks_per_tap;
    float4 ra[2];
    float rbv[BN * CBK / CTHREADS];

    auto load_tiles = [&](int ch) {
        const int tap = ch / chunks_per_tap;
        const int c0 = (ch - tap * chunks_per_tap) * CBK;
        const int dz = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
        const int c = c0 + kk4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int zz = rd[i] + dz, yy = rh[i] + dy, xx = rw[i] + dx;
            if (rok[i] && c < Cin && zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const int64_t off = ((((int64_t)rb[i] * D + zz) * H + yy) * W + xx) * Cin + c;
                v = *reinterpret_cast<const float4*>(x + off);
                if (scale != nullptr) {
                    const float4 sc = *reinterpret_cast<const float4*>(scale + (int64_t)rb[i] * Cin + c);
                    const float4 sh = *reinterpret_cast<const float4*>(shift + (int64_t)rb[i] * Cin + c);
                    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                    v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                }
            }
            ra[i] = v;
        }
        // B tile: Wt[tap][c0 + k][n0 + n], k = 0..15, n = 0..BN-1 (contiguous along n)
#pragma unroll
        for (int i = 0; i < BN * CBK / CTHREADS; ++i) {
            const int e = tid + CTHREADS * i;
            const int k = e / BN, n = e % BN;
            rbv[i] = (c0 + k < Cin && n0 + n < Cout) ? Wt[((int64_t)tap * Cin + c0 + k) * Cout + n0 + n] : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int m = (tid >> 2) + 64 * i;
            As[kk4 + 0][m] = ra[i].x; As[kk4 + 1][m] = ra[i].y; As[kk4 + 2][m] = ra[i].z; As[kk4 + 3][m] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < BN * CBK / CTHREADS; ++i) {
            const int e = tid + CTHREADS * i;
            Bs[e / BN][e % BN] = rbv[i];
        }
    };

    load_tiles(0);
    for (int ch = 0; ch < nchunks; ++ch) {
        store_tiles();
        __syncthreads();
        if (ch + 1 < nchunks) load_tiles(ch + 1);
#pragma unroll
        for (int kk = 0; kk < CBK; ++kk) {
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
    for (int i = 0; i < 8; ++i) {
        const int64_t gm = m0 + ty * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= Cout) continue;
            float v = acc[i][j];
            if (relu) v = fmaxf(v, 0.f);
            y[gm * Cout + n] = v;
        }
    }
}

// ---- pooling / upsampling / layout -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool2_kernel(const float* __restrict__ x, int B, int D, int H, int W, int C, float* __restrict__ y) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)B * Do * Ho * Wo * C;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C);
    int64_t v = t / C;
    const int w = (int)(v % Wo); v /= Wo;
    const int h = (int)(v % Ho); v /= Ho;
    const int d = (int)(v % Do);
    const int b = (int)(v / Do);
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int zz = 2 * d + (k >> 2), yy = 2 * h + ((k >> 1) & 1), xx = 2 * w + (k & 1);
        m = fmaxf(m, x[((((int64_t)b * D + zz) * H + yy) * W + xx) * C + c]);
    }
    y[t] = m;
}

__global__ void __launch_bounds__(256)
upsample_concat_kernel(const float* __restrict__ skip, int Cs, const float* __restrict__ x, int Cx, int B, int D,
                       int H, int W, int Dx, int Hx, int Wx, float* __restrict__ y) {
    const int C = Cs + Cx;
    const int64_t total = (int64_t)B * D * H * W * C;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C);
    int64_t v = t / C;
    if (c < Cs) { y[t] = skip[v * Cs + c]; return; }
    const int w = (int)(v % W); v /= W;
    const int h = (int)(v % H); v /= H;
    const int d = (int)(v % D);
    const int b = (int)(v / D);
    // F.interpolate(mode='nearest'): src = floor(dst * in/out) (integer arithmetic for exact ratios)
    const int zs = (int)(((int64_t)d * Dx) / D), ys = (int)(((int64_t)h * Hx) / H), xs = (int)(((int64_t)w * Wx) / W);
    y[t] = x[((((int64_t)b * Dx + zs) * Hx + ys) * Wx + xs) * Cx + (c - Cs)];
}

// Fast path (Cs % 4 == 0 and Cx % 4 == 0, every decoder level of the shipped UNet): one thread per float4 of the output,
// 16-byte coalesced loads and stores, 32-bit index arithmetic per voxel.
__global__ void __launch_bounds__(256)
upsample_concat_vec4_kernel(const float4* __restrict__ skip, int Qs, const float4* __restrict__ x, int Qx, int B, int D,
                            int H, int W, int Dx, int Hx, int Wx, float4* __restrict__ y) {
    const int Q = Qs + Qx;                                  // float4 quads per output voxel
    const int64_t total = (int64_t)B * D * H * W * Q;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int64_t v64 = t / Q;
    const int q = (int)(t - v64 * Q);
    if (q < Qs) { y[t] = __ldg(skip + v64 * Qs + q); return; }
    const int vox = D * H * W;
    const int b = (int)(v64 / vox);
    int v = (int)(v64 - (int64_t)b * vox);
    const int w = v % W; v /= W;
    const int h = v % H;
    const int d = v / H;
    const int zs = (d * Dx) / D, ys = (h * Hx) / H, xs = (w * Wx) / W;   // F.interpolate(mode='nearest')
    y[t] = __ldg(x + ((((int64_t)b * Dx + zs) * Hx + ys) * Wx + xs) * Qx + (q - Qs));
}

// strided NCDHW -> NDHWC through a 32x33 shared-memory transpose tile (voxels x channels)
__global__ void __launch_bounds__(256)
to_channels_last_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sd, int64_t sh, int64_t sw,
                        int C, int D, int H, int W, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int64_t vox = (int64_t)D * H * W;
    const int b = blockIdx.z;
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t v = v0 + tx;
        float val = 0.f;
        if (c < C && v < vox) {
            const int w = (int)(v % W), h = (int)((v / W) % H), d = (int)(v / ((int64_t)W * H));
            val = x[b * sb + c * sc + d * sd + h * sh + w * sw];
        }
        tile[j][tx] = val;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int64_t v = v0 + j;
        const int c = c0 + tx;
        if (c < C && v < vox) y[((int64_t)b * vox + v) * C + c] = tile[tx][j];
    }
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_groupnorm_stats(const float* x, int32_t B, int64_t voxels, int32_t C, int32_t groups, float eps,
                            const float* gamma, const float* beta, float* scale, float* shift, double* ws,
                            void* stream) {
    GNB_REQUIRE(x && scale && shift && ws, "gnb_groupnorm_stats: null pointer");
    GNB_REQUIRE(B > 0 && voxels > 0 && C > 0 && groups > 0 && C % groups == 0, "gnb_groupnorm_stats: bad shape");
    cudaStream_t st = as_stream(stream);
    GNB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)B * groups * 2, st));
    int chunks = (int)ceil_div<int64_t>(voxels * C, 256 * 64);
    const int target = ceil_div(8 * sm_count(), B);
    chunks = chunks < 1 ? 1 : (chunks > target ? target : chunks);
    const bool vec4 = C % 4 == 0 && (C / groups) % 4 == 0 && C / 4 <= 256 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (vec4) gn_partial_vec4_kernel<<<dim3(chunks, B), 256, 2 * groups * sizeof(double), st>>>(x, voxels, C, groups, ws);
    else gn_partial_kernel<<<dim3(chunks, B), 256, 2 * groups * sizeof(double), st>>>(x, voxels, C, groups, ws);
    gn_finalize_kernel<<<ceil_div(B * C, 256), 256, 0, st>>>(ws, B, voxels, C, groups, eps, gamma, beta, scale, shift);
    return check_launch("gnb_groupnorm_stats");
}

int32_t gnb_groupnorm_stats_cat(const float* skip, int32_t Cs, const float* x_low, int32_t Cx, int32_t B, int64_t voxels,
                                int32_t groups, float eps, const float* gamma, const float* beta, float* scale, float* shift,
                                double* ws, void* stream) {
    GNB_REQUIRE(skip && x_low && scale && shift && ws, "gnb_groupnorm_stats_cat: null pointer");
    const int C = Cs + Cx;
    GNB_REQUIRE(B > 0 && voxels > 0 && voxels % 8 == 0 && groups > 0 && C % groups == 0, "gnb_groupnorm_stats_cat: bad shape");
    GNB_REQUIRE(Cs % 4 == 0 && Cx % 4 == 0 && (C / groups) % 4 == 0 && Cs / 4 <= 256 && Cx / 4 <= 256 &&
                ((reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(x_low)) & 15) == 0,
                "gnb_groupnorm_stats_cat: channel counts must be multiples of 4 (groups of whole quads), 16-byte aligned sources");
    cudaStream_t st = as_stream(stream);
    GNB_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)B * groups * 2, st));
    const int target = ceil_div(8 * sm_count(), B);
    auto chunks_for = [&](int64_t vox, int c) {
        int ch = (int)ceil_div<int64_t>(vox * c, 256 * 64);
        return ch < 1 ? 1 : (ch > target ? target : ch);
    };
    // the statistics of cat(skip, upsample(x_low)) are the sums over skip plus 8 x the sums over x_low
    gn_partial_vec4_kernel<<<dim3(chunks_for(voxels, Cs), B), 256, 2 * groups * sizeof(double), st>>>(skip, voxels, Cs, groups, ws, 0, C, 1.0);
    gn_partial_vec4_kernel<<<dim3(chunks_for(voxels / 8, Cx), B), 256, 2 * groups * sizeof(double), st>>>(x_low, voxels / 8, Cx, groups, ws, Cs, C, 8.0);
    gn_finalize_kernel<<<ceil_div(B * C, 256), 256, 0, st>>>(ws, B, voxels, C, groups, eps, gamma, beta, scale, shift);
    return check_launch("gnb_groupnorm_stats_cat");
}

int32_t gnb_conv3d_k3(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin, const float* scale,
                      const float* shift, const float* Wt, int32_t Cout, int32_t relu, float* y, void* stream) {
    GNB_REQUIRE(x && Wt && y, "gnb_conv3d_k3: null pointer");
    GNB_REQUIRE((scale == nullptr) == (shift == nullptr), "gnb_conv3d_k3: scale/shift must come together");
    GNB_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "gnb_conv3d_k3: bad shape");
    GNB_REQUIRE(Cin % 4 == 0, "gnb_conv3d_k3: Cin must be a multiple of 4 (got %d)", Cin);
    const int64_t M = (int64_t)B * D * H * W;
    cudaStream_t st = as_stream(stream);
    if (Cout <= 32) {
        dim3 grid((unsigned)ceil_div<int64_t>(M, CBM), 1);
        conv3d_k3_kernel<32><<<grid, CTHREADS, 0, st>>>(x, B, D, H, W, Cin, scale, shift, Wt, Cout, relu, y);
    } else {
        dim3 grid((unsigned)ceil_div<int64_t>(M, CBM), (unsigned)ceil_div(Cout, 64));
        conv3d_k3_kernel<64><<<grid, CTHREADS, 0, st>>>(x, B, D, H, W, Cin, scale, shift, Wt, Cout, relu, y);
    }
    return check_launch("gnb_conv3d_k3");
}

int32_t gnb_maxpool3d_2(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, float* y,
                        void* stream) {
    GNB_REQUIRE(x && y, "gnb_maxpool3d_2: null pointer");
    GNB_REQUIRE(D >= 2 && H >= 2 && W >= 2, "gnb_maxpool3d_2: spatial size < 2");
    const int64_t total = (int64_t)B * (D / 2) * (H / 2) * (W / 2) * C;
    maxpool2_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(x, B, D, H, W, C, y);
    return check_launch("gnb_maxpool3d_2");
}

int32_t gnb_upsample_concat(const float* skip, int32_t Cs, const float* x, int32_t Cx, int32_t B, int32_t D,
                            int32_t H, int32_t W, int32_t Dx, int32_t Hx, int32_t Wx, float* y, void* stream) {
    GNB_REQUIRE(x && y && (Cs == 0 || skip), "gnb_upsample_concat: null pointer");
    const int64_t total = (int64_t)B * D * H * W * (Cs + Cx);
    if (total == 0) return GNB_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (Cs % 4 == 0 && Cx % 4 == 0 && aligned && (int64_t)D * H * W < (1ll << 30)) {
        upsample_concat_vec4_kernel<<<(unsigned)ceil_div<int64_t>(total / 4, 256), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(skip), Cs / 4, reinterpret_cast<const float4*>(x), Cx / 4, B, D, H, W, Dx, Hx, Wx,
            reinterpret_cast<float4*>(y));
        return check_launch("gnb_upsample_concat");
    }
    upsample_concat_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        skip, Cs, x, Cx, B, D, H, W, Dx, Hx, Wx, y);
    return check_launch("gnb_upsample_concat");
}

int32_t gnb_to_channels_last(const float* x, int64_t sb, int64_t sc, int64_t sd, int64_t sh, int64_t sw, int32_t B,
                             int32_t C, int32_t D, int32_t H, int32_t W, float* y, void* stream) {
    GNB_REQUIRE(x && y, "gnb_to_channels_last: null pointer");
    const int64_t vox = (int64_t)D * H * W;
    if (vox == 0 || B == 0 || C == 0) return GNB_OK;
    dim3 grid((unsigned)ceil_div<int64_t>(vox, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
    to_channels_last_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, sb, sc, sd, sh, sw, C, D, H, W, y);
    return check_launch("gnb_to_channels_last");
}

}  // extern "C"
