// PointConv message MLP + max aggregation in ONE kernel (ref components/pointnet2.py:30-31: PointConv(local_nn) with
// local_nn = MLP([Cin + 3, C1, C2, C3]) of Linear -> ReLU -> BatchNorm blocks (components/mlp.py:9-20), aggr = 'max').
//
// The unfused chain (gnb_pointconv_gather, three gnb_linear_tc launches, gnb_segment_max) already runs every pass at the HBM
// roofline -- and moves 11 GB per step for SA1 + SA2, because every [E, C] activation is written and read back.  Here a tile of
// 128 edges goes through the three layers without leaving the SM:
//
//   thread = edge row (tensor-memory lane), two threads per row split the channels (256 threads = 8 warps);
//   layer 1   Cin >= 16 (SA2): x_j is gathered straight into the A operand in TENSOR MEMORY (fp16 hi | lo columns, tcgen05.st),
//             one thread issues the MMAs against the resident W1x image; the three relative-position inputs enter in the fp32
//             epilogue (3 FMAs per channel) so that K stays a multiple of 64.   Cin < 16 (SA1, 6 inputs): fp32 FMAs, no MMA.
//   layer 2/3 accumulator -> registers (tcgen05.ld), bias + ReLU, fp16 hi/lo split riding on the conversions, written back to
//             tensor memory as the next A operand (BatchNorm1/2 follow a ReLU, so they are folded into W2 / W3 on the host);
//   output    bias + ReLU + BatchNorm3 per column, 64 columns at a time through a [column][row] staging tile in shared memory;
//             then one thread per (column, segment of equal targets) takes the maximum over the segment's rows (edges are sorted
//             by target, so segments are row ranges) and merges it into out[target] with an integer atomicMax on the order-
//             preserving encoding (segments straddle tiles) -- 128-byte coalesced reductions.  gnb_pointconv_mlp_max decodes the
//             buffer afterwards.  (A per-column warp reduction with a per-segment member mask compiles to a divergence-handling
//             loop of ~35 instructions per column and took two thirds of the kernel.)
//
// Stages of one tile are sequential inside a CTA (no warp specialisation); SA1 needs 256 TMEM columns and 48 KB of weights, so
// two CTAs share an SM and overlap each other's stages; SA2 needs all 512 columns: W1x (64 KB) stays resident, W2 (64 KB) and
// W3 (128 KB) alternate in one 128 KB region, each reload hidden behind the epilogue that follows the MMAs that last read it.
// Precision as in the decoders: fp16 hi + lo operands, hi*hi + lo*hi + hi*lo, fp32 accumulation.
#include "tc_common.cuh"

namespace gnb {
namespace sam {

constexpr int M = 128, THREADS = 256;
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// epilogue constants (copied to shared memory by every CTA, read as float4 broadcasts):
// [b1 C1 | w1 (KIN x C1, input-major: SA1 all 6 inputs, SA2 the 3 position inputs) | b2 C2 | b3 C3 | s3 C3 | t3 C3]
constexpr int STG_COLS = 64, STG_LD = M + 1;     // output staging tile: [64 columns][129] floats

struct Params {
    const float* x; int64_t ldx;      // [Nx, Cin] source features (may be null when Cin == 0)
    const float* pos_x;               // [Nx, 3]
    const float* pos_y;               // [My, 3]
    const int32_t* src;               // [E] source point of every edge
    const int32_t* dst;               // [E] target (centre) of every edge, ascending
    const int64_t* e_total;           // device scalar: E
    const uint8_t* w1_img;            // SA2: W1x image  [K/64][hi,lo][C1 rows x 128 B]
    const uint8_t* w2_img;            // [C1/64][hi,lo][C2 rows x 128 B]
    const uint8_t* w3_img;            // [C2/64][hi,lo][C3 rows x 128 B]
    float inv1, inv2, inv3;           // 2^-s of the three packed weights
    const float* consts;              // epilogue constants, see above
    uint32_t* out_enc;                // [My, C3] order-preserving encodings, zero-initialised
};

__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// ReLU + fp16 hi/lo split with the ReLU riding on the conversions (see decode_lattice.cu)
__device__ __forceinline__ void relu_split2(float2 h, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(h.y), "f"(h.x));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 r = sub2(h, hf);
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
__device__ __forceinline__ void split2(float2 h, uint32_t& hi, uint32_t& lo) {
    hi = cvt_f16x2_sat(h.x, h.y);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 r = sub2(h, hf);
    lo = cvt_f16x2_sat(r.x, r.y);
}
__device__ __forceinline__ uint32_t enc_f32(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int CIN, int C1, int C2, int C3>
struct Cfg {
    static constexpr bool MMA1 = CIN >= 16;
    static constexpr int KIN = MMA1 ? 3 : CIN + 3;                       // inputs applied with fp32 FMAs in the layer-1 epilogue
    static constexpr int O_B1 = 0, O_W1 = C1, O_B2 = C1 + KIN * C1, O_B3 = O_B2 + C2, O_S3 = O_B3 + C3, O_T3 = O_S3 + C3, N_CONST = O_T3 + C3;
    static constexpr int A_COLS = cmax(MMA1 ? CIN : 0, cmax(C1, C2));    // A operand of K channels: K/2 hi + K/2 lo columns
    static constexpr int ACC_COLS = cmax(MMA1 ? C1 : 0, C2);
    static constexpr int ACC12 = A_COLS, ACC3 = A_COLS + ACC_COLS;
    static constexpr int TMEM_USED = ACC3 + C3;
    static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;
    static constexpr int W1_BYTES = MMA1 ? C1 * CIN * 4 : 0, W2_BYTES = C2 * C1 * 4, W3_BYTES = C3 * C2 * 4;
    static constexpr bool RESIDENT = W1_BYTES + W2_BYTES + W3_BYTES <= 96 * 1024;
    // shared memory: resident -> [W2 | W3]; streamed -> [W1 resident | region RA = max(W2, W3): W2 and W3 alternate]
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = W1_BYTES;
    static constexpr int OFF_W3 = RESIDENT ? W1_BYTES + W2_BYTES : W1_BYTES;
    static constexpr int W_TOTAL = RESIDENT ? W1_BYTES + W2_BYTES + W3_BYTES : W1_BYTES + cmax(W2_BYTES, W3_BYTES);
    static constexpr int OFF_BARS = W_TOTAL;
    static constexpr int OFF_CONST = OFF_BARS + 64;
    static constexpr int OFF_SEG = OFF_CONST + N_CONST * 4;              // seg_row0[130] | seg_dst[130] | warp counts [4]
    static constexpr int STG_BYTES = STG_COLS * STG_LD * 4;
    // streamed weights: the staging tile lives in the upper half of region RA (free between the MMAs of layer 3 and the next W3)
    static constexpr int OFF_STG = RESIDENT ? ((OFF_SEG + 264 * 4 + 15) & ~15) : W1_BYTES + W2_BYTES;
    static constexpr int SMEM = RESIDENT ? OFF_STG + STG_BYTES : OFF_SEG + 264 * 4;
    static_assert(TMEM_USED <= 512 && SMEM + 1024 <= 227 * 1024, "budget");
    static_assert(RESIDENT || W2_BYTES + STG_BYTES <= W3_BYTES, "staging tile must fit beside W2 in region RA");
    static_assert(C1 % 64 == 0 && C2 % 64 == 0 && C3 % 64 == 0 && (!MMA1 || CIN % 64 == 0), "channel counts must be multiples of 64");
};

constexpr uint32_t idesc_n(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// MMAs of one layer: acc[128 x N] = A[128 x K] (TMEM: hi columns a_col.., lo columns a_col + K/2..) x W^T (image of K/64 x {hi, lo}
// pieces of N rows x 128 B).  The tensor core truncates its fp32 accumulator on every instruction (~0.5 ulp of the ACCUMULATOR,
// whatever the size of the addend); the small cross terms lo*hi and hi*lo therefore go FIRST, while the accumulator is still
// 2^-11 of its final size and its ulp negligible, and only the K/16 hi*hi instructions truncate at full scale -- the effect of
// gnb_linear_tc's separate cross-term accumulator without its TMEM columns (sharing one accumulator in the order hi*hi, lo*hi,
// hi*lo per K-step doubled the end-to-end error of the PointNet++ chain).
template <int K, int N>
__device__ __forceinline__ void issue_layer(uint32_t tmem_base, uint32_t a_col, uint32_t acc_col, uint32_t w_smem) {
    constexpr uint32_t PIECE = (uint32_t)N * 128u;
    const uint32_t acc = tmem_base + acc_col, ahi = tmem_base + a_col, alo = ahi + (uint32_t)(K / 2);
#pragma unroll
    for (int kc = 0; kc < K / 64; ++kc) {
        const uint32_t whi = w_smem + (uint32_t)(2 * kc) * PIECE, wlo = whi + PIECE;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ks = (uint32_t)(kc * 4 + kk);
            umma_f16_ta(acc, alo + ks * 8u, umma_desc(whi + kk * 32), idesc_n(N), ks != 0);
            umma_f16_ta(acc, ahi + ks * 8u, umma_desc(wlo + kk * 32), idesc_n(N), 1);
        }
    }
#pragma unroll
    for (int kc = 0; kc < K / 64; ++kc) {
        const uint32_t whi = w_smem + (uint32_t)(2 * kc) * PIECE;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16_ta(acc, ahi + (uint32_t)(kc * 4 + kk) * 8u, umma_desc(whi + kk * 32), idesc_n(N), 1);
    }
}

template <int CIN, int C1, int C2, int C3>
__global__ void __launch_bounds__(THREADS, (Cfg<CIN, C1, C2, C3>::TMEM_COLS == 256 ? 2 : 1))
sa_mlp_kernel(const Params p) {
    using C = Cfg<CIN, C1, C2, C3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qd = warp & 3, hsel = warp >> 2;          // TMEM lane quarter, channel half
    const int row = qd * 32 + lane;
    const uint32_t bar_w1 = sbase + C::OFF_BARS, bar_w2 = bar_w1 + 8, bar_w3 = bar_w1 + 16, bar_mma = bar_w1 + 24;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BARS + 32);
    const float* cst = reinterpret_cast<const float*>(smem + C::OFF_CONST);
    int* seg_row0 = reinterpret_cast<int*>(smem + C::OFF_SEG);          // [nseg + 1] first row of every segment of the tile
    int* seg_dst = seg_row0 + 130;                                       // [nseg] its target
    int* seg_cnt = seg_dst + 130;                                        // [4] segment starts per warp
    float* stg = reinterpret_cast<float*>(smem + C::OFF_STG);            // [64 columns][129]
    const bool issuer = threadIdx.x == 0;

    if (issuer) {
        mbar_init(bar_w1, 1); mbar_init(bar_w2, 1); mbar_init(bar_w3, 1); mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = threadIdx.x; t < C::N_CONST; t += THREADS) reinterpret_cast<float*>(smem + C::OFF_CONST)[t] = p.consts[t];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::OFF_BARS + 32), "r"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);   // this thread's row of the tensor memory

    const int64_t E = *p.e_total;
    const int64_t num_tiles = (E + M - 1) / M;
    if (issuer && (int64_t)blockIdx.x < num_tiles) {
        // weights that stay resident (all of them, or W1x) and the first W2
        if (C::RESIDENT) {
            mbar_expect_tx(bar_w1, C::W1_BYTES + C::W2_BYTES + C::W3_BYTES);
            if (C::MMA1) bulk_g2s(sbase + C::OFF_W1, p.w1_img, C::W1_BYTES, bar_w1);
            bulk_g2s(sbase + C::OFF_W2, p.w2_img, C::W2_BYTES, bar_w1);
            bulk_g2s(sbase + C::OFF_W3, p.w3_img, C::W3_BYTES, bar_w1);
        } else {
            mbar_expect_tx(bar_w1, C::W1_BYTES);
            bulk_g2s(sbase + C::OFF_W1, p.w1_img, C::W1_BYTES, bar_w1);
            mbar_expect_tx(bar_w2, C::W2_BYTES);
            bulk_g2s(sbase + C::OFF_W2, p.w2_img, C::W2_BYTES, bar_w2);
        }
    }
    uint32_t ph_mma = 0, ph_w = 0;   // phase bits: MMA completion (all threads), streamed weights (issuer)
    bool first = true;
    // this thread's channels of a layer of width CW: [hsel * CW/2, +CW/2); constants of channel c of a block at offset O: cst[O + c]
    auto c4 = [&](int off) { return *reinterpret_cast<const float4*>(cst + off); };

    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e = tile * M + row;
        const bool valid = e < E;
        const int j = valid ? p.src[e] : 0;
        const int i = valid ? p.dst[e] : -1;
        float d[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k) d[k] = __fsub_rn(__ldg(p.pos_x + (int64_t)j * 3 + k), __ldg(p.pos_y + (int64_t)i * 3 + k));
        }
        // segments of equal targets inside the tile (rows are sorted by target): start rows, compacted in row order
        unsigned start_mask = 0u;
        if (hsel == 0) {
            const bool start = valid && (row == 0 || p.dst[e - 1] != i);
            start_mask = __ballot_sync(0xffffffffu, start);
            if (lane == 0) seg_cnt[qd] = __popc(start_mask);
        }
        // ------------------------------------------------ layer 1 ------------------------------------------------
        if (C::MMA1) {
            // x_j -> A operand (this thread: channels [hsel * CIN/2, +CIN/2) of its row), 32 channels per round
            const float* xr = p.x + (int64_t)j * p.ldx + hsel * (CIN / 2);
#pragma unroll
            for (int r32 = 0; r32 < CIN / 64; ++r32) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int v4 = 0; v4 < 8; ++v4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) v = __ldg(reinterpret_cast<const float4*>(xr + r32 * 32 + v4 * 4));
                    split2(make_float2(v.x, v.y), hi[2 * v4], lo[2 * v4]);
                    split2(make_float2(v.z, v.w), hi[2 * v4 + 1], lo[2 * v4 + 1]);
                }
                const uint32_t col = (uint32_t)(hsel * (CIN / 4) + r32 * 16);
                tmem_st16(lane_addr + col, hi);
                tmem_st16(lane_addr + (uint32_t)(CIN / 2) + col, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncthreads();
            if (issuer) {
                if (first) mbar_wait(bar_w1, 0);
                tc_fence_after();
                issue_layer<CIN, C1>(tmem_base, 0u, (uint32_t)C::ACC12, sbase + C::OFF_W1);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
            tc_fence_after();
            // epilogue 1: h = relu(acc * inv1 + b1 + W1p d) -> A2 (channels [hsel * C1/2, +C1/2))
#pragma unroll
            for (int r32 = 0; r32 < C1 / 64; ++r32) {
                const int cb = hsel * (C1 / 2) + r32 * 32;   // first channel of this round
                uint32_t acc[32];
                tmem_ld32(lane_addr + (uint32_t)C::ACC12 + (uint32_t)cb, acc);
                tmem_ld_wait();
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 b = c4(C::O_B1 + cb + 4 * g), wx = c4(C::O_W1 + cb + 4 * g), wy = c4(C::O_W1 + C1 + cb + 4 * g),
                                 wz = c4(C::O_W1 + 2 * C1 + cb + 4 * g);
                    const float h0 = fmaf(wz.x, d[2], fmaf(wy.x, d[1], fmaf(wx.x, d[0], fmaf(__uint_as_float(acc[4 * g]), p.inv1, b.x))));
                    const float h1 = fmaf(wz.y, d[2], fmaf(wy.y, d[1], fmaf(wx.y, d[0], fmaf(__uint_as_float(acc[4 * g + 1]), p.inv1, b.y))));
                    const float h2 = fmaf(wz.z, d[2], fmaf(wy.z, d[1], fmaf(wx.z, d[0], fmaf(__uint_as_float(acc[4 * g + 2]), p.inv1, b.z))));
                    const float h3 = fmaf(wz.w, d[2], fmaf(wy.w, d[1], fmaf(wx.w, d[0], fmaf(__uint_as_float(acc[4 * g + 3]), p.inv1, b.w))));
                    relu_split2(make_float2(h0, h1), hi[2 * g], lo[2 * g]);
                    relu_split2(make_float2(h2, h3), hi[2 * g + 1], lo[2 * g + 1]);
                }
                // acc1 / A1 are dead (the MMAs completed, this thread's accumulator columns are in registers): A2 takes the A region
                const uint32_t col = (uint32_t)(hsel * (C1 / 4) + r32 * 16);
                tmem_st16(lane_addr + col, hi);
                tmem_st16(lane_addr + (uint32_t)(C1 / 2) + col, lo);
            }
        } else {
            // fp32 layer 1: h = relu(b1 + W1 [x_j, d]) for channels [hsel * C1/2, +C1/2)
            float in[C::KIN];
#pragma unroll
            for (int k = 0; k < CIN; ++k) in[k] = valid ? __ldg(p.x + (int64_t)j * p.ldx + k) : 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) in[CIN + k] = d[k];
#pragma unroll
            for (int r32 = 0; r32 < C1 / 64; ++r32) {
                const int cb = hsel * (C1 / 2) + r32 * 32;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float4 h = c4(C::O_B1 + cb + 4 * g);
#pragma unroll
                    for (int k = 0; k < C::KIN; ++k) {
                        const float4 w = c4(C::O_W1 + k * C1 + cb + 4 * g);
                        h.x = fmaf(w.x, in[k], h.x); h.y = fmaf(w.y, in[k], h.y); h.z = fmaf(w.z, in[k], h.z); h.w = fmaf(w.w, in[k], h.w);
                    }
                    relu_split2(make_float2(h.x, h.y), hi[2 * g], lo[2 * g]);
                    relu_split2(make_float2(h.z, h.w), hi[2 * g + 1], lo[2 * g + 1]);
                }
                const uint32_t col = (uint32_t)(hsel * (C1 / 4) + r32 * 16);
                tmem_st16(lane_addr + col, hi);
                tmem_st16(lane_addr + (uint32_t)(C1 / 2) + col, lo);
            }
        }
        // ------------------------------------------------ layer 2 ------------------------------------------------
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();          // also publishes seg_cnt
        if (issuer) {
            if (C::RESIDENT) { if (first) mbar_wait(bar_w1, 0); }
            else mbar_wait(bar_w2, ph_w);
            tc_fence_after();
            issue_layer<C1, C2>(tmem_base, 0u, (uint32_t)C::ACC12, sbase + C::OFF_W2);
            umma_commit(bar_mma);
        }
        int nseg;
        {   // segment table (while the tensor core works on layer 2)
            const int c0 = seg_cnt[0], c1 = seg_cnt[1], c2 = seg_cnt[2], c3 = seg_cnt[3];
            nseg = c0 + c1 + c2 + c3;
            if (hsel == 0) {
                const int base = qd == 0 ? 0 : (qd == 1 ? c0 : (qd == 2 ? c0 + c1 : c0 + c1 + c2));
                if ((start_mask >> lane) & 1u) {
                    const int idx = base + __popc(start_mask & ((1u << lane) - 1u));
                    seg_row0[idx] = row;
                    seg_dst[idx] = i;
                }
                if (threadIdx.x == 0) {
                    const int64_t left = E - tile * M;
                    seg_row0[nseg] = left < M ? (int)left : M;
                }
            }
        }
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
        tc_fence_after();
        if (!C::RESIDENT && issuer) {   // W2 has been consumed: W3 takes its place (lands while epilogue 2 runs)
            fence_proxy_async();        // the staging tile of the previous output stage (generic proxy) lies inside this region
            mbar_expect_tx(bar_w3, C::W3_BYTES);
            bulk_g2s(sbase + C::OFF_W3, p.w3_img, C::W3_BYTES, bar_w3);
        }
#pragma unroll
        for (int r32 = 0; r32 < C2 / 64; ++r32) {
            const int cb = hsel * (C2 / 2) + r32 * 32;
            uint32_t acc[32];
            tmem_ld32(lane_addr + (uint32_t)C::ACC12 + (uint32_t)cb, acc);
            tmem_ld_wait();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 b = c4(C::O_B2 + cb + 4 * g);
                relu_split2(make_float2(fmaf(__uint_as_float(acc[4 * g]), p.inv2, b.x), fmaf(__uint_as_float(acc[4 * g + 1]), p.inv2, b.y)), hi[2 * g], lo[2 * g]);
                relu_split2(make_float2(fmaf(__uint_as_float(acc[4 * g + 2]), p.inv2, b.z), fmaf(__uint_as_float(acc[4 * g + 3]), p.inv2, b.w)), hi[2 * g + 1], lo[2 * g + 1]);
            }
            const uint32_t col = (uint32_t)(hsel * (C2 / 4) + r32 * 16);
            tmem_st16(lane_addr + col, hi);
            tmem_st16(lane_addr + (uint32_t)(C2 / 2) + col, lo);
        }
        // ------------------------------------------------ layer 3 ------------------------------------------------
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();          // also publishes the segment table
        if (issuer) {
            if (!C::RESIDENT) mbar_wait(bar_w3, ph_w);
            tc_fence_after();
            issue_layer<C2, C3>(tmem_base, 0u, (uint32_t)C::ACC3, sbase + C::OFF_W3);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma); ph_mma ^= 1u;
        tc_fence_after();
        if (!C::RESIDENT && issuer) {   // W3 has been consumed: the next tile's W2 (lands while the output stage runs; the
            ph_w ^= 1u;                 // staging tile sits in the part of region RA that W2 does not cover)
            if (tile + gridDim.x < num_tiles) {
                mbar_expect_tx(bar_w2, C::W2_BYTES);
                bulk_g2s(sbase + C::OFF_W2, p.w2_img, C::W2_BYTES, bar_w2);
            }
        }
        first = false;
        // ------------------------------------------------ output ------------------------------------------------
        // v = relu(acc * inv3 + b3) * s3 + t3, 64 columns per round through the [column][row] staging tile, then per (column,
        // segment) the maximum over the segment's rows, merged into out[target] with an integer atomicMax
#pragma unroll 1
        for (int rnd = 0; rnd < C3 / STG_COLS; ++rnd) {
            const int cb = rnd * STG_COLS + hsel * 32;       // this thread's 32 columns of the round
            uint32_t acc[32];
            tmem_ld32(lane_addr + (uint32_t)C::ACC3 + (uint32_t)cb, acc);
            tmem_ld_wait();
            float* srow = stg + (hsel * 32) * STG_LD + row;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 b = c4(C::O_B3 + cb + 4 * g), sc = c4(C::O_S3 + cb + 4 * g), sh = c4(C::O_T3 + cb + 4 * g);
                srow[(4 * g) * STG_LD] = fmaf(fmaxf(fmaf(__uint_as_float(acc[4 * g]), p.inv3, b.x), 0.f), sc.x, sh.x);
                srow[(4 * g + 1) * STG_LD] = fmaf(fmaxf(fmaf(__uint_as_float(acc[4 * g + 1]), p.inv3, b.y), 0.f), sc.y, sh.y);
                srow[(4 * g + 2) * STG_LD] = fmaf(fmaxf(fmaf(__uint_as_float(acc[4 * g + 2]), p.inv3, b.z), 0.f), sc.z, sh.z);
                srow[(4 * g + 3) * STG_LD] = fmaf(fmaxf(fmaf(__uint_as_float(acc[4 * g + 3]), p.inv3, b.w), 0.f), sc.w, sh.w);
            }
            __syncthreads();
            {
                const int c = threadIdx.x & (STG_COLS - 1);
                const float* scol = stg + c * STG_LD;
                for (int sgi = threadIdx.x / STG_COLS; sgi < nseg; sgi += THREADS / STG_COLS) {
                    const int r0 = seg_row0[sgi], r1 = seg_row0[sgi + 1];
                    float m = scol[r0];
                    for (int r = r0 + 1; r < r1; ++r) m = fmaxf(m, scol[r]);
                    atomicMax(p.out_enc + (int64_t)seg_dst[sgi] * C3 + rnd * STG_COLS + c, enc_f32(m));
                }
            }
            __syncthreads();   // the staging tile is rewritten by the next round
        }
        // the accumulators and the A region are rewritten by the next tile only after every thread has left this stage
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
    }
}

// edge list of the grouped ball query: hits minus j == i, plus the self loop (flat point index i) appended last -- the rows
// gnb_pointconv_gather materialises, as (source, target) index pairs.  One warp per centroid.
__global__ void __launch_bounds__(256)
edge_index_kernel(const int64_t* __restrict__ nbr, const int32_t* __restrict__ cnt, const int64_t* __restrict__ eoffs, int64_t sumM, int K,
                  int32_t* __restrict__ src, int32_t* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= sumM) return;
    const int c = cnt[i];
    const int64_t e0 = eoffs[i];
    const int ne = (int)(eoffs[i + 1] - e0);
    int base_w = 0;
    for (int t0 = 0; t0 <= c; t0 += 32) {
        const int t = t0 + lane;
        int64_t j = i;                       // t == c: the self loop added by PointConv
        if (t < c) j = nbr[i * K + t];
        const bool keep = t <= c && !(t < c && j == i);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int w = base_w + __popc(mask & ((1u << lane) - 1u));
        if (keep && w < ne) { src[e0 + w] = (int32_t)j; dst[e0 + w] = (int32_t)i; }
        base_w += __popc(mask);
    }
}

__global__ void __launch_bounds__(256)
decode_max_kernel(uint32_t* __restrict__ buf, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t e = buf[t];
    // 0 = no edge reached this entry (cannot happen for PointConv targets: the self loop always exists) -> 0.0f like scatter-max's fill
    buf[t] = e == 0u ? 0u : ((e & 0x80000000u) ? (e ^ 0x80000000u) : ~e);
}

// W fp32 [N][ldw] (columns k0 .. k0+K-1) * 2^s -> [K/64][hi, lo][N rows x 128 B] fp16 K-major SWIZZLE_128B
__global__ void __launch_bounds__(256)
pack_pieces_kernel(const float* __restrict__ W, int N, int K, int ldw, int k0, float wscale, uint8_t* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * K) return;
    const int n = t / K, k = t % K;
    const float w = fminf(fmaxf(W[(int64_t)n * ldw + k0 + k] * wscale, -65504.f), 65504.f);
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const int kc = k / 64;
    const size_t piece = (size_t)N * 128;
    const uint32_t off = sw128_offset(n, k % 64);
    *reinterpret_cast<__half*>(out + (size_t)(2 * kc) * piece + off) = h;
    *reinterpret_cast<__half*>(out + (size_t)(2 * kc + 1) * piece + off) = l;
}

template <int CIN, int C1, int C2, int C3>
static int32_t launch(const Params& p, cudaStream_t st) {
    using C = Cfg<CIN, C1, C2, C3>;
    const int smem = C::SMEM + 1024;
    GNB_CUDA(cudaFuncSetAttribute(sa_mlp_kernel<CIN, C1, C2, C3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = sm_count() * (C::TMEM_COLS == 256 ? 2 : 1);
    sa_mlp_kernel<CIN, C1, C2, C3><<<grid, THREADS, smem, st>>>(p);
    return check_launch("gnb_pointconv_mlp_max");
}

}  // namespace sam
}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_pointconv_mlp_supported(int32_t Cin, int32_t C1, int32_t C2, int32_t C3) {
    return (Cin == 3 && C1 == 64 && C2 == 64 && C3 == 128) || (Cin == 128 && C1 == 128 && C2 == 128 && C3 == 256) ? 1 : 0;
}

int64_t gnb_pointconv_mlp_packed_bytes(int32_t Cin, int32_t C1, int32_t C2, int32_t C3) {
    const int64_t k1 = Cin >= 16 ? Cin : 0;
    return 4 * ((int64_t)C1 * k1 + (int64_t)C2 * C1 + (int64_t)C3 * C2);
}

int32_t gnb_pointconv_mlp_pack(const float* W1, const float* W2, const float* W3, int32_t Cin, int32_t C1, int32_t C2, int32_t C3,
                               int32_t s1, int32_t s2, int32_t s3, void* packed, void* stream) {
    GNB_REQUIRE(W1 && W2 && W3 && packed, "gnb_pointconv_mlp_pack: null pointer");
    GNB_REQUIRE(gnb_pointconv_mlp_supported(Cin, C1, C2, C3), "gnb_pointconv_mlp_pack: unsupported layer widths %d+3 -> %d -> %d -> %d", Cin, C1, C2, C3);
    cudaStream_t st = as_stream(stream);
    uint8_t* out = reinterpret_cast<uint8_t*>(packed);
    if (Cin >= 16) {
        sam::pack_pieces_kernel<<<ceil_div(C1 * Cin, 256), 256, 0, st>>>(W1, C1, Cin, Cin + 3, 0, ldexpf(1.f, s1), out);
        out += (size_t)C1 * Cin * 4;
    }
    sam::pack_pieces_kernel<<<ceil_div(C2 * C1, 256), 256, 0, st>>>(W2, C2, C1, C1, 0, ldexpf(1.f, s2), out);
    out += (size_t)C2 * C1 * 4;
    sam::pack_pieces_kernel<<<ceil_div(C3 * C2, 256), 256, 0, st>>>(W3, C3, C2, C2, 0, ldexpf(1.f, s3), out);
    return check_launch("gnb_pointconv_mlp_pack");
}

int32_t gnb_pointconv_mlp_max(const float* x, int64_t ldx, int32_t Cin, const float* pos_x, const float* pos_y, const int64_t* nbr,
                              const int32_t* cnt, const int64_t* eoffs, int64_t sumM, int32_t K, const void* packed, int32_t C1,
                              int32_t C2, int32_t C3, int32_t s1, int32_t s2, int32_t s3, const float* consts, int32_t* edge_ws,
                              float* out, void* stream) {
    GNB_REQUIRE(pos_x && pos_y && nbr && cnt && eoffs && packed && consts && edge_ws && out, "gnb_pointconv_mlp_max: null pointer");
    GNB_REQUIRE(Cin == 0 || x, "gnb_pointconv_mlp_max: null features");
    GNB_REQUIRE(gnb_pointconv_mlp_supported(Cin, C1, C2, C3), "gnb_pointconv_mlp_max: unsupported layer widths %d+3 -> %d -> %d -> %d", Cin, C1, C2, C3);
    GNB_REQUIRE(sumM * (int64_t)(K + 1) < (1ll << 31), "gnb_pointconv_mlp_max: more than 2^31 edges");
    if (sumM == 0) return GNB_OK;
    cudaStream_t st = as_stream(stream);
    const int64_t rows = sumM * (K + 1);
    int32_t* src = edge_ws;
    int32_t* dst = edge_ws + rows;
    sam::edge_index_kernel<<<(unsigned)ceil_div<int64_t>(sumM, 8), 256, 0, st>>>(nbr, cnt, eoffs, sumM, K, src, dst);
    GNB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)sumM * C3, st));
    sam::Params p;
    p.x = x; p.ldx = ldx; p.pos_x = pos_x; p.pos_y = pos_y; p.src = src; p.dst = dst; p.e_total = eoffs + sumM;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(packed);
    p.w1_img = img;
    if (Cin >= 16) img += (size_t)C1 * Cin * 4;
    p.w2_img = img;
    p.w3_img = img + (size_t)C2 * C1 * 4;
    p.inv1 = ldexpf(1.f, -s1); p.inv2 = ldexpf(1.f, -s2); p.inv3 = ldexpf(1.f, -s3);
    p.consts = consts;
    p.out_enc = reinterpret_cast<uint32_t*>(out);
    int32_t rc;
    if (Cin == 3) rc = sam::launch<3, 64, 64, 128>(p, st);
    else rc = sam::launch<128, 128, 128, 256>(p, st);
    if (rc != GNB_OK) return rc;
    sam::decode_max_kernel<<<(unsigned)ceil_div<int64_t>(sumM * C3, 256), 256, 0, st>>>(reinterpret_cast<uint32_t*>(out), sumM * C3);
    return check_launch("gnb_pointconv_mlp_max");
}

}  // extern "C"
