// Per-point Linear -> ReLU -> BatchNorm(eval) on the 5th-gen tensor cores (tcgen05 + TMEM), ref components/mlp.py:9-20.
//
//     Y[r, n] = post( sum_k X[r,k] * W[n,k] + bias[n] ),   post(v) = (relu ? max(v,0) : v) * bn_scale[n] + bn_shift[n]
//
// Every Linear of PointNet++ (SA edge MLPs with millions of rows, FP MLPs, heads), of the voxel aggregator and the
// decoders' hoisted first layer goes through this kernel.  With fp32 FFMA those layers were ALU-bound at ~27 TFLOP/s;
// here the contraction runs on tcgen05 and the layer becomes what it algorithmically is: a stream of X rows in and Y
// rows out (HBM-bound for K, N <= 256).
//
// Precision: like decode_tc.cu -- both operands are split into fp16 hi + lo and the product is formed as
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (~2^-21 relative per product), so the 1e-4 fp32 parity bound holds.
//
// One persistent warp-specialised CTA per SM, tiles of 128 rows x (<= 256 columns):
//   warps 0-7   A producers: coalesced 4-byte loads of the fp32 rows (lane = channel, no alignment requirement on ldx),
//               fp16 hi/lo split, stores in the UMMA canonical K-major SWIZZLE_128B layout; ring of 2-4 stages of one
//               64-channel chunk each, so loads of chunk c+1.. overlap the MMAs of chunk c.
//   warp 17     W loader: weights are pre-packed (gnb_linear_tc_pack) into shared-memory images per (column block,
//               K-chunk, hi|lo); when all of them fit next to the A ring they are loaded ONCE per CTA and stay resident,
//               otherwise cp.async.bulk streams them through a ring.
//   warp 16     MMA issuer (one elected thread, M=128, N=Npad, K=16 per instruction, 3 instructions per product).
//   warps 8-15  epilogue (two per TMEM lane quarter): tcgen05.ld (lane = row) -> bias / ReLU / BN affine -> 32 x 32 fp32
//               box in shared memory (SWIZZLE_128B, conflict-free) -> one TMA store per box (cp.async.bulk.tensor);
//               per-thread stores only for unaligned outputs and for the tile holding the device-side row count.
// Accumulators are double-buffered in TMEM (2 x 256 columns): the epilogue of tile t overlaps the MMAs of tile t+1.
#include "tc_common.cuh"
#include <cuda.h>

namespace gnb {

constexpr int LT_M = 128, LT_KC = 64, LT_THREADS = 576, LT_MAX_A = 4, LT_MAX_B = 20;
constexpr int LT_A_PART = LT_M * LT_KC * 2;   // 16 KB: one precision part of one A stage
constexpr int LT_A_STAGE = 2 * LT_A_PART;     // hi + lo
constexpr int LT_EPI_WARPS = 8;                  // two per TMEM lane quarter, alternating 32-column groups
constexpr int LT_STAGE_BYTES = 32 * 32 * 4;      // 4 KB per epilogue warp: one 32-row x 32-column fp32 box staged for a TMA store
constexpr int LT_SMEM_BUDGET = 225 * 1024 - LT_EPI_WARPS * LT_STAGE_BYTES;  // dynamic shared memory available to the A / W rings

// Per-column epilogue parameters (bias | bn_scale | bn_shift, padded) in constant memory: read as uniform constant-bank
// operands instead of 24 global loads per 32-column group (filled per launch, stream-ordered device-to-device copy).
constexpr int LT_MAX_COLS = 1024;
__constant__ float c_lt[3 * LT_MAX_COLS];

struct LtParams {
    const float* X;
    int64_t R, ldx;
    int K, N, Npad, n_blocks, nchunk, relu;
    const uint8_t* w_packed;   // [n_blocks][nchunk][hi,lo][Npad*128 B]
    const float* cparams;      // [3][n_blocks*Npad]: bias, bn_scale, bn_shift (padded columns: 0, 1, 0)
    float acc_scale;           // 2^-s: undoes the power-of-two scaling applied to the packed weights
    float* Y;
    int64_t ldy;
    const int64_t* rows_dev;   // nullable: number of valid rows on the device
    int na, nb, resident, piece_bytes;
    int64_t m_tiles;
    int vec_store;             // Y rows are 16-byte aligned: float4 stores
    int use_tma;               // Y rows are 16-byte aligned: full tiles leave through TMA stores of 128 x 32 boxes
    const int32_t* seg;        // segmented-max mode: segment id of every row (rows of a segment are contiguous), else NULL
    uint32_t* seg_out;         // [nseg, ldo] order-preserving encoded maxima (zero-initialised by the caller)
    int64_t ldo;
    int use_const;             // cparams fit in c_lt (n_blocks * Npad <= LT_MAX_COLS)
    int split;                 // Npad <= 128: the lo*hi + hi*lo cross terms accumulate in their own TMEM columns (+Npad)
    uint32_t* range_flag;      // nullable: raised when an OUTPUT value is outside +-65504 or not finite (gnb_linear_tc_flagged)
};

__global__ void __launch_bounds__(LT_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap y_map, const LtParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ __align__(8) uint64_t bars[2 * LT_MAX_A + 2 * LT_MAX_B + 4];
    __shared__ uint32_t tmem_ptr_smem;
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8 * s; };
    auto a_empty = [&](int s) { return bar0 + 8 * (LT_MAX_A + s); };
    auto b_full = [&](int s) { return bar0 + 8 * (2 * LT_MAX_A + s); };
    auto b_empty = [&](int s) { return bar0 + 8 * (2 * LT_MAX_A + LT_MAX_B + s); };
    auto d_full = [&](int s) { return bar0 + 8 * (2 * LT_MAX_A + 2 * LT_MAX_B + s); };
    auto d_empty = [&](int s) { return bar0 + 8 * (2 * LT_MAX_A + 2 * LT_MAX_B + 2 + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_MAX_A; ++s) { mbar_init(a_full(s), 128); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < LT_MAX_B; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), 32 * LT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);

    int64_t R = p.R;
    if (p.rows_dev != nullptr) { const int64_t rd = *p.rows_dev; R = rd < R ? rd : R; }
    const int64_t m_tiles = (R + LT_M - 1) / LT_M;         // tiles beyond the device-side row count are never touched
    const int64_t num_tiles = m_tiles * p.n_blocks;
    const uint32_t b_ring = sbase + p.na * LT_A_STAGE;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Npad >> 3) << 17) | ((uint32_t)(LT_M >> 4) << 24);

    if (warp < 8) {
        // =========================== A producers ===========================
        // Two groups of four warps take alternate stages (group g: stages with st % 2 == g; warp w of the group: rows
        // 32w..32w+31), so the global loads of two stages are in flight at the same time.
        const int grp = warp >> 2, wq = warp & 3;
        uint32_t st = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t r0 = (tile / p.n_blocks) * LT_M + wq * 32;
            for (int c = 0; c < p.nchunk; ++c, ++st) {
                if ((int)(st & 1) != grp) continue;
                const int slot = st % p.na;
                mbar_wait(a_empty(slot), ((st / p.na) & 1) ^ 1);
                uint8_t* ah = smem + slot * LT_A_STAGE;
                uint8_t* al = ah + LT_A_PART;
                // lane l owns the ADJACENT channels 2l, 2l+1 of the chunk: two 4-byte loads (no alignment requirement on
                // ldx), one saturating f16x2 conversion per precision part, one 4-byte swizzled store per part.  Only the
                // columns the MMAs of this chunk read (the first ceil16(K - 64c)) are produced.
                const int kc = min(LT_KC, p.K - c * LT_KC);
                const int kread = (kc + 15) & ~15;
                const int k0 = c * LT_KC + 2 * lane;
                const bool lane_on = 2 * lane < kread;
                const bool has0 = k0 < p.K, has1 = k0 + 1 < p.K;
                if (lane_on) {
#pragma unroll
                    for (int rb = 0; rb < 4; ++rb) {
                        float v0[8], v1[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int64_t r = r0 + rb * 8 + j;
                            const float* src = p.X + r * p.ldx + k0;
                            v0[j] = (r < R && has0) ? __ldg(src) : 0.f;
                            v1[j] = (r < R && has1) ? __ldg(src + 1) : 0.f;
                        }
                        // rows wq*32 + rb*8 + j, j = 0..7: one 8-row swizzle group
                        const uint32_t obase = (uint32_t)((wq * 4 + rb) * 1024 + (lane & 3) * 4);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t hi = cvt_f16x2_sat(v0[j], v1[j]);
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                            // inputs beyond the fp16 range saturate (hi = +-65504, lo = the saturated remainder): the layers
                            // of this network see activations of O(1e2) at most
                            const uint32_t lo = cvt_f16x2_sat(v0[j] - hf.x, v1[j] - hf.y);
                            const uint32_t o = obase + (uint32_t)(j * 128 + ((((lane >> 2) ^ j) & 7) << 4));
                            *reinterpret_cast<uint32_t*>(ah + o) = hi;
                            *reinterpret_cast<uint32_t*>(al + o) = lo;
                        }
                    }
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
                mbar_arrive(a_full(slot));
            }
        }
    } else if (warp < 16) {
        // =========================== epilogue ===========================
        const int q = warp & 3;          // TMEM lane quarter this warp may access (warp id mod 4)
        const int ghalf = (warp - 8) >> 2;  // this warp handles the 32-column groups g with (g & 1) == ghalf
        const int row = q * 32 + lane;
        const int ntot = p.n_blocks * p.Npad;
        // every epilogue warp owns a 4 KB staging box and issues its own TMA stores: no cross-warp synchronisation
        const uint32_t stage = b_ring + p.nb * p.piece_bytes + (uint32_t)(warp - 8) * LT_STAGE_BYTES;
        const uint32_t sdst = stage + (uint32_t)lane * 128;
        int it = 0;
        bool bad = false;   // an output outside the fp16 range of the consumer's operand split (or NaN), see range_flag
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int db = it & 1;
            const int nb = (int)(tile % p.n_blocks);
            const int64_t row0 = (tile / p.n_blocks) * LT_M;
            const int64_t grow = row0 + row;
            // Full tiles are written by the TMA unit (coalesced, asynchronous); the tile that contains the device-side
            // row count falls back to per-thread stores so rows >= R are never touched.
            const bool tma_tile = p.use_tma && row0 + LT_M <= R;
            // segmented-max mode (PointConv aggregation fused into the last edge-MLP layer): the tile never goes to HBM;
            // each warp reduces its 32 x 32 boxes column-wise over the runs of equal segment id and merges the run maxima
            // into the per-segment result with order-preserving integer atomics
            const bool seg_mode = p.seg != nullptr;
            int myseg = -1;
            unsigned bmask = 0;
            if (seg_mode) {
                if (grow < R) myseg = __ldg(p.seg + grow);
                const int prev = __shfl_up_sync(0xffffffffu, myseg, 1);
                bmask = __ballot_sync(0xffffffffu, lane == 0 || myseg != prev);   // run starts among this warp's 32 rows
            }
            mbar_wait_sleep(d_full(db), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(db * 256);
            const float* cb = p.cparams + nb * p.Npad;
            float* dst = p.Y + grow * p.ldy + (int64_t)nb * p.Npad;
            const int ncols = min(p.Npad, p.N - nb * p.Npad);  // valid columns of this block
#pragma unroll 1
            for (int n0 = ghalf * 32; n0 < ncols; n0 += 64) {
                uint32_t r[32];
                tmem_ld32(taddr + n0, r);
                if (p.split) {
                    uint32_t x[32];
                    tmem_ld32(taddr + p.Npad + n0, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int t = 0; t < 32; ++t) r[t] = __float_as_uint(__uint_as_float(r[t]) + __uint_as_float(x[t]));
                } else {
                    tmem_ld_wait();
                }
                if (tma_tile) {
                    // the previous store of this warp must have finished READING the staging box
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
                if (seg_mode) __syncwarp();   // the column reduction of the previous box has finished reading the stage
                if (tma_tile || seg_mode || grow < R) {
                    const int cidx = nb * p.Npad + n0;
#pragma unroll
                    for (int t = 0; t < 32; t += 4) {
                        float4 bb, sc, sh;
                        if (p.use_const) {
                            bb = make_float4(c_lt[cidx + t], c_lt[cidx + t + 1], c_lt[cidx + t + 2], c_lt[cidx + t + 3]);
                            sc = make_float4(c_lt[ntot + cidx + t], c_lt[ntot + cidx + t + 1], c_lt[ntot + cidx + t + 2], c_lt[ntot + cidx + t + 3]);
                            sh = make_float4(c_lt[2 * ntot + cidx + t], c_lt[2 * ntot + cidx + t + 1], c_lt[2 * ntot + cidx + t + 2], c_lt[2 * ntot + cidx + t + 3]);
                        } else {
                            bb = __ldg(reinterpret_cast<const float4*>(cb + n0 + t));
                            sc = __ldg(reinterpret_cast<const float4*>(cb + ntot + n0 + t));
                            sh = __ldg(reinterpret_cast<const float4*>(cb + 2 * ntot + n0 + t));
                        }
                        float4 o;
                        o.x = fmaf(__uint_as_float(r[t]), p.acc_scale, bb.x);
                        o.y = fmaf(__uint_as_float(r[t + 1]), p.acc_scale, bb.y);
                        o.z = fmaf(__uint_as_float(r[t + 2]), p.acc_scale, bb.z);
                        o.w = fmaf(__uint_as_float(r[t + 3]), p.acc_scale, bb.w);
                        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y);
                        o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
                        bad |= !(fabsf(o.x) <= 65504.f) | !(fabsf(o.y) <= 65504.f) | !(fabsf(o.z) <= 65504.f) | !(fabsf(o.w) <= 65504.f);
                        const int n = n0 + t;
                        if (tma_tile || seg_mode) {
                            // SWIZZLE_128B box: 16-byte chunk j of row r lives at chunk j ^ (r & 7) (bank-conflict free)
                            const uint32_t a = sdst + ((uint32_t)(((t >> 2) ^ (lane & 7)) & 7) << 4);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                        } else if (p.vec_store && n + 3 < ncols) {
                            *reinterpret_cast<float4*>(dst + n) = o;
                        } else {
                            if (n < ncols) dst[n] = o.x;
                            if (n + 1 < ncols) dst[n + 1] = o.y;
                            if (n + 2 < ncols) dst[n + 2] = o.z;
                            if (n + 3 < ncols) dst[n + 3] = o.w;
                        }
                    }
                }
                if (seg_mode) {
                    __syncwarp();
                    const int col = n0 + lane;                       // this lane's column of the box
                    const bool col_ok = col < ncols;
                    unsigned m = bmask;
                    while (m) {                                      // warp-uniform loop over the runs
                        const int start = __ffs(m) - 1;
                        m &= m - 1;
                        const int end = m ? __ffs(m) - 1 : 32;
                        const int sid = __shfl_sync(0xffffffffu, myseg, start);
                        if (sid < 0) continue;                       // rows beyond the device-side row count
                        float mx = -INFINITY;
                        for (int rr = start; rr < end; ++rr) {
                            float v;
                            const uint32_t a = stage + (uint32_t)rr * 128 + ((uint32_t)(((lane >> 2) ^ (rr & 7)) & 7) << 4) + (uint32_t)(lane & 3) * 4;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                            mx = fmaxf(mx, v);
                        }
                        if (col_ok) {
                            const unsigned b = __float_as_uint(mx);
                            atomicMax(p.seg_out + (int64_t)sid * p.ldo + (int64_t)nb * p.Npad + col, (b & 0x80000000u) ? ~b : (b | 0x80000000u));
                        }
                    }
                } else if (tma_tile) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&y_map),
                                     "r"(stage), "r"(nb * p.Npad + n0), "r"((int)row0 + q * 32)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(d_empty(db));  // accumulator buffer drained: the MMAs of tile it+2 may overwrite it
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (bad && p.range_flag != nullptr) *p.range_flag = 1u;
    } else if (warp == 16) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t st = 0, piece = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int db = it & 1;
                mbar_wait_sleep(d_empty(db), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(db * 256);
                for (int c = 0; c < p.nchunk; ++c, ++st, ++piece) {
                    const int slot = st % p.na;
                    int bslot;
                    if (p.resident) {
                        bslot = c;
                        mbar_wait_sleep(b_full(bslot), 0);       // completes once; stays complete for every later tile
                    } else {
                        bslot = piece % p.nb;
                        mbar_wait_sleep(b_full(bslot), (piece / p.nb) & 1);
                    }
                    mbar_wait_sleep(a_full(slot), (st / p.na) & 1);
                    tc_fence_after();
                    const uint32_t ahi = sbase + slot * LT_A_STAGE, alo = ahi + LT_A_PART;
                    const uint32_t bhi = b_ring + bslot * p.piece_bytes, blo = bhi + p.piece_bytes / 2;
                    const int kc = min(LT_KC, p.K - c * LT_KC);
                    const int nk = (kc + 15) >> 4;
                    for (int kk = 0; kk < nk; ++kk)
                        umma_f16(d_tmem, umma_desc(ahi + kk * 32), umma_desc(bhi + kk * 32), idesc, (c | kk) != 0);
                    // The tensor core's fp32 accumulator truncates on every accumulation (a systematic ~0.5 ulp bias per
                    // instruction): when TMEM has room the two small cross terms get their own accumulator, so the chain
                    // of full-magnitude addends is a third as long.
                    const uint32_t d_cross = p.split ? d_tmem + (uint32_t)p.Npad : d_tmem;
                    for (int kk = 0; kk < nk; ++kk)
                        umma_f16(d_cross, umma_desc(alo + kk * 32), umma_desc(bhi + kk * 32), idesc, !p.split || (c | kk) != 0);
                    for (int kk = 0; kk < nk; ++kk)
                        umma_f16(d_cross, umma_desc(ahi + kk * 32), umma_desc(blo + kk * 32), idesc, 1);
                    umma_commit(a_empty(slot));
                    if (!p.resident) umma_commit(b_empty(bslot));
                }
                umma_commit(d_full(db));
            }
        }
    } else {
        // =========================== W loader ===========================
        if (lane == 0) {
            if (p.resident) {
                if ((int64_t)blockIdx.x < num_tiles) {
                    for (int c = 0; c < p.nchunk; ++c) {
                        mbar_expect_tx(b_full(c), (uint32_t)p.piece_bytes);
                        bulk_g2s(b_ring + c * p.piece_bytes, p.w_packed + (size_t)c * p.piece_bytes, (uint32_t)p.piece_bytes, b_full(c));
                    }
                }
            } else {
                uint32_t piece = 0;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                    const int nb = (int)(tile % p.n_blocks);
                    for (int c = 0; c < p.nchunk; ++c, ++piece) {
                        const int slot = piece % p.nb;
                        mbar_wait_sleep(b_empty(slot), ((piece / p.nb) & 1) ^ 1);
                        mbar_expect_tx(b_full(slot), (uint32_t)p.piece_bytes);
                        bulk_g2s(b_ring + slot * p.piece_bytes, p.w_packed + ((size_t)nb * p.nchunk + c) * p.piece_bytes,
                                 (uint32_t)p.piece_bytes, b_full(slot));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

struct LtLayout { int n_blocks, Npad, nchunk, piece_bytes; };
static LtLayout lt_layout(int N, int K) {
    LtLayout l;
    // long contractions use 128-column blocks so that the split accumulators (see the MMA issuer) fit in TMEM
    const int max_cols = K > 512 ? 128 : 256;
    l.n_blocks = (N + max_cols - 1) / max_cols;
    l.Npad = ((N + l.n_blocks - 1) / l.n_blocks + 31) / 32 * 32;
    l.nchunk = (K + LT_KC - 1) / LT_KC;
    l.piece_bytes = 2 * l.Npad * 128;
    return l;
}

// W fp32 [N,K] -> [n_blocks][nchunk][hi,lo] images of [Npad rows x 64 K] fp16 (K-major, SWIZZLE_128B), zero padded;
// cparams [3][n_blocks*Npad] = bias | bn_scale | bn_shift with (0, 1, 0) in the padded columns.
__global__ void linear_tc_pack_kernel(const float* __restrict__ W, int N, int K, LtLayout l, float wscale,
                                      const float* __restrict__ bias, const float* __restrict__ bn_scale,
                                      const float* __restrict__ bn_shift, uint8_t* __restrict__ out, float* __restrict__ cparams) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ntot = l.n_blocks * l.Npad;
    const int64_t total = (int64_t)ntot * l.nchunk * LT_KC;
    if (t < ntot) {
        // column t of the padded layout is column (t / Npad) * Npad + t % Npad of W only if blocks are dense: they are,
        // block b covers W columns [b*Npad, (b+1)*Npad)
        const bool ok = t < N;
        cparams[t] = (ok && bias) ? bias[t] : 0.f;
        cparams[ntot + t] = (ok && bn_scale) ? bn_scale[t] : 1.f;
        cparams[2 * ntot + t] = (ok && bn_shift) ? bn_shift[t] : 0.f;
    }
    if (t >= total) return;
    const int kp = (int)(t % (l.nchunk * LT_KC));
    const int n = (int)(t / (l.nchunk * LT_KC));
    float w = 0.f;
    if (n < N && kp < K) w = W[(int64_t)n * K + kp] * wscale;
    w = fminf(fmaxf(w, -65504.f), 65504.f);
    const __half h = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(h));
    const int nb = n / l.Npad, nl = n % l.Npad, c = kp / LT_KC, kc = kp % LT_KC;
    uint8_t* piece = out + ((size_t)nb * l.nchunk + c) * l.piece_bytes;
    const uint32_t off = sw128_offset(nl, kc);
    *reinterpret_cast<__half*>(piece + off) = h;
    *reinterpret_cast<__half*>(piece + l.piece_bytes / 2 + off) = lo;
}

typedef CUresult (*LtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static LtEncodeFn lt_encode_fn() {
    static LtEncodeFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<LtEncodeFn>(ptr);
    return fn;
}

// seg[r] = i for the rows r in [offs[i], offs[i+1]) (one warp per segment)
__global__ void segment_ids_kernel(const int64_t* __restrict__ offs, int64_t nseg, int32_t* __restrict__ seg) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nseg) return;
    const int64_t e0 = offs[i], e1 = offs[i + 1];
    for (int64_t r = e0 + (threadIdx.x & 31); r < e1; r += 32) seg[r] = (int32_t)i;
}

// in place: order-preserving encoded maxima -> floats; a slot no row ever touched (still 0) becomes 0.0f (an empty
// segment aggregates to 0, like PyG's max aggregation)
__global__ void segmax_decode_kernel(uint32_t* __restrict__ enc, int64_t count) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t u = enc[t];
    enc[t] = u == 0u ? 0u : ((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int64_t gnb_linear_tc_packed_bytes(int32_t N, int32_t K) {
    if (N < 1 || K < 1) return 0;
    const LtLayout l = lt_layout(N, K);
    return (int64_t)l.n_blocks * l.nchunk * l.piece_bytes;
}

int64_t gnb_linear_tc_padded_cols(int32_t N, int32_t K) {
    if (N < 1 || K < 1) return 0;
    const LtLayout l = lt_layout(N, K);
    return (int64_t)l.n_blocks * l.Npad;
}

int32_t gnb_linear_tc_pack(const float* W, int32_t N, int32_t K, const float* bias, const float* bn_scale,
                           const float* bn_shift, int32_t scale_log2, void* packed, float* cparams, void* stream) {
    GNB_REQUIRE(W && packed && cparams, "gnb_linear_tc_pack: null pointer");
    GNB_REQUIRE(N >= 1 && K >= 1, "gnb_linear_tc_pack: bad shape %dx%d", N, K);
    const LtLayout l = lt_layout(N, K);
    const int64_t total = (int64_t)l.n_blocks * l.Npad * l.nchunk * LT_KC;
    linear_tc_pack_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        W, N, K, l, ldexpf(1.0f, scale_log2), bias, bn_scale, bn_shift, reinterpret_cast<uint8_t*>(packed), cparams);
    return check_launch("gnb_linear_tc_pack");
}

static int32_t linear_tc_launch(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed, const float* cparams,
                                int32_t scale_log2, int32_t N, int32_t relu, float* Y, int64_t ldy, const int64_t* rows_dev,
                                const int32_t* seg, uint32_t* seg_out, int64_t ldo, void* stream, uint32_t* range_flag = nullptr) {
    GNB_REQUIRE(X && packed && cparams && (Y || seg), "gnb_linear_tc: null pointer");
    GNB_REQUIRE(R >= 0 && K >= 1 && N >= 1 && ldx >= K && (seg || ldy >= N), "gnb_linear_tc: bad shape R=%lld K=%d N=%d ldx=%lld ldy=%lld",
                (long long)R, K, N, (long long)ldx, (long long)ldy);
    if (R == 0) return GNB_OK;
    const LtLayout l = lt_layout(N, K);
    LtParams p;
    p.X = X; p.R = R; p.ldx = ldx; p.K = K; p.N = N; p.Npad = l.Npad; p.n_blocks = l.n_blocks; p.nchunk = l.nchunk;
    p.relu = relu; p.w_packed = reinterpret_cast<const uint8_t*>(packed); p.cparams = cparams;
    p.acc_scale = ldexpf(1.0f, -scale_log2);
    p.Y = Y; p.ldy = ldy; p.rows_dev = rows_dev;
    p.seg = seg; p.seg_out = seg_out; p.ldo = ldo;
    p.range_flag = range_flag;
    p.piece_bytes = l.piece_bytes;
    p.m_tiles = ceil_div<int64_t>(R, LT_M);
    // separate cross-term accumulator whenever TMEM has room -- also for short contractions: measured on the PointNet++
    // chain (smoke configuration), sharing one accumulator doubles the end-to-end error (per-point logits 6.0e-5 ->
    // 1.15e-4 against the oracle) because EVERY tcgen05.mma truncates the full accumulator, however small its addend
    p.split = l.Npad <= 128 ? 1 : 0;
    p.vec_store = (!seg && (ldy % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) % 16) == 0 && (l.Npad % 4) == 0) ? 1 : 0;
    const int64_t resident_bytes = (int64_t)l.nchunk * l.piece_bytes;
    if (l.n_blocks == 1 && l.nchunk <= LT_MAX_B && resident_bytes + 2 * LT_A_STAGE <= LT_SMEM_BUDGET) {
        p.resident = 1;
        p.nb = l.nchunk;
        p.na = (int)((LT_SMEM_BUDGET - resident_bytes) / LT_A_STAGE);
    } else {
        p.resident = 0;
        p.na = 2;
        p.nb = (LT_SMEM_BUDGET - p.na * LT_A_STAGE) / l.piece_bytes;
        if (p.nb > 4) p.nb = 4;
        GNB_REQUIRE(p.nb >= 2, "gnb_linear_tc: weight piece of %d bytes does not fit the shared-memory ring", l.piece_bytes);
    }
    if (p.na > LT_MAX_A) p.na = LT_MAX_A;
    // TMA store of the output: [R rows x N columns] fp32 with row stride ldy, boxes of 32 rows x 32 columns (one per epilogue warp)
    CUtensorMap y_map;
    memset(&y_map, 0, sizeof(y_map));
    p.use_tma = 0;
    // (whole boxes only: N a multiple of 32 -- the TMA unit is never asked to clip)
    if (p.vec_store && (N % 32) == 0 && R >= LT_M && R < (1ll << 31)) {
        LtEncodeFn enc = lt_encode_fn();
        if (!enc) { set_error("gnb_linear_tc: cuTensorMapEncodeTiled is not available from this driver"); return GNB_ERR_CUDA; }
        const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)R};
        const cuuint64_t gstr[1] = {(cuuint64_t)ldy * 4};
        const cuuint32_t box[2] = {32, 32};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&y_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("gnb_linear_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return GNB_ERR_CUDA; }
        p.use_tma = 1;
    }
    p.use_const = l.n_blocks * l.Npad <= LT_MAX_COLS ? 1 : 0;
    ConstBankGuard guard(BANK_LINEAR_TC, as_stream(stream));   // the launch below reads c_lt
    if (p.use_const)
        GNB_CUDA(cudaMemcpyToSymbolAsync(c_lt, cparams, sizeof(float) * 3 * l.n_blocks * l.Npad, 0, cudaMemcpyDeviceToDevice, as_stream(stream)));
    const int smem = p.na * LT_A_STAGE + p.nb * l.piece_bytes + LT_EPI_WARPS * LT_STAGE_BYTES + 1024;
    GNB_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int64_t tiles = p.m_tiles * l.n_blocks;
    int grid = sm_count();
    if ((int64_t)grid > tiles) grid = (int)tiles;
    linear_tc_kernel<<<grid, LT_THREADS, smem, as_stream(stream)>>>(y_map, p);
    return check_launch("gnb_linear_tc");
}

int32_t gnb_linear_tc(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed, const float* cparams,
                      int32_t scale_log2, int32_t N, int32_t relu, float* Y, int64_t ldy, const int64_t* rows_dev,
                      void* stream) {
    GNB_REQUIRE(Y, "gnb_linear_tc: null pointer");
    return linear_tc_launch(X, R, K, ldx, packed, cparams, scale_log2, N, relu, Y, ldy, rows_dev, nullptr, nullptr, 0, stream);
}

int32_t gnb_linear_tc_flagged(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed, const float* cparams,
                              int32_t scale_log2, int32_t N, int32_t relu, float* Y, int64_t ldy, const int64_t* rows_dev,
                              void* stream) {
    GNB_REQUIRE(Y, "gnb_linear_tc_flagged: null pointer");
    uint32_t* flag = f16_flag_ptr();
    GNB_REQUIRE(flag != nullptr, "gnb_linear_tc_flagged: range flag allocation failed");
    return linear_tc_launch(X, R, K, ldx, packed, cparams, scale_log2, N, relu, Y, ldy, rows_dev, nullptr, nullptr, 0, stream, flag);
}

int32_t gnb_linear_tc_segmax(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed, const float* cparams,
                             int32_t scale_log2, int32_t N, int32_t relu, const int32_t* seg, void* enc_out, int64_t ldo,
                             const int64_t* rows_dev, void* stream) {
    GNB_REQUIRE(seg && enc_out && ldo >= N, "gnb_linear_tc_segmax: bad arguments");
    return linear_tc_launch(X, R, K, ldx, packed, cparams, scale_log2, N, relu, nullptr, 0, rows_dev, seg,
                            reinterpret_cast<uint32_t*>(enc_out), ldo, stream);
}

int32_t gnb_segment_ids(const int64_t* offs, int64_t nseg, int32_t* seg, void* stream) {
    GNB_REQUIRE(offs && seg && nseg >= 0, "gnb_segment_ids: bad arguments");
    if (nseg == 0) return GNB_OK;
    segment_ids_kernel<<<(unsigned)ceil_div<int64_t>(nseg, 8), 256, 0, as_stream(stream)>>>(offs, nseg, seg);
    return check_launch("gnb_segment_ids");
}

int32_t gnb_segmax_decode(void* enc, int64_t count, void* stream) {
    GNB_REQUIRE(enc && count >= 0, "gnb_segmax_decode: bad arguments");
    if (count == 0) return GNB_OK;
    segmax_decode_kernel<<<(unsigned)ceil_div<int64_t>(count, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<uint32_t*>(enc), count);
    return check_launch("gnb_segmax_decode");
}

}  // extern "C"
