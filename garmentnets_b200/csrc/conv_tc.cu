// 3x3x3 convolution of the 3D-UNet on the 5th-gen tensor cores: TMA-staged implicit GEMM (ref components/unet3d.py:43-72,
// 'gcr' SingleConv = GroupNorm -> Conv3d(pad 1, no bias) -> ReLU).
//
//   rows (M)  = 128 output voxels = one TMA box of the channels-last activation (e.g. 32 w x 4 h for a 32^3 grid)
//   cols (N)  = Cout (32 .. 256)
//   K         = 27 taps x Cin, consumed in stages of one tap x 64 channels
//
// The GroupNorm affine cannot be folded into the weights (zero padding is applied AFTER the normalisation and the
// statistics are per sample), so a light elementwise pass (gn_apply_split_kernel) writes the normalised activation once
// as fp16 hi + lo (channels padded to a multiple of 64).  The conv kernel then never touches the operand with a thread:
// for every (tap, channel chunk) ONE elected thread issues two cp.async.bulk.tensor (TMA) loads of the box shifted by
// the tap offset -- out-of-bounds rows/columns are zero-filled by the TMA unit, which is exactly the conv padding --
// straight into the UMMA K-major SWIZZLE_128B layout, plus two bulk copies of the pre-packed weight images.
// Like the decoder kernel, products are formed as hi*hi + lo*hi + hi*lo (fp32 accumulation in TMEM) to stay within the
// 1e-4 fp32 parity bound.  The tensor core's fp32 accumulator truncates on every accumulation, which is a systematic
// bias of ~0.5 ulp per tcgen05.mma: over the 27*Cin/16*3 = 648 instructions of a 128-channel layer that is ~4e-5
// relative per layer (measured: 2.3e-4 after the 15 layers of the UNet).  The accumulation chain is therefore cut:
// the hi*hi products rotate over THREE accumulators (stage mod 3) and the two small cross terms go to a fourth, so no
// accumulator sees more than 27*Cin/64*4/3 instructions of full-magnitude addends; the epilogue adds the four in fp32
// with round-to-nearest.  Warp roles: TMA/bulk loader, MMA issuer, 4 epilogue warps (tcgen05.ld -> ReLU -> fp32 NDHWC
// stores); accumulators are double-buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace gnb {

// ---- PTX wrappers (same conventions as decode_tc.cu) -------------------------------------------------------------------
namespace ctc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// e4m3 x e4m3 -> fp32: K = 32 elements (32 bytes) per instruction, twice the MACs of kind::f16 at the same operand bytes.
// The instruction descriptor is the f16 one (format code 0 = F16 there, = E4M3 here; fp32 accumulator).
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__host__ __device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + (c & 7) * 2);
}
// byte j (0..127) of row r of a K-major SWIZZLE_128B tile (8-bit operands)
__host__ __device__ __forceinline__ uint32_t sw128_byte_offset(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((j >> 4) ^ (r & 7)) & 7) << 4) + (j & 15));
}
}  // namespace ctc

// ---- cross-term precision ----------------------------------------------------------------------------------------------
// Mode 0 (default): a*w = hi*hi + lo*hi + hi*lo, three kind::f16 MMAs per K-step.
// Mode 1: the two cross terms as ONE kind::f8f6f4 MMA per K-step.  The "lo" activation tensor then holds, per channel pair
// (c, c+1), the four e4m3 bytes [lo(c)*2^12, lo(c+1)*2^12, a(c), a(c+1)] and the "lo" weight image the matching
// [w(c)*2^(s-12), w(c+1)*2^(s-12), w_lo(c)*2^s, w_lo(c+1)*2^s]: 64 channels = 128 one-byte K elements = the same 128-byte
// rows, the same TMA boxes and shared-memory descriptors, and  sum_K' a8*w8 = 2^s (lo*w + a*w_lo).  Two tensor
// pass-equivalents instead of three.  e4m3 keeps 4 significant bits of each cross term (relative error 2^-5 of something that is
// 2^-11 of the product): ~2e-5 relative per layer instead of ~1e-7 (tools/fp8_cross_term_sim.py); saturation (|a| > 448, or a
// remainder above 448 * 2^-12) degrades towards one-pass accuracy for that element instead of failing.
constexpr float CT_LO8_SCALE = 4096.f;   // 2^12
static int g_cross_fp8 = 0;

constexpr int CT_M = 128, CT_KC = 64, CT_THREADS = 192, CT_MAX_STAGES = 4;
constexpr int CT_NACC = 4;  // 3 rotating hi*hi accumulators + 1 for the cross terms
constexpr int CT_A_BYTES = CT_M * CT_KC * 2;  // 16 KB per precision part

struct ConvTcParams {
    int B, D, H, W, Cpad, Cout;
    int bw, bh, bd, bb;        // TMA box (voxels): bw*bh*bd*bb == 128
    int nchunk;                // Cpad / 64
    int nstages, stage_bytes, b_bytes;  // b_bytes = Cout*128 (one precision part of one weight piece)
    int relu;
    int nbuf;                  // 2: accumulator sets double-buffered in TMEM (4*acc_stride*2 <= 512 columns), else 1
    int acc_stride;            // TMEM columns between accumulators: Cout rounded up to a power of two (32/64/128)
    float out_scale;           // 2^-s: undoes the power-of-two scaling applied to the packed weights
    int cross_fp8;             // cross terms as one e4m3 MMA per K-step (see g_cross_fp8)
    const uint8_t* w_packed;   // [27][nchunk][hi,lo][Cout*128 B]
    float* y;                  // [B,D,H,W,Cout] fp32
    int64_t num_tiles;
};

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const ConvTcParams p) {
    using namespace ctc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ __align__(8) uint64_t bars[2 * CT_MAX_STAGES + 4];
    __shared__ uint32_t tmem_ptr_smem;
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8 * s; };
    auto empty = [&](int s) { return bar0 + 8 * (CT_MAX_STAGES + s); };
    auto d_full = [&](int s) { return bar0 + 8 * (2 * CT_MAX_STAGES + s); };
    auto d_empty = [&](int s) { return bar0 + 8 * (2 * CT_MAX_STAGES + 2 + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_MAX_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);

    const int nbw = p.W / p.bw, nbh = p.H / p.bh, nbd = p.D / p.bd;
    const int ksteps = 27 * p.nchunk;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);

    if (warp == 0) {
        // =========================== loader: TMA activations + bulk weights ===========================
        if (lane == 0) {
            uint32_t st = 0;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int64_t t = tile;
                const int iw = (int)(t % nbw); t /= nbw;
                const int ih = (int)(t % nbh); t /= nbh;
                const int id = (int)(t % nbd); t /= nbd;
                const int w0 = iw * p.bw, h0 = ih * p.bh, d0 = id * p.bd, b0 = (int)t * p.bb;
                for (int ks = 0; ks < ksteps; ++ks, ++st) {
                    const int tap = ks / p.nchunk, cc = ks - tap * p.nchunk;
                    const int dz = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
                    const int slot = st % p.nstages;
                    mbar_wait(empty(slot), ((st / p.nstages) & 1) ^ 1);
                    mbar_expect_tx(full(slot), (uint32_t)p.stage_bytes);
                    const uint32_t sa = sbase + slot * p.stage_bytes;
                    tma_load_5d(sa, &map_hi, full(slot), cc * CT_KC, w0 + dx, h0 + dy, d0 + dz, b0);
                    tma_load_5d(sa + CT_A_BYTES, &map_lo, full(slot), cc * CT_KC, w0 + dx, h0 + dy, d0 + dz, b0);
                    const uint8_t* wsrc = p.w_packed + (size_t)ks * 2 * p.b_bytes;
                    bulk_g2s(sa + 2 * CT_A_BYTES, wsrc, (uint32_t)p.b_bytes, full(slot));
                    bulk_g2s(sa + 2 * CT_A_BYTES + p.b_bytes, wsrc + p.b_bytes, (uint32_t)p.b_bytes, full(slot));
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t st = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int db = it % p.nbuf;
                mbar_wait(d_empty(db), ((it / p.nbuf) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_set = tmem_base + (uint32_t)(db * CT_NACC * p.acc_stride);
                const uint32_t d_cross = d_set + 3u * p.acc_stride;
                for (int ks = 0; ks < ksteps; ++ks, ++st) {
                    const int slot = st % p.nstages;
                    mbar_wait(full(slot), (st / p.nstages) & 1);
                    tc_fence_after();
                    const uint32_t ahi = sbase + slot * p.stage_bytes, alo = ahi + CT_A_BYTES;
                    const uint32_t bhi = ahi + 2 * CT_A_BYTES, blo = bhi + p.b_bytes;
                    const uint32_t d_main = d_set + (uint32_t)((ks % 3) * p.acc_stride);
#pragma unroll
                    for (int kk = 0; kk < CT_KC / 16; ++kk)
                        umma_f16(d_main, umma_desc(ahi + kk * 32), umma_desc(bhi + kk * 32), idesc, (ks >= 3) || kk != 0);
                    if (p.cross_fp8) {
#pragma unroll
                        for (int kk = 0; kk < CT_KC / 16; ++kk)   // 32 one-byte K elements per instruction: the same 32-byte steps
                            umma_f8(d_cross, umma_desc(alo + kk * 32), umma_desc(blo + kk * 32), idesc, (ks | kk) != 0);
                    } else {
#pragma unroll
                        for (int kk = 0; kk < CT_KC / 16; ++kk)
                            umma_f16(d_cross, umma_desc(alo + kk * 32), umma_desc(bhi + kk * 32), idesc, (ks | kk) != 0);
#pragma unroll
                        for (int kk = 0; kk < CT_KC / 16; ++kk)
                            umma_f16(d_cross, umma_desc(ahi + kk * 32), umma_desc(blo + kk * 32), idesc, 1);
                    }
                    umma_commit(empty(slot));
                }
                umma_commit(d_full(db));
            }
        }
    } else {
        // =========================== epilogue (warps 2..5) ===========================
        const int q = warp & 3;  // TMEM lane quarter
        const int row = q * 32 + lane;
        // row -> voxel inside the box (w fastest, then h, d, b: the order TMA writes the box)
        int r = row;
        const int lw = r % p.bw; r /= p.bw;
        const int lh = r % p.bh; r /= p.bh;
        const int ld = r % p.bd; r /= p.bd;
        const int lb = r;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            int64_t t = tile;
            const int iw = (int)(t % nbw); t /= nbw;
            const int ih = (int)(t % nbh); t /= nbh;
            const int id = (int)(t % nbd); t /= nbd;
            const int64_t vox = ((((int64_t)t * p.bb + lb) * p.D + id * p.bd + ld) * p.H + ih * p.bh + lh) * p.W + iw * p.bw + lw;
            float* dst = p.y + vox * p.Cout;
            const int db = it % p.nbuf;
            mbar_wait(d_full(db), (it / p.nbuf) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(db * CT_NACC * p.acc_stride);
            for (int n0 = 0; n0 < p.Cout; n0 += 32) {
                uint32_t v[32], u[32];
                tmem_ld32(taddr + n0, v);
                tmem_ld32(taddr + p.acc_stride + n0, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
                tmem_ld32(taddr + 2 * p.acc_stride + n0, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
                tmem_ld32(taddr + 3 * p.acc_stride + n0, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = __uint_as_float(v[j]) * p.out_scale; o.y = __uint_as_float(v[j + 1]) * p.out_scale;
                    o.z = __uint_as_float(v[j + 2]) * p.out_scale; o.w = __uint_as_float(v[j + 3]) * p.out_scale;
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(dst + n0 + j) = o;
                }
            }
            tc_fence_before();
            mbar_arrive(d_empty(db));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- stacked-dx variant for the narrow layers (Cout = 32) ---------------------------------------------------------------
// With N = Cout = 32 every tcgen05.mma reads a 4 KB A slab from shared memory for 64 K MACs: the layer is bound by the
// operand traffic (27 shifted copies of the activation through L2 -> smem -> tensor core), not by the tensor pipe.
// Here the three kw taps of a (kd, kh) pair share ONE unshifted activation box: their weights are stacked along N,
//     Z[v, j*32 + n] = sum_c X[v + (kd, kh, 0), c] * W[n, c, kd, kh, j]            (j = kw, N = 96)
// and the shift along W is applied to the OUTPUT in the epilogue,  y[w] = Z_0[w-1] + Z_1[w] + Z_2[w+1], with warp
// shuffles: a tile row is a whole W line (box width == W), so the neighbours are adjacent TMEM lanes of the same warp and
// the line ends are exactly the zero padding.  The fp16 split costs 2 MMAs per K-step instead of 3: A_hi x [W_hi | W_lo]
// (N = 192: hi*hi in columns 0..95, hi*lo in 96..191) and A_lo x W_hi (N = 96) accumulated onto columns 96..191.
// Cout = 64 stacks to N = 192 per precision part (three MMAs per K-step, one accumulator set, no double buffering).
// Per output voxel that is 9 activation boxes instead of 27 and 18 MMAs per 64 channels instead of 81; the accumulation
// chain of full-magnitude addends is 9*Cin/16 instructions long, so no accumulator rotation is needed.
struct ConvDxParams {
    int B, D, H, W, Cpad, Cout;
    int bw, bh, bd, bb;
    int nchunk, nstages, stage_bytes, b_bytes;   // b_bytes = 2 * 3 * Cout * 128 (hi rows then lo rows of one (kd,kh,chunk) piece)
    int relu;
    float out_scale;
    int cross_fp8;
    const uint8_t* w_packed;   // [9][nchunk][hi: 3*Cout rows, lo: 3*Cout rows][128 B]
    float* y;
    int64_t num_tiles;
};

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_dx_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                  const ConvDxParams p) {
    using namespace ctc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ __align__(8) uint64_t bars[2 * CT_MAX_STAGES + 4];
    __shared__ uint32_t tmem_ptr_smem;
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8 * s; };
    auto empty = [&](int s) { return bar0 + 8 * (CT_MAX_STAGES + s); };
    auto d_full = [&](int s) { return bar0 + 8 * (2 * CT_MAX_STAGES + s); };
    auto d_empty = [&](int s) { return bar0 + 8 * (2 * CT_MAX_STAGES + 2 + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < CT_MAX_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_ptr_smem);

    const int nbw = p.W / p.bw, nbh = p.H / p.bh, nbd = p.D / p.bd;
    const int ksteps = 9 * p.nchunk;
    const int NJ = 3 * p.Cout;   // stacked columns of one precision part (96 or 192)
    const bool wide = 2 * NJ <= 256;          // Cout = 32: [W_hi | W_lo] fits one N = 192 instruction
    const int nbuf = 4 * NJ <= 512 ? 2 : 1;   // accumulator sets (2*NJ columns each) double-buffered when TMEM has room
    const uint32_t idesc_wide = (1u << 4) | ((uint32_t)((2 * NJ) >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
    const uint32_t idesc_half = (1u << 4) | ((uint32_t)(NJ >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t st = 0;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int64_t t = tile;
                const int iw = (int)(t % nbw); t /= nbw;
                const int ih = (int)(t % nbh); t /= nbh;
                const int id = (int)(t % nbd); t /= nbd;
                const int w0 = iw * p.bw, h0 = ih * p.bh, d0 = id * p.bd, b0 = (int)t * p.bb;
                for (int ks = 0; ks < ksteps; ++ks, ++st) {
                    const int tap = ks / p.nchunk, cc = ks - tap * p.nchunk;   // tap = kd*3 + kh
                    const int dz = tap / 3 - 1, dy = tap % 3 - 1;
                    const int slot = st % p.nstages;
                    mbar_wait(empty(slot), ((st / p.nstages) & 1) ^ 1);
                    mbar_expect_tx(full(slot), (uint32_t)p.stage_bytes);
                    const uint32_t sa = sbase + slot * p.stage_bytes;
                    tma_load_5d(sa, &map_hi, full(slot), cc * CT_KC, w0, h0 + dy, d0 + dz, b0);
                    tma_load_5d(sa + CT_A_BYTES, &map_lo, full(slot), cc * CT_KC, w0, h0 + dy, d0 + dz, b0);
                    bulk_g2s(sa + 2 * CT_A_BYTES, p.w_packed + (size_t)ks * p.b_bytes, (uint32_t)p.b_bytes, full(slot));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t st = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int db = it % nbuf;
                mbar_wait(d_empty(db), ((it / nbuf) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_set = tmem_base + (uint32_t)(db * 256);
                for (int ks = 0; ks < ksteps; ++ks, ++st) {
                    const int slot = st % p.nstages;
                    mbar_wait(full(slot), (st / p.nstages) & 1);
                    tc_fence_after();
                    const uint32_t ahi = sbase + slot * p.stage_bytes, alo = ahi + CT_A_BYTES;
                    const uint32_t bw_ = ahi + 2 * CT_A_BYTES;   // rows 0..NJ-1 = hi, NJ..2NJ-1 = lo
                    if (p.cross_fp8) {
                        const uint32_t b8 = bw_ + (uint32_t)NJ * 128;   // the e4m3 image [w | w_lo] takes the place of the lo rows
#pragma unroll
                        for (int kk = 0; kk < CT_KC / 16; ++kk) {
                            umma_f16(d_set, umma_desc(ahi + kk * 32), umma_desc(bw_ + kk * 32), idesc_half, (ks | kk) != 0);
                            umma_f8(d_set + (uint32_t)NJ, umma_desc(alo + kk * 32), umma_desc(b8 + kk * 32), idesc_half, (ks | kk) != 0);
                        }
                    } else if (wide) {
#pragma unroll
                        for (int kk = 0; kk < CT_KC / 16; ++kk) {
                            umma_f16(d_set, umma_desc(ahi + kk * 32), umma_desc(bw_ + kk * 32), idesc_wide, (ks | kk) != 0);
                            umma_f16(d_set + (uint32_t)NJ, umma_desc(alo + kk * 32), umma_desc(bw_ + kk * 32), idesc_half, 1);
                        }
                    } else {
                        const uint32_t blo = bw_ + (uint32_t)NJ * 128;
#pragma unroll
                        for (int kk = 0; kk < CT_KC / 16; ++kk) {
                            umma_f16(d_set, umma_desc(ahi + kk * 32), umma_desc(bw_ + kk * 32), idesc_half, (ks | kk) != 0);
                            umma_f16(d_set + (uint32_t)NJ, umma_desc(alo + kk * 32), umma_desc(bw_ + kk * 32), idesc_half, (ks | kk) != 0);
                            umma_f16(d_set + (uint32_t)NJ, umma_desc(ahi + kk * 32), umma_desc(blo + kk * 32), idesc_half, 1);
                        }
                    }
                    umma_commit(empty(slot));
                }
                umma_commit(d_full(db));
            }
        }
    } else {
        // =========================== epilogue (warps 2..5): lane = voxel along W inside a line ===========================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int r = row;
        const int lw = r % p.bw; r /= p.bw;
        const int lh = r % p.bh; r /= p.bh;
        const int ld = r % p.bd; r /= p.bd;
        const int lb = r;
        const bool first = lw == 0, last = lw == p.bw - 1;   // line ends: the shifted neighbour is the zero padding
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            int64_t t = tile;
            const int iw = (int)(t % nbw); t /= nbw;
            const int ih = (int)(t % nbh); t /= nbh;
            const int id = (int)(t % nbd); t /= nbd;
            const int64_t vox = ((((int64_t)t * p.bb + lb) * p.D + id * p.bd + ld) * p.H + ih * p.bh + lh) * p.W + iw * p.bw + lw;
            float* dst = p.y + vox * p.Cout;
            const int db = it % nbuf;
            mbar_wait(d_full(db), (it / nbuf) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(db * 256);
            for (int n0 = 0; n0 < p.Cout; n0 += 32) {
                float acc[32];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint32_t v[32], u[32];
                    tmem_ld32(taddr + j * p.Cout + n0, v);
                    tmem_ld32(taddr + NJ + j * p.Cout + n0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float z = __uint_as_float(v[i]) + __uint_as_float(u[i]);
                        if (j == 0) {            // Z_0[w-1]
                            z = __shfl_up_sync(0xffffffffu, z, 1);
                            acc[i] = first ? 0.f : z;
                        } else if (j == 1) {     // Z_1[w]
                            acc[i] += z;
                        } else {                 // Z_2[w+1]
                            z = __shfl_down_sync(0xffffffffu, z, 1);
                            acc[i] += last ? 0.f : z;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = acc[j] * p.out_scale; o.y = acc[j + 1] * p.out_scale;
                    o.z = acc[j + 2] * p.out_scale; o.w = acc[j + 3] * p.out_scale;
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(dst + n0 + j) = o;
                }
            }
            tc_fence_before();
            mbar_arrive(d_empty(db));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// two floats -> two e4m3 bytes (first argument in the LOW byte), round to nearest, saturating at +-448
__device__ __forceinline__ uint32_t ct_cvt_e4m3x2(float lo_elem, float hi_elem) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi_elem), "f"(lo_elem));
    return (uint32_t)r;
}
// weight element of channel kc (0..63 inside its chunk) of row r in the e4m3 image: [w(c), w(c+1), w_lo(c), w_lo(c+1)] per pair
__device__ __forceinline__ void store_weight_e4m3(uint8_t* tile, int r, int kc, float w_scaled, float w_lo) {
    const uint32_t b = ct_cvt_e4m3x2(w_scaled * (1.0f / CT_LO8_SCALE), w_lo);   // low byte: w * 2^(s-12), high byte: w_lo * 2^s
    const int j = (kc >> 1) * 4 + (kc & 1);
    tile[ctc::sw128_byte_offset(r, j)] = (uint8_t)(b & 0xffu);
    tile[ctc::sw128_byte_offset(r, j + 2)] = (uint8_t)(b >> 8);
}

// W fp32 [Cout, Cin, 3,3,3] -> [9 (kd,kh)][Cpad/64][hi rows j*Cout+n (j = kw), then lo rows][64 K] fp16 SWIZZLE_128B images
__global__ void pack_conv_weights_dx_kernel(const float* __restrict__ W, int Cout, int Cin, int Cpad, float wscale, int cross_fp8,
                                            uint8_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)27 * Cpad * Cout;
    if (t >= total) return;
    const int n = (int)(t % Cout);
    const int k = (int)((t / Cout) % Cpad);
    const int tap = (int)(t / ((int64_t)Cout * Cpad));   // kd*9 + kh*3 + kw
    float w = 0.f;
    if (k < Cin) w = W[((int64_t)n * Cin + k) * 27 + tap] * wscale;
    w = fminf(fmaxf(w, -65504.f), 65504.f);
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const int nchunk = Cpad / CT_KC;
    const int cc = k / CT_KC, kc = k % CT_KC;
    const int pair = tap / 3, j = tap % 3;
    const int NJ = 3 * Cout;
    uint8_t* piece = out + ((size_t)pair * nchunk + cc) * (size_t)(2 * NJ * 128);
    *reinterpret_cast<__half*>(piece + ctc::sw128_offset(j * Cout + n, kc)) = h;
    if (cross_fp8) store_weight_e4m3(piece + (size_t)NJ * 128, j * Cout + n, kc, w, w - __half2float(h));
    else *reinterpret_cast<__half*>(piece + ctc::sw128_offset(NJ + j * Cout + n, kc)) = l;
}

// x fp32 [rows, C] (rows = B*voxels), scale/shift [B, C] -> xh, xl fp16 [rows, Cpad] (zero padded channels)
// One thread per FOUR channels: 16-byte loads, saturating f16x2 conversions, 8-byte stores (C % 4 == 0 fast path).
__device__ __forceinline__ uint32_t ct_cvt_f16x2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

__global__ void __launch_bounds__(256)
gn_apply_split_vec4_kernel(const float* __restrict__ x, int64_t rows, int64_t vox_per_sample, int C, int Cpad,
                           const float* __restrict__ scale, const float* __restrict__ shift, uint2* __restrict__ xh,
                           uint2* __restrict__ xl, uint32_t* __restrict__ range_flag, int cross_fp8) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per channel quad
    const int quads = Cpad / 4;
    if (t >= rows * quads) return;
    const int64_t r = t / quads;
    const int c = (int)(t - r * quads) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        v = __ldg(reinterpret_cast<const float4*>(x + r * C + c));
        if (scale != nullptr) {
            const int b = (int)(r / vox_per_sample);
            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)b * C + c));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)b * C + c));
            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        // the clamp below would hide an activation outside the fp16 range (or a NaN): record it (gnb_f16_overflow_fetch)
        if (!(fabsf(v.x) <= 65504.f) | !(fabsf(v.y) <= 65504.f) | !(fabsf(v.z) <= 65504.f) | !(fabsf(v.w) <= 65504.f)) *range_flag = 1u;
        v.x = fminf(fmaxf(v.x, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
        v.z = fminf(fmaxf(v.z, -65504.f), 65504.f); v.w = fminf(fmaxf(v.w, -65504.f), 65504.f);
    }
    uint2 h, l;
    h.x = ct_cvt_f16x2_sat(v.x, v.y);
    h.y = ct_cvt_f16x2_sat(v.z, v.w);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    if (cross_fp8) {   // per channel pair: [lo * 2^12, lo * 2^12, a, a] as e4m3 (see g_cross_fp8)
        l.x = ct_cvt_e4m3x2((v.x - f0.x) * CT_LO8_SCALE, (v.y - f0.y) * CT_LO8_SCALE) | (ct_cvt_e4m3x2(v.x, v.y) << 16);
        l.y = ct_cvt_e4m3x2((v.z - f1.x) * CT_LO8_SCALE, (v.w - f1.y) * CT_LO8_SCALE) | (ct_cvt_e4m3x2(v.z, v.w) << 16);
    } else {
        l.x = ct_cvt_f16x2_sat(v.x - f0.x, v.y - f0.y);
        l.y = ct_cvt_f16x2_sat(v.z - f1.x, v.w - f1.y);
    }
    xh[t] = h;
    xl[t] = l;
}

// The same pass on the VIRTUAL tensor cat(skip, nearest_upsample_2x(x_low)) (ref components/unet3d.py:291,325-330: the decoder's
// joining): channel quads below Cs come from skip [B,D,H,W,Cs], the others from x_low [B,D/2,H/2,W/2,Cx] at the parent voxel.
// The concatenated fp32 tensor is never written.
__global__ void __launch_bounds__(256)
gn_apply_split_cat_kernel(const float* __restrict__ skip, int Cs, const float* __restrict__ xlow, int Cx, int64_t rows, int D, int H,
                          int W, int Cpad, const float* __restrict__ scale, const float* __restrict__ shift, uint2* __restrict__ xh,
                          uint2* __restrict__ xl, uint32_t* __restrict__ range_flag, int cross_fp8) {
    // 32-bit index arithmetic (the entry point checks rows * quads < 2^32): the 64-bit divisions of the first version cost
    // more than the memory traffic
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per channel quad
    const unsigned quads = (unsigned)Cpad / 4u;
    const int C = Cs + Cx;
    if ((int64_t)t >= rows * quads) return;
    const unsigned r = t / quads;
    const int c = (int)(t - r * quads) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        const unsigned vox = (unsigned)D * H * W;
        const unsigned b = r / vox;
        if (c < Cs) {
            v = __ldg(reinterpret_cast<const float4*>(skip + (int64_t)r * Cs + c));
        } else {
            const unsigned rv = r - b * vox, row = rv / (unsigned)W, x = rv - row * (unsigned)W, z = row / (unsigned)H, y = row - z * (unsigned)H;
            const unsigned rl = ((b * (unsigned)(D / 2) + (z >> 1)) * (unsigned)(H / 2) + (y >> 1)) * (unsigned)(W / 2) + (x >> 1);
            v = __ldg(reinterpret_cast<const float4*>(xlow + (int64_t)rl * Cx + (c - Cs)));
        }
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)b * C + c));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)b * C + c));
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (!(fabsf(v.x) <= 65504.f) | !(fabsf(v.y) <= 65504.f) | !(fabsf(v.z) <= 65504.f) | !(fabsf(v.w) <= 65504.f)) *range_flag = 1u;
        v.x = fminf(fmaxf(v.x, -65504.f), 65504.f); v.y = fminf(fmaxf(v.y, -65504.f), 65504.f);
        v.z = fminf(fmaxf(v.z, -65504.f), 65504.f); v.w = fminf(fmaxf(v.w, -65504.f), 65504.f);
    }
    uint2 h, l;
    h.x = ct_cvt_f16x2_sat(v.x, v.y);
    h.y = ct_cvt_f16x2_sat(v.z, v.w);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    if (cross_fp8) {   // per channel pair: [lo * 2^12, lo * 2^12, a, a] as e4m3 (see g_cross_fp8)
        l.x = ct_cvt_e4m3x2((v.x - f0.x) * CT_LO8_SCALE, (v.y - f0.y) * CT_LO8_SCALE) | (ct_cvt_e4m3x2(v.x, v.y) << 16);
        l.y = ct_cvt_e4m3x2((v.z - f1.x) * CT_LO8_SCALE, (v.w - f1.y) * CT_LO8_SCALE) | (ct_cvt_e4m3x2(v.z, v.w) << 16);
    } else {
        l.x = ct_cvt_f16x2_sat(v.x - f0.x, v.y - f0.y);
        l.y = ct_cvt_f16x2_sat(v.z - f1.x, v.w - f1.y);
    }
    xh[t] = h;
    xl[t] = l;
}

// generic path (C % 4 != 0 or unaligned pointers): one thread per channel pair
__global__ void __launch_bounds__(256)
gn_apply_split_kernel(const float* __restrict__ x, int64_t rows, int64_t vox_per_sample, int C, int Cpad,
                      const float* __restrict__ scale, const float* __restrict__ shift, __half* __restrict__ xh,
                      __half* __restrict__ xl, uint32_t* __restrict__ range_flag, int cross_fp8) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per channel pair
    const int half_c = Cpad / 2;
    if (t >= rows * half_c) return;
    const int64_t r = t / half_c;
    const int c = (int)(t - r * half_c) * 2;
    float v0 = 0.f, v1 = 0.f;
    if (c < C) {
        const int b = (int)(r / vox_per_sample);
        const float2 xv = *reinterpret_cast<const float2*>(x + r * C + c);
        v0 = xv.x; v1 = xv.y;
        if (scale != nullptr) {
            const float2 sc = *reinterpret_cast<const float2*>(scale + (int64_t)b * C + c);
            const float2 sh = *reinterpret_cast<const float2*>(shift + (int64_t)b * C + c);
            v0 = fmaf(v0, sc.x, sh.x); v1 = fmaf(v1, sc.y, sh.y);
        }
        if (!(fabsf(v0) <= 65504.f) | !(fabsf(v1) <= 65504.f)) *range_flag = 1u;
        v0 = fminf(fmaxf(v0, -65504.f), 65504.f);
        v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
    }
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    *reinterpret_cast<__half2*>(xh + r * Cpad + c) = h;
    if (cross_fp8) {
        *reinterpret_cast<uint32_t*>(xl + r * Cpad + c) =
            ct_cvt_e4m3x2((v0 - hf.x) * CT_LO8_SCALE, (v1 - hf.y) * CT_LO8_SCALE) | (ct_cvt_e4m3x2(v0, v1) << 16);
    } else {
        *reinterpret_cast<__half2*>(xl + r * Cpad + c) = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    }
}

// W fp32 [Cout, Cin, 3,3,3] -> [27][Cpad/64][hi,lo][Cout rows x 64 K] fp16 K-major SWIZZLE_128B images
__global__ void pack_conv_weights_kernel(const float* __restrict__ W, int Cout, int Cin, int Cpad, float wscale, int cross_fp8,
                                         uint8_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)27 * Cpad * Cout;
    if (t >= total) return;
    const int n = (int)(t % Cout);
    const int k = (int)((t / Cout) % Cpad);
    const int tap = (int)(t / ((int64_t)Cout * Cpad));
    float w = 0.f;
    if (k < Cin) w = W[((int64_t)n * Cin + k) * 27 + tap] * wscale;  // [Cout][Cin][kd][kh][kw], tap = kd*9+kh*3+kw
    w = fminf(fmaxf(w, -65504.f), 65504.f);
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const int nchunk = Cpad / CT_KC;
    const int cc = k / CT_KC, kc = k % CT_KC;
    const size_t b_bytes = (size_t)Cout * 128;
    uint8_t* piece = out + ((size_t)tap * nchunk + cc) * 2 * b_bytes;
    const uint32_t off = ctc::sw128_offset(n, kc);
    *reinterpret_cast<__half*>(piece + off) = h;
    if (cross_fp8) store_weight_e4m3(piece + b_bytes, n, kc, w, w - __half2float(h));
    else *reinterpret_cast<__half*>(piece + b_bytes + off) = l;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_conv_tc_set_cross_precision(int32_t mode) {
    GNB_REQUIRE(mode == 0 || mode == 1, "gnb_conv_tc_set_cross_precision: mode must be 0 (fp16 cross terms) or 1 (e4m3 cross terms)");
    g_cross_fp8 = mode;
    return GNB_OK;
}

int32_t gnb_conv_tc_cross_precision(void) { return g_cross_fp8; }

int32_t gnb_conv3d_tc_pack_weights(const float* W, int32_t Cout, int32_t Cin, int32_t scale_log2, void* packed,
                                   void* stream) {
    GNB_REQUIRE(W && packed, "gnb_conv3d_tc_pack_weights: null pointer");
    GNB_REQUIRE(Cout % 32 == 0 && Cout >= 32 && Cout <= 128 && Cin > 0, "gnb_conv3d_tc_pack_weights: Cout must be a multiple of 32 in [32,128]");
    const int Cpad = ceil_div(Cin, CT_KC) * CT_KC;
    const int64_t total = (int64_t)27 * Cpad * Cout;
    pack_conv_weights_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        W, Cout, Cin, Cpad, ldexpf(1.0f, scale_log2), g_cross_fp8, reinterpret_cast<uint8_t*>(packed));
    return check_launch("gnb_conv3d_tc_pack_weights");
}

int32_t gnb_gn_apply_split(const float* x, int32_t B, int64_t voxels, int32_t C, const float* scale, const float* shift,
                           void* xh, void* xl, void* stream) {
    GNB_REQUIRE(x && xh && xl, "gnb_gn_apply_split: null pointer");
    GNB_REQUIRE(C % 2 == 0 && (scale == nullptr) == (shift == nullptr), "gnb_gn_apply_split: bad arguments");
    const int Cpad = ceil_div(C, CT_KC) * CT_KC;
    const int64_t rows = (int64_t)B * voxels;
    if (rows == 0) return GNB_OK;
    uint32_t* flag = f16_flag_ptr();
    GNB_REQUIRE(flag != nullptr, "gnb_gn_apply_split: range flag allocation failed");
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xh) | reinterpret_cast<uintptr_t>(xl) |
                           reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0;
    if (C % 4 == 0 && aligned) {
        gn_apply_split_vec4_kernel<<<(unsigned)ceil_div<int64_t>(rows * (Cpad / 4), 256), 256, 0, as_stream(stream)>>>(
            x, rows, voxels, C, Cpad, scale, shift, reinterpret_cast<uint2*>(xh), reinterpret_cast<uint2*>(xl), flag, g_cross_fp8);
        return check_launch("gnb_gn_apply_split");
    }
    gn_apply_split_kernel<<<(unsigned)ceil_div<int64_t>(rows * (Cpad / 2), 256), 256, 0, as_stream(stream)>>>(
        x, rows, voxels, C, Cpad, scale, shift, reinterpret_cast<__half*>(xh), reinterpret_cast<__half*>(xl), flag, g_cross_fp8);
    return check_launch("gnb_gn_apply_split");
}

int32_t gnb_gn_apply_split_cat(const float* skip, int32_t Cs, const float* x_low, int32_t Cx, int32_t B, int32_t D, int32_t H,
                               int32_t W, const float* scale, const float* shift, void* xh, void* xl, void* stream) {
    GNB_REQUIRE(skip && x_low && scale && shift && xh && xl, "gnb_gn_apply_split_cat: null pointer");
    GNB_REQUIRE(Cs % 4 == 0 && Cx % 4 == 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "gnb_gn_apply_split_cat: bad shape");
    GNB_REQUIRE((int64_t)B * D * H * W * (ceil_div(Cs + Cx, CT_KC) * CT_KC / 4) < (1ll << 32), "gnb_gn_apply_split_cat: more than 2^32 channel quads");
    GNB_REQUIRE(((reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(x_low) | reinterpret_cast<uintptr_t>(xh) |
                  reinterpret_cast<uintptr_t>(xl) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0,
                "gnb_gn_apply_split_cat: pointers must be 16-byte aligned");
    const int C = Cs + Cx, Cpad = ceil_div(C, CT_KC) * CT_KC;
    const int64_t rows = (int64_t)B * D * H * W;
    if (rows == 0) return GNB_OK;
    uint32_t* flag = f16_flag_ptr();
    GNB_REQUIRE(flag != nullptr, "gnb_gn_apply_split_cat: range flag allocation failed");
    gn_apply_split_cat_kernel<<<(unsigned)ceil_div<int64_t>(rows * (Cpad / 4), 256), 256, 0, as_stream(stream)>>>(
        skip, Cs, x_low, Cx, rows, D, H, W, Cpad, scale, shift, reinterpret_cast<uint2*>(xh), reinterpret_cast<uint2*>(xl), flag, g_cross_fp8);
    return check_launch("gnb_gn_apply_split_cat");
}

int32_t gnb_conv3d_tc_supported(int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!(pow2(D) && pow2(H) && pow2(W)) || B < 1) return 0;
    if ((int64_t)B * D * H * W < CT_M) return 0;
    // a 128-voxel tile spans bb samples when a sample has fewer than 128 voxels: the batch must be a multiple of bb
    // (any batch size works for the levels with >= 128 voxels per sample)
    const int64_t vox = (int64_t)D * H * W;
    const int bb = vox >= CT_M ? 1 : (int)(CT_M / vox);
    if (B % bb != 0) return 0;
    if (Cout % 32 != 0 || Cout < 32 || Cout > 128 || Cin < 1) return 0;  // 4 accumulators x Cout columns <= 512
    return 1;
}

int32_t gnb_conv3d_tc(const void* xh, const void* xl, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin,
                      const void* w_packed, int32_t scale_log2, int32_t Cout, int32_t relu, float* y, void* stream) {
    GNB_REQUIRE(xh && xl && w_packed && y, "gnb_conv3d_tc: null pointer");
    GNB_REQUIRE(gnb_conv3d_tc_supported(B, D, H, W, Cin, Cout), "gnb_conv3d_tc: unsupported shape B=%d D=%d H=%d W=%d Cin=%d Cout=%d", B, D, H, W, Cin, Cout);
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("gnb_conv3d_tc: cuTensorMapEncodeTiled is not available from this driver"); return GNB_ERR_CUDA; }
    ConvTcParams p;
    p.B = B; p.D = D; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu;
    p.out_scale = ldexpf(1.0f, -scale_log2);
    p.cross_fp8 = g_cross_fp8;
    p.acc_stride = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128);
    p.nbuf = (CT_NACC * p.acc_stride * 2 <= 512) ? 2 : 1;
    p.Cpad = ceil_div(Cin, CT_KC) * CT_KC;
    p.nchunk = p.Cpad / CT_KC;
    int rem = CT_M;
    p.bw = W < rem ? W : rem; rem /= p.bw;
    p.bh = H < rem ? H : rem; rem /= p.bh;
    p.bd = D < rem ? D : rem; rem /= p.bd;
    p.bb = rem;
    GNB_REQUIRE(p.bb <= B && B % p.bb == 0, "gnb_conv3d_tc: batch too small for a 128-voxel tile");
    p.b_bytes = Cout * 128;
    p.stage_bytes = 2 * CT_A_BYTES + 2 * p.b_bytes;
    p.nstages = (220 * 1024) / p.stage_bytes;
    if (p.nstages > CT_MAX_STAGES) p.nstages = CT_MAX_STAGES;
    p.w_packed = reinterpret_cast<const uint8_t*>(w_packed);
    p.y = y;
    p.num_tiles = (int64_t)B * D * H * W / CT_M;

    CUtensorMap maps[2];
    const cuuint64_t gdim[5] = {(cuuint64_t)p.Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)p.Cpad * 2, (cuuint64_t)W * p.Cpad * 2, (cuuint64_t)H * W * p.Cpad * 2,
                                (cuuint64_t)D * H * W * p.Cpad * 2};
    const cuuint32_t box[5] = {(cuuint32_t)CT_KC, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, (cuuint32_t)p.bb};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* srcs[2] = {xh, xl};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(srcs[i]), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("gnb_conv3d_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return GNB_ERR_CUDA; }
    }
    const int smem = p.nstages * p.stage_bytes + 1024;
    GNB_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = sm_count();
    if ((int64_t)grid > p.num_tiles) grid = (int)p.num_tiles;
    conv_tc_kernel<<<grid, CT_THREADS, smem, as_stream(stream)>>>(maps[0], maps[1], p);
    return check_launch("gnb_conv3d_tc");
}

int32_t gnb_conv3d_tc_dx_supported(int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout) {
    if (!gnb_conv3d_tc_supported(B, D, H, W, Cin, Cout)) return 0;
    // a tile row must be a whole W line (the shift along W is a warp shuffle) and the stacked accumulators must fit
    return ((Cout == 32 || Cout == 64) && (W == 32 || W == 16 || W == 8) && (int64_t)W <= CT_M) ? 1 : 0;
}

int32_t gnb_conv3d_tc_dx_pack_weights(const float* W, int32_t Cout, int32_t Cin, int32_t scale_log2, void* packed,
                                      void* stream) {
    GNB_REQUIRE(W && packed, "gnb_conv3d_tc_dx_pack_weights: null pointer");
    GNB_REQUIRE((Cout == 32 || Cout == 64) && Cin > 0, "gnb_conv3d_tc_dx_pack_weights: Cout must be 32 or 64");
    const int Cpad = ceil_div(Cin, CT_KC) * CT_KC;
    const int64_t total = (int64_t)27 * Cpad * Cout;
    pack_conv_weights_dx_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, as_stream(stream)>>>(
        W, Cout, Cin, Cpad, ldexpf(1.0f, scale_log2), g_cross_fp8, reinterpret_cast<uint8_t*>(packed));
    return check_launch("gnb_conv3d_tc_dx_pack_weights");
}

int32_t gnb_conv3d_tc_dx(const void* xh, const void* xl, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin,
                         const void* w_packed, int32_t scale_log2, int32_t Cout, int32_t relu, float* y, void* stream) {
    GNB_REQUIRE(xh && xl && w_packed && y, "gnb_conv3d_tc_dx: null pointer");
    GNB_REQUIRE(gnb_conv3d_tc_dx_supported(B, D, H, W, Cin, Cout), "gnb_conv3d_tc_dx: unsupported shape B=%d D=%d H=%d W=%d Cin=%d Cout=%d", B, D, H, W, Cin, Cout);
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("gnb_conv3d_tc_dx: cuTensorMapEncodeTiled is not available from this driver"); return GNB_ERR_CUDA; }
    ConvDxParams p;
    p.B = B; p.D = D; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu;
    p.out_scale = ldexpf(1.0f, -scale_log2);
    p.cross_fp8 = g_cross_fp8;
    p.Cpad = ceil_div(Cin, CT_KC) * CT_KC;
    p.nchunk = p.Cpad / CT_KC;
    int rem = CT_M;
    p.bw = W; rem /= p.bw;
    p.bh = H < rem ? H : rem; rem /= p.bh;
    p.bd = D < rem ? D : rem; rem /= p.bd;
    p.bb = rem;
    GNB_REQUIRE(p.bb <= B && B % p.bb == 0, "gnb_conv3d_tc_dx: batch too small for a 128-voxel tile");
    p.b_bytes = 2 * 3 * Cout * 128;
    p.stage_bytes = 2 * CT_A_BYTES + p.b_bytes;
    p.nstages = (220 * 1024) / p.stage_bytes;
    if (p.nstages > CT_MAX_STAGES) p.nstages = CT_MAX_STAGES;
    p.w_packed = reinterpret_cast<const uint8_t*>(w_packed);
    p.y = y;
    p.num_tiles = (int64_t)B * D * H * W / CT_M;
    CUtensorMap maps[2];
    const cuuint64_t gdim[5] = {(cuuint64_t)p.Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t gstr[4] = {(cuuint64_t)p.Cpad * 2, (cuuint64_t)W * p.Cpad * 2, (cuuint64_t)H * W * p.Cpad * 2,
                                (cuuint64_t)D * H * W * p.Cpad * 2};
    const cuuint32_t box[5] = {(cuuint32_t)CT_KC, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, (cuuint32_t)p.bb};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const void* srcs[2] = {xh, xl};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(srcs[i]), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("gnb_conv3d_tc_dx: cuTensorMapEncodeTiled failed (%d)", (int)r); return GNB_ERR_CUDA; }
    }
    const int smem = p.nstages * p.stage_bytes + 1024;
    GNB_CUDA(cudaFuncSetAttribute(conv_tc_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = sm_count();
    if ((int64_t)grid > p.num_tiles) grid = (int)p.num_tiles;
    conv_tc_dx_kernel<<<grid, CT_THREADS, smem, as_stream(stream)>>>(maps[0], maps[1], p);
    return check_launch("gnb_conv3d_tc_dx");
}

}  // extern "C"
