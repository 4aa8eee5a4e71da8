// Dense lattice decode, second generation (ref predict.py:145-158 + networks/conv_implicit_wnf.py:128-149).
//
// Same math as decode_tc_kernel<COUT, 1> (decode_tc.cu) -- H1 = BN1(ReLU(trilinear(U))) over the implicit 128^3 lattice,
// Linear2 on tcgen05 with fp16 hi/lo split operands, Linear3 folded into the epilogue -- restructured around what the
// first-generation kernel was actually bound by (ncu: tensor pipe 25 % active, producers stalled on L2 gathers, and
// 387 KB of L2->SM traffic per 128-row tile against a ~42 B/clk/SM L2 port):
//
//   * a work item is a PAIR of adjacent lattice lines (i, 2jp) and (i, 2jp+1): 2 x 128 rows, two TMEM accumulators
//     (2 x 256 columns = all of TMEM).  Every W2 piece streamed from L2 feeds both tiles (B traffic halves to 128 KB
//     per tile) and the two lines share their four (x, y) corner columns three times out of four (gather traffic
//     131 -> ~74 KB per tile).
//   * the pipeline is K-chunk major: a two-slot ring of 64 KB A chunks (2 tiles x {hi, lo} x 128 rows x 64 channels);
//     all eight producer warps fill chunk c+1 while the tensor core consumes chunk c, so the producers have the full
//     MMA time of a chunk (24 MMAs) instead of a quarter of a tile.
//   * producer mapping: an 8-lane group owns ONE D-cell of the feature grid for both lines and 8 consecutive channels
//     per lane: 16-byte gathers (16 per lane per chunk, all issued before the slot wait), separable x/y blend in
//     registers, one 16-byte swizzled shared-memory store per row and precision part.
//   * BN1 is folded into W2 / b2 on the host (it follows the ReLU, so it is linear in front of Linear2): the A operand
//     is ReLU(interp) only; the power-of-two weight scale is folded into b2 and W3 so the epilogue is add / max / fma.
//
// Warp roles (448 threads): 0-7 A producers, 8-11 epilogue (TMEM lane quarter = warp % 4), 12 MMA issuer, 13 B loader.
#include "tc_common.cuh"
#include <stdlib.h>

namespace gnb {

namespace dl2 {

constexpr int K = 256, N = 256, M = 128, KCHUNK = 64, NCHUNK = K / KCHUNK;
constexpr int PART_BYTES = M * KCHUNK * 2;          // 16 KB: one tile, one precision part, one K chunk
constexpr int SLOT_BYTES = 4 * PART_BYTES;          // 64 KB: {tile0 hi, tile0 lo, tile1 hi, tile1 lo}
constexpr int A_SLOTS = 2;
constexpr int B_PIECE_BYTES = N * KCHUNK * 2;       // 32 KB (one CTA); a CTA pair keeps one 16 KB half (128 of the 256 columns) per CTA
constexpr int B_SLOTS = 3;                          // single CTA: 3 x 32 KB; CTA pair: 6 x 16 KB in the same 96 KB
constexpr int THREADS = 448;
constexpr int MAX_G = 32;   // one cell pair per 16-lane producer group
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
constexpr uint32_t IDESC_PAIR = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * M) >> 4) << 24);   // M = 256 over two CTAs

struct Smem {
    static constexpr int a = 0;                                   // [2 slots][64 KB]
    static constexpr int b_ring = a + A_SLOTS * SLOT_BYTES;       // [3][32 KB]
    static constexpr int bars = b_ring + B_SLOTS * B_PIECE_BYTES;
    static constexpr int n_bars = 8 + 3 * (2 * B_SLOTS);   // a_full/a_empty[2], b_full/b_empty/b_peer[<=6], d_full/d_empty[2]
    static constexpr int tmem_ptr = bars + n_bars * 8;
    static constexpr int rowtab = tmem_ptr + 16;                  // [128] {float w0, float w1}: axis_cell weights of every lattice index
    static constexpr int kstart = rowtab + M * 8;                 // [G+1]
    static constexpr int pairs = kstart + (MAX_G + 2) * 4;        // [Q/2][2] u8: the two lines of every pair
    static constexpr int axtab = pairs + M;                       // [128] u8: axis_cell lower cell index of every lattice index
    static constexpr int total = axtab + M;
};
static_assert(Smem::total + 1024 <= 227 * 1024, "shared memory budget");

// Epilogue constants (b2s[256] | w3s[COUT][256]) live in constant memory: the epilogue warps consume them as immediate
// c[bank][offset] operands of FADD / FFMA -- no load instructions, no L1 traffic competing with the producers' gathers.
// Filled per launch by a stream-ordered device-to-device copy from the prep kernel's scratch (one stream at a time).
__constant__ float c_epi[4 * 256];

// (make PROFILE_KNOBS=1 PROFILE_NO_CLOCKS=1 keeps the knock-out switches but leaves the clock reads out: they slow the producers
// down by ~60 %, which makes every knock-out look producer-bound)
#if defined(GNB_PROFILE_KNOBS) && !defined(GNB_PROFILE_NO_CLOCKS)
#define GNB_PROFILE_CLOCKS
#endif
// Per-role wait-time attribution (profiling builds only, tools/decode_bench.py): cycles one thread of each role spends in
// its barrier waits.  [cta][16]: 0 MMA a_full | 1 MMA W2 | 2 MMA d_empty | 3 MMA total | 4 producer a_empty | 5 producer total |
// 6 producer row/store phase | 7 epilogue d_full | 8 epilogue total | 9 loader b_empty | 10 loader total |
// 11 producer x-blend (waits for the prefetched gathers) | 12 producer prefetch issue | 13 producer y-blend | 14 producer fence + arrive
#ifdef GNB_PROFILE_CLOCKS
__device__ unsigned long long g_prof[1024 * 16];
#define DL2_PROF_DECL unsigned long long prof_acc[6] = {0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64()
#define DL2_PROF(i, stmt) do { const long long t_ = clock64(); stmt; prof_acc[i] += (unsigned long long)(clock64() - t_); } while (0)
#define DL2_PROF_STORE(i, slot) g_prof[blockIdx.x * 16 + (slot)] = prof_acc[i]
#define DL2_PROF_TOTAL(slot) g_prof[blockIdx.x * 16 + (slot)] = (unsigned long long)(clock64() - prof_t0)
#else
#define DL2_PROF_DECL
#define DL2_PROF(i, stmt) do { stmt; } while (0)
#define DL2_PROF_STORE(i, slot)
#define DL2_PROF_TOTAL(slot)
#endif

struct Params {
    const float* U;            // [B,G,G,G,256] hoisted grid (Linear1 applied on the feature grid), fp32 channels-last
    int B, G, Q;
    const uint8_t* w2_packed;  // (W2 * diag(bn1_scale)) * 2^s, gnb_pack_f16_split layout
    const float* b2s;          // [256] (b2 + W2 bn1_shift) * 2^s          (scratch; the kernel reads the copy in c_epi)
    const float* w3s;          // [COUT][256] W3 * bn2_scale * 2^-s        (scratch; the kernel reads the copy in c_epi)
    const float* tail;         // [COUT][4] {c0, bn3_scale, bn3_shift, 0}
    float* out;                // [B, Q^3, COUT]
    int64_t num_pairs;         // B * Q * Q / 2
    int dbg;                   // profiling aid (GNB_DL2_DBG): 1 producers idle, 2 no W2 copies, 4 epilogue math skipped, 8 no A stores, 16 no MMAs
};

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 issue two fp32 operations per instruction) ---------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// two fp32 -> packed fp16x2 (first argument in the low half), saturating to +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// relu + fp16 hi/lo split of two fp32 values (ascending channel order inside the 32-bit words).  The ReLU rides on the
// conversions: hi = relu(x) TRUNCATED to fp16 (cvt.rz.relu), so the remainder x - hi is >= 0 for x >= 0 and equals x < 0
// otherwise, and lo = cvt.rn.relu(remainder) is the rounded remainder or 0 -- 5 instructions per channel pair instead of 7
// (no FMNMX).  Truncating hi costs one bit: hi + lo carries 21 significant bits (2^-22 relative) instead of 22.
__device__ __forceinline__ void relu_split2(float2 h, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(h.y), "f"(h.x));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 r = sub2(h, hf);
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}

// lattice index -> feature-grid cell and weights along one axis (same fp32 arithmetic as gnb_trilinear_sample_grid)
__device__ __forceinline__ void axis_cell(int idx, float sq, int G, int& c0, int& c1, float& w0, float& w1) {
    const float g = __fsub_rn(__fmul_rn(2.0f, __fmul_rn((float)idx, sq)), 1.0f);
    const float f = fminf((float)(G - 1), fmaxf(((g + 1.f) / 2.f) * (float)(G - 1), 0.f));
    const float fl = floorf(f);
    c0 = (int)fl;
    c1 = c0 + 1 < G ? c0 + 1 : G - 1;
    w1 = f - fl;
    w0 = (fl + 1.f) - f;
}

// Warp roles (448 threads): 0-7 A producers | 8-11 epilogue (TMEM lane quarter = warp % 4) | 12 MMA issuer | 13 B loader.
// Issue priority on sm_100 goes to the HIGHEST warp id: the MMA issuer and the loader (latency critical, a handful of
// instructions per MMA) come first, then the epilogue (it gates the reuse of the accumulators), and the producers fill
// the remaining issue slots.  The epilogue sleeps on its barrier instead of spinning through those slots.
// PAIR: the kernel runs as clusters of two CTAs (one TPC).  Each CTA keeps its own pair of lattice lines (A operand,
// accumulators, producers, epilogue exactly as in the single-CTA form) but only HALF of every W2 piece; the leader CTA
// issues tcgen05.mma.cta_group::2 (M = 256: both CTAs' tiles at once, N = 256 read as 128 columns from either CTA), so the
// per-SM shared-memory operand traffic of an MMA drops from 12 KB to 8 KB and the W2 stream per SM halves -- the port the
// producers' stores compete for.  Cross-CTA hand-shakes: the peer's producers / epilogue arrive remotely on the leader's
// a_full / d_empty barriers, the leader's commits are multicast to both CTAs' a_empty / b_empty / d_full barriers, and the
// peer reports the landing of its W2 halves on the leader's b_peer barriers.
// NPW: A-producer warps.  8 = 16 lanes x 4 channels per cell pair (one warp covers a 64-channel chunk of two cell pairs);
// 16 = the same cell pairs with 2 channels per lane, warps 8-15 taking the upper 32 channels of the chunk (80 registers per
// thread).  Measured: no gain from the extra warps, see g_producer_warps.
template <int COUT, bool PAIR, int NPW>
__global__ void __launch_bounds__(NPW == 16 ? 768 : THREADS, 1)
decode_lattice_kernel(const Params p) {
    constexpr int CPL = NPW == 16 ? 2 : 4;        // channels per producer lane
    constexpr int NH = CPL / 2;                   // packed channel pairs per lane
    constexpr int W_MMA = NPW + 4, W_LOAD = NPW + 5;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    constexpr int BS = PAIR ? 2 * B_SLOTS : B_SLOTS;                  // W2 ring slots
    constexpr int BPB = PAIR ? B_PIECE_BYTES / 2 : B_PIECE_BYTES;     // bytes of a W2 piece held by this CTA
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;              // 0 = leader (issues the MMAs)
    const uint32_t bar0 = sbase + Smem::bars;
    auto a_full = [&](int s) { return bar0 + 8 * s; };
    auto a_empty = [&](int s) { return bar0 + 8 * (2 + s); };
    auto b_full = [&](int s) { return bar0 + 8 * (4 + s); };
    auto b_empty = [&](int s) { return bar0 + 8 * (4 + BS + s); };
    auto d_full = [&](int t) { return bar0 + 8 * (4 + 2 * BS + t); };
    auto d_empty = [&](int t) { return bar0 + 8 * (6 + 2 * BS + t); };
    auto b_peer = [&](int s) { return bar0 + 8 * (8 + 2 * BS + s); };   // leader only: the peer's half of piece s has landed
    // arrive on a barrier of the LEADER (local for the leader itself)
    auto arrive_leader = [&](uint32_t bar) {
        if (PAIR && rank != 0) mbar_arrive_remote(mapa_cluster(bar, 0));
        else mbar_arrive(bar);
    };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Smem::tmem_ptr);
    const int Q = p.Q, G = p.G, QH = Q >> 1;
    const float sq = __fdiv_rn(1.0f, (float)(Q - 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_SLOTS; ++s) { mbar_init(a_full(s), PAIR ? 2 * NPW : NPW); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < BS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); mbar_init(b_peer(s), 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(d_full(t), 1); mbar_init(d_empty(t), PAIR ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_LOAD && lane == 0) {
        // pairing of the Q lattice lines of one i-plane: two lines of the same y-cell share their corner columns, so
        // pairs are formed inside the runs of equal y-cell; the few left-over lines are paired with each other
        uint8_t* pt = smem + Smem::pairs;
        uint8_t left[M];
        int np = 0, nl = 0, j = 0;
        while (j < Q) {
            int c0, c1, e = j;
            float w0, w1;
            axis_cell(j, sq, G, c0, c1, w0, w1);
            for (;;) {
                int d0, d1;
                if (e >= Q) break;
                axis_cell(e, sq, G, d0, d1, w0, w1);
                if (d0 != c0) break;
                ++e;
            }
            for (; j + 1 < e; j += 2) { pt[2 * np] = (uint8_t)j; pt[2 * np + 1] = (uint8_t)(j + 1); ++np; }
            if (j < e) left[nl++] = (uint8_t)j++;
        }
        for (int a = 0; a + 1 < nl; a += 2) { pt[2 * np] = left[a]; pt[2 * np + 1] = left[a + 1]; ++np; }
    }
    if (threadIdx.x < M) {
        // per-row blend along D (independent of the tile)
        const int k = threadIdx.x;
        int z0, z1;
        float w0, w1;
        axis_cell(k, sq, G, z0, z1, w0, w1);
        reinterpret_cast<float2*>(smem + Smem::rowtab)[k] = make_float2(w0, w1);
        (smem + Smem::axtab)[k] = (uint8_t)z0;
    }
    if (warp == W_MMA) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Smem::tmem_ptr), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Smem::tmem_ptr), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anybody arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x <= G) {
        // kstart[d] = first lattice row whose D-cell index is >= d (rows are monotone in k); kstart[G] = 128
        int k = 0;
        for (; k < M; ++k) {
            int z0, z1;
            float w0, w1;
            axis_cell(k, sq, G, z0, z1, w0, w1);
            if (z0 >= (int)threadIdx.x) break;
        }
        reinterpret_cast<int*>(smem + Smem::kstart)[threadIdx.x] = k;
    }
    __syncthreads();
    const uint8_t* pairs = smem + Smem::pairs;

    if (warp < NPW) {
        // =========================== A producers ===========================
        // A 16-lane group owns two adjacent D-cells (three slices) of both lines and 4 consecutive channels per lane.
        // The producers are the critical path of the kernel (the MMA issuer waited on a_full for a third of the run), and each
        // SM sub-partition runs only two of these warps, so the code is written for instruction-level parallelism: item
        // decoding by shifts (Q = 128), per-axis cell / weight table in shared memory, corner pointers per pair, and two rows
        // (eight independent blend -> ReLU -> split chains) computed before their eight stores.
        const int pw = warp & 7;
        const int cp = pw * 2 + (lane >> 4);      // cell pair: cells 2cp, 2cp+1
        const int l16 = lane & 15;
        const int cho = (warp >> 3) * 32 + l16 * CPL;   // first channel of this lane inside a 64-channel chunk
        const bool active = 2 * cp < G;           // G < 32: the surplus groups only pace the barriers
        const int dA = 2 * cp < G ? 2 * cp : G - 1;
        const uint32_t sd = (uint32_t)G * G * K;  // elements per D-slice (<= 2^18)
        const uint32_t soff[3] = {(uint32_t)dA * sd, (uint32_t)(dA + 1 < G ? dA + 1 : G - 1) * sd, (uint32_t)(dA + 2 < G ? dA + 2 : G - 1) * sd};
        // byte offset of this lane's 8-byte store inside a 128-byte swizzled row: 16-byte unit (l16 >> 1) ^ (k & 7)
        const uint32_t unit = (uint32_t)(cho >> 3), sub8 = (uint32_t)(cho & 7) * 2u;
        const uint32_t axtab = sbase + Smem::axtab, rowtab = sbase + Smem::rowtab, kstab = sbase + Smem::kstart;
        // read-only tables: plain (non-volatile) shared loads, free to be hoisted and interleaved by the compiler
        auto lds2 = [](uint32_t addr) { uint2 v; asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; };
        auto lds1 = [](uint32_t addr) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
        auto ldsb = [](uint32_t addr) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; };
        // axis_cell of lattice index idx from the tables: {w0, w1, c0, c1}
        auto axis = [&](int idx) {
            const uint2 w = lds2(rowtab + idx * 8);
            const uint32_t c0 = ldsb(axtab + idx);
            return make_uint4(w.x, w.y, c0, c0 + 1 < (uint32_t)G ? c0 + 1 : (uint32_t)G - 1);
        };
        const int k_lo[2] = {(int)lds1(kstab + 4 * (2 * cp < G ? 2 * cp : G)), (int)lds1(kstab + 4 * (2 * cp + 1 < G ? 2 * cp + 1 : G))};
        const int k_hi[2] = {k_lo[1], (int)lds1(kstab + 4 * (2 * cp + 2 < G ? 2 * cp + 2 : G))};

        struct Item { int b, i, j0, j1; };
        auto item_of = [&](uint32_t pr) {          // Q = 128: pr = (b * 128 + i) * 64 + jp
            Item it;
            const uint32_t jp = pr & (uint32_t)(M / 2 - 1);
            it.i = (int)((pr >> 6) & (uint32_t)(M - 1));
            it.b = (int)(pr >> 13);
            const uint32_t pj = lds1(sbase + Smem::pairs + (jp >> 1) * 4);   // four u8 per word: pairs 2w, 2w+1
            it.j0 = (int)((pj >> ((jp & 1) * 16)) & 0xffu);
            it.j1 = (int)((pj >> ((jp & 1) * 16 + 8)) & 0xffu);
            return it;
        };
        // the four (y, x) corner columns of lattice line (i, j) of sample b, at this lane's channels of chunk 0
        auto corners = [&](int b, int i, int j, const float* (&cptr)[4]) {
            const uint4 ax = axis(i), ay = axis(j);
            const float* ub = p.U + (size_t)b * G * sd + cho;
            cptr[0] = ub + (ay.z * (uint32_t)G + ax.z) * (uint32_t)K;
            cptr[1] = ub + (ay.z * (uint32_t)G + ax.w) * (uint32_t)K;
            cptr[2] = ub + (ay.w * (uint32_t)G + ax.z) * (uint32_t)K;
            cptr[3] = ub + (ay.w * (uint32_t)G + ax.w) * (uint32_t)K;
        };
        // gathers of one line for chunk c: [slice][y][x] 16-byte loads
        // this lane's CPL channels at p as NH packed pairs
        auto ldv = [](const float* ptr, float2 (&v)[NH]) {
            if (CPL == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(ptr));
                v[0] = make_float2(t.x, t.y);
                v[NH - 1] = make_float2(t.z, t.w);
            } else {
                v[0] = __ldg(reinterpret_cast<const float2*>(ptr));
            }
        };
        auto issue = [&](const float* const (&cptr)[4], int c, float2 (&r)[3][2][2][NH]) {
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const uint32_t o = soff[s] + (uint32_t)c * KCHUNK;
                ldv(cptr[0] + o, r[s][0][0]);
                ldv(cptr[1] + o, r[s][0][1]);
                ldv(cptr[2] + o, r[s][1][0]);
                ldv(cptr[3] + o, r[s][1][1]);
            }
        };

        float2 nxt[3][2][2][NH];
        DL2_PROF_DECL;
        uint32_t q = 0;  // running chunk counter (slot = q & 1)
        const uint32_t npairs = (uint32_t)p.num_pairs;
        uint32_t pair = blockIdx.x;
        const float* cA[4];   // corner pointers of line j0 of the current pair
        if (pair < npairs) {
            const Item it0 = item_of(pair);
            corners(it0.b, it0.i, it0.j0, cA);
            if (active && !(p.dbg & 1)) issue(cA, 0, nxt);
        }
        for (; pair < npairs; pair += gridDim.x) {
            const Item it = item_of(pair);
            const uint4 ax = axis(it.i), aya = axis(it.j0), ayb = axis(it.j1);
            const float wx0 = __uint_as_float(ax.x), wx1 = __uint_as_float(ax.y);
            const float wy0[2] = {__uint_as_float(aya.x), __uint_as_float(ayb.x)}, wy1[2] = {__uint_as_float(aya.y), __uint_as_float(ayb.y)};
            const bool same_y = aya.z == ayb.z;
            const bool has_next = pair + gridDim.x < npairs;

#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c, ++q) {
                const int slot = q & 1;
                if (active && !(p.dbg & 1)) {
#ifdef GNB_PROFILE_CLOCKS
                    const long long t_a = clock64();
#endif
                    // 1. x-blend of the prefetched corners: X[s][y] (4 channels as two packed pairs)
                    float2 X[3][2][NH];
                    const float2 vx0 = make_float2(wx0, wx0), vx1 = make_float2(wx1, wx1);
#pragma unroll
                    for (int s = 0; s < 3; ++s)
#pragma unroll
                        for (int yy = 0; yy < 2; ++yy)
#pragma unroll
                            for (int h2 = 0; h2 < NH; ++h2) X[s][yy][h2] = fma2(nxt[s][yy][1][h2], vx1, mul2(nxt[s][yy][0][h2], vx0));
#ifdef GNB_PROFILE_CLOCKS
                    // the clock read must follow the blend: make it depend on one blended value
                    long long t_b;
                    asm volatile("{ .reg .f32 t; mov.f32 t, %1; mov.u64 %0, %%clock64; }" : "=l"(t_b) : "f"(X[2][1][NH - 1].y) : "memory");
                    prof_acc[2] += (unsigned long long)(t_b - t_a);
#endif
                    // 2. prefetch the corners of the next chunk (or of the next pair's first chunk)
                    if (c + 1 < NCHUNK) issue(cA, c + 1, nxt);
                    else if (has_next) {   // cA is free from here on: it becomes the next pair's line j0
                        const Item ni = item_of(pair + gridDim.x);
                        corners(ni.b, ni.i, ni.j0, cA);
                        issue(cA, 0, nxt);
                    }
#ifdef GNB_PROFILE_CLOCKS
                    const long long t_c = clock64();
                    prof_acc[3] += (unsigned long long)(t_c - t_b);
#endif
                    // 3. (x, y)-blended slices of both lines: P[l][s] (4 channels as two packed pairs)
                    float2 P[2][3][NH];
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        const float2 vy0 = make_float2(wy0[l], wy0[l]), vy1 = make_float2(wy1[l], wy1[l]);
                        if (l == 0 || same_y) {
#pragma unroll
                            for (int s = 0; s < 3; ++s)
#pragma unroll
                                for (int h2 = 0; h2 < NH; ++h2) P[l][s][h2] = fma2(X[s][1][h2], vy1, mul2(X[s][0][h2], vy0));
                        } else {
                            // line 1 lies in another y-cell (left-over lines, ~3 % of the pairs): its own gathers
                            const float* cB[4];
                            corners(it.b, it.i, it.j1, cB);
#pragma unroll
                            for (int s = 0; s < 3; ++s) {
                                const uint32_t o = soff[s] + (uint32_t)c * KCHUNK;
                                float2 a00[NH], a10[NH], a01[NH], a11[NH];
                                ldv(cB[0] + o, a00); ldv(cB[1] + o, a10); ldv(cB[2] + o, a01); ldv(cB[3] + o, a11);
#pragma unroll
                                for (int h2 = 0; h2 < NH; ++h2) {
                                    const float2 xa = fma2(a10[h2], vx1, mul2(a00[h2], vx0)), xb = fma2(a11[h2], vx1, mul2(a01[h2], vx0));
                                    P[l][s][h2] = fma2(xb, vy1, mul2(xa, vy0));
                                }
                            }
                        }
                    }
#ifdef GNB_PROFILE_CLOCKS
                    long long t_d;
                    asm volatile("{ .reg .f32 t; mov.f32 t, %1; mov.u64 %0, %%clock64; }" : "=l"(t_d) : "f"(P[1][2][NH - 1].y) : "memory");
                    prof_acc[4] += (unsigned long long)(t_d - t_c);
#endif
                    // 4. the slot must have been consumed by the tensor core (chunk q - 2)
                    DL2_PROF(0, mbar_wait(a_empty(slot), ((q >> 1) & 1) ^ 1));
#ifdef GNB_PROFILE_CLOCKS
                    const long long t_rows = clock64();
#endif
                    const uint32_t a_addr = sbase + Smem::a + slot * SLOT_BYTES + sub8;  // + part * PART_BYTES + row offset
#pragma unroll
                    for (int cell = 0; cell < 2; ++cell) {
                        if (2 * cp + cell >= G) break;
                        float2 sl[2][NH];
#pragma unroll
                        for (int l = 0; l < 2; ++l)
#pragma unroll
                            for (int h2 = 0; h2 < NH; ++h2) sl[l][h2] = sub2(P[l][cell + 1][h2], P[l][cell][h2]);
                        // one row: z-blend, ReLU, fp16 hi/lo split of this lane's 4 channels of both lines -> v[l] = {hi0, hi1, lo0, lo1}
                        auto row_values = [&](const float wz, uint4 (&v)[2]) {
                            const float2 vz = make_float2(wz, wz);
#pragma unroll
                            for (int l = 0; l < 2; ++l) {
                                relu_split2(fma2(sl[l][0], vz, P[l][cell][0]), v[l].x, v[l].z);
                                if (NH == 2) relu_split2(fma2(sl[l][NH - 1], vz, P[l][cell][NH - 1]), v[l].y, v[l].w);
                            }
                        };
                        auto row_store = [&](const int k, const uint4 (&v)[2]) {   // row k: 8-row group k >> 3, 128-byte row k & 7, swizzled unit
                            const uint32_t addr = a_addr + (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u + ((unit ^ (uint32_t)(k & 7)) << 4);
#ifdef GNB_PROFILE_KNOBS
                            if (p.dbg & 8) return;   // knock-out: everything but the shared-memory stores
#endif
#pragma unroll
                            for (int l = 0; l < 2; ++l) {
                                if (NH == 2) {
                                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + (2 * l) * PART_BYTES), "r"(v[l].x), "r"(v[l].y) : "memory");
                                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + (2 * l + 1) * PART_BYTES), "r"(v[l].z), "r"(v[l].w) : "memory");
                                } else {
                                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr + (2 * l) * PART_BYTES), "r"(v[l].x) : "memory");
                                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr + (2 * l + 1) * PART_BYTES), "r"(v[l].z) : "memory");
                                }
                            }
                        };
                        const int k_end = k_hi[cell];
                        int k = k_lo[cell];
#pragma unroll 1
                        for (; k + 1 < k_end; k += 2) {
                            const float wz0 = __uint_as_float(lds1(rowtab + k * 8 + 4)), wz1 = __uint_as_float(lds1(rowtab + k * 8 + 12));
                            uint4 v0[2], v1[2];
                            row_values(wz0, v0);
                            row_values(wz1, v1);
                            row_store(k, v0);
                            row_store(k + 1, v1);
                        }
                        if (k < k_end) {
                            const float wz0 = __uint_as_float(lds1(rowtab + k * 8 + 4));
                            uint4 v0[2];
                            row_values(wz0, v0);
                            row_store(k, v0);
                        }
                    }
#ifdef GNB_PROFILE_CLOCKS
                    prof_acc[1] += (unsigned long long)(clock64() - t_rows);
#endif
                } else {
                    // no D-cell (G < 32): still pace on the slot, or an early arrival for the NEXT use of this slot would be
                    // counted into the current phase of a_full
                    mbar_wait(a_empty(slot), ((q >> 1) & 1) ^ 1);
                }
#ifdef GNB_PROFILE_CLOCKS
                const long long t_e = clock64();
#endif
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) arrive_leader(a_full(slot));
#ifdef GNB_PROFILE_CLOCKS
                prof_acc[5] += (unsigned long long)(clock64() - t_e);
#endif
            }
        }
        if (threadIdx.x == 0) {
            DL2_PROF_STORE(0, 4); DL2_PROF_STORE(1, 6); DL2_PROF_TOTAL(5);
            DL2_PROF_STORE(2, 11); DL2_PROF_STORE(3, 12); DL2_PROF_STORE(4, 13); DL2_PROF_STORE(5, 14);
        }
    } else if (warp < NPW + 4) {
        // =========================== epilogue ===========================
        const int qd = warp & 3;  // TMEM lane quarter this warp may access (warp id mod 4)
        const int row = qd * 32 + lane;
        float c_tail[COUT], bn3s[COUT], bn3h[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) { c_tail[o] = p.tail[o * 4]; bn3s[o] = p.tail[o * 4 + 1]; bn3h[o] = p.tail[o * 4 + 2]; }
        int it = 0;
        DL2_PROF_DECL;
        for (int64_t pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x, ++it) {
            const int jp = (int)(pair % QH);
            const int64_t plane = pair / QH;  // b * Q + i
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                DL2_PROF(0, mbar_wait_sleep(d_full(t), it & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(t * N);
                // software pipeline: the next EC columns are in flight (tcgen05.ld) while the current EC are folded.  The
                // column loop is fully unrolled so that every b2s / w3s value is an immediate constant-bank operand.
                constexpr int EC = NPW == 16 ? 16 : 32;   // 16 producer warps leave 80 registers per thread
                uint32_t r0[EC], r1[EC];
                float dsum[COUT][4];
#pragma unroll
                for (int o = 0; o < COUT; ++o)
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) dsum[o][q4] = 0.f;
                tmem_ldn(taddr, r0);
#pragma unroll
                for (int n0 = 0; n0 < N; n0 += 2 * EC) {
                    if ((p.dbg & 4) && n0 >= 64) break;
                    tmem_ld_wait();
                    tmem_ldn(taddr + n0 + EC, r1);
#pragma unroll
                    for (int u = 0; u < EC; ++u) {
                        const float v = fmaxf(__uint_as_float(r0[u]) + c_epi[n0 + u], 0.f);
#pragma unroll
                        for (int o = 0; o < COUT; ++o) dsum[o][u & 3] = fmaf(v, c_epi[(1 + o) * N + n0 + u], dsum[o][u & 3]);
                    }
                    tmem_ld_wait();
                    if (n0 + 2 * EC < N && !((p.dbg & 4) && n0 + 2 * EC >= 64)) {
                        tmem_ldn(taddr + n0 + 2 * EC, r0);
                    } else {
                        // every TMEM read of accumulator t has completed: the next pair's MMAs may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(d_empty(t));
                    }
#pragma unroll
                    for (int u = 0; u < EC; ++u) {
                        const float v = fmaxf(__uint_as_float(r1[u]) + c_epi[n0 + EC + u], 0.f);
#pragma unroll
                        for (int o = 0; o < COUT; ++o) dsum[o][u & 3] = fmaf(v, c_epi[(1 + o) * N + n0 + EC + u], dsum[o][u & 3]);
                    }
                }
                float2 dot[COUT][2];
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    dot[o][0] = make_float2(dsum[o][0], dsum[o][1]);
                    dot[o][1] = make_float2(dsum[o][2], dsum[o][3]);
                }
                const int64_t grow = (plane * Q + pairs[2 * jp + t]) * M + row;
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    const float2 dd = add2(dot[o][0], dot[o][1]);
                    p.out[grow * COUT + o] = fmaxf((dd.x + dd.y) + c_tail[o], 0.f) * bn3s[o] + bn3h[o];
                }
            }
        }
        if (warp == NPW && lane == 0) { DL2_PROF_STORE(0, 7); DL2_PROF_TOTAL(8); }
    } else if (warp == W_MMA) {
        // =========================== MMA issuer (leader CTA) / W2 landing notifier (peer CTA) ===========================
        if (lane == 0 && rank == 0) {
            int it = 0;
            uint32_t piece = 0, q = 0;
            DL2_PROF_DECL;
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
#ifdef GNB_PROFILE_KNOBS
                if (p.dbg & 16) return;      // knock-out: no MMAs (the commits arrive at once)
#endif
                if (PAIR) umma_f16_pair(d, da, db, IDESC_PAIR, acc);
                else umma_f16(d, da, db, IDESC, acc);
            };
            auto commit = [&](uint32_t bar) {
                if (PAIR) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            // barriers the peer CTA arrives on need a cluster-scope acquire
            auto wait_x = [&](uint32_t bar, uint32_t parity) {
                if (PAIR) mbar_wait_cluster(bar, parity);
                else mbar_wait(bar, parity);   // latency-critical: plain polling (the suspend-hint form brought nothing measurable)
            };
            auto wait_b = [&](uint32_t pc) {   // W2 piece pc (this CTA's part and, in a pair, the peer's)
                DL2_PROF(1, mbar_wait(b_full(pc % BS), (pc / BS) & 1); if (PAIR) mbar_wait_cluster(b_peer(pc % BS), (pc / BS) & 1));
                tc_fence_after();
            };
            for (int64_t pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x, ++it) {
                for (int c = 0; c < NCHUNK; ++c, ++q) {
                    const int slot = q & 1;
                    DL2_PROF(0, wait_x(a_full(slot), (q >> 1) & 1));
                    const uint32_t a_slot = sbase + Smem::a + slot * SLOT_BYTES;
                    const int s_hi = piece % BS, s_lo = (piece + 1) % BS;
                    const uint32_t b_hi = sbase + Smem::b_ring + s_hi * BPB, b_lo = sbase + Smem::b_ring + s_lo * BPB;
                    // hi*hi + lo*hi of tile t against the W2_hi piece; hi*lo against the W2_lo piece
                    // descriptors of the first K-step; the next K-steps add 32 bytes = 2 address units (no carry: the tiles
                    // are 1024-byte aligned and far below the 14-bit address field's wrap)
                    const uint64_t dbh = umma_desc(b_hi), dbl = umma_desc(b_lo);
                    auto mma_hi = [&](int t) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(t * N);
                        const uint64_t dah = umma_desc(a_slot + (2 * t) * PART_BYTES), dal = umma_desc(a_slot + (2 * t + 1) * PART_BYTES);
#pragma unroll
                        for (int kk = 0; kk < KCHUNK / 16; ++kk) mma(d_tmem, dah + 2 * kk, dbh + 2 * kk, (c | kk) != 0);
#pragma unroll
                        for (int kk = 0; kk < KCHUNK / 16; ++kk) mma(d_tmem, dal + 2 * kk, dbh + 2 * kk, 1);
                    };
                    auto mma_lo = [&](int t) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(t * N);
                        const uint64_t dah = umma_desc(a_slot + (2 * t) * PART_BYTES);
#pragma unroll
                        for (int kk = 0; kk < KCHUNK / 16; ++kk) mma(d_tmem, dah + 2 * kk, dbl + 2 * kk, 1);
                    };
                    wait_b(piece);
                    if (c == 0 || c == NCHUNK - 1) {
                        // tile-major: accumulator 0 is released to / taken from the epilogue a whole tile (12 MMAs) before
                        // accumulator 1, so draining one tile overlaps the other tile's MMAs
                        if (c == 0) { DL2_PROF(2, wait_x(d_empty(0), (it & 1) ^ 1)); tc_fence_after(); }
                        mma_hi(0);
                        wait_b(piece + 1);
                        mma_lo(0);
                        if (c == NCHUNK - 1) commit(d_full(0));
                        if (c == 0) { DL2_PROF(2, wait_x(d_empty(1), (it & 1) ^ 1)); tc_fence_after(); }
                        mma_hi(1);
                        commit(b_empty(s_hi));
                        mma_lo(1);
                        commit(b_empty(s_lo));
                        commit(a_empty(slot));
                        if (c == NCHUNK - 1) commit(d_full(1));
                    } else {
                        mma_hi(0);
                        mma_hi(1);
                        commit(b_empty(s_hi));
                        wait_b(piece + 1);
                        mma_lo(0);
                        mma_lo(1);
                        commit(b_empty(s_lo));
                        commit(a_empty(slot));  // every MMA that reads this A slot has been issued
                    }
                    piece += 2;
                }
            }
            DL2_PROF_STORE(0, 0); DL2_PROF_STORE(1, 1); DL2_PROF_STORE(2, 2); DL2_PROF_TOTAL(3);
        } else if (PAIR && lane == 0 && rank != 0) {
            // peer CTA: tell the leader when each of this CTA's W2 halves has landed (in order)
            uint32_t piece = 0;
            for (int64_t pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x)
                for (int pc = 0; pc < 2 * NCHUNK; ++pc, ++piece) {
                    mbar_wait(b_full(piece % BS), (piece / BS) & 1);
                    mbar_arrive_remote(mapa_cluster(b_peer(piece % BS), 0));
                }
        }
    } else {
        // =========================== B loader ===========================
        if (lane == 0) {
            uint32_t piece = 0;
            DL2_PROF_DECL;
            for (int64_t pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x) {
                for (int pc = 0; pc < 2 * NCHUNK; ++pc, ++piece) {
                    const int slot = piece % BS;
                    // (spinning: neither the suspend-hint form of try_wait nor a 64 ns back-off between probes changed the kernel time
                    // beyond run-to-run noise in same-run A/B measurements, although polling shows up as ~6 % of the shared-memory pipe)
                    DL2_PROF(0, mbar_wait(b_empty(slot), ((piece / BS) & 1) ^ 1));
                    if (p.dbg & 2) { mbar_arrive(b_full(slot)); continue; }
                    mbar_expect_tx(b_full(slot), BPB);
                    // a CTA pair splits every piece by output column: rows [rank * 128, rank * 128 + 128) of the K-major image
                    bulk_g2s(sbase + Smem::b_ring + slot * BPB, p.w2_packed + (size_t)pc * B_PIECE_BYTES + (size_t)rank * BPB, BPB,
                             b_full(slot));
                }
            }
            DL2_PROF_STORE(0, 9); DL2_PROF_TOTAL(10);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // both CTAs have drained their accumulators before the pair's TMEM is released
    if (warp == W_MMA) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// prep: b2s = (b2 + W2 bn1_shift) * 2^s ; w3s[o][n] = W3[o][n] * bn2_scale[n] * 2^-s ;
//       tail[o] = {sum_n bn2_shift[n] W3[o][n] + b3[o], bn3_scale[o], bn3_shift[o], 0}
__global__ void lattice_prep_kernel(const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ bn1_shift,
                                    const float* __restrict__ W3, const float* __restrict__ b3, const float* __restrict__ s2,
                                    const float* __restrict__ h2, const float* __restrict__ s3, const float* __restrict__ h3,
                                    int cout, float wscale, float* __restrict__ b2s, float* __restrict__ w3s,
                                    float* __restrict__ tail) {
    const int n = threadIdx.x;  // 256 threads
    __shared__ float red[N];
    if (blockIdx.x == 0) {
        float acc = b2[n];
        for (int k = 0; k < K; ++k) acc = fmaf(W2[n * K + k], bn1_shift[k], acc);
        b2s[n] = acc * wscale;
        return;
    }
    const int o = blockIdx.x - 1;
    const float w = W3[o * N + n];
    w3s[o * N + n] = w * (s2 ? s2[n] : 1.f) / wscale;
    red[n] = (h2 ? h2[n] : 0.f) * w;
    __syncthreads();
    for (int s = N / 2; s > 0; s >>= 1) {
        if (n < s) red[n] += red[n + s];
        __syncthreads();
    }
    if (n == 0) {
        tail[o * 4 + 0] = red[0] + (b3 ? b3[o] : 0.f);
        tail[o * 4 + 1] = s3 ? s3[o] : 1.f;
        tail[o * 4 + 2] = h3 ? h3[o] : 0.f;
        tail[o * 4 + 3] = 0.f;
    }
}

// CTA pairs need an even grid and an even number of line pairs (both CTAs of a cluster run the same number of items)
// The CTA-pair form (tcgen05.mma.cta_group::2) is correct and tested but measured SLOWER than one CTA per SM on B200
// (22.6 vs 19.8 ms at batch 32: the A producers are the critical path, and lock-stepping two CTAs per chunk adds their
// variance), so the single-CTA kernel is the default; gnb_decode_lattice_set_mode(1) selects the pair form.
static bool g_use_pair = false;
static bool g_want_pair = false;
// 16 producer warps (2 channels per lane) are bit-identical and were measured SLOWER than 8 (17.7 vs 17.1 ms): the row phase of
// the producers did not shrink with twice the warps -- it is bound by the shared-memory store path it shares with the MMA
// operand reads, not by per-warp latency.  gnb_decode_lattice_set_mode(2) selects the 16-warp form.
static int g_producer_warps = 8;

template <int COUT>
static int32_t launch(const Params& p, cudaStream_t st) {
    const int smem = Smem::total + 1024;
    int grid = sm_count();
    if ((int64_t)grid > p.num_pairs) grid = (int)p.num_pairs;
    if (g_use_pair && grid >= 2 && (p.num_pairs % 2) == 0) {
        grid &= ~1;
        GNB_CUDA(cudaFuncSetAttribute(decode_lattice_kernel<COUT, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        GNB_CUDA(cudaLaunchKernelEx(&cfg, decode_lattice_kernel<COUT, true, 8>, p));
        return check_launch("gnb_decode_lattice");
    }
    if (g_producer_warps == 16) {
        GNB_CUDA(cudaFuncSetAttribute(decode_lattice_kernel<COUT, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        decode_lattice_kernel<COUT, false, 16><<<grid, (16 + 6) * 32, smem, st>>>(p);
        return check_launch("gnb_decode_lattice");
    }
    GNB_CUDA(cudaFuncSetAttribute(decode_lattice_kernel<COUT, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    decode_lattice_kernel<COUT, false, 8><<<grid, THREADS, smem, st>>>(p);
    return check_launch("gnb_decode_lattice");
}

}  // namespace dl2
}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_decode_lattice(const float* U, int32_t B, int32_t G, int32_t Q, const float* W2, const void* w2f_packed,
                           int32_t w2f_scale_log2, const float* b2, const float* bn1_shift, const float* bn2_scale,
                           const float* bn2_shift, const float* W3, const float* b3, const float* bn3_scale,
                           const float* bn3_shift, int32_t Cout, float* scratch, float* out, void* stream) {
    GNB_REQUIRE(U && W2 && w2f_packed && b2 && bn1_shift && W3 && scratch && out, "gnb_decode_lattice: null pointer");
    GNB_REQUIRE(Cout >= 1 && Cout <= 3, "gnb_decode_lattice: Cout must be 1..3 (got %d)", Cout);
    GNB_REQUIRE(Q == dl2::M, "gnb_decode_lattice: volume_size must be 128 (one lattice line per 128-row tile)");
    GNB_REQUIRE(B > 0 && G >= 2 && G <= dl2::MAX_G, "gnb_decode_lattice: need B > 0 and a feature grid 2 <= G <= 32");
    GNB_REQUIRE((reinterpret_cast<uintptr_t>(U) & 15) == 0, "gnb_decode_lattice: U must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    float* b2s = scratch;                 // [256]
    float* w3s = scratch + dl2::N;        // [3][256]
    float* tail = scratch + 4 * dl2::N;   // [3][4]
    dl2::lattice_prep_kernel<<<1 + Cout, dl2::N, 0, st>>>(W2, b2, bn1_shift, W3, b3, bn2_scale, bn2_shift, bn3_scale, bn3_shift,
                                                         Cout, ldexpf(1.0f, w2f_scale_log2), b2s, w3s, tail);
    // b2s (256 floats) and w3s (Cout x 256) are contiguous in the scratch: one stream-ordered copy into constant memory
    // (guarded: another stream launching this family waits for this kernel before it refreshes the bank)
    ConstBankGuard guard(BANK_LATTICE, st);
    GNB_CUDA(cudaMemcpyToSymbolAsync(dl2::c_epi, b2s, sizeof(float) * dl2::N * (1 + Cout), 0, cudaMemcpyDeviceToDevice, st));
    dl2::Params p;
    p.U = U; p.B = B; p.G = G; p.Q = Q;
    p.w2_packed = reinterpret_cast<const uint8_t*>(w2f_packed);
    p.b2s = b2s; p.w3s = w3s; p.tail = tail; p.out = out;
    p.num_pairs = (int64_t)B * Q * (Q / 2);
    p.dbg = profile_knob("GNB_DL2_DBG");
    dl2::g_use_pair = dl2::g_want_pair;
    if (Cout == 1) return dl2::launch<1>(p, st);
    if (Cout == 2) return dl2::launch<2>(p, st);
    return dl2::launch<3>(p, st);
}

#ifdef GNB_PROFILE_CLOCKS
// profiling builds only (not in the header): copy the per-CTA wait-time table of the last launch to the host
__attribute__((visibility("default"))) int32_t gnb_prof_decode_lattice_read(unsigned long long* host_out, int32_t n) {
    return cudaMemcpyFromSymbol(host_out, dl2::g_prof, sizeof(unsigned long long) * (size_t)n) == cudaSuccess ? 0 : -1;
}
#endif

int32_t gnb_decode_lattice_set_mode(int32_t cta_pair) {
    dl2::g_want_pair = cta_pair == 1;
    dl2::g_producer_warps = cta_pair == 2 ? 16 : 8;
    return GNB_OK;
}

}  // extern "C"
