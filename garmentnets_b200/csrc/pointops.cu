// Point-set kernels of the PointNet++ stage: farthest point sampling, ball query, kNN + inverse-distance
// interpolation, PointConv grouping and segment max.  These are integer/index-heavy, latency- or L2-bound
// kernels: coalesced loads, shared-memory staging of a cloud, warp-ballot compaction and warp-shuffle reductions.
//
// Reference semantics restated (file:line relative to the reference repository):
//   components/pointnet2.py:26      fps(pos, batch, ratio)                       -> gnb_fps
//   components/pointnet2.py:28-29   radius(pos, pos[idx], r, ..., 64)           -> gnb_ball_query
//   components/pointnet2.py:72      knn_interpolate(x, pos, pos_skip, ..., k)   -> gnb_knn + gnb_knn_interpolate
//   components/pointnet2.py:30-31   PointConv(nn)(x, (pos, pos[idx]), edges)    -> gnb_pointconv_* + gnb_segment_max
#include "common.cuh"
#include <math_constants.h>

namespace gnb {

// ------------------------------------------------------------------------------------------------
// Farthest point sampling: one CTA per cloud, strictly sequential rounds.  The cloud lives in shared
// memory (centroid broadcast) and, on the fast path, each thread keeps its PPT points and their running
// min-distances in registers.  One barrier per round: per-warp (dist,index) keys are double-buffered and
// every warp redundantly reduces the 32 partials.
// key = dist_bits << 32 | (0xFFFFFFFF - index): max key == max distance, ties -> lowest index.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// Same result as warp_max_u64 (every lane gets the maximum key) in two hardware warp reductions (REDUX) instead of five
// 64-bit shuffle rounds: max of the distance bits, then max of the inverted index among the lanes that hold that distance.
__device__ __forceinline__ unsigned long long warp_argmax_key(unsigned long long v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long sub_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int T, int PPT>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr, const int64_t* __restrict__ start,
           const int64_t* __restrict__ out_ptr, int64_t* __restrict__ out, int stride) {
    extern __shared__ float sm[];
    __shared__ unsigned long long wbest[2][32];
    constexpr int NW = T / 32;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = ptr[b];
    const int n = (int)(ptr[b + 1] - p0);
    const int64_t o0 = out_ptr[b];
    const int m = (int)(out_ptr[b + 1] - o0);
    if (n <= 0 || m <= 0) return;
    float* sx = sm;
    float* sy = sm + stride;
    float* sz = sm + 2 * stride;
    float* sd = sm + 3 * stride;  // only used on the generic path (PPT == 0)
    for (int i = tid; i < n; i += T) {
        sx[i] = pos[(p0 + i) * 3 + 0];
        sy[i] = pos[(p0 + i) * 3 + 1];
        sz[i] = pos[(p0 + i) * 3 + 2];
        if (PPT == 0) sd[i] = CUDART_INF_F;
    }
    int cur = 0;
    if (start != nullptr) {
        long long s = start[b];
        cur = (int)(s < 0 ? 0 : (s >= n ? n - 1 : s));
    }
    if (tid == 0) out[o0] = p0 + cur;
    __syncthreads();

    float px[PPT > 0 ? PPT : 1], py[PPT > 0 ? PPT : 1], pz[PPT > 0 ? PPT : 1], pd[PPT > 0 ? PPT : 1];
    if (PPT > 0) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            int i = tid + j * T;
            bool ok = i < n;
            px[j] = ok ? sx[i] : 0.f;
            py[j] = ok ? sy[i] : 0.f;
            pz[j] = ok ? sz[i] : 0.f;
            pd[j] = ok ? CUDART_INF_F : -1.f;  // -1 never wins (real distances are >= 0)
        }
    }
    // the thread's points as register PAIRS for the packed fp32x2 arithmetic of the loop below
    constexpr int H = PPT > 0 ? (PPT + 1) / 2 : 1;
    unsigned long long qx[H], qy[H], qz[H];
    if (PPT > 0) {
#pragma unroll
        for (int j = 0; j + 1 < PPT; j += 2) {
            qx[j >> 1] = pack_f32x2(px[j], px[j + 1]);
            qy[j >> 1] = pack_f32x2(py[j], py[j + 1]);
            qz[j >> 1] = pack_f32x2(pz[j], pz[j + 1]);
        }
    }
    for (int s = 1; s < m; ++s) {
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        unsigned long long best = 0ull;
        if (PPT > 0) {
            // the running (distance, slot) pair stays in two 32-bit registers; the 64-bit key is built once per iteration.
            // Strict > with ascending j keeps the smallest index among equal distances, as the key order does.
            float bd = -1.f;   // -1 never wins (real distances are >= 0; padding slots hold -1)
            int bj = 0;
            // two points per packed fp32x2 instruction (sub / mul / add, each individually rounded: the same bits as
            // sqdist_nofma); the loop is bound by instruction issue on the one SM a cloud occupies
            const unsigned long long c2x = pack_f32x2(cx, cx), c2y = pack_f32x2(cy, cy), c2z = pack_f32x2(cz, cz);
#pragma unroll
            for (int j = 0; j < PPT; j += 2) {
                const unsigned long long dx = sub_f32x2(qx[j >> 1], c2x), dy = sub_f32x2(qy[j >> 1], c2y), dz = sub_f32x2(qz[j >> 1], c2z);
                // the sums stay scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (CUDA 12.9, also with
                // -fmad=false), which would round differently from the oracle
                const unsigned long long xx = mul_f32x2(dx, dx), yy = mul_f32x2(dy, dy), zz = mul_f32x2(dz, dz);
                const float d0 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)xx), __uint_as_float((unsigned)yy)), __uint_as_float((unsigned)zz));
                const float d1 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)(xx >> 32)), __uint_as_float((unsigned)(yy >> 32))),
                                           __uint_as_float((unsigned)(zz >> 32)));
                const float dm0 = fminf(pd[j], d0);
                const float dm1 = fminf(pd[j + 1], d1);
                pd[j] = dm0;
                pd[j + 1] = dm1;
                if (dm0 > bd) { bd = dm0; bj = j; }
                if (dm1 > bd) { bd = dm1; bj = j + 1; }
            }
            if (bd >= 0.f)
                best = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(tid + bj * T));
        } else {
            for (int i = tid; i < n; i += T) {
                float d = sqdist_nofma(sx[i], sy[i], sz[i], cx, cy, cz);
                float dm = fminf(sd[i], d);
                sd[i] = dm;
                unsigned long long key = ((unsigned long long)__float_as_uint(dm) << 32) |
                                         (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                best = key > best ? key : best;
            }
        }
        best = warp_argmax_key(best);
        // (measured alternatives for this second level, profiles/fps_variants_r02.txt: one shared 64-bit atomicMax per warp --
        // a CAS loop in SASS -- 1144 us, every thread folding the per-warp keys itself 1344 us, against 854 us for this)
        if (lane == 0) wbest[s & 1][warp] = best;
        __syncthreads();
        unsigned long long k2 = lane < NW ? wbest[s & 1][lane] : 0ull;
        k2 = warp_argmax_key(k2);
        cur = (int)(0xFFFFFFFFu - (unsigned)(k2 & 0xFFFFFFFFull));
        if (tid == 0) out[o0 + s] = p0 + cur;
    }
}

// ------------------------------------------------------------------------------------------------
// Ball query: one warp per query, the cloud is scanned in index order 32 points at a time; a ballot +
// popc prefix gives each in-radius point its ordered slot, so the first K hits in index order are kept.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ ptr, int B, int64_t i) {
    int lo = 0, hi = B;  // ptr[lo] <= i < ptr[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// (Four queries per warp sharing the candidate loads were measured slower, 490 vs 455 us for SA1: the loop is bound by
// the distance / ballot / rank arithmetic per query, not by the position loads.)
__global__ void __launch_bounds__(256)
ball_query_kernel(const float* __restrict__ x, const float* __restrict__ y, const int64_t* __restrict__ ptr_x,
                  const int64_t* __restrict__ ptr_y, int B, int64_t sumM, float r2, int K,
                  int64_t* __restrict__ nbr, int32_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= sumM) return;
    const int b = find_segment(ptr_y, B, q);
    const float qx = y[q * 3], qy = y[q * 3 + 1], qz = y[q * 3 + 2];
    const int64_t s = ptr_x[b], e = ptr_x[b + 1];
    int count = 0;
    for (int64_t base = s; base < e && count < K; base += 32) {
        const int64_t j = base + lane;
        bool hit = false;
        if (j < e) {
            float d = sqdist_nofma(x[j * 3], x[j * 3 + 1], x[j * 3 + 2], qx, qy, qz);
            hit = d < r2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        if (hit && count + rank < K) nbr[q * K + count + rank] = j;
        count += __popc(mask);
    }
    count = count < K ? count : K;
    for (int t = count + lane; t < K; t += 32) nbr[q * K + t] = -1;
    if (lane == 0) cnt[q] = count;
}

__global__ void radius_pairs_kernel(const int64_t* __restrict__ nbr, const int32_t* __restrict__ cnt,
                                    const int64_t* __restrict__ offs, int64_t sumM, int K,
                                    int64_t* __restrict__ row, int64_t* __restrict__ col) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= sumM) return;
    const int c = cnt[q];
    const int64_t o = offs[q];
    for (int t = lane; t < c; t += 32) {
        row[o + t] = q;
        col[o + t] = nbr[q * K + t];
    }
}

// single-CTA exclusive scan (inputs here are at most a few 10^4 per-centroid counts): every thread owns 16 consecutive
// elements per pass (four 16-byte loads, serial prefix in registers), so 65536 counts take four passes of one block-wide
// scan each instead of sixty-four (65 -> 9 us)
constexpr int SCAN_PT = 16;
__global__ void __launch_bounds__(1024)
scan_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    __shared__ long long wsum[32];
    __shared__ long long chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool vec = (reinterpret_cast<uintptr_t>(in) & 15) == 0, vec_out = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    long long carry = 0;  // identical in every thread
    for (int64_t base = 0; base < n; base += 1024 * SCAN_PT) {
        const int64_t i0 = base + (int64_t)tid * SCAN_PT;
        int v[SCAN_PT];
        if (vec && i0 + SCAN_PT <= n) {
#pragma unroll
            for (int j = 0; j < SCAN_PT; j += 4) {
                const int4 t = __ldg(reinterpret_cast<const int4*>(in + i0 + j));
                v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_PT; ++j) v[j] = i0 + j < n ? in[i0 + j] : 0;
        }
        long long pre[SCAN_PT], sum = 0;   // exclusive prefix inside the thread's run
#pragma unroll
        for (int j = 0; j < SCAN_PT; ++j) { pre[j] = sum; sum += (long long)v[j]; }
        long long incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = wsum[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;  // exclusive offset of each warp
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        const long long off = carry + wsum[warp] + incl - sum;
        if (vec_out && i0 + SCAN_PT <= n) {
#pragma unroll
            for (int j = 0; j < SCAN_PT; j += 2)   // i0 and j are even: 16-byte stores
                *reinterpret_cast<longlong2*>(out + i0 + j) = make_longlong2(off + pre[j], off + pre[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_PT; ++j)
                if (i0 + j < n) out[i0 + j] = off + pre[j];
        }
        carry += chunk_total;
        __syncthreads();  // wsum / chunk_total are rewritten by the next chunk
    }
    if (tid == 0) out[n] = carry;
}

// ------------------------------------------------------------------------------------------------
// kNN: one thread per query, sorted insertion over the cloud in index order (strict <, so equal
// distances keep the lower index first).
// ------------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(128)
knn_kernel(const float* __restrict__ x, const float* __restrict__ y, const int64_t* __restrict__ ptr_x,
           const int64_t* __restrict__ ptr_y, int B, int64_t Ny, int k, int64_t* __restrict__ idx,
           float* __restrict__ d2out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Ny) return;
    const int b = find_segment(ptr_y, B, q);
    const float qx = y[q * 3], qy = y[q * 3 + 1], qz = y[q * 3 + 2];
    float bd[KT];
    int bi[KT];   // index inside the cloud (clouds are far below 2^31 points); -1 = empty slot
#pragma unroll
    for (int p = 0; p < KT; ++p) { bd[p] = CUDART_INF_F; bi[p] = -1; }
    const int64_t s = ptr_x[b], e = ptr_x[b + 1];
    const int n = (int)(e - s);
    const float* __restrict__ xs = x + s * 3;
    auto insert = [&](float cd, int cj) {
        if (cd < bd[KT - 1] || KT > k) {
            bool ins = false;
#pragma unroll
            for (int p = 0; p < KT; ++p) {
                if (p < k) {
                    const bool sw = ins || (cd < bd[p]);
                    if (sw) {
                        const float td = bd[p]; bd[p] = cd; cd = td;
                        const int tj = bi[p]; bi[p] = cj; cj = tj;
                        ins = true;
                    }
                }
            }
        }
    };
    // four candidates per round: their twelve loads are in flight together (the loop was bound by the load -> distance ->
    // compare chain of one candidate at a time); insertion stays in index order, so ties keep the lower index
    int j = 0;
    for (; j + 3 < n; j += 4) {
        float c[12];
#pragma unroll
        for (int u = 0; u < 12; ++u) c[u] = __ldg(xs + j * 3 + u);
        float d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) d[u] = sqdist_nofma(c[3 * u], c[3 * u + 1], c[3 * u + 2], qx, qy, qz);
#pragma unroll
        for (int u = 0; u < 4; ++u) insert(d[u], j + u);
    }
    for (; j < n; ++j) insert(sqdist_nofma(__ldg(xs + j * 3), __ldg(xs + j * 3 + 1), __ldg(xs + j * 3 + 2), qx, qy, qz), j);
#pragma unroll
    for (int p = 0; p < KT; ++p) {
        if (p < k) {
            idx[q * k + p] = bi[p] < 0 ? -1 : s + bi[p];
            d2out[q * k + p] = bd[p];
        }
    }
}

// inverse-distance interpolation, one warp per output row, channels across lanes.  The k neighbour indices and weights are
// computed once per row (not once per 32 channels: the division dominated the first version); VEC: four channels per lane
// with 16-byte loads and stores.  Same operation order per channel as the oracle: sum_p(f_p * w_p) / sum_p(w_p), p ascending.
template <int KT, bool VEC>
__global__ void __launch_bounds__(256)
knn_interp_kernel(const float* __restrict__ feat, int64_t ldf, const int64_t* __restrict__ idx,
                  const float* __restrict__ d2, int64_t Ny, int k, int C, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Ny) return;
    int64_t jj[KT];
    float ww[KT];
    float den = 0.f;
    bool first = true;
#pragma unroll
    for (int p = 0; p < KT; ++p) {
        jj[p] = -1;
        ww[p] = 0.f;
        if (p < k) {
            jj[p] = idx[q * k + p];
            if (jj[p] >= 0) {
                ww[p] = __fdiv_rn(1.0f, fmaxf(d2[q * k + p], 1e-16f));
                den = first ? ww[p] : __fadd_rn(den, ww[p]);
                first = false;
            }
        }
    }
    if (VEC) {
        for (int c = lane * 4; c < C; c += 128) {
            float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
            bool f0 = true;
#pragma unroll
            for (int p = 0; p < KT; ++p) {
                if (jj[p] < 0) continue;
                const float4 f = __ldg(reinterpret_cast<const float4*>(feat + jj[p] * ldf + c));
                const float4 t = make_float4(__fmul_rn(f.x, ww[p]), __fmul_rn(f.y, ww[p]), __fmul_rn(f.z, ww[p]), __fmul_rn(f.w, ww[p]));
                if (f0) { num = t; f0 = false; }
                else { num.x = __fadd_rn(num.x, t.x); num.y = __fadd_rn(num.y, t.y); num.z = __fadd_rn(num.z, t.z); num.w = __fadd_rn(num.w, t.w); }
            }
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!first) o = make_float4(__fdiv_rn(num.x, den), __fdiv_rn(num.y, den), __fdiv_rn(num.z, den), __fdiv_rn(num.w, den));
            *reinterpret_cast<float4*>(out + q * ldo + c) = o;
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            float num = 0.f;
            bool f0 = true;
#pragma unroll
            for (int p = 0; p < KT; ++p) {
                if (jj[p] < 0) continue;
                const float t = __fmul_rn(feat[jj[p] * ldf + c], ww[p]);
                num = f0 ? t : __fadd_rn(num, t);
                f0 = false;
            }
            out[q * ldo + c] = first ? 0.f : __fdiv_rn(num, den);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// PointConv grouping
// ------------------------------------------------------------------------------------------------
__global__ void pointconv_edge_count_kernel(const int64_t* __restrict__ nbr, const int32_t* __restrict__ cnt,
                                            int64_t sumM, int K, int32_t* __restrict__ ecnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sumM) return;
    const int c = cnt[i];
    bool has_self = false;
    for (int t = 0; t < c; ++t) has_self |= (nbr[i * K + t] == i);
    ecnt[i] = c + (has_self ? 0 : 1);
}

__global__ void __launch_bounds__(256)
pointconv_gather_kernel(const float* __restrict__ xf, int64_t ldx, int Cin, const float* __restrict__ pos_x,
                        const float* __restrict__ pos_y, const int64_t* __restrict__ nbr,
                        const int32_t* __restrict__ cnt, const int64_t* __restrict__ eoffs, int64_t sumM, int K,
                        float* __restrict__ edge, int64_t lde) {
    // One warp per centroid.  The edge list (ball-query hits minus j == i, plus the self loop appended last) is compacted
    // with ballots, 32 candidate slots per round, so the index loads are coalesced and independent; narrow rows
    // (Cin <= 8, SA1) are then written one edge per lane, wide rows (SA2) are copied one row per warp iteration, four
    // rows in flight.
    __shared__ int64_t jlist[8][72];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (i >= sumM) return;
    const int c = cnt[i];
    const int64_t e0 = eoffs[i];
    const int ne = (int)(eoffs[i + 1] - e0);
    const float cx = pos_y[i * 3], cy = pos_y[i * 3 + 1], cz = pos_y[i * 3 + 2];
    const bool narrow = Cin <= 8;
    int base_w = 0;
    for (int t0 = 0; t0 <= c; t0 += 32) {
        const int t = t0 + lane;
        int64_t j = i;                       // t == c: the self loop added by PointConv (flat point index i, see header)
        if (t < c) j = nbr[i * K + t];
        const bool keep = t <= c && !(t < c && j == i);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        const int w = base_w + __popc(mask & ((1u << lane) - 1u));
        if (keep && w < ne) {
            if (narrow) {
                float* dst = edge + (e0 + w) * lde;
                for (int ch = 0; ch < Cin; ++ch) dst[ch] = xf[j * ldx + ch];
                dst[Cin + 0] = __fsub_rn(pos_x[j * 3 + 0], cx);
                dst[Cin + 1] = __fsub_rn(pos_x[j * 3 + 1], cy);
                dst[Cin + 2] = __fsub_rn(pos_x[j * 3 + 2], cz);
            } else if (w < 72) {
                jlist[wib][w] = j;
            }
        }
        base_w += __popc(mask);
    }
    if (narrow) return;
    __syncwarp();
    const int n = ne < base_w ? ne : base_w;
    for (int w0 = 0; w0 < n; w0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int w = w0 + u;
            if (w >= n) break;
            const int64_t j = jlist[wib][w];
            float* dst = edge + (e0 + w) * lde;
            for (int ch = lane; ch < Cin; ch += 32) dst[ch] = __ldg(xf + j * ldx + ch);
            if (lane < 3) {
                const float pc = lane == 0 ? cx : (lane == 1 ? cy : cz);
                dst[Cin + lane] = __fsub_rn(__ldg(pos_x + j * 3 + lane), pc);
            }
        }
    }
}

// one warp per (segment, block of 128 channels): each lane owns 4 consecutive channels (16-byte loads when the rows
// are 16-byte aligned), four rows in flight.  Empty segments produce 0 (torch_scatter's fill for 'max').
template <bool VEC>
__global__ void __launch_bounds__(256)
segment_max_kernel(const float* __restrict__ rows, int64_t ldr, const int64_t* __restrict__ offs, int64_t nseg,
                   int C, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nseg) return;
    const int c = blockIdx.y * 128 + lane * 4;
    if (c >= C) return;
    const int64_t e0 = offs[i], e1 = offs[i + 1];
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {
        if (e1 > e0) {
            const float* p = rows + c;
            float4 a = __ldg(reinterpret_cast<const float4*>(p + e0 * ldr));
            int64_t e = e0 + 1;
            for (; e + 3 < e1; e += 4) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p + e * ldr));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p + (e + 1) * ldr));
                const float4 b2 = __ldg(reinterpret_cast<const float4*>(p + (e + 2) * ldr));
                const float4 b3 = __ldg(reinterpret_cast<const float4*>(p + (e + 3) * ldr));
                a.x = fmaxf(fmaxf(fmaxf(a.x, b0.x), fmaxf(b1.x, b2.x)), b3.x);
                a.y = fmaxf(fmaxf(fmaxf(a.y, b0.y), fmaxf(b1.y, b2.y)), b3.y);
                a.z = fmaxf(fmaxf(fmaxf(a.z, b0.z), fmaxf(b1.z, b2.z)), b3.z);
                a.w = fmaxf(fmaxf(fmaxf(a.w, b0.w), fmaxf(b1.w, b2.w)), b3.w);
            }
            for (; e < e1; ++e) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p + e * ldr));
                a.x = fmaxf(a.x, b0.x); a.y = fmaxf(a.y, b0.y); a.z = fmaxf(a.z, b0.z); a.w = fmaxf(a.w, b0.w);
            }
            m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
        }
        *reinterpret_cast<float4*>(out + i * ldo + c) = make_float4(m[0], m[1], m[2], m[3]);
    } else {
        const int nc = C - c < 4 ? C - c : 4;
        if (e1 > e0) {
            for (int k = 0; k < nc; ++k) m[k] = rows[e0 * ldr + c + k];
            for (int64_t e = e0 + 1; e < e1; ++e)
                for (int k = 0; k < nc; ++k) m[k] = fmaxf(m[k], rows[e * ldr + c + k]);
        }
        for (int k = 0; k < nc; ++k) out[i * ldo + c + k] = m[k];
    }
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_fps(const float* pos, const int64_t* ptr, int32_t B, const int64_t* start, const int64_t* out_ptr,
                int64_t* out, int32_t max_n_host, void* stream) {
    GNB_REQUIRE(pos && ptr && out_ptr && out, "gnb_fps: null pointer");
    GNB_REQUIRE(B >= 0 && max_n_host >= 0, "gnb_fps: negative size");
    if (B == 0 || max_n_host == 0) return GNB_OK;
    cudaStream_t st = as_stream(stream);
    const int stride = (max_n_host + 31) & ~31;
    // threads x points-per-thread: the loop is a latency chain (centre load -> distances -> two-level reduction -> barrier) with a
    // fixed part of ~200 ns per selected point, so few warps with many register-resident points each win: measured for 32
    // clouds 4096 -> 2048: 512x8 854 us, 256x16 797 us, 128x32 785 us; 2048 -> 512: 512x8 (half padding) 226 us, 256x8 / 128x16 151 us
    if (max_n_host <= 256 * 8) {
        const size_t smem = (size_t)3 * stride * sizeof(float);
        GNB_CUDA(cudaFuncSetAttribute(fps_kernel<256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fps_kernel<256, 8><<<B, 256, smem, st>>>(pos, ptr, start, out_ptr, out, stride);
    } else if (max_n_host <= 128 * 32) {
        const size_t smem = (size_t)3 * stride * sizeof(float);
        GNB_CUDA(cudaFuncSetAttribute(fps_kernel<128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fps_kernel<128, 32><<<B, 128, smem, st>>>(pos, ptr, start, out_ptr, out, stride);
    } else if (max_n_host <= 1024 * 8) {
        const size_t smem = (size_t)3 * stride * sizeof(float);
        GNB_CUDA(cudaFuncSetAttribute(fps_kernel<1024, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fps_kernel<1024, 8><<<B, 1024, smem, st>>>(pos, ptr, start, out_ptr, out, stride);
    } else {
        const size_t smem = (size_t)4 * stride * sizeof(float);
        if (smem > 227 * 1024) {
            set_error("gnb_fps: cloud of %d points exceeds the shared-memory envelope (14336)", max_n_host);
            return GNB_ERR_UNSUPPORTED;
        }
        GNB_CUDA(cudaFuncSetAttribute(fps_kernel<1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fps_kernel<1024, 0><<<B, 1024, smem, st>>>(pos, ptr, start, out_ptr, out, stride);
    }
    return check_launch("gnb_fps");
}

int32_t gnb_ball_query(const float* x, const float* y, const int64_t* ptr_x, const int64_t* ptr_y, int32_t B,
                       int64_t sumM, double r, int32_t K, int64_t* nbr, int32_t* cnt, void* stream) {
    GNB_REQUIRE(x && y && ptr_x && ptr_y && nbr && cnt, "gnb_ball_query: null pointer");
    GNB_REQUIRE(B > 0 && K > 0 && sumM >= 0, "gnb_ball_query: bad size");
    if (sumM == 0) return GNB_OK;
    const float r2 = (float)(r * r);  // torch_cluster passes r*r (double) into a float kernel argument
    const int wpb = 8;
    ball_query_kernel<<<(unsigned)ceil_div<int64_t>(sumM, wpb), wpb * 32, 0, as_stream(stream)>>>(
        x, y, ptr_x, ptr_y, B, sumM, r2, K, nbr, cnt);
    return check_launch("gnb_ball_query");
}

int32_t gnb_radius_pairs(const int64_t* nbr, const int32_t* cnt, const int64_t* offs, int64_t sumM, int32_t K,
                         int64_t* row, int64_t* col, void* stream) {
    GNB_REQUIRE(nbr && cnt && offs && row && col, "gnb_radius_pairs: null pointer");
    if (sumM == 0) return GNB_OK;
    radius_pairs_kernel<<<(unsigned)ceil_div<int64_t>(sumM, 8), 256, 0, as_stream(stream)>>>(nbr, cnt, offs, sumM, K,
                                                                                          row, col);
    return check_launch("gnb_radius_pairs");
}

int32_t gnb_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* stream) {
    GNB_REQUIRE(out && (in || n == 0), "gnb_exclusive_scan_i32: null pointer");
    GNB_REQUIRE(n >= 0 && n <= (1ll << 24), "gnb_exclusive_scan_i32: n out of range");
    scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(in, n, out);
    return check_launch("gnb_exclusive_scan_i32");
}

int32_t gnb_knn(const float* x, const float* y, const int64_t* ptr_x, const int64_t* ptr_y, int32_t B, int64_t Ny,
                int32_t k, int64_t* idx, float* d2, void* stream) {
    GNB_REQUIRE(x && y && ptr_x && ptr_y && idx && d2, "gnb_knn: null pointer");
    GNB_REQUIRE(k >= 1 && k <= 16, "gnb_knn: k=%d outside [1,16]", k);
    if (Ny == 0) return GNB_OK;
    const unsigned grid = (unsigned)ceil_div<int64_t>(Ny, 128);
    cudaStream_t st = as_stream(stream);
    if (k == 1) knn_kernel<1><<<grid, 128, 0, st>>>(x, y, ptr_x, ptr_y, B, Ny, k, idx, d2);
    else if (k <= 3) knn_kernel<3><<<grid, 128, 0, st>>>(x, y, ptr_x, ptr_y, B, Ny, k, idx, d2);
    else if (k <= 8) knn_kernel<8><<<grid, 128, 0, st>>>(x, y, ptr_x, ptr_y, B, Ny, k, idx, d2);
    else knn_kernel<16><<<grid, 128, 0, st>>>(x, y, ptr_x, ptr_y, B, Ny, k, idx, d2);
    return check_launch("gnb_knn");
}

int32_t gnb_knn_interpolate(const float* feat, int64_t ldf, const int64_t* idx, const float* d2, int64_t Ny,
                            int32_t k, int32_t C, float* out, int64_t ldo, void* stream) {
    GNB_REQUIRE(feat && idx && d2 && out, "gnb_knn_interpolate: null pointer");
    if (Ny == 0 || C == 0) return GNB_OK;
    GNB_REQUIRE(k >= 1 && k <= 16, "gnb_knn_interpolate: k=%d outside [1,16]", k);
    const unsigned grid = (unsigned)ceil_div<int64_t>(Ny, 8);
    cudaStream_t st = as_stream(stream);
    const bool vec = C % 4 == 0 && ldf % 4 == 0 && ldo % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
#define GNB_INTERP(KT_)                                                                                                  \
    do {                                                                                                                 \
        if (vec) knn_interp_kernel<KT_, true><<<grid, 256, 0, st>>>(feat, ldf, idx, d2, Ny, k, C, out, ldo);              \
        else knn_interp_kernel<KT_, false><<<grid, 256, 0, st>>>(feat, ldf, idx, d2, Ny, k, C, out, ldo);                 \
    } while (0)
    if (k == 1) GNB_INTERP(1);
    else if (k <= 3) GNB_INTERP(3);
    else if (k <= 8) GNB_INTERP(8);
    else GNB_INTERP(16);
#undef GNB_INTERP
    return check_launch("gnb_knn_interpolate");
}

int32_t gnb_pointconv_edge_count(const int64_t* nbr, const int32_t* cnt, int64_t sumM, int32_t K, int32_t* ecnt,
                                 void* stream) {
    GNB_REQUIRE(nbr && cnt && ecnt, "gnb_pointconv_edge_count: null pointer");
    if (sumM == 0) return GNB_OK;
    pointconv_edge_count_kernel<<<(unsigned)ceil_div<int64_t>(sumM, 256), 256, 0, as_stream(stream)>>>(nbr, cnt, sumM,
                                                                                                    K, ecnt);
    return check_launch("gnb_pointconv_edge_count");
}

int32_t gnb_pointconv_gather(const float* x_feat, int64_t ldx, int32_t Cin, const float* pos_x, const float* pos_y,
                             const int64_t* nbr, const int32_t* cnt, const int64_t* eoffs, int64_t sumM, int32_t K,
                             float* edge, int64_t lde, void* stream) {
    GNB_REQUIRE(pos_x && pos_y && nbr && cnt && eoffs && edge, "gnb_pointconv_gather: null pointer");
    GNB_REQUIRE(Cin == 0 || x_feat, "gnb_pointconv_gather: null features");
    GNB_REQUIRE(lde >= Cin + 3, "gnb_pointconv_gather: lde too small");
    GNB_REQUIRE(K >= 1 && (K <= 71 || Cin <= 8), "gnb_pointconv_gather: at most 71 neighbours per centroid for feature rows wider than 8 (got K=%d)", K);
    if (sumM == 0) return GNB_OK;
    pointconv_gather_kernel<<<(unsigned)ceil_div<int64_t>(sumM, 8), 256, 0, as_stream(stream)>>>(
        x_feat, ldx, Cin, pos_x, pos_y, nbr, cnt, eoffs, sumM, K, edge, lde);
    return check_launch("gnb_pointconv_gather");
}

int32_t gnb_segment_max(const float* rows, int64_t ldr, const int64_t* offs, int64_t nseg, int32_t C, float* out,
                        int64_t ldo, void* stream) {
    GNB_REQUIRE(rows && offs && out, "gnb_segment_max: null pointer");
    if (nseg == 0 || C == 0) return GNB_OK;
    const dim3 grid((unsigned)ceil_div<int64_t>(nseg, 8), (unsigned)ceil_div(C, 128));
    const bool vec = C % 4 == 0 && ldr % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (vec) segment_max_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(rows, ldr, offs, nseg, C, out, ldo);
    else segment_max_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(rows, ldr, offs, nseg, C, out, ldo);
    return check_launch("gnb_segment_max");
}

}  // extern "C"
