// Mesh clean-up after marching cubes: the reference turns the closed winding-number
// iso-surface into the open garment by dropping every face that touches a vertex whose Gaussian gradient magnitude is
// below a threshold, then deleting the vertices no surviving face uses and re-indexing the faces
// (ref common/marching_cubes_util.py:19-35 inside wnf_to_mesh and :38-52 delete_invalid_verts; eval.py:532-546).
//
//     is_face_valid[f]   = on[f0] & on[f1] & on[f2]
//     raw_valid_vert_idx = unique(valid_faces)                      (ascending = original vertex order)
//     valid_faces        = rank_among_kept(valid_faces)
//
// Pure integer compaction, bit-exact by construction.  The whole batch is processed at once on the packed arrays that
// gnb_mc_emit_batch produces (faces hold per-sample LOCAL vertex ids; vptr / fptr are the per-sample row offsets):
//   1. mark    one thread per face: validity flag, and the three vertices of a valid face are flagged as used
//   2. scan    exclusive prefix sums of both flag arrays (three-kernel scan: tile sums, scan of the sums, tile scans)
//   3. bases   new per-sample offsets = prefix sums sampled at the old offsets -> 2(B+1) int64 the host reads ONCE
//   4. emit    kept vertex ids (ascending) and re-indexed faces, compacted
#include "common.cuh"

namespace gnb {

constexpr int MCU_TILE = 1024;  // elements per CTA in the scans (256 threads x 4)

__device__ __forceinline__ int mcu_find(const int64_t* __restrict__ ptr, int B, int64_t i) {
    int lo = 0, hi = B;  // ptr[lo] <= i < ptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
mcu_mark_kernel(const int32_t* __restrict__ faces, const int64_t* __restrict__ fptr, const int64_t* __restrict__ vptr, int B,
                int64_t F, const uint8_t* __restrict__ on, int32_t* __restrict__ fvalid, int32_t* __restrict__ used) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int64_t vb = vptr[mcu_find(fptr, B, f)];
    const int64_t g0 = vb + faces[f * 3], g1 = vb + faces[f * 3 + 1], g2 = vb + faces[f * 3 + 2];
    const int ok = (on[g0] != 0) & (on[g1] != 0) & (on[g2] != 0);
    fvalid[f] = ok;
    if (ok) { used[g0] = 1; used[g1] = 1; used[g2] = 1; }   // benign race: every writer stores 1
}

// ---- three-kernel exclusive scan of int32 flags (n up to 2^31) ---------------------------------------------------------
__global__ void __launch_bounds__(256)
mcu_tile_sum_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ tile_sum) {
    __shared__ int wsum[8];
    const int64_t base = (int64_t)blockIdx.x * MCU_TILE + threadIdx.x * 4;
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (base + k < n) ? in[base + k] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += wsum[w];
        tile_sum[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of the tile sums in place (int64 running carry), total -> *total
__global__ void __launch_bounds__(1024)
mcu_scan_sums_kernel(int32_t* __restrict__ tile_sum, int64_t ntiles, int64_t* __restrict__ tile_base, int64_t* __restrict__ total) {
    __shared__ long long wsum[32];
    __shared__ long long chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long carry = 0;
    for (int64_t base = 0; base < ntiles; base += 1024) {
        const int64_t i = base + tid;
        const long long v = i < ntiles ? (long long)tile_sum[i] : 0ll;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const long long w = wsum[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (i < ntiles) tile_base[i] = carry + wsum[warp] + incl - v;
        carry += chunk_total;
        __syncthreads();
    }
    if (tid == 0) *total = carry;
}

// exclusive prefix of every element: out[i] = tile_base[tile] + (sum of in[tile start .. i))
__global__ void __launch_bounds__(256)
mcu_tile_scan_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_base, int64_t* __restrict__ out) {
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * MCU_TILE + threadIdx.x * 4;
    int v[4];
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < warp) ? wsum[w] : 0;
    int64_t run = tile_base[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

// new per-sample offsets: rec[b] = vrank at vptr[b], rec[B+1+b] = frank at fptr[b]  (b = 0..B; rank at the end = total)
__global__ void mcu_bases_kernel(const int64_t* __restrict__ vptr, const int64_t* __restrict__ fptr, int B, int64_t V, int64_t F,
                                 const int64_t* __restrict__ vrank, const int64_t* __restrict__ frank,
                                 const int64_t* __restrict__ totals, int64_t* __restrict__ rec) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    const int64_t v = vptr[b], f = fptr[b];
    rec[b] = v >= V ? totals[0] : vrank[v];
    rec[B + 1 + b] = f >= F ? totals[1] : frank[f];
}

__global__ void __launch_bounds__(256)
mcu_emit_verts_kernel(const int32_t* __restrict__ used, const int64_t* __restrict__ vrank, int64_t V, int64_t* __restrict__ keep) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V && used[v]) keep[vrank[v]] = v;
}

__global__ void __launch_bounds__(256)
mcu_emit_faces_kernel(const int32_t* __restrict__ faces, const int64_t* __restrict__ fptr, const int64_t* __restrict__ vptr, int B,
                      int64_t F, const int32_t* __restrict__ fvalid, const int64_t* __restrict__ frank,
                      const int64_t* __restrict__ vrank, const int64_t* __restrict__ rec, int32_t* __restrict__ out) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F || !fvalid[f]) return;
    const int b = mcu_find(fptr, B, f);
    const int64_t vb = vptr[b], nb = rec[b];   // old / new first vertex of the sample
    const int64_t o = frank[f];
#pragma unroll
    for (int k = 0; k < 3; ++k) out[o * 3 + k] = (int32_t)(vrank[vb + faces[f * 3 + k]] - nb);
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int64_t gnb_mesh_cleanup_workspace_bytes(int64_t V, int64_t F) {
    if (V < 0 || F < 0) return 0;
    const int64_t tv = ceil_div<int64_t>(V > 0 ? V : 1, MCU_TILE), tf = ceil_div<int64_t>(F > 0 ? F : 1, MCU_TILE);
    // used i32[V] | fvalid i32[F] | vrank i64[V] | frank i64[F] | tile sums i32 + tile bases i64 for both | totals i64[2]
    int64_t b = 0;
    b += (V + 63) / 64 * 64 * 4 + (F + 63) / 64 * 64 * 4;
    b += (V + 7) / 8 * 8 * 8 + (F + 7) / 8 * 8 * 8;
    b += (tv + tf + 64) * 4 + (tv + tf + 64) * 8 + 64;
    return b + 1024;
}

int32_t gnb_mesh_cleanup_count(const int32_t* faces, const int64_t* fptr, const int64_t* vptr, int32_t B, int64_t V, int64_t F,
                               const uint8_t* on_surface, void* workspace, int64_t* rec, void* stream) {
    GNB_REQUIRE(fptr && vptr && workspace && rec && B >= 1 && V >= 0 && F >= 0, "gnb_mesh_cleanup_count: bad arguments");
    GNB_REQUIRE(V < (1ll << 31) && F < (1ll << 31), "gnb_mesh_cleanup_count: more than 2^31 vertices / faces");
    GNB_REQUIRE((F == 0 || faces) && (V == 0 || on_surface), "gnb_mesh_cleanup_count: null pointer");
    cudaStream_t st = as_stream(stream);
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    int32_t* used = reinterpret_cast<int32_t*>(p); p += (V + 63) / 64 * 64 * 4;
    int32_t* fvalid = reinterpret_cast<int32_t*>(p); p += (F + 63) / 64 * 64 * 4;
    int64_t* vrank = reinterpret_cast<int64_t*>(p); p += (V + 7) / 8 * 8 * 8;
    int64_t* frank = reinterpret_cast<int64_t*>(p); p += (F + 7) / 8 * 8 * 8;
    const int64_t tv = ceil_div<int64_t>(V > 0 ? V : 1, MCU_TILE), tf = ceil_div<int64_t>(F > 0 ? F : 1, MCU_TILE);
    int64_t* base_v = reinterpret_cast<int64_t*>(p); p += (tv + 32) * 8;
    int64_t* base_f = reinterpret_cast<int64_t*>(p); p += (tf + 32) * 8;
    int64_t* totals = reinterpret_cast<int64_t*>(p); p += 64;
    int32_t* sum_v = reinterpret_cast<int32_t*>(p); p += (tv + 32) * 4;
    int32_t* sum_f = reinterpret_cast<int32_t*>(p);
    GNB_CUDA(cudaMemsetAsync(used, 0, (size_t)((V + 63) / 64 * 64 * 4), st));
    GNB_CUDA(cudaMemsetAsync(totals, 0, 16, st));
    if (F > 0) mcu_mark_kernel<<<(unsigned)ceil_div<int64_t>(F, 256), 256, 0, st>>>(faces, fptr, vptr, B, F, on_surface, fvalid, used);
    if (V > 0) {
        mcu_tile_sum_kernel<<<(unsigned)tv, 256, 0, st>>>(used, V, sum_v);
        mcu_scan_sums_kernel<<<1, 1024, 0, st>>>(sum_v, tv, base_v, totals);
        mcu_tile_scan_kernel<<<(unsigned)tv, 256, 0, st>>>(used, V, base_v, vrank);
    }
    if (F > 0) {
        mcu_tile_sum_kernel<<<(unsigned)tf, 256, 0, st>>>(fvalid, F, sum_f);
        mcu_scan_sums_kernel<<<1, 1024, 0, st>>>(sum_f, tf, base_f, totals + 1);
        mcu_tile_scan_kernel<<<(unsigned)tf, 256, 0, st>>>(fvalid, F, base_f, frank);
    }
    mcu_bases_kernel<<<ceil_div(B + 1, 128), 128, 0, st>>>(vptr, fptr, B, V, F, vrank, frank, totals, rec);
    return check_launch("gnb_mesh_cleanup_count");
}

int32_t gnb_mesh_cleanup_emit(const int32_t* faces, const int64_t* fptr, const int64_t* vptr, int32_t B, int64_t V, int64_t F,
                              void* workspace, const int64_t* rec, int64_t* keep, int32_t* out_faces, void* stream) {
    GNB_REQUIRE(fptr && vptr && workspace && rec && B >= 1 && V >= 0 && F >= 0, "gnb_mesh_cleanup_emit: bad arguments");
    cudaStream_t st = as_stream(stream);
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const int32_t* used = reinterpret_cast<const int32_t*>(p); p += (V + 63) / 64 * 64 * 4;
    const int32_t* fvalid = reinterpret_cast<const int32_t*>(p); p += (F + 63) / 64 * 64 * 4;
    const int64_t* vrank = reinterpret_cast<const int64_t*>(p); p += (V + 7) / 8 * 8 * 8;
    const int64_t* frank = reinterpret_cast<const int64_t*>(p);
    if (V > 0 && keep) mcu_emit_verts_kernel<<<(unsigned)ceil_div<int64_t>(V, 256), 256, 0, st>>>(used, vrank, V, keep);
    if (F > 0 && out_faces)
        mcu_emit_faces_kernel<<<(unsigned)ceil_div<int64_t>(F, 256), 256, 0, st>>>(faces, fptr, vptr, B, F, fvalid, frank, vrank, rec, out_faces);
    return check_launch("gnb_mesh_cleanup_emit");
}

}  // extern "C"
