// Chamfer metrics on the device: the nearest-neighbour cores of the reference's
// eval.py -- `get_chamfer` inside compute_chamfer (:259-271) and inside compute_hybrid_chamfer (:381-401), which build
// two scipy cKDTrees per sample and query 10 k points each way.
//
//   forward[i]  = min_j |q_i - r_j|          (chamfer)            or  |qa_i - rb_{argmin_j |q_i - r_j|}|   (hybrid: the
//   match is found in NOCS space, the distance is measured between the corresponding simulation-space points)
//
// One thread per query point; the reference set of the query's sample is streamed through shared memory in tiles of
// 1024 points (brute force: 10^8 pair tests per direction and sample, far below a millisecond on a B200).  Squared
// distances in fp32 with every operation rounded (the same sqdist_nofma as ball query / kNN, ties -> lowest index); the
// reported distance of the chosen pair is recomputed in double from the coordinates, like cKDTree / np.linalg.norm on
// the float32 arrays promoted to double.  Per-sample sums are accumulated in double (warp shuffle + one atomic per warp).
#include "common.cuh"

namespace gnb {

constexpr int NN_THREADS = 128, NN_TILE = 1024;

__global__ void __launch_bounds__(NN_THREADS)
nn1_kernel(const float* __restrict__ q, const int64_t* __restrict__ ptr_q, const float* __restrict__ r,
           const int64_t* __restrict__ ptr_r, const float* __restrict__ qa, const float* __restrict__ rb, int64_t max_q,
           int64_t* __restrict__ idx_out, double* __restrict__ dist_out, double* __restrict__ sums) {
    __shared__ float sx[NN_TILE], sy[NN_TILE], sz[NN_TILE];
    const int b = blockIdx.y;
    const int64_t q0 = ptr_q[b], nq = ptr_q[b + 1] - q0;
    const int64_t r0 = ptr_r[b], nr = ptr_r[b + 1] - r0;
    const int64_t first = (int64_t)blockIdx.x * NN_THREADS;
    if (first >= nq) return;  // uniform per CTA
    const int64_t i = first + threadIdx.x;
    const bool active = i < nq;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (active) { px = q[(q0 + i) * 3]; py = q[(q0 + i) * 3 + 1]; pz = q[(q0 + i) * 3 + 2]; }
    float best = INFINITY;
    int64_t bj = -1;
    for (int64_t t0 = 0; t0 < nr; t0 += NN_TILE) {
        const int n = (int)((nr - t0) < NN_TILE ? (nr - t0) : NN_TILE);
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += NN_THREADS) {
            sx[k] = r[(r0 + t0 + k) * 3]; sy[k] = r[(r0 + t0 + k) * 3 + 1]; sz[k] = r[(r0 + t0 + k) * 3 + 2];
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int k = 0; k < n; ++k) {
                const float d = sqdist_nofma(sx[k], sy[k], sz[k], px, py, pz);
                if (d < best) { best = d; bj = t0 + k; }   // strict <: ties keep the lowest index
            }
        }
    }
    double dist = 0.0;
    if (active && bj >= 0) {
        const float* a = qa ? qa + (q0 + i) * 3 : q + (q0 + i) * 3;
        const float* c = rb ? rb + (r0 + bj) * 3 : r + (r0 + bj) * 3;
        const double dx = (double)a[0] - (double)c[0], dy = (double)a[1] - (double)c[1], dz = (double)a[2] - (double)c[2];
        dist = sqrt(dx * dx + dy * dy + dz * dz);
        if (idx_out) idx_out[q0 + i] = bj;
        if (dist_out) dist_out[q0 + i] = dist;
    }
    (void)max_q;
    double s = dist;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && sums) atomicAdd(&sums[b], s);
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_nn1_distance(const float* q, const int64_t* ptr_q, const float* r, const int64_t* ptr_r, int32_t B, int64_t max_q,
                         const float* qa, const float* rb, int64_t* idx, double* dist, double* sums, void* stream) {
    GNB_REQUIRE(q && ptr_q && r && ptr_r && B >= 1 && max_q >= 0, "gnb_nn1_distance: bad arguments");
    GNB_REQUIRE((qa == nullptr) == (rb == nullptr), "gnb_nn1_distance: qa and rb go together");
    if (sums) GNB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * B, as_stream(stream)));
    if (max_q == 0) return GNB_OK;
    const dim3 grid((unsigned)ceil_div<int64_t>(max_q, NN_THREADS), (unsigned)B);
    nn1_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(q, ptr_q, r, ptr_r, qa, rb, max_q, idx, dist, sums);
    return check_launch("gnb_nn1_distance");
}

}  // extern "C"
