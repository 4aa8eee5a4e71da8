// Mesh post-processing used by the reference's evaluation of a prediction:
//
//   * largest connected component of the open garment mesh (ref eval.py:538-546: igl.adjacency_matrix +
//     igl.connected_components, argmax of the component sizes, delete_invalid_verts with the membership mask);
//   * area-weighted surface sampling (ref common/geometry_util.py:184-223 mesh_sample_barycentric and :160-181
//     barycentric_interpolation, called at eval.py:222-243): face selection by inverse-CDF over the normalised double
//     areas with the uniform variates of numpy's RandomState (drawn on the host: they ARE the reference's random stream),
//     random barycentric coordinates, interpolation of any per-vertex field.
//
// Components: lock-free union-find over the faces (hook the larger root under the smaller one, path halving), so the
// representative of a component is its lowest vertex id -- the order in which igl numbers components -- and "largest,
// first on ties" is well defined.  All integer work; bit-exact against the scipy.sparse.csgraph restatement.
#include "common.cuh"

namespace gnb {

__device__ __forceinline__ int mo_find_sample(const int64_t* __restrict__ ptr, int B, int64_t i) {
    int lo = 0, hi = B;  // ptr[lo] <= i < ptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- connected components ---------------------------------------------------------------------------------------------
__global__ void lcc_init_kernel(int32_t* __restrict__ parent, int32_t* __restrict__ size, int64_t V) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) { parent[v] = (int32_t)v; size[v] = 0; }
}
__device__ __forceinline__ int32_t lcc_find(int32_t* parent, int32_t x) {
    // path halving; parent pointers only ever decrease, so concurrent updates are safe
    while (true) {
        const int32_t p = parent[x];
        if (p == x) return x;
        const int32_t g = parent[p];
        if (g != p) parent[x] = g;
        x = p;
    }
}
__device__ __forceinline__ void lcc_union(int32_t* parent, int32_t a, int32_t b) {
    while (true) {
        a = lcc_find(parent, a);
        b = lcc_find(parent, b);
        if (a == b) return;
        if (a < b) { const int32_t t = a; a = b; b = t; }   // a > b: hook root a under b
        const int32_t old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
        // somebody hooked a meanwhile: retry from the new roots
    }
}
__global__ void __launch_bounds__(256)
lcc_union_kernel(const int32_t* __restrict__ faces, const int64_t* __restrict__ fptr, const int64_t* __restrict__ vptr, int B,
                 int64_t F, int32_t* parent) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int64_t vb = vptr[mo_find_sample(fptr, B, f)];
    const int32_t g0 = (int32_t)(vb + faces[f * 3]), g1 = (int32_t)(vb + faces[f * 3 + 1]), g2 = (int32_t)(vb + faces[f * 3 + 2]);
    lcc_union(parent, g0, g1);
    lcc_union(parent, g1, g2);
}
__global__ void __launch_bounds__(256)
lcc_flatten_kernel(int32_t* parent, int32_t* __restrict__ size, int64_t V) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int32_t r = lcc_find(parent, (int32_t)v);
    parent[v] = r;   // roots keep pointing at themselves, so concurrent flattening stays consistent
    atomicAdd(&size[r], 1);
}
// one CTA per sample: the root with the largest size, lowest root id on ties; then the membership mask
__global__ void __launch_bounds__(256)
lcc_select_kernel(const int32_t* __restrict__ parent, const int32_t* __restrict__ size, const int64_t* __restrict__ vptr,
                  uint8_t* __restrict__ is_cc, int32_t* __restrict__ labels, int64_t* __restrict__ summary) {
    const int b = blockIdx.x;
    const int64_t v0 = vptr[b], v1 = vptr[b + 1];
    __shared__ long long s_best[256];
    __shared__ int s_ncc[256];
    // key = size << 32 | (0x7fffffff - root): max key = largest size, lowest root
    long long best = -1;
    int ncc = 0;
    for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x)
        if (parent[v] == (int32_t)v) {
            ++ncc;
            const long long key = ((long long)size[v] << 32) | (long long)(0x7fffffff - (int32_t)v);
            best = key > best ? key : best;
        }
    s_best[threadIdx.x] = best;
    s_ncc[threadIdx.x] = ncc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            s_best[threadIdx.x] = s_best[threadIdx.x + s] > s_best[threadIdx.x] ? s_best[threadIdx.x + s] : s_best[threadIdx.x];
            s_ncc[threadIdx.x] += s_ncc[threadIdx.x + s];
        }
        __syncthreads();
    }
    const long long k = s_best[0];
    const int32_t root = k < 0 ? -1 : (int32_t)(0x7fffffff - (int32_t)(k & 0xffffffffll));
    if (threadIdx.x == 0 && summary != nullptr) {
        summary[b * 3 + 0] = s_ncc[0];
        summary[b * 3 + 1] = k < 0 ? 0 : (k >> 32);
        summary[b * 3 + 2] = root < 0 ? -1 : (int64_t)root - v0;   // local id of the component's lowest vertex
    }
    for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        if (is_cc != nullptr) is_cc[v] = parent[v] == root ? 1 : 0;
        if (labels != nullptr) labels[v] = (int32_t)(parent[v] - v0);   // component label = local id of its lowest vertex
    }
}

// ---- area-weighted sampling -------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
face_doublearea_kernel(const T* __restrict__ verts, const int32_t* __restrict__ faces, int64_t F, double* __restrict__ area) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int64_t a = faces[f * 3], b = faces[f * 3 + 1], c = faces[f * 3 + 2];
    const double ax = (double)verts[a * 3], ay = (double)verts[a * 3 + 1], az = (double)verts[a * 3 + 2];
    const double rx = (double)verts[b * 3] - ax, ry = (double)verts[b * 3 + 1] - ay, rz = (double)verts[b * 3 + 2] - az;
    const double sx = (double)verts[c * 3] - ax, sy = (double)verts[c * 3 + 1] - ay, sz = (double)verts[c * 3 + 2] - az;
    // every product and sum rounded on its own (numpy's cross / sum order; nvcc would contract these into FMAs)
    const double cx = __dsub_rn(__dmul_rn(ry, sz), __dmul_rn(rz, sy)), cy = __dsub_rn(__dmul_rn(rz, sx), __dmul_rn(rx, sz)),
                 cz = __dsub_rn(__dmul_rn(rx, sy), __dmul_rn(ry, sx));
    area[f] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz)));   // |r x s| = twice the triangle area (igl.doublearea)
}
// single-CTA inclusive scan in double (the mesh of one garment has a few 1e5 faces): cdf[i] = sum_{j<=i} a[j]; then
// normalised by the total like numpy's RandomState.choice (p = a / sum(a); cdf = cumsum(p); cdf /= cdf[-1])
__global__ void __launch_bounds__(1024)
scan_f64_kernel(const double* __restrict__ a, int64_t n, double* __restrict__ cdf) {
    __shared__ double wsum[32];
    __shared__ double carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + tid;
        double x = i < n ? a[i] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            double w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const double carry = carry_s;
        const double incl = carry + (warp ? wsum[warp - 1] : 0.0) + x;
        if (i < n) cdf[i] = incl;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
normalise_cdf_kernel(double* __restrict__ cdf, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double total = cdf[n - 1];
    if (i < n - 1) cdf[i] = cdf[i] / total;   // the last element is normalised by the kernel below (it is still needed here)
}
__global__ void finish_cdf_kernel(double* __restrict__ cdf, int64_t n) { cdf[n - 1] = 1.0; }

// u_face f64[M] and uv f64[M,2]: uniform variates in [0,1) from the host generator
__global__ void __launch_bounds__(256)
sample_faces_kernel(const double* __restrict__ cdf, int64_t F, const double* __restrict__ u_face, const double* __restrict__ uv,
                    int64_t M, int64_t* __restrict__ face_idx, double* __restrict__ bary) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    // searchsorted(cdf, u, side='right'): number of entries <= u
    const double u = u_face[m];
    int64_t lo = 0, hi = F;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    face_idx[m] = lo < F ? lo : F - 1;
    double a = uv[m * 2], b = uv[m * 2 + 1];
    if (a + b >= 1.0) { a = 1.0 - a; b = 1.0 - b; }   // reflect into the triangle
    bary[m * 3 + 0] = a;
    bary[m * 3 + 1] = b;
    bary[m * 3 + 2] = 1.0 - (a + b);
}
// result[m, c] = sum_i bary[m, i] * field[faces[face_idx[m], i], c], accumulated in the field's dtype like the reference's
// in-place `result[:, c] += ...` on an array of verts.dtype (the product itself is formed in double)
template <typename T>
__global__ void __launch_bounds__(256)
barycentric_interp_kernel(const double* __restrict__ bary, const int64_t* __restrict__ face_idx, const int32_t* __restrict__ faces,
                          const T* __restrict__ field, int C, int64_t M, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * C) return;
    const int64_t m = i / C;
    const int c = (int)(i - m * C);
    const int64_t f = face_idx[m];
    T r = (T)0;
#pragma unroll
    for (int k = 0; k < 3; ++k)   // product and sum rounded separately like numpy (no FMA contraction)
        r = (T)__dadd_rn((double)r, __dmul_rn(bary[m * 3 + k], (double)field[(int64_t)faces[f * 3 + k] * C + c]));
    out[i] = r;
}

}  // namespace gnb

using namespace gnb;

extern "C" {

int32_t gnb_mesh_components(const int32_t* faces, const int64_t* fptr, const int64_t* vptr, int32_t B, int64_t V, int64_t F,
                            int32_t* parent_ws, int32_t* size_ws, uint8_t* is_largest, int32_t* labels, int64_t* summary,
                            void* stream) {
    GNB_REQUIRE(fptr && vptr && parent_ws && size_ws, "gnb_mesh_components: null pointer");
    GNB_REQUIRE(B >= 1 && V >= 0 && F >= 0 && V < (1ll << 31), "gnb_mesh_components: need B >= 1 and V < 2^31");
    GNB_REQUIRE(F == 0 || faces, "gnb_mesh_components: null faces");
    cudaStream_t st = as_stream(stream);
    if (V > 0) lcc_init_kernel<<<(unsigned)ceil_div<int64_t>(V, 256), 256, 0, st>>>(parent_ws, size_ws, V);
    if (F > 0) lcc_union_kernel<<<(unsigned)ceil_div<int64_t>(F, 256), 256, 0, st>>>(faces, fptr, vptr, B, F, parent_ws);
    if (V > 0) lcc_flatten_kernel<<<(unsigned)ceil_div<int64_t>(V, 256), 256, 0, st>>>(parent_ws, size_ws, V);
    lcc_select_kernel<<<B, 256, 0, st>>>(parent_ws, size_ws, vptr, is_largest, labels, summary);
    return check_launch("gnb_mesh_components");
}

int32_t gnb_mesh_sample_barycentric(const void* verts, int32_t verts_f64, const int32_t* faces, int64_t F, const double* areas_in,
                                    const double* u_face, const double* uv, int64_t M, double* cdf_ws, int64_t* face_idx,
                                    double* bary, void* stream) {
    GNB_REQUIRE(faces && u_face && uv && cdf_ws && face_idx && bary, "gnb_mesh_sample_barycentric: null pointer");
    GNB_REQUIRE(F >= 1 && M >= 0, "gnb_mesh_sample_barycentric: need at least one face");
    GNB_REQUIRE(areas_in || verts, "gnb_mesh_sample_barycentric: need verts or face areas");
    cudaStream_t st = as_stream(stream);
    const double* areas = areas_in;
    if (areas == nullptr) {
        double* tmp = cdf_ws + F;   // cdf_ws holds 2 F doubles: [cdf | areas]
        if (verts_f64) face_doublearea_kernel<double><<<(unsigned)ceil_div<int64_t>(F, 256), 256, 0, st>>>(
            reinterpret_cast<const double*>(verts), faces, F, tmp);
        else face_doublearea_kernel<float><<<(unsigned)ceil_div<int64_t>(F, 256), 256, 0, st>>>(
            reinterpret_cast<const float*>(verts), faces, F, tmp);
        areas = tmp;
    }
    scan_f64_kernel<<<1, 1024, 0, st>>>(areas, F, cdf_ws);
    normalise_cdf_kernel<<<(unsigned)ceil_div<int64_t>(F, 256), 256, 0, st>>>(cdf_ws, F);
    finish_cdf_kernel<<<1, 1, 0, st>>>(cdf_ws, F);
    if (M > 0) sample_faces_kernel<<<(unsigned)ceil_div<int64_t>(M, 256), 256, 0, st>>>(cdf_ws, F, u_face, uv, M, face_idx, bary);
    return check_launch("gnb_mesh_sample_barycentric");
}

int32_t gnb_barycentric_interpolation(const double* bary, const int64_t* face_idx, const int32_t* faces, const void* field,
                                      int32_t field_f64, int32_t C, int64_t M, void* out, void* stream) {
    GNB_REQUIRE(bary && face_idx && faces && field && out, "gnb_barycentric_interpolation: null pointer");
    GNB_REQUIRE(C >= 1 && M >= 0, "gnb_barycentric_interpolation: bad sizes");
    if (M == 0) return GNB_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned grid = (unsigned)ceil_div<int64_t>(M * C, 256);
    if (field_f64) barycentric_interp_kernel<double><<<grid, 256, 0, st>>>(bary, face_idx, faces, reinterpret_cast<const double*>(field), C, M,
                                                                         reinterpret_cast<double*>(out));
    else barycentric_interp_kernel<float><<<grid, 256, 0, st>>>(bary, face_idx, faces, reinterpret_cast<const float*>(field), C, M,
                                                               reinterpret_cast<float*>(out));
    return check_launch("gnb_barycentric_interpolation");
}

}  // extern "C"
