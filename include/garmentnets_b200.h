/*
 * garmentnets_b200 -- C-ABI of the B200-native GarmentNets dense-inference hot path.
 *
 * One `extern "C"` entry point per native component the reference reaches through
 * third-party binaries.  Every pointer is a DEVICE pointer unless
 * the parameter name ends in `_host`; sizes are plain integers; `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = legacy default stream).  The caller owns and
 * allocates every buffer (outputs worst-case sized).  All calls are asynchronous on
 * `stream` unless stated otherwise.  Return value: 0 = OK, negative = `gnb_status`;
 * `gnb_last_error()` returns a thread-local human-readable message.
 *
 * No torch types appear here: the Python host side (garmentnets_b200/_lib.py) binds this
 * file with ctypes and passes `tensor.data_ptr()` / `torch.cuda.current_stream().cuda_stream`.
 *
 * "ref:" citations are relative to the reference repository root (real-stanford/garmentnets).
 */
#ifndef GARMENTNETS_B200_H
#define GARMENTNETS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* every entry point below is exported; everything else in the library is hidden */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum gnb_status {
    GNB_OK = 0,
    GNB_ERR_INVALID = -1,      /* bad argument (shape, null pointer, unsupported size)       */
    GNB_ERR_CUDA = -2,         /* CUDA runtime error; message holds cudaGetErrorString        */
    GNB_ERR_UNSUPPORTED = -3,  /* valid request outside the compiled envelope                 */
    GNB_ERR_NO_SURFACE = -4    /* marching cubes: level outside [min,max] (skimage: ValueError) */
} gnb_status;

typedef enum gnb_reduce { GNB_REDUCE_SUM = 0, GNB_REDUCE_MEAN = 1, GNB_REDUCE_MAX = 2, GNB_REDUCE_MIN = 3 } gnb_reduce;

/* ---- library ------------------------------------------------------------------------ */
int32_t gnb_version(void);                 /* major*10000 + minor*100 + patch */
const char* gnb_last_error(void);          /* thread-local, never NULL        */
int32_t gnb_device_sm_count(void);         /* SMs of the current device; <0 on error */

/* ---- N1: farthest point sampling ------------------------------------------------------
 * ref: components/pointnet2.py:26  `idx = fps(pos, batch, ratio=self.ratio)` (torch_cluster 1.5.9).
 * pos   f32[sumN,3]; ptr i64[B+1] cloud offsets (batch vector is sorted, ref datasets/...:458-460);
 * start i64[B] LOCAL start index per cloud (the reference's random start, injected; NULL = 0);
 * out_ptr i64[B+1] offsets of the selected indices (m_b = ceil(ratio*n_b), computed by the caller);
 * out   i64[sumM] GLOBAL indices into pos, in selection order.
 * Distances are fp32 ((dx*dx + dy*dy) + dz*dz) without fused multiply-add; argmax ties -> lowest index.
 * max_n_host = largest cloud size (<= 14336 points; one CTA keeps a cloud + its distances in shared memory). */
int32_t gnb_fps(const float* pos, const int64_t* ptr, int32_t B, const int64_t* start,
                const int64_t* out_ptr, int64_t* out, int32_t max_n_host, void* stream);

/* ---- N2: ball query --------------------------------------------------------------------
 * ref: components/pointnet2.py:28-29 `radius(pos, pos[idx], r, batch, batch[idx], max_num_neighbors=64)`.
 * x f32[sumN,3] points, y f32[sumM,3] queries (centroids); ptr_x/ptr_y i64[B+1].
 * For query q of cloud b: the first K points j of cloud b IN INDEX ORDER with d2(j,q) < (float)(r*r).
 * nbr i64[sumM,K] global indices into x, -1 padded; cnt i32[sumM]. */
int32_t gnb_ball_query(const float* x, const float* y, const int64_t* ptr_x, const int64_t* ptr_y,
                       int32_t B, int64_t sumM, double r, int32_t K,
                       int64_t* nbr, int32_t* cnt, void* stream);

/* Compaction of (nbr,cnt) into the (row,col) pair list torch_cluster.radius returns
 * (row = query index, col = point index; sorted by row then ascending col).
 * offs i64[sumM+1] = exclusive prefix sum of cnt (see gnb_exclusive_scan_i32). */
int32_t gnb_radius_pairs(const int64_t* nbr, const int32_t* cnt, const int64_t* offs,
                         int64_t sumM, int32_t K, int64_t* row, int64_t* col, void* stream);

/* out i64[n+1] = exclusive prefix sum of in i32[n]; out[n] = total.  n <= 2^24. */
int32_t gnb_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* stream);

/* ---- N3: k nearest neighbours + inverse-distance interpolation --------------------------
 * ref: components/pointnet2.py:72 `knn_interpolate(x, pos, pos_skip, batch, batch_skip, k)` (PyG 1.7.2).
 * gnb_knn: for every query y_i the k nearest x_j of the same cloud, ascending d2, ties -> lower j.
 * idx i64[Ny,k] (-1 padded when the cloud has < k points), d2 f32[Ny,k].  k <= 16. */
int32_t gnb_knn(const float* x, const float* y, const int64_t* ptr_x, const int64_t* ptr_y,
                int32_t B, int64_t Ny, int32_t k, int64_t* idx, float* d2, void* stream);
/* out[i, 0:C] = sum_j(feat[idx_ij] * w_ij) / sum_j(w_ij),  w = 1/max(d2,1e-16); sums in rank order,
 * products rounded before adding (no FMA).  feat f32[Nx,C] row stride ldf; out row stride ldo (so the
 * result can be written straight into the `cat([x, x_skip])` buffer of ref components/pointnet2.py:73-74). */
int32_t gnb_knn_interpolate(const float* feat, int64_t ldf, const int64_t* idx, const float* d2,
                            int64_t Ny, int32_t k, int32_t C, float* out, int64_t ldo, void* stream);

/* ---- N4: PointConv grouping ------------------------------------------------------------
 * ref: components/pointnet2.py:30-31 `PointConv(nn)(x, (pos, pos[idx]), edge_index)` (PyG 1.7.2,
 * add_self_loops=True): the edge set of centroid i is ballquery(i) U {point with flat index i}.
 * gnb_pointconv_edge_count: ecnt i32[sumM] = cnt[i] + (i not in nbr[i,:cnt[i]]).
 * gnb_pointconv_gather: edge rows e in [eoffs[i], eoffs[i+1]) =
 *     [x_feat[j, 0:Cin], pos_x[j]-pos_y[i]]  (row stride lde >= Cin+3), neighbours first, self loop last.
 * gnb_segment_max: out[i, c] = max over the centroid's edge rows (0 when it has none). */
int32_t gnb_pointconv_edge_count(const int64_t* nbr, const int32_t* cnt, int64_t sumM, int32_t K,
                                 int32_t* ecnt, void* stream);
int32_t gnb_pointconv_gather(const float* x_feat, int64_t ldx, int32_t Cin, const float* pos_x,
                             const float* pos_y, const int64_t* nbr, const int32_t* cnt,
                             const int64_t* eoffs, int64_t sumM, int32_t K,
                             float* edge, int64_t lde, void* stream);
int32_t gnb_segment_max(const float* rows, int64_t ldr, const int64_t* offs, int64_t nseg, int32_t C,
                        float* out, int64_t ldo, void* stream);

/* ---- N10: per-point Linear -> ReLU -> BatchNorm(eval) ----------------------------------
 * ref: components/mlp.py:9-20 (one `Sequential(Linear, ReLU, PointBatchNorm1D)` block).
 * Y[r, n] = post( sum_k X[r,k] * W[n,k] + bias[n] );  post(v) = (relu ? max(v,0) : v) * bn_scale[n] + bn_shift[n]
 * (bn_scale/bn_shift NULL = identity; they are gamma/sqrt(var+eps) and beta - mean*scale).
 * X f32[R,K] row stride ldx; W f32[N,K] contiguous; Y f32[R,N] row stride ldy.
 * rows_dev (nullable): device i64 holding the number of valid rows (<= R); tiles beyond it exit early. */
int32_t gnb_linear(const float* X, int64_t R, int32_t K, int64_t ldx, const float* W, const float* bias,
                   int32_t N, int32_t relu, const float* bn_scale, const float* bn_shift,
                   float* Y, int64_t ldy, const int64_t* rows_dev, void* stream);

/* ---- N10 on the tensor cores: the same Linear -> ReLU -> BatchNorm block through tcgen05 ----------
 * ref: components/mlp.py:9-20.  Same contract as gnb_linear; the weight is pre-packed once per parameter version:
 * gnb_linear_tc_pack writes fp16 hi/lo shared-memory images of W (scaled by 2^scale_log2) into `packed`
 * (gnb_linear_tc_packed_bytes(N,K) bytes) and bias | bn_scale | bn_shift (NULL = 0 | 1 | 0) into
 * `cparams` f32[3 * gnb_linear_tc_padded_cols(N,K)].  Products are formed as hi*hi + lo*hi + hi*lo with fp32
 * accumulation (fp32-level accuracy).  No alignment requirement on X / ldx; Y rows are written with 16-byte
 * stores when Y and ldy allow it. */
int64_t gnb_linear_tc_packed_bytes(int32_t N, int32_t K);
int64_t gnb_linear_tc_padded_cols(int32_t N, int32_t K);
int32_t gnb_linear_tc_pack(const float* W, int32_t N, int32_t K, const float* bias, const float* bn_scale,
                           const float* bn_shift, int32_t scale_log2, void* packed, float* cparams,
                           void* stream);
int32_t gnb_linear_tc(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed,
                      const float* cparams, int32_t scale_log2, int32_t N, int32_t relu,
                      float* Y, int64_t ldy, const int64_t* rows_dev, void* stream);

/* gnb_linear_tc that also raises the per-device fp16 range flag (gnb_f16_overflow_fetch) when an OUTPUT value lies outside
 * +-65504 or is not finite: for outputs that feed a saturating fp16 operand split (the hoisted first decoder layer, whose
 * 1 GB result the pipeline would otherwise have to read once more with gnb_f16_range_check). */
int32_t gnb_linear_tc_flagged(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed,
                      const float* cparams, int32_t scale_log2, int32_t N, int32_t relu,
                      float* Y, int64_t ldy, const int64_t* rows_dev, void* stream);

/* PointConv aggregation fused into the last edge-MLP layer (ref components/pointnet2.py:31, PyG PointConv aggr='max'):
 * the rows of one segment (centroid) are contiguous; seg i32[R] holds the segment id of every row
 * (gnb_segment_ids fills it from the CSR offsets).  Instead of writing Y, every 128-row tile is reduced on chip and
 * merged into enc_out u32[nseg, ldo] with order-preserving integer atomic max (zero-initialise it, then
 * gnb_segmax_decode turns it into f32 in place; untouched slots decode to 0 like an empty max aggregation).
 * The [R, N] activation never goes to HBM and the separate gnb_segment_max pass disappears. */
int32_t gnb_linear_tc_segmax(const float* X, int64_t R, int32_t K, int64_t ldx, const void* packed,
                             const float* cparams, int32_t scale_log2, int32_t N, int32_t relu,
                             const int32_t* seg, void* enc_out, int64_t ldo, const int64_t* rows_dev,
                             void* stream);
int32_t gnb_segment_ids(const int64_t* offs, int64_t nseg, int32_t* seg, void* stream);
int32_t gnb_segmax_decode(void* enc, int64_t count, void* stream);

/* ---- N11: NOCS bin head ----------------------------------------------------------------
 * ref: networks/conv_implicit_wnf.py:222-231.  logits f32[R, bins*3] viewed [R,bins,3]:
 * bin i64[R,3] = argmax over bins (first max), conf f32[R,3] = softmax at the argmax,
 * nocs f32[R,3] = bin * (float)(1/(bins-1)). */
int32_t gnb_nocs_head(const float* logits, int64_t R, int32_t bins, int64_t* bin, float* conf,
                      float* nocs, void* stream);

/* ---- N5: scatter-reduce ----------------------------------------------------------------
 * ref: networks/conv_implicit_wnf.py:92-94 / components/gridding.py:32-35
 * `torch_scatter.scatter(src=features.T, index, dim=-1, dim_size, reduce)`.
 * src element (c, n) at src[c*src_sc + n*src_sn]; out element (c, m) at out[c*out_sc + m*out_sm];
 * index i64[N] in [0, dim_size).  Slots that receive nothing are 0.  `scratch` i32[dim_size]
 * (owner/count workspace).  MAX/MIN are bit-exact and order independent; SUM/MEAN use fp32 atomics. */
int32_t gnb_scatter_reduce(const float* src, int64_t src_sc, int64_t src_sn, const int64_t* index,
                           int64_t N, int32_t C, int64_t dim_size, int32_t reduce,
                           float* out, int64_t out_sc, int64_t out_sm, int32_t* scratch, void* stream);

/* Aggregator glue, ref networks/conv_implicit_wnf.py:62-85 + components/gridding.py:161-206,230-256:
 * voxel index of each point from its NOCS position (trunc((p-lc)*((G-1)/(uc-lc))), clamped), flat index
 * b*G^3 + i0*G^2 + i1*G + i2, and the per-point feature row
 * [feat(Cf) | p - voxel_origin (3), sim_points (3) if include_point_feature | confidence (3) if
 * include_confidence_feature] written with row stride ldo (conv_implicit_wnf.py:76-85).
 * lower_corner / upper_corner: HOST float[3], NULL = (0,0,0) / (1,1,1) as shipped
 * (config/train_pipeline_default.yaml:43-44).  sim_points / conf may be NULL when their flag is 0. */
int32_t gnb_aggregator_features(const float* feat, int64_t ldf, int32_t Cf, const float* nocs,
                                const float* sim_points, const float* conf, const int64_t* batch, int64_t N,
                                int32_t G, const float* lower_corner, const float* upper_corner,
                                int32_t include_point_feature, int32_t include_confidence_feature,
                                int64_t* flat_idx, float* out, int64_t ldo, void* stream);

/* ---- N6-N8: 3D-UNet layers on channels-last (NDHWC) activations ---------------------------
 * ref: components/unet3d.py:43-72 ('gcr' SingleConv = GroupNorm -> Conv3d(3x3x3, pad 1, no bias) -> ReLU),
 * :222 MaxPool3d(2), :325-330 nearest upsampling, :291 cat((encoder_features, x), 1), :437 final 1x1x1.
 * gnb_groupnorm_stats: x f32[B,D,H,W,C] -> scale/shift f32[B,C] so that GN(x)[b,..,c] = x*scale + shift
 *   (biased variance over the group's (C/groups)*D*H*W elements, eps).  ws f64[B*groups*2] workspace. */
int32_t gnb_groupnorm_stats(const float* x, int32_t B, int64_t voxels, int32_t C, int32_t groups, float eps,
                            const float* gamma, const float* beta, float* scale, float* shift,
                            double* ws, void* stream);
/* y[b,d,h,w,co] = relu?( sum_{tap,ci} (x[b,d+dz,h+dy,w+dx,ci]*scale[b,ci]+shift[b,ci]) * Wt[tap,ci,co] ) ; zero
 * padding is applied AFTER the affine (the reference pads the GroupNorm output).  Wt f32[27,Cin,Cout] is the
 * reference weight [Cout,Cin,3,3,3] permuted (tap = kd*9+kh*3+kw).  scale/shift NULL = identity. */
int32_t gnb_conv3d_k3(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin,
                      const float* scale, const float* shift, const float* Wt, int32_t Cout,
                      int32_t relu, float* y, void* stream);
/* Precision of the two cross terms (lo*hi + hi*lo) of the tensor-core 3x3x3 convolutions, process-wide:
 *   0 (default)  two fp16 MMAs per K-step (three tensor passes in total, ~1e-7 relative per layer);
 *   1            ONE e4m3 MMA per K-step over [lo*2^12 | a] x [w*2^(s-12) | w_lo*2^s] (two pass-equivalents, ~2e-5 relative
 *                per layer).
 * The mode is read by gnb_conv3d_tc*_pack_weights, gnb_gn_apply_split* (they emit the operands in the mode's format) and
 * gnb_conv3d_tc / gnb_conv3d_tc_dx: set it BEFORE packing weights and do not change it between a split and its convolution. */
int32_t gnb_conv_tc_set_cross_precision(int32_t mode);
int32_t gnb_conv_tc_cross_precision(void);

/* Tensor-core (TMA + tcgen05) version of gnb_conv3d_k3 for power-of-two grids with at least 128 voxels in the batch and
 * Cout in {32,64,96,128} (gnb_conv3d_tc_supported).  Three launches per 'gcr' SingleConv:
 *   gnb_groupnorm_stats -> gnb_gn_apply_split (x*scale+shift written once as fp16 hi + lo, channels zero-padded to a
 *   multiple of 64: xh, xl f16[B,D,H,W,Cpad]) -> gnb_conv3d_tc (per tap and 64-channel chunk two TMA box loads with
 *   hardware zero fill = the conv padding, weights from the images made by gnb_conv3d_tc_pack_weights
 *   (27 * Cpad * Cout * 4 bytes, weights pre-multiplied by 2^scale_log2 so their fp16 lo parts stay normal; the
 *   same scale_log2 is passed to gnb_conv3d_tc and undone exactly in its epilogue), products as hi*hi + lo*hi + hi*lo
 *   with fp32 accumulation in TMEM). */
int32_t gnb_conv3d_tc_supported(int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout);
int32_t gnb_conv3d_tc_pack_weights(const float* W, int32_t Cout, int32_t Cin, int32_t scale_log2, void* packed,
                                   void* stream);
int32_t gnb_gn_apply_split(const float* x, int32_t B, int64_t voxels, int32_t C, const float* scale, const float* shift,
                           void* xh, void* xl, void* stream);
/* The two passes above on the VIRTUAL tensor cat((skip, nearest_upsample_2x(x_low)), channel) -- the decoder's joining
 * (ref components/unet3d.py:291,325-330) -- without materialising it: skip f32[B,D,H,W,Cs], x_low f32[B,D/2,H/2,W/2,Cx];
 * voxels = D*H*W; scale / shift f32[B, Cs+Cx]; Cs, Cx and the group width multiples of 4. */
int32_t gnb_groupnorm_stats_cat(const float* skip, int32_t Cs, const float* x_low, int32_t Cx, int32_t B, int64_t voxels,
                                int32_t groups, float eps, const float* gamma, const float* beta, float* scale,
                                float* shift, double* ws, void* stream);
int32_t gnb_gn_apply_split_cat(const float* skip, int32_t Cs, const float* x_low, int32_t Cx, int32_t B, int32_t D,
                               int32_t H, int32_t W, const float* scale, const float* shift, void* xh, void* xl,
                               void* stream);
int32_t gnb_conv3d_tc(const void* xh, const void* xl, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin,
                      const void* w_packed, int32_t scale_log2, int32_t Cout, int32_t relu, float* y, void* stream);

/* Stacked-dx variant of gnb_conv3d_tc for the narrow layers (Cout in {32,64}, W in {8,16,32}): the three kw taps of a
 * (kd,kh) pair share one activation box, their weights are stacked along N and the shift along W is applied to the
 * output with warp shuffles; 9 instead of 27 activation boxes per tile and 2 instead of 3 MMAs per fp16-split product.
 * Same arguments and results as gnb_conv3d_tc; weights packed by gnb_conv3d_tc_dx_pack_weights (same byte size). */
int32_t gnb_conv3d_tc_dx_supported(int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout);
int32_t gnb_conv3d_tc_dx_pack_weights(const float* W, int32_t Cout, int32_t Cin, int32_t scale_log2, void* packed,
                                      void* stream);
int32_t gnb_conv3d_tc_dx(const void* xh, const void* xl, int32_t B, int32_t D, int32_t H, int32_t W, int32_t Cin,
                         const void* w_packed, int32_t scale_log2, int32_t Cout, int32_t relu, float* y, void* stream);
int32_t gnb_maxpool3d_2(const float* x, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, float* y, void* stream);
/* y[b,d,h,w, 0:Cs] = skip[b,d,h,w,:];  y[..., Cs:Cs+Cx] = x[b, d*Dx/D, h*Hx/H, w*Wx/W, :] (nearest). */
int32_t gnb_upsample_concat(const float* skip, int32_t Cs, const float* x, int32_t Cx, int32_t B,
                            int32_t D, int32_t H, int32_t W, int32_t Dx, int32_t Hx, int32_t Wx,
                            float* y, void* stream);
/* NCDHW (arbitrary element strides, in elements) -> contiguous NDHWC, and back. */
int32_t gnb_to_channels_last(const float* x, int64_t sb, int64_t sc, int64_t sd, int64_t sh, int64_t sw,
                             int32_t B, int32_t C, int32_t D, int32_t H, int32_t W, float* y, void* stream);

/* ---- N9+N10: implicit decoder ------------------------------------------------------------
 * ref: networks/conv_implicit_wnf.py:128-149 (grid_sample trilinear/border/align_corners + MLP) and the dense
 * 128^3 loop predict.py:145-158.
 * gnb_trilinear_sample: vol f32[B,D,H,W,C] channels-last; q f32[B,M,3];  out f32[B*M, C] row stride ldo.
 *   flip=0: q[...,0]->W, q[...,1]->H, q[...,2]->D  (the un-flipped convention of conv_implicit_wnf.py:137-142);
 *   flip=1: q[...,0]->D, q[...,1]->H, q[...,2]->W  (components/gridding.py:69-70 nocs_grid_sample).
 *   post: 0 = none, 1 = ReLU then *bn_scale+bn_shift (used when Linear1 is hoisted onto the feature grid).
 * gnb_trilinear_sample_grid: same with the implicit regular query lattice q[i,j,k] = (i,j,k)/(Q-1)
 *   (ref components/gridding.py:139-159), rows [m0, m0+M) of the flattened lattice of sample b. */
int32_t gnb_trilinear_sample(const float* vol, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C,
                             const float* q, int64_t M, int32_t flip, int32_t post,
                             const float* bn_scale, const float* bn_shift,
                             float* out, int64_t ldo, void* stream);
int32_t gnb_trilinear_sample_grid(const float* vol, int32_t b, int32_t D, int32_t H, int32_t W, int32_t C,
                                  int32_t Q, int64_t m0, int64_t M, int32_t post,
                                  const float* bn_scale, const float* bn_shift,
                                  float* out, int64_t ldo, void* stream);

/* Tensor-core (tcgen05 + TMEM) decoder tail for the shipped decoder shape [C, 256, 256, Cout<=3]:
 *     y = BN3(ReLU( BN2(ReLU(H1 * W2^T + b2)) * W3^T + b3 ))        (ref components/mlp.py:9-20 blocks 2 and 3)
 * with both GEMM operands split into fp16 hi+lo (three MMAs per product, fp32 accumulation in TMEM) so that the result
 * stays within the 1e-4 fp32 parity bound.  H1 is produced on the fly and never stored:
 *   Q  > 0 (lattice mode, needs Q == 128): H1 = BN1(ReLU(trilinear(U))) over the implicit regular query lattice
 *          (i,j,k)/(Q-1) of every sample b < B (ref predict.py:145-158); U f32[B,G,G,G,256] is Linear1 applied on the
 *          feature grid (it commutes with the interpolation); out f32[B, Q^3, Cout].
 *   Q == 0 (row mode): H1 = X f32[R,256] (row stride ldx, even) given explicitly; out f32[R, Cout].
 * w2_packed: W2 * 2^scale_log2 re-laid by gnb_pack_f16_split (N*K*4 bytes); the power-of-two scale (chosen by the
 * caller so that max|W2| * 2^s < 2^14) keeps the fp16 lo parts out of the subnormal range and is undone exactly in
 * the epilogue (w2_scale_log2).  scratch: f32[1024] workspace.  bn*_scale/shift may be
 * NULL (= identity) except bn1 in lattice mode. */
int32_t gnb_pack_f16_split(const float* W, int32_t N, int32_t K, int32_t scale_log2, void* packed, void* stream);
int32_t gnb_decode_tc(const float* U, int64_t ldx, int32_t B, int32_t G, int32_t Q, int64_t R,
                      const float* bn1_scale, const float* bn1_shift, const void* w2_packed, int32_t w2_scale_log2,
                      const float* b2,
                      const float* bn2_scale, const float* bn2_shift, const float* W3, const float* b3,
                      const float* bn3_scale, const float* bn3_shift, int32_t Cout, float* scratch, float* out,
                      void* stream);
/* Second-generation lattice kernel (same math as gnb_decode_tc with Q == 128, restructured for L2 traffic): a work item
 * is a PAIR of adjacent lattice lines (two TMEM accumulators), the pipeline is K-chunk major with a two-slot A ring, the
 * producers gather 16 bytes per lane and share the corner columns of the two lines, and BN1 is folded into W2:
 *   w2f_packed = gnb_pack_f16_split(W2 * diag(bn1_scale)) with its power-of-two scale w2f_scale_log2;
 *   W2 (unfolded, f32[256,256]) and bn1_shift are used once per call to form b2 + W2 bn1_shift.
 * U f32[B,G,G,G,256] (16-byte aligned), 2 <= G <= 64, Q == 128; scratch f32[2048]; out f32[B, Q^3, Cout]. */
int32_t gnb_decode_lattice(const float* U, int32_t B, int32_t G, int32_t Q, const float* W2, const void* w2f_packed,
                           int32_t w2f_scale_log2, const float* b2, const float* bn1_shift, const float* bn2_scale,
                           const float* bn2_shift, const float* W3, const float* b3, const float* bn3_scale,
                           const float* bn3_shift, int32_t Cout, float* scratch, float* out, void* stream);
/* Kernel variant of gnb_decode_lattice: 0 (default) = one CTA per SM, cta_group::1, 8 A-producer warps (4 channels per lane);
 * 2 = the same with 16 producer warps (2 channels per lane); 1 = clusters of two CTAs issuing tcgen05.mma.cta_group::2 (each SM
 * keeps half of every W2 piece).  Both alternatives were measured slower on B200 (17.7 / 18.7 vs 17.1 ms at batch 32) and are
 * kept for A/B measurements and tests; results are bit-identical across the three. */
int32_t gnb_decode_lattice_set_mode(int32_t cta_pair);
/* Query mode of the same kernel (the surface / warp-field decoder, ref predict.py:184-187 and
 * networks/conv_implicit_wnf.py:263-269): H1 = BN1(ReLU(trilinear(U[b], q_r))) for explicit query points q f32[R,3] in
 * [0,1]^3 (coordinate 0 -> W axis: the reference does not flip xyz, :135-142).  Rows are ragged per sample: rows
 * qptr[b] .. qptr[b+1]-1 (device i64[B+1], qptr[0] = 0, qptr[B] = R) sample U[b]; 1 <= B <= 128 per call.  The
 * 8-corner gather runs inside the A-operand producer warps, so neither the interpolated features nor the hidden
 * activations touch HBM.  out f32[R, Cout]. */
int32_t gnb_decode_tc_query(const float* U, int32_t B, int32_t G, const float* q, const int64_t* qptr, int64_t R,
                            const float* bn1_scale, const float* bn1_shift, const void* w2_packed,
                            int32_t w2_scale_log2, const float* b2, const float* bn2_scale, const float* bn2_shift,
                            const float* W3, const float* b3, const float* bn3_scale, const float* bn3_shift,
                            int32_t Cout, float* scratch, float* out, void* stream);

/* Query mode on the 32-channel feature grid (FUSED): the decoder's first Linear is applied per query inside the kernel
 * instead of being hoisted onto the grid -- trilinear interpolation commutes with the affine map, so
 * interp(X) W1^T + b1 == interp(X W1^T + b1).  X f32[B,G,G,G,32] (channels-last; the UNet's last decoder level when
 * final_conv is folded into W1), W1 f32[256,32], b1 f32[256].  The gather shrinks from 8 x 1 KB to 8 x 128 B per
 * query; everything else as gnb_decode_tc_query.  bn1_scale / bn1_shift may be NULL when the caller has folded
 * BatchNorm1 (it follows the ReLU, so it is linear in front of Linear2) into w2_packed / b2: W2' = W2 diag(scale),
 * b2' = b2 + W2 shift -- the fast path: BOTH contractions then run on tcgen05 (decode_query.cu: Linear1 as six N = 64 MMAs per
 * 64-channel chunk into a TMEM double buffer, a mid-epilogue warp group turns each chunk into the fp16 hi/lo operand of
 * Linear2 in shared memory).  With BatchNorm1 given, Linear1 is applied per query with FFMA2 in the producer warps.
 * scratch: f32[16384 + 512 * SMs] (tail constants, the power-of-two scale of W1 and its 32 KB fp16 shared-memory image, one
 * 128 x 4 row of partial dot products per CTA); 16384 + 131072 floats cover up to 256 SMs. */
int32_t gnb_decode_tc_query_fused(const float* X, int32_t B, int32_t G, int32_t C0, const float* W1, const float* b1,
                                  const float* q, const int64_t* qptr, int64_t R, const float* bn1_scale,
                                  const float* bn1_shift, const void* w2_packed, int32_t w2_scale_log2,
                                  const float* b2, const float* bn2_scale, const float* bn2_shift,
                                  const float* W3, const float* b3, const float* bn3_scale,
                                  const float* bn3_shift, int32_t Cout, float* scratch, float* out, void* stream);
/* Tests / A-B measurements: 1 forces the FFMA2-Linear1 kernel for every gnb_decode_tc_query_fused call, 0 (default) picks by
 * bn1_scale as described above. */
int32_t gnb_decode_query_set_mode(int32_t ffma_linear1);

/* ---- N13: gaussian gradient magnitude ------------------------------------------------------
 * ref: predict.py:162-163 `ni.gaussian_gradient_magnitude(wnf, sigma, mode="nearest")` (scipy 1.7).
 * v f32[D,H,W] -> out f32[D,H,W].  truncate = 4.0 (scipy default), so the filter radius is int(4*sigma + 0.5).
 * Radius <= 4 (sigma = 0.5 ships) runs as ONE fused pass over the volume and tmp may be NULL; larger radii run the
 * nine separable passes and need tmp f32[2*D*H*W]. */
int32_t gnb_gaussian_gradient_magnitude(const float* v, int32_t D, int32_t H, int32_t W, double sigma,
                                        float* out, float* tmp, void* stream);
/* nvol independent volumes v f32[nvol,D,H,W] in one launch; tmp f32[2*nvol*D*H*W] or NULL as above. */
int32_t gnb_gaussian_gradient_magnitude_batched(const float* v, int32_t nvol, int32_t D, int32_t H, int32_t W,
                                                double sigma, float* out, float* tmp, void* stream);

/* ---- N12: marching cubes -------------------------------------------------------------------
 * ref: predict.py:172-181 `marching_cubes(wnf, level, spacing, gradient_direction, method='lewiner')`
 * + the ggm lookup at trunc(vert/spacing).
 * MC33 structure (face test, interior test, tunnel tilings; PARITY UNPINNED vs scikit-image's tables).
 * Two phases.  gnb_mc_count classifies the (D-1)(H-1)(W-1) cells (one warp walks a strip of volume rows with a two-row
 * window of "value > level" bits: 16-byte loads, neighbour bits by shuffle, one ballot skips the rows the surface does not
 * cross), scans the per-row counts and returns the vertex / face / active-cell totals in counts_host[3]
 * (it synchronises `stream`).  gnb_mc_emit (n_active, n_verts = those totals) compacts the active cells, then writes
 *   verts f32[V,3] (axis0,axis1,axis2)*spacing, faces i32[F,3], normals f32[V,3], values f32[V],
 *   ggm_at_verts f32[V] (ggm may be NULL)
 * with one thread per vertex / per active cell.  normals and / or values may be NULL: they are then not computed (the
 * reference stores them in prediction.zarr, predict.py:193-200, but nothing downstream reads them).
 * Vertex numbering = first-use order of a sequential axis0->axis1->axis2 cell scan; faces in cell order.
 * ws: workspace of gnb_mc_workspace_bytes(D,H,W) bytes, shared by both calls.  D*H*W < 2^31. */
int64_t gnb_mc_workspace_bytes(int32_t D, int32_t H, int32_t W);
/* Byte offset, inside the workspace, of the 512-byte record {i64 V, i64 F, i64 A, ...} (+256: {u32 enc(min), u32 enc(max)}
 * order-preserving encodings of the data range).  gnb_mc_count with counts_host == NULL is fully asynchronous; a caller that processes
 * many volumes can then fetch all totals with ONE device->host copy instead of one synchronisation per volume. */
int64_t gnb_mc_totals_offset(int32_t D, int32_t H, int32_t W);
int32_t gnb_mc_count(const float* v, int32_t D, int32_t H, int32_t W, float level, void* ws,
                     int64_t* counts_host, void* stream);
int32_t gnb_mc_emit(const float* v, int32_t D, int32_t H, int32_t W, float level, const double* spacing_host,
                    int32_t ascent, const float* ggm, void* ws, int64_t n_active, int64_t n_verts, float* verts,
                    int32_t* faces, float* normals, float* values, float* ggm_at_verts, void* stream);
/* Batch form: N volumes v f32[N,D,H,W] (ggm likewise or NULL), one workspace block per volume `ws_stride` bytes apart
 * (a multiple of 256, >= gnb_mc_workspace_bytes).  gnb_mc_count_batch is asynchronous: it classifies every volume and scans
 * the counts, and leaves one 512-byte record per volume at gnb_mc_totals_offset inside its block:
 *   i64 V, F, A (active cells), vbase, fbase (first row of the volume in the concatenated outputs = exclusive prefix
 *   sums of V and F over the batch); byte 256: u32 enc(min), enc(max).
 * The caller fetches the N records with one strided device->host copy, sizes verts f32[sum V,3], faces i32[sum F,3],
 * normals, values, ggm_at_verts for the whole batch, and calls gnb_mc_emit_batch with max_active = max_i A_i and
 * max_verts = max_i V_i.  Face indices are local to their volume (0 .. V_i-1).  Seven launches per batch. */
int32_t gnb_mc_count_batch(const float* v, int32_t N, int32_t D, int32_t H, int32_t W, float level, void* ws,
                           int64_t ws_stride, void* stream);
int32_t gnb_mc_emit_batch(const float* v, int32_t N, int32_t D, int32_t H, int32_t W, float level,
                          const double* spacing_host, int32_t ascent, const float* ggm, void* ws, int64_t ws_stride,
                          int64_t max_active, int64_t max_verts, float* verts, int32_t* faces, float* normals,
                          float* values, float* ggm_at_verts, void* stream);

/* Host-side diagnostic (no device work): the tiling the kernels apply to ONE cell with the given eight corner values
 * (corner i at (x,y,z) = {000,100,110,010,001,101,111,011}, scikit-image's numbering): cube index, face-test and
 * interior-test decisions and the table entry they select.  code = index | face bits << 8 | tunnel << 14;
 * tri u8[3*14] vertex ids (0..11 cube edges, 12.. extra centre vertices), order u8[14] the ids in first-use order,
 * cen_n u8[2] / cen_loop u8[2*12] the cube edges whose iso-vertices a centre vertex averages.  Used by the CPU tests to
 * compare every (code, decisions) configuration with the oracle without a GPU. */
int32_t gnb_mc_cell_tiling_host(const float* corner_values, float level, int32_t* code_out, int32_t* ntri_out,
                                uint8_t* tri_out, int32_t* nvert_out, uint8_t* order_out, int32_t* ncen_out,
                                uint8_t* cen_n_out, uint8_t* cen_loop_out);

/* ---- next row: mesh clean-up after marching cubes ----------------------
 * ref: common/marching_cubes_util.py:19-35 (inside wnf_to_mesh) and :38-52 (delete_invalid_verts); eval.py:532-546.
 * A face survives iff all three of its vertices are flagged in on_surface u8[V]; vertices no surviving face uses are
 * deleted and the faces re-indexed (ascending original order, like np.unique).  Batched over the packed output of
 * gnb_mc_emit_batch: faces i32[F,3] hold per-sample LOCAL vertex ids, vptr / fptr i64[B+1] (device) are the row
 * offsets of the samples.  Two phases around the one host read of rec i64[2(B+1)] = new vptr | new fptr:
 * gnb_mesh_cleanup_count (mark, prefix sums) and gnb_mesh_cleanup_emit (keep i64[Vnew] = surviving GLOBAL vertex
 * rows in ascending order -- gather any per-vertex array with it -- and out_faces i32[Fnew,3] re-indexed, local). */
int64_t gnb_mesh_cleanup_workspace_bytes(int64_t V, int64_t F);
int32_t gnb_mesh_cleanup_count(const int32_t* faces, const int64_t* fptr, const int64_t* vptr, int32_t B, int64_t V,
                               int64_t F, const uint8_t* on_surface, void* workspace, int64_t* rec, void* stream);
int32_t gnb_mesh_cleanup_emit(const int32_t* faces, const int64_t* fptr, const int64_t* vptr, int32_t B, int64_t V,
                              int64_t F, void* workspace, const int64_t* rec, int64_t* keep, int32_t* out_faces,
                              void* stream);

/* Largest connected component of the (cleaned) meshes, ref eval.py:538-546 (igl.adjacency_matrix +
 * igl.connected_components + argmax of the component sizes).  Vertices are connected through faces; a component is
 * named by its lowest vertex id (the order igl numbers components in), "largest" = most vertices, first on ties.
 * Batched over packed meshes like the clean-up above.  parent_ws / size_ws: i32[V] scratch.  Outputs (each nullable):
 * is_largest u8[V] membership mask (feed it to gnb_mesh_cleanup_*), labels i32[V] = LOCAL id of the lowest vertex of the
 * vertex's component, summary i64[B,3] = {number of components, size of the largest, local id of its lowest vertex}. */
int32_t gnb_mesh_components(const int32_t* faces, const int64_t* fptr, const int64_t* vptr, int32_t B, int64_t V, int64_t F,
                            int32_t* parent_ws, int32_t* size_ws, uint8_t* is_largest, int32_t* labels, int64_t* summary,
                            void* stream);

/* Area-weighted surface sampling, ref common/geometry_util.py:184-223 (mesh_sample_barycentric) for ONE mesh: face m is
 * drawn with probability area_m / sum(area) by inverse CDF (numpy RandomState.choice: cdf = cumsum(p) / cumsum(p)[-1],
 * searchsorted(cdf, u, side='right')), u_face f64[M] and uv f64[M,2] being the uniform variates of the reference's own
 * host generator (RandomState(seed).random_sample / .uniform); barycentric (a, b, 1-a-b) with (a, b) reflected into the
 * triangle when a + b >= 1.  verts f32 or f64 [N,3] (verts_f64 selects), faces i32[F,3]; areas_in f64[F] optional
 * (else twice the triangle areas are computed: igl.doublearea).  cdf_ws: f64[2F] scratch.  Outputs face_idx i64[M],
 * bary f64[M,3]. */
int32_t gnb_mesh_sample_barycentric(const void* verts, int32_t verts_f64, const int32_t* faces, int64_t F, const double* areas_in,
                                    const double* u_face, const double* uv, int64_t M, double* cdf_ws, int64_t* face_idx,
                                    double* bary, void* stream);
/* ref common/geometry_util.py:160-181 (barycentric_interpolation): out[m,c] = sum_i bary[m,i] * field[faces[face_idx[m],i], c],
 * accumulated in the field's dtype (f32 or f64, field_f64 selects) like the reference's in-place += on verts.dtype. */
int32_t gnb_barycentric_interpolation(const double* bary, const int64_t* face_idx, const int32_t* faces, const void* field,
                                      int32_t field_f64, int32_t C, int64_t M, void* out, void* stream);

/* ---- next row: chamfer / hybrid chamfer nearest-neighbour core -----------
 * ref: eval.py:259-271 (get_chamfer in compute_chamfer) and :381-401 (get_chamfer in compute_hybrid_chamfer), which use
 * scipy cKDTree.query(k=1).  For every query point q_i of sample b (rows ptr_q[b]..ptr_q[b+1]-1 of q f32[.,3]) the
 * nearest reference point r_j of the same sample (ties -> lowest index) -> idx i64 (local index j, nullable),
 * dist f64 (nullable) and sums f64[B] = sum_i dist_i (nullable; divide by the query count for the chamfer term).
 * dist_i = |q_i - r_j| in double, or |qa_i - rb_j| when qa / rb (f32, same row layout) are given (hybrid chamfer:
 * match in NOCS space, distance in simulation space).  max_q = largest query count of a sample (sizes the grid). */
int32_t gnb_nn1_distance(const float* q, const int64_t* ptr_q, const float* r, const int64_t* ptr_r, int32_t B,
                         int64_t max_q, const float* qa, const float* rb, int64_t* idx, double* dist, double* sums,
                         void* stream);

#if defined(__GNUC__)
/* ---- PointConv message MLP + max aggregation in one kernel (ref components/pointnet2.py:30-31 with
 * local_nn = MLP([Cin+3, C1, C2, C3]), components/mlp.py:9-20; aggr = 'max') -----------------------------------------
 * Replaces gnb_pointconv_gather + 3 x gnb_linear_tc + gnb_segment_max for the two local set-abstraction levels: the [E, C]
 * activations of the edge MLP never reach HBM (11 GB per step at batch 32).  Supported widths (gnb_pointconv_mlp_supported):
 * 3+3 -> 64 -> 64 -> 128 (SA1) and 128+3 -> 128 -> 128 -> 256 (SA2).
 *   gnb_pointconv_mlp_pack: W1 [C1, Cin+3], W2' [C2, C1], W3' [C3, C2] fp32 (BatchNorm1 / 2 folded into W2' / W3' by the caller:
 *     they follow a ReLU) -> fp16 hi/lo shared-memory images, scaled by 2^s1 / 2^s2 / 2^s3; packed: gnb_pointconv_mlp_packed_bytes.
 *   gnb_pointconv_mlp_max: edges = gnb_ball_query output (nbr i64[sumM,K], cnt) with PointConv's self loops, offsets eoffs
 *     i64[sumM+1] from gnb_pointconv_edge_count + scan (as for gnb_pointconv_gather); consts f32 = [b1 C1 | w1 KIN x C1 | b2' C2 |
 *     b3' C3 | bn3_scale C3 | bn3_shift C3] where w1 holds, input-major, the layer-1 weights applied in fp32 (SA1: all 6 inputs,
 *     SA2: the 3 relative-position inputs, KIN = 3); edge_ws i32[2 * sumM * (K+1)] scratch; out f32[sumM, C3]. */
int32_t gnb_pointconv_mlp_supported(int32_t Cin, int32_t C1, int32_t C2, int32_t C3);
int64_t gnb_pointconv_mlp_packed_bytes(int32_t Cin, int32_t C1, int32_t C2, int32_t C3);
int32_t gnb_pointconv_mlp_pack(const float* W1, const float* W2, const float* W3, int32_t Cin, int32_t C1, int32_t C2,
                               int32_t C3, int32_t s1, int32_t s2, int32_t s3, void* packed, void* stream);
int32_t gnb_pointconv_mlp_max(const float* x, int64_t ldx, int32_t Cin, const float* pos_x, const float* pos_y,
                              const int64_t* nbr, const int32_t* cnt, const int64_t* eoffs, int64_t sumM, int32_t K,
                              const void* packed, int32_t C1, int32_t C2, int32_t C3, int32_t s1, int32_t s2, int32_t s3,
                              const float* consts, int32_t* edge_ws, float* out, void* stream);

/* ---- fp16 operand range flag --------------------------------------------------------------------------------------
 * The tensor-core kernels split fp32 operands into fp16 hi + lo with saturating conversions (+-65504).  With trained
 * checkpoints an activation outside that range would silently be clamped; one flag per device records it instead:
 *   - gnb_gn_apply_split sets it for the operands of the 3D-UNet convolutions (no extra pass);
 *   - gnb_f16_range_check(x, n) sets it when any of the n fp32 values is outside the range (or not finite): the pipeline
 *     runs it on the grids the implicit decoders interpolate (a convex combination never exceeds its inputs);
 *   - not covered: the inputs of the PointNet++ Linear blocks (gnb_linear_tc).
 * gnb_f16_overflow_fetch synchronises `stream`, returns 1 if the flag is set (0 otherwise, < 0 on error) and clears it when
 * reset != 0. */
int32_t gnb_f16_range_check(const float* x, int64_t n, void* stream);
int32_t gnb_f16_overflow_fetch(int32_t reset, void* stream);
/* Stream-ordered form: stores the flag into pinned host memory from a one-thread kernel (valid once `stream` has reached this point, e.g. after the
 * caller's next synchronisation) and clears it when reset != 0; no synchronisation of its own. */
int32_t gnb_f16_overflow_fetch_async(uint32_t* pinned_host_out, int32_t reset, void* stream);

/* Small stream-ordered device -> host transfer WITHOUT the copy engine: a kernel stores `rows` records of `width_bytes`
 * (pitches in bytes; everything a multiple of 4) straight into pinned, device-addressable host memory.  For the few hundred
 * bytes the host is waiting for in the middle of a batch (the marching-cubes totals, predict.py:172-177 hands the same
 * numbers back through skimage's return values): a cudaMemcpy of that size queues behind the bulk result transfers of the
 * previous batch on the copy engine -- milliseconds when eight ranks share one host.  Valid on the host once `stream` has
 * reached this point.  GNB_ERR_INVALID for pageable destinations and for transfers above 16 MiB. */
int32_t gnb_copy_to_pinned_host(const void* src, int64_t src_pitch, void* pinned_host_dst, int64_t dst_pitch,
                                int64_t width_bytes, int64_t rows, void* stream);

#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* GARMENTNETS_B200_H */
