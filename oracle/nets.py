"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the network stages of the hot path, driven by a flat
``state_dict`` that uses the reference's parameter names.

ATen-backed arithmetic (Linear, BatchNorm eval, GroupNorm, Conv3d, max_pool3d, nearest interpolate, grid_sample)
runs through torch CPU ops, which have the semantics the reference relies on; the composition follows
    components/mlp.py:3-20                        -> mlp
    components/pointnet2.py:22-33,44-52,70-76     -> sa_module / global_sa_module / fp_module
    networks/pointnet2_nocs.py:134-166            -> pointnet2_nocs_forward
    networks/conv_implicit_wnf.py:213-240         -> nocs_head
    networks/conv_implicit_wnf.py:43-100          -> volume_feature_aggregator
    components/unet3d.py:43-72,195-293,449-474    -> unet3d_forward   (PINNED against the reference module, tests/golden)
    networks/conv_implicit_wnf.py:128-149         -> implicit_decoder (PINNED: F.grid_sample + reference MLP)
    predict.py:145-158                            -> dense_decode
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import pointops as P

SD = Dict[str, torch.Tensor]

# When True, ``mlp`` overwrites each BatchNorm's running statistics in ``sd`` with the statistics of its input (the
# CPU twin of garmentnets_b200.synthetic.calibrate_bn_, used by ``bench.py --impl reference`` which has no GPU model).
CALIBRATE = False


def _t(a) -> torch.Tensor:
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))


def mlp(sd: SD, prefix: str, x: torch.Tensor, batch_norm: bool = True) -> torch.Tensor:
    """components/mlp.py:9-20 -- [Linear, ReLU, PointBatchNorm1D] per layer, ReLU+BN also after the last layer;
    BN in eval mode (running statistics, eps 1e-5)."""
    layer = 0
    shape = x.shape
    x = x.reshape(-1, shape[-1])
    while f"{prefix}{layer}.0.weight" in sd:
        p = f"{prefix}{layer}."
        x = F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"])
        x = F.relu(x)
        if batch_norm and (p + "2.weight") in sd:
            if CALIBRATE:  # synthetic-weight support: running statistics := statistics of the activations seen
                sd[p + "2.running_mean"] = x.mean(0)
                sd[p + "2.running_var"] = x.var(0, unbiased=False).clamp_min(1e-2)
            x = F.batch_norm(x, sd[p + "2.running_mean"], sd[p + "2.running_var"], sd[p + "2.weight"],
                             sd[p + "2.bias"], training=False, eps=1e-5)
        layer += 1
    return x.reshape(*shape[:-1], x.shape[-1])


def sa_module(sd: SD, prefix: str, x, pos, ptr, ratio: float, r: float, start=None):
    """components/pointnet2.py:22-33.  x [N,C] or None, pos [N,3] (numpy f32), ptr i64[B+1]."""
    idx = P.fps(pos, ptr, ratio, start)
    counts = P.fps_counts(ptr, ratio)
    ptr_y = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    pos_y = pos[idx]
    nbr, cnt = P.ball_query(pos, pos_y, ptr, ptr_y, r, 64)
    offs, src = P.pointconv_edges(nbr, cnt)
    edge = P.pointconv_edge_features(x, pos, pos_y, offs, src)
    msg = mlp(sd, prefix + "conv.local_nn.", _t(edge)).numpy()
    out = P.segment_max(msg, offs)
    aux = {"idx": idx, "nbr": nbr, "cnt": cnt, "eoffs": offs, "esrc": src}
    return out, pos_y, ptr_y, aux


def global_sa_module(sd: SD, prefix: str, x, pos, ptr):
    """components/pointnet2.py:44-52."""
    h = mlp(sd, prefix + "nn.", _t(np.concatenate([x, pos], axis=1))).numpy()
    B = len(ptr) - 1
    out = np.zeros((B, h.shape[1]), dtype=np.float32)
    for b in range(B):
        if ptr[b + 1] > ptr[b]:
            out[b] = h[ptr[b]:ptr[b + 1]].max(axis=0)
    return out, np.zeros((B, 3), dtype=np.float32), np.arange(B + 1, dtype=np.int64)


def fp_module(sd: SD, prefix: str, k: int, x, pos, ptr, x_skip, pos_skip, ptr_skip):
    """components/pointnet2.py:70-76."""
    idx, d2 = P.knn(pos, pos_skip, ptr, ptr_skip, k)
    y = P.knn_interpolate(x, idx, d2)
    if x_skip is not None:
        y = np.concatenate([y, x_skip], axis=1)
    return mlp(sd, prefix + "nn.", _t(y)).numpy(), {"knn_idx": idx, "knn_d2": d2, "interp": y}


def pointnet2_nocs_forward(sd: SD, hp: dict, x, pos, ptr, starts=None, prefix: str = ""):
    """networks/pointnet2_nocs.py:134-166 (eval: dropout is the identity)."""
    x = np.ascontiguousarray(x, np.float32)
    pos = np.ascontiguousarray(pos, np.float32)
    s1 = s2 = None
    if starts is not None:
        s1, s2 = starts
    sa1_x, sa1_pos, sa1_ptr, aux1 = sa_module(sd, prefix + "sa1_module.", x, pos, ptr, hp["sa1_ratio"], hp["sa1_r"], s1)
    sa2_x, sa2_pos, sa2_ptr, aux2 = sa_module(sd, prefix + "sa2_module.", sa1_x, sa1_pos, sa1_ptr, hp["sa2_ratio"],
                                              hp["sa2_r"], s2)
    sa3_x, sa3_pos, sa3_ptr = global_sa_module(sd, prefix + "sa3_module.", sa2_x, sa2_pos, sa2_ptr)
    fp3_x, _ = fp_module(sd, prefix + "fp3_module.", hp["fp3_k"], sa3_x, sa3_pos, sa3_ptr, sa2_x, sa2_pos, sa2_ptr)
    fp2_x, _ = fp_module(sd, prefix + "fp2_module.", hp["fp2_k"], fp3_x, sa2_pos, sa2_ptr, sa1_x, sa1_pos, sa1_ptr)
    fp1_x, _ = fp_module(sd, prefix + "fp1_module.", hp["fp1_k"], fp2_x, sa1_pos, sa1_ptr, x, pos, ptr)
    h = _t(fp1_x)
    h = F.relu(F.linear(h, sd[prefix + "lin1.weight"], sd[prefix + "lin1.bias"]))
    features = F.linear(h, sd[prefix + "lin2.weight"], sd[prefix + "lin2.bias"])
    logits = F.linear(features, sd[prefix + "lin3.weight"], sd[prefix + "lin3.bias"])
    g = F.relu(_t(sa3_x))
    g = F.linear(g, sd[prefix + "global_lin1.weight"], sd[prefix + "global_lin1.bias"])
    global_logits = F.linear(g, sd[prefix + "global_lin2.weight"], sd[prefix + "global_lin2.bias"])
    return {
        "per_point_features": features.numpy(), "per_point_logits": logits.numpy(),
        "global_logits": global_logits.numpy(), "global_feature": sa3_x,
        "sa1": (sa1_x, sa1_pos, sa1_ptr, aux1), "sa2": (sa2_x, sa2_pos, sa2_ptr, aux2),
        "fp3_x": fp3_x, "fp2_x": fp2_x, "fp1_x": fp1_x,
    }


def nocs_head(logits: np.ndarray, bins: int):
    """networks/conv_implicit_wnf.py:222-231: reshape (N,bins,3), argmax / softmax over bins, bin -> NOCS point."""
    lg = _t(logits).reshape(-1, bins, 3)
    b = torch.argmax(lg, dim=1)
    conf = torch.gather(F.softmax(lg, dim=1), 1, b.unsqueeze(1)).squeeze(1)
    scale = torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(bins - 1), dtype=torch.float32)
    nocs = b.to(torch.float32) * scale
    return b.numpy(), conf.numpy(), nocs.numpy()


def points_grid_idxs(points: np.ndarray, G: int, lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 1.0)) -> np.ndarray:
    """components/gridding.py:161-186 (lower corner 0 / upper corner 1 unless given)."""
    p = _t(points)
    lc, uc = torch.tensor(lower, dtype=torch.float32), torch.tensor(upper, dtype=torch.float32)
    scales = (torch.tensor([G] * 3, dtype=torch.float32) - 1) / (uc - lc)
    f = (p + (-lc)) * scales
    return torch.clamp(f.to(torch.int64), 0, G - 1).numpy()


def aggregator_rows(feat, nocs, sim_points, conf, batch, G: int, lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 1.0),
                    include_point_feature: bool = True, include_confidence_feature: bool = True):
    """networks/conv_implicit_wnf.py:62-85: flat voxel index and the concatenated per-point feature rows."""
    idx3 = points_grid_idxs(nocs, G, lower, upper)
    flat = (_t(batch).to(torch.int64) * G ** 3 + _t(idx3[:, 0]) * G ** 2 + _t(idx3[:, 1]) * G + _t(idx3[:, 2])).numpy()
    parts = [_t(feat)]
    if include_point_feature:
        lc, uc = torch.tensor(lower, dtype=torch.float32), torch.tensor(upper, dtype=torch.float32)
        scales = (uc - lc) / (torch.tensor([G] * 3, dtype=torch.float32) - 1)   # components/gridding.py:249-255
        origin = _t(idx3) * scales + lc
        parts += [_t(nocs) - origin, _t(sim_points)]
    if include_confidence_feature:
        parts.append(_t(conf))
    return flat, torch.cat(parts, dim=-1)


def volume_feature_aggregator(sd: SD, prefix: str, feat, nocs, sim_points, conf, batch, B: int, G: int, reduce: str = "max",
                              **flags):
    """networks/conv_implicit_wnf.py:43-100 (shipped: include_point_feature / include_confidence_feature, reduce max).
    Returns (volume [B,C,G,G,G], flat_idx, pre-MLP features)."""
    flat, feats = aggregator_rows(feat, nocs, sim_points, conf, batch, G, **flags)
    h = mlp(sd, prefix + "local_nn.", feats)
    vol_flat = P.scatter(h.numpy().T, flat, B * G ** 3, reduce)
    C = h.shape[1]
    vol = vol_flat.reshape(C, B, G, G, G).transpose(1, 0, 2, 3, 4)
    return np.ascontiguousarray(vol), flat, feats.numpy(), h.numpy()


def _single_conv(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """components/unet3d.py:43-72 with order 'gcr': GroupNorm -> Conv3d(no bias, pad 1) -> ReLU."""
    C = x.shape[1]
    g = groups if C >= groups else 1
    x = F.group_norm(x, g, sd[p + "groupnorm.weight"], sd[p + "groupnorm.bias"], eps=1e-5)
    x = F.conv3d(x, sd[p + "conv.weight"], None, padding=1)
    return F.relu(x)


def unet3d_forward(sd: SD, prefix: str, x, num_levels: int = 4, groups: int = 8) -> torch.Tensor:
    """components/unet3d.py:449-474 (DoubleConv, nearest upsampling, concat (skip, up), final 1x1x1, no activation)."""
    x = _t(x)
    feats = []
    for i in range(num_levels):
        if i > 0:
            x = F.max_pool3d(x, 2)
        p = f"{prefix}encoders.{i}.basic_module."
        x = _single_conv(sd, p + "SingleConv1.", x, groups)
        x = _single_conv(sd, p + "SingleConv2.", x, groups)
        feats.insert(0, x)
    feats = feats[1:]
    for i, skip in enumerate(feats):
        x = F.interpolate(x, size=skip.shape[2:], mode="nearest")
        x = torch.cat((skip, x), dim=1)
        p = f"{prefix}decoders.{i}.basic_module."
        x = _single_conv(sd, p + "SingleConv1.", x, groups)
        x = _single_conv(sd, p + "SingleConv2.", x, groups)
    return F.conv3d(x, sd[prefix + "final_conv.weight"], sd[prefix + "final_conv.bias"])


def implicit_decoder(sd: SD, prefix: str, features_grid, query_points) -> torch.Tensor:
    """networks/conv_implicit_wnf.py:128-149.  features_grid [B,C,D,H,W], query_points [B,M,3] -> [B,M,Cout]."""
    fg, q = _t(features_grid), _t(query_points)
    qn = 2.0 * q - 1.0
    s = F.grid_sample(fg, qn.view(*qn.shape[:2], 1, 1, 3), mode="bilinear", padding_mode="border", align_corners=True)
    s = s.view(s.shape[:3]).permute(0, 2, 1)
    return mlp(sd, prefix + "mlp.", s)


def grid_points(Q: int) -> torch.Tensor:
    """components/gridding.py:139-159 (include_batch=False): [Q,Q,Q,3] fp32, point = idx * (1/(Q-1))."""
    ax = torch.arange(Q, dtype=torch.int64)
    idx = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1)
    scales = (torch.ones(3) - torch.zeros(3)) / (torch.tensor([Q] * 3, dtype=torch.float32) - 1)
    return idx.to(torch.float32) * scales + (-torch.zeros(3))


def dense_decode(sd: SD, prefix: str, features_grid, Q: int = 128, chunk: int = 64, max_chunks: Optional[int] = None):
    """predict.py:145-158: Q^3 lattice decoded in chunk^3 blocks (ArraySlicer order: last axis fastest).
    features_grid [1,C,D,H,W].  Returns the [Q,Q,Q] volume (zeros where chunks were skipped by max_chunks)."""
    gp = grid_points(Q)
    out = torch.zeros((Q, Q, Q), dtype=torch.float32)
    n = -(-Q // chunk)
    done = 0
    for a in range(n):
        for b in range(n):
            for c in range(n):
                if max_chunks is not None and done >= max_chunks:
                    return out
                sl = (slice(a * chunk, min(Q, (a + 1) * chunk)), slice(b * chunk, min(Q, (b + 1) * chunk)),
                      slice(c * chunk, min(Q, (c + 1) * chunk)))
                q = gp[sl]
                val = implicit_decoder(sd, prefix, features_grid, q.reshape(1, -1, 3))
                out[sl] = val.view(*q.shape[:-1])
                done += 1
    return out
