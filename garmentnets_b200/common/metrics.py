"""Chamfer metrics -- the nearest-neighbour cores of the reference's ``eval.py`` (``get_chamfer`` inside
``compute_chamfer`` :259-271 and inside ``compute_hybrid_chamfer`` :381-401) on the device.

The reference builds two scipy cKDTrees per sample and queries 10 k points each way; here one brute-force kernel
(``gnb_nn1_distance``) serves a whole batch of point-set pairs.  Same result keys as the reference's dicts.  Inputs are
CUDA float32 tensors; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib, ops


def _pack(sets: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    sizes = np.array([int(s.shape[0]) for s in sets], dtype=np.int64)
    host = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    flat = torch.cat([ops._req(s, torch.float32, "points").reshape(-1, 3) for s in sets]) if len(sets) > 1 else \
        ops._req(sets[0], torch.float32, "points").reshape(-1, 3)
    return flat.contiguous(), torch.from_numpy(host).to(flat.device), sizes


def nearest_neighbor(query: Sequence[torch.Tensor], ref: Sequence[torch.Tensor],
                     query_other: Optional[Sequence[torch.Tensor]] = None, ref_other: Optional[Sequence[torch.Tensor]] = None):
    """For every point of ``query[b]`` its nearest point in ``ref[b]`` (cKDTree.query(k=1) semantics, ties -> lowest index).
    Returns ``(idx i64 [sum Nq], dist f64 [sum Nq], mean f64 [B], sizes)``; with ``query_other`` / ``ref_other`` the distance
    is measured between the corresponding points of those sets instead (hybrid chamfer)."""
    q, ptr_q, nq = _pack(query)
    r, ptr_r, nr = _pack(ref)
    if np.any((nr == 0) & (nq > 0)):
        raise ValueError("nearest_neighbor: empty reference set")
    B = len(nq)
    qa = rb = None
    if query_other is not None:
        qa, _, na = _pack(query_other)
        rb, _, nb = _pack(ref_other)
        if not (np.array_equal(na, nq) and np.array_equal(nb, nr)):
            raise ValueError("nearest_neighbor: the *_other sets must match the sizes of query / ref")
    dev = q.device
    idx = torch.empty((q.shape[0],), dtype=torch.int64, device=dev)
    dist = torch.empty((q.shape[0],), dtype=torch.float64, device=dev)
    sums = torch.empty((B,), dtype=torch.float64, device=dev)
    _lib.call("gnb_nn1_distance", q.data_ptr(), ptr_q.data_ptr(), r.data_ptr(), ptr_r.data_ptr(), B, int(nq.max()) if B else 0,
              ops._ptr(qa), ops._ptr(rb), idx.data_ptr(), dist.data_ptr(), sums.data_ptr(), ops._stream())
    mean = sums / torch.from_numpy(np.maximum(nq, 1).astype(np.float64)).to(dev)
    return idx, dist, mean, nq


def chamfer_batch(pred_points: Sequence[torch.Tensor], gt_points: Sequence[torch.Tensor]) -> List[Dict[str, float]]:
    """ref eval.py:259-271 for a batch of (pred, gt) point-set pairs; one host read for all results."""
    _, _, fwd, _ = nearest_neighbor(pred_points, gt_points)
    _, _, bwd, _ = nearest_neighbor(gt_points, pred_points)
    both = torch.stack([fwd, bwd]).cpu().numpy()
    return [{"chamfer_forward": float(f), "chamfer_backward": float(b), "chamfer_symmetrical": float(np.mean([f, b]))}
            for f, b in both.T]


def chamfer(pred_points: torch.Tensor, gt_points: torch.Tensor) -> Dict[str, float]:
    return chamfer_batch([pred_points], [gt_points])[0]


def hybrid_chamfer_batch(pred_nocs: Sequence[torch.Tensor], gt_nocs: Sequence[torch.Tensor], pred_sim: Sequence[torch.Tensor],
                         gt_sim: Sequence[torch.Tensor]) -> List[Dict[str, float]]:
    """ref eval.py:381-401: nearest neighbours in NOCS space, distances between the matched simulation-space points."""
    _, _, fwd, _ = nearest_neighbor(pred_nocs, gt_nocs, pred_sim, gt_sim)
    _, _, bwd, _ = nearest_neighbor(gt_nocs, pred_nocs, gt_sim, pred_sim)
    both = torch.stack([fwd, bwd]).cpu().numpy()
    return [{"hybrid_chamfer_forward": float(f), "hybrid_chamfer_backward": float(b),
             "hybrid_chamfer_symmetrical": float(np.mean([f, b]))} for f, b in both.T]


def hybrid_chamfer(pred_nocs_points, gt_nocs_points, pred_sim_points, gt_sim_points) -> Dict[str, float]:
    return hybrid_chamfer_batch([pred_nocs_points], [gt_nocs_points], [pred_sim_points], [gt_sim_points])[0]


def optimal_gradient_threshold(gt_mc_verts: torch.Tensor, gt_is_on_surface: torch.Tensor, pred_mc_verts: torch.Tensor,
                               pred_mc_gm: torch.Tensor, precision_weight: float = 0.85) -> Dict[str, float]:
    """ref eval.py:58-102 (``compute_optimal_gradient_treshold``): label every predicted marching-cubes vertex with the
    on-surface flag of its nearest ground-truth vertex (``gnb_nn1_distance``), then pick the gradient-magnitude threshold
    (a decision stump over the sorted magnitudes) that maximises ``precision * w + recall * (1 - w)``.  The sort / prefix
    sums run on the device in float64 like numpy's; one host read for the result."""
    idx, _, _, _ = nearest_neighbor([pred_mc_verts], [gt_mc_verts])
    nn_is_on_surface = ops._req(gt_is_on_surface, torch.bool, "gt_is_on_surface")[idx]
    gm = ops._req(pred_mc_gm, torch.float32, "pred_mc_gm")
    sorted_idx = torch.argsort(gm, stable=True)
    s = nn_is_on_surface[sorted_idx]
    false_negative = torch.cumsum(s.to(torch.int64), 0)
    true_positive = torch.flip(torch.cumsum(torch.flip(s, [0]).to(torch.int64), 0), [0])
    false_positive = torch.flip(torch.cumsum(torch.flip(~s, [0]).to(torch.int64), 0), [0])
    precision = true_positive.double() / (true_positive + false_positive).double()
    recall = true_positive.double() / (true_positive + false_negative).double()
    score = precision * precision_weight + recall * (1 - precision_weight)
    finite = torch.isfinite(score)
    if bool(finite.any()):
        # np.argmax treats NaN as the maximum; the reference only reaches argmax when some score is finite, and NaN
        # (0/0) can only occur where true_positive + false_positive == 0, which cannot happen (the suffix is never empty)
        max_score_idx = int(torch.argmax(torch.where(finite, score, torch.full_like(score, -float("inf")))).item())
        thr = float(gm[sorted_idx[max_score_idx]].item())
    else:
        thr = float(gm.min().item())
    return {"optimal_wnf_gradient_threshold": thr}
