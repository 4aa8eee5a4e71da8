"""Oracle (TEST INFRASTRUCTURE): the reference's per-sample predict loop (predict.py:138-187) composed from the
stage restatements in ``oracle.nets`` / ``oracle.pointops`` / ``oracle.postproc``, driven by a flat state_dict with
the reference's ``ConvImplicitWNFPipeline`` key names.  Also the thing ``bench.py`` times as the CPU baseline."""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np
import torch

from . import nets as N
from . import pointops as P
from . import postproc


def to_cpu_state_dict(module_or_sd) -> Dict[str, torch.Tensor]:
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    return {k: v.detach().to("cpu") for k, v in sd.items()}


def stage1(sd, hp, x, pos, batch, B, fps_starts=None):
    """pointnet2_forward (networks/conv_implicit_wnf.py:213-240)."""
    ptr = P.batch_to_ptr(np.asarray(batch), B)
    res = N.pointnet2_nocs_forward(sd, hp["pointnet2"], x, pos, ptr, fps_starts, prefix="pointnet2_nocs.")
    bins = hp["pointnet2"]["nocs_bins"]
    b, conf, nocs = N.nocs_head(res["per_point_logits"], bins)
    res.update({"nocs_bin": b, "pred_confidence": conf, "pred_nocs": nocs, "ptr": ptr})
    return res


def stage2(sd, hp, s1, pos, batch, B):
    """unet3d_forward (networks/conv_implicit_wnf.py:242-251)."""
    G = hp["volume_agg"]["grid_shape"][0]
    vol_in, flat, feats, h = N.volume_feature_aggregator(sd, "volume_agg.", s1["per_point_features"], s1["pred_nocs"],
                                                         pos, s1["pred_confidence"], batch, B, G)
    out = N.unet3d_forward(sd, "unet_3d.abstract_3d_unet.", vol_in, hp["unet3d"]["num_levels"], hp["unet3d"]["num_groups"])
    return {"in_feature_volume": vol_in, "out_feature_volume": out.numpy(), "flat_idx": flat, "agg_features": feats}


def predict_sample(sd, hp, x, pos, fps_starts=None, max_chunks: Optional[int] = None, timings: Optional[dict] = None):
    """One garment through predict.py:138-187.  ``max_chunks`` bounds the dense decode (bench sampling)."""
    pr = hp["prediction"]
    n = len(pos)
    batch = np.zeros(n, dtype=np.int64)
    t0 = time.perf_counter()
    s1 = stage1(sd, hp, x, pos, batch, 1, fps_starts)
    t1 = time.perf_counter()
    s2 = stage2(sd, hp, s1, pos, batch, 1)
    t2 = time.perf_counter()
    wnf = N.dense_decode(sd, "volume_decoder.", s2["out_feature_volume"], pr["volume_size"], 64, max_chunks).numpy()
    t3 = time.perf_counter()
    out = {"stage1": s1, "stage2": s2, "wnf_volume": wnf}
    if max_chunks is None:
        tail = postproc.predict_tail(wnf, pr["gradient_sigma"], pr["iso_surface_level"], pr["gradient_direction"])
        t4 = time.perf_counter()
        q = torch.from_numpy(tail["verts"].astype(np.float32)).view(1, -1, 3)
        warp = N.implicit_decoder(sd, "surface_decoder.", s2["out_feature_volume"], q).view(-1, 3).numpy()
        t5 = time.perf_counter()
        tail["warp_field"] = warp.astype(np.float32)
        out["mesh"] = tail
    else:
        t4 = t5 = t3
    if timings is not None:
        timings.update({"pointnet2": t1 - t0, "unet3d": t2 - t1, "dense_decode": t3 - t2, "ggm_mc": t4 - t3,
                        "surface_decode": t5 - t4})
    return out
