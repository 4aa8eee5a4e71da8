"""Synthetic inputs and weights for tests and benchmarks (no dataset / checkpoint is reachable offline).

* ``make_cloud`` / ``make_batch``: seeded garment-like point clouds per SURVEY.md section 8(d) config 5 -- thin
  two-sheet shells hanging below the gripper (origin) in the gripper frame, metres, rotated about z, rgb ~ U(0,1);
  the input contract of ref datasets/conv_implicit_wnf_dataset.py:205-228 (x = rgb, pos = sim points, batch sorted).
* ``HPARAMS``: the only shipped hyper-parameters (ref config/train_pointnet2_default.yaml:30-43,
  config/train_pipeline_default.yaml:40-61, config/predict_default.yaml:43-47).
* ``randomize_``: seeded non-trivial BatchNorm running statistics / affine parameters on top of torch's default
  init, so that eval-mode BN and GroupNorm affine are actually exercised.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

CATEGORIES = ["Dress", "Tshirt", "Trousers", "Jumpsuit", "Skirt", "Top"]
# (x half extent, y half extent, z length) in metres
EXTENTS = {
    "Dress": (0.30, 0.12, 1.10), "Tshirt": (0.35, 0.10, 0.70), "Trousers": (0.22, 0.10, 1.00),
    "Jumpsuit": (0.28, 0.12, 1.40), "Skirt": (0.30, 0.15, 0.60), "Top": (0.30, 0.10, 0.50),
}

HPARAMS = {
    "pointnet2": dict(feature_dim=128, batch_norm=True, dropout=True, sa1_ratio=0.5, sa1_r=0.05, sa2_ratio=0.25,
                      sa2_r=0.1, fp3_k=1, fp2_k=3, fp1_k=3, nocs_bins=64),
    "volume_agg": dict(nn_channels=[137, 137, 128], batch_norm=True, lower_corner=(0, 0, 0), upper_corner=(1, 1, 1),
                       grid_shape=(32, 32, 32), reduce_method="max", include_point_feature=True,
                       include_confidence_feature=True),
    "unet3d": dict(in_channels=128, out_channels=128, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4),
    "volume_decoder": dict(nn_channels=(128, 256, 256, 1), batch_norm=True),
    "surface_decoder": dict(nn_channels=(128, 256, 256, 3), batch_norm=True),
    "prediction": dict(volume_size=128, gradient_sigma=0.5, iso_surface_level=0.5, gradient_direction="ascent"),
}


def make_cloud(category: str = "Tshirt", n: int = 4096, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Return (pos f32[n,3], rgb f32[n,3]) for one synthetic hanging garment."""
    cat_id = CATEGORIES.index(category)
    rng = np.random.default_rng(1000 + cat_id + 7919 * seed)
    hx, hy, zl = EXTENTS[category]
    u = rng.uniform(-1.0, 1.0, n)
    t = rng.uniform(0.0, 1.0, n)
    side = rng.integers(0, 2, n) * 2 - 1
    # width tapers towards the grasp point, the two sheets bulge apart in the middle and wrinkle a little
    width = hx * (0.25 + 0.75 * np.sqrt(t))
    x = u * width
    bulge = hy * np.sin(np.pi * t) * np.sqrt(np.clip(1.0 - u * u, 0.0, 1.0))
    y = side * (0.004 + bulge) + 0.004 * np.sin(18.0 * x + 5.0 * t)
    z = -zl * t
    theta = rng.uniform(-np.pi, np.pi)
    c, s = np.cos(theta), np.sin(theta)
    pos = np.stack([c * x - s * y, s * x + c * y, z], axis=1).astype(np.float32)
    rgb = rng.uniform(0.0, 1.0, (n, 3)).astype(np.float32)
    return pos, rgb


def make_batch(B: int, n: int = 4096, category: str = "Tshirt", seed: int = 0) -> Dict[str, np.ndarray]:
    """PyG-style flat batch: x [B*n,3] rgb, pos [B*n,3], batch i64[B*n] sorted."""
    ps, xs = zip(*[make_cloud(category, n, seed * 1000 + b) for b in range(B)])
    return {"x": np.concatenate(xs), "pos": np.concatenate(ps),
            "batch": np.repeat(np.arange(B, dtype=np.int64), n)}


def randomize_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Seeded perturbation of norm layers (BN running stats ~ N(0,0.1) / U(0.5,1.5); affine ~ N(1,0.1) / N(0,0.1))."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, torch.nn.GroupNorm):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    return module


def build_pipeline(seed: int = 0, device=None, hparams=None):
    """Seeded random-init pipeline with the shipped architecture (eval mode, deterministic FPS start)."""
    from .pipeline import ConvImplicitWNFPipeline
    hp = hparams or HPARAMS
    torch.manual_seed(seed)
    model = ConvImplicitWNFPipeline.from_hparams(hp)
    randomize_(model, seed + 1)
    model.eval().requires_grad_(False)
    model.pointnet2_nocs.set_random_start(False)
    if device is not None:
        model = model.to(device)
    return model


@torch.no_grad()
def calibrate_bn_(model, data, index=None, n_queries: int = 4096, seed: int = 0):
    """Give every BatchNorm of the random-init pipeline the running statistics of the activations it actually sees
    on ``data`` (what training would have produced).  Without this a random deep point network is numerically
    almost constant across points: every point lands in the same NOCS bin / voxel and the pipeline degenerates."""
    from .components.mlp import _Block
    g = torch.Generator().manual_seed(seed)
    _Block.calibrating = True
    try:
        p = model.pointnet2_forward(data, index=index)
        u = model.unet3d_forward(p)
        B = u["out_feature_volume"].shape[0]
        q = torch.rand(B, n_queries, 3, generator=g).to(u["out_feature_volume"].device)
        model.volume_decoder_forward(u, q)
        model.surface_decoder_forward(u, q)
    finally:
        _Block.calibrating = False


def prepare_model_(model, data, index=None, inside_fraction: float = 0.12):
    """Seeded weights -> usable synthetic network: BN statistics from data, then the WNF level calibration."""
    calibrate_bn_(model, data, index)
    return calibrate_wnf_(model, data, index, inside_fraction)


@torch.no_grad()
def calibrate_wnf_(model, data, index=None, inside_fraction: float = 0.12, level: float = 0.5, probe_size: int = 32):
    """Random weights give a winding-number field that never crosses the iso level, so marching cubes would find
    nothing.  Re-centre the LAST BatchNorm of ``volume_decoder`` (running_mean := the (1-inside_fraction) quantile of
    its input on a probe lattice of sample 0, var := 1, gamma := 1, beta := level) so that ``inside_fraction`` of the
    volume lies above ``level``.  Deterministic for fixed seeds; the calibrated state_dict is what both the CUDA path
    and the oracle consume."""
    bn = model.volume_decoder.mlp[-1][2]
    bn.running_mean.zero_()
    bn.running_var.fill_(1.0 - bn.eps)
    bn.weight.fill_(1.0)
    bn.bias.zero_()
    p = model.pointnet2_forward(data, index=index)
    u = model.unet3d_forward(p)
    vol = model.dense_decode(u["out_feature_volume"][:1], probe_size)
    vals = vol.reshape(-1).float().cpu()
    k = max(1, int(round((1.0 - inside_fraction) * vals.numel())))
    q = torch.kthvalue(vals, k).values.item()
    if not q > 0:  # more than (1-inside_fraction) of the ReLU outputs are zero: put the level just above zero
        q = float(vals[vals > 0].min().item()) if bool((vals > 0).any()) else 0.0
    bn.running_mean.fill_(q)
    bn.bias.fill_(level)
    return q
