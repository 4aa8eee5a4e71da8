"""Per-point MLP -- mirror of the reference's ``components/mlp.py`` (``MLP`` :9-20, ``PointBatchNorm1D`` :3-7).

``MLP(channels, batch_norm)`` returns an ``nn.Sequential`` of ``Sequential(Linear, ReLU, PointBatchNorm1D)`` blocks
with the reference's ``state_dict`` keys (``{l}.0.weight``, ``{l}.2.running_mean`` ...).  In eval mode every block is
ONE launch of ``gnb_linear`` (GEMM + bias + ReLU + folded BatchNorm affine).  Training-mode forward is outside the
inference hot path and is refused rather than silently served by another backend.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from .. import ops


class FoldedBatchNorm:
    """Mixin: eval-mode BatchNorm as a per-channel affine that the preceding GEMM's epilogue applies."""

    def folded_affine(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Eval-mode BN as y = x*scale + shift (scale = gamma/sqrt(var+eps), shift = beta - mean*scale); cached
        until a parameter or running statistic changes."""
        key = (self.running_mean._version, self.running_var._version, self.weight._version, self.bias._version,
               self.running_mean.data_ptr(), self.weight.data_ptr())
        cached = getattr(self, "_gnb_fold", None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                scale = self.weight * torch.rsqrt(self.running_var + self.eps)
                shift = self.bias - self.running_mean * scale
            cached = (key, scale.contiguous(), shift.contiguous())
            self._gnb_fold = cached
        return cached[1], cached[2]



class PointBatchNorm1D(FoldedBatchNorm, nn.BatchNorm1d):
    """BatchNorm1d over the last dim of an arbitrarily shaped [..., C] tensor (ref components/mlp.py:3-7)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("garmentnets_b200 implements the inference hot path only (BatchNorm in eval mode)")
        # Stand-alone BN never occurs on the hot path (it is fused into the Linear block); kept for API completeness
        # as the same fused kernel with an identity weight.
        scale, shift = self.folded_affine()
        flat = x.reshape(-1, x.shape[-1])
        eye = torch.eye(flat.shape[1], dtype=torch.float32, device=flat.device)
        return ops.linear(flat, eye, None, False, scale, shift).view(x.shape)


class _Block(nn.Sequential):
    """Linear -> ReLU -> (PointBatchNorm1D): one fused kernel launch."""

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("garmentnets_b200 implements the inference hot path only (eval mode)")
        lin = self[0]
        lead = x.shape[:-1]
        flat = x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x
        if _Block.calibrating and len(self) > 2:
            y = self._calibrate(flat, out, rows_dev)
            return y if x.dim() == 2 else y.view(*lead, y.shape[-1])
        bn = self[2] if len(self) > 2 else None
        if ops.USE_LINEAR_TC and flat.shape[0] >= ops.LINEAR_TC_MIN_ROWS:
            w = ops.packed_linear_for(self, "block", lin.weight, lin.bias, bn)
            y = ops.linear_tc(flat, w, True, out=out, rows_dev=rows_dev)
        else:
            scale, shift = bn.folded_affine() if bn is not None else (None, None)
            y = ops.linear(flat, lin.weight, lin.bias, True, scale, shift, out=out, rows_dev=rows_dev)
        return y if x.dim() == 2 else y.reshape(*lead, y.shape[-1])

    # Synthetic-weight support (garmentnets_b200.synthetic.calibrate_bn_): set this block's BatchNorm running
    # statistics to the statistics of the activations it actually sees, the way training would have.  Not a compute
    # path: it only edits parameters, once, before any measurement.
    calibrating = False

    @torch.no_grad()
    def _calibrate(self, flat, out, rows_dev):
        lin, bn = self[0], self[2]
        y = ops.linear(flat, lin.weight, lin.bias, True, None, None, rows_dev=rows_dev)
        n = y.shape[0] if rows_dev is None else int(rows_dev.item())
        bn.running_mean.copy_(y[:n].mean(0))
        bn.running_var.copy_(y[:n].var(0, unbiased=False).clamp_min(1e-2))
        scale, shift = bn.folded_affine()
        res = y * scale + shift
        if out is not None:
            out.copy_(res)
            return out
        return res


class FusedMLP(nn.Sequential):
    def forward(self, x: torch.Tensor, rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        for block in self:
            x = block(x, rows_dev=rows_dev)
        return x

    @property
    def channels(self):
        return [self[0][0].in_features] + [b[0].out_features for b in self]


def MLP(channels, batch_norm=True):
    blocks = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        parts = [nn.Linear(cin, cout), nn.ReLU()]
        if batch_norm:
            parts.append(PointBatchNorm1D(cout))
        blocks.append(_Block(*parts))
    return FusedMLP(*blocks)
