"""Voxel-grid helpers -- mirror of the reference's ``components/gridding.py`` (``batch_to_volume`` :8-42,
``nocs_grid_sample`` :45-98, ``VirtualGrid`` :101-256, ``ceil_div`` :259, ``ArraySlicer`` :262-298).

``VirtualGrid`` / ``ArraySlicer`` are host-side index arithmetic (a handful of tiny tensor expressions, written to
round exactly like the reference: fp32 ``(shape-1)/(uc-lc)`` scales, truncation toward zero, per-axis clamp).  The
two functions that move real data dispatch to the sm_100a kernels: ``batch_to_volume`` -> ``gnb_scatter_reduce``,
``nocs_grid_sample`` -> ``gnb_trilinear_sample`` (flipped zyx convention).
"""
from __future__ import annotations

import itertools
from typing import Tuple

import numpy as np
import torch

from .. import ops


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    """``torch_scatter.scatter`` for the call shape the reference uses: ``src`` [C,N], ``index`` [N], ``dim=-1``
    (ref networks/conv_implicit_wnf.py:92-94, components/gridding.py:32-35).  Empty slots are 0.  The result is a
    logical [C, dim_size] tensor stored channels-last, so the caller's reshape/permute yields an NDHWC volume."""
    if out is not None:
        raise NotImplementedError("scatter(out=...) is not used on the GarmentNets hot path")
    if src.dim() != 2 or index.dim() != 1 or dim not in (-1, 1):
        raise NotImplementedError("scatter: only src [C,N], index [N], dim=-1 is implemented (the reference's call)")
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    return ops.scatter_reduce(src, index, int(dim_size), reduce, channels_last=True)


def batch_to_volume(batch, volume_size, reduce='mean'):
    """Scatter per-point features into a [B,C,G,G,G] volume (older gridding API; imported by the reference pipeline
    at networks/conv_implicit_wnf.py:16 but never called).  Voxel = clamp(trunc(pos * G), 0, G-1)."""
    pts, feats, bidx = batch.pos, batch.x, batch.batch
    B = int(batch.num_graphs)
    G = int(volume_size)
    with torch.no_grad():
        ijk = torch.clamp((pts * G).to(torch.int64), 0, G - 1)
        flat = ((bidx * G + ijk[:, 0]) * G + ijk[:, 1]) * G + ijk[:, 2]
    vol = scatter(feats.to(torch.float32).t(), flat, dim=-1, dim_size=B * G ** 3, reduce=reduce)
    C = feats.shape[1]
    return vol.reshape(C, B, G, G, G).permute(1, 0, 2, 3, 4)


def nocs_grid_sample(feature_volume: torch.Tensor, query_points: torch.Tensor, mode: str = 'bilinear',
                     padding_mode: str = 'border', align_corners: bool = True) -> torch.Tensor:
    """Trilinear lookup with xyz query points against a (D,H,W)=(x,y,z)-indexed volume (the reference flips the query
    to zyx before ``grid_sample``, components/gridding.py:69-70).
    feature_volume (N,C,D,H,W) | (N,D,H,W) | (D,H,W); query_points (N,M,3) | (M,3) -> (N,M,C) | (M,C)."""
    if (mode, padding_mode, align_corners) != ('bilinear', 'border', True):
        raise NotImplementedError("nocs_grid_sample: only bilinear / border / align_corners=True is implemented")
    if query_points.dim() not in (2, 3):
        raise RuntimeError("Invalid query_points shape {}".format(str(query_points.shape)))
    q = query_points if query_points.dim() == 3 else query_points.unsqueeze(0)
    fv = feature_volume
    if fv.dim() == 3:
        fv = fv[None, None]
    elif fv.dim() == 4:
        fv = fv[:, None]
    vol = ops.to_channels_last(fv)
    out = ops.trilinear_sample(vol, q.contiguous(), flip=True).view(q.shape[0], q.shape[1], fv.shape[1])
    return out if query_points.dim() == 3 else out[0]


class VirtualGrid:
    """Axis-aligned lattice of ``grid_shape`` nodes spanning [lower_corner, upper_corner] (nodes ON the corners)."""

    def __init__(self, lower_corner=(0, 0, 0), upper_corner=(1, 1, 1), grid_shape=(32, 32, 32), batch_size=8,
                 device=torch.device('cpu'), int_dtype=torch.int64, float_dtype=torch.float32):
        self.lower_corner = tuple(lower_corner)
        self.upper_corner = tuple(upper_corner)
        self.grid_shape = tuple(grid_shape)
        self.batch_size = int(batch_size)
        self.device = device
        self.int_dtype = int_dtype
        self.float_dtype = float_dtype

    # -- helpers -------------------------------------------------------------------------------------------
    def _corners(self, device):
        kw = dict(dtype=self.float_dtype, device=device)
        lc = torch.tensor(self.lower_corner, **kw)
        uc = torch.tensor(self.upper_corner, **kw)
        last = torch.tensor(self.grid_shape, **kw) - 1  # index of the last node per axis, as float
        return lc, uc, last

    def _dims(self, include_batch: bool) -> Tuple[int, ...]:
        return ((self.batch_size,) if include_batch else ()) + self.grid_shape

    @staticmethod
    def _row_major_strides(dims) -> Tuple[int, ...]:
        strides = [1]
        for d in reversed(dims[1:]):
            strides.append(strides[-1] * int(d))
        return tuple(reversed(strides))

    # -- API ---------------------------------------------------------------------------------------------------
    @property
    def num_grids(self):
        return int(np.prod(self._dims(True)))

    def get_grid_idxs(self, include_batch=True):
        axes = [torch.arange(n, device=self.device, dtype=self.int_dtype) for n in self._dims(include_batch)]
        return torch.stack(torch.meshgrid(*axes, indexing='ij'), dim=-1)

    def get_grid_points(self, include_batch=True):
        lc, uc, last = self._corners(self.device)
        idxs = self.get_grid_idxs(include_batch=include_batch)
        if include_batch:
            idxs = idxs[..., 1:]
        return idxs.to(self.float_dtype) * ((uc - lc) / last) + (-lc)

    def get_points_grid_idxs(self, points, batch_idx=None):
        lc, uc, last = self._corners(self.device)
        cell = ((points + (-lc)) * (last / (uc - lc))).to(dtype=self.int_dtype)  # truncation toward zero
        hi = torch.tensor(self.grid_shape, dtype=self.int_dtype, device=cell.device) - 1
        cell = torch.minimum(torch.clamp(cell, min=0), hi)
        if batch_idx is None:
            return cell
        b = batch_idx.view(*points.shape[:-1], 1).to(dtype=cell.dtype)
        return torch.cat([b, cell], dim=-1)

    def flatten_idxs(self, idxs, keepdim=False):
        width = idxs.shape[-1]
        if width not in (3, 4):
            raise RuntimeError("Invalid shape {}".format(str(idxs.shape)))
        strides = self._row_major_strides(self._dims(width == 4))
        w = torch.tensor(strides, dtype=idxs.dtype, device=idxs.device)
        return (idxs * w).sum(dim=-1, keepdim=keepdim, dtype=idxs.dtype)

    def unflatten_idxs(self, flat_idxs, include_batch=True):
        strides = self._row_major_strides(self._dims(include_batch))
        if flat_idxs.shape[-1:] == (1,):
            flat_idxs = flat_idxs[..., 0]
        parts, rest = [], flat_idxs
        for s in strides:
            parts.append(torch.div(rest, s, rounding_mode='floor'))
            rest = rest % s
        return torch.stack(parts, dim=-1)

    def idxs_to_points(self, idxs):
        if idxs.shape[-1] not in (3, 4):
            raise RuntimeError("Invalid shape {}".format(tuple(idxs.shape)))
        lc, uc, last = self._corners(idxs.device)
        cell = idxs[..., 1:] if idxs.shape[-1] == 4 else idxs
        return cell * ((uc - lc) / last) + lc


def ceil_div(a, b):
    return -(-a // b)


class ArraySlicer:
    """Iterate an array in ``chunks``-sized blocks over its leading axes, last chunked axis fastest."""

    def __init__(self, shape: tuple, chunks: tuple):
        assert len(chunks) <= len(shape)
        self.relevent_shape = tuple(shape[:len(chunks)])
        self.chunks = tuple(chunks)
        self.chunk_size = tuple(ceil_div(n, c) for n, c in zip(self.relevent_shape, self.chunks))

    def __len__(self):
        return int(np.prod(self.chunk_size))

    def __getitem__(self, idx):
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        block = np.unravel_index(idx, self.chunk_size)
        return [slice(int(b) * c, min(n, (int(b) + 1) * c))
                for b, c, n in zip(block, self.chunks, self.relevent_shape)]

    def __iter__(self):
        for block in itertools.product(*[range(n) for n in self.chunk_size]):
            yield [slice(b * c, min(n, (b + 1) * c)) for b, c, n in zip(block, self.chunks, self.relevent_shape)]
