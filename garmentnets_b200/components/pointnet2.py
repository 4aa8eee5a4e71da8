"""PointNet++ modules -- mirror of the reference's ``components/pointnet2.py`` (``SAModule`` :11-33,
``GlobalSAModule`` :36-52, ``FPModule`` :61-76, module-level ``MLP`` :55-59) on the sm_100a kernels.

Same constructor signatures, attribute paths (``.conv.local_nn`` etc., hence the same checkpoint keys) and
``forward`` contracts.  The third-party ops the reference calls (torch_cluster ``fps`` / ``radius`` / ``knn``, PyG
``PointConv`` / ``knn_interpolate`` / ``global_max_pool``) are replaced by ``gnb_*`` entry points; their functional
forms are also exported here under the PyG names for the import shims.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops
from .mlp import FoldedBatchNorm, FusedMLP, _Block


# ------------------------------------------------------------------------------------------------ functional
class CloudIndex:
    """CSR description of a PyG-style flat batch: device ``ptr`` plus its host copy (sizes the launches)."""

    def __init__(self, ptr: torch.Tensor, ptr_host: np.ndarray):
        self.ptr = ptr
        self.ptr_host = np.asarray(ptr_host, dtype=np.int64)

    @property
    def num_graphs(self) -> int:
        return len(self.ptr_host) - 1

    @property
    def max_n(self) -> int:
        return int(np.diff(self.ptr_host).max()) if self.num_graphs else 0

    @property
    def total(self) -> int:
        return int(self.ptr_host[-1])

    @staticmethod
    def from_batch(batch: torch.Tensor) -> "CloudIndex":
        # one device->host read, like the reference's `int(batch.max()) + 1` inside torch_cluster
        ptr = ops.batch_to_ptr(batch)
        return CloudIndex(ptr, ptr.cpu().numpy())

    # Device copies of small host offset arrays, keyed by content.  A pageable host->device copy synchronises the stream
    # before it starts, so creating the same CSR offsets again on every forward (three times per step: SA1, SA2 and the
    # global level) drained the launch queue each time; with the cache the steady state issues no such copy.
    _device_cache: dict = {}

    @staticmethod
    def _to_device(host: np.ndarray, device) -> torch.Tensor:
        host = np.ascontiguousarray(host, dtype=np.int64)
        key = (host.tobytes(), str(device))
        hit = CloudIndex._device_cache.get(key)
        if hit is None:
            if len(CloudIndex._device_cache) >= 512:
                CloudIndex._device_cache.clear()
            hit = torch.from_numpy(host.copy()).to(device)
            CloudIndex._device_cache[key] = hit
        return hit

    @staticmethod
    def uniform(B: int, n: int, device) -> "CloudIndex":
        host = np.arange(B + 1, dtype=np.int64) * n
        return CloudIndex(CloudIndex._to_device(host, device), host)

    def subsample(self, ratio: float) -> "CloudIndex":
        counts = ops.fps_counts(self.ptr_host, ratio)
        host = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        return CloudIndex(CloudIndex._to_device(host, self.ptr.device), host)

    def batch_vector(self) -> torch.Tensor:
        counts = torch.from_numpy(np.diff(self.ptr_host)).to(self.ptr.device)
        return torch.repeat_interleave(torch.arange(self.num_graphs, device=self.ptr.device), counts)


def fps(pos, batch=None, ratio=0.5, random_start=True, *, index: Optional[CloudIndex] = None,
        start: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch_geometric.nn.fps (ref components/pointnet2.py:26).  ``start`` injects the per-cloud start index."""
    if index is None:
        index = CloudIndex.from_batch(batch if batch is not None else pos.new_zeros(pos.shape[0], dtype=torch.int64))
    sub = index.subsample(ratio)
    if start is None and random_start:
        sizes = np.maximum(np.diff(index.ptr_host), 1)
        start = torch.from_numpy((np.random.randint(0, 2 ** 31, size=len(sizes)) % sizes).astype(np.int64)).to(pos.device)
    return ops.fps(pos, index.ptr, sub.ptr, index.max_n, sub.total, start)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, *, index_x=None, index_y=None):
    """torch_geometric.nn.radius (ref components/pointnet2.py:28-29): returns stacked [row(y idx), col(x idx)]."""
    index_x = index_x or CloudIndex.from_batch(batch_x if batch_x is not None else x.new_zeros(len(x), dtype=torch.int64))
    index_y = index_y or CloudIndex.from_batch(batch_y if batch_y is not None else y.new_zeros(len(y), dtype=torch.int64))
    nbr, cnt = ops.ball_query(x, y, index_x.ptr, index_y.ptr, r, max_num_neighbors)
    row, col = ops.radius_pairs(nbr, cnt)
    return torch.stack([row, col], dim=0)


def global_max_pool(x, batch, size=None, *, index: Optional[CloudIndex] = None):
    """torch_geometric.nn.global_max_pool (ref components/pointnet2.py:49); ``batch`` is sorted."""
    index = index or CloudIndex.from_batch(batch)
    return ops.segment_max(x, index.ptr)


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, num_workers=1, *, index_x=None, index_y=None,
                    out: Optional[torch.Tensor] = None):
    """torch_geometric.nn.knn_interpolate (ref components/pointnet2.py:72)."""
    index_x = index_x or CloudIndex.from_batch(batch_x if batch_x is not None else pos_x.new_zeros(len(pos_x), dtype=torch.int64))
    index_y = index_y or CloudIndex.from_batch(batch_y if batch_y is not None else pos_y.new_zeros(len(pos_y), dtype=torch.int64))
    idx, d2 = ops.knn(pos_x, pos_y, index_x.ptr, index_y.ptr, k)
    if out is None:
        out = torch.empty((pos_y.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    ops.knn_interpolate_into(x, idx, d2, out)
    return out


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class PointConv(nn.Module):
    """torch_geometric.nn.PointConv (1.7.2) restricted to what the reference uses: ``local_nn``, max aggregation,
    ``add_self_loops=True``, no ``global_nn``."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True, **kwargs):
        super().__init__()
        if global_nn is not None:
            raise NotImplementedError("PointConv(global_nn=...) is not used on the GarmentNets hot path")
        self.local_nn = local_nn
        self.global_nn = None
        self.add_self_loops = add_self_loops

    def forward_grouped(self, x, pos_x, pos_y, nbr, cnt) -> torch.Tensor:
        """Edge set = ball query U self loop (see include/garmentnets_b200.h, N4); message MLP; segment max."""
        if not self.add_self_loops:
            raise NotImplementedError("PointConv(add_self_loops=False) is not used on the GarmentNets hot path")
        M, K = nbr.shape
        eoffs = ops.pointconv_edges(nbr, cnt)
        cin = 0 if x is None else x.shape[1]
        fused = self._fused_layers(cin)
        if fused is not None and x is not None and x.dtype == torch.float32 and x.stride(1) == 1:
            # one kernel: gather + the three Linear -> ReLU -> BatchNorm blocks + max aggregation, activations stay on chip
            return ops.pointconv_mlp_max(x, _f32c(pos_x), _f32c(pos_y), nbr, cnt, eoffs, ops.pointconv_mlp_pack(self, fused))
        rows = M * (K + 1)  # worst case; kernels stop at the device-side edge total eoffs[M]
        edge = torch.empty((rows, cin + 3), dtype=torch.float32, device=pos_x.device)
        ops.pointconv_gather(x, pos_x, pos_y, nbr, cnt, eoffs, edge)
        total = eoffs[M:]
        h = edge
        blocks = list(self.local_nn)
        last = blocks[-1]
        fuse = (self.fuse_aggregation and ops.USE_LINEAR_TC and rows >= ops.LINEAR_TC_MIN_ROWS and not _Block.calibrating
                and not last.training and isinstance(last, _Block))
        for block in (blocks[:-1] if fuse else blocks):
            h = block(h, rows_dev=total)
        if not fuse:
            return ops.segment_max(h, eoffs)
        # max aggregation fused into the last layer's epilogue: its [E, C] output never reaches HBM
        lin = last[0]
        w = ops.packed_linear_for(last, "block", lin.weight, lin.bias, last[2] if len(last) > 2 else None)
        return ops.linear_tc_segmax(h, w, ops.segment_ids(eoffs, rows), M, relu=True, rows_dev=total)

    def _fused_layers(self, cin: int):
        """(weight, bias, bn_scale, bn_shift) of the three blocks when gnb_pointconv_mlp_max can run this MLP, else None."""
        if not (ops.USE_SA_MLP and self.add_self_loops) or _Block.calibrating:
            return None
        blocks = list(self.local_nn) if self.local_nn is not None else []
        if len(blocks) != 3 or not all(isinstance(b, _Block) and len(b) > 2 and not b.training for b in blocks):
            return None
        ch = [blocks[0][0].in_features] + [b[0].out_features for b in blocks]
        if ch[0] != cin + 3 or not ops.pointconv_mlp_supported(cin, ch[1], ch[2], ch[3]):
            return None
        return [(b[0].weight, b[0].bias, *b[2].folded_affine()) for b in blocks]

    # Off by default: measured on B200 (batch 32) the fused epilogue -- column-wise run maxima out of a shared-memory box --
    # costs more than it saves (SA1 last layer 1140 us fused vs 525 + 217 us for gnb_linear_tc + gnb_segment_max, both of
    # which already run at 3.5-4 TB/s and at the HBM roofline respectively).  Kept (and tested) as the starting point for
    # a register-level segmented reduction.
    fuse_aggregation = False

    def forward(self, x, pos, edge_index):
        """The PyG call shape the reference's own ``SAModule`` uses (ref components/pointnet2.py:30-31):
        ``x`` [Nx,C] or None (or a pair), ``pos`` = (pos_x [Nx,3], pos_y [M,3]) or one tensor, ``edge_index`` i64[2,E] =
        [source j into x / pos_x, target i into pos_y].  PyG 1.7.2 semantics: with ``add_self_loops`` the edges whose two
        INDICES coincide are dropped and (i, i) is appended for i < M -- also in the bipartite case, where source i is a
        different point than centre i; message = local_nn([x_j, pos_j - pos_i]); max aggregation over the targets.
        Arbitrary edge order is accepted (edges are grouped by target with a stable device sort); the edge MLP and the
        segmented max are the same kernels ``forward_grouped`` uses."""
        if isinstance(x, (tuple, list)):
            x = x[0]
        pos_x, pos_y = (pos, pos) if isinstance(pos, torch.Tensor) else pos
        M = pos_y.shape[0]
        src, dst = edge_index[0], edge_index[1]
        if self.add_self_loops:
            keep = src != dst
            loop = torch.arange(M, dtype=src.dtype, device=src.device)
            src, dst = torch.cat([src[keep], loop]), torch.cat([dst[keep], loop])
        dst, perm = torch.sort(dst, stable=True)
        src = src[perm]
        eoffs = ops.batch_to_ptr(dst, M)
        msg = pos_x[src] - pos_y[dst]
        edge = msg if x is None else torch.cat([x[src], msg], dim=1)
        h = edge.contiguous()
        for block in self.local_nn:
            h = block(h)
        out = ops.segment_max(h, eoffs)
        if not self.add_self_loops:   # PyG's max aggregation leaves 0 for targets without any edge
            empty = (eoffs[1:] == eoffs[:-1])
            out[empty] = 0.0
        return out


# ------------------------------------------------------------------------------------------------ modules
class SAModule(nn.Module):
    """Local set abstraction: FPS -> ball query (<=64) -> PointConv (ref components/pointnet2.py:11-33)."""

    def __init__(self, ratio, r, nn):
        super().__init__()
        self.ratio = ratio
        self.r = r
        self.conv = PointConv(nn)
        self.random_start = True  # the reference's fps default; set False (or pass fps_start) for determinism

    def forward(self, x, pos, batch, *, index: Optional[CloudIndex] = None, fps_start: Optional[torch.Tensor] = None,
                return_index: bool = False):
        index = index or CloudIndex.from_batch(batch)
        sub = index.subsample(self.ratio)
        idx = fps(pos, None, self.ratio, self.random_start, index=index, start=fps_start)
        pos_y = pos[idx]
        nbr, cnt = ops.ball_query(pos, pos_y, index.ptr, sub.ptr, self.r, 64)
        out = self.conv.forward_grouped(x, pos, pos_y, nbr, cnt)
        batch_y = batch[idx] if batch is not None else sub.batch_vector()
        if return_index:
            return out, pos_y, batch_y, sub, {"idx": idx, "nbr": nbr, "cnt": cnt}
        return out, pos_y, batch_y


class GlobalSAModule(nn.Module):
    """Global set abstraction (ref components/pointnet2.py:36-52)."""

    def __init__(self, nn):
        super().__init__()
        self.nn = nn

    def forward(self, x, pos, batch, *, index: Optional[CloudIndex] = None):
        index = index or CloudIndex.from_batch(batch)
        h = torch.cat([x, pos], dim=1)
        blocks = list(self.nn) if isinstance(self.nn, nn.Sequential) else None
        last = blocks[-1] if blocks else None
        rows = h.shape[0]
        fuse = (blocks is not None and ops.USE_LINEAR_TC and rows >= ops.LINEAR_TC_MIN_ROWS and not _Block.calibrating
                and isinstance(last, _Block) and not last.training)
        if fuse:
            # global max pooling fused into the last layer's epilogue (as in PointConv): its [rows, C] output never reaches HBM
            # and the pooling does not run as 32 CTAs of its own
            for block in blocks[:-1]:
                h = block(h)
            lin = last[0]
            w = ops.packed_linear_for(last, "block", lin.weight, lin.bias, last[2] if len(last) > 2 else None)
            out = ops.linear_tc_segmax(h, w, ops.segment_ids(index.ptr, rows), index.num_graphs, relu=True)
        else:
            out = ops.segment_max(self.nn(h), index.ptr)
        B = out.shape[0]
        return out, pos.new_zeros((B, 3)), torch.arange(B, device=pos.device)


def MLP(channels, batch_norm=True):
    """The unused module-level helper of the reference (components/pointnet2.py:55-59): Linear, ReLU, BatchNorm1d."""
    blocks = [_Block(nn.Linear(channels[i - 1], channels[i]), nn.ReLU(), _PlainBN(channels[i]))
              for i in range(1, len(channels))]
    return FusedMLP(*blocks)


class _PlainBN(FoldedBatchNorm, nn.BatchNorm1d):
    pass


class FPModule(nn.Module):
    """Feature propagation: kNN inverse-distance interpolation + skip concat + MLP (ref components/pointnet2.py:61-76)."""

    def __init__(self, k, nn):
        super().__init__()
        self.k = k
        self.nn = nn

    def forward(self, x, pos, batch, x_skip, pos_skip, batch_skip, *, index: Optional[CloudIndex] = None,
                index_skip: Optional[CloudIndex] = None):
        index = index or CloudIndex.from_batch(batch)
        index_skip = index_skip or CloudIndex.from_batch(batch_skip)
        c = x.shape[1]
        cs = 0 if x_skip is None else x_skip.shape[1]
        # row stride padded to a multiple of 4 floats: the interpolation kernel then takes its 16-byte path (FP1 concatenates
        # 128 + 3 channels) and the Linear block reads rows at any stride
        buf = ops.padded_rows(pos_skip.shape[0], c + cs, x.device)
        knn_interpolate(x, pos, pos_skip, k=self.k, index_x=index, index_y=index_skip, out=buf)
        if x_skip is not None:
            buf[:, c:] = x_skip
        return self.nn(buf), pos_skip, batch_skip
