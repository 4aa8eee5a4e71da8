"""Tensor-level wrappers around the C-ABI (``include/garmentnets_b200.h``).

PyTorch is plumbing here: it owns device memory and the current stream; every computation below is a call into
``libgarmentnets_b200.so``.  Inputs must live on a CUDA device -- there is no CPU path.
"""
from __future__ import annotations

import math
import threading
from typing import Dict, Optional, Tuple

import torch

from . import _lib

REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2, "min": 3}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype: torch.dtype, name: str, contiguous: bool = True) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.GarmentNetsB200Error(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.GarmentNetsB200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        t = t.contiguous()
    return t


def _rows(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """2-D fp32 tensor whose last dim is contiguous; returns (tensor, row stride)."""
    t = _req(t, torch.float32, name, contiguous=False)
    if t.dim() != 2:
        raise _lib.GarmentNetsB200Error(f"{name}: expected a 2-D tensor")
    if t.shape[1] > 1 and t.stride(1) != 1 or t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t, int(ld)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------- point ops
def batch_to_ptr(batch: torch.Tensor, num_graphs: Optional[int] = None) -> torch.Tensor:
    """Sorted PyG batch vector -> CSR offsets i64[B+1] (plumbing; one bincount + cumsum)."""
    if num_graphs is None:
        num_graphs = int(batch.max().item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=num_graphs)
    ptr = torch.zeros(num_graphs + 1, dtype=torch.int64, device=batch.device)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


def fps_counts(ptr_host, ratio: float):
    """m_b = ceil(fp32(n_b) * fp32(ratio)) (torch_cluster 1.5.9 `fps`: deg.float() * ratio, ceil)."""
    import numpy as np
    n = np.diff(np.asarray(ptr_host, dtype=np.int64)).astype(np.float32)
    return np.ceil(n * np.float32(ratio)).astype(np.int64)


def fps(pos: torch.Tensor, ptr: torch.Tensor, out_ptr: torch.Tensor, max_n: int, total_m: int,
        start: Optional[torch.Tensor] = None) -> torch.Tensor:
    pos = _req(pos, torch.float32, "pos")
    ptr = _req(ptr, torch.int64, "ptr")
    out_ptr = _req(out_ptr, torch.int64, "out_ptr")
    if start is not None:
        start = _req(start, torch.int64, "start")
    out = torch.empty(total_m, dtype=torch.int64, device=pos.device)
    _lib.call("gnb_fps", pos.data_ptr(), ptr.data_ptr(), ptr.numel() - 1, _ptr(start), out_ptr.data_ptr(),
              out.data_ptr(), int(max_n), _stream())
    return out


def ball_query(x: torch.Tensor, y: torch.Tensor, ptr_x: torch.Tensor, ptr_y: torch.Tensor, r: float,
               K: int = 64) -> Tuple[torch.Tensor, torch.Tensor]:
    x = _req(x, torch.float32, "x")
    y = _req(y, torch.float32, "y")
    M = y.shape[0]
    nbr = torch.empty((M, K), dtype=torch.int64, device=x.device)
    cnt = torch.empty((M,), dtype=torch.int32, device=x.device)
    _lib.call("gnb_ball_query", x.data_ptr(), y.data_ptr(), ptr_x.data_ptr(), ptr_y.data_ptr(), ptr_x.numel() - 1, M,
              float(r), int(K), nbr.data_ptr(), cnt.data_ptr(), _stream())
    return nbr, cnt


def exclusive_scan(cnt: torch.Tensor) -> torch.Tensor:
    cnt = _req(cnt, torch.int32, "cnt")
    out = torch.empty(cnt.numel() + 1, dtype=torch.int64, device=cnt.device)
    _lib.call("gnb_exclusive_scan_i32", cnt.data_ptr(), cnt.numel(), out.data_ptr(), _stream())
    return out


def radius_pairs(nbr: torch.Tensor, cnt: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(row, col) pair list in torch_cluster.radius order; sizes the output with one host read of the total."""
    offs = exclusive_scan(cnt)
    total = int(offs[-1].item())
    row = torch.empty(total, dtype=torch.int64, device=nbr.device)
    col = torch.empty(total, dtype=torch.int64, device=nbr.device)
    _lib.call("gnb_radius_pairs", nbr.data_ptr(), cnt.data_ptr(), offs.data_ptr(), nbr.shape[0], nbr.shape[1],
              row.data_ptr(), col.data_ptr(), _stream())
    return row, col


def knn(x: torch.Tensor, y: torch.Tensor, ptr_x: torch.Tensor, ptr_y: torch.Tensor, k: int):
    x = _req(x, torch.float32, "x")
    y = _req(y, torch.float32, "y")
    Ny = y.shape[0]
    idx = torch.empty((Ny, k), dtype=torch.int64, device=x.device)
    d2 = torch.empty((Ny, k), dtype=torch.float32, device=x.device)
    _lib.call("gnb_knn", x.data_ptr(), y.data_ptr(), ptr_x.data_ptr(), ptr_y.data_ptr(), ptr_x.numel() - 1, Ny, int(k),
              idx.data_ptr(), d2.data_ptr(), _stream())
    return idx, d2


def knn_interpolate_into(feat: torch.Tensor, idx: torch.Tensor, d2: torch.Tensor, out: torch.Tensor) -> None:
    """out[:, :C] = inverse-distance interpolation; ``out`` may be wider (the concat buffer)."""
    feat, ldf = _rows(feat, "feat")
    assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1
    _lib.call("gnb_knn_interpolate", feat.data_ptr(), ldf, idx.data_ptr(), d2.data_ptr(), idx.shape[0], idx.shape[1],
              feat.shape[1], out.data_ptr(), out.stride(0), _stream())


def pointconv_edges(nbr: torch.Tensor, cnt: torch.Tensor) -> torch.Tensor:
    """CSR offsets i64[M+1] of the PointConv edge set (ball query U self loop)."""
    ecnt = torch.empty_like(cnt)
    _lib.call("gnb_pointconv_edge_count", nbr.data_ptr(), cnt.data_ptr(), nbr.shape[0], nbr.shape[1], ecnt.data_ptr(),
              _stream())
    return exclusive_scan(ecnt)


def pointconv_gather(x_feat: Optional[torch.Tensor], pos_x: torch.Tensor, pos_y: torch.Tensor, nbr: torch.Tensor,
                     cnt: torch.Tensor, eoffs: torch.Tensor, edge: torch.Tensor) -> None:
    if x_feat is not None:
        x_feat, ldx = _rows(x_feat, "x_feat")
        cin = x_feat.shape[1]
    else:
        ldx, cin = 0, 0
    _lib.call("gnb_pointconv_gather", _ptr(x_feat), ldx, cin, pos_x.data_ptr(), pos_y.data_ptr(), nbr.data_ptr(),
              cnt.data_ptr(), eoffs.data_ptr(), nbr.shape[0], nbr.shape[1], edge.data_ptr(), edge.stride(0), _stream())


def segment_max(rows: torch.Tensor, offs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows, ldr = _rows(rows, "rows")
    nseg = offs.numel() - 1
    if out is None:
        out = torch.empty((nseg, rows.shape[1]), dtype=torch.float32, device=rows.device)
    _lib.call("gnb_segment_max", rows.data_ptr(), ldr, offs.data_ptr(), nseg, rows.shape[1], out.data_ptr(),
              out.stride(0), _stream())
    return out


# ---------------------------------------------------------------------------------------------- dense ops
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
           bn_scale: Optional[torch.Tensor] = None, bn_shift: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    x, ldx = _rows(x, "x")
    weight = _req(weight, torch.float32, "weight")
    R, K = x.shape
    N = weight.shape[0]
    assert weight.shape[1] == K, f"weight {tuple(weight.shape)} vs input K={K}"
    if out is None:
        out = torch.empty((R, N), dtype=torch.float32, device=x.device)
    assert out.stride(1) == 1 or N == 1
    _lib.call("gnb_linear", x.data_ptr(), R, K, ldx, weight.data_ptr(), _ptr(bias), N, int(relu), _ptr(bn_scale),
              _ptr(bn_shift), out.data_ptr(), out.stride(0), _ptr(rows_dev), _stream())
    return out


class PackedLinear:
    """Weight of one Linear(+bias, +folded BatchNorm affine) packed for ``gnb_linear_tc`` (fp16 hi/lo shared-memory
    images + padded per-column epilogue parameters)."""
    __slots__ = ("packed", "cparams", "scale_log2", "N", "K")

    def __init__(self, packed, cparams, scale_log2, N, K):
        self.packed, self.cparams, self.scale_log2, self.N, self.K = packed, cparams, scale_log2, N, K


def pack_linear_tc(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn_scale: Optional[torch.Tensor] = None,
                   bn_shift: Optional[torch.Tensor] = None) -> PackedLinear:
    weight = _req(weight, torch.float32, "weight")
    N, K = weight.shape
    lib = _lib.load()
    s = _pow2_scale(weight)
    packed = torch.empty(int(lib.gnb_linear_tc_packed_bytes(N, K)), dtype=torch.uint8, device=weight.device)
    cparams = torch.empty(3 * int(lib.gnb_linear_tc_padded_cols(N, K)), dtype=torch.float32, device=weight.device)
    _lib.call("gnb_linear_tc_pack", weight.data_ptr(), N, K, _ptr(bias), _ptr(bn_scale), _ptr(bn_shift), s,
              packed.data_ptr(), cparams.data_ptr(), _stream())
    return PackedLinear(packed, cparams, s, N, K)


def _version_key(*tensors):
    return tuple(None if t is None else (t.data_ptr(), t._version) for t in tensors)


def packed_linear_for(owner, slot: str, weight, bias=None, bn=None) -> PackedLinear:
    """``pack_linear_tc`` cached on ``owner`` (any object) under ``slot`` until a parameter / running statistic of the
    Linear or of its BatchNorm (an object with ``folded_affine()``) changes."""
    bn_t = () if bn is None else (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = _version_key(weight, bias, *bn_t)
    cache = owner.__dict__.setdefault("_gnb_packed_linear", {})
    hit = cache.get(slot)
    if hit is None or hit[0] != key:
        sc, sh = bn.folded_affine() if bn is not None else (None, None)
        hit = (key, pack_linear_tc(weight, bias, sc, sh))
        cache[slot] = hit
    return hit[1]


USE_LINEAR_TC = True       # False: every Linear block runs on the fp32 FFMA kernel (gnb_linear)
LINEAR_TC_MIN_ROWS = 1024   # below this the fp32 kernel is used (a 128-row tensor-core tile per SM would idle most SMs)


def padded_rows(R: int, N: int, device) -> torch.Tensor:
    """[R, N] fp32 view whose row stride is a multiple of 4 floats (16-byte aligned rows for vector stores)."""
    ld = (N + 3) // 4 * 4
    buf = torch.empty((R, ld), dtype=torch.float32, device=device)
    return buf if ld == N else buf[:, :N]


def linear_tc(x: torch.Tensor, w: PackedLinear, relu: bool = False, out: Optional[torch.Tensor] = None,
              rows_dev: Optional[torch.Tensor] = None, flag_range: bool = False) -> torch.Tensor:
    """Tensor-core Linear block (``gnb_linear_tc``): x [R,K] (any row stride) -> [R,N].  ``flag_range``: the epilogue also
    raises the fp16 range flag for outputs outside +-65504 (``gnb_linear_tc_flagged``)."""
    x, ldx = _rows(x, "x")
    R, K = x.shape
    assert K == w.K, f"packed weight K={w.K} vs input K={K}"
    if out is None:
        out = padded_rows(R, w.N, x.device)
    assert out.stride(1) == 1 or w.N == 1
    _lib.call("gnb_linear_tc_flagged" if flag_range else "gnb_linear_tc", x.data_ptr(), R, K, ldx, w.packed.data_ptr(),
              w.cparams.data_ptr(), w.scale_log2, w.N, int(relu), out.data_ptr(), out.stride(0), _ptr(rows_dev), _stream())
    return out


def segment_ids(offs: torch.Tensor, rows: int) -> torch.Tensor:
    """CSR offsets i64[nseg+1] -> i32[rows] segment id of every row (rows beyond offs[-1] are left unset)."""
    offs = _req(offs, torch.int64, "offs")
    seg = torch.empty((rows,), dtype=torch.int32, device=offs.device)
    _lib.call("gnb_segment_ids", offs.data_ptr(), offs.numel() - 1, seg.data_ptr(), _stream())
    return seg


def linear_tc_segmax(x: torch.Tensor, w: PackedLinear, seg: torch.Tensor, nseg: int, relu: bool = True,
                     rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``segment_max(linear_tc(x), offs)`` in one kernel (``gnb_linear_tc_segmax``): x [R,K] with contiguous segments,
    seg i32[R] -> [nseg, N] fp32.  The [R,N] activation is reduced on chip and never written."""
    x, ldx = _rows(x, "x")
    seg = _req(seg, torch.int32, "seg")
    R, K = x.shape
    assert K == w.K and seg.numel() >= R
    out = torch.zeros((nseg, w.N), dtype=torch.int32, device=x.device)   # order-preserving encoded maxima
    _lib.call("gnb_linear_tc_segmax", x.data_ptr(), R, K, ldx, w.packed.data_ptr(), w.cparams.data_ptr(), w.scale_log2, w.N,
              int(relu), seg.data_ptr(), out.data_ptr(), out.stride(0), _ptr(rows_dev), _stream())
    _lib.call("gnb_segmax_decode", out.data_ptr(), out.numel(), _stream())
    return out.view(torch.float32)


def linear_module(owner, slot: str, x: torch.Tensor, weight, bias=None, relu: bool = False,
                  flag_range: bool = False) -> torch.Tensor:
    """A plain ``nn.Linear`` (optionally + ReLU) applied to rows: tensor cores for large row counts, fp32 kernel below.
    ``flag_range``: outputs outside the fp16 range raise the range flag (in the epilogue on the tensor-core path, with a
    separate pass over the result otherwise)."""
    if USE_LINEAR_TC and x.shape[0] >= LINEAR_TC_MIN_ROWS:
        return linear_tc(x, packed_linear_for(owner, slot, weight, bias), relu, flag_range=flag_range)
    y = linear(x, weight, bias, relu)
    if flag_range:
        f16_range_check(y)
    return y


def nocs_head(logits: torch.Tensor, bins: int):
    logits = _req(logits, torch.float32, "logits")
    R = logits.shape[0]
    dev = logits.device
    b = torch.empty((R, 3), dtype=torch.int64, device=dev)
    conf = torch.empty((R, 3), dtype=torch.float32, device=dev)
    nocs = torch.empty((R, 3), dtype=torch.float32, device=dev)
    _lib.call("gnb_nocs_head", logits.data_ptr(), R, int(bins), b.data_ptr(), conf.data_ptr(), nocs.data_ptr(), _stream())
    return b, conf, nocs


def scatter_reduce(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str,
                   channels_last: bool = True) -> torch.Tensor:
    """``src`` logical [C, N] (any strides), ``index`` i64[N] -> logical [C, dim_size].

    With ``channels_last`` the result is physically [dim_size, C] (returned as a transposed view), which is what
    lets the unchanged reference caller hand a channels-last volume straight to the UNet
    (ref networks/conv_implicit_wnf.py:92-99)."""
    src = _req(src, torch.float32, "src", contiguous=False)
    index = _req(index, torch.int64, "index")
    C, N = src.shape
    dev = src.device
    scratch = torch.empty(max(dim_size, 1), dtype=torch.int32, device=dev)
    if channels_last:
        phys = torch.empty((dim_size, C), dtype=torch.float32, device=dev)
        out_sc, out_sm = 1, C
    else:
        phys = torch.empty((C, dim_size), dtype=torch.float32, device=dev)
        out_sc, out_sm = dim_size, 1
    _lib.call("gnb_scatter_reduce", src.data_ptr(), src.stride(0), src.stride(1), index.data_ptr(), N, C, dim_size,
              REDUCE[reduce], phys.data_ptr(), out_sc, out_sm, scratch.data_ptr(), _stream())
    return phys.t() if channels_last else phys


def aggregator_features(feat, nocs, sim_points, conf, batch, G: int, lower_corner=(0.0, 0.0, 0.0),
                        upper_corner=(1.0, 1.0, 1.0), include_point_feature: bool = True,
                        include_confidence_feature: bool = True):
    """Per-point aggregator rows [feat | offset in voxel, sim_points | confidence] and flat voxel indices
    (ref networks/conv_implicit_wnf.py:62-85); the optional blocks follow the reference's two flags."""
    import ctypes
    feat, ldf = _rows(feat, "feat")
    nocs = _req(nocs, torch.float32, "nocs")
    sim_points = _req(sim_points, torch.float32, "sim_points") if include_point_feature else None
    conf = _req(conf, torch.float32, "conf") if include_confidence_feature else None
    batch = _req(batch, torch.int64, "batch")
    N, Cf = feat.shape
    cols = Cf + (6 if include_point_feature else 0) + (3 if include_confidence_feature else 0)
    out = torch.empty((N, cols), dtype=torch.float32, device=feat.device)
    flat = torch.empty((N,), dtype=torch.int64, device=feat.device)
    lc = (ctypes.c_float * 3)(*[float(v) for v in lower_corner])
    uc = (ctypes.c_float * 3)(*[float(v) for v in upper_corner])
    _lib.call("gnb_aggregator_features", feat.data_ptr(), ldf, Cf, nocs.data_ptr(), _ptr(sim_points), _ptr(conf),
              batch.data_ptr(), N, int(G), ctypes.cast(lc, ctypes.c_void_p).value, ctypes.cast(uc, ctypes.c_void_p).value,
              1 if include_point_feature else 0, 1 if include_confidence_feature else 0, flat.data_ptr(), out.data_ptr(),
              out.stride(0), _stream())
    return out, flat


# ---------------------------------------------------------------------------------------------- UNet ops (NDHWC)
def groupnorm_stats(x: torch.Tensor, groups: int, eps: float, gamma, beta):
    """x contiguous [B,D,H,W,C] -> (scale, shift) f32[B,C] with GN(x) = x*scale + shift."""
    B, C = x.shape[0], x.shape[-1]
    voxels = x.numel() // (B * C)
    scale = torch.empty((B, C), dtype=torch.float32, device=x.device)
    shift = torch.empty((B, C), dtype=torch.float32, device=x.device)
    ws = torch.empty((B * groups * 2,), dtype=torch.float64, device=x.device)
    _lib.call("gnb_groupnorm_stats", x.data_ptr(), B, voxels, C, int(groups), float(eps), _ptr(gamma), _ptr(beta),
              scale.data_ptr(), shift.data_ptr(), ws.data_ptr(), _stream())
    return scale, shift


def groupnorm_stats_cat(skip: torch.Tensor, x_low: torch.Tensor, groups: int, eps: float, gamma, beta):
    """GroupNorm scale / shift of the virtual tensor cat((skip, nearest_upsample_2x(x_low)), channel) without materialising it."""
    B, D, H, W, Cs = skip.shape
    Cx = x_low.shape[-1]
    C = Cs + Cx
    scale = torch.empty((B, C), dtype=torch.float32, device=skip.device)
    shift = torch.empty((B, C), dtype=torch.float32, device=skip.device)
    ws = torch.empty((B * groups * 2,), dtype=torch.float64, device=skip.device)
    _lib.call("gnb_groupnorm_stats_cat", skip.data_ptr(), Cs, x_low.data_ptr(), Cx, B, D * H * W, int(groups), float(eps),
              _ptr(gamma), _ptr(beta), scale.data_ptr(), shift.data_ptr(), ws.data_ptr(), _stream())
    return scale, shift


def gn_apply_split_cat(skip: torch.Tensor, x_low: torch.Tensor, scale, shift):
    """cat((skip, upsample(x_low))) * scale + shift as fp16 hi + lo [B,D,H,W,Cpad] (see gn_apply_split)."""
    B, D, H, W, Cs = skip.shape
    Cx = x_low.shape[-1]
    cpad = (Cs + Cx + 63) // 64 * 64
    xh = torch.empty((B, D, H, W, cpad), dtype=torch.float16, device=skip.device)
    xl = torch.empty((B, D, H, W, cpad), dtype=torch.float16, device=skip.device)
    _lib.call("gnb_gn_apply_split_cat", skip.data_ptr(), Cs, x_low.data_ptr(), Cx, B, D, H, W, scale.data_ptr(), shift.data_ptr(),
              xh.data_ptr(), xl.data_ptr(), _stream())
    return xh, xl


def conv3d_k3(x: torch.Tensor, wt: torch.Tensor, scale=None, shift=None, relu: bool = True) -> torch.Tensor:
    """x [B,D,H,W,Cin] contiguous, wt [27,Cin,Cout] -> [B,D,H,W,Cout]."""
    B, D, H, W, Cin = x.shape
    Cout = wt.shape[2]
    y = torch.empty((B, D, H, W, Cout), dtype=torch.float32, device=x.device)
    _lib.call("gnb_conv3d_k3", x.data_ptr(), B, D, H, W, Cin, _ptr(scale), _ptr(shift), wt.data_ptr(), Cout, int(relu),
              y.data_ptr(), _stream())
    return y


def maxpool3d_2(x: torch.Tensor) -> torch.Tensor:
    B, D, H, W, C = x.shape
    y = torch.empty((B, D // 2, H // 2, W // 2, C), dtype=torch.float32, device=x.device)
    _lib.call("gnb_maxpool3d_2", x.data_ptr(), B, D, H, W, C, y.data_ptr(), _stream())
    return y


def upsample_concat(skip: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    B, D, H, W, Cs = skip.shape
    _, Dx, Hx, Wx, Cx = x.shape
    y = torch.empty((B, D, H, W, Cs + Cx), dtype=torch.float32, device=x.device)
    _lib.call("gnb_upsample_concat", skip.data_ptr(), Cs, x.data_ptr(), Cx, B, D, H, W, Dx, Hx, Wx, y.data_ptr(),
              _stream())
    return y


def to_channels_last(x: torch.Tensor) -> torch.Tensor:
    """logical NCDHW tensor with arbitrary strides -> contiguous [B,D,H,W,C] (no copy if it already is one)."""
    x = _req(x, torch.float32, "x", contiguous=False)
    B, C, D, H, W = x.shape
    y_view = x.permute(0, 2, 3, 4, 1)
    if y_view.is_contiguous():
        return y_view
    y = torch.empty((B, D, H, W, C), dtype=torch.float32, device=x.device)
    sb, sc, sd, sh, sw = x.stride()
    _lib.call("gnb_to_channels_last", x.data_ptr(), sb, sc, sd, sh, sw, B, C, D, H, W, y.data_ptr(), _stream())
    return y


# ---------------------------------------------------------------------------------------------- decoder ops
def trilinear_sample(vol: torch.Tensor, q: torch.Tensor, flip: bool = False, bn_scale=None, bn_shift=None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """vol [B,D,H,W,C] contiguous, q [B,M,3] -> [B*M, C]."""
    B, D, H, W, C = vol.shape
    q = _req(q, torch.float32, "q")
    M = q.shape[1]
    if out is None:
        out = torch.empty((B * M, C), dtype=torch.float32, device=vol.device)
    post = bn_scale is not None
    _lib.call("gnb_trilinear_sample", vol.data_ptr(), B, D, H, W, C, q.data_ptr(), M, int(flip), int(post),
              _ptr(bn_scale), _ptr(bn_shift), out.data_ptr(), out.stride(0), _stream())
    return out


def trilinear_sample_grid(vol: torch.Tensor, b: int, Q: int, m0: int, M: int, bn_scale=None, bn_shift=None,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _, D, H, W, C = vol.shape
    if out is None:
        out = torch.empty((M, C), dtype=torch.float32, device=vol.device)
    post = bn_scale is not None
    _lib.call("gnb_trilinear_sample_grid", vol.data_ptr(), int(b), D, H, W, C, int(Q), int(m0), int(M), int(post),
              _ptr(bn_scale), _ptr(bn_shift), out.data_ptr(), out.stride(0), _stream())
    return out


def _ggm_tmp(v: torch.Tensor, sigma: float) -> Optional[torch.Tensor]:
    """Workspace of the multi-pass form; the fused single-pass kernel (filter radius int(4*sigma + 0.5) <= 4, which
    covers the shipped sigma = 0.5) needs none."""
    if int(4.0 * float(sigma) + 0.5) <= 4:
        return None
    return torch.empty((2,) + tuple(v.shape), dtype=torch.float32, device=v.device)


def gaussian_gradient_magnitude(v: torch.Tensor, sigma: float) -> torch.Tensor:
    v = _req(v, torch.float32, "v")
    D, H, W = v.shape
    out = torch.empty_like(v)
    tmp = _ggm_tmp(v, sigma)
    _lib.call("gnb_gaussian_gradient_magnitude", v.data_ptr(), D, H, W, float(sigma), out.data_ptr(), _ptr(tmp),
              _stream())
    return out


def marching_cubes(volume: torch.Tensor, level: float, spacing=(1.0, 1.0, 1.0), gradient_direction: str = "ascent",
                   ggm: Optional[torch.Tensor] = None):
    """Device marching cubes (call shape of ``skimage.measure.marching_cubes``, ref predict.py:172-177): MC33 structure
    (face test, interior test, tunnel tilings) with scikit-image's conventions.  PARITY UNPINNED vs scikit-image: triangle
    order / vertex numbering inside a cell and the tunnel tilings can differ from its Lewiner tables (INTEGRATION.md
    section 5), so per-vertex outputs are not index-compatible with a scikit-image run.

    Returns (verts f32[V,3], faces i32[F,3], normals f32[V,3], values f32[V], ggm_at_verts f32[V] or None), all on the
    device.  Raises ValueError if ``level`` is outside the data range and RuntimeError if no surface is found, like
    scikit-image.  One host synchronisation (the vertex / face totals size the outputs)."""
    import ctypes
    volume = _req(volume, torch.float32, "volume")
    if volume.dim() != 3:
        raise ValueError("Input volume should be a 3D array.")
    if gradient_direction not in ("ascent", "descent"):
        raise ValueError("Incorrect input %s in `gradient_direction`" % gradient_direction)
    D, H, W = volume.shape
    dev = volume.device
    ws_bytes = _lib.load().gnb_mc_workspace_bytes(D, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    counts = (ctypes.c_int64 * 3)()
    _lib.call("gnb_mc_count", volume.data_ptr(), D, H, W, float(level), ws.data_ptr(),
              ctypes.cast(counts, ctypes.c_void_p).value, _stream())
    V, Fc, A = int(counts[0]), int(counts[1]), int(counts[2])
    if V == 0:
        raise RuntimeError("No surface found at the given iso value.")
    verts = torch.empty((V, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((Fc, 3), dtype=torch.int32, device=dev)
    normals = torch.empty((V, 3), dtype=torch.float32, device=dev)
    values = torch.empty((V,), dtype=torch.float32, device=dev)
    ggm_at = torch.empty((V,), dtype=torch.float32, device=dev) if ggm is not None else None
    sp = (ctypes.c_double * 3)(*[float(s) for s in spacing])
    _lib.call("gnb_mc_emit", volume.data_ptr(), D, H, W, float(level), ctypes.cast(sp, ctypes.c_void_p).value,
              1 if gradient_direction == "ascent" else 0, _ptr(ggm), ws.data_ptr(), A, V, verts.data_ptr(), faces.data_ptr(),
              normals.data_ptr(), values.data_ptr(), _ptr(ggm_at), _stream())
    return verts, faces, normals, values, ggm_at


# ---------------------------------------------------------------------------------------------- tensor-core decoder tail
def _pow2_scale(weight: torch.Tensor) -> int:
    """Power-of-two exponent s with max|w| * 2^s in [2^12, 2^13): keeps the fp16 lo parts of the split out of the
    subnormal range (one host read at pack time; packed weights are cached by the modules)."""
    m = float(weight.abs().max().item())
    return 0 if m == 0.0 else int(12 - math.floor(math.log2(m)))


def pack_f16_split(weight: torch.Tensor):
    """[256,256] fp32 weight -> (fp16 hi/lo shared-memory images for ``decode_tc`` uint8[N*K*4], scale_log2)."""
    weight = _req(weight, torch.float32, "weight")
    N, K = weight.shape
    s = _pow2_scale(weight)
    packed = torch.empty(N * K * 4, dtype=torch.uint8, device=weight.device)
    _lib.call("gnb_pack_f16_split", weight.data_ptr(), N, K, s, packed.data_ptr(), _stream())
    return packed, s


def decode_tc(w2_packed, b2, bn2, W3, b3, bn3, *, U=None, Q=0, bn1=None, X=None, out=None) -> torch.Tensor:
    w2_packed, w2_s = w2_packed
    """Tensor-core decoder tail (``gnb_decode_tc``).  Lattice mode: ``U`` [B,G,G,G,256], ``Q`` = 128, ``bn1`` =
    (scale, shift) -> [B, Q^3, Cout].  Row mode: ``X`` [R,256] -> [R, Cout].  bn2 / bn3 are (scale, shift) or None."""
    Cout = W3.shape[0]
    dev = W3.device
    scratch = torch.empty(1024, dtype=torch.float32, device=dev)
    s2, h2 = bn2 if bn2 is not None else (None, None)
    s3, h3 = bn3 if bn3 is not None else (None, None)
    if U is not None:
        B, G = U.shape[0], U.shape[1]
        assert U.is_contiguous() and U.shape[-1] == 256
        if out is None:
            out = torch.empty((B, Q ** 3, Cout), dtype=torch.float32, device=dev)
        _lib.call("gnb_decode_tc", U.data_ptr(), 0, B, G, int(Q), 0, bn1[0].data_ptr(), bn1[1].data_ptr(),
                  w2_packed.data_ptr(), w2_s, b2.data_ptr(), _ptr(s2), _ptr(h2), W3.data_ptr(), _ptr(b3), _ptr(s3), _ptr(h3),
                  Cout, scratch.data_ptr(), out.data_ptr(), _stream())
        return out
    X, ldx = _rows(X, "X")
    R = X.shape[0]
    if out is None:
        out = torch.empty((R, Cout), dtype=torch.float32, device=dev)
    _lib.call("gnb_decode_tc", X.data_ptr(), ldx, 0, 0, 0, R, None, None, w2_packed.data_ptr(), w2_s, b2.data_ptr(), _ptr(s2),
              _ptr(h2), W3.data_ptr(), _ptr(b3), _ptr(s3), _ptr(h3), Cout, scratch.data_ptr(), out.data_ptr(), _stream())
    return out


def decode_lattice(W2, w2f_packed, b2, bn1_shift, bn2, W3, b3, bn3, *, U, Q, out=None) -> torch.Tensor:
    """Pair-tile lattice kernel (``gnb_decode_lattice``): ``U`` [B,G,G,G,256] hoisted grid, ``w2f_packed`` =
    ``pack_f16_split(W2 * bn1_scale[None, :])`` -> [B, Q^3, Cout]."""
    w2f, w2f_s = w2f_packed
    Cout = W3.shape[0]
    dev = W3.device
    B, G = U.shape[0], U.shape[1]
    assert U.is_contiguous() and U.shape[-1] == 256
    if out is None:
        out = torch.empty((B, Q ** 3, Cout), dtype=torch.float32, device=dev)
    s2, h2 = bn2 if bn2 is not None else (None, None)
    s3, h3 = bn3 if bn3 is not None else (None, None)
    scratch = torch.empty(2048, dtype=torch.float32, device=dev)
    _lib.call("gnb_decode_lattice", U.data_ptr(), B, G, int(Q), W2.data_ptr(), w2f.data_ptr(), w2f_s, b2.data_ptr(),
              bn1_shift.data_ptr(), _ptr(s2), _ptr(h2), W3.data_ptr(), _ptr(b3), _ptr(s3), _ptr(h3), Cout,
              scratch.data_ptr(), out.data_ptr(), _stream())
    return out


def decode_tc_query(w2_packed, b2, bn2, W3, b3, bn3, *, U, q, qptr, bn1, out=None) -> torch.Tensor:
    """Query mode of the tensor-core decoder (``gnb_decode_tc_query``): ``U`` [B,G,G,G,256] hoisted grids, ``q`` [R,3]
    query points of all samples back to back, ``qptr`` device i64[B+1] row offsets -> [R, Cout]."""
    w2_packed, w2_s = w2_packed
    Cout = W3.shape[0]
    dev = W3.device
    q = _req(q, torch.float32, "q")
    qptr = _req(qptr, torch.int64, "qptr")
    B, G = U.shape[0], U.shape[1]
    assert U.is_contiguous() and U.shape[-1] == 256 and qptr.numel() == B + 1
    R = q.shape[0]
    if out is None:
        out = torch.empty((R, Cout), dtype=torch.float32, device=dev)
    s2, h2 = bn2 if bn2 is not None else (None, None)
    s3, h3 = bn3 if bn3 is not None else (None, None)
    scratch = torch.empty(1024, dtype=torch.float32, device=dev)
    _lib.call("gnb_decode_tc_query", U.data_ptr(), B, G, q.data_ptr(), qptr.data_ptr(), R, bn1[0].data_ptr(),
              bn1[1].data_ptr(), w2_packed.data_ptr(), w2_s, b2.data_ptr(), _ptr(s2), _ptr(h2), W3.data_ptr(), _ptr(b3),
              _ptr(s3), _ptr(h3), Cout, scratch.data_ptr(), out.data_ptr(), _stream())
    return out


def decode_tc_query_fused(W1, b1, w2_packed, b2, bn2, W3, b3, bn3, *, X, q, qptr, bn1, out=None) -> torch.Tensor:
    """Query mode on the 32-channel grid (``gnb_decode_tc_query_fused``): ``X`` [B,G,G,G,32] channels-last, ``W1`` [256,32]
    / ``b1`` [256] = the decoder's first Linear (with the UNet's final_conv folded in), applied per query inside the
    kernel; ``q`` [R,3] query points of all samples back to back, ``qptr`` device i64[B+1] -> [R, Cout]."""
    w2_packed, w2_s = w2_packed
    Cout = W3.shape[0]
    dev = W3.device
    q = _req(q, torch.float32, "q")
    qptr = _req(qptr, torch.int64, "qptr")
    X = _req(X, torch.float32, "X")
    W1 = _req(W1, torch.float32, "W1")
    B, G, C0 = X.shape[0], X.shape[1], X.shape[-1]
    assert W1.shape == (256, C0) and qptr.numel() == B + 1
    R = q.shape[0]
    if out is None:
        out = torch.empty((R, Cout), dtype=torch.float32, device=dev)
    s2, h2 = bn2 if bn2 is not None else (None, None)
    s3, h3 = bn3 if bn3 is not None else (None, None)
    scratch = torch.empty(16384 + 256 * 512, dtype=torch.float32, device=dev)
    s1, h1 = bn1 if bn1 is not None else (None, None)   # None: BatchNorm1 already folded into w2_packed / b2
    _lib.call("gnb_decode_tc_query_fused", X.data_ptr(), B, G, C0, W1.data_ptr(), b1.data_ptr(), q.data_ptr(), qptr.data_ptr(),
              R, _ptr(s1), _ptr(h1), w2_packed.data_ptr(), w2_s, b2.data_ptr(), _ptr(s2), _ptr(h2),
              W3.data_ptr(), _ptr(b3), _ptr(s3), _ptr(h3), Cout, scratch.data_ptr(), out.data_ptr(), _stream())
    return out


# ---------------------------------------------------------------------------------------------- tensor-core 3x3x3 conv
_CROSS_PRECISION = [None]   # mirror of the library's process-wide mode (avoids a ctypes call per layer)


def conv_tc_cross_precision() -> int:
    """0: fp16 cross terms (three tensor passes), 1: e4m3 cross terms (two pass-equivalents); see
    ``gnb_conv_tc_set_cross_precision`` in the header."""
    if _CROSS_PRECISION[0] is None:
        _CROSS_PRECISION[0] = int(_lib.load().gnb_conv_tc_cross_precision())
    return _CROSS_PRECISION[0]


def conv_tc_set_cross_precision(mode: int) -> None:
    """Process-wide; packed weights are re-packed on their next use (the module caches are keyed on the mode)."""
    _lib.call("gnb_conv_tc_set_cross_precision", int(mode))
    _CROSS_PRECISION[0] = int(mode)


def conv3d_tc_supported(B, D, H, W, Cin, Cout) -> bool:
    return bool(_lib.call("gnb_conv3d_tc_supported", int(B), int(D), int(H), int(W), int(Cin), int(Cout)))


def conv3d_tc_pack_weights(weight: torch.Tensor):
    """[Cout,Cin,3,3,3] fp32 -> (fp16 hi/lo shared-memory images per (tap, 64-channel chunk), scale_log2)."""
    weight = _req(weight, torch.float32, "weight")
    Cout, Cin = weight.shape[:2]
    cpad = (Cin + 63) // 64 * 64
    s = _pow2_scale(weight)
    packed = torch.empty(27 * cpad * Cout * 4, dtype=torch.uint8, device=weight.device)
    _lib.call("gnb_conv3d_tc_pack_weights", weight.data_ptr(), Cout, Cin, s, packed.data_ptr(), _stream())
    return packed, s


def conv3d_tc_dx_supported(B, D, H, W, Cin, Cout) -> bool:
    return bool(_lib.call("gnb_conv3d_tc_dx_supported", int(B), int(D), int(H), int(W), int(Cin), int(Cout)))


def conv3d_tc_dx_pack_weights(weight: torch.Tensor):
    """[32|64,Cin,3,3,3] fp32 -> (stacked-kw fp16 hi/lo images per ((kd,kh), 64-channel chunk), scale_log2)."""
    weight = _req(weight, torch.float32, "weight")
    Cout, Cin = weight.shape[:2]
    cpad = (Cin + 63) // 64 * 64
    s = _pow2_scale(weight)
    packed = torch.empty(27 * cpad * Cout * 4, dtype=torch.uint8, device=weight.device)
    _lib.call("gnb_conv3d_tc_dx_pack_weights", weight.data_ptr(), Cout, Cin, s, packed.data_ptr(), _stream())
    return packed, s


def conv3d_tc_dx(xh: torch.Tensor, xl: torch.Tensor, cin: int, w_packed, cout: int, relu: bool = True):
    """Stacked-dx tensor-core convolution for Cout in {32, 64} (``gnb_conv3d_tc_dx``); same contract as ``conv3d_tc``."""
    w_packed, w_s = w_packed
    B, D, H, W, _ = xh.shape
    y = torch.empty((B, D, H, W, cout), dtype=torch.float32, device=xh.device)
    _lib.call("gnb_conv3d_tc_dx", xh.data_ptr(), xl.data_ptr(), B, D, H, W, int(cin), w_packed.data_ptr(), w_s, int(cout),
              int(relu), y.data_ptr(), _stream())
    return y


def gn_apply_split(x: torch.Tensor, scale, shift):
    """x [B,D,H,W,C] fp32 -> (xh, xl) fp16 [B,D,H,W,Cpad] holding x*scale+shift as hi + lo."""
    B, D, H, W, C = x.shape
    cpad = (C + 63) // 64 * 64
    xh = torch.empty((B, D, H, W, cpad), dtype=torch.float16, device=x.device)
    xl = torch.empty((B, D, H, W, cpad), dtype=torch.float16, device=x.device)
    _lib.call("gnb_gn_apply_split", x.data_ptr(), B, D * H * W, C, _ptr(scale), _ptr(shift), xh.data_ptr(), xl.data_ptr(),
              _stream())
    return xh, xl


def conv3d_tc(xh: torch.Tensor, xl: torch.Tensor, cin: int, w_packed, cout: int, relu: bool = True):
    w_packed, w_s = w_packed
    B, D, H, W, _ = xh.shape
    y = torch.empty((B, D, H, W, cout), dtype=torch.float32, device=xh.device)
    _lib.call("gnb_conv3d_tc", xh.data_ptr(), xl.data_ptr(), B, D, H, W, int(cin), w_packed.data_ptr(), w_s, int(cout),
              int(relu), y.data_ptr(), _stream())
    return y


# ---------------------------------------------------------------------------------------------- batched predict tail
def gaussian_gradient_magnitude_batched(v: torch.Tensor, sigma: float) -> torch.Tensor:
    """v [N,D,H,W]: N independent volumes in one set of launches."""
    v = _req(v, torch.float32, "v")
    N, D, H, W = v.shape
    out = torch.empty_like(v)
    tmp = _ggm_tmp(v, sigma)
    _lib.call("gnb_gaussian_gradient_magnitude_batched", v.data_ptr(), N, D, H, W, float(sigma), out.data_ptr(),
              _ptr(tmp), _stream())
    return out


USE_SA_MLP = True          # False: PointConv runs as gather + three Linear blocks + segment max (the unfused chain)


def pointconv_mlp_supported(cin: int, c1: int, c2: int, c3: int) -> bool:
    return bool(_lib.call("gnb_pointconv_mlp_supported", int(cin), int(c1), int(c2), int(c3)))


def pointconv_mlp_pack(owner, layers):
    """``layers`` = three (weight, bias, bn_scale, bn_shift) tuples of a PointConv message MLP -> cached (packed images,
    scales, constants) for :func:`pointconv_mlp_max`.  BatchNorm1 / 2 (they follow a ReLU) are folded into the next layer."""
    key = _version_key(*[t for layer in layers for t in layer])
    cached = getattr(owner, "_gnb_sa_mlp", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    (w1, b1, s1, t1), (w2, b2, s2, t2), (w3, b3, s3, t3) = layers
    with torch.no_grad():
        w2f = (w2 * s1[None, :]).contiguous()
        b2f = b2 + w2 @ t1
        w3f = (w3 * s2[None, :]).contiguous()
        b3f = b3 + w3 @ t2
        c1, cin = w1.shape[0], w1.shape[1] - 3
        c2, c3 = w2.shape[0], w3.shape[0]
        small = w1[:, cin:] if cin >= 16 else w1                       # layer-1 weights applied in fp32: [C1, KIN]
        consts = torch.cat([b1, small.t().contiguous().view(-1), b2f, b3f, s3, t3]).contiguous().float()
        scales = (_pow2_scale(w1[:, :cin]) if cin >= 16 else 0, _pow2_scale(w2f), _pow2_scale(w3f))
        packed = torch.empty(int(_lib.load().gnb_pointconv_mlp_packed_bytes(cin, c1, c2, c3)), dtype=torch.uint8, device=w1.device)
        _lib.call("gnb_pointconv_mlp_pack", w1.contiguous().data_ptr(), w2f.data_ptr(), w3f.data_ptr(), cin, c1, c2, c3, *scales,
                  packed.data_ptr(), _stream())
    val = (packed, scales, consts, (cin, c1, c2, c3))
    owner._gnb_sa_mlp = (key, val)
    return val


def pointconv_mlp_max(x, pos_x, pos_y, nbr, cnt, eoffs, pack) -> torch.Tensor:
    """Fused PointConv (``gnb_pointconv_mlp_max``): message MLP over the grouped edges + max aggregation -> [M, C3]."""
    packed, scales, consts, (cin, c1, c2, c3) = pack
    M, K = nbr.shape
    out = torch.empty((M, c3), dtype=torch.float32, device=pos_x.device)
    ws = torch.empty(2 * M * (K + 1), dtype=torch.int32, device=pos_x.device)
    _lib.call("gnb_pointconv_mlp_max", _ptr(x), x.stride(0) if x is not None else 0, cin, pos_x.data_ptr(), pos_y.data_ptr(),
              nbr.data_ptr(), cnt.data_ptr(), eoffs.data_ptr(), M, K, packed.data_ptr(), c1, c2, c3, *scales, consts.data_ptr(),
              ws.data_ptr(), out.data_ptr(), _stream())
    return out


def f16_range_check(t: torch.Tensor) -> None:
    """Queue a check of ``t`` (fp32, any shape) against the fp16 range of the tensor-core operand split; the outcome is read
    with :func:`f16_overflow` (one flag per device, see include/garmentnets_b200.h)."""
    t = _req(t, torch.float32, "t")
    _lib.call("gnb_f16_range_check", t.data_ptr(), t.numel(), _stream())


def f16_overflow(reset: bool = True) -> bool:
    """True if an fp32 value outside +-65504 (or a non-finite one) reached a saturating fp16 split since the last reset.
    Synchronises the current stream."""
    return _lib.call("gnb_f16_overflow_fetch", 1 if reset else 0, _stream()) == 1


def f16_overflow_async(pinned: torch.Tensor, reset: bool = True) -> None:
    """Stream-ordered copy of the flag into ``pinned`` (int32[1], pinned host memory): valid after the next synchronisation of
    the current stream."""
    assert pinned.is_pinned() and pinned.dtype == torch.int32 and pinned.numel() >= 1
    _lib.call("gnb_f16_overflow_fetch_async", pinned.data_ptr(), 1 if reset else 0, _stream())


_MC_RECORDS: Dict[Tuple[int, int], torch.Tensor] = {}


def _mc_record_buffer(dev: torch.device, n: int) -> torch.Tensor:
    """Pinned [>= n, 512] byte buffer for the per-volume marching-cubes records (one per device and host thread, grown on
    demand; the caller consumes it before it returns)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), threading.get_ident())
    buf = _MC_RECORDS.get(key)
    if buf is None or buf.shape[0] < n:
        buf = _MC_RECORDS[key] = torch.zeros((max(n, 64), 512), dtype=torch.uint8).pin_memory()
    return buf


def marching_cubes_batch(volumes: torch.Tensor, level: float, spacing=(1.0, 1.0, 1.0), gradient_direction: str = "ascent",
                         ggm: Optional[torch.Tensor] = None, return_packed: bool = False, with_normals: bool = True,
                         lazy_views: bool = False):
    """Marching cubes of N volumes [N,D,H,W] in seven launches and ONE host synchronisation: classify / scan for the
    whole batch, the N 512-byte records come back in a single device->host copy, then compaction, one vertex launch (a
    thread per vertex) and one face launch (a thread per active cell) write every mesh into shared [sum V] / [sum F]
    buffers.
    Returns a list of (verts, faces, normals, values, ggm_at) views or the exception skimage would raise for that volume
    (ValueError: level outside the data range, RuntimeError: no surface).  ``return_packed`` adds the shared buffers:
    ``{"verts": f32[sum V,3], "vptr": host i64[N+1] row offsets, "faces": i32[sum F,3], "fptr", "normals", "values",
    "ggm_at"}`` (volumes without a mesh own zero rows).  ``with_normals=False`` skips the per-vertex normals and values
    (``None`` in their place): the reference stores them (predict.py:193-200) but nothing downstream reads them.
    ``lazy_views`` (with ``return_packed``) leaves ``None`` instead of the per-volume views: 5 x N tensor slices cost ~1 ms of
    host time for a batch of 32, which the caller should spend AFTER queueing the kernels that consume the packed buffers.
    The triangulation follows the MC33 structure but is NOT pinned to scikit-image's Lewiner tables: vertex / face order and
    the tiling of ambiguous cells can differ from ``skimage.measure.marching_cubes``."""
    import ctypes
    import numpy as np
    volumes = _req(volumes, torch.float32, "volumes")
    if ggm is not None:
        ggm = _req(ggm, torch.float32, "ggm")
    if gradient_direction not in ("ascent", "descent"):
        raise ValueError("Incorrect input %s in `gradient_direction`" % gradient_direction)
    N, D, H, W = volumes.shape
    dev = volumes.device
    lib = _lib.load()
    ws_bytes = (int(lib.gnb_mc_workspace_bytes(D, H, W)) + 255) // 256 * 256
    off = int(lib.gnb_mc_totals_offset(D, H, W))
    ws = torch.empty((N, ws_bytes), dtype=torch.uint8, device=dev)
    _lib.call("gnb_mc_count_batch", volumes.data_ptr(), N, D, H, W, float(level), ws.data_ptr(), ws_bytes, _stream())
    # the one synchronisation: the N records are stored into pinned host memory by a kernel (not a cudaMemcpy, which would queue
    # behind the previous batch's bulk result transfers on the copy engine) and the host waits on an event behind it
    rec_host = _mc_record_buffer(dev, N)
    _lib.call("gnb_copy_to_pinned_host", ws.data_ptr() + off, ws_bytes, rec_host.data_ptr(), 512, 512, N, _stream())
    done = torch.cuda.Event()
    done.record()
    # everything that does not need the totals happens before the wait: the device is idle from here until the emission launch
    sp = (ctypes.c_double * 3)(*[float(x) for x in spacing])
    sp_ptr = ctypes.cast(sp, ctypes.c_void_p).value
    ascent = 1 if gradient_direction == "ascent" else 0
    stream = _stream()
    rec = rec_host[:N].numpy()
    done.synchronize()
    totals = rec[:, :40].copy().view(np.int64)  # V, F, A, vbase, fbase
    sumV, sumF = int(totals[:, 0].sum()), int(totals[:, 1].sum())
    verts = torch.empty((sumV, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((sumF, 3), dtype=torch.int32, device=dev)
    normals = torch.empty((sumV, 3), dtype=torch.float32, device=dev) if with_normals else None
    values = torch.empty((sumV,), dtype=torch.float32, device=dev) if with_normals else None
    ggm_at = torch.empty((sumV,), dtype=torch.float32, device=dev) if ggm is not None else None
    if sumV > 0:
        _lib.call("gnb_mc_emit_batch", volumes.data_ptr(), N, D, H, W, float(level), sp_ptr, ascent, _ptr(ggm), ws.data_ptr(),
                  ws_bytes, int(totals[:, 2].max()), int(totals[:, 0].max()), verts.data_ptr(), faces.data_ptr(), _ptr(normals),
                  _ptr(values), _ptr(ggm_at), stream)
    # (the value range of every volume, for skimage's ValueError: decoded while the emission kernels run)
    enc = rec[:, 256:264].copy().view(np.uint32)
    dec = np.where(enc & 0x80000000, enc & 0x7FFFFFFF, ~enc).astype(np.uint32).view(np.float32)
    out = []
    for i in range(N):
        V, Fc, _, vb, fb = (int(t) for t in totals[i])
        if level < float(dec[i, 0]) or level > float(dec[i, 1]):
            out.append(ValueError("Surface level must be within volume data range."))
        elif V == 0:
            out.append(RuntimeError("No surface found at the given iso value."))
        elif lazy_views:
            out.append(None)     # the caller slices the packed buffers itself (after it has queued its next kernels)
        else:
            out.append((verts[vb:vb + V], faces[fb:fb + Fc], normals[vb:vb + V] if with_normals else None,
                        values[vb:vb + V] if with_normals else None, ggm_at[vb:vb + V] if ggm_at is not None else None))
    if return_packed:
        vptr = np.zeros(N + 1, np.int64)
        np.cumsum(totals[:, 0], out=vptr[1:])
        fptr = np.zeros(N + 1, np.int64)
        np.cumsum(totals[:, 1], out=fptr[1:])
        return out, {"verts": verts, "vptr": vptr, "faces": faces, "fptr": fptr, "normals": normals, "values": values,
                     "ggm_at": ggm_at}
    return out
