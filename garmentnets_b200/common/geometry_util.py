"""Mesh geometry helpers -- mirror of the parts of the reference's ``common/geometry_util.py`` that evaluation of a
prediction uses (``barycentric_interpolation`` :160-181, ``mesh_sample_barycentric`` :184-223, called at eval.py:222-243)
plus the connected-component step of eval.py:538-546 (``igl.adjacency_matrix`` / ``igl.connected_components``), on the
sm_100a kernels.  Same names, argument meaning and results; tensors live on the device and there
is no CPU fallback.  The only host work is what the reference itself does on the host: drawing the uniform variates from
``numpy.random.RandomState(seed)`` -- they ARE its random stream (M + 2M doubles)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib, ops


def _faces_i32(faces: torch.Tensor) -> torch.Tensor:
    if not isinstance(faces, torch.Tensor) or not faces.is_cuda:
        raise _lib.GarmentNetsB200Error("faces: expected a CUDA tensor (the hot path has no CPU fallback)")
    return (faces if faces.dtype == torch.int32 else faces.to(torch.int32)).contiguous()


def _field(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.GarmentNetsB200Error(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype not in (torch.float32, torch.float64):
        raise _lib.GarmentNetsB200Error(f"{name}: expected float32 or float64, got {t.dtype}")
    return t.contiguous(), int(t.dtype == torch.float64)


def barycentric_interpolation(query_coords: torch.Tensor, verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """ref common/geometry_util.py:160-181.  ``query_coords`` f64[M,3] barycentric coordinates, ``verts`` [N,C] any
    per-vertex field, ``faces`` [M,3] the vertex ids of the face each query lies in (1:1 with ``query_coords``)."""
    verts, f64 = _field(verts, "verts")
    faces = _faces_i32(faces)
    bary = ops._req(query_coords, torch.float64, "query_coords")
    M, C = bary.shape[0], verts.shape[1]
    out = torch.empty((M, C), dtype=verts.dtype, device=verts.device)
    ident = torch.arange(M, dtype=torch.int64, device=verts.device)
    _lib.call("gnb_barycentric_interpolation", bary.data_ptr(), ident.data_ptr(), faces.data_ptr(), verts.data_ptr(), f64, C, M,
              out.data_ptr(), ops._stream())
    return out


def interpolate_on_faces(bary: torch.Tensor, face_idx: torch.Tensor, faces: torch.Tensor, field: torch.Tensor) -> torch.Tensor:
    """``barycentric_interpolation(bary, field, faces[face_idx])`` without materialising ``faces[face_idx]``."""
    field, f64 = _field(field, "field")
    faces = _faces_i32(faces)
    bary = ops._req(bary, torch.float64, "bary")
    face_idx = ops._req(face_idx, torch.int64, "face_idx")
    M, C = bary.shape[0], field.shape[1]
    out = torch.empty((M, C), dtype=field.dtype, device=field.device)
    _lib.call("gnb_barycentric_interpolation", bary.data_ptr(), face_idx.data_ptr(), faces.data_ptr(), field.data_ptr(), f64, C, M,
              out.data_ptr(), ops._stream())
    return out


def mesh_sample_barycentric(verts: torch.Tensor, faces: torch.Tensor, num_samples: int, seed: Optional[int] = None,
                            face_areas: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """ref common/geometry_util.py:184-223: uniform samples on the surface as ``(barycentric_all f64[num_samples,3],
    selected_face_idx [num_samples] in faces.dtype)``.  The random stream is numpy's ``RandomState(seed)`` exactly as the
    reference consumes it (``choice`` draws ``random_sample(num_samples)``, then ``uniform(0, 1, (num_samples, 2))``)."""
    verts, f64 = _field(verts, "verts")
    f32i = _faces_i32(faces)
    F = f32i.shape[0]
    if F < 1:
        raise ValueError("mesh_sample_barycentric: the mesh has no faces")
    rs = np.random.RandomState(seed=seed)
    u_face = rs.random_sample(num_samples)
    uv = rs.uniform(0, 1, size=(num_samples, 2))
    dev = verts.device
    host = torch.from_numpy(np.concatenate([u_face, uv.reshape(-1)])).to(dev)
    cdf_ws = torch.empty(2 * F, dtype=torch.float64, device=dev)
    face_idx = torch.empty(num_samples, dtype=torch.int64, device=dev)
    bary = torch.empty((num_samples, 3), dtype=torch.float64, device=dev)
    areas = None if face_areas is None else ops._req(face_areas.to(torch.float64), torch.float64, "face_areas")
    _lib.call("gnb_mesh_sample_barycentric", verts.data_ptr(), f64, f32i.data_ptr(), F, ops._ptr(areas), host.data_ptr(),
              host[num_samples:].data_ptr(), int(num_samples), cdf_ws.data_ptr(), face_idx.data_ptr(), bary.data_ptr(),
              ops._stream())
    return bary, face_idx.to(faces.dtype)


def connected_components_batch(faces: torch.Tensor, vptr_host: Sequence[int], fptr_host: Sequence[int]):
    """Packed meshes of a batch (``faces`` i32[sum F,3] with per-sample LOCAL ids, host offsets).  Returns
    ``(is_largest bool[sum V], labels i32[sum V], summary i64[B,3] host)``: membership in the sample's largest component
    (most vertices, first on ties), component label = local id of the component's lowest vertex, and per sample
    {number of components, size of the largest, its label}."""
    faces = _faces_i32(faces)
    vptr_host = np.asarray(vptr_host, dtype=np.int64)
    fptr_host = np.asarray(fptr_host, dtype=np.int64)
    B = len(vptr_host) - 1
    V, F = int(vptr_host[-1]), int(fptr_host[-1])
    if B < 1 or len(fptr_host) != B + 1 or faces.shape[0] != F:
        raise ValueError("connected_components_batch: inconsistent offsets / array sizes")
    dev = faces.device
    ptrs = torch.from_numpy(np.concatenate([vptr_host, fptr_host])).to(dev)
    ws = torch.empty((2, max(V, 1)), dtype=torch.int32, device=dev)
    is_largest = torch.empty(max(V, 1), dtype=torch.uint8, device=dev)
    labels = torch.empty(max(V, 1), dtype=torch.int32, device=dev)
    summary = torch.empty((B, 3), dtype=torch.int64, device=dev)
    _lib.call("gnb_mesh_components", faces.data_ptr(), ptrs[B + 1:].data_ptr(), ptrs[:B + 1].data_ptr(), B, V, F, ws[0].data_ptr(),
              ws[1].data_ptr(), is_largest.data_ptr(), labels.data_ptr(), summary.data_ptr(), ops._stream())
    return is_largest[:V].view(torch.bool), labels[:V], summary.cpu().numpy()


def connected_components(faces: torch.Tensor, num_verts: Optional[int] = None):
    """``igl.connected_components(igl.adjacency_matrix(faces))`` for one mesh (ref eval.py:538-539): returns
    ``(num_cc, cc_idxs i32[V], cc_sizes i64[num_cc])`` with components numbered 0.. in order of their lowest vertex."""
    n = int(num_verts) if num_verts is not None else (int(faces.max().item()) + 1 if faces.numel() else 0)
    _, labels, summary = connected_components_batch(faces, [0, n], [0, faces.shape[0]])
    roots, cc_idxs = torch.unique(labels, sorted=True, return_inverse=True)   # plumbing: relabel lowest-vertex ids 0..num_cc-1
    cc_sizes = torch.bincount(cc_idxs, minlength=roots.numel())
    return int(summary[0, 0]), cc_idxs.to(torch.int32), cc_sizes


def largest_component_mask(faces: torch.Tensor, num_verts: int) -> torch.Tensor:
    """``is_cc_vert`` of eval.py:538-541: membership of every vertex in the component with the most vertices."""
    mask, _, _ = connected_components_batch(faces, [0, int(num_verts)], [0, faces.shape[0]])
    return mask
