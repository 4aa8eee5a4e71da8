"""Mesh clean-up after marching cubes -- mirror of the reference's ``common/marching_cubes_util.py``
(``wnf_to_mesh`` :5-35, ``delete_invalid_verts`` :38-52) on the sm_100a kernels.

The reference turns the closed winding-number iso-surface into the open garment by keeping only the faces whose three
vertices all pass a per-vertex test (Gaussian gradient magnitude above a threshold in ``wnf_to_mesh``, predicted value
above a threshold in eval.py:532-546), deleting the vertices no kept face uses and re-indexing the faces.  Same
signatures and results (vertex order = ascending original index, like ``np.unique``); tensors stay on the device and
there is no CPU fallback.  ``delete_invalid_verts_batch`` does a whole batch of meshes with one host synchronisation.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib, ops


def delete_invalid_verts_batch(faces: torch.Tensor, vptr_host: Sequence[int], fptr_host: Sequence[int],
                               is_vert_on_surface: torch.Tensor):
    """Packed meshes of a batch: ``faces`` i32[sum F, 3] with per-sample LOCAL vertex ids, ``vptr_host`` / ``fptr_host``
    host i64[B+1] row offsets, ``is_vert_on_surface`` bool[sum V].  Returns ``(keep, valid_faces, new_vptr, new_fptr)``:
    ``keep`` i64[sum V'] = surviving GLOBAL vertex rows in ascending order (gather any per-vertex array with it),
    ``valid_faces`` i32[sum F', 3] re-indexed (local), and the new host offsets."""
    faces = ops._req(faces, torch.int32, "faces")
    on = ops._req(is_vert_on_surface, torch.bool, "is_vert_on_surface")
    vptr_host = np.asarray(vptr_host, dtype=np.int64)
    fptr_host = np.asarray(fptr_host, dtype=np.int64)
    B = len(vptr_host) - 1
    V, F = int(vptr_host[-1]), int(fptr_host[-1])
    if B < 1 or len(fptr_host) != B + 1 or on.numel() != V or faces.shape[0] != F:
        raise ValueError("delete_invalid_verts_batch: inconsistent offsets / array sizes")
    dev = faces.device
    lib = _lib.load()
    ws = torch.empty(int(lib.gnb_mesh_cleanup_workspace_bytes(V, F)), dtype=torch.uint8, device=dev)
    ptrs = torch.from_numpy(np.concatenate([vptr_host, fptr_host])).to(dev)
    vptr, fptr = ptrs[:B + 1], ptrs[B + 1:]
    rec = torch.empty(2 * (B + 1), dtype=torch.int64, device=dev)
    on_u8 = on.view(torch.uint8)
    _lib.call("gnb_mesh_cleanup_count", faces.data_ptr(), fptr.data_ptr(), vptr.data_ptr(), B, V, F, on_u8.data_ptr(),
              ws.data_ptr(), rec.data_ptr(), ops._stream())
    rec_host = rec.cpu().numpy()                       # the one synchronisation: new per-sample offsets
    new_vptr, new_fptr = rec_host[:B + 1].copy(), rec_host[B + 1:].copy()
    keep = torch.empty((int(new_vptr[-1]),), dtype=torch.int64, device=dev)
    valid_faces = torch.empty((int(new_fptr[-1]), 3), dtype=torch.int32, device=dev)
    _lib.call("gnb_mesh_cleanup_emit", faces.data_ptr(), fptr.data_ptr(), vptr.data_ptr(), B, V, F, ws.data_ptr(),
              rec.data_ptr(), keep.data_ptr(), valid_faces.data_ptr(), ops._stream())
    return keep, valid_faces, new_vptr, new_fptr


def delete_invalid_verts(mc_verts: torch.Tensor, mc_faces: torch.Tensor, is_vert_on_surface: torch.Tensor
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """ref common/marching_cubes_util.py:38-52: ``(valid_verts, valid_faces)`` of one mesh."""
    faces = mc_faces if mc_faces.dtype == torch.int32 else mc_faces.to(torch.int32)
    keep, valid_faces, _, _ = delete_invalid_verts_batch(faces.contiguous(), [0, mc_verts.shape[0]], [0, faces.shape[0]],
                                                         is_vert_on_surface)
    return mc_verts.index_select(0, keep), valid_faces.to(mc_faces.dtype)


def wnf_to_mesh(wnf_volume: torch.Tensor, iso_surface_level: float = 0.5, gradient_threshold: float = 0.25,
                sigma: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """ref common/marching_cubes_util.py:5-35: winding-number volume [D,H,W] -> open garment mesh
    ``(valid_verts f32[V',3], valid_faces i32[F',3])``.  Like the reference, the Gaussian gradient magnitude always uses
    sigma = 0.5 (:7-8 ignores the ``sigma`` argument)."""
    volume_size = wnf_volume.shape[-1]
    wnf_ggm = ops.gaussian_gradient_magnitude(wnf_volume, 0.5)
    voxel_spacing = 1 / (volume_size - 1)
    mc_verts, mc_faces, _, _, mc_verts_ggm = ops.marching_cubes(wnf_volume, iso_surface_level, (voxel_spacing,) * 3, "ascent",
                                                                wnf_ggm)
    return delete_invalid_verts(mc_verts, mc_faces, mc_verts_ggm > gradient_threshold)


def clean_batch(results: List[dict], vptr_host, fptr_host, packed: dict, is_vert_on_surface: torch.Tensor) -> List[dict]:
    """Apply the clean-up to the packed output of ``ConvImplicitWNFPipeline.predict`` (``model._last_packed``): returns one
    dict per sample with every per-vertex array gathered and the faces re-indexed."""
    keep, faces, nv, nf = delete_invalid_verts_batch(packed["faces"], vptr_host, fptr_host, is_vert_on_surface)
    per_vertex = {k: v.index_select(0, keep) for k, v in (("verts", packed["verts"]), ("normals", packed["normals"]),
                                                          ("volume_value", packed["values"]),
                                                          ("volume_gradient_magnitude", packed["ggm_at"]),
                                                          ("warp_field", packed.get("warp_field"))) if v is not None}
    out = []
    for b in range(len(nv) - 1):
        r = {k: v[int(nv[b]):int(nv[b + 1])] for k, v in per_vertex.items()}
        r["faces"] = faces[int(nf[b]):int(nf[b + 1])]
        out.append(r)
    return out
