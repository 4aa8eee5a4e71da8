"""Oracle (TEST INFRASTRUCTURE): the CPU tail of the reference's predict loop, ref predict.py:160-187.

* ``gaussian_gradient_magnitude``: scipy.ndimage itself (scipy is installed; the reference pins 1.7.0) -- PINNED.
* ``marching_cubes``: ctypes wrapper over ``mc_oracle.c`` (PARITY UNPINNED vs scikit-image, see that file), followed
  by exactly the arithmetic skimage / predict.py apply to the raw vertices: ``vertices * np.r_[spacing]`` in float64,
  the ggm lookup at ``(verts / spacing).astype(np.uint32)`` and the float32 / int32 casts of predict.py:192-199.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmc_oracle.so")


class _McResult(ctypes.Structure):
    _fields_ = [("coords", ctypes.POINTER(ctypes.c_float)), ("faces", ctypes.POINTER(ctypes.c_int32)),
                ("normals", ctypes.POINTER(ctypes.c_float)), ("values", ctypes.POINTER(ctypes.c_float)),
                ("nv", ctypes.c_int64), ("nf", ctypes.c_int64), ("status", ctypes.c_int)]


def _lib():
    if not os.path.exists(_SO):
        import subprocess
        subprocess.check_call(["make", "-C", _HERE])
    lib = ctypes.CDLL(_SO)
    lib.mc_oracle.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                              ctypes.POINTER(_McResult)]
    lib.mc_oracle.restype = ctypes.c_int
    lib.mc_free.argtypes = [ctypes.POINTER(_McResult)]
    return lib


def gaussian_gradient_magnitude(volume: np.ndarray, sigma: float) -> np.ndarray:
    """predict.py:162-163."""
    import scipy.ndimage as ni
    return ni.gaussian_gradient_magnitude(volume, sigma=sigma, mode="nearest")


def marching_cubes(volume: np.ndarray, level: float, spacing=(1.0, 1.0, 1.0), gradient_direction: str = "ascent"):
    """predict.py:172-177 call shape.  Returns (verts float64 [V,3], faces int32 [F,3], normals, values) like skimage
    (verts are float64 because skimage multiplies float32 vertices by a float64 spacing vector).
    Raises ValueError when ``level`` is outside the data range, RuntimeError when no surface is found."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    if vol.ndim != 3 or min(vol.shape) < 2:
        raise ValueError("Input volume should be a 3D numpy array.")
    lib = _lib()
    res = _McResult()
    rc = lib.mc_oracle(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2], float(level),
                       1 if gradient_direction == "ascent" else 0, ctypes.byref(res))
    if rc == -4:
        raise ValueError("Surface level must be within volume data range.")
    try:
        nv, nf = int(res.nv), int(res.nf)
        if nv == 0:
            raise RuntimeError("No surface found at the given iso value.")
        coords = np.ctypeslib.as_array(res.coords, shape=(nv, 3)).copy()
        faces = np.ctypeslib.as_array(res.faces, shape=(nf, 3)).copy() if nf else np.zeros((0, 3), np.int32)
        normals = np.ctypeslib.as_array(res.normals, shape=(nv, 3)).copy()
        values = np.ctypeslib.as_array(res.values, shape=(nv,)).copy()
    finally:
        lib.mc_free(ctypes.byref(res))
    verts = coords * np.r_[spacing]  # float32 * float64 -> float64, as in skimage
    return verts, faces, normals, values


def predict_tail(wnf_volume: np.ndarray, sigma: float = 0.5, level: float = 0.5, gradient_direction: str = "ascent"):
    """predict.py:160-181 + the dtype casts of :192-197 (everything but the surface decoder)."""
    volume_size = wnf_volume.shape[-1]
    ggm = gaussian_gradient_magnitude(wnf_volume, sigma)
    voxel_spacing = 1 / (volume_size - 1)
    verts, faces, normals, values = marching_cubes(wnf_volume, level, (voxel_spacing,) * 3, gradient_direction)
    nn_idx = (verts / voxel_spacing).astype(np.uint32)
    verts_ggm = ggm[nn_idx[:, 0], nn_idx[:, 1], nn_idx[:, 2]]
    return {"verts": verts.astype(np.float32), "faces": faces.astype(np.int32), "normals": normals.astype(np.float32),
            "volume_value": values.astype(np.float32), "volume_gradient_magnitude": verts_ggm.astype(np.float32),
            "ggm": ggm}


def delete_invalid_verts(mc_verts: np.ndarray, mc_faces: np.ndarray, is_vert_on_surface: np.ndarray):
    """ref common/marching_cubes_util.py:38-52, restated line for line (``np.bool`` -> ``bool``: the alias was removed from
    numpy; the reference file cannot be imported here because its module top imports scikit-image)."""
    is_face_valid = np.ones(len(mc_faces), dtype=bool)
    for i in range(3):
        is_face_valid_i = is_vert_on_surface[mc_faces[:, i]]
        is_face_valid = is_face_valid & is_face_valid_i
    raw_valid_faces = mc_faces[is_face_valid]
    raw_valid_vert_idx = np.unique(raw_valid_faces.flatten())
    valid_verts = mc_verts[raw_valid_vert_idx]
    valid_vert_idx = np.arange(len(valid_verts))
    vert_raw_idx_valid_idx_map = np.zeros(len(mc_verts), dtype=mc_faces.dtype)
    vert_raw_idx_valid_idx_map[raw_valid_vert_idx] = valid_vert_idx
    valid_faces = vert_raw_idx_valid_idx_map[raw_valid_faces]
    return valid_verts, valid_faces


def wnf_to_mesh(wnf_volume: np.ndarray, iso_surface_level=0.5, gradient_threshold=0.25, sigma=0.5):
    """ref common/marching_cubes_util.py:5-35 on the oracle's marching cubes (sigma is ignored there too, :7-8)."""
    volume_size = wnf_volume.shape[-1]
    wnf_ggm = gaussian_gradient_magnitude(wnf_volume, 0.5)
    voxel_spacing = 1 / (volume_size - 1)
    mc_verts, mc_faces, _, _ = marching_cubes(wnf_volume, iso_surface_level, (voxel_spacing,) * 3, "ascent")
    idx = (mc_verts / voxel_spacing).astype(np.uint32)
    mc_verts_ggm = wnf_ggm[idx[:, 0], idx[:, 1], idx[:, 2]]
    return delete_invalid_verts(mc_verts, mc_faces, mc_verts_ggm > gradient_threshold)


def chamfer(pred_points: np.ndarray, gt_points: np.ndarray) -> dict:
    """ref eval.py:259-271 (get_chamfer inside compute_chamfer), restated; scipy's cKDTree is installed."""
    from scipy.spatial import cKDTree
    forward_distance, _ = cKDTree(gt_points).query(pred_points, k=1)
    backward_distance, _ = cKDTree(pred_points).query(gt_points, k=1)
    forward_chamfer, backward_chamfer = np.mean(forward_distance), np.mean(backward_distance)
    return {"chamfer_forward": forward_chamfer, "chamfer_backward": backward_chamfer,
            "chamfer_symmetrical": np.mean([forward_chamfer, backward_chamfer])}


def hybrid_chamfer(pred_nocs_points, gt_nocs_points, pred_sim_points, gt_sim_points) -> dict:
    """ref eval.py:381-401 (get_chamfer inside compute_hybrid_chamfer), restated."""
    from scipy.spatial import cKDTree
    _, forward_nn_idx = cKDTree(gt_nocs_points).query(pred_nocs_points, k=1)
    _, backward_nn_idx = cKDTree(pred_nocs_points).query(gt_nocs_points, k=1)
    forward_distance = np.linalg.norm(pred_sim_points - gt_sim_points[forward_nn_idx], axis=1)
    backward_distance = np.linalg.norm(gt_sim_points - pred_sim_points[backward_nn_idx], axis=1)
    forward_chamfer, backward_chamfer = np.mean(forward_distance), np.mean(backward_distance)
    return {"hybrid_chamfer_forward": forward_chamfer, "hybrid_chamfer_backward": backward_chamfer,
            "hybrid_chamfer_symmetrical": np.mean([forward_chamfer, backward_chamfer])}


def optimal_gradient_threshold(gt_mc_verts, gt_mc_is_on_surface, pred_mc_verts, pred_mc_gm, precision_weight=0.85):
    """ref eval.py:58-102 (compute_optimal_gradient_treshold) on plain arrays, restated line for line."""
    from scipy.spatial import cKDTree
    _, nn_vert_idx = cKDTree(gt_mc_verts).query(pred_mc_verts, k=1)
    nn_is_on_surface = gt_mc_is_on_surface[nn_vert_idx]
    sorted_idx = np.argsort(pred_mc_gm)
    sorted_nn_is_on_surface = nn_is_on_surface[sorted_idx]
    false_negative = np.cumsum(sorted_nn_is_on_surface)
    true_positive = np.cumsum(sorted_nn_is_on_surface[::-1])[::-1]
    false_positive = np.cumsum(~sorted_nn_is_on_surface[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = true_positive / (true_positive + false_positive)
        recall = true_positive / (true_positive + false_negative)
    score = precision * precision_weight + recall * (1 - precision_weight)
    if np.any(np.isfinite(score)):
        max_score_threshold = pred_mc_gm[sorted_idx[np.argmax(score)]]
    else:
        max_score_threshold = pred_mc_gm.min()
    return {"optimal_wnf_gradient_threshold": max_score_threshold}


def connected_components(faces: np.ndarray, num_verts: int):
    """ref eval.py:538-541 restated with scipy (igl is not installable offline): components of the vertex adjacency
    graph of the faces, numbered in order of their lowest vertex like igl.connected_components; returns
    (num_cc, cc_idxs, cc_sizes, is_cc_vert) where is_cc_vert marks the component with the most vertices (np.argmax:
    first on ties)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components as cc
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    rows = np.concatenate([f[:, 0], f[:, 1], f[:, 2]])
    cols = np.concatenate([f[:, 1], f[:, 2], f[:, 0]])
    adj = sp.coo_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(num_verts, num_verts))
    num_cc, idx = cc(adj, directed=False)
    # scipy labels components in order of first appearance along the vertex index = order of their lowest vertex
    sizes = np.bincount(idx, minlength=num_cc)
    return num_cc, idx, sizes, idx == int(np.argmax(sizes))


def doublearea(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """igl.doublearea restated: twice the triangle areas, float64."""
    v = np.asarray(verts, dtype=np.float64)
    r = v[faces[:, 1]] - v[faces[:, 0]]
    s = v[faces[:, 2]] - v[faces[:, 0]]
    return np.linalg.norm(np.cross(r, s), axis=1)
