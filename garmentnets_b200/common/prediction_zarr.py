"""``prediction.zarr`` writer (ref predict.py:75-84 the store, :192-279 the per-sample groups).

The reference writes every per-sample array through ``zarr`` 2.8 into a ``DirectoryStore`` with
``Blosc(zstd, 6, BITSHUFFLE)`` and one chunk per array (``chunks=data.shape``); ``eval.py`` reads it back with
``zarr.open``.  zarr / numcodecs are not installable here, so this module writes the **zarr v2 on-disk format** itself
(the published spec: ``.zgroup`` / ``.zattrs`` / ``.zarray`` JSON documents, C-order little-endian chunk files named by
their chunk index, dimension separator ``.``) with a codec every stock zarr installation decodes:

* ``compressor="zlib"``  -> ``{"id": "zlib", "level": L}`` (numcodecs.Zlib; Python's ``zlib`` here), the default;
* ``compressor=None``    -> raw chunks (``"compressor": null``).

Blosc-zstd itself cannot be produced without the Blosc library; the group / array names, dtypes, shapes, chunking and
attributes are the reference's, so ``eval.py``'s ``zarr.open(path)['samples'][key]['marching_cubes_mesh']['verts'][:]``
reads the same arrays.  PARITY NOTE: the layout is checked against the v2 spec by ``read_array`` / ``tests/test_zarr.py``
(an independent reader in this file), not against the zarr package (absent offline).

Group layout per sample (ref predict.py:192-279)::

    samples/<group_key>/marching_cubes_mesh/{verts f32[V,3], faces i32[F,3], normals f32[V,3], volume_value f32[V],
                                             volume_gradient_magnitude f32[V], warp_field f32[V,3]}
    samples/<group_key>/point_cloud/{pred_nocs, pred_nocs_confidence, pred_nocs_logits, input_points, input_rgb u8, gt_nocs}
    samples/<group_key>/misc/{pred_nocs_grip_point, pred_global_nocs_grip_point, pred_global_confidence, global_feature, ...}
"""
from __future__ import annotations

import json
import os
import zlib
from typing import Dict, Mapping, Optional

import numpy as np

MC_KEYS = ("verts", "faces", "normals", "volume_value", "volume_gradient_magnitude", "warp_field")
_MC_DTYPES = {"verts": np.float32, "faces": np.int32, "normals": np.float32, "volume_value": np.float32,
              "volume_gradient_magnitude": np.float32, "warp_field": np.float32}


def _write_json(path: str, doc) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(doc, f, indent=4, sort_keys=True)
    os.replace(tmp, path)


def _dtype_str(dt: np.dtype) -> str:
    """zarr v2 dtype string: numpy's array-protocol typestr ('<f4', '|u1', '|b1', ...)."""
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise ValueError("big-endian arrays are not written")
    return dt.str


class ZarrGroup:
    """A zarr v2 group in a directory store (the subset of ``zarr.Group`` predict.py uses)."""

    def __init__(self, path: str, compressor: Optional[str] = "zlib", level: int = 1, overwrite: bool = False):
        if compressor not in (None, "zlib"):
            raise ValueError("compressor must be None or 'zlib' (Blosc is not available offline)")
        self.path, self.compressor, self.level = path, compressor, int(level)
        os.makedirs(path, exist_ok=True)
        meta = os.path.join(path, ".zgroup")
        if overwrite or not os.path.exists(meta):
            _write_json(meta, {"zarr_format": 2})

    def require_group(self, name: str, overwrite: bool = False) -> "ZarrGroup":
        return ZarrGroup(os.path.join(self.path, name), self.compressor, self.level, overwrite=overwrite)

    def put_attrs(self, attrs: Mapping) -> None:
        _write_json(os.path.join(self.path, ".zattrs"), dict(attrs))

    def array(self, name: str, data, chunks=None, overwrite: bool = True) -> None:
        """One chunk per array (``chunks=data.shape``, as predict.py:213-216 does), C order."""
        a = np.ascontiguousarray(np.asarray(data))
        if chunks is not None and tuple(chunks) != a.shape:
            raise ValueError("only chunks == data.shape is supported (what the reference writes)")
        d = os.path.join(self.path, name)
        if os.path.exists(os.path.join(d, ".zarray")) and not overwrite:
            raise ValueError(f"array {name!r} exists")
        os.makedirs(d, exist_ok=True)
        # a zero-length dimension has no chunks at all; a 0-d array has the single chunk "0"
        chunk_shape = [max(int(n), 1) for n in a.shape]
        comp = None if self.compressor is None else {"id": "zlib", "level": self.level}
        _write_json(os.path.join(d, ".zarray"), {
            "chunks": chunk_shape, "compressor": comp, "dtype": _dtype_str(a.dtype), "fill_value": 0 if a.dtype.kind != "b" else False,
            "filters": None, "order": "C", "shape": [int(n) for n in a.shape], "zarr_format": 2})
        if a.size == 0:
            return
        key = ".".join("0" for _ in a.shape) or "0"
        raw = a.tobytes()
        if self.compressor == "zlib":
            raw = zlib.compress(raw, self.level)
        tmp = os.path.join(d, key + ".tmp")
        with open(tmp, "wb") as f:
            f.write(raw)
        os.replace(tmp, os.path.join(d, key))


def read_array(path: str) -> np.ndarray:
    """Independent reader of one single-chunk zarr v2 array written by any implementation (tests; spec section 'Arrays')."""
    meta = json.load(open(os.path.join(path, ".zarray")))
    assert meta["zarr_format"] == 2 and meta["order"] == "C" and not meta["filters"]
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dt = np.dtype(meta["dtype"])
    if int(np.prod(shape, dtype=np.int64)) == 0:
        return np.zeros(shape, dt)
    assert all(c >= s for c, s in zip(chunks, shape)), "single-chunk arrays only"
    key = ".".join("0" for _ in shape) or "0"
    raw = open(os.path.join(path, key), "rb").read()
    comp = meta["compressor"]
    if comp is not None:
        assert comp["id"] == "zlib"
        raw = zlib.decompress(raw)
    full = np.frombuffer(raw, dt).reshape(chunks)
    return full[tuple(slice(0, s) for s in shape)].copy()


class PredictionZarrWriter:
    """``prediction.zarr`` of one prediction run (ref predict.py:75-84): root group with the ``subset`` attribute and a
    ``samples`` group; ``write_sample`` stores what one iteration of the reference loop stores for its sample."""

    def __init__(self, path: str, subset: str = "test", compressor: Optional[str] = "zlib", level: int = 1):
        self.root = ZarrGroup(path, compressor, level)
        self.root.put_attrs({"subset": subset})
        self.samples = self.root.require_group("samples")

    def write_sample(self, group_key: str, mc_data: Mapping[str, np.ndarray], pc_data: Optional[Mapping[str, np.ndarray]] = None,
                     misc_data: Optional[Mapping[str, np.ndarray]] = None, attrs: Optional[Mapping] = None) -> ZarrGroup:
        g = self.samples.require_group(group_key)
        if attrs:
            g.put_attrs(attrs)
        mc = g.require_group("marching_cubes_mesh")
        for key, data in mc_data.items():   # predict.py:192-200 casts: float32 everywhere, int32 faces
            a = np.asarray(data)
            if key in _MC_DTYPES:
                a = a.astype(_MC_DTYPES[key], copy=False)
            mc.array(key, a, chunks=a.shape)
        if pc_data:
            pc = g.require_group("point_cloud")
            for key, data in pc_data.items():
                a = np.asarray(data)
                pc.array(key, a, chunks=a.shape)
        if misc_data:
            misc = g.require_group("misc")
            for key, data in misc_data.items():
                a = np.asarray(data)
                misc.array(key, a, chunks=a.shape)
        return g

    def write_batch(self, group_keys, results, point_outputs: Optional[Dict[str, np.ndarray]] = None, num_points=None,
                    inputs: Optional[Dict[str, np.ndarray]] = None) -> None:
        """Everything ``HostPredictor.result`` / ``point_outputs`` returned for one batch: sample ``b`` goes to
        ``samples/<group_keys[b]>``.  ``num_points`` (per-sample point counts) slices the flat per-point arrays;
        ``inputs`` may hold the flat ``pos`` / ``x`` of the batch (stored as input_points / input_rgb u8 like
        predict.py:222-225)."""
        off = np.concatenate([[0], np.cumsum(num_points)]) if num_points is not None else None
        for b, (key, r) in enumerate(zip(group_keys, results)):
            pc = None
            if point_outputs is not None and off is not None:
                sl = slice(int(off[b]), int(off[b + 1]))
                pc = {"pred_nocs": point_outputs["pred_nocs"][sl]}
                if "pred_confidence" in point_outputs:
                    pc["pred_nocs_confidence"] = point_outputs["pred_confidence"][sl]
                if inputs is not None:
                    pc["input_points"] = np.asarray(inputs["pos"])[sl]
                    pc["input_rgb"] = (np.asarray(inputs["x"])[sl] * 255).astype(np.uint8)
            self.write_sample(key, r, pc, attrs={"batch_idx": b})
