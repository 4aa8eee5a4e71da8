"""Golden fixture for the area-weighted sampling of row f3: executes the REFERENCE's own ``mesh_sample_barycentric`` and
``barycentric_interpolation`` (/root/reference/common/geometry_util.py:160-223), cut out of the module with ``ast`` (the
module's top imports ``igl``, which is not installable offline) and run unmodified; the one ``igl`` call inside,
``igl.doublearea``, is served by the oracle's numpy restatement.

    python oracle/make_golden_mesh_sampling.py     # rewrites tests/golden/mesh_sampling.npz
"""
import ast
import os
import sys
import types

import numpy as np

REF = "/root/reference/common/geometry_util.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postproc  # noqa: E402


def _cut(names):
    src = open(REF).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names)
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"np": np, "Optional": __import__("typing").Optional, "Tuple": __import__("typing").Tuple,
          "igl": types.SimpleNamespace(doublearea=postproc.doublearea)}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def main():
    ns = _cut(["mesh_sample_barycentric", "barycentric_interpolation"])
    rng = np.random.default_rng(5)
    # a bumpy open sheet: 30 x 40 grid of vertices, two triangles per quad, very uneven face areas
    gx, gy = np.meshgrid(np.linspace(0, 1, 30) ** 2, np.linspace(0, 1, 40), indexing="ij")
    verts = np.stack([gx, gy, 0.1 * np.sin(7 * gx) * np.cos(5 * gy)], -1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(30 * 40).reshape(30, 40)
    q = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 4)
    faces = np.concatenate([q[:, [0, 1, 2]], q[:, [0, 2, 3]]]).astype(np.int32)
    field = rng.normal(size=(len(verts), 3)).astype(np.float32)
    out = {"verts": verts, "faces": faces, "field": field}
    for seed, n in ((0, 10000), (7, 257)):
        bc, fi = ns["mesh_sample_barycentric"](verts, faces, num_samples=n, seed=seed)
        pts = ns["barycentric_interpolation"](bc, verts, faces[fi])
        fld = ns["barycentric_interpolation"](bc, field, faces[fi])
        out.update({f"bc_{seed}": bc, f"fi_{seed}": fi, f"pts_{seed}": pts, f"fld_{seed}": fld})
    # float64 vertices (eval.py normalises the vertices in float64 before sampling)
    v64 = verts.astype(np.float64) * 1.7 - 0.3
    bc, fi = ns["mesh_sample_barycentric"](v64, faces, num_samples=500, seed=3)
    out.update({"v64": v64, "bc_64": bc, "fi_64": fi, "pts_64": ns["barycentric_interpolation"](bc, v64, faces[fi])})
    path = os.path.join(ROOT, "tests", "golden", "mesh_sampling.npz")
    np.savez_compressed(path, **out)
    print("written", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
