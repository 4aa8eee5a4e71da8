"""Golden fixture for the mesh clean-up produced by the REFERENCE's own function.

/root/reference/common/marching_cubes_util.py cannot be imported here (its module top imports scikit-image's removed
`marching_cubes_lewiner`), so the source of `delete_invalid_verts` (:38-52) is cut out of the file with `ast` and
executed unmodified; the only shim is `np.bool = bool` (the alias was removed from numpy >= 1.24).

    python oracle/make_golden_mesh_cleanup.py      # rewrites tests/golden/mesh_cleanup.npz
"""
import ast
import os

import numpy as np

REF_FILE = "/root/reference/common/marching_cubes_util.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mesh_cleanup.npz")


def reference_function():
    src = open(REF_FILE).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "delete_invalid_verts")
    code = ast.get_source_segment(src, node)
    if not hasattr(np, "bool"):
        np.bool = bool  # noqa: the 2021 reference uses the removed alias
    ns = {"np": np}
    exec(compile(code, REF_FILE, "exec"), ns)
    return ns["delete_invalid_verts"]


def main():
    fn = reference_function()
    rng = np.random.default_rng(2021)
    out = {}
    for i, (V, F, p) in enumerate([(40, 90, 0.7), (3000, 7000, 0.5), (500, 1500, 0.95), (64, 100, 0.0)]):
        verts = rng.normal(size=(V, 3)).astype(np.float32)
        faces = rng.integers(0, V, size=(F, 3)).astype(np.int32)
        on = rng.random(V) < p
        v, f = fn(verts, faces, on)
        out.update({f"verts{i}": verts, f"faces{i}": faces, f"on{i}": on, f"valid_verts{i}": v,
                    f"valid_faces{i}": np.asarray(f).reshape(-1, 3).astype(np.int32)})
    out["cases"] = np.int64(4)
    np.savez_compressed(OUT, **out)
    print("written", OUT, {k: v.shape for k, v in out.items() if k.startswith("valid")})


if __name__ == "__main__":
    main()
