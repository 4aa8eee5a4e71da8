"""Golden fixture for the gradient-threshold search (ref eval.py:58-102) produced by the
REFERENCE's own `compute_optimal_gradient_treshold`: its source is cut out of eval.py with `ast` and executed unmodified
(scipy cKDTree is installed); the zarr groups it reads are replaced by nested dicts of numpy arrays (`group[key][:]`
works on both).

    python oracle/make_golden_threshold.py      # rewrites tests/golden/gradient_threshold.npz
"""
import ast
import os

import numpy as np
from scipy.spatial import ckdtree

REF_FILE = "/root/reference/eval.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gradient_threshold.npz")


def reference_function():
    src = open(REF_FILE).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "compute_optimal_gradient_treshold")
    ns = {"np": np, "ckdtree": ckdtree}
    exec(compile(ast.get_source_segment(src, node), REF_FILE, "exec"), ns)
    return ns["compute_optimal_gradient_treshold"]


def main():
    fn = reference_function()
    rng = np.random.default_rng(11)
    out = {}
    for i, (n_gt, n_pred) in enumerate([(400, 600), (6000, 8000), (50, 30)]):
        gt_verts = rng.random((n_gt, 3)).astype(np.float32)
        gt_on = gt_verts[:, 2] > 0.35                                  # the "open" part of the ground-truth surface
        pred_verts = (rng.random((n_pred, 3))).astype(np.float32)
        # gradient magnitude correlated with being on the surface, plus noise
        pred_gm = (0.2 + 0.5 * (pred_verts[:, 2] > 0.35) + rng.normal(scale=0.2, size=n_pred)).astype(np.float32)
        groups = {"s": {"gt_marching_cubes_mesh": {"marching_cube_verts": gt_verts, "is_vertex_on_surface": gt_on},
                        "marching_cubes_mesh": {"verts": pred_verts, "volume_gradient_magnitude": pred_gm}}}
        for w in (0.85, 0.5):
            r = fn("s", groups, precision_weight=w)
            out[f"thr{i}_{int(w * 100)}"] = np.float64(r["optimal_wnf_gradient_threshold"])
        out.update({f"gt_verts{i}": gt_verts, f"gt_on{i}": gt_on, f"pred_verts{i}": pred_verts, f"pred_gm{i}": pred_gm})
    out["cases"] = np.int64(3)
    np.savez_compressed(OUT, **out)
    print("written", OUT, {k: float(v) for k, v in out.items() if k.startswith("thr")})


if __name__ == "__main__":
    main()
