"""Golden fixture for the chamfer metrics produced by the REFERENCE's own code.

eval.py cannot be imported here (hydra, zarr, igl ... are missing), and the two `get_chamfer` helpers are nested inside
compute_chamfer (:259-271) and compute_hybrid_chamfer (:381-401); their sources are cut out of the file with `ast` and
executed unmodified against scipy's cKDTree (installed).

    python oracle/make_golden_chamfer.py      # rewrites tests/golden/chamfer.npz
"""
import ast
import os

import numpy as np
from scipy.spatial import ckdtree

REF_FILE = "/root/reference/eval.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "chamfer.npz")


def reference_functions():
    src = open(REF_FILE).read()
    tree = ast.parse(src)
    out = {}
    for outer in ("compute_chamfer", "compute_hybrid_chamfer"):
        fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == outer)
        inner = next(n for n in ast.walk(fn) if isinstance(n, ast.FunctionDef) and n.name == "get_chamfer")
        code = ast.get_source_segment(src, inner)
        lines = code.split("\n")
        indent = len(lines[1]) - len(lines[1].lstrip()) - 4      # nested def: remove the outer indentation
        code = "\n".join([lines[0]] + [l[indent:] if l.strip() else l for l in lines[1:]])
        ns = {"np": np, "ckdtree": ckdtree}
        exec(compile(code, REF_FILE, "exec"), ns)
        out[outer] = ns["get_chamfer"]
    return out


def main():
    fns = reference_functions()
    rng = np.random.default_rng(7)
    out = {}
    for i, (n_pred, n_gt) in enumerate([(500, 700), (10000, 10000), (3, 1)]):
        pred_nocs = rng.random((n_pred, 3)).astype(np.float32)
        gt_nocs = (rng.random((n_gt, 3)) * 0.9 + 0.05).astype(np.float32)
        pred_sim = (pred_nocs * [0.7, 0.2, 1.1] + rng.normal(scale=0.01, size=(n_pred, 3))).astype(np.float32)
        gt_sim = (gt_nocs * [0.7, 0.2, 1.1]).astype(np.float32)
        c = fns["compute_chamfer"](pred_nocs, gt_nocs)
        h = fns["compute_hybrid_chamfer"](pred_nocs, gt_nocs, pred_sim, gt_sim)
        out.update({f"pred_nocs{i}": pred_nocs, f"gt_nocs{i}": gt_nocs, f"pred_sim{i}": pred_sim, f"gt_sim{i}": gt_sim,
                    f"chamfer{i}": np.float64(c["chamfer_symmetrical"]),
                    f"hybrid{i}": np.array([h["hybrid_chamfer_forward"], h["hybrid_chamfer_backward"],
                                            h["hybrid_chamfer_symmetrical"]], dtype=np.float64)})
    out["cases"] = np.int64(3)
    np.savez_compressed(OUT, **out)
    print("written", OUT, {k: v for k, v in out.items() if k.startswith(("chamfer", "hybrid"))})


if __name__ == "__main__":
    main()
