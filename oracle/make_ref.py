"""Oracle (TEST INFRASTRUCTURE): stage the reference's UNCHANGED Python sources for the GPU box.

``/root/reference`` exists only in the build container; the ``-m gpu`` tests run on a box that does not have it.
This recipe copies, byte for byte, the reference modules whose forward the drop-in boundary must serve

    networks/conv_implicit_wnf.py, networks/pointnet2_nocs.py          (L4: the pipeline that must run unchanged)
    common/torch_util.py, common/visualization_util.py, common/rendering_util.py   (their module-load imports)
    components/pointnet2.py -> ref_components/pointnet2.py             (the reference's own SA/FP modules, run against
                                                                        our torch_geometric.nn stand-ins)

into ``oracle/_ref/`` -- git-ignored (never part of the history), not gpurun-ignored (it travels with the snapshot like
``oracle/_build``).  ``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present.  Nothing under
``garmentnets_b200/`` reads ``oracle/_ref``; only ``tests/test_reference_forward.py`` does (as the thing under test is
OUR components; the reference files are the unmodified caller).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

FILES = {
    "networks/conv_implicit_wnf.py": "networks/conv_implicit_wnf.py",
    "networks/pointnet2_nocs.py": "networks/pointnet2_nocs.py",
    "common/torch_util.py": "common/torch_util.py",
    "common/visualization_util.py": "common/visualization_util.py",
    "common/rendering_util.py": "common/rendering_util.py",
    "components/pointnet2.py": "ref_components/pointnet2.py",
}


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "networks")):
        print("make_ref: /root/reference not mounted; keeping whatever oracle/_ref holds")
        return 0
    manifest = {}
    for src, dst in FILES.items():
        s, d = os.path.join(REF, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(s, "rb") as f:
            manifest[dst] = {"source": src, "sha256": hashlib.sha256(f.read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(f"make_ref: staged {len(FILES)} unmodified reference files under {OUT}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
