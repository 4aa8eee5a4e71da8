"""Golden fixture for row a9: runs the REFERENCE's own ``batch_to_volume`` (/root/reference/components/gridding.py:8-42,
imported unmodified) on CPU tensors.  Its only third-party call, ``torch_scatter.scatter``, is served by the oracle's
numpy restatement (``oracle.pointops.scatter``; torch_scatter itself is not installable offline), so what this pins is
the reference function's own arithmetic: voxel index = clamp(trunc(pos * G), 0, G-1), flat index layout, dtype cast,
reshape / permute.

    python oracle/make_golden_batch_to_volume.py      # rewrites tests/golden/batch_to_volume.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pointops as P  # noqa: E402


def _scatter(src, index, dim=-1, dim_size=None, reduce="sum", out=None):
    assert dim == -1 and out is None
    return torch.from_numpy(P.scatter(src.numpy(), index.numpy(), int(dim_size), reduce))


def main():
    stub = types.ModuleType("torch_scatter")
    stub.scatter = _scatter
    sys.modules["torch_scatter"] = stub
    sys.path.insert(0, REF)
    for name in [m for m in sys.modules if m == "components" or m.startswith("components.")]:
        del sys.modules[name]
    from components import gridding
    sys.path.remove(REF)
    assert gridding.__file__.startswith(REF)

    class Bag:
        pass

    g = torch.Generator().manual_seed(77)
    B, N, C, G = 3, 500, 5, 6
    batch = Bag()
    batch.pos = torch.rand(N, 3, generator=g) * 1.3 - 0.15          # some points outside [0,1): exercised clamp
    batch.pos[:7] = torch.tensor([0.0, 1.0, 0.999999, 1.0 / 6, 0.5, -0.0, 2.0])[:, None]
    batch.x = torch.randn(N, C, generator=g)
    batch.batch = torch.sort(torch.randint(0, B, (N,), generator=g)).values
    batch.num_graphs = B
    out = {"pos": batch.pos.numpy(), "x": batch.x.numpy(), "batch": batch.batch.numpy(), "G": np.int64(G), "B": np.int64(B)}
    for reduce in ("mean", "max", "sum", "min"):
        vol = gridding.batch_to_volume(batch, G, reduce=reduce)
        assert tuple(vol.shape) == (B, C, G, G, G)
        out["vol_" + reduce] = vol.contiguous().numpy()
    path = os.path.join(ROOT, "tests", "golden", "batch_to_volume.npz")
    np.savez_compressed(path, **out)
    print("written", path)


if __name__ == "__main__":
    main()
