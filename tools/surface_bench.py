"""Device time of the fused query-mode decoder (surface / warp-field decode) at the benchmark size: 32 samples x ~188k mesh
vertices each, scan-ordered like marching-cubes output.   python tools/surface_bench.py [verts_per_sample]
GNB_TC_DBG=1|2|4 knocks out the gather / the Linear1 math / the epilogue math (bottleneck analysis)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from garmentnets_b200 import synthetic
from garmentnets_b200.pipeline import ImplicitWNFDecoder

dev = torch.device("cuda:0")
V = int(sys.argv[1]) if len(sys.argv) > 1 else 188000
B = 32
torch.manual_seed(0)
dec = synthetic.randomize_(ImplicitWNFDecoder(nn_channels=(128, 256, 256, 3)), 1).eval().requires_grad_(False).to(dev)
final_conv = torch.nn.Conv3d(32, 128, 1).to(dev).requires_grad_(False)
x32 = torch.randn(B, 32, 32, 32, 32, device=dev)
# scan-ordered points on a wavy sheet: consecutive vertices are spatial neighbours, as after marching cubes
g = torch.Generator().manual_seed(1)
q = torch.rand(B, V, 3, generator=g)
q = q[:, torch.argsort((q[0, :, 0] * 128).floor() * 16384 + (q[0, :, 1] * 128).floor() * 128 + (q[0, :, 2] * 128).floor())]
q_all = q.reshape(-1, 3).contiguous().to(dev)
qptr = np.arange(B + 1, dtype=np.int64) * V
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, n=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"{name:52s} {best:8.3f} ms   {B * V / best / 1e3:8.1f} Mquery/s")
    return best


timeit("fused query mode (32-ch gather + Linear1 in kernel)", lambda: dec.forward_fused_ragged(x32, final_conv, q_all, qptr))
u = dec.hoisted_folded(x32, final_conv)
timeit("hoisted query mode (256-ch gather)", lambda: dec.forward_hoisted_ragged(u, q_all, qptr))
