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


from garmentnets_b200 import _lib
timeit("query decoder, Linear1 + Linear2 on tcgen05 (default)", lambda: dec.forward_fused_ragged(x32, final_conv, q_all, qptr))
_lib.call("gnb_decode_query_set_mode", 1)
timeit("query decoder, Linear1 with FFMA2 in the producers", lambda: dec.forward_fused_ragged(x32, final_conv, q_all, qptr))
_lib.call("gnb_decode_query_set_mode", 0)
u = dec.hoisted_folded(x32, final_conv)
timeit("hoisted query mode (256-ch gather)", lambda: dec.forward_hoisted_ragged(u, q_all, qptr))

# profiling build only (GNB_B200_LIBRARY=.../libgarmentnets_b200_prof.so): per-role wait-time attribution of the query decoder
lib = _lib.load()
if hasattr(lib, "gnb_prof_decode_query_read"):
    import ctypes
    names = ["mma:a1_full", "mma:acc1_empty", "mma:d_empty", "mma:a2_full", "mma:b_full", "mma:total", "gather:a1_empty", "gather:total",
             "mid:acc1_full", "mid:a2_empty", "mid:total", "epi:d_full", "epi:total", "load:b_empty"]
    timeit("query decoder (profiling build)", lambda: dec.forward_fused_ragged(x32, final_conv, q_all, qptr), n=2)
    buf = np.zeros(1024 * 16, np.uint64)
    lib.gnb_prof_decode_query_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_int32(buf.size))
    t = buf.reshape(1024, 16)[:148].astype(np.float64)
    print("   Mcycles/CTA: " + " ".join(f"{n}={t[:, i].mean() / 1e6:6.2f}" for i, n in enumerate(names)))
