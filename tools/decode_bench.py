"""Device time of the dense lattice decode kernels at the benchmark size (B x 128^3 queries), CUDA events, best of 5.
    python tools/decode_bench.py [B]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from garmentnets_b200 import ops, synthetic
from garmentnets_b200.pipeline import ImplicitWNFDecoder

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
dec = synthetic.randomize_(ImplicitWNFDecoder(nn_channels=(128, 256, 256, 1)), 1).eval().requires_grad_(False).to(dev)
u = torch.randn(B, 32, 32, 32, 256, device=dev)
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
    flop = B * 128 ** 3 * (2 * 256 * 256 + 512)
    print(f"{name:44s} {best:8.3f} ms   {flop / best / 1e9:7.1f} TFLOP/s algorithmic  ({3 * flop / best / 1e9:7.1f} executed)")
    return best


out1 = ops.decode_tc(*dec._tc_args(), U=u, Q=128, bn1=dec.mlp[0][2].folded_affine())
out2 = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
print("max |pair - gen1| =", (out1 - out2).abs().max().item())
timeit("decode_tc lattice (gen 1)", lambda: ops.decode_tc(*dec._tc_args(), U=u, Q=128, bn1=dec.mlp[0][2].folded_affine()))
from garmentnets_b200 import _lib
_lib.call("gnb_decode_lattice_set_mode", 2)
out3 = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
print("max |8 producer warps - 16 producer warps| =", (out3 - out2).abs().max().item())
timeit("decode_lattice (one CTA / SM, 16 producer warps)", lambda: ops.decode_lattice(*dec._lattice_args(), U=u, Q=128))
_lib.call("gnb_decode_lattice_set_mode", 1)
timeit("decode_lattice (cta_group::2 pairs)", lambda: ops.decode_lattice(*dec._lattice_args(), U=u, Q=128))
_lib.call("gnb_decode_lattice_set_mode", 0)
timeit("decode_lattice (one CTA / SM, 8 producer warps, default)", lambda: ops.decode_lattice(*dec._lattice_args(), U=u, Q=128))

# ---- profiling build only (GNB_B200_LIBRARY=garmentnets_b200/lib/libgarmentnets_b200_prof.so): knock-outs and per-role
# wait-time attribution of both kernel forms
lib = _lib.load()
has_clocks = hasattr(lib, "gnb_prof_decode_lattice_read")   # absent in PROFILE_NO_CLOCKS=1 builds (knock-outs only)
if os.environ.get("GNB_B200_LIBRARY"):
    import ctypes
    import numpy as np
    names = ["mma:a_full", "mma:w2", "mma:d_empty", "mma:total", "prod:a_empty", "prod:total", "prod:rows", "epi:d_full",
             "epi:total", "load:b_empty", "load:total", "prod:xblend", "prod:issue", "prod:yblend", "prod:fence"]

    def prof(label):
        if not has_clocks:
            return
        buf = np.zeros(1024 * 16, np.uint64)
        lib.gnb_prof_decode_lattice_read(ctypes.c_void_p(buf.ctypes.data), ctypes.c_int32(buf.size))
        t = buf.reshape(1024, 16)[:148].astype(np.float64)
        lead, peer = t[0::2], t[1::2]
        fmt = lambda a: " ".join(f"{n}={a[:, i].mean() / 1e6:6.2f}" for i, n in enumerate(names))
        print(f"   [{label}] Mcycles/CTA even: {fmt(lead)}")
        print(f"   [{label}] Mcycles/CTA odd : {fmt(peer)}")

    for mode, mname in ((0, "8 warps"),):
        _lib.call("gnb_decode_lattice_set_mode", mode)
        for dbg in (0, 4, 8, 16, 24, 28):
            os.environ["GNB_DL2_DBG"] = str(dbg)
            timeit(f"{mname} dbg={dbg}", lambda: ops.decode_lattice(*dec._lattice_args(), U=u, Q=128), n=3)
            prof(f"{mname} dbg={dbg}")
    os.environ["GNB_DL2_DBG"] = "0"
_lib.call("gnb_decode_lattice_set_mode", 0)
