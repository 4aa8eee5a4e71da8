"""Where does the pair-lattice kernel's time go?  Runs it with parts disabled (GNB_DL2_DBG bit mask)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from garmentnets_b200 import ops, synthetic
from garmentnets_b200.pipeline import ImplicitWNFDecoder
dev = torch.device("cuda:0")
B = 8
dec = synthetic.randomize_(ImplicitWNFDecoder(nn_channels=(128, 256, 256, 1)), 1).eval().requires_grad_(False).to(dev)
u = torch.randn(B, 32, 32, 32, 256, device=dev)
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    os.environ["GNB_DL2_DBG"] = str(dbg)
    ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.decode_lattice(*dec._lattice_args(), U=u, Q=128); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    tiles = B * 128 * 128
    print(f"dbg={dbg} (producers {'off' if dbg & 1 else 'on '}, W2 copies {'off' if dbg & 2 else 'on '}, epilogue {'off' if dbg & 4 else 'on '}): "
          f"{best:7.3f} ms  {best * 1e3 * 148 / tiles:6.2f} us/tile")
