"""Device time of farthest point sampling at the benchmark sizes (32 clouds: 4096 -> 2048 and 2048 -> 512 points) and a checksum of
the selected indices (variants must select the same points).   python tools/fps_bench.py
With the profiling library (GNB_B200_LIBRARY=.../libgarmentnets_b200_prof.so) GNB_FPS_VARIANT=1|2 selects 256x16 / 128x32
threads x points-per-thread for the 4096-point level."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from garmentnets_b200 import ops

dev = torch.device("cuda:0")
B = 32
g = torch.Generator().manual_seed(0)
for n, ratio in ((4096, 0.5), (2048, 0.25)):
    pos = torch.rand(B * n, 3, generator=g).to(dev)
    ptr_h = np.arange(B + 1, dtype=np.int64) * n
    m = ops.fps_counts(ptr_h, ratio)
    out_ptr_h = np.concatenate([[0], np.cumsum(m)]).astype(np.int64)
    ptr, out_ptr = torch.from_numpy(ptr_h).to(dev), torch.from_numpy(out_ptr_h).to(dev)
    start = torch.zeros(B, dtype=torch.int64, device=dev)
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        idx = ops.fps(pos, ptr, out_ptr, n, int(out_ptr_h[-1]), start)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"fps {n} -> {int(m[0])}: {best * 1e3:8.1f} us   checksum {int((idx * torch.arange(1, idx.numel() + 1, device=dev)).sum() % 1000000007)}")
