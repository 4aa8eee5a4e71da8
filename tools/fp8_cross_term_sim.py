"""CPU simulation of the tensor-core operand splits on the synthetic decoders (no GPU, no CUDA library): end-to-end error of
  * the shipped 3-pass fp16 split            hi*hi + lo*hi + hi*lo,
  * fp16 hi*hi + fp8 (e4m3) cross terms      [lo * 2^10 | a] x [W * 2^(c-10) | W_lo * 2^c]  (2 tensor pass-equivalents),
  * a single fp16 pass,
against float64, through Linear2 -> BN2 -> Linear3 -> BN3 of ImplicitWNFDecoder (findings table).
    python tools/fp8_cross_term_sim.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from garmentnets_b200 import synthetic
from garmentnets_b200.pipeline import ImplicitWNFDecoder
torch.manual_seed(0)
for cout, seed in ((1, 1), (3, 2), (1, 5)):
    dec = synthetic.randomize_(ImplicitWNFDecoder(nn_channels=(128, 256, 256, cout)), seed).eval().double()
    R = 200000
    # interpolated features: convex combination of 8 randn vectors
    w = torch.rand(R, 8, dtype=torch.float64); w = w / w.sum(1, keepdim=True)
    feats = torch.randn(R, 8, 128, dtype=torch.float64)
    x = (w[..., None] * feats).sum(1)
    l1, bn1 = dec.mlp[0][0], dec.mlp[0][2]
    l2, bn2 = dec.mlp[1][0], dec.mlp[1][2]
    l3, bn3 = dec.mlp[2][0], dec.mlp[2][2]
    def bn(m, v): return (v - m.running_mean) / torch.sqrt(m.running_var + m.eps) * m.weight + m.bias
    h1 = torch.relu(l1(x)).float()           # A operand (fp32 in the kernel)
    # BN1 folded into W2
    s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps); t1 = bn1.bias - bn1.running_mean * s1
    W2 = (l2.weight * s1[None, :]).float(); b2 = (l2.bias + l2.weight @ t1)
    def tail(h2pre):
        h2 = bn(bn2, torch.relu(h2pre + b2))
        return bn(bn3, torch.relu(l3(h2)))
    ref = tail(h1.double() @ W2.double().T)
    def rz16(x):
        hf = x.half().float(); over = hf > x
        return torch.where(over, torch.nextafter(hf.half(), torch.zeros_like(hf).half()).float(), hf)
    f8 = lambda v: v.to(torch.float8_e4m3fn).float()
    c = 2.0 ** math.floor(math.log2(60000 / W2.abs().max().item()))
    Ahi = rz16(h1); Alo = h1 - Ahi
    Whi = (W2 * c).half().float(); Wlo = W2 * c - Whi
    main = Ahi.double() @ Whi.double().T
    s = 2.0 ** 10
    p3 = main + Alo.half().double() @ Whi.double().T + Ahi.double() @ Wlo.half().double().T
    p8 = main + f8(Alo * s).double() @ f8(W2 * c / s).double().T + f8(h1).double() @ f8(Wlo).double().T
    Ahn = h1.half().float(); Aln = h1 - Ahn
    p8n = Ahn.double() @ Whi.double().T + f8(Aln * s).double() @ f8(W2 * c / s).double().T + f8(h1).double() @ f8(Wlo).double().T
    print(f"cout {cout} seed {seed}: c=2^{math.log2(c):.0f} max h1 {h1.max().item():.2f} out rms {ref.pow(2).mean().sqrt().item():.3f} max {ref.abs().max().item():.3f}")
    for name, p in (("3-pass f16", p3), ("f16+2xfp8 (rz hi)", p8), ("f16+2xfp8 (rn hi)", p8n), ("1-pass", main)):
        e = (tail(p / c) - ref).abs()
        print(f"   {name:20s} max abs err {e.max().item():.3e}  rms {e.pow(2).mean().sqrt().item():.3e}  frac>1e-4 {(e > 1e-4).double().mean().item():.2e}")
