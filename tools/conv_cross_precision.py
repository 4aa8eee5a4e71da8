"""fp16 vs e4m3 cross terms of the tensor-core convolutions (gnb_conv_tc_set_cross_precision): error against float64 for single
layers and for the whole 3D-UNet, and device time of the UNet at the benchmark size.   python tools/conv_cross_precision.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from garmentnets_b200 import ops, synthetic
from garmentnets_b200.components import unet3d

dev = torch.device("cuda:0")


def layer(B, G, Cin, Cout, dx):
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn(B, Cin, G, G, G, generator=g) * 1.5 + 0.3
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    ref = F.relu(F.conv3d(x.double(), w.double(), None, padding=1)).permute(0, 2, 3, 4, 1)
    out = []
    for mode in (0, 1):
        ops.conv_tc_set_cross_precision(mode)
        xh, xl = ops.gn_apply_split(ops.to_channels_last(x.to(dev)), None, None)
        if dx:
            y = ops.conv3d_tc_dx(xh, xl, Cin, ops.conv3d_tc_dx_pack_weights(w.to(dev)), Cout, relu=True)
        else:
            y = ops.conv3d_tc(xh, xl, Cin, ops.conv3d_tc_pack_weights(w.to(dev)), Cout, relu=True)
        e = (y.cpu().double() - ref).abs()
        out.append(f"mode {mode}: max {e.max().item():.2e} rms {e.pow(2).mean().sqrt().item():.2e}")
    print(f"conv{'_dx' if dx else '   '} B={B} G={G} {Cin}->{Cout}: " + " | ".join(out) + f"  (output rms {ref.pow(2).mean().sqrt().item():.2f})")


for args in ((1, 8, 64, 64, False), (2, 4, 128, 128, False), (1, 16, 96, 32, True), (2, 16, 192, 64, True), (1, 16, 32, 64, True)):
    layer(*args)

# whole UNet against the float64 oracle
from oracle import nets as ON
torch.manual_seed(4)
net = synthetic.randomize_(unet3d.Abstract3DUNet(128, 128, False, unet3d.DoubleConv, f_maps=32, layer_order="gcr", num_groups=8,
                                                 num_levels=4, is_segmentation=False), 5).eval().to(dev).requires_grad_(False)
g = torch.Generator().manual_seed(6)
x = torch.randn(2, 128, 32, 32, 32, generator=g) * (torch.rand(2, 128, 32, 32, 32, generator=g) < 0.1)
sd = {k: v.cpu().double() for k, v in net.state_dict().items()}
ref64 = ON.unet3d_forward(sd, "", x.double()).numpy()
for mode in (0, 1):
    ops.conv_tc_set_cross_precision(mode)
    y = net(x.to(dev)).cpu().numpy()
    e = np.abs(y - ref64)
    print(f"UNet mode {mode}: max abs err {e.max():.2e}  rms {np.sqrt((e ** 2).mean()):.2e}   (output max {np.abs(ref64).max():.2f}, rms {np.sqrt((ref64 ** 2).mean()):.2f})")

# device time at the benchmark size
xb = torch.randn(32, 128, 32, 32, 32, device=dev) * (torch.rand(32, 128, 32, 32, 32, device=dev) < 0.1)
xb = ops.to_channels_last(xb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in (0, 1, 0, 1):
    ops.conv_tc_set_cross_precision(mode)
    net.forward_ndhwc(xb, apply_final=False)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        net.forward_ndhwc(xb, apply_final=False)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"UNet batch 32 mode {mode}: {best:.3f} ms")
ops.conv_tc_set_cross_precision(0)
