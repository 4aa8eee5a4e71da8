"""TMA + tcgen05 3x3x3 convolution vs torch CPU (GroupNorm -> Conv3d -> ReLU), all UNet layer shapes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

TOL = 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("B,G,Cin,Cout", [(1, 8, 64, 64), (2, 8, 32, 32), (1, 16, 96, 32), (2, 4, 128, 128),
                                          (1, 8, 192, 64), (4, 4, 256, 96), (1, 32, 128, 128), (1, 16, 32, 64),
                                          (5, 8, 64, 128), (3, 16, 32, 96), (6, 4, 64, 128)])   # batches that are not powers of two
def test_conv_tc_matches_torch(dev, B, G, Cin, Cout):
    from garmentnets_b200 import ops
    assert ops.conv3d_tc_supported(B, G, G, G, Cin, Cout)
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn(B, Cin, G, G, G, generator=g) * 1.5 + 0.3
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    gamma, beta = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.1
    ref = F.relu(F.conv3d(F.group_norm(x, 8, gamma, beta, 1e-5), w, None, padding=1)).permute(0, 2, 3, 4, 1)
    x_cl = ops.to_channels_last(x.to(dev))
    scale, shift = ops.groupnorm_stats(x_cl, 8, 1e-5, gamma.to(dev), beta.to(dev))
    xh, xl = ops.gn_apply_split(x_cl, scale, shift)
    cpad = (Cin + 63) // 64 * 64
    assert xh.shape[-1] == cpad
    gn = (x_cl * scale[:, None, None, None, :] + shift[:, None, None, None, :])
    assert ((xh[..., :Cin].float() + xl[..., :Cin].float()) - gn).abs().max().item() < 2e-6
    assert cpad == Cin or torch.all(xh[..., Cin:] == 0)
    y = ops.conv3d_tc(xh, xl, Cin, ops.conv3d_tc_pack_weights(w.to(dev)), Cout, relu=True)
    err = (y.cpu() - ref).abs().max().item()
    assert err < TOL, err
    y32 = ops.conv3d_k3(x_cl, w.permute(2, 3, 4, 1, 0).reshape(27, Cin, Cout).contiguous().to(dev), scale, shift, relu=True)
    assert (y - y32).abs().max().item() < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("B,D,H,W,Cin,Cout", [(2, 8, 8, 8, 32, 32), (1, 16, 16, 16, 96, 32), (1, 32, 32, 32, 128, 32),
                                              (4, 2, 4, 16, 40, 32), (1, 4, 8, 32, 64, 32), (2, 16, 16, 16, 192, 64),
                                              (1, 16, 16, 16, 32, 64), (2, 8, 8, 8, 64, 64), (3, 16, 16, 16, 96, 32), (5, 8, 8, 8, 64, 64)])
def test_conv_tc_stacked_dx_matches_torch(dev, B, D, H, W, Cin, Cout):
    """Stacked-kw kernel (Cout = 32 / 64): the shift along W applied to the output must reproduce zero padding at both ends of
    every line, for every tile geometry (W = 8 / 16 / 32)."""
    from garmentnets_b200 import ops
    assert ops.conv3d_tc_dx_supported(B, D, H, W, Cin, Cout)
    g = torch.Generator().manual_seed(Cin * 5 + W)
    x = torch.randn(B, Cin, D, H, W, generator=g) * 1.5 + 0.3
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    ref = F.relu(F.conv3d(x.double(), w.double(), None, padding=1)).float().permute(0, 2, 3, 4, 1)
    xh, xl = ops.gn_apply_split(ops.to_channels_last(x.to(dev)), None, None)
    y = ops.conv3d_tc_dx(xh, xl, Cin, ops.conv3d_tc_dx_pack_weights(w.to(dev)), Cout, relu=True)
    err = (y.cpu() - ref).abs().max().item()
    assert err < 2e-5, err
    y3 = ops.conv3d_tc(xh, xl, Cin, ops.conv3d_tc_pack_weights(w.to(dev)), Cout, relu=True)
    assert (y - y3).abs().max().item() < 2e-5


@pytest.mark.gpu
def test_unet_tc_equals_fp32_path(dev):
    from garmentnets_b200 import synthetic
    from garmentnets_b200.components import unet3d
    from oracle import nets as ON
    torch.manual_seed(4)
    net = synthetic.randomize_(unet3d.Abstract3DUNet(128, 128, False, unet3d.DoubleConv, f_maps=32, layer_order="gcr",
                                                     num_groups=8, num_levels=4, is_segmentation=False), 5).eval().to(dev)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 128, 32, 32, 32, generator=g) * (torch.rand(2, 128, 32, 32, 32, generator=g) < 0.1)
    unet3d.USE_TENSOR_CORES = True
    y_tc = net(x.to(dev))
    unet3d.USE_TENSOR_CORES = False
    try:
        y_32 = net(x.to(dev))
    finally:
        unet3d.USE_TENSOR_CORES = True
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    ref = ON.unet3d_forward(sd, "", x).numpy()
    ref64 = ON.unet3d_forward({k: v.double() for k, v in sd.items()}, "", x.double()).numpy()
    scale = max(1.0, float(np.abs(ref64).max()))
    err_tc = np.abs(y_tc.cpu().numpy() - ref64).max()
    err_32 = np.abs(y_32.cpu().numpy() - ref64).max()
    err_cpu = np.abs(ref - ref64).max()
    print(f"UNet max-abs error vs float64: tensor-core {err_tc:.2e}, fp32 FFMA {err_32:.2e}, torch CPU fp32 {err_cpu:.2e}")
    assert err_tc < TOL * scale and err_32 < TOL * scale
    assert np.abs(y_tc.cpu().numpy() - ref).max() < TOL * scale  # and against the fp32 oracle itself


@pytest.mark.gpu
@pytest.mark.parametrize("B,G,Cin,Cout,dx", [(1, 8, 64, 64, False), (2, 4, 128, 128, False), (1, 16, 96, 32, True), (2, 16, 192, 64, True)])
def test_conv_tc_e4m3_cross_terms(dev, B, G, Cin, Cout, dx):
    """Opt-in mode 1 of gnb_conv_tc_set_cross_precision: hi*hi in fp16 plus ONE e4m3 MMA per K-step for lo*w + a*w_lo.
    Two tensor pass-equivalents instead of three; the cross terms keep 4 significant bits, so a layer is good to ~1e-5
    relative rms (1e-4 max) instead of ~1e-6 -- measured 4.4e-4 over the 14 layers of the UNet, which is why the default stays
    mode 0.  The mode must not leak: mode 0 results are bit-identical before and after."""
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(Cin * 7 + Cout)
    x = torch.randn(B, Cin, G, G, G, generator=g) * 1.5 + 0.3
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    ref = F.relu(F.conv3d(x.double(), w.double(), None, padding=1)).permute(0, 2, 3, 4, 1)

    def run():
        xh, xl = ops.gn_apply_split(ops.to_channels_last(x.to(dev)), None, None)
        if dx:
            return ops.conv3d_tc_dx(xh, xl, Cin, ops.conv3d_tc_dx_pack_weights(w.to(dev)), Cout, relu=True)
        return ops.conv3d_tc(xh, xl, Cin, ops.conv3d_tc_pack_weights(w.to(dev)), Cout, relu=True)

    assert ops.conv_tc_cross_precision() == 0
    y0 = run()
    try:
        ops.conv_tc_set_cross_precision(1)
        assert ops.conv_tc_cross_precision() == 1
        y1 = run()
    finally:
        ops.conv_tc_set_cross_precision(0)
    e0 = (y0.cpu().double() - ref).abs()
    e1 = (y1.cpu().double() - ref).abs()
    assert e0.max().item() < 3e-5
    assert e1.max().item() < 2e-4 and e1.pow(2).mean().sqrt().item() < 3e-5, (e1.max().item(), e1.pow(2).mean().sqrt().item())
    assert not torch.equal(y0, y1)          # the mode really changed the arithmetic
    assert torch.equal(run(), y0)           # and mode 0 is back, bit for bit
    with pytest.raises(Exception):
        ops.conv_tc_set_cross_precision(2)
