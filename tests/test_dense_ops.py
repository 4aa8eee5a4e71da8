"""Dense / gridding kernels vs the CPU oracle (torch CPU ATen ops + numpy), called through the C-ABI wrappers."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nets as ON
from oracle import pointops as P

TOL = 1e-4  # north_star: fp32 fields within 1e-4 max-abs


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("R,K,N,relu,bn", [(1000, 6, 64, True, True), (257, 131, 128, True, True), (4096, 259, 256, True, True),
                                           (64, 1280, 256, True, True), (513, 128, 192, False, False), (300, 256, 1, True, True),
                                           (300, 256, 3, True, True), (1, 16, 16, True, False), (129, 137, 137, True, True)])
def test_linear_block(dev, R, K, N, relu, bn):
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    sc = torch.rand(N, generator=g) + 0.5
    sh = torch.randn(N, generator=g)
    ref = F.linear(x, w, b)
    if relu:
        ref = F.relu(ref)
    if bn:
        ref = ref * sc + sh
    got = ops.linear(x.to(dev), w.to(dev), b.to(dev), relu, sc.to(dev) if bn else None, sh.to(dev) if bn else None)
    assert (got.cpu() - ref).abs().max().item() < 2e-5
    # strided input / output and device-side row count
    wide = torch.zeros(R, K + 5)
    wide[:, :K] = x
    out = torch.full((R, N + 3), 7.0, device=dev)
    rows = torch.tensor([max(R - 3, 0)], dtype=torch.int64, device=dev)
    ops.linear(wide.to(dev)[:, :K], w.to(dev), b.to(dev), relu, sc.to(dev) if bn else None, sh.to(dev) if bn else None,
               out=out[:, :N], rows_dev=rows)
    assert (out[:max(R - 3, 0), :N].cpu() - ref[:max(R - 3, 0)]).abs().max().item() < 2e-5 if R > 3 else True
    assert torch.all(out[max(R - 3, 0):, :N] == 7.0) and torch.all(out[:, N:] == 7.0)


@pytest.mark.gpu
def test_mlp_module_matches_oracle(dev):
    from garmentnets_b200 import synthetic
    from garmentnets_b200.components.mlp import MLP
    torch.manual_seed(0)
    m = synthetic.randomize_(MLP([137, 137, 128]), 1).eval()
    x = torch.randn(777, 137)
    ref = ON.mlp(m.state_dict(), "", x)
    got = m.to(dev)(x.to(dev))
    assert (got.cpu() - ref).abs().max().item() < 2e-5
    x3 = torch.randn(2, 50, 137)
    assert (m(x3.to(dev)).cpu() - ON.mlp({k: v.cpu() for k, v in m.state_dict().items()}, "", x3)).abs().max() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("reduce", ["max", "min", "sum", "mean"])
@pytest.mark.parametrize("channels_last", [True, False])
def test_scatter_reduce(dev, reduce, channels_last):
    from garmentnets_b200 import ops
    rng = np.random.default_rng(5)
    N, C, S = 3000, 19, 2000
    feat = rng.normal(size=(N, C)).astype(np.float32)
    feat[::7] = -np.abs(feat[::7])  # make sure all-negative slots exist (empty must stay 0, not -inf)
    index = rng.integers(0, S, N).astype(np.int64)
    index[index % 5 == 0] = 11  # heavy collisions
    ref = P.scatter(feat.T, index, S, reduce)
    got = ops.scatter_reduce(_t(feat, dev).t(), _t(index, dev), S, reduce, channels_last=channels_last)
    assert got.shape == (C, S)
    if reduce in ("max", "min"):
        assert np.array_equal(got.cpu().numpy(), ref)  # bit-exact, order independent
    else:
        # atomic accumulation order is nondeterministic: slot 11 sums ~600 values, so the fp32 rounding error scales with
        # the magnitude of the sums (same max(1, max|ref|) convention as the model-level tests)
        assert np.abs(got.cpu().numpy() - ref).max() < 1e-4 * max(1.0, float(np.abs(ref).max()))
    assert np.all(got.cpu().numpy()[:, np.setdiff1d(np.arange(S), index)] == 0)


@pytest.mark.gpu
def test_scatter_empty_input(dev):
    from garmentnets_b200 import ops
    got = ops.scatter_reduce(torch.zeros((4, 0), device=dev), torch.zeros(0, dtype=torch.int64, device=dev), 10, "max")
    assert got.shape == (4, 10) and torch.all(got == 0)


@pytest.mark.gpu
def test_nocs_head_and_aggregator_features(dev):
    from garmentnets_b200 import ops
    rng = np.random.default_rng(3)
    N, bins, G, B = 2000, 64, 32, 3
    logits = rng.normal(size=(N, bins * 3)).astype(np.float32) * 3
    logits[5, 10 * 3 + 1] = logits[5, 40 * 3 + 1] = 50.0  # exact tie -> first max
    b_ref, conf_ref, nocs_ref = ON.nocs_head(logits, bins)
    b, conf, nocs = ops.nocs_head(_t(logits, dev), bins)
    assert np.array_equal(b.cpu().numpy(), b_ref) and b_ref[5, 1] == 10
    assert np.array_equal(nocs.cpu().numpy(), nocs_ref)
    assert np.abs(conf.cpu().numpy() - conf_ref).max() < 1e-6
    # KATs from the reference's VirtualGrid
    assert nocs_ref.max() <= 1.0 and np.float32(1) * (np.float32(1.0) / np.float32(63.0)) in nocs_ref
    feat = rng.normal(size=(N, 128)).astype(np.float32)
    sim = rng.normal(size=(N, 3)).astype(np.float32)
    batch = np.sort(rng.integers(0, B, N)).astype(np.int64)
    out, flat = ops.aggregator_features(_t(feat, dev), nocs, _t(sim, dev), conf, _t(batch, dev), G)
    idx3 = ON.points_grid_idxs(nocs_ref, G)
    k = np.arange(64)
    assert np.array_equal(np.unique(ON.points_grid_idxs(np.stack([k * np.float32(1 / 63)] * 3, 1).astype(np.float32), 32)[:, 0]),
                          np.arange(32))
    flat_ref = batch * G ** 3 + idx3[:, 0] * G ** 2 + idx3[:, 1] * G + idx3[:, 2]
    assert np.array_equal(flat.cpu().numpy(), flat_ref)
    scales = (torch.ones(3) / (torch.tensor([G] * 3, dtype=torch.float32) - 1))
    local = torch.from_numpy(nocs_ref) - torch.from_numpy(idx3) * scales
    ref = np.concatenate([feat, local.numpy(), sim, conf.cpu().numpy()], 1)
    assert np.array_equal(out.cpu().numpy(), ref)
    # the reference's other configurations (conv_implicit_wnf.py:27-32,76-85): flags off, a box that is not the unit cube
    pts = (rng.random((N, 3)).astype(np.float32) * 2.4 - 1.2)   # some points outside the box: clamped indices
    for lower, upper, ip, ic in (((0, 0, 0), (1, 1, 1), False, False), ((0, 0, 0), (1, 1, 1), True, False),
                                 ((0, 0, 0), (1, 1, 1), False, True), ((-1.0, -0.5, -1.0), (1.0, 1.5, 0.25), True, True)):
        flat_ref, rows_ref = ON.aggregator_rows(feat, pts, sim, conf.cpu().numpy(), batch, G, lower, upper, ip, ic)
        out, flat = ops.aggregator_features(_t(feat, dev), _t(pts, dev), _t(sim, dev), conf, _t(batch, dev), G, lower_corner=lower,
                                            upper_corner=upper, include_point_feature=ip, include_confidence_feature=ic)
        assert out.shape[1] == 128 + 6 * ip + 3 * ic
        assert np.array_equal(flat.cpu().numpy(), flat_ref) and np.array_equal(out.cpu().numpy(), rows_ref.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("flip", [False, True])
def test_trilinear_sample_matches_grid_sample(dev, flip):
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, C, D, H, W, M = 2, 24, 5, 6, 7, 500
    vol = torch.randn(B, C, D, H, W, generator=g)
    q = torch.rand(B, M, 3, generator=g) * 1.2 - 0.1  # includes out-of-range -> border clamp
    q[0, 0] = torch.tensor([0.0, 0.0, 1.0])
    q[0, 1] = torch.tensor([1.0, 1.0, 1.0])
    qn = 2.0 * q - 1.0
    if flip:
        qn = qn.flip(-1)
    ref = F.grid_sample(vol, qn.view(B, M, 1, 1, 3), mode="bilinear", padding_mode="border", align_corners=True)
    ref = ref.view(B, C, M).permute(0, 2, 1).reshape(B * M, C)
    vol_cl = vol.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    got = ops.trilinear_sample(vol_cl, q.to(dev), flip=flip)
    assert (got.cpu() - ref).abs().max().item() < 1e-5
    if not flip:  # axis convention probe: un-flipped (0,0,1) reads volume[..., D=last, 0, 0]
        assert torch.allclose(got[0].cpu(), vol[0, :, D - 1, 0, 0], atol=1e-6)


@pytest.mark.gpu
def test_trilinear_grid_lattice_matches_explicit_points(dev):
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(6)
    vol = torch.randn(2, 4, 4, 4, 16, generator=g).to(dev)
    Q = 9
    gp = ON.grid_points(Q).reshape(1, -1, 3)
    assert torch.equal(gp[0, 1 * 81 + 2 * 9 + 3], torch.tensor([1.0, 2.0, 3.0]) * (torch.tensor(1.0) / torch.tensor(8.0)))
    ref = ops.trilinear_sample(vol[1:2].contiguous(), gp.to(dev))
    got = ops.trilinear_sample_grid(vol, 1, Q, 100, 500)
    assert torch.equal(got, ref[100:600])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,sigma", [((16, 17, 18), 0.5), ((32, 32, 32), 0.5), ((9, 8, 7), 1.0)])
def test_ggm_matches_scipy(dev, shape, sigma):
    import scipy.ndimage as ni
    from garmentnets_b200 import ops
    v = np.random.default_rng(8).normal(size=shape).astype(np.float32)
    ref = ni.gaussian_gradient_magnitude(v, sigma=sigma, mode="nearest")
    got = ops.gaussian_gradient_magnitude(_t(v, dev), sigma).cpu().numpy()
    assert ref.dtype == np.float32
    assert np.abs(got - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max())
    assert (got == ref).mean() > 0.99  # double accumulation in the same order: bit-exact up to rare rounding ties


@pytest.mark.gpu
def test_ggm_impulse_kat(dev):
    from garmentnets_b200 import ops
    v = np.zeros((9, 9, 9), np.float32)
    v[4, 4, 4] = 1.0
    got = ops.gaussian_gradient_magnitude(_t(v, dev), 0.5).cpu().numpy()
    assert abs(got[4, 4, 5] - 0.26344162) < 1e-7 and got[4, 4, 4] == 0 and abs(got[5, 5, 5] - 0.008357321) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("B,G,Cin,Cout,groups", [(2, 8, 32, 64, 8), (1, 6, 96, 32, 8), (1, 4, 128, 128, 8), (2, 5, 4, 16, 1)])
def test_gn_conv_relu(dev, B, G, Cin, Cout, groups):
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, G, G, G + 1, generator=g) * 2 + 0.5
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    gamma, beta = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.1
    ref = F.relu(F.conv3d(F.group_norm(x, groups, gamma, beta, 1e-5), w, None, padding=1))
    x_cl = ops.to_channels_last(x.to(dev))
    assert torch.equal(x_cl.cpu(), x.permute(0, 2, 3, 4, 1).contiguous())
    scale, shift = ops.groupnorm_stats(x_cl, groups, 1e-5, gamma.to(dev), beta.to(dev))
    gn = x_cl * scale[:, None, None, None, :] + shift[:, None, None, None, :]
    assert (gn.cpu() - F.group_norm(x, groups, gamma, beta, 1e-5).permute(0, 2, 3, 4, 1)).abs().max() < 2e-5
    wt = w.permute(2, 3, 4, 1, 0).reshape(27, Cin, Cout).contiguous().to(dev)
    y = ops.conv3d_k3(x_cl, wt, scale, shift, relu=True)
    assert (y.cpu() - ref.permute(0, 2, 3, 4, 1)).abs().max().item() < 5e-5


@pytest.mark.gpu
def test_pool_upsample_concat(dev):
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 12, 8, 6, 4, generator=g)
    x_cl = ops.to_channels_last(x.to(dev))
    p = ops.maxpool3d_2(x_cl)
    assert torch.equal(p.cpu(), F.max_pool3d(x, 2).permute(0, 2, 3, 4, 1))
    skip = torch.randn(2, 5, 8, 6, 4, generator=g)
    up = F.interpolate(F.max_pool3d(x, 2), size=(8, 6, 4), mode="nearest")
    ref = torch.cat((skip, up), 1).permute(0, 2, 3, 4, 1)
    got = ops.upsample_concat(ops.to_channels_last(skip.to(dev)), p)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("Cs,Cx,size,small", [(32, 64, (8, 8, 8), (4, 4, 4)), (64, 128, (4, 6, 2), (2, 3, 1)), (8, 4, (5, 7, 3), (2, 3, 1))])
def test_upsample_concat_vectorised(dev, Cs, Cx, size, small):
    """16-byte path of gnb_upsample_concat (channel counts divisible by 4; ref components/unet3d.py:291,325-330),
    including non-integer nearest-neighbour ratios."""
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(Cs + Cx)
    skip = torch.randn(3, Cs, *size, generator=g)
    low = torch.randn(3, Cx, *small, generator=g)
    ref = torch.cat((skip, F.interpolate(low, size=size, mode="nearest")), 1).permute(0, 2, 3, 4, 1)
    got = ops.upsample_concat(ops.to_channels_last(skip.to(dev)), ops.to_channels_last(low.to(dev)))
    assert torch.equal(got.cpu(), ref)


@pytest.mark.gpu
def test_f16_range_flag(dev):
    """Hardening: values outside +-65504 (or NaN) reaching a saturating fp16 operand split raise the per-device flag --
    through the explicit check the pipeline runs on the decoder grids and through the UNet's normalise-and-split pass."""
    from garmentnets_b200 import ops
    ops.f16_overflow(reset=True)
    x = torch.randn(3, 1000, 7, device=dev) * 100
    ops.f16_range_check(x)
    assert not ops.f16_overflow()
    x[1, 17, 3] = 7.0e4
    ops.f16_range_check(x)
    assert ops.f16_overflow(reset=False) and ops.f16_overflow()      # read without and with reset
    assert not ops.f16_overflow()
    x[1, 17, 3] = float("nan")
    ops.f16_range_check(x.view(-1)[1:])                               # unaligned start: scalar path
    assert ops.f16_overflow()
    # the split pass of the UNet (GroupNorm scale / shift applied, then hi + lo)
    v = torch.randn(2, 4, 4, 4, 32, device=dev)
    sc, sh = torch.ones(2, 32, device=dev), torch.zeros(2, 32, device=dev)
    ops.gn_apply_split(v, sc, sh)
    assert not ops.f16_overflow()
    sc[1, 5] = 1.0e6
    ops.gn_apply_split(v, sc, sh)
    assert ops.f16_overflow()
    # a Linear whose OUTPUT feeds a split (the hoisted first decoder layer): checked by the epilogue that writes it, on the
    # tensor-core path (full TMA tiles and the ragged last tile) and on the fp32 path
    g = torch.Generator().manual_seed(0)
    w, b = (torch.randn(256, 32, generator=g) * 0.2).to(dev), torch.randn(256, generator=g).to(dev)
    for rows in (5000, 300):
        x = torch.randn(rows, 32, generator=g).to(dev)
        y = ops.linear_module(torch.nn.Module(), "t", x, w, b, flag_range=True)
        assert not ops.f16_overflow() and torch.allclose(y, x @ w.t() + b, atol=1e-4)
        x[rows - 3, 7] = 3.0e6        # one output row far outside the fp16 range (last, partial tile)
        ops.linear_module(torch.nn.Module(), "t", x, w, b, flag_range=True)
        assert ops.f16_overflow()
        ops.linear_module(torch.nn.Module(), "t", x, w, b)     # unflagged entry point: no check
        assert not ops.f16_overflow()


@pytest.mark.gpu
def test_copy_to_pinned_host_and_async_flag(dev):
    """The small device -> host path that bypasses the copy engine (marching-cubes totals, range flag): a strided 2-D record copy
    stored by a kernel into pinned memory equals the source after a stream synchronisation; pageable destinations, odd sizes
    and bulk sizes are refused; the asynchronous flag fetch lands in pinned memory and resets the device flag."""
    from garmentnets_b200 import _lib, ops
    from garmentnets_b200._lib import GarmentNetsB200Error
    src = torch.arange(7 * 64, dtype=torch.int32, device=dev).view(7, 64)
    dst = torch.full((7, 16), -1, dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    # columns 8..19 of every row -> columns 2..13 of the pinned rows
    _lib.call("gnb_copy_to_pinned_host", src.data_ptr() + 8 * 4, 64 * 4, dst.data_ptr() + 2 * 4, 16 * 4, 12 * 4, 7, st)
    torch.cuda.synchronize()
    assert torch.equal(dst[:, 2:14], src[:, 8:20].cpu())
    assert (dst[:, :2] == -1).all() and (dst[:, 14:] == -1).all()
    _lib.call("gnb_copy_to_pinned_host", src.data_ptr(), 256, dst.data_ptr(), 64, 0, 7, st)   # empty: no-op
    pageable = torch.zeros((7, 16), dtype=torch.int32)
    with pytest.raises(GarmentNetsB200Error):
        _lib.call("gnb_copy_to_pinned_host", src.data_ptr(), 256, pageable.data_ptr(), 64, 48, 7, st)
    with pytest.raises(GarmentNetsB200Error):
        _lib.call("gnb_copy_to_pinned_host", src.data_ptr(), 256, dst.data_ptr(), 64, 46, 7, st)        # width not a multiple of 4
    with pytest.raises(GarmentNetsB200Error):
        _lib.call("gnb_copy_to_pinned_host", src.data_ptr(), 256, dst.data_ptr(), 64, 2 << 20, 7, st)   # not a small record
    # the flag, stream-ordered
    ops.f16_overflow(reset=True)
    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    x = torch.full((64,), 1.0e5, device=dev)
    ops.f16_range_check(x)
    ops.f16_overflow_async(flag)
    torch.cuda.synchronize()
    assert int(flag[0]) == 1 and not ops.f16_overflow()    # fetched and reset
    ops.f16_overflow_async(flag)
    torch.cuda.synchronize()
    assert int(flag[0]) == 0
    with pytest.raises(GarmentNetsB200Error):
        _lib.call("gnb_f16_overflow_fetch_async", torch.zeros(1, dtype=torch.int32).data_ptr(), 1, st)   # pageable
