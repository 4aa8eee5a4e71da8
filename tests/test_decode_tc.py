"""Tensor-core (tcgen05) decoder tail vs fp32 references: the bf16 hi/lo split must keep the result within the 1e-4
parity bound; lattice mode is checked against the oracle at the full BASELINE size (128^3)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from garmentnets_b200 import synthetic
from oracle import nets as ON

TOL = 1e-4


def _decoder(dev, cout, seed):
    from garmentnets_b200.pipeline import ImplicitWNFDecoder
    torch.manual_seed(seed)
    dec = synthetic.randomize_(ImplicitWNFDecoder(nn_channels=(128, 256, 256, cout)), seed + 1).eval().requires_grad_(False)
    return dec.to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("R,cout", [(1, 1), (127, 3), (128, 1), (1000, 3), (40000, 2), (148 * 128 * 3 + 5, 1)])
def test_row_mode_matches_fp32(dev, R, cout):
    from garmentnets_b200 import ops
    dec = _decoder(dev, cout, 10 + cout)
    g = torch.Generator().manual_seed(R)
    X = torch.randn(R, 256, generator=g) * 1.5
    sd = {k: v.cpu() for k, v in dec.state_dict().items()}
    # reference: blocks 2 and 3 of the MLP in fp64 (what fp32 approximates), and in fp32 (the oracle)
    ref32 = X
    ref64 = X.double()
    for l in (1, 2):
        w, b = sd[f"mlp.{l}.0.weight"], sd[f"mlp.{l}.0.bias"]
        ref32 = F.batch_norm(F.relu(F.linear(ref32, w, b)), sd[f"mlp.{l}.2.running_mean"], sd[f"mlp.{l}.2.running_var"],
                             sd[f"mlp.{l}.2.weight"], sd[f"mlp.{l}.2.bias"], False, 0.0, 1e-5)
        ref64 = F.batch_norm(F.relu(F.linear(ref64, w.double(), b.double())), sd[f"mlp.{l}.2.running_mean"].double(),
                             sd[f"mlp.{l}.2.running_var"].double(), sd[f"mlp.{l}.2.weight"].double(),
                             sd[f"mlp.{l}.2.bias"].double(), False, 0.0, 1e-5)
    got = ops.decode_tc(*dec._tc_args(), X=X.to(dev)).cpu()
    assert got.shape == (R, cout)
    err32 = (got - ref32).abs().max().item()
    err64 = (got.double() - ref64).abs().max().item()
    assert err32 < TOL and err64 < TOL, (err32, err64)
    # strided input rows (an [R, 300] buffer)
    wide = torch.zeros(R, 300, device=dev)
    wide[:, :256] = X.to(dev)
    got2 = ops.decode_tc(*dec._tc_args(), X=wide[:, :256]).cpu()
    assert torch.equal(got, got2)


@pytest.mark.gpu
def test_tail_dispatch_equals_fp32_path(dev):
    dec = _decoder(dev, 3, 5)
    g = torch.Generator().manual_seed(0)
    fg = torch.randn(2, 128, 6, 6, 6, generator=g).to(dev)
    q = torch.rand(2, 333, 3, generator=g).to(dev)
    dec.use_tensor_cores = True
    y_tc = dec(fg, q)
    dec.use_tensor_cores = False
    y_32 = dec(fg, q)
    assert (y_tc - y_32).abs().max().item() < TOL
    ref = ON.implicit_decoder({k: v.cpu() for k, v in dec.state_dict().items()}, "", fg.cpu(), q.cpu())
    assert (y_tc.cpu() - ref).abs().max().item() < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("pair_kernel", [True, False])
def test_lattice_mode_full_size_vs_oracle(dev, pair_kernel):
    """128^3 lattice (BASELINE size), one volume: fused tcgen05 kernels (pair-tile generation and the first one) vs the
    oracle's chunked grid_sample + MLP."""
    from garmentnets_b200 import ops
    from garmentnets_b200.pipeline import ConvImplicitWNFPipeline
    torch.manual_seed(1)
    model = synthetic.randomize_(ConvImplicitWNFPipeline.from_hparams(synthetic.HPARAMS), 2).eval().requires_grad_(False).to(dev)
    g = torch.Generator().manual_seed(3)
    fvol = (torch.randn(2, 128, 32, 32, 32, generator=g) * 0.7).to(dev)
    model.volume_decoder.use_tensor_cores = True
    model.volume_decoder.use_pair_lattice = pair_kernel
    wnf_tc = model.dense_decode(fvol, 128)
    model.volume_decoder.use_tensor_cores = False
    wnf_32 = model.dense_decode(fvol[1:2], 128)
    assert wnf_tc.shape == (2, 128, 128, 128)
    assert (wnf_tc[1] - wnf_32[0]).abs().max().item() < TOL
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref = ON.dense_decode(sd, "volume_decoder.", fvol[:1].cpu().numpy(), 128, 64)
    err = (wnf_tc[0].cpu() - ref).abs().max().item()
    assert err < TOL, err


@pytest.mark.gpu
@pytest.mark.parametrize("counts,cout", [([333, 0, 1, 700], 3), ([5], 1), ([128, 128], 2), ([0, 0, 9000], 3)])
def test_query_mode_matches_oracle_and_row_mode(dev, counts, cout):
    """Ragged query mode (gather fused into the tcgen05 producer) == trilinear_sample + row mode, and within 1e-4 of the
    oracle's grid_sample + MLP; samples without rows and tiles that straddle two samples included."""
    from garmentnets_b200 import ops
    dec = _decoder(dev, cout, 20 + cout)
    B = len(counts)
    g = torch.Generator().manual_seed(sum(counts) + cout)
    fg = (torch.randn(B, 128, 6, 6, 6, generator=g) * 0.8).to(dev)
    qs = [torch.rand(n, 3, generator=g) for n in counts]
    # a few points exactly on / outside the border (padding_mode='border' clamps them)
    if counts[-1] >= 4:
        qs[-1][:4] = torch.tensor([[0, 0, 0], [1, 1, 1], [1.2, -0.1, 0.5], [0.5, 1.0, 0.0]])
    q_all = torch.cat(qs).to(dev)
    qptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    u = dec.hoisted(ops.to_channels_last(fg))
    got = dec.forward_hoisted_ragged(u, q_all, qptr).cpu()
    assert got.shape == (sum(counts), cout)
    sd = {k: v.cpu() for k, v in dec.state_dict().items()}
    for b, n in enumerate(counts):
        if n == 0:
            continue
        sl = slice(int(qptr[b]), int(qptr[b + 1]))
        ref = ON.implicit_decoder(sd, "", fg[b:b + 1].cpu(), qs[b].view(1, -1, 3))[0]
        assert (got[sl] - ref).abs().max().item() < TOL
        rows = dec.forward_hoisted(u[b:b + 1], qs[b].view(1, -1, 3).to(dev)).view(n, cout).cpu()
        assert (got[sl] - rows).abs().max().item() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("counts,cout,G", [([333, 0, 1, 700], 3, 6), ([5], 1, 4), ([128, 128], 2, 8), ([0, 0, 9000], 3, 32),
                                           ([20000, 30000], 3, 32)])
def test_fused_query_mode_matches_oracle(dev, counts, cout, G):
    """FUSED query mode (32-channel gather + Linear1 applied per query in the producers + tcgen05 tail) against the oracle
    run the reference's way: final_conv (1x1x1) -> grid_sample -> MLP (ref components/unet3d.py:467,
    networks/conv_implicit_wnf.py:128-149), and against the hoisted 256-channel query kernel."""
    from garmentnets_b200 import ops
    dec = _decoder(dev, cout, 30 + cout)
    B = len(counts)
    g = torch.Generator().manual_seed(sum(counts) + cout)
    x32 = (torch.randn(B, G, G, G, 32, generator=g) * 0.8).to(dev)           # channels-last last UNet level
    final_conv = torch.nn.Conv3d(32, 128, 1).to(dev).requires_grad_(False)
    qs = [torch.rand(n, 3, generator=g) for n in counts]
    if counts[-1] >= 4:
        qs[-1][:4] = torch.tensor([[0, 0, 0], [1, 1, 1], [1.2, -0.1, 0.5], [0.5, 1.0, 0.0]])
    q_all = torch.cat(qs).to(dev)
    qptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    assert dec.fused_query_ready(x32)
    got = dec.forward_fused_ragged(x32, final_conv, q_all, qptr).cpu()
    assert got.shape == (sum(counts), cout)
    hoisted = dec.forward_hoisted_ragged(dec.hoisted_folded(x32, final_conv), q_all, qptr).cpu()
    assert (got - hoisted).abs().max().item() < 5e-5
    sd = {k: v.cpu() for k, v in dec.state_dict().items()}
    fv = F.conv3d(x32.permute(0, 4, 1, 2, 3).cpu(), final_conv.weight.cpu(), final_conv.bias.cpu())   # [B,128,G,G,G]
    for b, n in enumerate(counts):
        if n == 0:
            continue
        sl = slice(int(qptr[b]), int(qptr[b + 1]))
        ref = ON.implicit_decoder(sd, "", fv[b:b + 1], qs[b].view(1, -1, 3))[0]
        assert (got[sl] - ref).abs().max().item() < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("counts,cout", [([4000, 1, 0, 2500], 3), ([777], 1)])
def test_tensor_core_linear1_equals_ffma_linear1(dev, counts, cout):
    """gnb_decode_tc_query_fused with Linear1 on tcgen05 (decode_query.cu, the default when BatchNorm1 is folded) against the
    first-generation kernel that applies Linear1 per query with FFMA2: same math, fp16 hi/lo split products vs fp32 FMAs."""
    from garmentnets_b200 import _lib
    dec = _decoder(dev, cout, 50 + cout)
    B, G = len(counts), 16
    g = torch.Generator().manual_seed(11 + cout)
    x32 = (torch.randn(B, G, G, G, 32, generator=g) * 0.8).to(dev)
    final_conv = torch.nn.Conv3d(32, 128, 1).to(dev).requires_grad_(False)
    q_all = torch.rand(sum(counts), 3, generator=g).to(dev)
    qptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    try:
        _lib.call("gnb_decode_query_set_mode", 1)
        ffma = dec.forward_fused_ragged(x32, final_conv, q_all, qptr)
        _lib.call("gnb_decode_query_set_mode", 0)
        tc = dec.forward_fused_ragged(x32, final_conv, q_all, qptr)
        again = dec.forward_fused_ragged(x32, final_conv, q_all, qptr)
    finally:
        _lib.call("gnb_decode_query_set_mode", 0)
    torch.cuda.synchronize()
    assert torch.equal(tc, again)
    assert (tc - ffma).abs().max().item() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("G,B", [(32, 3), (8, 1), (16, 2), (5, 2), (31, 1)])
def test_pair_lattice_equals_first_generation(dev, G, B):
    """Pair-tile lattice kernel vs decode_tc lattice mode on other grid sizes (G < 32: idle producer groups; G > 32:
    two D-cells per group; odd G) and an odd batch (odd number of work items per CTA)."""
    from garmentnets_b200 import ops
    dec = _decoder(dev, 1, 31)
    g = torch.Generator().manual_seed(G * 7 + B)
    u = (torch.randn(B, G, G, G, 256, generator=g) * 0.9).to(dev)
    if G % 2 == 0:
        ref = ops.decode_tc(*dec._tc_args(), U=u, Q=128, bn1=dec.mlp[0][2].folded_affine())
    else:  # the first-generation kernel needs an even G: reference = explicit sampling + row mode
        ref = torch.stack([dec.forward_lattice(u, b, 128, 0, 128 ** 3) for b in range(B)])
    got = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
    assert got.shape == (B, 128 ** 3, 1)
    assert (got - ref.view_as(got)).abs().max().item() < 2e-5


@pytest.mark.gpu
def test_two_streams_do_not_race_on_the_constant_operand_banks(dev):
    """The decoders / Linear blocks keep their per-column epilogue operands in __constant__ banks refreshed in front of
    every launch.  Two streams launching the same family with DIFFERENT operands (volume vs surface decoder, two Linear
    blocks) must still each see their own: the library orders such launches on the device (ConstBankGuard)."""
    from garmentnets_b200 import ops
    dec_a, dec_b = _decoder(dev, 1, 31), _decoder(dev, 3, 32)
    g = torch.Generator().manual_seed(1)
    Xa = (torch.randn(60000, 256, generator=g) * 1.5).to(dev)
    Xb = (torch.randn(50000, 256, generator=g) * 1.5).to(dev)
    wa = ops.pack_linear_tc(dec_a.mlp[1][0].weight, dec_a.mlp[1][0].bias, *dec_a.mlp[1][2].folded_affine())
    wb = ops.pack_linear_tc(dec_b.mlp[1][0].weight, dec_b.mlp[1][0].bias, *dec_b.mlp[1][2].folded_affine())
    args_a, args_b = dec_a._tc_args(), dec_b._tc_args()
    ref_a, ref_b = ops.decode_tc(*args_a, X=Xa), ops.decode_tc(*args_b, X=Xb)
    lin_a, lin_b = ops.linear_tc(Xa, wa, relu=True), ops.linear_tc(Xb, wb, relu=True)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(sa):
            ya = ops.decode_tc(*args_a, X=Xa)
            la = ops.linear_tc(Xa, wa, relu=True)
        with torch.cuda.stream(sb):
            yb = ops.decode_tc(*args_b, X=Xb)
            lb = ops.linear_tc(Xb, wb, relu=True)
        torch.cuda.synchronize()
        assert torch.equal(ya, ref_a) and torch.equal(yb, ref_b)
        assert torch.equal(la, lin_a) and torch.equal(lb, lin_b)


@pytest.mark.gpu
@pytest.mark.parametrize("G,B,cout", [(32, 2, 1), (16, 1, 3), (7, 3, 2)])
def test_cta_pair_lattice_kernel_is_bit_identical_to_single_cta(dev, G, B, cout):
    """gnb_decode_lattice as clusters of two CTAs (tcgen05.mma.cta_group::2, each SM holds half of every W2 piece) against the
    one-CTA-per-SM form: same products, same accumulation order -> identical bits."""
    from garmentnets_b200 import _lib, ops
    dec = _decoder(dev, cout, 40 + cout)
    g = torch.Generator().manual_seed(G + B)
    u = (torch.randn(B, G, G, G, 256, generator=g) * 0.9).to(dev)
    try:
        _lib.call("gnb_decode_lattice_set_mode", 0)
        single = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
        _lib.call("gnb_decode_lattice_set_mode", 2)      # 16 producer warps x 2 channels per lane
        eight = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
        _lib.call("gnb_decode_lattice_set_mode", 1)
        pair = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
        again = ops.decode_lattice(*dec._lattice_args(), U=u, Q=128)
    finally:
        _lib.call("gnb_decode_lattice_set_mode", 0)   # back to the default (one CTA per SM)
    torch.cuda.synchronize()
    assert torch.equal(pair, single) and torch.equal(eight, single)
    assert torch.equal(pair, again)
