"""gnb_linear_tc (tcgen05 Linear -> ReLU -> BatchNorm block, ref components/mlp.py:9-20) against torch on the CPU.
The fp16 hi/lo split keeps fp32-level accuracy: the bound below is the same 2e-5 the fp32 kernel is held to."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(5000, 6, 64, True, True),       # SA1 layer 1 (K below one MMA step)
          (3000, 64, 64, True, True), (3000, 64, 128, True, True),
          (2500, 131, 128, True, True),    # odd K / odd row stride
          (2000, 128, 256, True, True),
          (1500, 1280, 256, True, True),   # FP3: 20 K-chunks, streamed weights
          (1300, 259, 256, True, True), (1100, 256, 512, True, True), (1030, 512, 1024, True, True),   # column blocks
          (4096, 137, 137, True, True),    # aggregator: padded columns, scalar tail stores
          (2048, 32, 256, False, False),   # hoisted first decoder layer
          (1024, 128, 192, False, False), (129, 256, 256, True, False), (1, 16, 16, True, True), (127, 384, 256, True, True)]


@pytest.mark.parametrize("R,K,N,relu,bn", SHAPES)
def test_linear_tc_block(dev, R, K, N, relu, bn):
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    sc = torch.rand(N, generator=g) + 0.5
    sh = torch.randn(N, generator=g)
    ref = F.linear(x.double(), w.double(), b.double())
    if relu:
        ref = F.relu(ref)
    if bn:
        ref = ref * sc.double() + sh.double()
    ref = ref.float()
    pk = ops.pack_linear_tc(w.to(dev), b.to(dev), sc.to(dev) if bn else None, sh.to(dev) if bn else None)
    got = ops.linear_tc(x.to(dev), pk, relu)
    assert got.shape == (R, N)
    assert (got.cpu() - ref).abs().max().item() < 2e-5
    # strided input / output and device-side row count: rows beyond it and columns beyond N stay untouched
    wide = torch.zeros(R, K + 5)
    wide[:, :K] = x
    out = torch.full((R, N + 3), 7.0, device=dev)
    nrows = max(R - 3, 0)
    rows = torch.tensor([nrows], dtype=torch.int64, device=dev)
    ops.linear_tc(wide.to(dev)[:, :K], pk, relu, out=out[:, :N], rows_dev=rows)
    if nrows:
        assert (out[:nrows, :N].cpu() - ref[:nrows]).abs().max().item() < 2e-5
    assert torch.all(out[nrows:, :N] == 7.0) and torch.all(out[:, N:] == 7.0)


def test_linear_tc_large_magnitudes(dev):
    """Activations of O(1e2) and weights of O(1e-3 .. 1e1) (what the synthetic BatchNorm statistics produce)."""
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4096, 256, generator=g) * 100.0
    w = torch.randn(256, 256, generator=g) * torch.logspace(-3, 1, 256)[:, None] / 16
    ref = F.linear(x.double(), w.double()).float()
    got = ops.linear_tc(x.to(dev), ops.pack_linear_tc(w.to(dev)), False)
    err = (got.cpu() - ref).abs().max().item()
    assert err < 1e-5 * ref.abs().max().item(), err


def test_mlp_block_routes_to_tensor_cores(dev):
    """components.mlp.MLP in eval mode: >= LINEAR_TC_MIN_ROWS rows go through gnb_linear_tc and agree with the fp32 kernel."""
    from garmentnets_b200 import _lib, ops
    from garmentnets_b200.components.mlp import MLP
    torch.manual_seed(0)
    m = MLP([131, 128, 128, 256]).to(dev).eval()
    for blk in m:
        blk[2].running_mean.normal_(0, 0.1)
        blk[2].running_var.uniform_(0.5, 1.5)
    x = torch.randn(3000, 131, device=dev)
    calls = []
    orig = _lib.call
    _lib.call = lambda name, *a: (calls.append(name), orig(name, *a))[1]
    try:
        y_tc = m(x)
        ops.USE_LINEAR_TC = False
        y_fp32 = m(x)
    finally:
        ops.USE_LINEAR_TC = True
        _lib.call = orig
    assert calls.count("gnb_linear_tc") == 3 and calls.count("gnb_linear") == 3
    assert (y_tc - y_fp32).abs().max().item() < 2e-5 * max(1.0, y_fp32.abs().max().item())
    # a changed BatchNorm statistic invalidates the cached pack
    m[0][2].running_mean.add_(1.0)
    y2 = m(x)
    assert (y2 - y_tc).abs().max().item() > 1e-3


@pytest.mark.parametrize("nseg,K,N,maxlen", [(700, 64, 128, 65), (300, 128, 256, 65), (5000, 64, 64, 9), (3, 131, 128, 2000),
                                             (1200, 259, 1024, 40)])
def test_linear_tc_segmax_equals_linear_then_segment_max(dev, nseg, K, N, maxlen):
    """Aggregation fused into the epilogue (gnb_linear_tc_segmax) == gnb_linear_tc followed by gnb_segment_max, bit for
    bit (max is order independent and both paths produce the same per-row values); segments straddle warps, tiles and
    CTAs, rows beyond the device-side row count are ignored, all-negative segments keep their sign."""
    from garmentnets_b200 import ops
    g = torch.Generator().manual_seed(nseg + K + N)
    lens = torch.randint(1, maxlen + 1, (nseg,), generator=g)
    offs = torch.zeros(nseg + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens, 0)
    E = int(offs[-1])
    rows = E + 37                                   # worst-case allocation, like PointConv.forward_grouped
    x = torch.randn(rows, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) - 1.0
    sc = torch.rand(N, generator=g) + 0.5
    sh = torch.randn(N, generator=g) - 2.0           # many all-negative segments
    pk = ops.pack_linear_tc(w.to(dev), b.to(dev), sc.to(dev), sh.to(dev))
    xd, od = x.to(dev), offs.to(dev)
    total = od[nseg:]
    full = ops.linear_tc(xd, pk, True, rows_dev=total)
    want = ops.segment_max(full, od)
    got = ops.linear_tc_segmax(xd, pk, ops.segment_ids(od, rows), nseg, relu=True, rows_dev=total)
    assert got.shape == (nseg, N) and got.dtype == torch.float32
    assert torch.equal(got, want)
    ref = torch.relu(x[:E].double() @ w.double().t() + b.double()) * sc.double() + sh.double()
    ref = torch.stack([ref[offs[i]:offs[i + 1]].max(0).values for i in range(nseg)]).float()
    assert (got.cpu() - ref).abs().max().item() < 2e-5
