"""Point-set ops: CUDA kernels vs the CPU oracle (bit-exact indices), oracle vs scipy cKDTree (sets)."""
import numpy as np
import pytest
import torch

from garmentnets_b200 import synthetic
from oracle import pointops as P


def _cloud_batch(B, n, seed=0, ragged=False):
    d = synthetic.make_batch(B, n, "Tshirt", seed)
    if ragged:  # drop a different number of points per cloud
        keep = np.ones(B * n, dtype=bool)
        for b in range(B):
            keep[b * n + n - 17 * (b + 1): (b + 1) * n] = False
        d = {k: v[keep] for k, v in d.items()}
    return d


# ------------------------------------------------------------------------------------------------ CPU: oracle sanity
def test_oracle_ball_query_matches_kdtree_sets():
    from scipy.spatial import cKDTree
    d = _cloud_batch(1, 1024)
    pos = d["pos"]
    ptr = np.array([0, 1024])
    idx = P.fps(pos, ptr, 0.5)
    assert len(idx) == 512 and len(set(idx.tolist())) == 512 and idx[0] == 0
    nbr, cnt = P.ball_query(pos, pos[idx], ptr, np.array([0, 512]), 0.05, 1 << 20 if False else 1024)
    tree = cKDTree(pos.astype(np.float64))
    mism = 0
    for q in range(512):
        ref = set(tree.query_ball_point(pos[idx[q]].astype(np.float64), 0.05))
        got = set(nbr[q, :cnt[q]].tolist())
        mism += len(ref ^ got)  # only boundary points (|d - r| ~ 1e-8) may differ between fp32 and fp64
    assert mism <= 2


def test_oracle_knn_matches_kdtree():
    from scipy.spatial import cKDTree
    d = _cloud_batch(1, 512)
    pos = d["pos"]
    src = pos[::4]
    idx, d2 = P.knn(src, pos, np.array([0, len(src)]), np.array([0, len(pos)]), 3)
    _, ref = cKDTree(src.astype(np.float64)).query(pos.astype(np.float64), k=3)
    assert (np.sort(idx, 1) == np.sort(ref, 1)).mean() > 0.999
    assert np.all(np.diff(d2, axis=1) >= 0)


def test_oracle_fps_counts_and_ragged():
    ptr = np.array([0, 100, 100, 357])
    assert P.fps_counts(ptr, 0.25).tolist() == [25, 0, 65]
    rng = np.random.default_rng(0)
    pos = rng.normal(size=(357, 3)).astype(np.float32)
    idx = P.fps(pos, ptr, 0.25, start=np.array([3, 0, 9]))
    assert idx[0] == 3 and idx[25] == 100 + 9 and len(idx) == 90
    assert np.all(idx[:25] < 100) and np.all(idx[25:] >= 100)


def test_oracle_pointconv_edges_self_loop_rule():
    nbr = np.array([[0, 2, -1], [0, 2, 3], [3, -1, -1]])
    cnt = np.array([2, 3, 1], dtype=np.int32)
    offs, src = P.pointconv_edges(nbr, cnt)
    # centroid 0: {2} + self 0 ; centroid 1: {0,2,3} + self 1 ; centroid 2: {3} + self 2
    assert offs.tolist() == [0, 2, 6, 8]
    assert src.tolist() == [2, 0, 0, 2, 3, 1, 3, 2]


def test_oracle_scatter_empty_is_zero():
    src = np.array([[-1.0, -5.0, 2.0]], dtype=np.float32)
    out = P.scatter(src, np.array([1, 1, 3]), 5, "max")
    assert out.tolist() == [[0.0, -1.0, 0.0, 2.0, 0.0]]
    out = P.scatter(src, np.array([1, 1, 3]), 5, "mean")
    assert out.tolist() == [[0.0, -3.0, 0.0, 2.0, 0.0]]


# ------------------------------------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("B,n,ragged", [(2, 1024, False), (3, 700, True), (1, 4096, False), (1, 5000, False), (1, 9000, False)])
def test_fps_bit_exact(dev, B, n, ragged):
    from garmentnets_b200 import ops
    from garmentnets_b200.components.pointnet2 import CloudIndex
    d = _cloud_batch(B, n, seed=1, ragged=ragged)
    ptr = P.batch_to_ptr(d["batch"], B)
    start = np.array([(7 * b + 3) % (ptr[b + 1] - ptr[b]) for b in range(B)], dtype=np.int64)
    ref = P.fps(d["pos"], ptr, 0.5, start)
    index = CloudIndex(torch.from_numpy(ptr).to(dev), ptr)
    sub = index.subsample(0.5)
    got = ops.fps(torch.from_numpy(d["pos"]).to(dev), index.ptr, sub.ptr, index.max_n, sub.total,
                  torch.from_numpy(start).to(dev))
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,r,ragged", [(2, 2048, 0.05, False), (3, 900, 0.1, True), (1, 4096, 0.05, False)])
def test_ball_query_bit_exact(dev, B, n, r, ragged):
    from garmentnets_b200 import ops
    d = _cloud_batch(B, n, seed=2, ragged=ragged)
    ptr = P.batch_to_ptr(d["batch"], B)
    idx = P.fps(d["pos"], ptr, 0.5)
    ptr_y = np.concatenate([[0], np.cumsum(P.fps_counts(ptr, 0.5))])
    nbr_ref, cnt_ref = P.ball_query(d["pos"], d["pos"][idx], ptr, ptr_y, r, 64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nbr, cnt = ops.ball_query(t(d["pos"]), t(d["pos"][idx]), t(ptr), t(ptr_y), r, 64)
    assert np.array_equal(cnt.cpu().numpy(), cnt_ref)
    assert np.array_equal(nbr.cpu().numpy(), nbr_ref)
    assert cnt_ref.max() == 64  # truncation is exercised
    row, col = ops.radius_pairs(nbr, cnt)
    row_ref, col_ref = P.radius_pairs(nbr_ref, cnt_ref)
    assert np.array_equal(row.cpu().numpy(), row_ref) and np.array_equal(col.cpu().numpy(), col_ref)
    # PointConv edge set
    offs_ref, src_ref = P.pointconv_edges(nbr_ref, cnt_ref)
    offs = ops.pointconv_edges(nbr, cnt)
    assert np.array_equal(offs.cpu().numpy(), offs_ref)
    feat = np.random.default_rng(0).normal(size=(len(d["pos"]), 5)).astype(np.float32)
    edge = torch.zeros((len(idx) * 65, 8), device=dev)
    ops.pointconv_gather(t(feat), t(d["pos"]), t(d["pos"][idx]), nbr, cnt, offs, edge)
    edge_ref = P.pointconv_edge_features(feat, d["pos"], d["pos"][idx], offs_ref, src_ref)
    assert np.array_equal(edge[:len(edge_ref)].cpu().numpy(), edge_ref)
    seg = ops.segment_max(edge[:len(edge_ref)], offs)
    assert np.array_equal(seg.cpu().numpy(), P.segment_max(edge_ref, offs_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 3, 5, 12])
def test_knn_interpolate(dev, k):
    from garmentnets_b200 import ops
    d = _cloud_batch(2, 1024, seed=3)
    ptr = P.batch_to_ptr(d["batch"], 2)
    idx_c = P.fps(d["pos"], ptr, 0.25)
    ptr_c = np.concatenate([[0], np.cumsum(P.fps_counts(ptr, 0.25))])
    src = d["pos"][idx_c]
    feat = np.random.default_rng(1).normal(size=(len(src), 37)).astype(np.float32)
    idx_ref, d2_ref = P.knn(src, d["pos"], ptr_c, ptr, k)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    idx, d2 = ops.knn(t(src), t(d["pos"]), t(ptr_c), t(ptr), k)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    assert np.array_equal(d2.cpu().numpy(), d2_ref)
    out = torch.zeros((len(d["pos"]), 40), device=dev)
    ops.knn_interpolate_into(t(feat), idx, d2, out)
    ref = P.knn_interpolate(feat, idx_ref, d2_ref)
    assert np.array_equal(out[:, :37].cpu().numpy(), ref)  # same op order, no FMA -> bit-exact
    assert torch.all(out[:, 37:] == 0)
    # 16-byte channel path (C, both row strides multiples of 4): 256 channels written into a wider concat buffer, and 130
    # channels (two passes of 128 per warp, the second one partial)
    for C in (256, 132):
        featv = np.random.default_rng(2).normal(size=(len(src), C)).astype(np.float32)
        outv = torch.full((len(d["pos"]), C + 8), -7.0, device=dev)
        ops.knn_interpolate_into(t(featv), idx, d2, outv)
        assert np.array_equal(outv[:, :C].cpu().numpy(), P.knn_interpolate(featv, idx_ref, d2_ref))
        assert torch.all(outv[:, C:] == -7.0)


@pytest.mark.gpu
def test_knn_fewer_points_than_k(dev):
    from garmentnets_b200 import ops
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    src = np.zeros((2, 3), np.float32)  # one coarse point per cloud (FP3 configuration, k=1 there; k=3 here)
    y = np.random.default_rng(0).normal(size=(10, 3)).astype(np.float32)
    idx, d2 = ops.knn(t(src), t(y), t(np.array([0, 1, 2])), t(np.array([0, 4, 10])), 3)
    idx_ref, d2_ref = P.knn(src, y, np.array([0, 1, 2]), np.array([0, 4, 10]), 3)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    feat = np.array([[1.0, 2.0], [3.0, 4.0]], np.float32)
    out = torch.zeros((10, 2), device=dev)
    ops.knn_interpolate_into(t(feat), idx, d2, out)
    assert np.array_equal(out.cpu().numpy(), P.knn_interpolate(feat, idx_ref, d2_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 1000, 1024, 16384, 16385, 65536, 70000])
def test_exclusive_scan(dev, n):
    """Pass boundaries of the 16-elements-per-thread scan (16384 per pass), ragged tails, an input that does not start on a
    16-byte boundary (scalar loads), and sums beyond 32 bits."""
    from garmentnets_b200 import ops
    v = torch.randint(0, 66, (n,), dtype=torch.int32, device=dev)
    out = ops.exclusive_scan(v).cpu().numpy()
    ref = np.concatenate([[0], np.cumsum(v.cpu().numpy().astype(np.int64))])
    assert np.array_equal(out, ref)
    if n > 1:
        w = torch.randint(0, 66, (n + 1,), dtype=torch.int32, device=dev)[1:]      # 4-byte aligned view
        assert np.array_equal(ops.exclusive_scan(w).cpu().numpy(), np.concatenate([[0], np.cumsum(w.cpu().numpy().astype(np.int64))]))
        big = torch.full((n,), 2 ** 31 - 1, dtype=torch.int32, device=dev)
        assert int(ops.exclusive_scan(big)[-1]) == n * (2 ** 31 - 1)


@pytest.mark.gpu
@pytest.mark.parametrize("nseg,C,maxlen", [(32, 1024, 512), (1000, 128, 70), (17, 6, 9), (5, 130, 40)])
def test_segment_max_shapes(dev, nseg, C, maxlen):
    """Segment max over CSR rows (PointConv aggregation and global_max_pool, ref components/pointnet2.py:31,49): wide
    channel counts, channel counts that are not a multiple of 4, empty segments (-> 0) -- bit-exact."""
    from garmentnets_b200 import ops
    rng = np.random.default_rng(nseg + C)
    lens = rng.integers(0, maxlen + 1, nseg)
    lens[0] = 0
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rows = rng.normal(size=(int(offs[-1]), C)).astype(np.float32)
    got = ops.segment_max(torch.from_numpy(rows).to(dev), torch.from_numpy(offs).to(dev)).cpu().numpy()
    ref = np.zeros((nseg, C), np.float32)
    for i in range(nseg):
        if lens[i]:
            ref[i] = rows[offs[i]:offs[i + 1]].max(0)
    assert np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,channels,r,sizes", [(3, [6, 64, 64, 128], 0.08, [1500, 900]), (128, [131, 128, 128, 256], 0.15, [700, 1100, 300]),
                                                  (3, [6, 64, 64, 128], 0.6, [400])])
def test_fused_pointconv_matches_unfused_chain(dev, cin, channels, r, sizes):
    """gnb_pointconv_mlp_max (gather + three Linear->ReLU->BN blocks on tcgen05 + max aggregation in one kernel, activations in
    tensor memory) against the unfused chain gather / gnb_linear_tc x 3 / gnb_segment_max, which the oracle tests pin
    (ref components/pointnet2.py:30-31).  The third case saturates the 64-neighbour limit (65 edges per centre with the self loop)."""
    from garmentnets_b200 import ops, synthetic
    from garmentnets_b200.components.mlp import MLP
    from garmentnets_b200.components.pointnet2 import CloudIndex, PointConv, fps
    torch.manual_seed(cin)
    conv = PointConv(synthetic.randomize_(MLP(channels), 3)).eval().requires_grad_(False).to(dev)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    index = CloudIndex(torch.from_numpy(ptr).to(dev), ptr)
    g = torch.Generator().manual_seed(sum(sizes))
    pos = torch.rand(int(ptr[-1]), 3, generator=g).to(dev)
    x = (torch.randn(int(ptr[-1]), cin, generator=g) * (1.0 if cin == 3 else 0.7)).to(dev)
    sub = index.subsample(0.5)
    idx = fps(pos, None, 0.5, False, index=index)
    pos_y = pos[idx]
    nbr, cnt = ops.ball_query(pos, pos_y, index.ptr, sub.ptr, r, 64)
    if r > 0.5:
        assert int(cnt.max()) == 64
    assert conv._fused_layers(cin) is not None
    try:
        ops.USE_SA_MLP = False
        ref = conv.forward_grouped(x, pos, pos_y, nbr, cnt)
        ops.USE_SA_MLP = True
        got = conv.forward_grouped(x, pos, pos_y, nbr, cnt)
        again = conv.forward_grouped(x, pos, pos_y, nbr, cnt)
    finally:
        ops.USE_SA_MLP = True
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (pos_y.shape[0], channels[-1])
    assert torch.equal(got, again)                       # integer atomics: deterministic
    tol = 3e-5 * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) < tol, (float((got - ref).abs().max()), tol)
