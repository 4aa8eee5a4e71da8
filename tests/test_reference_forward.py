"""The reference's UNCHANGED ``networks/conv_implicit_wnf.py`` (staged byte for byte under ``oracle/_ref`` by
``oracle/make_ref.py``) executed on the B200 on top of ``garmentnets_b200.components`` through ``shims.install()``:
``pointnet2_forward / unet3d_forward / volume_decoder_forward / surface_decoder_forward / forward``
(ref networks/conv_implicit_wnf.py:213-269,314-338) against the CPU oracle and against our own ``pipeline``;
plus the reference's own ``components/pointnet2.py`` SAModule / FPModule / GlobalSAModule (ref :11-76) driven through our
``torch_geometric.nn`` stand-ins (``fps`` / ``radius`` / ``PointConv.forward(x, (pos, pos_y), edge_index)`` /
``knn_interpolate`` / ``global_max_pool``)."""
import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from garmentnets_b200 import synthetic
from oracle import nets as ON
from oracle import pipeline as OP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
TOL = 1e-4

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isfile(os.path.join(REFDIR, "networks", "conv_implicit_wnf.py")),
                                 reason="oracle/_ref not staged (run oracle/make_ref.py in the build container)")]


def close(got, ref, tol=TOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref))) < tol * max(1.0, float(np.max(np.abs(ref))))


@pytest.fixture(scope="module")
def ref_networks():
    from garmentnets_b200 import shims
    shims.install()
    saved = list(sys.path)
    sys.path.insert(0, REFDIR)
    try:
        for name in [m for m in sys.modules if m.split(".")[0] in ("networks", "common")]:
            del sys.modules[name]
        ciw = importlib.import_module("networks.conv_implicit_wnf")
        assert os.path.realpath(ciw.__file__).startswith(os.path.realpath(REFDIR))
        yield ciw
    finally:
        sys.path[:] = saved


@pytest.fixture(scope="module")
def setup(dev, ref_networks):
    from garmentnets_b200.components.pointnet2 import CloudIndex
    from garmentnets_b200.pipeline import Batch
    import copy
    hp = copy.deepcopy(synthetic.HPARAMS)
    B, n = 2, 2048
    d = synthetic.make_batch(B, n, "Tshirt", seed=21)
    model = synthetic.build_pipeline(seed=4, device=dev, hparams=hp)
    data = Batch(x=torch.from_numpy(d["x"]).to(dev), pos=torch.from_numpy(d["pos"]).to(dev),
                 batch=torch.from_numpy(d["batch"]).to(dev))
    index = CloudIndex.uniform(B, n, dev)
    synthetic.prepare_model_(model, data, index)
    ref = ref_networks.ConvImplicitWNFPipeline(
        pointnet2_params=hp["pointnet2"], volume_agg_params=hp["volume_agg"], unet3d_params=hp["unet3d"],
        volume_decoder_params=hp["volume_decoder"], surface_decoder_params=hp["surface_decoder"])
    ref.load_state_dict(model.state_dict(), strict=True)   # a reference-layout checkpoint: identical keys
    ref = ref.to(dev).eval()
    ref.requires_grad_(False)
    for m in (ref.pointnet2_nocs.sa1_module, ref.pointnet2_nocs.sa2_module):
        m.random_start = False   # torch_cluster draws the FPS start at random; pin it to point 0 of every cloud
    sd = OP.to_cpu_state_dict(model)
    starts = (np.zeros(B, np.int64), np.zeros(B, np.int64))
    s1 = OP.stage1(sd, hp, d["x"], d["pos"], d["batch"], B, starts)
    s2 = OP.stage2(sd, hp, s1, d["pos"], d["batch"], B)
    g = torch.Generator().manual_seed(5)
    vq = torch.rand(B, 900, 3, generator=g)
    sq = torch.rand(B, 500, 3, generator=g) * 1.1 - 0.05
    return dict(hp=hp, B=B, n=n, d=d, model=model, ref=ref, data=data, sd=sd, s1=s1, s2=s2, dev=dev, vq=vq, sq=sq)


def test_reference_classes_are_the_unmodified_files(ref_networks):
    import hashlib
    import json
    with open(os.path.join(REFDIR, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for rel, meta in manifest.items():
        with open(os.path.join(REFDIR, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == meta["sha256"], rel
    assert ref_networks.ConvImplicitWNFPipeline.__module__ == "networks.conv_implicit_wnf"


def test_reference_pointnet2_forward_on_gpu(setup):
    """ref networks/conv_implicit_wnf.py:213-240 (argmax / softmax / gather are the reference's own torch calls)."""
    s = setup
    res = s["ref"].pointnet2_forward(s["data"])
    s1 = s["s1"]
    assert res["per_point_logits"].is_cuda
    assert close(res["per_point_features"].cpu().numpy(), s1["per_point_features"])
    assert close(res["per_point_logits"].cpu().numpy(), s1["per_point_logits"])
    assert close(res["global_logits"].cpu().numpy(), s1["global_logits"])
    assert close(res["global_feature"].cpu().numpy(), s1["global_feature"])
    nd = res["nocs_data"]
    same = (nd.pos.cpu().numpy() == s1["pred_nocs"]).all(axis=1)
    assert same.mean() > 0.995
    assert np.abs(nd.pred_confidence.cpu().numpy() - s1["pred_confidence"])[same].max() < TOL
    assert nd.num_graphs == s["B"]
    # against our own stage function: the SA / FP stack is the same kernels (bit-identical features in front of the head);
    # the reference's head calls nn.Linear / argmax / softmax directly (ATen), ours the fused kernels -> tolerance there
    ours = s["model"].pointnet2_forward(s["data"])
    assert torch.equal(ours["global_feature"], res["global_feature"])
    assert close(ours["per_point_logits"].cpu().numpy(), res["per_point_logits"].cpu().numpy())
    assert (ours["nocs_data"].pos == nd.pos).all(dim=1).float().mean().item() > 0.995


def test_reference_unet3d_forward_on_gpu(setup):
    """ref networks/conv_implicit_wnf.py:242-251 incl. the reference VolumeFeatureAggregator.forward (:43-100) with its
    VirtualGrid arithmetic, torch.cat, torch_scatter.scatter(src=features.T, ...) and reshape/permute, on the ORACLE's
    stage-1 outputs."""
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev, s1, s2 = s["dev"], s["s1"], s["s2"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nocs = Batch(x=t(s1["per_point_features"]), pos=t(s1["pred_nocs"]), batch=s["data"].batch, sim_points=s["data"].pos,
                 pred_confidence=t(s1["pred_confidence"]))
    vin = s["ref"].volume_agg(nocs)
    assert tuple(vin.shape) == (s["B"], 128, 32, 32, 32)
    assert np.array_equal(vin.cpu().numpy() == 0, s2["in_feature_volume"] == 0)
    assert close(vin.cpu().numpy(), s2["in_feature_volume"])
    out = s["ref"].unet3d_forward({"nocs_data": nocs})["out_feature_volume"]
    assert tuple(out.shape) == (s["B"], 128, 32, 32, 32)
    assert close(out.cpu().numpy(), s2["out_feature_volume"]), np.abs(out.cpu().numpy() - s2["out_feature_volume"]).max()


def test_reference_decoder_forwards_on_gpu(setup):
    """ref networks/conv_implicit_wnf.py:253-269 -> ImplicitWNFDecoder.forward (:128-149)."""
    s = setup
    dev, s2, sd = s["dev"], s["s2"], s["sd"]
    u = {"out_feature_volume": torch.from_numpy(s2["out_feature_volume"]).to(dev)}
    got_v = s["ref"].volume_decoder_forward(u, s["vq"].to(dev))
    got_s = s["ref"].surface_decoder_forward(u, s["sq"].to(dev))
    ref_v = ON.implicit_decoder(sd, "volume_decoder.", s2["out_feature_volume"], s["vq"])
    ref_s = ON.implicit_decoder(sd, "surface_decoder.", s2["out_feature_volume"], s["sq"])
    assert got_v["pred_volume_value"].shape == (s["B"], 900)
    assert (got_v["out_features"].cpu() - ref_v).abs().max().item() < TOL
    assert (got_s["out_features"].cpu() - ref_s).abs().max().item() < TOL


def test_reference_full_forward_on_gpu(setup):
    """ref networks/conv_implicit_wnf.py:314-338: the whole ``forward(data)`` of the unchanged class, against our own
    ``ConvImplicitWNFPipeline.forward`` and, continuing the oracle from the CUDA stage-1
    outputs (identical voxel assignment by construction), against the oracle within 1e-4."""
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev = s["dev"]
    data = Batch(x=s["data"].x, pos=s["data"].pos, batch=s["data"].batch, volume_query_points=s["vq"].to(dev),
                 surf_query_points=s["sq"].to(dev))
    res = s["ref"](data)
    ours = s["model"].forward(data)
    # (the reference decoder calls F.grid_sample itself and then our MLP; ours is the fused gather + MLP kernel)
    for k in ("volume_decoder_result", "surface_decoder_result"):
        assert (res[k]["out_features"] - ours[k]["out_features"]).abs().max().item() < TOL, k
    p = res["pointnet2_result"]
    s1_gpu = {"per_point_features": p["per_point_features"].cpu().numpy(), "pred_nocs": p["nocs_data"].pos.cpu().numpy(),
              "pred_confidence": p["nocs_data"].pred_confidence.cpu().numpy()}
    s2c = OP.stage2(s["sd"], s["hp"], s1_gpu, s["d"]["pos"], s["d"]["batch"], s["B"])
    assert close(res["unet3d_result"]["out_feature_volume"].cpu().numpy(), s2c["out_feature_volume"])
    ref_v = ON.implicit_decoder(s["sd"], "volume_decoder.", s2c["out_feature_volume"], s["vq"])
    ref_s = ON.implicit_decoder(s["sd"], "surface_decoder.", s2c["out_feature_volume"], s["sq"])
    dv = (res["volume_decoder_result"]["out_features"].cpu() - ref_v).abs().max().item()
    ds = (res["surface_decoder_result"]["out_features"].cpu() - ref_s).abs().max().item()
    # (the synthetic last BatchNorm of the surface decoder yields a field of magnitude up to ~20: bound scaled by max|ref|)
    assert dv < TOL * max(1.0, ref_v.abs().max().item()) and ds < TOL * max(1.0, ref_s.abs().max().item()), (dv, ds)


def test_reference_forward_volume_task_space(setup, ref_networks):
    """ref networks/conv_implicit_wnf.py:279-310,320-322: with ``volume_task_space`` the aggregator grids the per-point features at
    the AABB-normalised simulation coordinates.  Our ``get_aabb_scale_offset`` / ``apply_volume_task_space`` against the
    reference's own methods (bit-equal: a handful of elementwise operations) and the two full forwards against each other."""
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev = s["dev"]
    g = torch.Generator().manual_seed(9)
    lo = -(torch.rand(s["B"], 3, generator=g) * 0.6 + 0.2)
    hi = torch.rand(s["B"], 3, generator=g) * 0.6 + 0.2
    aabb = torch.stack([lo, hi], dim=1).to(dev)                      # [B, 2, 3]
    sc_r, of_r = ref_networks.ConvImplicitWNFPipeline.get_aabb_scale_offset(aabb)
    sc_o, of_o = s["model"].get_aabb_scale_offset(aabb)
    assert torch.equal(sc_r, sc_o) and torch.equal(of_r, of_o)
    data = Batch(x=s["data"].x, pos=s["data"].pos, batch=s["data"].batch, volume_query_points=s["vq"].to(dev),
                 surf_query_points=s["sq"].to(dev), cloth_sim_aabb=aabb)
    plain = s["model"].forward(data)
    try:
        s["ref"].volume_task_space = True
        s["model"].volume_task_space = True
        res = s["ref"](data)
        ours = s["model"].forward(data)
    finally:
        s["ref"].volume_task_space = False
        s["model"].volume_task_space = False
    assert torch.equal(res["pointnet2_result"]["nocs_data"].pos, ours["pointnet2_result"]["nocs_data"].pos)
    assert close(ours["unet3d_result"]["out_feature_volume"].cpu().numpy(), res["unet3d_result"]["out_feature_volume"].cpu().numpy())
    for k in ("volume_decoder_result", "surface_decoder_result"):
        assert (res[k]["out_features"] - ours[k]["out_features"]).abs().max().item() < TOL, k
    # and the flag really changes the volume
    assert (plain["unet3d_result"]["out_feature_volume"] - ours["unet3d_result"]["out_feature_volume"]).abs().max().item() > 1e-3


def test_reference_forward_mc_surface_decoder(setup, ref_networks):
    """ref networks/conv_implicit_wnf.py:195-199,270-276,334-337: with ``mc_surface_loss_weight > 0`` the class owns a third implicit
    decoder and ``forward`` returns ``mc_surface_decoder_result``.  Same constructor call for the reference class and ours, strict
    ``state_dict`` load, outputs of the extra decoder against each other."""
    from garmentnets_b200.pipeline import Batch, ConvImplicitWNFPipeline
    s = setup
    dev, hp = s["dev"], s["hp"]
    kw = dict(pointnet2_params=hp["pointnet2"], volume_agg_params=hp["volume_agg"], unet3d_params=hp["unet3d"],
              volume_decoder_params=hp["volume_decoder"], surface_decoder_params=hp["surface_decoder"],
              mc_surface_decoder_params=hp["surface_decoder"], mc_surface_loss_weight=1.0)
    ours = ConvImplicitWNFPipeline(**kw)
    sd = dict(s["model"].state_dict())
    for k, v in list(sd.items()):                                   # the third decoder: a perturbed copy of the surface decoder
        if k.startswith("surface_decoder."):
            sd["mc_" + k] = v.clone() * 1.03 if v.dtype.is_floating_point else v.clone()
    ours.load_state_dict(sd, strict=True)
    ours = ours.to(dev).eval().requires_grad_(False)
    ours.pointnet2_nocs.set_random_start(False)
    ref = ref_networks.ConvImplicitWNFPipeline(**kw)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref = ref.to(dev).eval().requires_grad_(False)
    for m in (ref.pointnet2_nocs.sa1_module, ref.pointnet2_nocs.sa2_module):
        m.random_start = False
    data = Batch(x=s["data"].x, pos=s["data"].pos, batch=s["data"].batch, volume_query_points=s["vq"].to(dev),
                 surf_query_points=s["sq"].to(dev), mc_surf_query_points=s["sq"].to(dev).flip(1))
    a, b = ref(data), ours.forward(data)
    assert set(a.keys()) == set(b.keys()) and "mc_surface_decoder_result" in b
    for k in ("volume_decoder_result", "surface_decoder_result", "mc_surface_decoder_result"):
        ra, rb = a[k]["out_features"], b[k]["out_features"]
        assert (ra - rb).abs().max().item() < TOL * max(1.0, ra.abs().max().item()), k
    assert (b["mc_surface_decoder_result"]["out_features"] - b["surface_decoder_result"]["out_features"].flip(1)).abs().max().item() > 1e-3


def _load_ref_pointnet2():
    from garmentnets_b200 import shims
    shims.install()
    spec = importlib.util.spec_from_file_location("ref_components_pointnet2", os.path.join(REFDIR, "ref_components", "pointnet2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_own_sa_fp_modules_on_our_pyg_stand_ins(setup):
    """The reference's OWN ``components/pointnet2.py`` (not our mirror): SAModule.forward calls fps / radius /
    PointConv(x, (pos, pos[idx]), edge_index), FPModule calls knn_interpolate, GlobalSAModule global_max_pool -- all
    served by our ``torch_geometric.nn`` stand-ins.  Same weights as our mirror modules -> same outputs as the mirror
    (index sets bit-exact, features identical up to the order of the max) and as the oracle."""
    rp = _load_ref_pointnet2()
    s = setup
    model, data = s["model"], s["data"]
    net = model.pointnet2_nocs
    np.random.seed(0)
    sa1_ref = rp.SAModule(net.sa1_module.ratio, net.sa1_module.r, net.sa1_module.conv.local_nn).to(s["dev"]).eval()
    import garmentnets_b200.components.pointnet2 as ours
    # pin the random FPS start of the stand-in to point 0 of every cloud (torch_cluster draws it at random)
    orig_fps = ours.fps
    rp.fps = lambda pos, batch=None, ratio=0.5, **k: orig_fps(pos, batch, ratio, random_start=False)
    x1, pos1, b1 = sa1_ref(data.x, data.pos, data.batch)
    mx1, mpos1, mb1 = net.sa1_module(data.x, data.pos, data.batch)
    assert torch.equal(pos1, mpos1) and torch.equal(b1, mb1)
    # the reference module reaches PointConv.forward(x, pos, edge_index) (gather / three Linear blocks / segment max); our
    # SAModule runs the fused gnb_pointconv_mlp_max: same edges and weights, fp16 hi/lo tensor-core products both ways
    assert close(x1.cpu().numpy(), mx1.cpu().numpy())
    assert close(x1.cpu().numpy(), s["s1"]["sa1"][0]) and close(mx1.cpu().numpy(), s["s1"]["sa1"][0])
    sa2_ref = rp.SAModule(net.sa2_module.ratio, net.sa2_module.r, net.sa2_module.conv.local_nn).to(s["dev"]).eval()
    x2, pos2, b2 = sa2_ref(x1, pos1, b1)
    mx2, mpos2, mb2 = net.sa2_module(mx1, mpos1, mb1)
    assert torch.equal(pos2, mpos2) and close(x2.cpu().numpy(), mx2.cpu().numpy())
    assert close(x2.cpu().numpy(), s["s1"]["sa2"][0])
    sa3_ref = rp.GlobalSAModule(net.sa3_module.nn)
    x3, pos3, b3 = sa3_ref(x2, pos2, b2)
    assert close(x3.cpu().numpy(), s["s1"]["global_feature"])
    fp3_ref = rp.FPModule(net.fp3_module.k, net.fp3_module.nn)
    f3, _, _ = fp3_ref(x3, pos3, b3, x2, pos2, b2)
    assert close(f3.cpu().numpy(), s["s1"]["fp3_x"])
    fp2_ref = rp.FPModule(net.fp2_module.k, net.fp2_module.nn)
    f2, _, _ = fp2_ref(f3, pos2, b2, x1, pos1, b1)
    assert close(f2.cpu().numpy(), s["s1"]["fp2_x"])
    fp1_ref = rp.FPModule(net.fp1_module.k, net.fp1_module.nn)
    f1, _, _ = fp1_ref(f2, pos1, b1, data.x, data.pos, data.batch)
    assert close(f1.cpu().numpy(), s["s1"]["fp1_x"])


def test_pointconv_forward_edge_index_arbitrary_order(setup):
    """PointConv.forward accepts any edge order and reproduces PyG's remove_self_loops / add_self_loops quirk in the
    bipartite case (index equality, not point identity)."""
    from garmentnets_b200.components.pointnet2 import PointConv
    from oracle import pointops as P
    s = setup
    dev = s["dev"]
    conv = s["model"].pointnet2_nocs.sa1_module.conv
    g = torch.Generator().manual_seed(3)
    Nx, M, E = 300, 40, 900
    x = torch.rand(Nx, 3, generator=g)
    pos_x = torch.rand(Nx, 3, generator=g)
    pos_y = pos_x[torch.randperm(Nx, generator=g)[:M]].contiguous()
    src = torch.randint(0, Nx, (E,), generator=g)
    dst = torch.randint(0, M, (E,), generator=g)
    src[:10] = dst[:10]   # index-equal edges: dropped, then re-added once as the self loop
    out = conv(x.to(dev), (pos_x.to(dev), pos_y.to(dev)), torch.stack([src, dst]).to(dev))
    # oracle: per target the set {src of its edges with src != dst} U {dst}
    lists = [[int(j) for j, i in zip(src.tolist(), dst.tolist()) if i == t and j != t] + [t] for t in range(M)]
    offs = np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64)
    flat = np.concatenate([np.asarray(l, np.int64) for l in lists])
    edge = P.pointconv_edge_features(x.numpy(), pos_x.numpy(), pos_y.numpy(), offs, flat)
    h = ON.mlp(s["sd"], "pointnet2_nocs.sa1_module.conv.local_nn.", torch.from_numpy(edge)).numpy()
    want = P.segment_max(h, offs)
    assert close(out.cpu().numpy(), want)
