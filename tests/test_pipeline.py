"""Model-level parity: the CUDA pipeline (through the C-ABI) vs the CPU oracle on the same seeded weights/inputs.
Stage-wise: every stage is fed the ORACLE's inputs so that a flipped argmax upstream cannot hide or fake an error."""
import numpy as np
import pytest
import torch

from garmentnets_b200 import synthetic
from oracle import nets as ON
from oracle import pipeline as OP
from oracle import pointops as P

TOL = 1e-4  # north_star tolerance for the fp32 output fields (winding number, warp field): max-abs


def close(got, ref, tol=TOL):
    """Intermediate feature tensors are not O(1) (the synthetic BatchNorm statistics give activations up to ~1e2 and
    amplify fp32 summation-order noise accordingly), so they are compared with the 1e-4 bound scaled by
    max(1, max|ref|) of the tensor; the output fields (winding number, warp field) use the plain max-abs bound."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref))) < tol * max(1.0, float(np.max(np.abs(ref))))


def _small_hparams():
    import copy
    hp = copy.deepcopy(synthetic.HPARAMS)
    hp["prediction"]["volume_size"] = 32
    return hp


@pytest.fixture(scope="module")
def setup(dev):
    from garmentnets_b200.components.pointnet2 import CloudIndex
    from garmentnets_b200.pipeline import Batch
    hp = _small_hparams()
    B, n = 2, 1024
    d = synthetic.make_batch(B, n, "Tshirt", seed=5)
    model = synthetic.build_pipeline(seed=3, device=dev, hparams=hp)
    data = Batch(x=torch.from_numpy(d["x"]).to(dev), pos=torch.from_numpy(d["pos"]).to(dev),
                 batch=torch.from_numpy(d["batch"]).to(dev))
    index = CloudIndex.uniform(B, n, dev)
    synthetic.prepare_model_(model, data, index)
    sd = OP.to_cpu_state_dict(model)
    starts = (np.array([5, 17]), np.array([3, 0]))
    s1 = OP.stage1(sd, hp, d["x"], d["pos"], d["batch"], B, starts)
    s2 = OP.stage2(sd, hp, s1, d["pos"], d["batch"], B)
    return dict(hp=hp, B=B, n=n, d=d, model=model, data=data, index=index, sd=sd, starts=starts, s1=s1, s2=s2, dev=dev)


@pytest.mark.gpu
def test_state_dict_keys_follow_reference_names(setup):
    keys = set(setup["sd"].keys())
    for k in ["pointnet2_nocs.sa1_module.conv.local_nn.0.0.weight", "pointnet2_nocs.sa2_module.conv.local_nn.2.2.running_var",
              "pointnet2_nocs.sa3_module.nn.1.0.bias", "pointnet2_nocs.fp3_module.nn.0.0.weight", "pointnet2_nocs.lin3.weight",
              "pointnet2_nocs.global_lin2.bias", "volume_agg.local_nn.1.2.num_batches_tracked",
              "unet_3d.abstract_3d_unet.encoders.0.basic_module.SingleConv1.groupnorm.weight",
              "unet_3d.abstract_3d_unet.decoders.2.basic_module.SingleConv2.conv.weight",
              "unet_3d.abstract_3d_unet.final_conv.bias", "volume_decoder.mlp.2.0.weight", "surface_decoder.mlp.2.2.bias"]:
        assert k in keys, k
    assert sum(v.numel() for k, v in setup["sd"].items() if k.startswith("unet_3d") ) == 4624640


@pytest.mark.gpu
def test_pointnet2_stage(setup):
    s = setup
    dev = s["dev"]
    starts = tuple(torch.from_numpy(a.astype(np.int64)).to(dev) for a in s["starts"])
    res = s["model"].pointnet2_forward(s["data"], index=s["index"], fps_starts=starts, return_aux=True)
    s1 = s["s1"]
    for name in ("sa1", "sa2"):
        _, _, aux = res["aux"][name]
        ref_aux = s1[name][3]
        assert np.array_equal(aux["idx"].cpu().numpy(), ref_aux["idx"]), name       # FPS bit-exact
        assert np.array_equal(aux["cnt"].cpu().numpy(), ref_aux["cnt"]), name       # ball query bit-exact
        assert np.array_equal(aux["nbr"].cpu().numpy(), ref_aux["nbr"]), name
        assert close(res["aux"][name][0].cpu().numpy(), s1[name][0]), name
    for name in ("fp3", "fp2", "fp1"):
        assert close(res["aux"][name].cpu().numpy(), s1[name + "_x"]), name
    assert close(res["global_feature"].cpu().numpy(), s1["global_feature"])
    assert close(res["per_point_features"].cpu().numpy(), s1["per_point_features"])
    assert close(res["per_point_logits"].cpu().numpy(), s1["per_point_logits"])
    assert close(res["global_logits"].cpu().numpy(), s1["global_logits"])
    nd = res["nocs_data"]
    same = (nd.pos.cpu().numpy() == s1["pred_nocs"]).all(axis=1)
    assert same.mean() > 0.995  # argmax of near-tied logits may flip for a handful of points
    assert np.abs(nd.pred_confidence.cpu().numpy() - s1["pred_confidence"])[same].max() < TOL


@pytest.mark.gpu
def test_aggregator_and_unet_stage(setup):
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev, s1, s2 = s["dev"], s["s1"], s["s2"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nocs_data = Batch(x=t(s1["per_point_features"]), pos=t(s1["pred_nocs"]), batch=s["data"].batch,
                      sim_points=s["data"].pos, pred_confidence=t(s1["pred_confidence"]))
    nocs_data.num_graphs = s["B"]
    out = s["model"].unet3d_forward({"nocs_data": nocs_data})
    vin = out["in_feature_volume"]
    assert tuple(vin.shape) == (s["B"], 128, 32, 32, 32)
    assert vin.permute(0, 2, 3, 4, 1).is_contiguous()  # channels-last all the way, no transposes
    ref_in = s2["in_feature_volume"]
    assert close(vin.cpu().numpy(), ref_in)
    assert np.array_equal(vin.cpu().numpy() == 0, ref_in == 0)  # same occupancy pattern; empty voxels are exactly 0
    # UNet on the ORACLE's input volume
    vout = s["model"].unet_3d(t(ref_in))
    assert tuple(vout.shape) == (s["B"], 128, 32, 32, 32)
    assert close(vout.cpu().numpy(), s2["out_feature_volume"]), np.abs(vout.cpu().numpy() - s2["out_feature_volume"]).max()


@pytest.mark.gpu
def test_decoders(setup):
    s = setup
    dev, s2, sd = s["dev"], s["s2"], s["sd"]
    fvol = torch.from_numpy(s2["out_feature_volume"]).to(dev)
    g = torch.Generator().manual_seed(0)
    q = torch.rand(s["B"], 700, 3, generator=g) * 1.1 - 0.05
    ref_v = ON.implicit_decoder(sd, "volume_decoder.", s2["out_feature_volume"], q)
    ref_s = ON.implicit_decoder(sd, "surface_decoder.", s2["out_feature_volume"], q)
    u = {"out_feature_volume": fvol}
    got_v = s["model"].volume_decoder_forward(u, q.to(dev))
    got_s = s["model"].surface_decoder_forward(u, q.to(dev))
    assert got_v["pred_volume_value"].shape == (s["B"], 700)
    assert (got_v["out_features"].cpu() - ref_v).abs().max().item() < TOL
    assert (got_s["out_features"].cpu() - ref_s).abs().max().item() < TOL
    # dense lattice decode (predict.py:145-158) at Q=32 against the chunked oracle loop
    Q = 32
    ref_d = ON.dense_decode(sd, "volume_decoder.", s2["out_feature_volume"][:1], Q, 16)
    got_d = s["model"].dense_decode(fvol[:1], Q)[0]
    assert (got_d.cpu() - ref_d).abs().max().item() < TOL
    assert (ref_d > 0.5).float().mean().item() > 0.02  # calibration gives a surface


@pytest.mark.gpu
def test_predict_end_to_end(setup):
    """Whole predict loop vs the oracle on sample 0 (meshes compared on the ORACLE's wnf volume for exactness, then
    end to end with tolerance-level statistics)."""
    from garmentnets_b200 import ops
    from garmentnets_b200.components.pointnet2 import CloudIndex
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev, hp, d, n = s["dev"], s["hp"], s["d"], s["n"]
    ref = OP.predict_sample(s["sd"], hp, d["x"][:n], d["pos"][:n], (s["starts"][0][:1], s["starts"][1][:1]))
    mesh = ref["mesh"]
    # (1) CUDA tail on the oracle's volume: bit-exact topology
    wnf = torch.from_numpy(ref["wnf_volume"]).to(dev)
    ggm = ops.gaussian_gradient_magnitude(wnf, 0.5)
    verts, faces, normals, values, ggm_at = ops.marching_cubes(wnf, 0.5, (1 / 31,) * 3, "ascent", ggm)
    assert np.array_equal(faces.cpu().numpy(), mesh["faces"])
    assert np.array_equal(verts.cpu().numpy(), mesh["verts"])
    assert np.abs(ggm_at.cpu().numpy() - mesh["volume_gradient_magnitude"]).max() < 1e-6
    fv = torch.from_numpy(ref["stage2"]["out_feature_volume"]).to(dev)
    warp = s["model"].surface_decoder(fv, verts.view(1, -1, 3)).view(-1, 3)
    assert np.abs(warp.cpu().numpy() - mesh["warp_field"]).max() < TOL
    # (2) full CUDA predict for the single sample.  A NOCS argmax whose two best logits tie within rounding may land on
    # the other bin; that point then feeds another voxel and, through GroupNorm, perturbs the whole volume.  So: (a) every
    # flipped bin must be such a near-tie in the ORACLE's logits, and (b) the oracle continues from the CUDA stage-1
    # outputs (identical voxel assignment by construction) and must reproduce the CUDA volume.
    one = Batch(x=s["data"].x[:n], pos=s["data"].pos[:n], batch=s["data"].batch[:n])
    starts = tuple(torch.from_numpy(a[:1].astype(np.int64)).to(dev) for a in s["starts"])
    index1 = CloudIndex.uniform(1, n, dev)
    out = s["model"].predict(one, volume_size=32, index=index1, fps_starts=starts, keep_volume=True)[0]
    p1 = s["model"].pointnet2_forward(one, index=index1, fps_starts=starts)
    nocs_gpu = p1["nocs_data"].pos.cpu().numpy()
    s1_ref = ref["stage1"]
    logits_ref = s1_ref["per_point_logits"].reshape(n, 64, 3)
    bins_gpu = np.rint(nocs_gpu * 63).astype(np.int64)
    flipped = np.argwhere(bins_gpu != s1_ref["nocs_bin"])
    assert len(flipped) <= 0.01 * n
    scale = max(1.0, float(np.abs(logits_ref).max()))
    for pt, ax in flipped:
        gap = logits_ref[pt, :, ax].max() - logits_ref[pt, bins_gpu[pt, ax], ax]
        assert gap < TOL * scale, (pt, ax, gap)
    s1_gpu = {"per_point_features": p1["per_point_features"].cpu().numpy(), "pred_nocs": nocs_gpu,
              "pred_confidence": p1["nocs_data"].pred_confidence.cpu().numpy()}
    batch0 = np.zeros(n, np.int64)
    s2_cont = OP.stage2(s["sd"], hp, s1_gpu, d["pos"][:n], batch0, 1)
    wnf_cont = ON.dense_decode(s["sd"], "volume_decoder.", s2_cont["out_feature_volume"], 32, 16).numpy()
    diff = np.abs(out["wnf_volume"].cpu().numpy() - wnf_cont)
    assert np.median(diff) < 1e-5 and (diff > TOL).mean() < 0.02 and diff.max() < 1e-3, (np.median(diff), (diff > TOL).mean(), diff.max())
    if len(flipped) == 0:
        assert abs(len(out["verts"]) - len(mesh["verts"])) <= 0.05 * len(mesh["verts"]) + 8
    assert out["faces"].dtype == torch.int32 and out["warp_field"].shape == (len(out["verts"]), 3)


@pytest.mark.gpu
def test_folded_final_conv_equals_two_step(setup):
    """predict() folds the UNet's 1x1x1 final_conv into each decoder's first Linear (two affine maps = one): the folded
    32 -> 256 map must reproduce hoisted(final_conv(x)) (ref components/unet3d.py:467, networks/conv_implicit_wnf.py:148)."""
    from garmentnets_b200 import ops
    s = setup
    model, dev = s["model"], s["dev"]
    unet = model.unet_3d.abstract_3d_unet
    vol_in = torch.from_numpy(np.ascontiguousarray(s["s2"]["in_feature_volume"])).to(dev)
    x_last = unet.forward_ndhwc(ops.to_channels_last(vol_in), apply_final=False)
    full = unet.forward_ndhwc(ops.to_channels_last(vol_in))
    assert x_last.shape[-1] == 32 and full.shape[-1] == 128
    for dec in (model.volume_decoder, model.surface_decoder):
        two_step = dec.hoisted(full)
        folded = dec.hoisted_folded(x_last, unet.final_conv)
        assert close(folded.cpu().numpy(), two_step.cpu().numpy(), 2e-5)


@pytest.mark.gpu
def test_host_predictor_matches_device_predict(setup):
    """The host-buffer API (pinned staging, copy stream, double buffering) returns exactly what predict() leaves on the
    device, for every sample and for consecutive overlapping submissions."""
    from garmentnets_b200.pipeline import HostPredictor
    model, d, index = setup["model"], setup["d"], setup["index"]
    ref = model.predict(setup["data"], volume_size=32, index=index)   # synthetic.build_pipeline fixes the FPS starts
    ref = [{k: v.cpu().numpy() for k, v in r.items()} for r in ref]
    nocs_ref = model._last_point_outputs["pred_nocs"].cpu().numpy()
    hp = HostPredictor(model, depth=2, volume_size=32)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
    tickets = [hp.submit(host["x"], host["pos"], host["batch"], index=index) for _ in range(2)]
    for t in tickets:
        res = hp.result(t)
        assert len(res) == len(ref)
        for got, want in zip(res, ref):
            for k in ("verts", "faces", "normals", "volume_value", "volume_gradient_magnitude", "warp_field"):
                assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k]), k
        assert np.array_equal(hp.point_outputs(t)["pred_nocs"], nocs_ref)
        assert t["d2h_bytes"] == sum(v.nbytes for r in ref for v in r.values()) + 2 * nocs_ref.nbytes


@pytest.mark.gpu
@pytest.mark.parametrize("category", synthetic.CATEGORIES)
def test_pointnet2_stage_six_categories_ragged(setup, category):
    """BASELINE.json configs[4] (six-category sweep) as a parity case: the category generators only change the
    neighbour-count statistics (ball-query truncation rate).  Ragged batch (two clouds of different size): FPS and
    ball-query indices bit-exact against the oracle, features / logits within tolerance."""
    from garmentnets_b200.components.pointnet2 import CloudIndex
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev, hp = s["dev"], s["hp"]
    cat_id = synthetic.CATEGORIES.index(category)
    sizes = [900 + 17 * cat_id, 640]
    clouds = [synthetic.make_cloud(category, n, 40 + i) for i, n in enumerate(sizes)]
    pos = np.concatenate([c[0] for c in clouds]).astype(np.float32)
    rgb = np.concatenate([c[1] for c in clouds]).astype(np.float32)
    batch = np.repeat(np.arange(2), sizes).astype(np.int64)
    starts = (np.array([3, 11]), np.array([0, 5]))
    s1 = OP.stage1(s["sd"], hp, rgb, pos, batch, 2, starts)
    data = Batch(x=torch.from_numpy(rgb).to(dev), pos=torch.from_numpy(pos).to(dev), batch=torch.from_numpy(batch).to(dev))
    index = CloudIndex.from_batch(data.batch)
    st = tuple(torch.from_numpy(a.astype(np.int64)).to(dev) for a in starts)
    res = s["model"].pointnet2_forward(data, index=index, fps_starts=st, return_aux=True)
    for name in ("sa1", "sa2"):
        _, _, aux = res["aux"][name]
        ref_aux = s1[name][3]
        assert np.array_equal(aux["idx"].cpu().numpy(), ref_aux["idx"]), name
        assert np.array_equal(aux["cnt"].cpu().numpy(), ref_aux["cnt"]), name
        assert np.array_equal(aux["nbr"].cpu().numpy(), ref_aux["nbr"]), name
        assert close(res["aux"][name][0].cpu().numpy(), s1[name][0]), name
    assert close(res["per_point_features"].cpu().numpy(), s1["per_point_features"])
    assert close(res["per_point_logits"].cpu().numpy(), s1["per_point_logits"])


@pytest.mark.gpu
def test_config1_full_size_pointnet_and_gridding(setup):
    """BASELINE.json configs[1] as a parity case at its full size: batch = 16 clouds x 4096 points through PointNet++
    SA/FP and the 32^3 scatter-gridding.  FPS and ball-query indices bit-exact against the oracle for all 16 clouds
    (2048 + 512 serial FPS rounds and 64-neighbour truncation at the real density); the gridded volume on the oracle's
    per-point outputs: identical occupancy, values within tolerance."""
    from garmentnets_b200.components.pointnet2 import CloudIndex
    from garmentnets_b200.pipeline import Batch
    s = setup
    dev, hp = s["dev"], s["hp"]
    B, n = 16, 4096
    d = synthetic.make_batch(B, n, "Tshirt", seed=77)
    rng = np.random.default_rng(1)
    starts = (rng.integers(0, n, B), rng.integers(0, n // 2, B))
    s1 = OP.stage1(s["sd"], hp, d["x"], d["pos"], d["batch"], B, starts)
    data = Batch(x=torch.from_numpy(d["x"]).to(dev), pos=torch.from_numpy(d["pos"]).to(dev),
                 batch=torch.from_numpy(d["batch"]).to(dev))
    index = CloudIndex.uniform(B, n, dev)
    st = tuple(torch.from_numpy(a.astype(np.int64)).to(dev) for a in starts)
    res = s["model"].pointnet2_forward(data, index=index, fps_starts=st, return_aux=True)
    for name, m in (("sa1", 2048), ("sa2", 512)):
        _, _, aux = res["aux"][name]
        ref_aux = s1[name][3]
        assert aux["idx"].shape[0] == B * m
        assert np.array_equal(aux["idx"].cpu().numpy(), ref_aux["idx"]), name
        assert np.array_equal(aux["cnt"].cpu().numpy(), ref_aux["cnt"]), name
        assert np.array_equal(aux["nbr"].cpu().numpy(), ref_aux["nbr"]), name
    assert int(res["aux"]["sa1"][2]["cnt"].max()) == 64          # the truncation at 64 neighbours is exercised
    assert close(res["per_point_logits"].cpu().numpy(), s1["per_point_logits"])
    # gridding on the oracle's stage-1 outputs (a flipped argmax upstream would move a point to another voxel)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nocs_data = Batch(x=t(s1["per_point_features"]), pos=t(s1["pred_nocs"]), batch=data.batch, sim_points=data.pos,
                      pred_confidence=t(s1["pred_confidence"]))
    nocs_data.num_graphs = B
    vin = s["model"].volume_agg(nocs_data)
    assert tuple(vin.shape) == (B, 128, 32, 32, 32)
    ref_in = ON.volume_feature_aggregator(s["sd"], "volume_agg.", s1["per_point_features"], s1["pred_nocs"], d["pos"],
                                          s1["pred_confidence"], d["batch"], B, 32)[0]
    got = vin.cpu().numpy()
    assert np.array_equal(got == 0, ref_in == 0)
    assert close(got, ref_in)
