"""Pin the oracle (and the host-side gridding mirror) against fixtures generated from the REFERENCE's own modules by
oracle/make_golden.py (components/unet3d.py, components/mlp.py, components/gridding.py run in the build container)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nets as ON

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(G, name))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, sd


META = json.load(open(os.path.join(G, "meta.json")))


def test_oracle_unet_matches_reference_module():
    z, sd = _load("unet3d_small.npz")
    y = ON.unet3d_forward(sd, "", z["x"], num_levels=3, groups=8)
    assert np.abs(y.numpy() - z["y"]).max() < 1e-5


def test_oracle_mlp_matches_reference_module():
    z, sd = _load("mlp_small.npz")
    y = ON.mlp(sd, "", torch.from_numpy(z["x"]))
    assert np.abs(y.numpy() - z["y"]).max() < 1e-6
    assert META["mlp_keys"][:3] == ["0.0.weight", "0.0.bias", "0.2.weight"]


def test_oracle_decoder_matches_reference_grid_sample_plus_mlp():
    z, sd = _load("decoder_small.npz")
    y = ON.implicit_decoder(sd, "", z["fg"], z["q"])
    assert np.abs(y.numpy() - z["y"]).max() < 1e-6


def test_oracle_grid_helpers_match_reference_virtual_grid():
    z, _ = _load("virtual_grid.npz")
    assert np.array_equal(ON.points_grid_idxs(z["pts"], 32), z["cell"])
    assert z["cell"][:, 0].tolist() == [(31 * k) // 63 for k in range(64)]
    b, conf, nocs = ON.nocs_head(np.eye(64, dtype=np.float32)[:, :, None].repeat(3, 2).reshape(64, 192) * 5, 64)
    assert np.array_equal(nocs, z["pts"]) and np.array_equal(b, z["bins"])
    gp = ON.grid_points(9).numpy()
    assert np.array_equal(gp, z["gp9"])
    gp128 = ON.grid_points(128)
    assert np.array_equal(gp128[1, 2, 3].numpy(), z["gp128_123"]) and np.array_equal(gp128[127, 127, 127].numpy(), z["gp128_last"])
    assert float(z["pts"][1, 0]) == 0.01587301678955555 and float(z["pts"][63, 0]) == 1.0


def test_product_virtual_grid_and_slicer_match_reference():
    """The host-side index helpers of the product are plain tensor arithmetic and run on CPU tensors."""
    from garmentnets_b200.components.gridding import ArraySlicer, VirtualGrid, ceil_div
    z, _ = _load("virtual_grid.npz")
    vg64 = VirtualGrid(grid_shape=(64,) * 3, batch_size=1)
    assert np.array_equal(vg64.idxs_to_points(torch.from_numpy(z["bins"])).numpy(), z["pts"])
    vg32 = VirtualGrid(grid_shape=(32,) * 3, batch_size=4)
    assert vg32.num_grids == int(z["num_grids"]) == 4 * 32 ** 3
    assert np.array_equal(vg32.get_points_grid_idxs(torch.from_numpy(z["pts"])).numpy(), z["cell"])
    cell = vg32.get_points_grid_idxs(torch.from_numpy(z["rnd"]), batch_idx=torch.from_numpy(z["bidx"]))
    assert np.array_equal(cell.numpy(), z["cell_rnd"])
    assert np.array_equal(vg32.flatten_idxs(cell).numpy(), z["flat_rnd"])
    assert np.array_equal(vg32.idxs_to_points(cell).numpy(), z["origin_rnd"])
    assert vg32.flatten_idxs(torch.tensor([[1, 2, 3, 4]])).item() == 34916 == int(z["flat_kat"][0])
    assert np.array_equal(vg32.unflatten_idxs(torch.from_numpy(z["flat_rnd"])).numpy(), z["cell_rnd"])
    odd = VirtualGrid(lower_corner=(-1, 0, 0.5), upper_corner=(1, 2, 1.5), grid_shape=(5, 6, 7), batch_size=2)
    c_odd = odd.get_points_grid_idxs(torch.from_numpy(z["rnd"]))
    assert np.array_equal(c_odd.numpy(), z["cell_odd"])
    assert np.array_equal(odd.idxs_to_points(c_odd).numpy(), z["pts_odd"])
    assert np.array_equal(VirtualGrid(grid_shape=(9,) * 3).get_grid_points(include_batch=False).numpy(), z["gp9"])
    assert VirtualGrid(grid_shape=(3, 4, 5), batch_size=2).get_grid_idxs().shape == (2, 3, 4, 5, 4)
    sl = ArraySlicer((128, 128, 128, 3), (64, 64, 64))
    assert len(sl) == META["array_slicer_len"] == 8
    assert [[(s.start, s.stop) for s in x] for x in sl] == [[tuple(t) for t in x] for x in META["array_slicer_128_64"]]
    odd_sl = ArraySlicer((10, 7, 3), (4, 7))
    assert [[(s.start, s.stop) for s in odd_sl[i]] for i in range(len(odd_sl))] == [[tuple(t) for t in x] for x in META["array_slicer_odd"]]
    assert ceil_div(7, 2) == 4


def test_product_unet_module_tree_matches_reference_keys():
    from garmentnets_b200.components.unet3d import Abstract3DUNet, DoubleConv
    m = Abstract3DUNet(128, 128, False, DoubleConv, f_maps=32, layer_order="gcr", num_groups=8, num_levels=4,
                       is_segmentation=False)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == META["unet_full_keys"]
    assert sum(p.numel() for p in m.parameters()) == META["unet_full_param_count"] == 4624640


def test_oracle_ggm_matches_scipy_fixture():
    from oracle import postproc
    z = np.load(os.path.join(G, "ggm.npz"))
    assert np.array_equal(postproc.gaussian_gradient_magnitude(z["vol"], 0.5), z["vol_ggm"])
    assert abs(z["impulse"][4, 4, 5] - 0.26344162) < 1e-7 and abs(z["impulse"][5, 5, 5] - 0.008357321) < 1e-8


@pytest.mark.gpu
def test_cuda_unet_and_decoder_match_reference_fixtures(dev):
    """The CUDA path against the REFERENCE module outputs directly (not just against the oracle)."""
    from garmentnets_b200.components import mlp as PM
    from garmentnets_b200.components.gridding import nocs_grid_sample
    from garmentnets_b200.components.unet3d import Abstract3DUNet, DoubleConv
    from garmentnets_b200.pipeline import ImplicitWNFDecoder
    z, sd = _load("unet3d_small.npz")
    net = Abstract3DUNet(16, 8, False, DoubleConv, f_maps=8, layer_order="gcr", num_groups=8, num_levels=3,
                         is_segmentation=False)
    net.load_state_dict(sd)
    y = net.to(dev).eval()(torch.from_numpy(z["x"]).to(dev))
    assert tuple(y.shape) == z["y"].shape
    assert np.abs(y.cpu().numpy() - z["y"]).max() < 1e-4
    z, sd = _load("mlp_small.npz")
    m = PM.MLP([7, 16, 5])
    m.load_state_dict(sd)
    assert np.abs(m.to(dev).eval()(torch.from_numpy(z["x"]).to(dev)).cpu().numpy() - z["y"]).max() < 1e-5
    z, sd = _load("decoder_small.npz")
    dec = ImplicitWNFDecoder(nn_channels=(6, 12, 12, 2))
    dec.load_state_dict(sd)
    y = dec.to(dev).eval()(torch.from_numpy(z["fg"]).to(dev), torch.from_numpy(z["q"]).to(dev))
    assert np.abs(y.cpu().numpy() - z["y"]).max() < 1e-4
    ys = nocs_grid_sample(torch.from_numpy(z["fg"]).to(dev), torch.from_numpy(z["q"]).to(dev))
    assert np.abs(ys.cpu().numpy() - z["nocs_sampled"]).max() < 1e-5


def _b2v_case():
    z = np.load(os.path.join(G, "batch_to_volume.npz"))
    return z


def test_oracle_batch_to_volume_index_arithmetic_matches_reference():
    """Row a9: the restated voxel-index arithmetic of ref components/gridding.py:18-27 against the reference function's
    own outputs (occupancy pattern and values, all four reduce modes)."""
    from oracle import pointops as P
    z = _b2v_case()
    Gs, B = int(z["G"]), int(z["B"])
    ijk = np.clip((z["pos"] * np.float32(Gs)).astype(np.int64), 0, Gs - 1)
    flat = z["batch"] * Gs ** 3 + ijk[:, 0] * Gs ** 2 + ijk[:, 1] * Gs + ijk[:, 2]
    for reduce in ("mean", "max", "sum", "min"):
        vol = P.scatter(z["x"].T, flat, B * Gs ** 3, reduce).reshape(-1, B, Gs, Gs, Gs).transpose(1, 0, 2, 3, 4)
        assert np.array_equal(vol, z["vol_" + reduce]), reduce


@pytest.mark.gpu
@pytest.mark.parametrize("reduce", ["mean", "max", "sum", "min"])
def test_cuda_batch_to_volume_matches_reference_function(dev, reduce):
    """Row a9 on the device: ``components.gridding.batch_to_volume`` (same name / signature as the reference's) against
    the reference function's own output."""
    from garmentnets_b200.components.gridding import batch_to_volume
    from garmentnets_b200.pipeline import Batch
    z = _b2v_case()
    batch = Batch(x=torch.from_numpy(z["x"]).to(dev), pos=torch.from_numpy(z["pos"]).to(dev),
                  batch=torch.from_numpy(z["batch"]).to(dev))
    batch.num_graphs = int(z["B"])
    vol = batch_to_volume(batch, int(z["G"]), reduce=reduce)
    ref = z["vol_" + reduce]
    assert tuple(vol.shape) == ref.shape
    got = vol.cpu().numpy()
    assert np.array_equal(got == 0, ref == 0)
    if reduce in ("max", "min"):
        assert np.array_equal(got, ref)
    else:
        assert np.abs(got - ref).max() < 1e-5
