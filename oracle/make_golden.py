"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own importable modules
(/root/reference/components/{unet3d,mlp,gridding}.py) in this container.  The reference cannot travel to the GPU box,
so the vectors are committed together with this script.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz / *.json

Only torch CPU + the reference files are executed here; `torch_scatter` (needed only by the dead-code
`batch_to_volume`) is satisfied with an empty stub module so that `components.gridding` imports.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    sys.modules.setdefault("torch_scatter", types.ModuleType("torch_scatter"))
    sys.path.insert(0, REF)
    for name in [m for m in sys.modules if m == "components" or m.startswith("components.")]:
        del sys.modules[name]
    from components import gridding, mlp, unet3d
    sys.path.remove(REF)
    assert gridding.__file__.startswith(REF)
    return gridding, mlp, unet3d


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            if isinstance(m, (torch.nn.modules.batchnorm._BatchNorm, torch.nn.GroupNorm)):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def _sd_np(module, prefix):
    return {prefix + k: v.numpy() for k, v in module.state_dict().items()}


def main():
    gridding, mlp, unet3d = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(1234)

    # 1. 3D-UNet (reference module, small configuration) ------------------------------------------------------
    net = unet3d.Abstract3DUNet(in_channels=16, out_channels=8, final_sigmoid=False, basic_module=unet3d.DoubleConv,
                                f_maps=8, layer_order="gcr", num_groups=8, num_levels=3, is_segmentation=False).eval()
    _randomize(net, 1)
    x = torch.randn(2, 16, 8, 8, 8)
    with torch.no_grad():
        y = net(x)
    np.savez_compressed(os.path.join(OUT, "unet3d_small.npz"), x=x.numpy(), y=y.numpy(), **_sd_np(net, "sd."))
    full = unet3d.Abstract3DUNet(in_channels=128, out_channels=128, final_sigmoid=False,
                                 basic_module=unet3d.DoubleConv, f_maps=32, layer_order="gcr", num_groups=8,
                                 num_levels=4, is_segmentation=False)
    meta = {"unet_full_param_count": sum(p.numel() for p in full.parameters()),
            "unet_full_keys": {k: list(v.shape) for k, v in full.state_dict().items()}}

    # 2. MLP (reference components/mlp.py) ------------------------------------------------------------------------
    m = mlp.MLP([7, 16, 5], batch_norm=True).eval()
    _randomize(m, 2)
    xm = torch.randn(3, 11, 7)
    with torch.no_grad():
        ym = m(xm)
    np.savez_compressed(os.path.join(OUT, "mlp_small.npz"), x=xm.numpy(), y=ym.numpy(), **_sd_np(m, "sd."))
    meta["mlp_keys"] = list(m.state_dict().keys())

    # 3. VirtualGrid / ArraySlicer known answers -------------------------------------------------------------------
    vg64 = gridding.VirtualGrid(grid_shape=(64,) * 3, batch_size=1)
    bins = torch.stack([torch.arange(64)] * 3, dim=1)
    pts = vg64.idxs_to_points(bins)
    vg32 = gridding.VirtualGrid(grid_shape=(32,) * 3, batch_size=4)
    cell = vg32.get_points_grid_idxs(pts)
    rnd = torch.rand(200, 3) * 1.4 - 0.2
    bidx = torch.randint(0, 4, (200,))
    cell_rnd = vg32.get_points_grid_idxs(rnd, batch_idx=bidx)
    flat_rnd = vg32.flatten_idxs(cell_rnd)
    origin_rnd = vg32.idxs_to_points(cell_rnd)
    vg_odd = gridding.VirtualGrid(lower_corner=(-1, 0, 0.5), upper_corner=(1, 2, 1.5), grid_shape=(5, 6, 7), batch_size=2)
    cell_odd = vg_odd.get_points_grid_idxs(rnd)
    pts_odd = vg_odd.idxs_to_points(cell_odd)
    gp9 = gridding.VirtualGrid(grid_shape=(9,) * 3).get_grid_points(include_batch=False)
    gp128 = gridding.VirtualGrid(grid_shape=(128,) * 3).get_grid_points(include_batch=False)
    slicer = gridding.ArraySlicer((128, 128, 128, 3), (64, 64, 64))
    slices = [[(int(s.start), int(s.stop)) for s in sl] for sl in slicer]
    slicer_odd = gridding.ArraySlicer((10, 7, 3), (4, 7))
    slices_odd = [[(int(s.start), int(s.stop)) for s in slicer_odd[i]] for i in range(len(slicer_odd))]
    np.savez_compressed(os.path.join(OUT, "virtual_grid.npz"), bins=bins.numpy(), pts=pts.numpy(), cell=cell.numpy(),
                        rnd=rnd.numpy(), bidx=bidx.numpy(), cell_rnd=cell_rnd.numpy(), flat_rnd=flat_rnd.numpy(),
                        origin_rnd=origin_rnd.numpy(), cell_odd=cell_odd.numpy(), pts_odd=pts_odd.numpy(),
                        gp9=gp9.numpy(), gp128_123=gp128[1, 2, 3].numpy(), gp128_last=gp128[127, 127, 127].numpy(),
                        flat_kat=vg32.flatten_idxs(torch.tensor([[1, 2, 3, 4]])).numpy(),
                        num_grids=np.int64(vg32.num_grids))
    meta["array_slicer_128_64"] = slices
    meta["array_slicer_odd"] = slices_odd
    meta["array_slicer_len"] = len(slicer)

    # 4. implicit decoder = F.grid_sample (un-flipped) + reference MLP, restating conv_implicit_wnf.py:128-149 with
    #    the reference's own MLP module (the LightningModule wrapper itself needs pytorch_lightning, not installed)
    dec = mlp.MLP([6, 12, 12, 2], batch_norm=True).eval()
    _randomize(dec, 3)
    fg = torch.randn(2, 6, 4, 5, 6)
    q = torch.rand(2, 50, 3) * 1.2 - 0.1
    qn = 2.0 * q - 1.0
    with torch.no_grad():
        sf = F.grid_sample(input=fg, grid=qn.view(*(qn.shape[:2] + (1, 1, 3))), mode="bilinear", padding_mode="border",
                           align_corners=True)
        sf = sf.view(sf.shape[:3]).permute(0, 2, 1)
        yd = dec(sf)
        ys = gridding.nocs_grid_sample(fg, q)  # the flipped (zyx) variant, components/gridding.py:45-98
    np.savez_compressed(os.path.join(OUT, "decoder_small.npz"), fg=fg.numpy(), q=q.numpy(), y=yd.numpy(),
                        sampled=sf.numpy(), nocs_sampled=ys.numpy(), **_sd_np(dec, "sd.mlp."))

    # 5. gaussian gradient magnitude taps (scipy present; reference predict.py:162-163)
    import scipy.ndimage as ni
    imp = np.zeros((9, 9, 9), np.float32)
    imp[4, 4, 4] = 1
    ggm = ni.gaussian_gradient_magnitude(imp, sigma=0.5, mode="nearest")
    vol = np.random.default_rng(0).normal(size=(12, 13, 14)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "ggm.npz"), impulse=ggm, vol=vol,
                        vol_ggm=ni.gaussian_gradient_magnitude(vol, sigma=0.5, mode="nearest"))

    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
