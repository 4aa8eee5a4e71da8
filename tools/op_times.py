"""Device time of individual hot-path ops at the benchmark sizes (CUDA events, best of 5).  Development aid:
    python tools/op_times.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from garmentnets_b200 import ops, synthetic
from garmentnets_b200.components.pointnet2 import CloudIndex
from garmentnets_b200.pipeline import Batch

dev = torch.device("cuda:0")


def timeit(name, fn, n=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"{name:50s} {best:9.3f} ms")
    return best


B, N = 32, 4096
d = synthetic.make_batch(B, N, "Tshirt", 1)
model = synthetic.build_pipeline(0, dev)
data = Batch(x=torch.from_numpy(d["x"]).to(dev), pos=torch.from_numpy(d["pos"]).to(dev), batch=torch.from_numpy(d["batch"]).to(dev))
index = CloudIndex.uniform(B, N, dev)
synthetic.prepare_model_(model, data, index)
pn = model.pointnet2_nocs

sub = index.subsample(0.5)
timeit("fps SA1 (32 x 4096 -> 2048)", lambda: ops.fps(data.pos, index.ptr, sub.ptr, N, sub.total, None))
idx = ops.fps(data.pos, index.ptr, sub.ptr, N, sub.total, None)
pos1 = data.pos[idx]
timeit("ball query SA1", lambda: ops.ball_query(data.pos, pos1, index.ptr, sub.ptr, 0.05, 64))
nbr, cnt = ops.ball_query(data.pos, pos1, index.ptr, sub.ptr, 0.05, 64)
print("   mean neighbours SA1:", float(cnt.float().mean()), "max", int(cnt.max()))
timeit("SA1 pointconv (gather + 3 linear + segmax)", lambda: pn.sa1_module.conv.forward_grouped(data.x, data.pos, pos1, nbr, cnt))
x1 = pn.sa1_module.conv.forward_grouped(data.x, data.pos, pos1, nbr, cnt)
sub2 = sub.subsample(0.25)
timeit("fps SA2 (32 x 2048 -> 512)", lambda: ops.fps(pos1, sub.ptr, sub2.ptr, 2048, sub2.total, None))
idx2 = ops.fps(pos1, sub.ptr, sub2.ptr, 2048, sub2.total, None)
pos2 = pos1[idx2]
nbr2, cnt2 = ops.ball_query(pos1, pos2, sub.ptr, sub2.ptr, 0.1, 64)
print("   mean neighbours SA2:", float(cnt2.float().mean()))
timeit("SA2 pointconv", lambda: pn.sa2_module.conv.forward_grouped(x1, pos1, pos2, nbr2, cnt2))
timeit("pointnet2_forward total", lambda: model.pointnet2_forward(data, index=index))
p = model.pointnet2_forward(data, index=index)
timeit("aggregator", lambda: model.volume_agg(p["nocs_data"]))
vol_in = model.volume_agg(p["nocs_data"])
timeit("unet3d", lambda: model.unet_3d(vol_in))
fv = model.unet_3d(vol_in)
x_cl = ops.to_channels_last(vol_in)
enc0 = model.unet_3d.abstract_3d_unet.encoders[0].basic_module.SingleConv1
timeit("  E0.c1 groupnorm_stats", lambda: ops.groupnorm_stats(x_cl, 8, 1e-5, enc0.groupnorm.weight, enc0.groupnorm.bias))
sc, sh = ops.groupnorm_stats(x_cl, 8, 1e-5, enc0.groupnorm.weight, enc0.groupnorm.bias)
timeit("  E0.c1 gn_apply_split", lambda: ops.gn_apply_split(x_cl, sc, sh))
xh, xl = ops.gn_apply_split(x_cl, sc, sh)
wp = enc0.packed_weight_tc()
timeit("  E0.c1 conv3d_tc 128->128 @32^3 x32", lambda: ops.conv3d_tc(xh, xl, 128, wp, 128, True))
timeit("dense_decode (tcgen05, 32 x 128^3)", lambda: model.dense_decode(fv, 128))
wnf = model.dense_decode(fv, 128)
timeit("ggm batched (32 x 128^3)", lambda: ops.gaussian_gradient_magnitude_batched(wnf, 0.5))
ggm = ops.gaussian_gradient_magnitude_batched(wnf, 0.5)
timeit("marching_cubes_batch (32)", lambda: ops.marching_cubes_batch(wnf, 0.5, (1 / 127,) * 3, "ascent", ggm))
lib_ws = torch.empty(int(__import__("garmentnets_b200")._lib.load().gnb_mc_workspace_bytes(128, 128, 128)), dtype=torch.uint8, device=dev)
from garmentnets_b200 import _lib
timeit("  mc classify+scan (1 volume, async)", lambda: _lib.call("gnb_mc_count", wnf[0].data_ptr(), 128, 128, 128, 0.5, lib_ws.data_ptr(), None, torch.cuda.current_stream().cuda_stream))
mcs = ops.marching_cubes_batch(wnf, 0.5, (1 / 127,) * 3, "ascent", ggm)
verts = mcs[0][0]
print("   verts/faces sample 0:", len(verts), len(mcs[0][1]))
dec = model.surface_decoder
fvol = ops.to_channels_last(fv)
timeit("surface hoisted linear (32 x 32^3 x 128->256)", lambda: dec.hoisted(fvol))
u = dec.hoisted(fvol)
timeit("  surface decode 1 sample (interp + tc tail)", lambda: dec.forward_hoisted(u[:1], verts.view(1, -1, 3)))
sc1, sh1 = dec.mlp[0][2].folded_affine()
timeit("    trilinear_sample 1 sample", lambda: ops.trilinear_sample(u[:1], verts.view(1, -1, 3), False, sc1, sh1))
