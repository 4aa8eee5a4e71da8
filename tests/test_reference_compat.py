"""The reference's UNCHANGED networks/*.py import and construct on top of garmentnets_b200 through the shims, with
parameter names identical to our pipeline's (so reference checkpoints load).  Needs /root/reference (build container
only; skipped on the GPU box, where the reference does not exist)."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "networks")), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref_networks():
    from garmentnets_b200 import shims
    shims.install()
    saved = list(sys.path)
    sys.path.insert(0, REF)
    try:
        for name in [m for m in sys.modules if m.split(".")[0] in ("networks", "common")]:
            del sys.modules[name]
        import networks.conv_implicit_wnf as ciw
        assert ciw.__file__.startswith(REF)
        yield ciw
    finally:
        sys.path[:] = saved


def test_reference_pipeline_constructs_on_our_components(ref_networks):
    from garmentnets_b200 import synthetic
    from garmentnets_b200.components import mlp as our_mlp, unet3d as our_unet
    from garmentnets_b200.pipeline import ConvImplicitWNFPipeline
    hp = synthetic.HPARAMS
    ref = ref_networks.ConvImplicitWNFPipeline(
        pointnet2_params=hp["pointnet2"], volume_agg_params=hp["volume_agg"], unet3d_params=hp["unet3d"],
        volume_decoder_params=hp["volume_decoder"], surface_decoder_params=hp["surface_decoder"])
    ours = ConvImplicitWNFPipeline.from_hparams(hp)
    ref_sd, our_sd = ref.state_dict(), ours.state_dict()
    assert list(ref_sd.keys()) == list(our_sd.keys())
    assert all(ref_sd[k].shape == our_sd[k].shape for k in ref_sd)
    # the reference modules are built from OUR components (this is the drop-in boundary)
    assert isinstance(ref.unet_3d.abstract_3d_unet, our_unet.Abstract3DUNet)
    assert isinstance(ref.volume_decoder.mlp, our_mlp.FusedMLP)
    assert type(ref.pointnet2_nocs.sa1_module).__module__ == "garmentnets_b200.components.pointnet2"
    # and a reference-layout checkpoint loads into our pipeline key for key
    ours.load_state_dict(ref_sd, strict=True)
    assert ref.hparams["volume_loss_weight"] == 1.0 and ref.pointnet2_nocs.nocs_bins == 64


def test_reference_forward_refuses_cpu_tensors(ref_networks):
    """No silent CPU fallback: the reference forward on CPU tensors must fail loudly inside our components."""
    import torch
    from garmentnets_b200 import GarmentNetsB200Error, synthetic
    from garmentnets_b200.pipeline import Batch
    hp = synthetic.HPARAMS
    ref = ref_networks.ConvImplicitWNFPipeline(
        pointnet2_params=hp["pointnet2"], volume_agg_params=hp["volume_agg"], unet3d_params=hp["unet3d"],
        volume_decoder_params=hp["volume_decoder"], surface_decoder_params=hp["surface_decoder"]).eval()
    d = synthetic.make_batch(1, 256)
    data = Batch(x=torch.from_numpy(d["x"]), pos=torch.from_numpy(d["pos"]), batch=torch.from_numpy(d["batch"]))
    with pytest.raises((GarmentNetsB200Error, RuntimeError, AssertionError)):
        ref.pointnet2_forward(data)
