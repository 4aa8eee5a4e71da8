"""Import shims that let the reference's UNCHANGED ``networks/conv_implicit_wnf.py`` / ``networks/pointnet2_nocs.py``
run on top of ``garmentnets_b200`` ("extra names L4/L5 import directly").

``install()`` registers, in ``sys.modules``:

* ``components`` (+ ``.mlp .pointnet2 .gridding .unet3d .loss .symmetry``) -> ``garmentnets_b200.components``;
* ``torch_scatter.scatter``                     -> ``gnb_scatter_reduce``  (ref networks/conv_implicit_wnf.py:10,92);
* ``torch_geometric.data.Batch`` / ``.nn.*``    -> ``garmentnets_b200.pipeline.Batch`` / the PyG-named functions of
  ``components.pointnet2``                         (ref networks/conv_implicit_wnf.py:11,232);
* ``pytorch_lightning.LightningModule``         -> a thin ``nn.Module`` subclass with ``save_hyperparameters``,
  ``hparams``, ``device``, ``log`` and ``load_from_checkpoint`` (ref networks/conv_implicit_wnf.py:23,34; predict.py:101);
* ``matplotlib.cm`` / ``skimage.transform`` / ``skimage.measure`` stand-ins for the visualisation imports executed at
  module load (ref common/rendering_util.py:2-4) -- only if the real packages are absent.  ``skimage.measure.
  marching_cubes`` is served by the device kernel.

Only modules that are NOT importable are shimmed; a real installation of any of them is left alone (except
``components``, which is the point of the exercise).
"""
from __future__ import annotations

import importlib
import sys
import types


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__garmentnets_b200_shim__ = True
    sys.modules[name] = m
    return m


def _not_on_hot_path(what):
    def f(*a, **k):
        raise NotImplementedError(f"{what} is outside the GarmentNets inference hot path (garmentnets_b200 shim)")
    return f


def install(force: bool = False) -> None:
    import torch
    from torch import nn

    from .. import components as C
    from ..components import gridding, mlp, pointnet2, unet3d
    from ..pipeline import Batch

    # ---- components -------------------------------------------------------------------------------------------
    sys.modules["components"] = C
    for sub in (gridding, mlp, pointnet2, unet3d):
        sys.modules["components." + sub.__name__.rsplit(".", 1)[1]] = sub

    class MirrorMSELoss(nn.Module):  # training-only (ref components/loss.py); constructible, not runnable
        forward = _not_on_hot_path("MirrorMSELoss")

    _module("components.loss", MirrorMSELoss=MirrorMSELoss)
    _module("components.symmetry", mirror_nocs_points_by_axis=_not_on_hot_path("mirror_nocs_points_by_axis"))

    # ---- torch_scatter ------------------------------------------------------------------------------------------
    if force or _missing("torch_scatter"):
        _module("torch_scatter", scatter=gridding.scatter)

    # ---- torch_geometric ----------------------------------------------------------------------------------------
    if force or _missing("torch_geometric"):
        tg = _module("torch_geometric")
        tg.data = _module("torch_geometric.data", Batch=Batch, Data=Batch, DataLoader=_not_on_hot_path("DataLoader"))
        tg.nn = _module("torch_geometric.nn", PointConv=pointnet2.PointConv, fps=pointnet2.fps, radius=pointnet2.radius,
                        global_max_pool=pointnet2.global_max_pool, knn_interpolate=pointnet2.knn_interpolate)
        tg.datasets = _module("torch_geometric.datasets", ModelNet=_not_on_hot_path("ModelNet"))
        tg.transforms = _module("torch_geometric.transforms")

    # ---- pytorch_lightning ----------------------------------------------------------------------------------------
    if force or _missing("pytorch_lightning"):
        import inspect

        class _HParams(dict):
            __getattr__ = dict.get

        class LightningModule(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
                self._hparams = _HParams()
                self.logger = None
                self.global_step = 0

            def save_hyperparameters(self, *a, **k):
                frame = inspect.currentframe().f_back
                args = inspect.getargvalues(frame)
                hp = {n: args.locals[n] for n in args.args if n != "self"}
                if args.keywords:
                    hp.update(args.locals.get(args.keywords, {}))
                self._hparams = _HParams(hp)

            @property
            def hparams(self):
                return self._hparams

            @property
            def device(self):
                p = next(self.parameters(), None)
                return p.device if p is not None else torch.device("cpu")

            def log(self, *a, **k):
                pass

            @classmethod
            def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
                ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
                hp = dict(ckpt.get("hyper_parameters", {}))
                hp.update(kwargs)
                model = cls(**hp)
                model.load_state_dict(ckpt["state_dict"], strict=strict)
                return model

        pl = _module("pytorch_lightning", LightningModule=LightningModule, Trainer=_not_on_hot_path("Trainer"))
        pl.callbacks = _module("pytorch_lightning.callbacks", ModelCheckpoint=_not_on_hot_path("ModelCheckpoint"))
        pl.loggers = _module("pytorch_lightning.loggers", WandbLogger=_not_on_hot_path("WandbLogger"))

    # ---- visualisation imports executed at module load ------------------------------------------------------------
    if force or _missing("matplotlib"):
        mpl = _module("matplotlib")
        mpl.cm = _module("matplotlib.cm", get_cmap=_not_on_hot_path("matplotlib.cm.get_cmap"))
        mpl.pyplot = _module("matplotlib.pyplot")
    if force or _missing("skimage"):
        from .. import ops

        def marching_cubes(volume, level=None, spacing=(1.0, 1.0, 1.0), gradient_direction="descent", step_size=1,
                           allow_degenerate=True, method="lewiner", mask=None):
            """``skimage.measure.marching_cubes`` call shape on the device kernel (ref predict.py:172-177).  Same MC33
            structure and conventions as scikit-image's Lewiner implementation, NOT its exact tables: vertex / face
            numbering can differ (parity unpinned).
            Accepts a CUDA tensor or a numpy array (copied to the current device); returns numpy arrays like skimage:
            float64 verts (float32 * float64 spacing), int32 faces, float32 normals / values."""
            import numpy as np
            if step_size != 1 or mask is not None or method not in ("lewiner", "lorensen"):
                raise NotImplementedError("marching_cubes shim: step_size=1, mask=None only")
            vol = volume if isinstance(volume, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(volume, np.float32)).cuda()
            if level is None:
                level = 0.5 * (float(vol.min()) + float(vol.max()))
            v, f, n, val, _ = ops.marching_cubes(vol, level, (1.0, 1.0, 1.0), gradient_direction)
            verts = v.cpu().numpy() * np.r_[spacing]
            return verts, f.cpu().numpy(), n.cpu().numpy(), val.cpu().numpy()

        sk = _module("skimage")
        sk.measure = _module("skimage.measure", marching_cubes=marching_cubes)
        sk.transform = _module("skimage.transform", resize=_not_on_hot_path("skimage.transform.resize"))
