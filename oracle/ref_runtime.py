"""Runs the reference's OWN training loop (``vision_mtl/training_lit.py:run_pipe`` driving
``vision_mtl/lit_module.py:MTLModule``) on the host cores: the CPU arm of ``bench.py``.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference tree is taken from ``/root/reference`` (build container) or from the git-ignored
``baseline/_ref`` copy made by ``oracle/install_reference.py`` (the GPU box).  Its modules are imported
UNMODIFIED; what is missing from this image is supplied as import stubs that contribute no arithmetic,
with two stated exceptions:

* ``torchmetrics`` (0.7.3 pinned by the reference, not installable): the four metric classes are stubs
  whose ``forward`` computes the restated 0.7.3 definitions of ``oracle/metrics_np.py`` from one confusion
  matrix.  That is CHEAPER than the real package (which one-hot expands predictions and targets, SURVEY
  2.1), so the CPU baseline is, if anything, flattered.
* ``torch.optim.lr_scheduler.ReduceLROnPlateau``: torch >= 2.7 rejects the ``verbose`` argument the
  reference passes; the shim drops it.
* ``segmentation_models_pytorch`` / ``timm`` are absent: csnet task networks are the stand-in backbone
  (``vision_mtl_b200/utils/standin_backbone.py``), as everywhere else in this repository.
"""
from __future__ import annotations

import argparse
import inspect
import os
import sys
import time
import types

import numpy as np
import torch
import torch.nn as nn

from . import metrics_np as MN

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root() -> str:
    for cand in (os.environ.get("VMTL_REFERENCE_ROOT"), "/root/reference", os.path.join(REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "vision_mtl", "models")):
            return cand
    raise RuntimeError("reference tree not found (neither /root/reference nor baseline/_ref; run "
                       "`python -m oracle.install_reference` in the build container)")


def _module(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


class _LightningModule(nn.Module):
    """The slice of pl.LightningModule the reference touches under its manual loop (no Trainer)."""

    def save_hyperparameters(self, ignore=()):
        frame = inspect.currentframe().f_back
        args = {k: v for k, v in frame.f_locals.items() if k not in ("self", "__class__", *ignore)}
        self.hparams = _AttrDict(args)

    def log(self, *a, **k):
        return None

    def log_dict(self, *a, **k):
        return None


class _SegMetric:
    """torchmetrics 0.7.3 semantics restated: forward() returns the batch-local value."""

    fn = None

    def __init__(self, num_classes=None, **_):
        self.num_classes = num_classes

    def to(self, *_a, **_k):
        return self

    def __call__(self, preds, target):
        cm = MN.confusion_matrix(preds.detach().cpu().numpy(), target.detach().cpu().numpy(), self.num_classes)
        return torch.tensor(type(self).fn(cm), dtype=torch.float32)


class _Accuracy(_SegMetric):
    fn = staticmethod(MN.accuracy_micro)


class _FBeta(_SegMetric):
    fn = staticmethod(MN.fbeta_weighted)


class _Jaccard(_SegMetric):
    fn = staticmethod(MN.jaccard_macro_absent0)


class _MAE:
    def to(self, *_a, **_k):
        return self

    def __call__(self, preds, target):
        return (preds.detach() - target).abs().mean()


class _Noop:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return a[0] if a else None


def install_stubs() -> str:
    root = reference_root()
    smp = _module("segmentation_models_pytorch", Unet=None)
    smp.base = _module("segmentation_models_pytorch.base", SegmentationHead=type("SegmentationHead", (nn.Sequential,), {}))
    pl = _module("pytorch_lightning", LightningModule=_LightningModule, LightningDataModule=object)
    pl.loggers = _module("pytorch_lightning.loggers", TensorBoardLogger=_Noop)
    _module("torchmetrics", Accuracy=_Accuracy, FBetaScore=_FBeta, JaccardIndex=_Jaccard, MeanAbsoluteError=_MAE)
    alb = _module("albumentations", Compose=_Noop, Resize=_Noop)
    alb.pytorch = _module("albumentations.pytorch", ToTensorV2=_Noop)
    _module("omegaconf", MISSING="???")
    comet = _module("comet_ml", Experiment=_Noop, ExistingExperiment=_Noop, API=_Noop)
    comet.api = _module("comet_ml.api", API=_Noop)
    try:
        import matplotlib.patches  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _module("matplotlib", use=lambda *a, **k: None)
        mpl.__path__ = []  # a package: `matplotlib.patches` resolves to the stub below
        mpl.pyplot = _module("matplotlib.pyplot", show=lambda *a, **k: None, close=lambda *a, **k: None, Figure=object, Axes=object)
        mpl.patches = _module("matplotlib.patches", Rectangle=_Noop)
    for name in ("h5py", "optuna", "cv2"):
        try:
            __import__(name)
        except Exception:
            _module(name)
    try:
        from PIL import Image  # noqa: F401
    except Exception:
        pil = _module("PIL")
        pil.__path__ = []
        pil.Image = _module("PIL.Image")
    try:
        import torchvision  # noqa: F401  (cfg.py builds torchvision transforms at import)
    except Exception:
        tv = _module("torchvision")
        tv.transforms = _module("torchvision.transforms", Compose=_Noop, ToTensor=_Noop, Resize=_Noop)
    try:
        import dotenv  # noqa: F401
    except Exception:
        _module("dotenv", load_dotenv=lambda *a, **k: None)
    sched = torch.optim.lr_scheduler
    if "verbose" not in inspect.signature(sched.ReduceLROnPlateau.__init__).parameters:
        base = sched.ReduceLROnPlateau

        class ReduceLROnPlateau(base):  # torch >= 2.7 dropped `verbose`
            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        sched.ReduceLROnPlateau = ReduceLROnPlateau
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def load() -> dict:
    root = install_stubs()
    from vision_mtl import training_lit
    from vision_mtl.lit_module import MTLModule
    from vision_mtl.models.cross_stitch_model import CSNet
    from vision_mtl.models.mtan_model import MTANMiniUnet

    return {"root": root, "run_pipe": training_lit.run_pipe, "MTLModule": MTLModule, "CSNet": CSNet,
            "MTANMiniUnet": MTANMiniUnet}


class _TimedBatches:
    """The datamodule surface run_pipe touches; stamps the time every batch is handed out."""

    benchmark_batch = None

    def __init__(self, batch: dict, n: int):
        self.batch, self.n, self.stamps = batch, n, []

    def train_dataloader(self):
        for _ in range(self.n):
            self.stamps.append(time.perf_counter())
            yield {k: v.clone() for k, v in self.batch.items()}
        self.stamps.append(time.perf_counter())

    def val_dataloader(self):
        return iter(())


class _NullLogger:
    log_dir = "."

    def log_metrics(self, *a, **k):
        return None


def build_module(ref: dict, model: str, num_classes: int, lr: float):
    """Models as ``utils/pipeline_utils.py:build_model`` configures them (:99-136)."""
    if model == "mtan":
        net = ref["MTANMiniUnet"](in_channels=3, map_tasks_to_num_channels={"depth": 1, "segm": num_classes},
                                  task_subnets_hidden_channels=128, encoder_first_channel=32, encoder_num_channels=4)
    elif model == "csnet":
        from vision_mtl_b200.utils.model_utils import get_model_with_dense_preds  # stand-in backbone (no smp/timm)

        nets = {"depth": get_model_with_dense_preds(1, None, dict(encoder_weights=None)),
                "segm": get_model_with_dense_preds(num_classes, None, dict(encoder_weights=None))}
        net = ref["CSNet"](nets, channel_wise_stitching=True)
    else:
        raise ValueError(model)
    return ref["MTLModule"](model=net, num_classes=num_classes, lr=lr, device="cpu")


def time_reference_loop(model: str, batch: dict, num_classes: int, lr: float, steps: int, warmup: int,
                        quiet: bool = True):
    """(images/s, seconds/step, threads) of the reference's run_pipe train loop on ``batch``."""
    import contextlib
    import io

    ref = load()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(11)
    module = build_module(ref, model, num_classes, lr)
    dm = _TimedBatches(batch, warmup + steps)
    args = argparse.Namespace(lr=lr, val_epoch_freq=10**9, save_epoch_freq=10**9, num_epochs=10**9, do_show_preds=False)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink if quiet else sys.stdout), contextlib.redirect_stderr(sink if quiet else sys.stderr):
        hist = ref["run_pipe"](args, module, dm, 1, "cpu", None, _NullLogger())
    st = dm.stamps
    dt = st[-1] - st[warmup]
    bs = batch["img"].shape[0]
    loss = hist["train"]["train/loss"][0]
    if not np.isfinite(loss):
        raise RuntimeError("reference loop produced a non-finite loss")
    return bs * steps / dt, dt / steps, threads
