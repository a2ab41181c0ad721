"""Import shims so the UNMODIFIED reference modules under ``/root/reference`` import in this
container: ``segmentation_models_pytorch`` is missing and is only needed because
``vision_mtl/utils/model_utils.py:3,6`` imports it at module scope.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` and by tests that skip when the
reference tree is absent (it does not exist on the GPU box).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VMTL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vision_mtl"))


def install() -> None:
    """Put the reference on sys.path behind a stub for segmentation_models_pytorch."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "segmentation_models_pytorch" not in sys.modules:
        import torch.nn as nn

        smp = types.ModuleType("segmentation_models_pytorch")
        base = types.ModuleType("segmentation_models_pytorch.base")

        class SegmentationHead(nn.Sequential):  # never instantiated by the golden generator
            pass

        base.SegmentationHead = SegmentationHead
        smp.base = base
        smp.Unet = None
        sys.modules["segmentation_models_pytorch"] = smp
        sys.modules["segmentation_models_pytorch.base"] = base
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Returns the reference classes the hot path is pinned against."""
    install()
    from vision_mtl.losses import SILogLoss
    from vision_mtl.models.cross_stitch_model import CrossStitchLayer, CSNet
    from vision_mtl.models.mtan_model import MTANMiniUnet
    from vision_mtl.utils import loss_utils

    return {"SILogLoss": SILogLoss, "CrossStitchLayer": CrossStitchLayer, "CSNet": CSNet,
            "MTANMiniUnet": MTANMiniUnet, "loss_utils": loss_utils}
