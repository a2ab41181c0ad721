"""Model-building helpers mirroring ``vision_mtl/utils/model_utils.py`` of the reference.

Everything here is glue around stock cuDNN/ATen ops (3x3 convs, BN, ReLU); none of it is on
the hand-written hot path.  ``Backbone`` is the smp ``Unet`` encoder/decoder pair when
``segmentation_models_pytorch`` is importable and otherwise the stand-in from
``standin_backbone.py`` that honours the module-naming contract ``CSNet`` relies on
(SURVEY F9 / Appendix D).
"""
from __future__ import annotations

import typing as t

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .standin_backbone import StandinUnet

try:  # pragma: no cover - not installable offline
    import segmentation_models_pytorch as smp  # type: ignore
    from segmentation_models_pytorch.base import SegmentationHead as _SmpHead  # type: ignore
except Exception:  # ImportError or a partially stubbed module
    smp = None
    _SmpHead = None


class _Activation(nn.Module):
    """smp's ``Activation(None)``: a wrapper whose only child is ``activation``."""

    def __init__(self, activation=None):
        super().__init__()
        if activation is None or activation == "identity":
            self.activation = nn.Identity()
        elif activation == "sigmoid":
            self.activation = nn.Sigmoid()
        elif callable(activation):
            self.activation = activation()
        else:
            raise ValueError(f"unsupported activation {activation!r}")

    def forward(self, x):
        return self.activation(x)


class SegmentationHead(nn.Sequential):
    """Same layout as smp 0.3.3 ``SegmentationHead``: ``0`` conv, ``1`` upsampling, ``2`` activation
    (reference call sites: basic_model.py:30-41, model_utils.py:125-130)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, activation=None, upsampling=1):
        conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=kernel_size // 2)
        up = nn.UpsamplingBilinear2d(scale_factor=upsampling) if upsampling > 1 else nn.Identity()
        super().__init__(conv, up, _Activation(activation))


class Backbone(nn.Module):
    """Encoder + Unet decoder (reference: model_utils.py:10-43)."""

    def __init__(
        self,
        encoder_name: str = "timm-mobilenetv3_large_100",
        encoder_weights: t.Optional[str] = "imagenet",
        decoder_first_channel: int = 256,
        num_decoder_layers: int = 5,
        in_channels: int = 3,
    ):
        super().__init__()
        self.decoder_channels = [decoder_first_channel // (2**i) for i in range(num_decoder_layers)]
        if smp is not None and getattr(smp, "Unet", None) is not None:  # pragma: no cover
            net = smp.Unet(
                encoder_name=encoder_name,
                encoder_weights=encoder_weights,
                in_channels=in_channels,
                encoder_depth=len(self.decoder_channels),
                decoder_channels=self.decoder_channels,
            )
        else:
            # offline: random-init stand-in with the same module names and channel plan
            net = StandinUnet(in_channels=in_channels, decoder_channels=self.decoder_channels)
        self.encoder = net.encoder
        self.decoder = net.decoder

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.decoder(*self.encoder(x))


def concat_slightly_diff_sized_tensors(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """Zero-pad ``x1`` (centred) to the spatial size of ``x2`` and return ``cat([x2, x1], 1)``
    (reference: model_utils.py:46-58)."""
    dh = x2.shape[2] - x1.shape[2]
    dw = x2.shape[3] - x1.shape[3]
    if dh or dw:
        x1 = F.pad(x1, [dw // 2, dw - dw // 2, dh // 2, dh - dh // 2])
    return torch.cat([x2, x1], dim=1)


class DoubleConv(nn.Module):
    """Two (3x3 conv, BN, ReLU) stages; child layout ``double_conv.{0,1,3,4}`` as in the reference
    (model_utils.py:61-80) so checkpoints load unchanged."""

    def __init__(self, in_channels: int, out_channels: int, mid_channels: t.Optional[int] = None):
        super().__init__()
        mid = mid_channels or out_channels
        stages = []
        for cin, cout in ((in_channels, mid), (mid, out_channels)):
            stages += [nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        self.double_conv = nn.Sequential(*stages)

    def forward(self, x):
        dc = self.double_conv
        for conv, bn, act in ((dc[0], dc[1], dc[2]), (dc[3], dc[4], dc[5])):
            x = conv(x)
            # BN (batch or running statistics) + ReLU as one statistics pass + one apply pass of the library;
            # shapes / devices the kernels do not cover keep the ATen modules
            x = ops.batch_norm_relu(x, bn) if ops.bn_supported(bn, x) else act(bn(x))
        return x


def _is_stitch_level(parts: t.List[str]) -> bool:
    """Encoder block containers are 5-part names (``0.encoder.model.blocks.N``), decoder block
    containers 4-part names (``0.decoder.blocks.N``) -- model_utils.py:83-98."""
    return ("encoder" in parts and len(parts) == 5) or ("decoder" in parts and len(parts) == 4)


def get_joint_layer_names(all_layer_names: t.List[str]) -> t.List[str]:
    return [n for n in all_layer_names if _is_stitch_level(n.split("."))]


def get_joint_layer_names_before_stitch_for_unet(joint_layer_names: t.List[str]) -> t.List[str]:
    """Names after whose (container) visit a cross-stitch unit is applied: every decoder block and
    every encoder block except block 0 (model_utils.py:100-115)."""
    out = []
    for name in joint_layer_names:
        parts = name.split(".")
        if "encoder" in parts and len(parts) == 5:
            if int(parts[-1]) != 0:
                out.append(name)
        elif "decoder" in parts and len(parts) == 4:
            out.append(name)
    return out


def get_model_with_dense_preds(
    segm_classes: int = 10, activation: t.Any = None, backbone_params: t.Optional[dict] = None
) -> nn.Module:
    """``Sequential(Backbone, SegmentationHead)`` -> module names prefixed ``0.`` / ``1.``
    (model_utils.py:118-132)."""
    backbone = Backbone(in_channels=3, **(backbone_params or {}))
    head = SegmentationHead(backbone.decoder_channels[-1], segm_classes, kernel_size=3, activation=activation)
    return nn.Sequential(backbone, head)
