"""Hard-parameter-sharing baseline, drop-in for ``vision_mtl/models/basic_model.py``: one
backbone, two 3x3 ``SegmentationHead``s (basic_model.py:30-41).  The heads stay on cuDNN; the
step module feeds their logits to the fused CE / SILog kernels."""
from __future__ import annotations

import typing as t

import torch
import torch.nn as nn

from ..utils.model_utils import Backbone, SegmentationHead


class BasicMTLModel(nn.Module):
    def __init__(
        self,
        segm_classes: int,
        activation: t.Any = None,
        encoder_name: str = "timm-mobilenetv3_large_100",
        encoder_weights: t.Optional[str] = "imagenet",
        decoder_first_channel: int = 256,
        num_decoder_layers: int = 5,
        in_channels: int = 3,
    ):
        super().__init__()
        self.backbone = Backbone(
            encoder_name=encoder_name,
            encoder_weights=encoder_weights,
            decoder_first_channel=decoder_first_channel,
            num_decoder_layers=num_decoder_layers,
            in_channels=in_channels,
        )
        width = self.backbone.decoder_channels[-1]
        self.segm_head = SegmentationHead(width, segm_classes, kernel_size=3, activation=activation)
        self.depth_head = SegmentationHead(width, 1, kernel_size=3, activation=activation)

    def forward(self, x: torch.Tensor) -> t.Dict[str, torch.Tensor]:
        shared = self.backbone(x)
        return {"depth": self.depth_head(shared), "segm": self.segm_head(shared)}

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> t.Dict[str, torch.Tensor]:
        if self.training:
            self.eval()
        return self.forward(x)
