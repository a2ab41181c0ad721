"""Stand-in for ``smp.Unet(encoder_name="timm-mobilenetv3_large_100")``.

segmentation-models-pytorch 0.3.3 and timm cannot be installed offline, so the ``basic`` and
``csnet`` models run on this random-init substitute.  It keeps exactly what the reference
relies on:

* module names -- ``encoder.model.{conv_stem,bn1,act1,blocks.0..6}`` and
  ``decoder.{center,blocks.N.{conv1.{0,1,2},attention1.attention,conv2.{0,1,2},attention2.attention}}``
  -- which is CSNet's only contract with the backbone (regexes at cross_stitch_model.py:48-49);
* the MobileNetV3-Large stage plan: output channels 16,24,40,80,112,160,960 with strides
  1,2,2,2,1,2,1 after a stride-2 stem, feature taps (3,16,24,40,112,960);
* the Unet decoder channel arithmetic (in = previous + skip).

It is NOT numerically the timm network (no squeeze-excite, plain BN + activation leaves);
backbone numerics are out of the hot path's scope and are stated as unpinned.
"""
from __future__ import annotations

import typing as t

import torch
import torch.nn.functional as F
from torch import nn


def _divisible(v: float, d: int = 8) -> int:
    return max(d, int(v + d / 2) // d * d)


class _SeparableBlock(nn.Module):
    """depthwise 3x3 -> pointwise (stage 0)."""

    def __init__(self, cin, cout, k, stride, act):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.act1 = act()
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = self.bn2(self.conv_pw(self.act1(self.bn1(self.conv_dw(x)))))
        return x + y if self.skip else y


class _InvertedResidual(nn.Module):
    """pointwise expand -> depthwise kxk -> pointwise project."""

    def __init__(self, cin, cout, k, stride, expand, act):
        super().__init__()
        mid = _divisible(cin * expand)
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.act1 = act()
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.act2 = act()
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_pw(x)))
        y = self.act2(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return x + y if self.skip else y


class _ConvBnAct(nn.Module):
    def __init__(self, cin, cout, act):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.act1 = act()

    def forward(self, x):
        return self.act1(self.bn1(self.conv(x)))


# (kind, kernel, stride, expand, out_channels, repeats, activation) per stage
_STAGES = [
    ("ds", 3, 1, 1.0, 16, 1, nn.ReLU),
    ("ir", 3, 2, (4.0, 3.0), 24, 2, nn.ReLU),
    ("ir", 5, 2, 3.0, 40, 3, nn.ReLU),
    ("ir", 3, 2, (6.0, 2.5, 2.3, 2.3), 80, 4, nn.Hardswish),
    ("ir", 3, 1, 6.0, 112, 2, nn.Hardswish),
    ("ir", 5, 2, 6.0, 160, 3, nn.Hardswish),
    ("cn", 1, 1, 1.0, 960, 1, nn.Hardswish),
]
_TAPS = (0, 1, 2, 4, 6)


class _MobileNetV3LargeFeatures(nn.Module):
    def __init__(self, in_channels: int = 3):
        super().__init__()
        self.conv_stem = nn.Conv2d(in_channels, 16, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.act1 = nn.Hardswish()
        stages, cin = [], 16
        for kind, k, stride, expand, cout, reps, act in _STAGES:
            blocks = []
            for r in range(reps):
                e = expand[r] if isinstance(expand, tuple) else expand
                s = stride if r == 0 else 1
                if kind == "ds":
                    blocks.append(_SeparableBlock(cin, cout, k, s, act))
                elif kind == "ir":
                    blocks.append(_InvertedResidual(cin, cout, k, s, e, act))
                else:
                    blocks.append(_ConvBnAct(cin, cout, act))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)

    def forward(self, x) -> t.List[torch.Tensor]:
        x = self.act1(self.bn1(self.conv_stem(x)))
        feats = []
        for i, stage in enumerate(self.blocks):
            x = stage(x)
            if i in _TAPS:
                feats.append(x)
        return feats


class StandinEncoder(nn.Module):
    out_channels = (3, 16, 24, 40, 112, 960)

    def __init__(self, in_channels: int = 3):
        super().__init__()
        self.model = _MobileNetV3LargeFeatures(in_channels)

    def forward(self, x):
        return [x] + self.model(x)


class _NoAttention(nn.Module):
    def __init__(self):
        super().__init__()
        self.attention = nn.Identity()

    def forward(self, x):
        return self.attention(x)


def _conv_bn_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _conv_bn_relu(cin + cskip, cout)
        self.attention1 = _NoAttention()
        self.conv2 = _conv_bn_relu(cout, cout)
        self.attention2 = _NoAttention()

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = self.attention1(torch.cat([x, skip], dim=1))
        return self.attention2(self.conv2(self.conv1(x)))


class StandinDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]  # 960, 112, 40, 24, 16
        ins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList(_DecoderBlock(i, s, o) for i, s, o in zip(ins, skips, decoder_channels))

    def forward(self, *features):
        feats = features[1:][::-1]
        x = self.center(feats[0])
        skips = feats[1:]
        for i, block in enumerate(self.blocks):
            x = block(x, skips[i] if i < len(skips) else None)
        return x


class StandinUnet(nn.Module):
    def __init__(self, in_channels: int = 3, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        self.encoder = StandinEncoder(in_channels)
        self.decoder = StandinDecoder(self.encoder.out_channels, list(decoder_channels))
