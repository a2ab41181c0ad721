"""MTAN mini-UNet, drop-in for ``vision_mtl/models/mtan_model.py``.

Class names, constructor arguments, sub-module names (hence ``state_dict`` keys)
and outputs follow the reference.  What changes is the attention gate -- the reference's
``conv2 -> bn2 -> sigmoid -> * shared`` chain (mtan_model.py:71-75, :158-162), five ATen
kernels and four materialised intermediates per site -- which runs here as the fused
tensor-core kernels behind ``ops.attention_gate`` in NHWC.  Everything around it (3x3 convs,
BN/ReLU, pooling, up-sampling) stays on cuDNN/ATen in ``channels_last``.

``MTANMiniUnet.forward_features`` returns the per-task features in front of the 1x1 heads so
the step module can fuse the heads with their losses; ``forward`` keeps returning logits.
"""
from __future__ import annotations

import typing as t

import torch
from torch import nn

from .. import ops
from ..utils.model_utils import DoubleConv, concat_slightly_diff_sized_tensors


class _GatedAttention(nn.Module):
    """Shared pieces of the encoder / decoder attention modules: the 1x1 squeeze to the hidden
    width (``conv1, bn1, relu1``) and the fused gate over (``conv2, bn2, sigmoid``)."""

    gate_precision: t.Optional[str] = None  # None -> ops.default_gate_precision

    def _build_gate(self, in_channels: int, hidden: int, gated_channels: int) -> None:
        self.conv1 = nn.Conv2d(in_channels, hidden, kernel_size=1, padding=0)
        self.bn1 = nn.BatchNorm2d(hidden)
        self.relu1 = nn.ReLU()
        self.conv2 = nn.Conv2d(hidden, gated_channels, kernel_size=1, padding=0)
        self.bn2 = nn.BatchNorm2d(gated_channels)
        self.sigmoid = nn.Sigmoid()

    @staticmethod
    def _bn_relu(x: torch.Tensor, bn: nn.BatchNorm2d, act: nn.Module, pool: t.Optional[nn.Module] = None) -> torch.Tensor:
        """``[pool(] act(bn(x)) [)]`` through the fused BatchNorm kernels where they apply."""
        if ops.bn_supported(bn, x) and isinstance(act, nn.ReLU):
            return ops.batch_norm_relu(x, bn, relu=True, pool=pool is not None)
        y = act(bn(x))
        return y if pool is None else pool(y)

    @classmethod
    def _conv_bn_relu(cls, x: torch.Tensor, conv: nn.Conv2d, bn: nn.BatchNorm2d, act: nn.Module,
                      pool: t.Optional[nn.Module] = None) -> torch.Tensor:
        """``[pool(] act(bn(conv(x))) [)]``.  Under batch statistics the convolution runs WITHOUT its bias (the
        normalisation cancels it; ``ops.conv_without_bias``) and the BatchNorm kernels take the bias for the
        running mean and return its (analytically zero) gradient: no bias-add pass, no bias-gradient reduction."""
        y, cb = ops.conv_without_bias(conv, x, bn) if isinstance(act, nn.ReLU) else (conv(x), None)
        if cb is not None:
            return ops.batch_norm_relu(y, bn, relu=True, pool=pool is not None, conv_bias=cb)
        return cls._bn_relu(y, bn, act, pool)

    def _gate(self, merged: torch.Tensor, shared: torch.Tensor) -> torch.Tensor:
        folded = (isinstance(self.relu1, nn.ReLU) and merged.is_cuda and self.conv1.out_channels % 4 == 0
                  and ops.folded_gate_supported(torch.empty((0, self.conv1.out_channels, 1, 1), device=merged.device),
                                                shared, self.bn1, self.bn2, self.gate_precision))
        if folded:
            # bn1 + relu1 are folded into the gate kernels: the hidden activation never reaches HBM; conv1's bias is
            # folded into bn1 (batch statistics)
            squeezed, cb = ops.conv_without_bias(self.conv1, merged, self.bn1)
            return ops.attention_gate_folded(squeezed, self.bn1, shared, self.conv2, self.bn2, self.gate_precision, cb)
        squeezed = self.conv1(merged)
        hidden = self._bn_relu(squeezed, self.bn1, self.relu1)
        bn = self.bn2
        use_batch_stats = bn.training or bn.running_mean is None
        ops._bump_batch_counter(bn)
        momentum = 0.1 if bn.momentum is None else bn.momentum
        return ops.attention_gate(
            hidden, shared, self.conv2.weight, self.conv2.bias, bn.weight, bn.bias,
            bn.running_mean if bn.track_running_stats else None,
            bn.running_var if bn.track_running_stats else None,
            use_batch_stats, momentum, bn.eps, self.gate_precision,
        )


class AttentionModuleEncoder(_GatedAttention):
    def __init__(
        self,
        shared_1_channels: int,
        out_channels: int,
        shared_2_channels: int,
        prev_layer_out_channels: t.Optional[int] = None,
        hidden_channels: int = 64,
    ):
        super().__init__()
        self.is_first = prev_layer_out_channels is None
        self.prev_layer_out_channels = prev_layer_out_channels or 0
        self.shared_1_channels = shared_1_channels
        self.out_channels = out_channels
        self.shared_2_channels = shared_2_channels
        self.hidden_channels = hidden_channels
        self._build_gate(shared_1_channels + self.prev_layer_out_channels, hidden_channels, shared_2_channels)
        self.conv3 = nn.Conv2d(shared_2_channels, out_channels, kernel_size=3, padding=1)
        self.bn3 = nn.BatchNorm2d(out_channels)
        self.relu2 = nn.ReLU()
        self.maxpool = nn.MaxPool2d(kernel_size=2)

    def forward(self, conv1_shared, conv2_shared, prev_layer_outs=None):
        if self.is_first:
            merged = conv1_shared
        else:
            if prev_layer_outs is None:
                raise ValueError("prev_layer_outs must be provided for a non-first AttentionModuleEncoder")
            merged = torch.cat((conv1_shared, prev_layer_outs), dim=1)
        gated = self._gate(merged, conv2_shared)
        return self._conv_bn_relu(gated, self.conv3, self.bn3, self.relu2, self.maxpool)


class AttentionModuleDecoder(_GatedAttention):
    def __init__(
        self,
        shared_1_channels: int,
        shared_2_channels: int,
        prev_layer_out_channels: int,
        out_channels: int,
        hidden_channels: int = 64,
    ):
        super().__init__()
        self.shared_1_channels = shared_1_channels
        self.shared_2_channels = shared_2_channels
        self.prev_layer_out_channels = prev_layer_out_channels
        self.out_channels = out_channels
        self.hidden_channels = hidden_channels
        self._build_gate(shared_1_channels + hidden_channels, hidden_channels, shared_2_channels)
        # brings the previous attention output to the hidden width before the merge
        self.conv3 = nn.Conv2d(prev_layer_out_channels, hidden_channels, kernel_size=3, padding=1)
        self.bn3 = nn.BatchNorm2d(hidden_channels)
        self.relu2 = nn.ReLU()
        self.maxpool = nn.MaxPool2d(kernel_size=2)
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv_out = nn.Conv2d(shared_2_channels, out_channels, kernel_size=3, padding=1)
        self.bn_out = nn.BatchNorm2d(out_channels)
        self.relu_out = nn.ReLU()

    def forward(self, conv1_shared, prev_layer_outs, conv2_shared):
        prev = self._conv_bn_relu(prev_layer_outs, self.conv3, self.bn3, self.relu2)
        if conv1_shared.shape[2:] != conv2_shared.shape[2:]:
            raise ValueError("conv1_shared and conv2_shared must share their spatial size")
        up = self.up
        if (conv1_shared.shape[2:] != prev.shape[2:] and isinstance(up, nn.Upsample) and up.mode == "bilinear"
                and up.align_corners and up.scale_factor == 2 and ops.upsample2_cat_supported(conv1_shared, prev)):
            merged = ops.upsample2_cat(conv1_shared, prev)  # up-sampled straight into its slice of the concatenation
        else:
            if conv1_shared.shape[2:] != prev.shape[2:]:
                prev = up(prev)
            merged = torch.cat((conv1_shared, prev), dim=1)
        gated = self._gate(merged, conv2_shared)
        return self._conv_bn_relu(gated, self.conv_out, self.bn_out, self.relu_out)


class MTANDown(nn.Module):
    """Shared double conv + one encoder attention module per task."""

    def __init__(self, in_channels, out_channels, task_attn_modules, apply_pool: bool = True):
        super().__init__()
        self.dconv = DoubleConv(in_channels, out_channels)
        self.pool = nn.MaxPool2d(2) if apply_pool else nn.Identity()
        self.task_attn_modules = task_attn_modules

    def forward(self, x, prev_layer_outs=None):
        shared = self.dconv(x)
        outs = [
            attn(conv1_shared=x, conv2_shared=shared, prev_layer_outs=prev_layer_outs[i] if prev_layer_outs else None)
            for i, attn in enumerate(self.task_attn_modules)
        ]
        return self.pool(shared), outs


class MTANUp(nn.Module):
    """Transposed-conv up-sampling, skip merge, shared double conv + decoder attention per task."""

    def __init__(self, in_channels, out_channels, task_attn_modules):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels)
        self.task_attn_modules = task_attn_modules
        self.out_channels = out_channels
        self.in_channels = in_channels

    def forward(self, x1, x2, task_attn_prev_outs):
        merged = concat_slightly_diff_sized_tensors(self.up(x1), x2)
        shared = self.conv(merged)
        outs = [
            attn(conv1_shared=merged, prev_layer_outs=task_attn_prev_outs[i], conv2_shared=shared)
            for i, attn in enumerate(self.task_attn_modules)
        ]
        return shared, outs


class MTANMiniUnet(nn.Module):
    def __init__(
        self,
        in_channels: int,
        map_tasks_to_num_channels: t.Dict[str, int],
        task_subnets_hidden_channels: int = 128,
        encoder_first_channel: int = 64,
        encoder_num_channels: int = 4,
    ):
        super().__init__()
        self.num_tasks = len(map_tasks_to_num_channels)
        self.in_channels = in_channels
        T, hid = self.num_tasks, task_subnets_hidden_channels

        # shared (global) sub-network channel plan (reference: mtan_model.py:262-300)
        enc_out = [encoder_first_channel * 2**i for i in range(encoder_num_channels)]
        enc_in = [in_channels] + enc_out[:-1]
        dec_out = enc_out[::-1]
        dec_in = [enc_out[-1] * 2] + dec_out[:-1]
        self.global_subnet_enc_out_channels, self.global_subnet_enc_in_channels = enc_out, enc_in
        self.global_subnet_dec_out_channels, self.global_subnet_dec_in_channels = dec_out, dec_in
        # task (attention) sub-networks follow the shared widths level by level
        self.task_attn_out_channels_enc = list(enc_out)
        self.task_attn_prev_layer_out_channels_enc = [None] + enc_out[:-1]
        self.task_attn_in_channels_enc = list(enc_in)
        self.task_subnet_out_channels_dec = list(dec_out)
        self.task_attn_in_channels_dec = list(dec_in)
        self.task_attn_prev_layer_out_channels_dec = [enc_out[-1]] + dec_out[:-1]

        self.bottleneck = DoubleConv(enc_out[-1], enc_out[-1] * 2)

        def enc_attn(i):
            return nn.ModuleList(
                AttentionModuleEncoder(
                    shared_1_channels=enc_in[i], shared_2_channels=enc_out[i], out_channels=enc_out[i],
                    prev_layer_out_channels=self.task_attn_prev_layer_out_channels_enc[i], hidden_channels=hid,
                ) for _ in range(T))

        def dec_attn(i):
            return nn.ModuleList(
                AttentionModuleDecoder(
                    shared_1_channels=dec_in[i], shared_2_channels=dec_out[i],
                    prev_layer_out_channels=self.task_attn_prev_layer_out_channels_dec[i],
                    out_channels=dec_out[i], hidden_channels=hid,
                ) for _ in range(T))

        self.enc_layers = nn.ModuleList(
            MTANDown(enc_in[i], enc_out[i], enc_attn(i), apply_pool=False) for i in range(len(enc_in)))
        self.dec_layers = nn.ModuleList(MTANUp(dec_in[i], dec_out[i], dec_attn(i)) for i in range(len(dec_in)))
        self.pool = nn.MaxPool2d(2)
        # 1x1 task heads on the last decoder attention output (mtan_model.py:367-376)
        self.map_tasks_to_heads = nn.ModuleDict(
            {task: nn.Conv2d(dec_out[-1], c, kernel_size=1) for task, c in map_tasks_to_num_channels.items()})

    def forward_features(self, x: torch.Tensor) -> t.Dict[str, torch.Tensor]:
        """Per-task ``[B, C_last, H, W]`` features in front of the heads (NHWC memory)."""
        x = x.contiguous(memory_format=torch.channels_last)
        attn_outs, skips = None, []
        for down in self.enc_layers:
            x, attn_outs = down(x, attn_outs)
            skips.append(x)
            x = self.pool(x)
        x = self.bottleneck(x)
        for i, up in enumerate(self.dec_layers):
            x, attn_outs = up(x, skips[-(i + 1)], attn_outs)
        return {task: attn_outs[i] for i, task in enumerate(self.map_tasks_to_heads.keys())}

    def forward(self, x: torch.Tensor, features_only: bool = False) -> t.Dict[str, torch.Tensor]:
        feats = self.forward_features(x)
        if features_only:  # fused head+loss route of MTLModule (works through a DDP wrapper too)
            return feats
        return {task: head(feats[task]) for task, head in self.map_tasks_to_heads.items()}
