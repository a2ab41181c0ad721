"""Plain-PyTorch CPU restatement of the reference models and of one training step.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA path and the
timed CPU baseline of ``bench.py`` -- never imported by ``vision_mtl_b200``.

* ``mtan_forward`` is a *functional* restatement of ``MTANMiniUnet`` driven by a
  ``state_dict`` (the key layout of SURVEY A.4 is the contract shared by the reference, this
  port and the product), citing vision_mtl/models/mtan_model.py.
* ``CSNetOracle`` restates ``CSNet.forward``'s flat walk (cross_stitch_model.py:102-157) over
  caller-supplied task networks.
* ``step_losses_and_metrics`` restates ``MTLModule.shared_step`` (lit_module.py:75-144).

Pinned against outputs of the unmodified reference by ``oracle/make_golden.py`` ->
``tests/golden/*.npz`` (checked in ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import re
import typing as t

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels_ref as K
from . import metrics_np as MN


# ------------------------------------------------------------------------------------------
# functional MTAN
# ------------------------------------------------------------------------------------------
def _conv(p, pre, x, padding=0):
    return F.conv2d(x, p[pre + ".weight"], p.get(pre + ".bias"), padding=padding)


def _bn(p, pre, x, training, momentum=0.1, eps=1e-5):
    if training and (pre + ".num_batches_tracked") in p:
        p[pre + ".num_batches_tracked"] += 1
    return F.batch_norm(x, p[pre + ".running_mean"], p[pre + ".running_var"], p[pre + ".weight"],
                        p[pre + ".bias"], training, momentum, eps)


def _double_conv(p, pre, x, training):
    """model_utils.py:61-80: (3x3 conv no bias, BN, ReLU) x 2 under ``double_conv.{0,1,3,4}``."""
    x = F.relu(_bn(p, pre + ".double_conv.1", _conv(p, pre + ".double_conv.0", x, 1), training))
    return F.relu(_bn(p, pre + ".double_conv.4", _conv(p, pre + ".double_conv.3", x, 1), training))


def _gate(p, pre, merged, shared, training):
    """mtan_model.py:65-75 / :152-162: 1x1 squeeze, then conv2 -> bn2 -> sigmoid -> * shared."""
    hidden = F.relu(_bn(p, pre + ".bn1", _conv(p, pre + ".conv1", merged), training))
    attn = torch.sigmoid(_bn(p, pre + ".bn2", _conv(p, pre + ".conv2", hidden), training))
    return shared * attn


def _attn_encoder(p, pre, conv1_shared, conv2_shared, prev, training):
    """AttentionModuleEncoder.forward (mtan_model.py:49-83)."""
    merged = conv1_shared if prev is None else torch.cat((conv1_shared, prev), dim=1)
    g = _gate(p, pre, merged, conv2_shared, training)
    g = F.relu(_bn(p, pre + ".bn3", _conv(p, pre + ".conv3", g, 1), training))
    return F.max_pool2d(g, 2)


def _attn_decoder(p, pre, conv1_shared, prev, conv2_shared, training):
    """AttentionModuleDecoder.forward (mtan_model.py:133-169)."""
    prev = F.relu(_bn(p, pre + ".bn3", _conv(p, pre + ".conv3", prev, 1), training))
    if conv1_shared.shape[2:] != prev.shape[2:]:
        prev = F.interpolate(prev, scale_factor=2, mode="bilinear", align_corners=True)
    g = _gate(p, pre, torch.cat((conv1_shared, prev), dim=1), conv2_shared, training)
    return F.relu(_bn(p, pre + ".bn_out", _conv(p, pre + ".conv_out", g, 1), training))


def _pad_cat(x1, x2):
    """concat_slightly_diff_sized_tensors (model_utils.py:46-58)."""
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return torch.cat([x2, x1], dim=1)


def mtan_tasks(p: dict) -> t.List[str]:
    return [m.group(1) for k in p for m in [re.match(r"map_tasks_to_heads\.(\w+)\.weight$", k)] if m]


def mtan_features(p: dict, x: torch.Tensor, training: bool = True) -> t.Dict[str, torch.Tensor]:
    """Everything of MTANMiniUnet.forward (mtan_model.py:378-404) except the 1x1 heads."""
    levels = len({m.group(1) for k in p for m in [re.match(r"enc_layers\.(\d+)\.dconv", k)] if m})
    tasks = mtan_tasks(p)
    attn, skips = None, []
    for i in range(levels):  # MTANDown (mtan_model.py:187-201), pooling done by the caller (:388)
        shared = _double_conv(p, f"enc_layers.{i}.dconv", x, training)
        attn = [_attn_encoder(p, f"enc_layers.{i}.task_attn_modules.{ti}", x, shared,
                              None if attn is None else attn[ti], training) for ti in range(len(tasks))]
        skips.append(shared)
        x = F.max_pool2d(shared, 2)
    x = _double_conv(p, "bottleneck", x, training)
    for i in range(levels):  # MTANUp (mtan_model.py:222-243)
        up = F.conv_transpose2d(x, p[f"dec_layers.{i}.up.weight"], p[f"dec_layers.{i}.up.bias"], stride=2)
        merged = _pad_cat(up, skips[-(i + 1)])
        x = _double_conv(p, f"dec_layers.{i}.conv", merged, training)
        attn = [_attn_decoder(p, f"dec_layers.{i}.task_attn_modules.{ti}", merged, attn[ti], x, training)
                for ti in range(len(tasks))]
    return {task: attn[ti] for ti, task in enumerate(tasks)}


def mtan_forward(p: dict, x: torch.Tensor, training: bool = True) -> t.Dict[str, torch.Tensor]:
    feats = mtan_features(p, x, training)
    return {task: _conv(p, f"map_tasks_to_heads.{task}", f) for task, f in feats.items()}


# ------------------------------------------------------------------------------------------
# CSNet flat walk
# ------------------------------------------------------------------------------------------
class CSNetOracle(nn.Module):
    """Restates CSNet (cross_stitch_model.py:40-201) over caller-supplied identical task nets.
    ``state_dict`` keys match the reference: ``models.<task>...`` and
    ``cross_stitch_layers.<name with '.'->'_'>.weights``."""

    ENC = r"0.encoder.model.blocks.(\d+)$"
    DEC = r"0.decoder.blocks.(\d+)$"

    def __init__(self, models: dict, channel_wise_stitching: bool = False, mode: str = "reference_diag"):
        super().__init__()
        self.tasks = list(models.keys())
        self.models = nn.ModuleDict(models)
        self.mode = mode
        proto = self.models[self.tasks[0]]
        self.names = [n for n, _ in list(proto.named_modules())[1:]]
        self.n_enc = len(list(self._get(proto, "0.encoder.model.blocks").children()))
        self.n_dec = len(list(self._get(proto, "0.decoder.blocks").children()))
        self.sites = [n for n in self.names if self._is_site(n)]
        T = len(self.tasks)
        chans = self._site_channels(proto) if channel_wise_stitching else [None] * len(self.sites)

        class _Alpha(nn.Module):
            def __init__(self, c):
                super().__init__()
                self.weights = nn.Parameter(torch.rand(T, T, c) if c is not None else torch.rand(T, T))

        self.cross_stitch_layers = nn.ModuleDict({n.replace(".", "_"): _Alpha(c) for n, c in zip(self.sites, chans)})

    @staticmethod
    def _get(m, name):
        for part in name.split("."):
            m = getattr(m, part)
        return m

    @staticmethod
    def _is_site(name):  # model_utils.py:100-115
        parts = name.split(".")
        if "encoder" in parts and len(parts) == 5:
            return int(parts[-1]) != 0
        return "decoder" in parts and len(parts) == 4

    def _enc_saved(self, idx):  # cross_stitch_model.py:116-120
        return idx not in (0, self.n_enc - 1, self.n_dec - 1)

    def _site_channels(self, proto):  # cross_stitch_model.py:171-201
        named = list(proto.named_modules())[1:]
        pos = {n: i for i, (n, _) in enumerate(named)}
        out, enc_c = [], []
        for site in self.sites:
            i = pos[site] - 1
            while not isinstance(named[i][1], nn.Conv2d):
                i -= 1
            c = named[i][1].out_channels
            m = re.match(self.ENC, site)
            if m and self._enc_saved(int(m.group(1))):
                enc_c.append(c)
            m = re.match(self.DEC, site)
            if m and int(m.group(1)) != self.n_dec - 1:
                c += enc_c[-int(m.group(1)) - 1]
            out.append(c)
        return out

    def forward(self, x):
        feats = {t_: x.clone() for t_ in self.tasks}
        saved = {t_: [] for t_ in self.tasks}
        for name in self.names:
            for t_ in self.tasks:
                layer = self._get(self.models[t_], name)
                m = re.match(self.ENC, name)
                if m and self._enc_saved(int(m.group(1))):
                    saved[t_].append(feats[t_].clone())
                m = re.match(self.DEC, name)
                if m:
                    idx = int(m.group(1))
                    if idx != self.n_dec - 1:
                        feats[t_] = _pad_cat(feats[t_], saved[t_][-idx - 1])
                    else:
                        feats[t_] = F.interpolate(feats[t_], scale_factor=2, mode="nearest")
                if next(layer.children(), None) is not None:
                    continue  # containers are skipped; only leaves are applied
                feats[t_] = layer(feats[t_])
            if name in self.sites:
                w = self.cross_stitch_layers[name.replace(".", "_")].weights
                stacked = torch.stack([feats[t_] for t_ in self.tasks], dim=0)
                mixed = K.xstitch_reference_diag(w, stacked) if self.mode == "reference_diag" else K.xstitch_full_mix(w, stacked)
                feats = {t_: mixed[i] for i, t_ in enumerate(self.tasks)}
        return feats


# ------------------------------------------------------------------------------------------
# one step: losses + metrics (MTLModule.shared_step, lit_module.py:75-144)
# ------------------------------------------------------------------------------------------
def step_losses_and_metrics(raw_out: dict, gt_mask, gt_depth, num_classes: int,
                            w_segm: float = 1.0, w_depth: float = 1.0) -> dict:
    logits, depth_logits = raw_out["segm"], raw_out["depth"]
    preds = K.segm_predictions(logits)
    depth_pred = K.depth_predictions(depth_logits)
    loss_segm = K.cross_entropy(logits, gt_mask)
    loss_depth = K.silog(depth_pred, gt_depth)
    cm = MN.confusion_matrix(preds.numpy(), gt_mask.numpy(), num_classes)
    out = {"loss": K.total_loss(loss_segm, loss_depth, w_segm, w_depth), "loss_segm": loss_segm,
           "loss_depth": loss_depth, "confusion": cm, "segm_predictions": preds,
           "depth_predictions": depth_pred, "mae": float(K.depth_mae(depth_pred.detach(), gt_depth))}
    out.update(MN.all_seg_metrics(cm))
    return out
