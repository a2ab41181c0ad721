"""Per-kernel CPU restatements of the reference arithmetic (torch, fp32 or fp64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites
the reference lines it restates (paths relative to ``/root/reference``).
Layouts here are the reference's own (NCHW, stacked ``[T,B,C,H,W]``); the
product kernels work on NHWC and the tests convert.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# Cross-stitch  (vision_mtl/models/cross_stitch_model.py:32-37)
# --------------------------------------------------------------------------
def xstitch_reference_diag(weights: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """What the reference einsum ``"aa,abcij->abcij"`` / ``"aac,abcij->abcij"``
    actually computes: the repeated index takes the diagonal of alpha, so
    ``y[a] = alpha[a,a(,c)] * x[a]`` (cross_stitch_model.py:33-36)."""
    T = x.shape[0]
    idx = torch.arange(T)
    if weights.dim() == 3:  # channel-wise, weights [T,T,C]
        d = weights[idx, idx, :]  # [T,C]
        return d[:, None, :, None, None] * x
    d = weights[idx, idx]  # [T]
    return d[:, None, None, None, None] * x


def xstitch_full_mix(weights: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """The cross-stitch unit of Misra et al. that BASELINE.json's north_star
    names: ``y[a] = sum_b alpha[a,b(,c)] * x[b]``."""
    if weights.dim() == 3:
        return torch.einsum("abc,bncij->ancij", weights, x)
    return torch.einsum("ab,bncij->ancij", weights, x)


# --------------------------------------------------------------------------
# MTAN attention gate (vision_mtl/models/mtan_model.py:71-75 and :158-162)
# --------------------------------------------------------------------------
def gate_forward(
    h: torch.Tensor,  # [B,K,H,W]   relu(bn1(conv1(.)))
    s: torch.Tensor,  # [B,N,H,W]   shared features (conv2_shared)
    weight: torch.Tensor,  # [N,K,1,1]
    bias: torch.Tensor,  # [N]
    gamma: torch.Tensor,
    beta: torch.Tensor,
    running_mean: torch.Tensor | None,
    running_var: torch.Tensor | None,
    training: bool,
    momentum: float = 0.1,
    eps: float = 1e-5,
) -> torch.Tensor:
    """``conv2 -> bn2 -> sigmoid -> * conv2_shared`` exactly as the reference
    sequences the ATen ops (mtan_model.py:71-75)."""
    z = F.conv2d(h, weight, bias)
    u = F.batch_norm(z, running_mean, running_var, gamma, beta, training, momentum, eps)
    return s * torch.sigmoid(u)


# --------------------------------------------------------------------------
# Heads, post-processing and losses (mtan_model.py:401-404, lit_module.py:120-144,
# losses.py:14-36)
# --------------------------------------------------------------------------
def head_project(feat: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """1x1 task head, ``nn.Conv2d(32, C, kernel_size=1)`` (mtan_model.py:367-376)."""
    return F.conv2d(feat, weight, bias)


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """``nn.CrossEntropyLoss()`` defaults (lit_module.py:31,123)."""
    return F.cross_entropy(logits, target, ignore_index=ignore_index)


def segm_predictions(logits: torch.Tensor) -> torch.Tensor:
    """``argmax(softmax(logits, 1), 1)`` (lit_module.py:137-138)."""
    return torch.argmax(F.softmax(logits, dim=1), dim=1)


def depth_predictions(depth_logits: torch.Tensor) -> torch.Tensor:
    """``sigmoid(depth_logits).permute(0, 2, 3, 1)`` -> [B,H,W,1] (lit_module.py:139)."""
    return torch.sigmoid(depth_logits).permute(0, 2, 3, 1)


def silog(pred: torch.Tensor, target: torch.Tensor, min_depth: float = 1e-3, mask=None) -> torch.Tensor:
    """SILog loss (losses.py:22-36).  The bilinear ``interpolate`` call at
    losses.py:24-27 resizes ``pred`` to its own trailing two dims for the
    (B,H,W,1) layout, i.e. it is the identity, and is skipped here.  An explicit ``mask``
    replaces the ``target > min_depth`` one (losses.py:29-30)."""
    if mask is None:
        mask = target > min_depth
    g = torch.log(pred[mask]) - torch.log(target[mask])
    dg = torch.var(g) + 0.15 * torch.pow(torch.mean(g), 2)
    return 10 * torch.sqrt(dg)


def total_loss(loss_segm, loss_depth, w_segm: float = 1.0, w_depth: float = 1.0):
    """lit_module.py:126."""
    return w_segm * loss_segm + w_depth * loss_depth


def depth_mae(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """torchmetrics ``MeanAbsoluteError`` over all elements, no mask
    (lit_module.py:68,112)."""
    return (pred - target).abs().sum() / pred.numel()


def depth_abs_rel(pred: torch.Tensor, target: torch.Tensor, min_depth: float = 1e-3) -> torch.Tensor:
    """Mean of |p-t|/t over t > min_depth.  Not in the reference; named by
    north_star ("depth abs/rel error sums")."""
    m = target > min_depth
    return ((pred[m] - target[m]).abs() / target[m]).sum() / m.sum()


# --------------------------------------------------------------------------
# conv -> BatchNorm2d -> ReLU (-> MaxPool2d(2)) tails
# (vision_mtl/utils/model_utils.py:61-80; vision_mtl/models/mtan_model.py:65-69, :77-81, :141-142, :165-167)
# --------------------------------------------------------------------------
def bn_relu(x, gamma, beta, running_mean, running_var, training: bool, relu: bool = True, pool: bool = False,
            momentum: float = 0.1, eps: float = 1e-5) -> torch.Tensor:
    """``nn.BatchNorm2d`` then (optionally) ``nn.ReLU`` then (optionally) ``nn.MaxPool2d(2)``, as the
    reference sequences the ATen ops."""
    y = F.batch_norm(x, running_mean, running_var, gamma, beta, training, momentum, eps)
    if relu:
        y = F.relu(y)
    if pool:
        y = F.max_pool2d(y, 2)
    return y
