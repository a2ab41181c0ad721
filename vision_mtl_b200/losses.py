"""Depth loss, drop-in for ``vision_mtl/losses.py``.

``SILogLoss.forward(pred, target)`` takes the reference's arguments: ``pred`` = sigmoid
depth predictions and ``target`` in the ``(B,H,W,1)`` layout.  On that layout the reference's
bilinear ``interpolate`` (losses.py:23-27) resizes to the tensor's own size, i.e. it is the
identity, so it is not executed.  The arithmetic runs in ``csrc/head_loss.cu``: one streaming
pass for the masked moments ``(n, sum g, sum g^2)`` in fp64 and one pass for the gradient,
instead of two boolean-mask gathers (each a host sync), two logs, var, mean, pow and sqrt.
"""
from __future__ import annotations

import typing as t

import torch
import torch.nn as nn

from . import ops


class SILogLoss(nn.Module):
    def __init__(self, min_depth: float = 1e-3):
        super().__init__()
        self.min_depth = min_depth

    def forward(
        self,
        pred: torch.Tensor,
        target: torch.Tensor,
        mask: t.Optional[torch.Tensor] = None,
        interpolate: bool = True,
        min_depth: t.Optional[float] = None,
    ) -> torch.Tensor:
        """Reference-compatible entry point on sigmoid predictions.

        The fused training path does not come through here (``MTLModule`` feeds depth *logits* or
        head features to ``ops.head_silog``).  For predictions the logit is recovered as
        ``log(p) - log1p(-p)``; sigmoid of that reproduces ``p`` to 1 ulp, and autograd chains the
        kernel's d/dlogit back through it.
        """
        md = self.min_depth if min_depth is None else min_depth
        tgt = target.reshape(-1)
        if mask is not None:
            # losses.py:29-33 with an explicit mask: every selected pixel counts, whatever its depth.  The kernel
            # masks by `target > min_depth`, so deselected pixels get a target below a zero threshold.  (Selected
            # pixels with target <= 0 make the reference's log NaN; here they drop out.)
            tgt = torch.where(mask.reshape(-1).to(torch.bool), tgt, torch.full_like(tgt, -1.0))
            md = 0.0
        # fp32 sigmoid saturates to exactly 0 / 1: keep the recovered logit (and its derivative) finite
        p = pred.reshape(pred.shape[0], 1, -1, 1).clamp(1e-37, 1.0 - 2.0 ** -24)
        logit = torch.log(p) - torch.log1p(-p)
        silog, _, _, _ = ops.head_silog(logit, None, None, tgt, md, want_pred=False)
        return silog
