"""Metric objects with the call surface the reference uses from torchmetrics 0.7.3
(``metric(preds, target)`` returning the per-batch value, ``.to(device)``; configured at
vision_mtl/lit_module.py:48-69).

All three segmentation metrics are functions of one integer ``C x C`` confusion matrix
(``cm[target, pred]``), accumulated by ``csrc/metrics.cu`` in a single pass over the label maps
(16 B/pixel) instead of torchmetrics' one-hot expansion.  Like the
reference, ``forward`` returns the batch-local value and additionally accumulates a running
state that ``compute()`` reduces.
"""
from __future__ import annotations

import typing as t

import torch

from . import ops


class _ConfusionMetric:
    _index = 0  # position in ops.seg_metrics output

    def __init__(self, num_classes: int, ignore_index: t.Optional[int] = None, **_unused):
        self.num_classes = num_classes
        self.ignore_index = -100 if ignore_index is None else ignore_index
        self.device = torch.device("cpu")
        self.confmat: t.Optional[torch.Tensor] = None

    def to(self, device):
        self.device = torch.device(device)
        if self.confmat is not None:
            self.confmat = self.confmat.to(self.device)
        return self

    def reset(self):
        self.confmat = None

    def from_confusion(self, conf: torch.Tensor) -> torch.Tensor:
        return ops.seg_metrics(conf)[self._index]

    def update_state(self, conf: torch.Tensor) -> None:
        self.confmat = conf.clone() if self.confmat is None else self.confmat + conf

    def __call__(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        conf = ops.confusion_accumulate(preds, target, self.num_classes, None, self.ignore_index)
        self.update_state(conf)
        return self.from_confusion(conf)

    forward = __call__

    def compute(self) -> torch.Tensor:
        if self.confmat is None:
            raise RuntimeError("compute() called before any update")
        return self.from_confusion(self.confmat)


class Accuracy(_ConfusionMetric):
    """``average="micro"``: sum(tp) / #pixels."""

    _index = 0

    def __init__(self, threshold: float = 0.5, num_classes: t.Optional[int] = None,
                 ignore_index: t.Optional[int] = None, average: str = "micro", **kw):
        if average != "micro":
            raise NotImplementedError("only average='micro' (the reference configuration)")
        super().__init__(num_classes, ignore_index)


class JaccardIndex(_ConfusionMetric):
    """Mean IoU over all classes, absent classes score 0."""

    _index = 1

    def __init__(self, num_classes: int, threshold: float = 0.5, ignore_index: t.Optional[int] = None, **kw):
        super().__init__(num_classes, ignore_index)


class FBetaScore(_ConfusionMetric):
    """``beta=1, average="weighted", mdmc_average="global"``: support-weighted F1."""

    _index = 2

    def __init__(self, beta: float = 1.0, threshold: float = 0.5, num_classes: t.Optional[int] = None,
                 average: str = "weighted", ignore_index: t.Optional[int] = None,
                 mdmc_average: t.Optional[str] = "global", **kw):
        if beta != 1.0 or average != "weighted":
            raise NotImplementedError("only beta=1, average='weighted' (the reference configuration)")
        super().__init__(num_classes, ignore_index)


class MeanAbsoluteError:
    """sum |p - t| / numel over all elements (no mask)."""

    def __init__(self):
        self.device = torch.device("cpu")
        self.sums: t.Optional[torch.Tensor] = None

    def to(self, device):
        self.device = torch.device(device)
        return self

    def reset(self):
        self.sums = None

    def __call__(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        s = ops.depth_error_sums(preds.reshape(-1), target.reshape(-1))
        self.sums = s.clone() if self.sums is None else self.sums + s
        return (s[1] / s[0]).to(torch.float32)

    forward = __call__

    def compute(self) -> torch.Tensor:
        return (self.sums[1] / self.sums[0]).to(torch.float32)
