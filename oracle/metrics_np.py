"""numpy restatement of the torchmetrics==0.7.3 metrics the reference uses.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED: torchmetrics 0.7.3 is a ``requirements.txt:10`` dependency
whose source is absent from /root/reference and from this image; the reference
has no test holding its outputs.  The definitions below are the published ones
for the configuration at vision_mtl/lit_module.py:48-69 and are cross-checked
against scikit-learn in ``tests/test_oracle_metrics.py``.

All inputs are integer label maps; everything derives from one integer
confusion matrix ``cm[target, pred]`` which is the bit-exact artefact.
"""
from __future__ import annotations

import numpy as np


def confusion_matrix(pred, target, num_classes: int, ignore_index: int | None = None) -> np.ndarray:
    """``bincount(target*C + pred, minlength=C*C).reshape(C, C)`` (rows = target)."""
    p = np.asarray(pred).reshape(-1).astype(np.int64)
    t = np.asarray(target).reshape(-1).astype(np.int64)
    keep = (t >= 0) & (t < num_classes) & (p >= 0) & (p < num_classes)
    if ignore_index is not None:
        keep &= t != ignore_index
    idx = t[keep] * num_classes + p[keep]
    return np.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def accuracy_micro(cm: np.ndarray) -> float:
    """``Accuracy(average="micro")``, mdmc "global": sum(tp) / #pixels (lit_module.py:49-54)."""
    return float(np.trace(cm)) / float(cm.sum())


def jaccard_macro_absent0(cm: np.ndarray) -> float:
    """``JaccardIndex(num_classes=C)``: absent_score=0.0, elementwise_mean over all C
    classes (lit_module.py:63-67)."""
    tp = np.diag(cm).astype(np.float64)
    union = cm.sum(0).astype(np.float64) + cm.sum(1).astype(np.float64) - tp
    iou = np.where(union > 0, tp / np.where(union > 0, union, 1.0), 0.0)
    return float(iou.mean())


def fbeta_weighted(cm: np.ndarray, beta: float = 1.0) -> float:
    """``FBetaScore(beta=1, average="weighted", mdmc_average="global")``
    (lit_module.py:55-62): per-class F weighted by support (= tp + fn)."""
    tp = np.diag(cm).astype(np.float64)
    fp = cm.sum(0).astype(np.float64) - tp
    fn = cm.sum(1).astype(np.float64) - tp
    b2 = beta * beta
    num = (1.0 + b2) * tp
    den = (1.0 + b2) * tp + b2 * fn + fp
    f = np.where(den > 0, num / np.where(den > 0, den, 1.0), 0.0)
    support = tp + fn
    return float((f * support).sum() / support.sum())


def mean_abs_error(pred, target) -> float:
    """``MeanAbsoluteError``: sum |p - t| / numel (lit_module.py:68)."""
    p = np.asarray(pred, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    return float(np.abs(p - t).sum() / p.size)


def all_seg_metrics(cm: np.ndarray) -> dict:
    return {
        "accuracy": accuracy_micro(cm),
        "jaccard_index": jaccard_macro_absent0(cm),
        "fbeta_score": fbeta_weighted(cm),
    }
