"""Epoch summaries and metric printing (reference: vision_mtl/utils/loss_utils.py)."""
from __future__ import annotations

import numbers
import typing as t

import torch


def _as_scalar_tensor(v) -> torch.Tensor:
    return v.detach().float().reshape(()) if isinstance(v, torch.Tensor) else torch.tensor(float(v))


def summarize_epoch_metrics(step_results: dict, metric_name_prefix: t.Optional[str] = None) -> dict:
    """Unweighted mean of the per-batch scalars of every key, then clear the lists
    (loss_utils.py:27-44).  All keys are reduced on the device and fetched with ONE copy, where
    the reference's ``torch.tensor([...])`` synchronises once per stored element."""
    prefix = "" if metric_name_prefix is None else metric_name_prefix + "/"
    keys = list(step_results.keys())
    means = []
    for k in keys:
        vals = step_results[k]
        if len(vals) == 0:
            means.append(torch.tensor(float("nan")))
        else:
            means.append(torch.stack([_as_scalar_tensor(v) for v in vals]).mean())
    dev = next((m.device for m in means if m.is_cuda), torch.device("cpu"))
    host = torch.stack([m.to(dev) for m in means]).tolist()
    for k in keys:
        step_results[k].clear()
    return {f"{prefix}{k}": v for k, v in zip(keys, host)}


def print_metrics(prefix: str, train_epoch_metrics: dict) -> str:
    """Format (and print) the latest value of every metric (loss_utils.py:47-64)."""
    latest = {}
    for k, v in train_epoch_metrics.items():
        if isinstance(v, torch.Tensor):
            latest[k] = v.reshape(-1)[-1] if v.numel() > 1 else v.reshape(())
        elif isinstance(v, numbers.Number):
            latest[k] = v
        else:
            latest[k] = v[-1]
    tens = [x for x in latest.values() if isinstance(x, torch.Tensor)]
    if tens:  # one device->host copy for all tensor-valued entries
        host = iter(torch.stack([x.detach().float() for x in tens]).tolist())
        latest = {k: (next(host) if isinstance(x, torch.Tensor) else x) for k, x in latest.items()}
    parts = []
    for k, value in latest.items():
        print(f"{prefix}/{k}: {value:.3f} ")
        parts.append(f"{k}: {value:.3f} ")
    return "".join(parts)
