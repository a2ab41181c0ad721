"""Synthetic Cityscapes- / NYUv2-shaped batches: the benchmark and the smoke loop
have no datasets.  Layouts are the ones the reference datasets produce:
``img [B,3,H,W] f32 U(0,1)``, ``mask [B,H,W] int64``, ``depth [B,H,W,1] f32``."""
from __future__ import annotations

import torch


def make_batch(batch_size: int, height: int, width: int, num_classes: int, dataset: str = "cityscapes",
               seed: int = 11, device="cpu", pin: bool = False) -> dict:
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch_size, 3, height, width, generator=g)
    mask = torch.randint(0, num_classes, (batch_size, height, width), generator=g)
    if dataset == "nyuv2":  # absolute depth [0,10] m normalised by max_depth (nyuv2.py:128-133)
        depth = torch.rand(batch_size, height, width, 1, generator=g)
    else:  # Cityscapes disparity-like depth in [0,0.5] with ~20 % exact zeros (invalid pixels)
        depth = torch.rand(batch_size, height, width, 1, generator=g) * 0.5
        depth[torch.rand(batch_size, height, width, 1, generator=g) < 0.2] = 0.0
    batch = {"img": img, "mask": mask, "depth": depth}
    if pin and torch.cuda.is_available():
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if str(device) != "cpu":
        batch = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
        batch["img"] = batch["img"].contiguous(memory_format=torch.channels_last)
    return batch
