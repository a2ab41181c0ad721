"""Deterministic, RNG-free test data shared by the golden generator, the oracle tests and the
GPU parity tests.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Values come from NumPy's frozen legacy MT19937 stream seeded by a CRC of the tensor's NAME, so
the same tensors are rebuilt bit-for-bit on any machine without storing them: golden files only
have to hold the reference's OUTPUTS.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

_M = 2**31


def lcg_uniform(n: int, seed: int) -> np.ndarray:
    """n values in [-0.5, 0.5), float64.  ``numpy.random.RandomState`` (MT19937) is NumPy's frozen
    legacy generator: its stream is guaranteed bit-identical across NumPy versions and platforms."""
    return np.random.RandomState(seed % (2**32)).random_sample(n) - 0.5


def key_seed(key: str, salt: int = 0) -> int:
    return (zlib.crc32(key.encode()) + 7919 * salt) % _M


def tensor(shape, key: str, scale: float = 1.0, salt: int = 0) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    v = lcg_uniform(n, key_seed(key, salt)) * (2.0 * scale)
    return torch.from_numpy(v.astype(np.float32)).reshape(tuple(shape))


def labels(shape, num_classes: int, key: str) -> torch.Tensor:
    n = int(np.prod(shape))
    v = (lcg_uniform(n, key_seed(key)) + 0.5) * num_classes
    return torch.from_numpy(np.clip(v.astype(np.int64), 0, num_classes - 1)).reshape(tuple(shape))


def fill_state_dict(sd: dict, salt: int = 0) -> dict:
    """A well-conditioned value for every entry of a state_dict, keyed by NAME (independent of
    module construction order, which differs between the reference and the product)."""
    out = {}
    for k, v in sd.items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_mean"):
            out[k] = tensor(shape, k, 0.1, salt)
        elif k.endswith("running_var"):
            out[k] = 1.0 + tensor(shape, k, 0.3, salt).abs()
        elif v.dim() == 1 and k.endswith("weight"):  # BatchNorm scale
            out[k] = 1.0 + tensor(shape, k, 0.2, salt)
        elif v.dim() == 1:  # any bias
            out[k] = tensor(shape, k, 0.1, salt)
        elif k.endswith("weights"):  # cross-stitch alphas: U(0,1) like reset_parameters
            out[k] = tensor(shape, k, 0.5, salt) + 0.5
        else:  # conv / transposed-conv kernels: variance-preserving uniform
            fan_in = int(np.prod(shape[1:])) if v.dim() > 1 else shape[0]
            out[k] = tensor(shape, k, float(np.sqrt(3.0 / max(fan_in, 1))), salt)
        out[k] = out[k].to(v.dtype)
    return out


def image_batch(B: int, H: int, W: int, num_classes: int, key: str = "batch", depth_zero_frac: float = 0.2,
                depth_max: float = 0.5) -> dict:
    """Synthetic Cityscapes-shaped batch: img U(0,1), mask randint(0,C), depth U(0,0.5)
    with a fraction of exact zeros, layouts as the reference datasets produce them.  NYUv2-shaped:
    ``depth_zero_frac=0, depth_max=1`` (depth U(0,10)/10, no zeros)."""
    img = tensor((B, 3, H, W), key + "/img", 0.5) + 0.5
    mask = labels((B, H, W), num_classes, key + "/mask")
    depth = (tensor((B, H, W, 1), key + "/depth", depth_max / 2) + depth_max / 2).clamp_min(0.0)
    zero = (tensor((B, H, W, 1), key + "/zero", 0.5) + 0.5) < depth_zero_frac
    depth = torch.where(zero, torch.zeros_like(depth), depth)
    return {"img": img, "mask": mask, "depth": depth}


def summarize(t: torch.Tensor, k: int = 8) -> np.ndarray:
    """Compact, order-sensitive fingerprint of a tensor: [sum, l2, first k, strided k]."""
    f = t.detach().cpu().double().reshape(-1)
    n = f.numel()
    head = f[:k]
    stride = max(n // k, 1)
    samp = f[::stride][:k]
    pad = lambda x: torch.cat([x, torch.zeros(k - x.numel(), dtype=torch.float64)])  # noqa: E731
    return torch.cat([f.sum().reshape(1), f.norm().reshape(1), pad(head), pad(samp)]).numpy()
