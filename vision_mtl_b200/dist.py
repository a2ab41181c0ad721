"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

The reference is single-process; data parallelism is the only strategy that fits
this workload: every hot-path kernel is per-pixel independent, so a step shards over the batch
with exactly one exchange -- the gradient all-reduce (bucketed by DDP, overlapped with the
backward) -- plus ONE small all-reduce of the step statistics:

    float64 [C*C + 8] = confusion matrix (integers are exact in fp64 below 2^53) ++
                        {loss_sum_weight, n_steps_unused, sum|p-t|, n_pixels, sum rel, n_valid_depth, loss, 1}

The global confusion matrix is the exact sum of the per-rank matrices, so the metrics derived
from it equal a single-process evaluation of the concatenated batch bit for bit.  BatchNorm and
SILog keep standard per-replica semantics (what wrapping the reference in DDP would give).
"""
from __future__ import annotations

import os
import typing as t

import torch
import torch.distributed as dist

from . import ops


def env_world() -> t.Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process -> 0,0,1)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_distributed(backend: t.Optional[str] = None) -> t.Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)  # binds the communicator, no barrier() warning
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Contiguous batch shard of every tensor in a ``{"img","mask","depth"}`` dict."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        if n % world != 0:
            raise ValueError(f"global batch {n} is not divisible by world size {world}")
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def wrap_data_parallel(module, device_index: t.Optional[int] = None, **ddp_kwargs):
    """Wrap ``module.model`` (an ``MTLModule``'s network) in DDP; gradients of every parameter --
    including those produced by the fused kernels' backward -- land in ordinary ``.grad`` tensors,
    so DDP buckets and all-reduces them while the backward is still running."""
    if not is_distributed():
        return module
    from torch.nn.parallel import DistributedDataParallel as DDP

    kwargs = dict(gradient_as_bucket_view=True, broadcast_buffers=False)
    kwargs.update(ddp_kwargs)
    if device_index is not None:
        kwargs.setdefault("device_ids", [device_index])
    if device_index is not None and torch.cuda.is_available():
        # whole-step CUDA-graph capture needs DDP constructed on a side stream (torch CUDA-graphs notes)
        side = torch.cuda.Stream(device=device_index)
        side.wait_stream(torch.cuda.current_stream(device_index))
        with torch.cuda.stream(side):
            module.model = DDP(module.model, **kwargs)
        torch.cuda.current_stream(device_index).wait_stream(side)
    else:
        module.model = DDP(module.model, **kwargs)
    return module


def enable_stat_sync(module, group=None) -> None:
    """Global-batch-exact mode: call BEFORE ``wrap_data_parallel``.  The library's BatchNorm /
    gate / SILog kernels all-reduce their moments between their two halves (``ops.set_stat_sync``); BatchNorm
    layers that stay on ATen (the csnet / basic backbone leaves) become ``torch.nn.SyncBatchNorm``.  Cross-entropy
    stays a per-replica mean: with equal shards and no ignored pixels the average over replicas is the global mean."""
    if not is_distributed():
        return
    ops.set_stat_sync(True, group)
    inner = module.model
    if not hasattr(inner, "forward_features"):  # MTAN runs every BatchNorm through the library
        module.model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(inner, group)


STAT_EXTRA = 8


def pack_step_stats(conf: torch.Tensor, loss: torch.Tensor, depth_sums: t.Optional[torch.Tensor] = None) -> torch.Tensor:
    """One flat float64 buffer for the single metric all-reduce of a step."""
    dev = conf.device
    if depth_sums is None:  # {P, sum|p-t|, n_valid, sum rel}
        depth_sums = torch.zeros(4, dtype=torch.float64, device=dev)
    d = depth_sums.to(torch.float64)
    head = torch.zeros(2, dtype=torch.float64, device=dev)
    tail = torch.stack([loss.detach().to(torch.float64).reshape(()), torch.ones((), dtype=torch.float64, device=dev)])
    # layout of the extras: [0, 0, sum|p-t|, P, sum rel, n_valid, loss, 1]
    # (no list-indexing here: it would build a host index tensor, i.e. a copy inside graph capture)
    return torch.cat([conf.reshape(-1).to(torch.float64), head, torch.stack([d[1], d[0], d[3], d[2]]), tail])


def unpack_step_stats(buf: torch.Tensor, num_classes: int) -> dict:
    CC = num_classes * num_classes
    conf = buf[:CC].round().to(torch.int64).reshape(num_classes, num_classes)
    e = buf[CC:]
    # no host reads here: this runs inside the captured step graph
    return {"confusion": conf, "loss": e[6] / e[7], "replicas": e[7],
            "mae": e[2] / torch.clamp(e[3], min=1.0), "abs_rel": e[4] / torch.clamp(e[5], min=1.0)}


def allreduce_step_stats(conf: torch.Tensor, loss: torch.Tensor, depth_sums: t.Optional[torch.Tensor] = None) -> dict:
    """Global (all ranks) confusion matrix / mean loss / depth errors with ONE collective."""
    buf = pack_step_stats(conf, loss, depth_sums)
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = unpack_step_stats(buf, conf.shape[0])
    if conf.is_cuda:
        seg = ops.seg_metrics(out["confusion"])
        out.update(accuracy=seg[0], jaccard_index=seg[1], fbeta_score=seg[2])
    return out
