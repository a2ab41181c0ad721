"""Whole-step CUDA-graph capture of the training step ("host loop de-sync +
CUDA-graphed step").

The reference loop issues several thousand tiny kernels per step from Python (csnet: ~190 leaf
modules x 2 tasks, forward and backward) and synchronises six times per step; on a B200 that is
host-bound.  ``GraphedTrainStep`` captures forward + fused losses/metrics + backward + Adam into
ONE graph on static buffers: per step the host copies the batch into the static inputs, replays
the graph and (optionally) reads back the packed step scalars -- one launch, one D2H copy.

All hand-written kernels are capture-safe: they only enqueue work on the current stream, their
workspaces come from torch's (graph-pool aware) caching allocator and TMA descriptors are by-value
kernel parameters over static addresses.

The optimizer must be ``capturable`` and its learning rate a DEVICE tensor (``make_optimizer``):
a Python-float lr would be baked into the graph and a later ``ReduceLROnPlateau`` step would be
silently ignored on replay.
"""
from __future__ import annotations

import typing as t

import torch

from .lit_module import STEP_KEYS, MTLModule


def make_optimizer(params, lr: float, device) -> torch.optim.Optimizer:
    """Adam as the reference builds it (training_lit.py:51) on the library's multi-tensor kernel
    (``optim.Adam``: one launch per step, device-side step counter), with the learning rate held in a device
    tensor so schedulers (which ``fill_`` it) act on graph replays."""
    from .optim import Adam

    return Adam(params, lr=torch.tensor(float(lr), dtype=torch.float32, device=device))


class GraphedTrainStep:
    def __init__(self, module: MTLModule, optimizer: torch.optim.Optimizer, example_batch: dict,
                 warmup: int = 3, after_backward: t.Optional[t.Callable[[], None]] = None,
                 profile: bool = False, preserve_state: bool = False, record_step_outputs: bool = False):
        """``example_batch`` fixes shapes/dtypes; ``after_backward`` (e.g. a metric all-reduce) is
        captured between backward and the optimizer step.  ``profile=True`` captures a pair of external
        CUDA events around every library call, so ``kernel_stats()`` reports per-op device time of the
        last replay (an instrumented copy for measurement; the event nodes cost a little step time).
        ``preserve_state=True`` (training loops): parameters, buffers and optimizer state are restored
        in place after the eager warm-up steps, so building the graph does not train.
        ``record_step_outputs=True``: every replay appends its five scalars to
        ``module.step_outputs["train"]`` like an eager ``training_step``."""
        self.module, self.optimizer = module, optimizer
        for group in optimizer.param_groups:
            # optimizers with a host-side step counter (Adam & co.) expose `capturable`; SGD-like ones have no
            # such state and capture as they are
            if "capturable" in optimizer.defaults and not group.get("capturable", False):
                raise ValueError("GraphedTrainStep needs a capturable optimizer (graph_step.make_optimizer)")
            if preserve_state and not isinstance(group["lr"], torch.Tensor):
                raise ValueError("GraphedTrainStep in a training loop needs the learning rate in a device tensor "
                                 "(graph_step.make_optimizer): a float lr is baked into the graph")
        dev = example_batch["img"].device
        self.static = {k: torch.empty_like(v) for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v)
        self._after_backward = after_backward
        self._record = record_step_outputs
        self._keep = {k: len(v) for k, v in module.step_outputs["train"].items()}
        snapshot = self._snapshot() if preserve_state else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # eager warm-up on a side stream (allocator, cuDNN autotune, DDP buckets)
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        self.kernel_records: list = []
        if profile:
            from . import ops

            with ops.kernel_timing() as records, torch.cuda.graph(self.graph):
                self._capture()
            self.kernel_records = list(records)
        else:
            with torch.cuda.graph(self.graph):
                self._capture()
        self._trim_step_outputs()
        if snapshot is not None:
            self._restore(snapshot)

    # -- state kept across the warm-up steps (addresses stay: the graph holds them) -------------------
    def _state_tensors(self) -> t.List[torch.Tensor]:
        ts = [p for p in self.module.model.parameters()] + [b for b in self.module.model.buffers()]
        return ts

    def _snapshot(self):
        return [x.detach().clone() for x in self._state_tensors()]

    def _restore(self, snapshot) -> None:
        with torch.no_grad():
            for x, saved in zip(self._state_tensors(), snapshot):
                x.copy_(saved)
            for st in self.optimizer.state.values():  # Adam moments and step counters created by the warm-up
                for v in st.values():
                    if isinstance(v, torch.Tensor):
                        v.zero_()

    def _capture(self) -> None:
        self.loss = self._eager_step()
        self.scalars = self.module.last_step_scalars
        self.confusion = self.module.last_confusion
        self.global_stats = getattr(self.module, "last_global_stats", None)

    def kernel_stats(self) -> dict:
        """name -> {calls, ms, bytes, gbps} of the last replay (``profile=True``; synchronise first)."""
        from . import ops

        return ops.summarize_timing(self.kernel_records)

    def _eager_step(self) -> torch.Tensor:
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.module.training_step(self.static, 0)
        loss.backward()
        if self._after_backward is not None:
            self._after_backward()
        self.optimizer.step()
        return loss

    def _trim_step_outputs(self) -> None:
        # warm-up / capture appended device scalars that replays will overwrite in place
        for k, n in self._keep.items():
            del self.module.step_outputs["train"][k][n:]

    def load(self, batch: dict) -> None:
        """Copy a (host or device) batch into the static inputs (async on the current stream)."""
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)

    def __call__(self, batch: t.Optional[dict] = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        self.module.last_step_scalars = self.scalars
        self.module.last_confusion = self.confusion
        if self.global_stats is not None:
            self.module.last_global_stats = self.global_stats
        if self._record:  # the static scalars are overwritten by the next replay: keep this step's copy
            vals = self.scalars.clone()
            rec = self.module.step_outputs["train"]
            for i, k in enumerate(STEP_KEYS):
                rec[k].append(vals[i])
        return self.loss
