"""Adam as the reference builds it (``torch.optim.Adam(params, lr)``, training_lit.py:51 /
lit_module.py:194), run by ONE multi-tensor kernel per parameter group (``vmtl_adam_step``).

Drop-in for ``torch.optim.Adam`` where the reference uses it: same constructor keywords, ``param_groups``,
``state_dict()`` layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), works with
``ReduceLROnPlateau``.  Differences in mechanism, not in arithmetic:

* the moments of a group live in two flat fp32 buffers (the per-parameter state tensors are views into them);
* the step counter is ONE device scalar per group, advanced by the kernel itself, and the learning rate may be a
  device tensor -- so a step captured into a CUDA graph keeps counting and follows the scheduler;
* there is no CPU path: parameters must be fp32 CUDA tensors (``VmtlError`` otherwise).
"""
from __future__ import annotations

import ctypes
import typing as t

import torch

from . import _lib, ops


def _dense(x: torch.Tensor) -> bool:
    return x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))


def _same_layout(p: torch.Tensor, g: torch.Tensor) -> bool:
    """Same element order in memory (strides of size-1 dimensions do not matter)."""
    return _dense(p) and _dense(g) and all(ps == gs for n, ps, gs in zip(p.shape, p.stride(), g.stride()) if n > 1)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: t.Union[float, torch.Tensor] = 1e-3, betas: t.Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, amsgrad: bool = False, *, foreach=None,
                 maximize: bool = False, capturable: bool = True, differentiable: bool = False, fused=None):
        if amsgrad or maximize or differentiable:
            raise NotImplementedError("vision_mtl_b200.optim.Adam covers the reference configuration: "
                                      "amsgrad=False, maximize=False, differentiable=False")
        if not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameters")
        if isinstance(lr, torch.Tensor) and lr.numel() != 1:
            raise ValueError("a tensor learning rate must have one element")
        # `capturable` / `fused` / `foreach` are accepted for API compatibility: every step of this class is capture-safe
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=True, differentiable=False, fused=True)
        super().__init__(params, defaults)
        self._flat: t.Dict[int, dict] = {}  # group index -> flat moment buffers, offsets, step, ticket, cached table

    # ------------------------------------------------------------------ state
    def _group_state(self, gi: int, group: dict) -> dict:
        fs = self._flat.get(gi)
        if fs is not None and fs["nparams"] == len(group["params"]):
            return fs
        params = group["params"]
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise _lib.VmtlError("vision_mtl_b200.optim.Adam runs on fp32 CUDA parameters only; got "
                                     f"{p.dtype} on {p.device}")
        dev = params[0].device
        offsets, total = [], 0
        for p in params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4  # every tensor's moments start 16-byte aligned
        fs = {
            "exp_avg": torch.zeros(max(total, 4), dtype=torch.float32, device=dev),
            "exp_avg_sq": torch.zeros(max(total, 4), dtype=torch.float32, device=dev),
            "step": torch.zeros((), dtype=torch.float32, device=dev),
            "ticket": torch.zeros(1, dtype=torch.int32, device=dev),
            "offsets": offsets, "key": None, "table": None, "nparams": len(params),
        }
        for p, off in zip(params, offsets):
            st = self.state[p]
            n = p.numel()
            # views with the parameter's own strides: element k of the flat slice pairs with element k of p's storage
            m_view = torch.as_strided(fs["exp_avg"], p.size(), p.stride(), off)
            v_view = torch.as_strided(fs["exp_avg_sq"], p.size(), p.stride(), off)
            if "exp_avg" in st:  # state loaded / created before: move it into the flat buffers
                m_view.copy_(st["exp_avg"])
                v_view.copy_(st["exp_avg_sq"])
                fs["step"].copy_(torch.as_tensor(st["step"], dtype=torch.float32))
            st["step"] = fs["step"]  # one counter per group, shared by its parameters
            st["exp_avg"], st["exp_avg_sq"] = m_view, v_view
        self._flat[gi] = fs
        return fs

    def state_dict(self) -> dict:
        """torch.optim.Adam's layout; every parameter gets its OWN copy of the (group-wide) step counter, so the
        dictionary can be loaded into a stock optimizer that advances the counters one by one."""
        sd = super().state_dict()
        sd["state"] = {k: {**v, "step": v["step"].clone()} if "step" in v else dict(v) for k, v in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict: dict) -> None:
        super().load_state_dict(state_dict)
        self._flat.clear()  # the loaded per-parameter tensors are re-packed by the next step
        for gi, group in enumerate(self.param_groups):
            self._group_state(gi, group)

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure: t.Optional[t.Callable] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            fs = self._group_state(gi, group)
            live = []
            for i, p in enumerate(group["params"]):
                g = p.grad
                if g is None or p.numel() == 0:
                    continue
                if g.is_sparse or g.dtype != torch.float32 or not g.is_cuda:
                    raise _lib.VmtlError("vision_mtl_b200.optim.Adam needs dense fp32 CUDA gradients")
                if not _same_layout(p, g):  # rare (a gradient produced in another memory format): one re-layout copy
                    g = torch.empty_like(p).copy_(g)
                live.append((i, p, g))
            key = tuple((i, p.data_ptr(), g.data_ptr()) for i, p, g in live)
            if key != fs["key"]:
                n = len(live)
                P, G = (ctypes.c_void_p * max(n, 1))(), (ctypes.c_void_p * max(n, 1))()
                N, O = (ctypes.c_int64 * max(n, 1))(), (ctypes.c_int64 * max(n, 1))()
                for j, (i, p, g) in enumerate(live):
                    P[j], G[j], N[j], O[j] = p.data_ptr(), g.data_ptr(), p.numel(), fs["offsets"][i]
                fs["key"], fs["table"] = key, (P, G, N, O, n)
                fs["nbytes"] = 28 * sum(p.numel() for _, p, _ in live)
            P, G, N, O, n = fs["table"]
            lr = group["lr"]
            lr_dev = lr if isinstance(lr, torch.Tensor) and lr.is_cuda else None
            lr_host = 0.0 if lr_dev is not None else float(lr)
            if lr_dev is not None and lr_dev.dtype != torch.float32:
                raise _lib.VmtlError("a device learning rate must be fp32")
            beta1, beta2 = group["betas"]
            ops._Prof.launches += max(1, -(-n // lib.vmtl_adam_max_tensors_per_launch())) - 1
            ops._call("adam_step", fs["nbytes"], P, G, N, O, n, ops._p(fs["exp_avg"]), ops._p(fs["exp_avg_sq"]),
                      ops._p(lr_dev), lr_host, ops._p(fs["step"]), ops._p(fs["ticket"]), float(beta1), float(beta2),
                      float(group["eps"]), float(group["weight_decay"]), ops._stream())
        return loss
