"""Host-side operators over the C ABI of ``libvmtl_b200.so``.

Every function here enqueues hand-written sm_100a kernels on torch's current CUDA stream;
PyTorch is used for device memory, streams and autograd plumbing only.  There is no CPU or
eager fallback: non-CUDA tensors raise.

Layout: feature maps are NHWC (``torch.channels_last``); a ``[B,C,H,W]`` channels-last
tensor is the row-major ``[B*H*W, C]`` matrix the kernels stream.
"""
from __future__ import annotations

import ctypes
from contextlib import contextmanager
from typing import Optional, Sequence

import torch

from . import _lib

XS_REFERENCE_DIAG = 0
XS_FULL_MIX = 1
GATE_FP32_FFMA = 0
GATE_TC_3XTF32 = 1
GATE_TC_TF32 = 2
LAYOUT_NCHW = 0
LAYOUT_NHWC = 1

_XS_MODES = {"reference_diag": XS_REFERENCE_DIAG, "full_mix": XS_FULL_MIX}
_GATE_PRECISIONS = {"fp32_ffma": GATE_FP32_FFMA, "tc_3xtf32": GATE_TC_3XTF32, "tc_tf32": GATE_TC_TF32}

# default contraction precision of the gate; tests and bench may override
default_gate_precision = "tc_3xtf32"


# --------------------------------------------------------------------------------------
# launch accounting + optional per-launch CUDA-event timing (bench.py roofline numbers)
# --------------------------------------------------------------------------------------
class _Prof:
    enabled = False
    records: list = []  # (name, algorithmic_bytes, issued_bytes, start_event, end_event)
    launches = 0  # kernels launched by this library since the last reset
    nvtx = False  # NVTX range around every library call (tools/site_bench.py --nvtx, profiling runs)


def enable_nvtx(on: bool = True) -> None:
    _Prof.nvtx = bool(on)


# kernels enqueued by one C call (for the `gpu_launches` claim in bench.py)
_KERNELS_PER_CALL = {
    "xstitch_fwd": 1, "xstitch_bwd": 2, "gate_fwd": 3, "gate_bwd": 3, "head_ce_fwd": 2,
    "head_ce_bwd": 2, "ce_logits_fwd": 2, "ce_logits_bwd": 1, "head_silog_fwd": 2,
    "head_silog_bwd": 2, "confusion_accum": 1, "depth_err_sums": 2, "seg_metrics": 1,
    "xstitch_cat_fwd": 1, "xstitch_cat_bwd": 2,
    "bnrelu_fwd": 3, "bnrelu_bwd": 3, "bnrelu_pool_fwd": 3, "bnrelu_pool_bwd": 3,
    "bn_moments": 2, "bnrelu_fwd_global": 2, "bnrelu_bwd_moments": 2, "bnrelu_bwd_global": 2,
    "gate_fwd_moments": 2, "gate_fwd_global": 2, "gate_bwd_moments": 2, "gate_bwd_global": 2, "silog_finalize": 1,
}


def reset_launch_count() -> None:
    _Prof.launches = 0


def launch_count() -> int:
    return _Prof.launches


@contextmanager
def kernel_timing():
    """Record a CUDA-event pair around every library call made inside the block."""
    _Prof.enabled = True
    _Prof.records = []
    try:
        yield _Prof.records
    finally:
        _Prof.enabled = False


def summarize_timing(records) -> dict:
    """name -> {calls, ms, bytes, issued_bytes, gbps}; call after torch.cuda.synchronize().
    ``bytes`` are the ALGORITHMIC bytes of SURVEY 8(d) (what ``gbps`` and every roofline fraction are
    computed from); ``issued_bytes`` are what the kernels of the call actually request from HBM."""
    out: dict = {}
    for name, nbytes, issued, e0, e1 in records:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0, "issued_bytes": 0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["bytes"] += nbytes
        d["issued_bytes"] += issued
    for d in out.values():
        d["gbps"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    return out


def _call(name: str, algo_bytes: int, *args, issued: Optional[int] = None) -> None:
    fn = getattr(_lib.load(), "vmtl_" + name)
    _Prof.launches += _KERNELS_PER_CALL.get(name, 1)
    if _Prof.nvtx:
        torch.cuda.nvtx.range_push("vmtl_" + name)
        try:
            rc = fn(*args)
        finally:
            torch.cuda.nvtx.range_pop()
        _lib.check(rc, "vmtl_" + name)
        return
    if _Prof.enabled:
        # inside a stream capture the events must be "external" to become timing-capable graph nodes
        # (they are then re-recorded by every replay)
        ext = torch.cuda.is_current_stream_capturing()
        e0 = torch.cuda.Event(enable_timing=True, external=ext)
        e1 = torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        rc = fn(*args)
        e1.record()
        _Prof.records.append((name, algo_bytes, algo_bytes if issued is None else issued, e0, e1))
    else:
        rc = fn(*args)
    _lib.check(rc, "vmtl_" + name)


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VmtlError(
                "vision_mtl_b200 ops run on CUDA (sm_100a) only; got a tensor on " + str(t.device)
            )


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """Channels-last view of a [B,C,H,W] fp32 tensor (no copy when already NHWC)."""
    if x.dim() != 4:
        raise ValueError("expected a [B,C,H,W] tensor")
    if x.dtype != torch.float32:
        raise TypeError("vision_mtl_b200 kernels are fp32")
    return x.contiguous(memory_format=torch.channels_last)


def _ptr_array(ts: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    return arr


# --------------------------------------------------------------------------------------
# Global-batch statistics under data parallelism
# --------------------------------------------------------------------------------------
class _Sync:
    enabled = False
    group = None


def set_stat_sync(enabled: bool, group=None) -> None:
    """Global-batch-exact mode: every batch statistic of the hot path (BatchNorm moments of the gate, of the
    hidden layer and of every conv -> BN -> ReLU chain, forward and backward; the SILog moments) is all-reduced
    over ``group`` between the two halves of its op, so an N-GPU step computes what the reference computes on
    the concatenated batch.  Needs an initialised ``torch.distributed`` process group (NCCL) and equal shards."""
    _Sync.enabled = bool(enabled)
    _Sync.group = group


def stat_sync_world() -> int:
    """Replicas whose statistics are combined (1: local statistics, the default data-parallel semantics)."""
    if not _Sync.enabled:
        return 1
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(_Sync.group)


def _allreduce_moments(t: torch.Tensor) -> None:
    import torch.distributed as dist

    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_Sync.group)


# --------------------------------------------------------------------------------------
# Cross-stitch
# --------------------------------------------------------------------------------------
def xstitch_forward(xs: Sequence[torch.Tensor], alpha: torch.Tensor, mode: int) -> list:
    T = len(xs)
    xs = [_nhwc(x) for x in xs]
    _need_cuda(alpha, *xs)
    B, C, H, W = xs[0].shape
    ys = [torch.empty_like(x) for x in xs]
    npix = B * H * W
    cw = 1 if alpha.dim() == 3 else 0
    a = alpha.detach().contiguous()
    xa, ya = _ptr_array(xs), _ptr_array(ys)
    _call("xstitch_fwd", 2 * 4 * T * npix * C, xa, ya, _p(a), T, npix, C, cw, mode, _stream())
    return ys


def xstitch_backward(dys, xs, alpha: torch.Tensor, mode: int, need_dx: bool = True):
    T = len(xs)
    dys = [_nhwc(d) for d in dys]
    B, C, H, W = xs[0].shape
    npix = B * H * W
    cw = 1 if alpha.dim() == 3 else 0
    a = alpha.detach().contiguous()
    dalpha = torch.empty_like(a)
    dxs = [torch.empty_like(x) for x in xs] if need_dx else None
    lib = _lib.load()
    ws = _workspace(lib.vmtl_xstitch_bwd_workspace_bytes(T, npix, C, cw), a.device)
    dya, xa = _ptr_array(dys), _ptr_array(xs)
    dxa = _ptr_array(dxs) if need_dx else None
    _call("xstitch_bwd", (3 if need_dx else 2) * 4 * T * npix * C, dya, xa, dxa, _p(a), _p(dalpha), T,
          npix, C, cw, mode, _p(ws), ws.numel(), _stream())
    return dxs, dalpha


class CrossStitchFunction(torch.autograd.Function):
    """y[a] = sum_b alpha[a,b(,c)] x[b] (or the reference's diagonal-only variant)."""

    @staticmethod
    def forward(ctx, alpha, mode, *xs):
        xs = [_nhwc(x) for x in xs]
        ys = xstitch_forward(xs, alpha, mode)
        ctx.mode = mode
        ctx.need_dx = any(ctx.needs_input_grad[2:])
        ctx.save_for_backward(alpha, *xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        alpha, *xs = ctx.saved_tensors
        dys = [torch.zeros_like(x) if d is None else d for d, x in zip(dys, xs)]
        dxs, dalpha = xstitch_backward(dys, xs, alpha, ctx.mode, ctx.need_dx)
        if dxs is None:
            dxs = [None] * len(xs)
        return (dalpha, None, *dxs)


def cross_stitch(xs: Sequence[torch.Tensor], alpha: torch.Tensor, mode: str = "reference_diag"):
    return list(CrossStitchFunction.apply(alpha, _XS_MODES[mode], *xs))


class CrossStitchCatFunction(torch.autograd.Function):
    """Cross-stitch over ``cat([skip, zero_pad(x)], 1)`` (``up2=False``) or over the nearest x2 up-sampling of
    ``x`` (``up2=True``) per task, gathered by the kernel: the assembled tensor is never materialised."""

    @staticmethod
    def forward(ctx, alpha, mode, up2, n_skip, *tensors):
        skips = [_nhwc(t) for t in tensors[:n_skip]]
        xs = [_nhwc(t) for t in tensors[n_skip:]]
        T = len(xs)
        _need_cuda(alpha, *xs, *skips)
        B, Cx, Hi, Wi = xs[0].shape
        if skips:
            _, Cs, Ho, Wo = skips[0].shape
        else:
            Cs, Ho, Wo = 0, (2 * Hi if up2 else Hi), (2 * Wi if up2 else Wi)
        C = Cs + Cx
        ys = [torch.empty((B, C, Ho, Wo), dtype=torch.float32, device=xs[0].device, memory_format=torch.channels_last)
              for _ in range(T)]
        cw = 1 if alpha.dim() == 3 else 0
        a = alpha.detach().contiguous()
        geom = (T, B, Ho, Wo, Cs, Hi, Wi, Cx, 1 if up2 else 0, cw, mode)
        nbytes = 4 * T * (B * Ho * Wo * (C + Cs) + B * Hi * Wi * Cx)
        _call("xstitch_cat_fwd", nbytes, _ptr_array(skips) if skips else None, _ptr_array(xs), _ptr_array(ys), _p(a),
              *geom, _stream())
        ctx.geom = geom
        ctx.n_skip = n_skip
        ctx.save_for_backward(alpha, *skips, *xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        alpha, *rest = ctx.saved_tensors
        n_skip = ctx.n_skip
        skips, xs = rest[:n_skip], rest[n_skip:]
        T, B, Ho, Wo, Cs, Hi, Wi, Cx, up2, cw, mode = ctx.geom
        dev = xs[0].device
        dys = [torch.zeros((B, Cs + Cx, Ho, Wo), device=dev, memory_format=torch.channels_last)
               if d is None else _nhwc(d) for d in dys]
        a = alpha.detach().contiguous()
        dalpha = torch.empty_like(a)
        need = ctx.needs_input_grad[4:]
        dskips = [torch.empty_like(t) if need[i] else None for i, t in enumerate(skips)]
        dxs = [torch.empty_like(t) if need[n_skip + i] else None for i, t in enumerate(xs)]
        lib = _lib.load()
        ws = _workspace(lib.vmtl_xstitch_cat_bwd_workspace_bytes(T, B, Ho, Wo, Cs, Hi, Wi, Cx, up2, cw), dev)

        def ptrs(ts):
            return (ctypes.c_void_p * len(ts))(*[0 if t is None else t.data_ptr() for t in ts])

        nbytes = 4 * T * (B * Ho * Wo * (Cs + Cx) + 2 * (B * Ho * Wo * Cs + B * Hi * Wi * Cx))  # R dy, R + W (skip, x)
        _call("xstitch_cat_bwd", nbytes, _ptr_array(dys), _ptr_array(skips) if skips else None, _ptr_array(xs),
              ptrs(dskips) if skips else None, ptrs(dxs), _p(a), _p(dalpha), T, B, Ho, Wo, Cs, Hi, Wi, Cx, up2, cw, mode,
              _p(ws), ws.numel(), _stream())
        return (dalpha, None, None, None, *dskips, *dxs)


def cross_stitch_cat(skips: Sequence[torch.Tensor], xs: Sequence[torch.Tensor], alpha: torch.Tensor,
                     mode: str = "reference_diag", up2: bool = False):
    """``stitch([cat([skip_t, zero_pad(x_t)], 1) for t])`` / ``stitch([upsample2(x_t) for t])`` in one kernel."""
    skips = list(skips or [])
    return list(CrossStitchCatFunction.apply(alpha, _XS_MODES[mode], bool(up2), len(skips), *skips, *xs))


# --------------------------------------------------------------------------------------
# BatchNorm2d (+ ReLU, + 2x2 max-pool)
# --------------------------------------------------------------------------------------
class _Counters:
    pending: Optional[dict] = None  # id(counter) -> [counter, increments] while deferred_batch_counters() is active


@contextmanager
def deferred_batch_counters():
    """Collect the ``num_batches_tracked += 1`` of every BatchNorm call made inside the block and apply them with
    ONE multi-tensor add at exit (the reference bumps each counter with its own kernel: ~100 launches per MTAN
    forward).  Only for modules with a fixed momentum -- a cumulative average needs its count at call time."""
    if _Counters.pending is not None:  # nested: the outermost block flushes
        yield
        return
    _Counters.pending = {}
    try:
        yield
    finally:
        pending, _Counters.pending = _Counters.pending, None
        once = [c for c, n in pending.values() if n == 1]
        if once:
            torch._foreach_add_(once, 1)
        for c, n in pending.values():
            if n > 1:
                c.add_(n)


def _bump_batch_counter(bn) -> None:
    if not (bn.training and bn.track_running_stats and bn.num_batches_tracked is not None):
        return
    if _Counters.pending is not None and bn.momentum is not None:
        slot = _Counters.pending.setdefault(id(bn.num_batches_tracked), [bn.num_batches_tracked, 0])
        slot[1] += 1
    else:
        bn.num_batches_tracked.add_(1)


def _bn_forward(x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, pool, y, stats, world,
                conv_bias=None):
    """Statistics + finalize (+ apply when ``y`` is given) of one BatchNorm; ``world > 1`` all-reduces the moments.
    ``conv_bias`` (batch statistics only): the bias the producing convolution did NOT add -- it only moves the
    running mean."""
    cb = _p(conv_bias.detach()) if conv_bias is not None else _p(None)
    B, C, H, W = x.shape
    M = B * H * W
    dev = x.device
    ws = _workspace(_lib.load().vmtl_bnrelu_workspace_bytes(M, C), dev)
    out_rows = 0 if y is None else y.numel() // C
    if training and world > 1:
        mom = torch.empty((2, C), dtype=torch.float64, device=dev)
        _call("bn_moments", 4 * C * M, _p(x), M, C, _p(mom), _p(ws), ws.numel(), _stream())
        _allreduce_moments(mom)
        _call("bnrelu_fwd_global", 4 * C * (M + out_rows) if y is not None else 0, _p(x), _p(gamma.detach()),
              _p(beta.detach()), _p(running_mean), _p(running_var), float(momentum), float(eps), 1 if relu else 0,
              1 if pool else 0, B, H, W, C, _p(mom), M * world, _p(y), _p(stats[0]), _p(stats[1]), _p(stats[2:]), cb,
              _p(ws), ws.numel(), _stream())
        if y is None:
            _Prof.launches -= 1
        return
    nbytes = 4 * C * ((2 * M if training else M) + out_rows) if y is not None else (4 * C * M if training else 0)
    if pool:
        _call("bnrelu_pool_fwd", nbytes, _p(x), _p(gamma.detach()), _p(beta.detach()), _p(running_mean),
              _p(running_var), float(momentum), float(eps), 1 if training else 0, 1 if relu else 0, B, H, W, C,
              _p(y), _p(stats[0]), _p(stats[1]), _p(stats[2:]), cb, _p(ws), ws.numel(), _stream())
    else:
        _call("bnrelu_fwd", nbytes, _p(x), _p(gamma.detach()), _p(beta.detach()), _p(running_mean),
              _p(running_var), float(momentum), float(eps), 1 if training else 0, 1 if relu else 0, M, C,
              _p(y), _p(stats[0]), _p(stats[1]), _p(stats[2:]), cb, _p(ws), ws.numel(), _stream())
    _Prof.launches -= (0 if training else 1) + (0 if y is not None else 1)  # no statistics / no apply pass


def _bn_backward(dy, x, stats, training, relu, pool, need_dx, world, want_dconv_bias=False):
    """-> (dx | None, dgamma, dbeta[, dconv_bias]) of one BatchNorm (+ReLU, +pool); ``world > 1``: global
    normalisation terms.  ``dconv_bias`` = sum over pixels of dx, the gradient of a folded convolution bias."""
    B, C, H, W = x.shape
    M = B * H * W
    dev = x.device
    dx = torch.empty_like(x) if need_dx else None
    ws = _workspace(_lib.load().vmtl_bnrelu_workspace_bytes(M, C), dev)
    dy_rows = dy.numel() // C
    if training and world > 1:
        mom = torch.empty((2, C), dtype=torch.float64, device=dev)
        _call("bnrelu_bwd_moments", 4 * C * (M + dy_rows), _p(dy), _p(x), _p(stats[2:]), _p(stats[0]), _p(stats[1]),
              1 if relu else 0, 1 if pool else 0, B, H, W, C, _p(mom), _p(ws), ws.numel(), _stream())
        local = mom.to(torch.float32)  # this replica's (dbeta, dgamma): averaged with every other gradient
        _allreduce_moments(mom)
        if need_dx:
            _call("bnrelu_bwd_global", 4 * C * (2 * M + dy_rows), _p(dy), _p(x), _p(stats[2:]), _p(stats[0]),
                  _p(stats[1]), 1 if relu else 0, 1 if pool else 0, B, H, W, C, _p(mom), M * world, _p(dx), _p(ws),
                  ws.numel(), _stream())
        if want_dconv_bias:  # this replica's sum of dx: A (sum_local g - M_local c1_global); sums to zero over replicas
            dcb = stats[2] * (local[0] - (mom[0] * (float(M) / float(M * world))).to(torch.float32))
            return dx, local[1], local[0], dcb
        return dx, local[1], local[0]
    dgb = torch.empty((3 if want_dconv_bias else 2, C), dtype=torch.float32, device=dev)
    dcb_p = _p(dgb[2]) if want_dconv_bias else _p(None)
    nbytes = 4 * C * (2 * (M + dy_rows) + (M if need_dx else 0))
    if pool:
        _call("bnrelu_pool_bwd", nbytes, _p(dy), _p(x), _p(stats[2:]), _p(stats[0]), _p(stats[1]),
              1 if training else 0, 1 if relu else 0, B, H, W, C, _p(dx), _p(dgb[0]), _p(dgb[1]), dcb_p, _p(ws),
              ws.numel(), _stream())
    else:
        _call("bnrelu_bwd", nbytes, _p(dy), _p(x), _p(stats[2:]), _p(stats[0]), _p(stats[1]),
              1 if training else 0, 1 if relu else 0, M, C, _p(dx), _p(dgb[0]), _p(dgb[1]), dcb_p, _p(ws), ws.numel(),
              _stream())
    if not need_dx:
        _Prof.launches -= 1
    if want_dconv_bias:
        return dx, dgb[0], dgb[1], dgb[2]
    return dx, dgb[0], dgb[1]


class BNReLUFunction(torch.autograd.Function):
    """``[maxpool2(] relu( batch_norm(x) ) [)]`` in NHWC: statistics pass + apply pass (+ their backward)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, pool, conv_bias=None):
        """``conv_bias``: bias of the convolution that produced ``x`` WITHOUT adding it (batch statistics only)."""
        x = _nhwc(x)
        _need_cuda(x, gamma, beta)
        B, C, H, W = x.shape
        dev = x.device
        if pool:
            y = torch.empty((B, C, H // 2, W // 2), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        else:
            y = torch.empty_like(x)
        stats = torch.empty((4, C), dtype=torch.float32, device=dev)  # save_mean, save_invstd, A, B
        world = stat_sync_world() if training else 1
        if conv_bias is not None and not training:
            raise ValueError("a folded convolution bias needs batch statistics (keep the bias in the convolution otherwise)")
        _bn_forward(x, gamma, beta, running_mean, running_var, training, momentum, eps, relu, pool, y, stats, world,
                    conv_bias)
        ctx.cfg = (bool(training), bool(relu), bool(pool), world, conv_bias is not None)
        ctx.save_for_backward(x, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        training, relu, pool, world, has_cb = ctx.cfg
        out = _bn_backward(_nhwc(dy), x, stats, training, relu, pool, ctx.needs_input_grad[0], world, has_cb)
        dx, dgamma, dbeta = out[:3]
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, (out[3] if has_cb else None)


def bn_supported(bn: torch.nn.Module, x: torch.Tensor) -> bool:
    """Shapes / configurations the fused BatchNorm kernels cover (anything else stays on ATen)."""
    return (isinstance(bn, torch.nn.BatchNorm2d) and bn.affine and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
            and x.shape[1] % 4 == 0 and 4 <= x.shape[1] <= 1024 and (bn.training or bn.running_mean is not None))


def conv_without_bias(conv: torch.nn.Conv2d, x: torch.Tensor, bn: torch.nn.Module):
    """``(conv(x) without its bias, bias)`` when ``bn`` -- which consumes the result next -- normalises with batch
    statistics: those cancel a per-channel shift exactly, so the reference's bias-add pass over the output and its
    backward reduction over all pixels are dead work; the BatchNorm kernels account for the bias where it matters (the
    running mean; its gradient).  Otherwise ``(conv(x), None)``."""
    C = getattr(conv, "out_channels", 0)
    if (isinstance(conv, torch.nn.Conv2d) and conv.bias is not None and isinstance(bn, torch.nn.BatchNorm2d) and bn.affine
            and (bn.training or bn.running_mean is None) and x.is_cuda and x.dtype == torch.float32
            and C % 4 == 0 and 4 <= C <= 1024 and bn.num_features == C):
        return conv._conv_forward(x, conv.weight, None), conv.bias
    return conv(x), None


def batch_norm_relu(x: torch.Tensor, bn: torch.nn.BatchNorm2d, relu: bool = True, pool: bool = False,
                    conv_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``maxpool2(relu(bn(x)))`` (each optional) with ``bn``'s parameters, mode and running statistics
    (updated in place like ``nn.BatchNorm2d.forward``).  ``conv_bias``: see ``conv_without_bias``."""
    use_batch_stats = bn.training or bn.running_mean is None
    _bump_batch_counter(bn)
    if bn.momentum is None:  # cumulative moving average
        momentum = 1.0 / float(bn.num_batches_tracked) if bn.training and bn.track_running_stats else 0.0
    else:
        momentum = bn.momentum
    return BNReLUFunction.apply(x, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                                bn.running_var if bn.track_running_stats else None, use_batch_stats, momentum,
                                bn.eps, relu, pool, conv_bias)


# --------------------------------------------------------------------------------------
# Bilinear x2 up-sampling fused with the concatenation behind it
# --------------------------------------------------------------------------------------
class UpsampleCatFunction(torch.autograd.Function):
    """``torch.cat((first, upsample2_bilinear_align_corners(x)), dim=1)`` in NHWC (mtan_model.py:143-145): the
    up-sampled half is written straight into its channel slice of the result, and in the backward its gradient is
    gathered straight out of that slice (no repacking copy); ``first`` gets its slice of the gradient as a view."""

    @staticmethod
    def forward(ctx, first, x):
        first, x = _nhwc(first), _nhwc(x)
        _need_cuda(first, x)
        B, C1, Ho, Wo = first.shape
        _, Cx, Hi, Wi = x.shape
        if (Ho, Wo) != (2 * Hi, 2 * Wi) or C1 % 4 or Cx % 4:
            raise ValueError("upsample_cat needs first [B,C1,2H,2W], x [B,Cx,H,W] with C1, Cx multiples of 4")
        out = torch.empty((B, C1 + Cx, Ho, Wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        out[:, :C1].copy_(first)
        _call("up2_bilinear_fwd", 20 * x.numel(), _p(x), ctypes.c_void_p(out.data_ptr() + 4 * C1), B, Hi, Wi, Cx,
              C1 + Cx, _stream())
        ctx.geom = (B, C1, Cx, Hi, Wi)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C1, Cx, Hi, Wi = ctx.geom
        dout = _nhwc(dout)
        dfirst = dout[:, :C1] if ctx.needs_input_grad[0] else None
        dx = None
        if ctx.needs_input_grad[1]:
            dx = torch.empty((B, Cx, Hi, Wi), dtype=torch.float32, device=dout.device, memory_format=torch.channels_last)
            _call("up2_bilinear_bwd", 20 * dx.numel(), ctypes.c_void_p(dout.data_ptr() + 4 * C1), C1 + Cx, _p(dx), B, Hi,
                  Wi, Cx, _stream())
        return dfirst, dx


def upsample2_cat(first: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``cat((first, bilinear_x2(x, align_corners=True)), dim=1)``; see ``UpsampleCatFunction``."""
    return UpsampleCatFunction.apply(first, x)


def upsample2_cat_supported(first: torch.Tensor, x: torch.Tensor) -> bool:
    return (first.is_cuda and x.is_cuda and first.dtype == torch.float32 and x.dtype == torch.float32
            and first.dim() == 4 and x.dim() == 4 and first.shape[0] == x.shape[0]
            and first.shape[2] == 2 * x.shape[2] and first.shape[3] == 2 * x.shape[3]
            and first.shape[1] % 4 == 0 and x.shape[1] % 4 == 0)


# --------------------------------------------------------------------------------------
# MTAN attention gate
# --------------------------------------------------------------------------------------
def _gate_bwd_passes(M: int, N: int):
    """(reads of h in pass 1, dh launches) of the tensor-core backward (csrc/gate_tc_bwd_tma.cuh): pass 1 is one
    launch whose CTAs split wide gates into 64-column chunks, each chunk reading h once (from L2 after the first)."""
    if N <= 64:
        return 1, 1
    one_tile = (M + 127) // 128 <= _lib.load().vmtl_sm_count()
    n_dh = -(-N // 256) if one_tile else N // 64
    return N // 64, n_dh


def gate_bytes(M: int, K: int, N: int, training: bool, backward: bool, stored_z: bool = True):
    """(algorithmic, issued) HBM bytes of one gate call.

    Algorithmic = SURVEY 8(d): forward ``4M(K + 2N)`` in eval mode, ``4M(K + 2N + min(K, 2N))`` with batch
    statistics (z must be stored or recomputed for the second phase); backward
    ``4M(2K + 4N + min(K, 2N))``.  Issued = what the kernels of this library request: the forward stores and
    re-reads z (``4M(K + 4N)``); the backward reads (dy, s, z) in the statistics + dW pass and again in the dh
    pass (``4M(2K + 7N)``; wide gates re-read h / re-add dh once per 64-column pass)."""
    if not backward:
        algo = 4 * M * (K + 2 * N + (min(K, 2 * N) if training else 0))
        issued = 4 * M * (K + (4 * N if training or stored_z else 2 * N))
        return algo, issued
    n_p1, n_dh = _gate_bwd_passes(M, N) if K == 128 else (1, 1)
    algo = 4 * M * (2 * K + 4 * N + min(K, 2 * N))
    issued = 4 * M * (7 * N + K * (n_p1 + n_dh))
    return algo, issued


def _gate_forward(h, h_coef, s, w2, bias, gamma, beta, running_mean, running_var, training, momentum, eps,
                  precision, need_z, world):
    """-> (y, z | None, mean, invstd); ``world > 1`` (training): the batch statistics of z are global."""
    B, K, H, W = h.shape
    N = s.shape[1]
    M = B * H * W
    dev = h.device
    y = torch.empty_like(s)
    z = torch.empty_like(s) if need_z else None
    st = torch.empty((2, N), dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = _workspace(lib.vmtl_gate_workspace_bytes(M, K, N, precision, 0), dev)
    algo, issued = gate_bytes(M, K, N, training, backward=False, stored_z=z is not None)
    if training and world > 1:
        mom = torch.empty((2, N), dtype=torch.float64, device=dev)
        _call("gate_fwd_moments", 4 * M * (K + N), _p(h), _p(h_coef), _p(w2), _p(bias.detach()), precision, M, K, N,
              _p(z), _p(mom), _p(ws), ws.numel(), _stream())
        _allreduce_moments(mom)
        _call("gate_fwd_global", algo - 4 * M * (K + N), _p(s), _p(z), _p(gamma.detach()), _p(beta.detach()),
              _p(running_mean), _p(running_var), float(momentum), float(eps), precision, M, K, N, _p(mom), M * world,
              _p(y), _p(st[0]), _p(st[1]), _p(ws), ws.numel(), _stream())
        _Prof.launches -= 1  # moments + global = contraction, 2 x finalize, gate pass
        return y, z, st[0], st[1]
    _call("gate_fwd", algo, _p(h), _p(h_coef), _p(s), _p(w2), _p(bias.detach()), _p(gamma.detach()),
          _p(beta.detach()), _p(running_mean), _p(running_var), float(momentum), float(eps),
          1 if training else 0, precision, M, K, N, _p(y), _p(z), _p(st[0]), _p(st[1]), _p(ws),
          ws.numel(), _stream(), issued=issued)
    return y, z, st[0], st[1]


def _gate_backward(dy, h, h_coef, s, z, w2, gamma, beta, mean, invstd, training, precision, need_dh, need_ds, world):
    """-> (dh | None, ds | None, dW [N,K], dbias, dgamma, dbeta)."""
    B, K, H, W = h.shape
    N = s.shape[1]
    M = B * H * W
    dev = h.device
    dh = torch.empty_like(h) if need_dh else None
    ds = torch.empty_like(s) if need_ds else None
    dW = torch.empty_like(w2)
    small = torch.empty((3, N), dtype=torch.float32, device=dev)  # dbias, dgamma, dbeta
    lib = _lib.load()
    ws = _workspace(lib.vmtl_gate_workspace_bytes(M, K, N, precision, 1), dev)
    algo, issued = gate_bytes(M, K, N, training, backward=True)
    args = (_p(dy), _p(h), _p(h_coef), _p(s), _p(z), _p(w2), _p(gamma.detach()), _p(beta.detach()), _p(mean),
            _p(invstd))
    if training and world > 1:
        mom = torch.empty((2, N), dtype=torch.float64, device=dev)
        _call("gate_bwd_moments", 4 * M * (K + 4 * N), *args, precision, M, K, N, _p(ds), _p(mom), _p(ws), ws.numel(),
              _stream())
        _allreduce_moments(mom)
        _call("gate_bwd_global", algo - 4 * M * (K + 4 * N), *args, precision, M, K, N, _p(mom), M * world, _p(dh),
              _p(dW), _p(small[0]), _p(small[1]), _p(small[2]), _p(ws), ws.numel(), _stream())
    else:
        _call("gate_bwd", algo, *args, 1 if training else 0, precision, M, K, N, _p(dh), _p(ds), _p(dW),
              _p(small[0]), _p(small[1]), _p(small[2]), _p(ws), ws.numel(), _stream(), issued=issued)
    if precision != GATE_FP32_FFMA and N > 64 and bool(lib.vmtl_gate_tc_supported(K, N)):
        # wider gates run extra dh launches: per 64 columns, or per 256 when every CTA owns a single 128-row tile
        _Prof.launches += _gate_bwd_passes(M, N)[1] - 1
    return dh, ds, dW, small[0], small[1], small[2]


class GateFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, s, weight, bias, gamma, beta, running_mean, running_var, training, momentum,
                eps, precision, h_coef=None):
        """``h_coef`` ([2,K], no gradient): ``h`` is then the hidden layer's pre-activation and the kernels fold
        ``max(A1 h + B1, 0)`` into their operand conversion; the backward's first gradient is w.r.t. that
        activation (``FoldedGateFunction`` chains the BatchNorm backward behind it)."""
        h = _nhwc(h)
        s = _nhwc(s)
        _need_cuda(h, s, weight)
        K, N = h.shape[1], s.shape[1]
        w2 = weight.detach().reshape(N, K).contiguous()
        # z = conv output is kept for the backward; pure inference on the tensor-core path
        # fuses everything into one pass and never stores it
        need_z = training or precision == GATE_FP32_FFMA or any(ctx.needs_input_grad)
        world = stat_sync_world() if training else 1
        y, z, mean, invstd = _gate_forward(h, h_coef, s, w2, bias, gamma, beta, running_mean, running_var, training,
                                           momentum, eps, precision, need_z, world)
        ctx.cfg = (bool(training), precision, weight.shape, world)
        ctx.save_for_backward(h, s, z, w2, gamma, beta, mean, invstd, *([h_coef] if h_coef is not None else []))
        return y

    @staticmethod
    def backward(ctx, dy):
        h, s, z, w2, gamma, beta, mean, invstd, *pre = ctx.saved_tensors
        training, precision, wshape, world = ctx.cfg
        dh, ds, dW, dbias, dgamma, dbeta = _gate_backward(
            _nhwc(dy), h, pre[0] if pre else None, s, z, w2, gamma, beta, mean, invstd, training, precision,
            ctx.needs_input_grad[0], ctx.needs_input_grad[1], world)
        return (dh, ds, dW.reshape(wshape), dbias, dgamma, dbeta, None, None, None, None, None, None, None)


class FoldedGateFunction(torch.autograd.Function):
    """``s * sigmoid(bn2(conv2( relu(bn1(c)) )))`` with the hidden activation ``relu(bn1(c))`` never written to
    HBM: one statistics pass over ``c`` (the 1x1 squeeze's output) produces bn1's folded
    coefficients, the gate kernels apply them while converting their operand, and the backward chains the gate
    backward (gradient w.r.t. the activation) into the fused BatchNorm + ReLU backward over ``c``."""

    @staticmethod
    def forward(ctx, c, g1, b1, rm1, rv1, train1, mom1, eps1, s, weight, bias, g2, b2, rm2, rv2, train2, mom2, eps2,
                precision, conv1_bias=None):
        c = _nhwc(c)
        s = _nhwc(s)
        _need_cuda(c, s, weight)
        K, N = c.shape[1], s.shape[1]
        world1 = stat_sync_world() if train1 else 1
        world2 = stat_sync_world() if train2 else 1
        stats1 = torch.empty((4, K), dtype=torch.float32, device=c.device)  # mean1, invstd1, A1, B1
        if conv1_bias is not None and not train1:
            raise ValueError("a folded convolution bias needs batch statistics")
        _bn_forward(c, g1, b1, rm1, rv1, train1, mom1, eps1, True, False, None, stats1, world1, conv1_bias)
        w2 = weight.detach().reshape(N, K).contiguous()
        need_z = train2 or any(ctx.needs_input_grad)
        y, z, mean2, invstd2 = _gate_forward(c, stats1[2:], s, w2, bias, g2, b2, rm2, rv2, train2, mom2, eps2,
                                             precision, need_z, world2)
        ctx.cfg = (bool(train1), bool(train2), precision, weight.shape, world1, world2, conv1_bias is not None)
        ctx.save_for_backward(c, stats1, s, z, w2, g2, b2, mean2, invstd2)
        return y

    @staticmethod
    def backward(ctx, dy):
        c, stats1, s, z, w2, g2, b2, mean2, invstd2 = ctx.saved_tensors
        train1, train2, precision, wshape, world1, world2, has_cb = ctx.cfg
        need_dc = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or has_cb
        dh, ds, dW, dbias, dg2, db2 = _gate_backward(_nhwc(dy), c, stats1[2:], s, z, w2, g2, b2, mean2, invstd2, train2,
                                                     precision, need_dc, ctx.needs_input_grad[8], world2)
        dc = dg1 = db1 = dcb = None
        if need_dc:  # relu + bn1 backward over (dh, c): statistics pass, finalize, apply pass
            out = _bn_backward(dh, c, stats1, train1, True, False, ctx.needs_input_grad[0], world1, has_cb)
            dc, dg1, db1 = out[:3]
            dcb = out[3] if has_cb else None
        return (dc, dg1, db1, None, None, None, None, None, ds, dW.reshape(wshape), dbias, dg2, db2,
                None, None, None, None, None, None, dcb)


def _bn_mode(bn):
    """(use batch statistics, momentum) of an ``nn.BatchNorm2d`` call; bumps ``num_batches_tracked``."""
    use_batch_stats = bn.training or bn.running_mean is None
    _bump_batch_counter(bn)
    if bn.momentum is None:
        momentum = 1.0 / float(bn.num_batches_tracked) if bn.training and bn.track_running_stats else 0.0
    else:
        momentum = bn.momentum
    return use_batch_stats, momentum


def folded_gate_supported(c: torch.Tensor, s: torch.Tensor, bn1, bn2, precision: Optional[str] = None) -> bool:
    prec = _GATE_PRECISIONS[precision or default_gate_precision]
    return (prec != GATE_FP32_FFMA and c.is_cuda and bn_supported(bn1, c) and bn2.affine
            and (bn2.training or bn2.running_mean is not None)
            and bool(_lib.load().vmtl_gate_tc_supported(c.shape[1], s.shape[1])))


def attention_gate_folded(c, bn1, s, conv2, bn2, precision: Optional[str] = None, conv1_bias=None):
    """The whole tail of an MTAN attention module after its 1x1 squeeze ``c = conv1(merged)``:
    ``s * sigmoid(bn2(conv2(relu(bn1(c)))))`` (mtan_model.py:66-75 / :153-162).  ``conv1_bias``: ``c`` was produced
    without its bias (``conv_without_bias``)."""
    t1, m1 = _bn_mode(bn1)
    t2, m2 = _bn_mode(bn2)
    rs = lambda bn, name: getattr(bn, name) if bn.track_running_stats else None  # noqa: E731
    return FoldedGateFunction.apply(c, bn1.weight, bn1.bias, rs(bn1, "running_mean"), rs(bn1, "running_var"), t1, m1,
                                    bn1.eps, s, conv2.weight, conv2.bias, bn2.weight, bn2.bias, rs(bn2, "running_mean"),
                                    rs(bn2, "running_var"), t2, m2, bn2.eps,
                                    _GATE_PRECISIONS[precision or default_gate_precision], conv1_bias)


def attention_gate(h, s, weight, bias, gamma, beta, running_mean, running_var, training: bool,
                   momentum: float = 0.1, eps: float = 1e-5, precision: Optional[str] = None):
    """``s * sigmoid(batch_norm(conv1x1(h)))`` in NHWC; updates running stats in place."""
    prec = _GATE_PRECISIONS[precision or default_gate_precision]
    return GateFunction.apply(h, s, weight, bias, gamma, beta, running_mean, running_var, training,
                              momentum, eps, prec)


# --------------------------------------------------------------------------------------
# Heads fused with losses / metrics
# --------------------------------------------------------------------------------------
def _loss_ws(P: int, device) -> torch.Tensor:
    return _workspace(_lib.load().vmtl_loss_workspace_bytes(P), device)


class HeadCEFunction(torch.autograd.Function):
    """1x1 seg head + mean cross-entropy (+ argmax, + confusion matrix accumulation)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, target, ignore_index, conf, want_pred):
        feat = _nhwc(feat)
        _need_cuda(feat, weight, target)
        B, Cin, H, W = feat.shape
        C = weight.shape[0]
        P = B * H * W
        w2 = weight.detach().reshape(C, Cin).contiguous()
        tgt = target.contiguous()
        if tgt.dtype != torch.int64:
            raise TypeError("segmentation target must be int64")
        out = torch.empty(2, dtype=torch.float64, device=feat.device)
        loss = torch.empty((), dtype=torch.float32, device=feat.device)
        pred = torch.empty((B, H, W), dtype=torch.uint8, device=feat.device) if want_pred else None
        ws = _loss_ws(P, feat.device)
        _call("head_ce_fwd", P * (4 * Cin + 8 + (1 if want_pred else 0)), _p(feat), _p(w2),
              _p(bias.detach()), _p(tgt), P, Cin, C, int(ignore_index), _p(out), _p(loss), _p(pred),
              _p(conf), _p(ws), ws.numel(), _stream())
        ctx.ignore_index = int(ignore_index)
        ctx.wshape = weight.shape
        ctx.save_for_backward(feat, w2, bias, tgt, out)
        if pred is None:
            pred = torch.empty(0, dtype=torch.uint8, device=feat.device)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, gloss, _gpred):
        feat, w2, bias, tgt, out = ctx.saved_tensors
        B, Cin, H, W = feat.shape
        C = w2.shape[0]
        P = B * H * W
        g = gloss.detach().to(torch.float32).contiguous()
        dfeat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(w2)
        db = torch.empty(C, dtype=torch.float32, device=feat.device)
        ws = _loss_ws(P, feat.device)
        _call("head_ce_bwd", P * (8 * Cin + 8), _p(feat), _p(w2), _p(bias.detach()), _p(tgt), P, Cin, C,
              ctx.ignore_index, _p(out), _p(g), _p(dfeat), _p(dW), _p(db), _p(ws), ws.numel(), _stream())
        return dfeat, dW.reshape(ctx.wshape), db, None, None, None, None


def head_cross_entropy(feat, weight, bias, target, ignore_index: int = -100,
                       conf: Optional[torch.Tensor] = None, want_pred: bool = True):
    """Returns (loss, pred uint8 [B,H,W]); ``conf`` (int64 [C,C]) is accumulated in place.

    Head shapes outside the fused kernels' set (Cin != 32, e.g. the reference-default ``encoder_first_channel=64``
    MTAN) run the 1x1 projection through cuDNN and the loss kernels on its logits -- still on the GPU, still one
    pass over the logits; more than 32 classes raise ``VmtlError``."""
    Cin, C = feat.shape[1], weight.shape[0]
    if not _lib.load().vmtl_head_supported(0, Cin, C):
        logits = torch.nn.functional.conv2d(_nhwc(feat), weight, bias)
        return cross_entropy_logits(logits, target, ignore_index, conf, want_pred)
    return HeadCEFunction.apply(feat, weight, bias, target, ignore_index, conf, want_pred)


def _logits_layout(logits: torch.Tensor):
    if logits.dim() != 4 or logits.dtype != torch.float32:
        raise TypeError("expected fp32 [B,C,H,W] logits")
    if logits.is_contiguous():
        return logits, LAYOUT_NCHW
    if logits.is_contiguous(memory_format=torch.channels_last):
        return logits, LAYOUT_NHWC
    return logits.contiguous(), LAYOUT_NCHW


class CELogitsFunction(torch.autograd.Function):
    """Mean cross-entropy (+ argmax, + confusion) on precomputed logits, NCHW or NHWC."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index, conf, want_pred):
        logits, layout = _logits_layout(logits)
        _need_cuda(logits, target)
        B, C, H, W = logits.shape
        P, HW = B * H * W, H * W
        tgt = target.contiguous()
        if tgt.dtype != torch.int64:
            raise TypeError("segmentation target must be int64")
        out = torch.empty(2, dtype=torch.float64, device=logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        pred = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device) if want_pred else None
        ws = _loss_ws(P, logits.device)
        _call("ce_logits_fwd", P * (4 * C + 8 + (1 if want_pred else 0)), _p(logits), _p(tgt), P, HW, C,
              layout, int(ignore_index), _p(out), _p(loss), _p(pred), _p(conf), _p(ws), ws.numel(),
              _stream())
        ctx.ignore_index = int(ignore_index)
        ctx.layout = layout
        ctx.save_for_backward(logits, tgt, out)
        if pred is None:
            pred = torch.empty(0, dtype=torch.uint8, device=logits.device)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, gloss, _gpred):
        logits, tgt, out = ctx.saved_tensors
        B, C, H, W = logits.shape
        P, HW = B * H * W, H * W
        g = gloss.detach().to(torch.float32).contiguous()
        dlogits = torch.empty_like(logits)
        _call("ce_logits_bwd", P * (8 * C + 8), _p(logits), _p(tgt), P, HW, C, ctx.layout,
              ctx.ignore_index, _p(out), _p(g), _p(dlogits), _stream())
        return dlogits, None, None, None, None


def cross_entropy_logits(logits, target, ignore_index: int = -100,
                         conf: Optional[torch.Tensor] = None, want_pred: bool = True):
    return CELogitsFunction.apply(logits, target, ignore_index, conf, want_pred)


class HeadSilogFunction(torch.autograd.Function):
    """1x1 depth head + sigmoid + SILog (+ MAE, abs-rel, predictions).

    ``weight is None``: ``feat`` is already the [B,1,H,W] depth logit (basic / csnet)."""

    @staticmethod
    def forward(ctx, feat, weight, bias, target, min_depth, want_pred):
        has_head = weight is not None
        feat = _nhwc(feat) if has_head else feat.contiguous()
        _need_cuda(feat, target)
        B, Cin, H, W = feat.shape
        P = B * H * W
        tgt = target.contiguous()
        if tgt.dtype != torch.float32 or tgt.numel() != P:
            raise TypeError("depth target must be fp32 with B*H*W elements")
        w1 = weight.detach().reshape(Cin).contiguous() if has_head else None
        out = torch.empty(8, dtype=torch.float64, device=feat.device)
        scalars = torch.empty(3, dtype=torch.float32, device=feat.device)
        pred = torch.empty((B, H, W, 1), dtype=torch.float32, device=feat.device) if want_pred else None
        ws = _loss_ws(P, feat.device)
        _call("head_silog_fwd", P * (4 * Cin + 4 + (4 if want_pred else 0)), _p(feat), _p(w1),
              _p(bias.detach() if has_head else None), _p(tgt), P, Cin, float(min_depth), _p(out),
              _p(scalars), _p(pred), _p(ws), ws.numel(), _stream())
        world = stat_sync_world()
        if world > 1:  # SILog of the GLOBAL batch: all-reduce the sums, re-derive mean / D / scalars
            _allreduce_moments(out)
            _call("silog_finalize", 0, _p(out), _p(scalars), _stream())
        ctx.world = world
        ctx.min_depth = float(min_depth)
        ctx.has_head = has_head
        ctx.wshape = weight.shape if has_head else None
        if has_head:
            ctx.save_for_backward(feat, tgt, out, w1, bias)
        else:
            ctx.save_for_backward(feat, tgt, out)
        if pred is None:
            pred = torch.empty(0, dtype=torch.float32, device=feat.device)
        ctx.mark_non_differentiable(pred, out)
        return scalars, pred, out

    @staticmethod
    def backward(ctx, gscalars, _gpred, _gout):
        gsilog = gscalars[0:1]  # only silog (scalars[0]) is differentiable; mae/abs_rel are metrics
        if ctx.has_head:
            feat, tgt, out, w1, bias = ctx.saved_tensors
        else:
            feat, tgt, out = ctx.saved_tensors
            w1 = bias = None
        B, Cin, H, W = feat.shape
        P = B * H * W
        g = gsilog.detach().to(torch.float32).contiguous()
        if ctx.world > 1:  # the data-parallel wrapper averages gradients; the global loss is not a per-replica mean
            g = g * float(ctx.world)
        need_dfeat = ctx.needs_input_grad[0] or not ctx.has_head
        dfeat = torch.empty_like(feat) if need_dfeat else None
        dw = torch.empty(Cin, dtype=torch.float32, device=feat.device) if ctx.has_head else None
        db = torch.empty(1, dtype=torch.float32, device=feat.device) if ctx.has_head else None
        ws = _loss_ws(P, feat.device)
        _call("head_silog_bwd", P * (8 * Cin + 4), _p(feat), _p(w1),
              _p(bias.detach() if bias is not None else None), _p(tgt), P, Cin, ctx.min_depth, _p(out),
              _p(g), _p(dfeat), _p(dw), _p(db), _p(ws), ws.numel(), _stream())
        if ctx.has_head:
            return dfeat, dw.reshape(ctx.wshape), db, None, None, None
        return dfeat, None, None, None, None, None


def head_silog(feat, weight, bias, target, min_depth: float = 1e-3, want_pred: bool = True,
               return_moments: bool = False):
    """Returns (silog, mae, abs_rel, pred [B,H,W,1]); only silog carries a gradient.  ``return_moments``
    appends the float64 [8] moment vector of ``vmtl_head_silog_fwd`` ({n, sum g, sum g^2, sum|p-t|,
    sum|p-t|/t, mean g, D, P}): what a data-parallel step all-reduces."""
    if weight is not None and not _lib.load().vmtl_head_supported(1, feat.shape[1], 1):
        # projection widths the fused kernel does not cover: cuDNN 1x1 conv, then the loss on the depth logit
        feat, weight, bias = torch.nn.functional.conv2d(_nhwc(feat), weight, bias), None, None
    scalars, pred, moments = HeadSilogFunction.apply(feat, weight, bias, target, min_depth, want_pred)
    out = (scalars[0], scalars[1].detach(), scalars[2].detach(), pred)
    return out + (moments,) if return_moments else out


# --------------------------------------------------------------------------------------
# Validation reductions
# --------------------------------------------------------------------------------------
def confusion_accumulate(pred: torch.Tensor, target: torch.Tensor, num_classes: int,
                         conf: Optional[torch.Tensor] = None, ignore_index: int = -100) -> torch.Tensor:
    """conf[target, pred] += 1 (int64 [C,C], rows = target); bit-exact integer arithmetic."""
    _need_cuda(pred, target)
    if conf is None:
        conf = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=pred.device)
    pred = pred.contiguous()
    target = target.contiguous()
    if target.dtype != torch.int64 or pred.dtype not in (torch.uint8, torch.int64):
        raise TypeError("pred must be uint8/int64 and target int64")
    P = target.numel()
    is_u8 = pred.dtype == torch.uint8
    _call("confusion_accum", P * (9 if is_u8 else 16), _p(pred), 1 if is_u8 else 0, _p(target), P,
          num_classes, int(ignore_index), _p(conf), _stream())
    return conf


def depth_error_sums(pred: torch.Tensor, target: torch.Tensor, min_depth: float = 1e-3) -> torch.Tensor:
    """float64 [4] = {P, sum|p-t|, n(t>min_depth), sum|p-t|/t}."""
    _need_cuda(pred, target)
    pred = pred.contiguous()
    target = target.contiguous()
    P = pred.numel()
    out = torch.empty(4, dtype=torch.float64, device=pred.device)
    ws = _loss_ws(P, pred.device)
    _call("depth_err_sums", P * 8, _p(pred), _p(target), P, float(min_depth), _p(out), _p(ws), ws.numel(),
          _stream())
    return out


def seg_metrics(conf: torch.Tensor) -> torch.Tensor:
    """float32 [3] = {accuracy (micro), jaccard (macro, absent=0), F1 (weighted)} on device."""
    _need_cuda(conf)
    C = conf.shape[0]
    m = torch.empty(3, dtype=torch.float32, device=conf.device)
    _call("seg_metrics", 8 * C * C, _p(conf.contiguous()), C, _p(m), _stream())
    return m
