"""fp32-grade convolutions on the TF32 tensor cores: the 3xTF32 split of the gate kernels (csrc/gate_tc.cu) applied
to cuDNN's convolutions.  EXPERIMENTAL, off by default (``enable()`` / ``bench.py --conv-3xtf32``).

cuDNN's strict-fp32 convolution path is 87 % of the MTAN step (FFT convolutions + SIMT complex GEMMs + NCHW<->NHWC
transposes, ``profiles/r2_step_breakdown_mtan.txt``); its TF32 path runs on the tensor cores but rounds both
operands to 11 mantissa bits (1e-3 relative error: outside the 1e-4 parity bar).  With ``a = a_hi + a_lo``
(``a_hi`` exactly representable in TF32, ``a_lo`` the exact remainder)

    conv(x, w) = conv(x_hi, w_hi) + conv(x_lo, w_hi) + conv(x_hi, w_lo) + O(2^-22),

and because a convolution is linear in its input channels the three terms are ONE TF32 convolution over the
channel-stacked operands ``[x_hi | x_lo | x_hi]`` and ``[w_hi | w_hi | w_lo]`` -- the products are exact (TF32 x TF32
fits fp32), the accumulation is fp32 inside cuDNN.  Backward the same way: ``dx`` stacks along the OUTPUT channels
(``conv_transpose2d([dy_hi | dy_lo | dy_hi], [w_hi ; w_hi ; w_lo])``), ``dW`` along the BATCH (the weight gradient
sums over it).  Stride-1, groups-1 ``nn.Conv2d`` only (every convolution of the MTAN network); anything else keeps
its stock path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_state = {"enabled": False, "orig": None}


def tf32_split(x: torch.Tensor):
    """(hi, lo): hi = x rounded to TF32 (nearest), exactly representable, lo = x - hi exactly."""
    hi = ((x.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
    return hi, x - hi


class _TF32:
    def __enter__(self):
        self.prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32 = self.prev


class Conv3xTF32Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, padding, dilation):
        xh, xl = tf32_split(x)
        wh, wl = tf32_split(w)
        with _TF32():
            y = F.conv2d(torch.cat((xh, xl, xh), 1), torch.cat((wh, wh, wl), 1), bias, 1, padding, dilation, 1)
        ctx.save_for_backward(x, w)
        ctx.cfg = (padding, dilation, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        padding, dilation, has_bias = ctx.cfg
        dyh, dyl = tf32_split(dy.contiguous(memory_format=torch.channels_last))
        dx = dw = db = None
        with _TF32():
            if ctx.needs_input_grad[0]:
                wh, wl = tf32_split(w)
                dx = F.conv_transpose2d(torch.cat((dyh, dyl, dyh), 1), torch.cat((wh, wh, wl), 0), None, 1, padding, 0, 1,
                                        dilation)
            if ctx.needs_input_grad[1]:
                xh, xl = tf32_split(x)
                dw = torch.nn.grad.conv2d_weight(torch.cat((xh, xh, xl), 0), w.shape, torch.cat((dyh, dyl, dyh), 0), 1,
                                                 padding, dilation, 1)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 2, 3))
        return dx, dw, db, None, None


def _conv_forward(self, input, weight, bias):
    if (_state["enabled"] and input.is_cuda and input.dtype == torch.float32 and self.groups == 1
            and self.stride == (1, 1) and self.padding_mode == "zeros" and not isinstance(self.padding, str)
            and self.in_channels >= 8):
        return Conv3xTF32Function.apply(input, weight, bias, self.padding, self.dilation)
    return _state["orig"](self, input, weight, bias)


def enable(on: bool = True) -> None:
    """Route every eligible ``nn.Conv2d`` through the 3xTF32 scheme (process-wide)."""
    if _state["orig"] is None:
        _state["orig"] = torch.nn.Conv2d._conv_forward
        torch.nn.Conv2d._conv_forward = _conv_forward
    _state["enabled"] = bool(on)
