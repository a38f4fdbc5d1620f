"""Convolution dispatch for the RSSFormer path (NHWC activations, fp32 master weights).

Two engines, chosen per layer shape (DESIGN.md lists which layer uses which):
  * `igemm`  — the hand-written tcgen05/TMA implicit-GEMM kernels of csrc/conv_igemm.cu (bf16 operands,
               fp32 TMEM accumulation) through the C ABI;
  * `lib`    — ATen's convolution (cuDNN) for the shapes the igemm kernels do not cover yet.
Both take the low-precision copy of the weight from the optimiser's bf16 shadow buffer when one is
registered (no per-call cast kernels), and return fp32 weight gradients to the master parameter.
"""
import torch

from . import ops

CL = torch.channels_last
_SHADOW = {}          # id(param) -> bf16 view kept fresh by the fused optimiser step (trainer.py)
ENGINE = {"igemm": False}


def register_shadow(param, view):
    _SHADOW[id(param)] = view


def clear_shadows():
    _SHADOW.clear()


def _lowp(w, dtype):
    if w is None or w.dtype == dtype:
        return w
    s = _SHADOW.get(id(w))
    if s is not None and s.dtype == dtype:
        return s.view(w.shape)
    return w.detach().to(dtype)


class _ConvLib(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation):
        x = ops.nhwc(x)
        w = _lowp(weight, x.dtype).contiguous(memory_format=CL)
        b = _lowp(bias, x.dtype)
        y = torch.ops.aten.convolution(x, w, b, [stride, stride], [padding, padding], [dilation, dilation], False, [0, 0], 1)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, dilation, bias is not None, weight.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding, dilation, has_bias, wdtype = ctx.cfg
        dy = ops.nhwc(dy)
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        mask = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]]
        dx, dw, db = torch.ops.aten.convolution_backward(dy, x, w, [w.shape[0]] if has_bias else None, [stride, stride],
                                                         [padding, padding], [dilation, dilation], False, [0, 0], 1, mask)
        if dw is not None:
            dw = dw.to(wdtype)
        if db is not None:
            db = db.to(wdtype)
        return dx, dw, db, None, None, None


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1):
    return _ConvLib.apply(x, weight, bias, stride, padding, dilation)
