"""``op.upfirdn2d`` - public surface of src/op/upfirdn2d.py:149-165 over the C ABI.

``upfirdn2d(input[N,C,H,W], kernel[kh,kw], up=1, down=1, pad=(0,0))``: ``up``/``down`` int or
``(x, y)``; ``pad`` ``(p0, p1)`` (both axes) or ``(x0, x1, y0, y1)``; negative pads crop.
Differentiable w.r.t. ``input`` to any order: the gradient of an upfirdn2d is another upfirdn2d
with the flipped kernel, ``up``/``down`` swapped and the pads of src/op/upfirdn2d.py:112-115.
CPU tensors are rejected (no CPU fallback in this package).
"""
from __future__ import annotations

from collections import abc

import torch
from torch.autograd import Function

from lfp_native import capi
from lfp_native.torch_glue import dtype_code, ptr, require_cuda, stream_ptr


def _native(x4: torch.Tensor, kernel: torch.Tensor, up, down, pad) -> torch.Tensor:
    """The native module's ``upfirdn2d(input[major,in_h,in_w,minor], kernel, ...)``
    (src/op/upfirdn2d.cpp:17-27)."""
    require_cuda(x4, "input")
    require_cuda(kernel, "kernel")
    if not x4.is_contiguous():
        raise RuntimeError("input must be contiguous")
    if not kernel.is_contiguous():
        raise RuntimeError("kernel must be contiguous")
    if kernel.dtype != x4.dtype:
        kernel = kernel.to(x4.dtype)
    major, in_h, in_w, minor = x4.shape
    kh, kw = kernel.shape
    L = capi.lib()
    import ctypes as C
    oh, ow = C.c_int(), C.c_int()
    capi.check(L.lfp_upfirdn2d_out_size(in_h, in_w, kh, kw, up[0], up[1], down[0], down[1], *pad,
                                        C.byref(oh), C.byref(ow)), "upfirdn2d")
    out = torch.empty((major, max(oh.value, 0), max(ow.value, 0), minor), dtype=x4.dtype, device=x4.device)
    if out.numel() == 0 or x4.numel() == 0:
        return out.zero_()
    with torch.cuda.device(x4.device):
        capi.check(L.lfp_upfirdn2d(ptr(x4), ptr(kernel), ptr(out), dtype_code(x4), major, in_h, in_w, minor,
                                   kh, kw, up[0], up[1], down[0], down[1], *pad, stream_ptr(x4.device)),
                   "upfirdn2d")
    return out


class UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        n, c, in_h, in_w = input.shape
        kh, kw = kernel.shape
        out = _native(input.reshape(-1, in_h, in_w, 1), kernel, up, down, pad)
        out_h, out_w = out.shape[1], out.shape[2]
        # pads of the adjoint (src/op/upfirdn2d.py:112-115)
        ctx.g_pad = (kw - pad[0] - 1, in_w * up[0] - out_w * down[0] + pad[0] - up[0] + 1,
                     kh - pad[2] - 1, in_h * up[1] - out_h * down[1] + pad[2] - up[1] + 1)
        ctx.up, ctx.down = up, down
        ctx.save_for_backward(kernel)
        return out.view(n, c, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        (kernel,) = ctx.saved_tensors
        grad_input = None
        if ctx.needs_input_grad[0]:
            grad_input = UpFirDn2d.apply(grad_output.contiguous(), torch.flip(kernel, [0, 1]), ctx.down, ctx.up,
                                         ctx.g_pad)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    require_cuda(input, "input")
    return UpFirDn2d.apply(input.contiguous(), kernel.contiguous(), tuple(int(u) for u in up),
                           tuple(int(d) for d in down), tuple(int(p) for p in pad))
