"""``op.fused_leaky_relu`` / ``op.FusedLeakyReLU`` - public surface of src/op/fused_act.py:93-127
over the C ABI: ``leaky_relu(input + bias[c], negative_slope) * scale`` with the bias broadcast
over dim 1; backward uses the saved *output* (``out > 0``), as the reference does (:29-31, :73).
CPU tensors are rejected (no CPU fallback in this package).
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from lfp_native import capi
from lfp_native.torch_glue import dtype_code, ptr, require_cuda, stream_ptr


def _native(x: torch.Tensor, bias, ref, act: int, grad: int, alpha: float, scale: float) -> torch.Tensor:
    """The native module's ``fused_bias_act(input, bias, refer, act, grad, alpha, scale)``
    (src/op/fused_bias_act.cpp:18-28); ``None`` stands for the reference's empty tensor."""
    require_cuda(x, "input")
    if not x.is_contiguous():
        raise RuntimeError("input must be contiguous")
    if bias is not None and bias.numel() == 0:
        bias = None
    if ref is not None and ref.numel() == 0:
        ref = None
    if bias is not None:
        require_cuda(bias, "bias")
        bias = bias.to(x.dtype).contiguous()
    if ref is not None:
        ref = ref.to(x.dtype).contiguous()
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    step_b = 1
    for d in x.shape[2:]:
        step_b *= d
    size_b = bias.numel() if bias is not None else 0
    with torch.cuda.device(x.device):
        capi.check(capi.lib().lfp_fused_bias_act(ptr(x), ptr(bias), ptr(ref), ptr(out), dtype_code(x), x.numel(),
                                                 step_b, size_b, act, grad, float(alpha), float(scale),
                                                 stream_ptr(x.device)), "fused_bias_act")
    return out


def _grad_bias(grad_input: torch.Tensor) -> torch.Tensor:
    """Sum over every dim but 1 (src/op/fused_act.py:34-40); deterministic native reduction in fp32."""
    dims = [0] + list(range(2, grad_input.ndim))
    # the native reduction covers fp32, non-empty tensors and up to 65535 channels (its grid.y); everything else
    # takes the reference's own expression, grad_input.sum(dim)
    if (grad_input.dtype != torch.float32 or grad_input.ndim < 2 or grad_input.numel() == 0
            or grad_input.shape[1] > 65535):
        return grad_input.sum(dims)
    outer, size_b = grad_input.shape[0], grad_input.shape[1]
    step_b = grad_input.numel() // max(outer * size_b, 1)
    L = capi.lib()
    nbytes = L.lfp_bias_grad_reduce_scratch(outer, size_b, step_b)
    scratch = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=grad_input.device)
    out = torch.empty(size_b, dtype=torch.float32, device=grad_input.device)
    with torch.cuda.device(grad_input.device):
        capi.check(L.lfp_bias_grad_reduce(ptr(grad_input), ptr(out), outer, size_b, step_b, ptr(scratch), nbytes,
                                          stream_ptr(grad_input.device)), "bias_grad_reduce")
    return out


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = _native(grad_output.contiguous(), None, out, 3, 1, negative_slope, scale)
        grad_bias = _grad_bias(grad_input).detach() if has_bias else grad_output.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        gradgrad_out = _native(gradgrad_input.contiguous(), gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.has_bias = bias is not None
        out = _native(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.has_bias,
                                                                     ctx.negative_slope, ctx.scale)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    require_cuda(input, "input")
    return FusedLeakyReLUFunction.apply(input.contiguous(), bias, negative_slope, scale)
