"""Torch-facing wrapper of the stand-alone modulated convolution (C-ABI group 5 of include/lfp_sg2.h).

``ModConvPlan`` owns one native handle per ``model.ModulatedConv2d`` module and device; ``modulated_conv2d`` is the
differentiable call the module's ``forward`` makes: input ``[B, Cin, H, W]`` and style ``[B, style_dim]`` in, output
``[B, Cout, H', W']`` out, gradients to the input and the style (layer parameters are frozen constants, as on the
whole-synthesis path; src/model.py:169-302 is what it replaces).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import capi
from .torch_glue import ptr, require_cuda, stream_ptr


class ModConvPlan:
    def __init__(self, in_channel: int, out_channel: int, kernel_size: int, style_dim: int, demodulate: bool = True,
                 upsample: bool = False, blur_kernel: Sequence[float] = (1, 3, 3, 1), device=None):
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("ModConvPlan needs a CUDA device (no CPU fallback)")
        self.cin, self.cout, self.k, self.style_dim, self.upsample = in_channel, out_channel, kernel_size, style_dim, upsample
        self._h = C.c_void_p()
        taps = (C.c_float * len(blur_kernel))(*[float(v) for v in blur_kernel])
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_modconv_create(C.byref(self._h), in_channel, out_channel, kernel_size, style_dim,
                                                     1 if demodulate else 0, 1 if upsample else 0, taps, len(blur_kernel)),
                       "modconv_create")
        self._sig = None

    def __del__(self):
        try:
            if self._h:
                capi.lib().lfp_modconv_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load(self, weight: torch.Tensor, mod_weight: torch.Tensor, mod_bias: torch.Tensor) -> None:
        L = capi.lib()
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            keep = []
            for name, t in (("weight", weight), ("modulation.weight", mod_weight), ("modulation.bias", mod_bias)):
                t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                capi.check(L.lfp_modconv_set_param(self._h, name.encode(), ptr(t), t.numel(), st), "modconv_set_param")
            capi.check(L.lfp_modconv_finalize(self._h, st), "modconv_finalize")
            torch.cuda.current_stream(self.device).synchronize()

    def sync(self, weight, mod_weight, mod_bias) -> None:
        sig = tuple((t.data_ptr(), t._version) for t in (weight, mod_weight, mod_bias))
        if sig != self._sig:
            self.load(weight, mod_weight, mod_bias)
            self._sig = sig

    def out_size(self, h: int, w: int):
        return (2 * h, 2 * w) if self.upsample else (h, w)

    def new_workspace(self, batch: int, h: int, w: int) -> torch.Tensor:
        n = int(capi.lib().lfp_modconv_workspace_bytes(self._h, batch, h, w))
        return torch.empty(n + 256, dtype=torch.uint8, device=self.device)

    @staticmethod
    def _aligned(ws: torch.Tensor) -> int:
        return (ws.data_ptr() + 255) // 256 * 256

    def forward(self, x: torch.Tensor, style: torch.Tensor, ws: torch.Tensor, precision: int) -> torch.Tensor:
        require_cuda(x, "input")
        require_cuda(style, "style")
        B, cin, H, W = x.shape
        if cin != self.cin or tuple(style.shape) != (B, self.style_dim):
            raise RuntimeError(f"modulated conv expects input [B, {self.cin}, H, W] and style [B, {self.style_dim}], got "
                               f"{tuple(x.shape)} and {tuple(style.shape)}")
        x = x.to(torch.float32).contiguous()
        style = style.to(torch.float32).contiguous()
        oh, ow = self.out_size(H, W)
        out = torch.empty((B, self.cout, oh, ow), dtype=torch.float32, device=self.device)
        base = self._aligned(ws)
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_modconv_forward(self._h, B, H, W, ptr(x), ptr(style), ptr(out), base,
                                                      ws.data_ptr() + ws.numel() - base, precision, stream_ptr(self.device)),
                       "modconv_forward")
        self._keep = (x, style)
        self._last_ws = ws
        return out

    def backward(self, d_out: torch.Tensor, shape, ws: torch.Tensor, precision: int, need_dx: bool = True, need_ds: bool = True):
        B, _, H, W = shape
        d_out = d_out.to(torch.float32).contiguous()
        dx = torch.empty(shape, dtype=torch.float32, device=self.device) if need_dx else None
        ds = torch.empty((B, self.style_dim), dtype=torch.float32, device=self.device) if need_ds else None
        base = self._aligned(ws)
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_modconv_backward(self._h, B, H, W, ptr(d_out), ptr(dx), ptr(ds), base,
                                                       ws.data_ptr() + ws.numel() - base, precision, stream_ptr(self.device)),
                       "modconv_backward")
        return dx, ds


class _ModConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, style, plan, precision):
        ws = plan.new_workspace(x.shape[0], x.shape[2], x.shape[3])
        out = plan.forward(x, style, ws, precision)
        ctx.plan, ctx.ws, ctx.precision, ctx.shape = plan, ws, precision, tuple(x.shape)
        ctx.x_dtype, ctx.s_dtype = x.dtype, style.dtype
        ctx.save_for_backward(x, style)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, style = ctx.saved_tensors
        plan = ctx.plan
        # one forward in flight per handle: if another forward ran on the plan since, redo this one on its own workspace
        if getattr(plan, "_last_ws", None) is not ctx.ws:
            plan.forward(x, style, ctx.ws, ctx.precision)
        dx, ds = plan.backward(d_out, ctx.shape, ctx.ws, ctx.precision, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return (dx.to(ctx.x_dtype) if dx is not None else None, ds.to(ctx.s_dtype) if ds is not None else None, None, None)


def modulated_conv2d(plan: ModConvPlan, x: torch.Tensor, style: torch.Tensor, precision: int) -> torch.Tensor:
    return _ModConv.apply(x, style, plan, precision)
