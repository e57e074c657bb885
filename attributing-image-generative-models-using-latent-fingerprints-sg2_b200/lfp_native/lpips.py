"""Torch-facing wrapper of the native perceptual loss (C-ABI group 7 of include/lfp_sg2.h): LPIPS v0.1 / VGG16,
forward + backward to the estimated image with cached target features (src/custom_lpips/networks_basic.py:27-91,
src/utils.py:44-50)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import capi
from .torch_glue import ptr, require_cuda, stream_ptr

VGG_FEATURE_IDX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)
VGG_SLICE = (1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5)


def lpips_param_names():
    names = []
    for s, i in zip(VGG_SLICE, VGG_FEATURE_IDX):
        names += [f"net.slice{s}.{i}.weight", f"net.slice{s}.{i}.bias"]
    return names + [f"lin{k}.model.1.weight" for k in range(5)]


class LpipsPlan:
    def __init__(self, height: int, width: int = None, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("LpipsPlan needs a CUDA device (no CPU fallback)")
        self.h, self.w = height, width if width is not None else height
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_lpips_create(C.byref(self._h), self.h, self.w), "lpips_create")
        self._ws = None
        self.target_batch = 0

    def __del__(self):
        try:
            if self._h:
                capi.lib().lfp_lpips_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load(self, params: Dict[str, torch.Tensor]) -> None:
        """``params``: PNetLin state_dict entries (backbone convs + linear heads)."""
        L = capi.lib()
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            keep = []
            for name in lpips_param_names():
                if name not in params:
                    raise KeyError(f"LPIPS parameter '{name}' missing")
                t = params[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                capi.check(L.lfp_lpips_set_param(self._h, name.encode(), ptr(t), t.numel(), st), "lpips_set_param")
            capi.check(L.lfp_lpips_finalize(self._h, st), "lpips_finalize")
            torch.cuda.current_stream(self.device).synchronize()

    def workspace(self, batch: int) -> torch.Tensor:
        need = int(capi.lib().lfp_lpips_workspace_bytes(self._h, batch)) + 256
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _call(self, fn, what, batch, *args):
        ws = self.workspace(batch)
        base = (ws.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(self.device):
            capi.check(fn(self._h, batch, *args, base, ws.data_ptr() + ws.numel() - base), what)

    def set_target(self, target: torch.Tensor, precision: int) -> None:
        require_cuda(target, "target")
        target = target.to(torch.float32).contiguous()
        if tuple(target.shape[1:]) != (3, self.h, self.w):
            raise RuntimeError(f"target must be [B, 3, {self.h}, {self.w}], got {tuple(target.shape)}")
        L = capi.lib()
        self._call(lambda h, b, t, base, n: L.lfp_lpips_set_target(h, b, t, base, n, precision, stream_ptr(self.device)),
                   "lpips_set_target", target.shape[0], ptr(target))
        self.target_batch = target.shape[0]

    def loss_grad(self, est: torch.Tensor, precision: int, need_grad: bool = True):
        """(loss [B], d loss / d est [B, 3, H, W] or None)."""
        require_cuda(est, "est")
        est = est.to(torch.float32).contiguous()
        B = est.shape[0]
        if tuple(est.shape[1:]) != (3, self.h, self.w):
            raise RuntimeError(f"est must be [B, 3, {self.h}, {self.w}], got {tuple(est.shape)}")
        loss = torch.empty(B, dtype=torch.float32, device=self.device)
        d_est = torch.empty_like(est) if need_grad else None
        L = capi.lib()
        self._call(lambda h, b, e, lo, de, base, n: L.lfp_lpips_loss_grad(h, b, e, lo, de, base, n, precision, stream_ptr(self.device)),
                   "lpips_loss_grad", B, ptr(est), ptr(loss), ptr(d_est))
        return loss, d_est
