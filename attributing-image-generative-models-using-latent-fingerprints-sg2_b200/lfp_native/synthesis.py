"""Torch-facing wrapper of the whole-synthesis plan (C-ABI group 3 of include/lfp_sg2.h).

``SynthesisPlan`` owns one native plan per (generator, device); ``synthesize`` is the
differentiable call used by ``model.Generator.forward``: latent ``[B, n_latent, 512]`` and the
noise list in, image ``[B, 3, S, S]`` out, gradient to the latent only (generator parameters are
treated as frozen, as in the attribution loop, src/main.py:58).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import capi
from .torch_glue import ptr, require_cuda, stream_ptr

_PARAM_SUFFIXES = (".conv.weight", ".conv.modulation.weight", ".conv.modulation.bias", ".noise.weight",
                   ".activate.bias", ".bias")


def plan_param_names(size: int) -> List[str]:
    """state_dict names the native plan consumes."""
    import math
    names = ["input.input"]
    n_blocks = int(math.log2(size)) - 2
    for p in ["conv1"] + [f"convs.{i}" for i in range(2 * n_blocks)]:
        names += [p + s for s in _PARAM_SUFFIXES[:5]]
    for p in ["to_rgb1"] + [f"to_rgbs.{j}" for j in range(n_blocks)]:
        names += [p + ".conv.weight", p + ".conv.modulation.weight", p + ".conv.modulation.bias", p + ".bias"]
    return names


class SynthesisPlan:
    def __init__(self, size: int, style_dim: int = 512, channel_multiplier: int = 2,
                 blur_kernel: Sequence[float] = (1, 3, 3, 1), device=None):
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("SynthesisPlan needs a CUDA device (no CPU fallback)")
        self.size, self.style_dim = size, style_dim
        self._h = C.c_void_p()
        taps = (C.c_float * len(blur_kernel))(*[float(v) for v in blur_kernel])
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_synth_create(C.byref(self._h), size, style_dim, channel_multiplier, taps,
                                                   len(blur_kernel)), "synth_create")
        self.n_latent = capi.lib().lfp_synth_n_latent(self._h)
        self.num_noise = capi.lib().lfp_synth_num_noise(self._h)
        self.generation = 0          # bumped by every forward; backward checks it
        self._ws_cache: Optional[torch.Tensor] = None
        self._sig = None

    def __del__(self):
        try:
            if self._h:
                capi.lib().lfp_synth_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- parameters -----------------------------------------------------------------------
    def load(self, params: Dict[str, torch.Tensor]) -> None:
        """Upload tensors keyed by reference state_dict names and rebuild the derived tables."""
        L = capi.lib()
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            keep = []
            for name in plan_param_names(self.size):
                if name not in params:
                    raise KeyError(f"generator parameter '{name}' missing")
                t = params[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                capi.check(L.lfp_synth_set_param(self._h, name.encode(), ptr(t), t.numel(), st), "synth_set_param")
            capi.check(L.lfp_synth_finalize(self._h, st), "synth_finalize")
            torch.cuda.current_stream(self.device).synchronize()  # staged copies are done with `keep`

    def sync_from_module(self, module: torch.nn.Module) -> None:
        """Re-upload when any consumed parameter changed (version counter or storage)."""
        sd = {k: v for k, v in module.named_parameters()}
        names = plan_param_names(self.size)
        sig = tuple((sd[n].data_ptr(), sd[n]._version) for n in names)
        if sig != self._sig:
            self.load({n: sd[n] for n in names})
            self._sig = sig

    # ---- execution ------------------------------------------------------------------------
    def workspace_bytes(self, batch: int, forward_only: bool = False) -> int:
        L = capi.lib()
        return int(L.lfp_synth_generate_workspace_bytes(self._h, batch) if forward_only
                   else L.lfp_synth_workspace_bytes(self._h, batch))

    def new_workspace(self, batch: int, forward_only: bool = False) -> torch.Tensor:
        return torch.empty(self.workspace_bytes(batch, forward_only) + 256, dtype=torch.uint8, device=self.device)

    def shared_workspace(self, batch: int) -> torch.Tensor:
        """Grow-only scratch for forward-only calls (lfp_synth_generate): nothing is kept for a backward."""
        need = self.workspace_bytes(batch, forward_only=True) + 256
        if self._ws_cache is None or self._ws_cache.numel() < need:
            self._ws_cache = None
            self._ws_cache = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws_cache

    @staticmethod
    def _aligned(ws: torch.Tensor) -> int:
        return (ws.data_ptr() + 255) // 256 * 256

    def generate(self, latent: torch.Tensor, noise: Sequence[torch.Tensor], precision: int = capi.PREC_FP32) -> torch.Tensor:
        """Forward only, on the plan's shared forward-only workspace (fingerprinted generation, target rendering)."""
        return self.forward(latent, noise, self.shared_workspace(latent.shape[0]), precision, forward_only=True)

    def forward(self, latent: torch.Tensor, noise: Sequence[torch.Tensor], ws: torch.Tensor,
                precision: int = capi.PREC_FP32, forward_only: bool = False) -> torch.Tensor:
        require_cuda(latent, "latent")
        B = latent.shape[0]
        if tuple(latent.shape) != (B, self.n_latent, self.style_dim):
            raise RuntimeError(f"latent must be [B, {self.n_latent}, {self.style_dim}], got {tuple(latent.shape)}")
        if len(noise) != self.num_noise:
            raise RuntimeError(f"expected {self.num_noise} noise maps, got {len(noise)}")
        latent = latent.to(torch.float32).contiguous()
        nz = []
        for i, n in enumerate(noise):
            res = 4 if i == 0 else 8 << ((i - 1) // 2)
            require_cuda(n, f"noise[{i}]")
            n = n.to(torch.float32).contiguous()
            if n.numel() not in (res * res, B * res * res):
                raise RuntimeError(f"noise[{i}] must be [1 or {B}, 1, {res}, {res}], got {tuple(n.shape)}")
            nz.append(n)
        image = torch.empty((B, 3, self.size, self.size), dtype=torch.float32, device=self.device)
        nptr = (C.c_void_p * self.num_noise)(*[n.data_ptr() for n in nz])
        nb = (C.c_int * self.num_noise)(*[n.numel() // ((4 if i == 0 else 8 << ((i - 1) // 2)) ** 2)
                                          for i, n in enumerate(nz)])
        base = self._aligned(ws)
        fn = capi.lib().lfp_synth_generate if forward_only else capi.lib().lfp_synth_forward
        with torch.cuda.device(self.device):
            capi.check(fn(self._h, B, ptr(latent), nptr, nb, ptr(image), base, ws.data_ptr() + ws.numel() - base, precision,
                          stream_ptr(self.device)), "synth_generate" if forward_only else "synth_forward")
        self.generation += 1
        self._keep = (latent, nz)  # pointers recorded by the plan must outlive the backward
        return image

    def read_activations(self, batch: int, ws: torch.Tensor) -> List[torch.Tensor]:
        """Saved StyledConv outputs of the last forward on ``ws`` as [B, C, res, res] tensors."""
        L = capi.lib()
        outs = []
        base = self._aligned(ws)
        with torch.cuda.device(self.device):
            for i in range(L.lfp_synth_num_convs(self._h)):
                ch, res = C.c_int(), C.c_int()
                capi.check(L.lfp_synth_read_activation(self._h, batch, i, None, None, C.byref(ch), C.byref(res), None))
                t = torch.empty((batch, ch.value, res.value, res.value), dtype=torch.float32, device=self.device)
                capi.check(L.lfp_synth_read_activation(self._h, batch, i, base, ptr(t), None, None,
                                                       stream_ptr(self.device)), "read_activation")
                outs.append(t)
        return outs

    def backward(self, d_image: torch.Tensor, batch: int, ws: torch.Tensor,
                 precision: int = capi.PREC_FP32) -> torch.Tensor:
        d_image = d_image.to(torch.float32).contiguous()
        d_latent = torch.empty((batch, self.n_latent, self.style_dim), dtype=torch.float32, device=self.device)
        base = self._aligned(ws)
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_synth_backward(self._h, batch, ptr(d_image), ptr(d_latent), base,
                                                     ws.data_ptr() + ws.numel() - base, precision,
                                                     stream_ptr(self.device)), "synth_backward")
        return d_latent


class _Synthesize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent, plan, precision, *noise):
        needs_grad = latent.requires_grad
        ws = plan.new_workspace(latent.shape[0]) if needs_grad else plan.shared_workspace(latent.shape[0])
        image = plan.forward(latent, noise, ws, precision, forward_only=not needs_grad)
        if needs_grad:
            ctx.plan, ctx.ws, ctx.precision = plan, ws, precision
            ctx.generation = plan.generation
            ctx.batch = latent.shape[0]
            ctx.save_for_backward(latent, *noise)
        return image

    @staticmethod
    def backward(ctx, d_image):
        plan = ctx.plan
        latent, *noise = ctx.saved_tensors
        if plan.generation != ctx.generation:
            # another forward ran on this plan in between: rebuild this call's activations
            plan.forward(latent, noise, ctx.ws, ctx.precision)
        d_latent = plan.backward(d_image, ctx.batch, ctx.ws, ctx.precision)
        plan.generation += 1  # the workspace of this call is spent
        return (d_latent.to(latent.dtype), None, None) + (None,) * len(noise)


def synthesize(plan: SynthesisPlan, latent: torch.Tensor, noise: Sequence[torch.Tensor],
               precision: int = capi.PREC_FP32) -> torch.Tensor:
    return _Synthesize.apply(latent, plan, precision, *noise)
