"""Batched latent-optimisation engine: the reference's hot loop (``optimization`` in
src/main.py:45-89) for B trajectories at once on one GPU, every step through the native path.

One trajectory = one (image, Latin-hypercube guess) pair.  Trajectories never interact: Adam
state is per element, the lr schedule ``lr * exp(-0.001 (i+1))`` (src/main.py:42-43) is shared,
the loss is per trajectory, and the native kernels reduce in a fixed order per sample, so a
trajectory's result is independent of what else is in its batch or on which rank it runs.

Per step (src/main.py:58-70):
    w0 = U^T alpha + mu ; wx = w0 + sd V^T diag(sigma) sigmoid(key)      lfp_embed_forward
    est = G(wx, noise)                                                   lfp_synth_forward
    loss = MSE(target, est) + 0.1 * alpha_bound(alpha)                   lfp_mse_loss_grad, lfp_attrib_bound_loss
    backward to (alpha, key) + Adam(lr_i)                                lfp_synth_backward, lfp_attrib_adam_update
Losses (src/utils.py:44-50): ``loss="mse"`` is the reference's own MSE alternative (:46-47); ``loss="lpips"`` is its
default perceptual loss, LPIPS v0.1 / VGG16, on the native kernels (lfp_lpips_*: target features cached per image,
forward + backward to the image; backbone weights are whatever ``lpips_params`` holds - no pretrained VGG offline).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from lfp_native import capi
from lfp_native.synthesis import SynthesisPlan
from lfp_native.torch_glue import ptr, stream_ptr


def get_lr(step: int, lr0: float = 0.2) -> float:
    """src/main.py:42-43."""
    return lr0 * math.exp(-0.001 * (step + 1))


class AttributionEngine:
    def __init__(self, plan: SynthesisPlan, noise: Sequence[torch.Tensor], pc: torch.Tensor,
                 sigma_512: torch.Tensor, latent_mean: torch.Tensor, key_len: int = 64, shift: int = 448,
                 sigma: float = 1.0, sd: float = 1.0, lr: float = 0.2, precision: int = capi.PREC_FP32,
                 loss: str = "mse", lpips_params: Optional[dict] = None):
        self.plan, self.device = plan, plan.device
        if loss not in ("mse", "lpips"):
            raise ValueError(f"loss must be 'mse' or 'lpips', got {loss!r}")
        self.loss_kind, self.lpips, self._lpips_target = loss, None, None
        if loss == "lpips":
            if lpips_params is None:
                raise ValueError("loss='lpips' needs lpips_params (PNetLin state_dict: VGG16 convs + linear heads)")
            from lfp_native.lpips import LpipsPlan
            self.lpips = LpipsPlan(plan.size, plan.size, device=plan.device)
            self.lpips.load(lpips_params)
        dev, f32 = self.device, torch.float32
        dim = pc.shape[0]
        self.dim, self.key_len, self.n_main = dim, key_len, dim - key_len
        pc = pc.to(dev, f32)
        s512 = sigma_512.to(dev, f32).reshape(-1)
        # get_uv / get_alpha_bound (src/main.py:23-40)
        self.V = pc[shift:shift + key_len].contiguous()
        self.U = torch.cat([pc[:shift], pc[shift + key_len:dim]], 0).contiguous()
        self.sigma_key = torch.full((key_len,), float(sigma), device=dev, dtype=f32)
        self.sigma_main = torch.cat([s512[:shift], s512[shift + key_len:dim]], 0).contiguous()
        self.max_alpha, self.min_alpha = 3 * self.sigma_main, -3 * self.sigma_main
        self.mu = latent_mean.to(dev, f32).reshape(-1).contiguous()
        self.sd, self.lr0, self.precision = float(sd), float(lr), precision
        self.noise = [n.to(dev, f32).contiguous() for n in noise]
        self._ws = None
        self._mse_scratch = None

    # ---- pieces -------------------------------------------------------------------------------
    def embed(self, alpha: torch.Tensor, key: torch.Tensor):
        """alpha [B, n_main], key logits [B, key_len] -> (w0, wx) [B, dim]."""
        B = alpha.shape[0]
        w0 = torch.empty(B, self.dim, device=self.device)
        wx = torch.empty_like(w0)
        capi.check(capi.lib().lfp_embed_forward(ptr(alpha), ptr(key), ptr(self.U), ptr(self.V), ptr(self.sigma_key),
                                                ptr(self.mu), self.sd, B, self.n_main, self.key_len, self.dim,
                                                ptr(w0), ptr(wx), stream_ptr(self.device)), "embed_forward")
        return w0, wx

    def embed_with_key(self, alpha: torch.Tensor, key_bits: torch.Tensor):
        """generate_with_alpha's latent (src/generator.py:83-89): binary key instead of sigmoid(logits)."""
        w0 = alpha @ self.U + self.mu
        wx = w0 + self.sd * ((self.sigma_key * key_bits.to(w0.dtype)) @ self.V)
        return w0, wx

    def render(self, wx: torch.Tensor) -> torch.Tensor:
        """generate_image (src/generator.py:170-174) for B latents, no gradient kept."""
        B = wx.shape[0]
        latent = wx[:, None, :].expand(B, self.plan.n_latent, self.dim).contiguous()
        return self.plan.generate(latent, self.noise, self.precision)

    def loss_and_grad(self, wx: torch.Tensor, target: torch.Tensor):
        """MSE(target, G(wx)) per trajectory and its gradient w.r.t. wx.  target [1 or B, 3, S, S]."""
        B = wx.shape[0]
        L = capi.lib()
        if self._ws is None or self._ws.numel() < self.plan.workspace_bytes(B) + 256:
            self._ws = None
            self._ws = self.plan.new_workspace(B)
        latent = wx[:, None, :].expand(B, self.plan.n_latent, self.dim).contiguous()
        img = self.plan.forward(latent, self.noise, self._ws, self.precision)
        if self.loss_kind == "lpips":
            key = (target.data_ptr(), target._version, tuple(target.shape), self.precision)
            if key != self._lpips_target:      # the target's VGG features are computed once per image (and arithmetic)
                self.lpips.set_target(target, self.precision)
                self._lpips_target = key
            loss, d_img = self.lpips.loss_grad(img, self.precision)
        else:
            numel = img[0].numel()
            nb = L.lfp_mse_scratch_bytes(B, numel)
            if self._mse_scratch is None or self._mse_scratch.numel() < nb:
                self._mse_scratch = torch.empty(max(nb, 4), dtype=torch.uint8, device=self.device)
            loss = torch.empty(B, device=self.device)
            d_img = torch.empty_like(img)
            capi.check(L.lfp_mse_loss_grad(ptr(img), ptr(target), target.shape[0], B, numel, ptr(loss), ptr(d_img),
                                           ptr(self._mse_scratch), self._mse_scratch.numel(), stream_ptr(self.device)),
                       "mse_loss_grad")
        d_latent = self.plan.backward(d_img, B, self._ws, self.precision)
        # backward of the repeat over the latent slots (src/model.py:531-535): ascending-slot sum, the order the native
        # step (slot_sum_kernel) uses, so that both drivers give the same bits
        d_wx = d_latent[:, 0].clone()
        for slot in range(1, d_latent.shape[1]):
            d_wx += d_latent[:, slot]
        return loss, d_wx, img

    def loss_and_grad_host(self, wx_host: torch.Tensor, target: torch.Tensor, loss_host: torch.Tensor,
                           dwx_host: torch.Tensor) -> None:
        """End-to-end call with HOST (pinned) buffers: H2D of the latents, synthesis forward +
        backward, D2H of the per-trajectory loss and d(loss)/d(wx); synchronous."""
        wx = wx_host.to(self.device, non_blocking=True)
        loss, dwx, _ = self.loss_and_grad(wx, target)
        loss_host.copy_(loss, non_blocking=True)
        dwx_host.copy_(dwx, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    # ---- the loop -----------------------------------------------------------------------------
    def init_state(self, alpha0: torch.Tensor, optimise_alpha: bool = True):
        """alpha0 [B, n_main]; key logits start at zero (src/utils.py:19-21)."""
        alpha = alpha0.to(self.device, torch.float32).clone().contiguous()
        key = torch.zeros(alpha.shape[0], self.key_len, device=self.device)
        z = torch.zeros_like
        return dict(alpha=alpha, key=key, m_a=z(alpha), v_a=z(alpha), m_k=z(key), v_k=z(key), step=0,
                    optimise_alpha=optimise_alpha, loss=None)

    def alpha0_from_lhs(self, u: torch.Tensor) -> torch.Tensor:
        """LHS sample in (0,1) -> ``2 u sigma - sigma`` (src/main.py:52)."""
        u = u.to(self.device, torch.float32)
        return 2 * u * self.sigma_main - self.sigma_main

    def step(self, st: dict, target: torch.Tensor) -> None:
        """One Adam step of every trajectory in ``st`` (no host synchronisation)."""
        alpha, key = st["alpha"], st["key"]
        B = alpha.shape[0]
        w0, wx = self.embed(alpha, key)
        mse, d_wx, _ = self.loss_and_grad(wx, target)
        L = capi.lib()
        loss = torch.empty_like(mse)
        capi.check(L.lfp_attrib_bound_loss(ptr(alpha), ptr(self.max_alpha), ptr(self.min_alpha), ptr(mse), B, self.n_main,
                                           0.1, ptr(loss), stream_ptr(self.device)), "attrib_bound_loss")   # src/main.py:65
        st["loss"] = loss
        st["w0"] = w0
        i = st["step"]
        lr = get_lr(i, self.lr0)                                            # src/main.py:67
        t, b1, b2 = i + 1, 0.9, 0.999
        # backward of the embed + bound sub-gradient + Adam (src/main.py:69-70) in one kernel, in place
        capi.check(L.lfp_attrib_adam_update(ptr(d_wx), ptr(alpha), ptr(key), ptr(self.U), ptr(self.V), ptr(self.sigma_key),
                                            ptr(self.max_alpha), ptr(self.min_alpha), self.sd, 0.1, ptr(st["m_a"]),
                                            ptr(st["v_a"]), ptr(st["m_k"]), ptr(st["v_k"]), B, self.n_main, self.key_len,
                                            self.dim, lr / (1 - b1 ** t), math.sqrt(1 - b2 ** t), b1, b2, 1 - b1, 1 - b2, 1e-8,
                                            1 if st["optimise_alpha"] else 0, stream_ptr(self.device)), "attrib_adam_update")
        st["step"] = i + 1

    STATE_KEYS = ("alpha", "key", "m_a", "v_a", "m_k", "v_k")

    def host_state(self, st: dict) -> dict:
        """Pinned host copy of a trajectory state (what a host-side caller of the step owns)."""
        hs = {k: st[k].detach().cpu().pin_memory() for k in self.STATE_KEYS}
        hs["loss"] = torch.zeros(st["alpha"].shape[0]).pin_memory()
        hs["step"], hs["optimise_alpha"] = st["step"], st["optimise_alpha"]
        return hs

    def step_host(self, hs: dict, target: torch.Tensor, st: Optional[dict] = None, stepper=None) -> dict:
        """One full Adam step with the trajectory state in HOST (pinned) buffers: H2D of alpha, key logits and the Adam
        moments, embed -> synthesis forward -> loss -> synthesis backward -> Adam on the device, D2H of the updated
        state and the per-trajectory loss; synchronous.  ``st`` (device buffers of the same shapes) is reused when
        given; ``stepper`` (a NativeStepper bound to ``st``) runs the step as one graph launch.  This is the
        end-to-end call bench.py times as ``e2e``."""
        if st is None:
            st = {k: torch.empty(hs[k].shape, device=self.device) for k in self.STATE_KEYS}
        for k in self.STATE_KEYS:
            st[k].copy_(hs[k], non_blocking=True)
        st["step"], st["optimise_alpha"] = hs["step"], hs["optimise_alpha"]
        if stepper is not None:      # ``st`` must be the state the stepper is bound to
            stepper.run(1)
        else:
            self.step(st, target)
        for k in self.STATE_KEYS:
            hs[k].copy_(st[k], non_blocking=True)
        hs["loss"].copy_(st["loss"], non_blocking=True)
        hs["step"] = st["step"]
        torch.cuda.current_stream(self.device).synchronize()
        return st

    def native_stepper(self, st: dict, target: torch.Tensor, max_steps: int = 4096) -> "NativeStepper":
        """Bind ``st`` and ``target`` to the native whole-step (lfp_attrib_*, CUDA-graph replay)."""
        return NativeStepper(self, st, target, max_steps)

    def run(self, alpha0: torch.Tensor, target: torch.Tensor, steps: int, optimise_alpha: bool = True,
            native: bool = True):
        """``steps`` Adam steps from ``alpha0`` (src/main.py:51-81 for every trajectory of the batch).  ``native``: the
        whole step is one captured CUDA graph replayed per step; otherwise every step is driven from Python through the
        same kernels (bit-identical results, tests/test_attribution_gpu.py)."""
        st = self.init_state(alpha0, optimise_alpha)
        if native and steps > 0:
            self.native_stepper(st, target, max_steps=max(steps, 1)).run(steps)
            return st
        for _ in range(steps):
            self.step(st, target)
        return st

    @staticmethod
    def decode(key_logits: torch.Tensor) -> torch.Tensor:
        """``round(sigmoid(key))`` (src/main.py:72)."""
        return torch.round(torch.sigmoid(key_logits))


class NativeStepper:
    """One trajectory batch bound to the native whole-step (include/lfp_sg2.h group 6).  ``run(n)`` enqueues ``n`` Adam
    steps: the first eagerly, the rest as replays of one captured CUDA graph, on a private stream that is ordered after
    and before the caller's current stream.  ``st`` is updated in place (``alpha``, ``key``, moments, ``loss``, ``step``)."""

    def __init__(self, eng: AttributionEngine, st: dict, target: torch.Tensor, max_steps: int = 4096):
        import ctypes as C
        self.eng, self.st, self.device = eng, st, eng.device
        L = capi.lib()
        B = st["alpha"].shape[0]
        self.B = B
        self.target = target.to(self.device, torch.float32).contiguous()
        self.loss = torch.zeros(B, device=self.device)
        self._h = C.c_void_p()
        self.max_steps = max_steps + st["step"]
        with torch.cuda.device(self.device):
            capi.check(L.lfp_attrib_create(C.byref(self._h), eng.plan._h, eng.plan.size, B, eng.n_main, eng.key_len, eng.dim,
                                           ptr(eng.U), ptr(eng.V), ptr(eng.sigma_key), ptr(eng.mu), ptr(eng.max_alpha),
                                           ptr(eng.min_alpha), eng.sd, 0.1, eng.lr0, self.max_steps, eng.precision), "attrib_create")
            nbytes = int(L.lfp_attrib_workspace_bytes(self._h))
            self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = (self.ws.data_ptr() + 255) // 256 * 256
            nz = eng.noise
            self._nptr = (C.c_void_p * len(nz))(*[n.data_ptr() for n in nz])
            self._nb = (C.c_int * len(nz))(*[n.shape[0] for n in nz])
            capi.check(L.lfp_attrib_bind(self._h, self._nptr, self._nb, ptr(self.target), self.target.shape[0], ptr(st["alpha"]),
                                         ptr(st["key"]), ptr(st["m_a"]), ptr(st["v_a"]), ptr(st["m_k"]), ptr(st["v_k"]),
                                         ptr(self.loss), 1 if st["optimise_alpha"] else 0, base,
                                         self.ws.data_ptr() + self.ws.numel() - base), "attrib_bind")
            if eng.loss_kind == "lpips":    # the perceptual loss inside the captured step
                eng.lpips.set_target(self.target, eng.precision)
                n = int(L.lfp_lpips_workspace_bytes(eng.lpips._h, B))
                self.lpips_ws = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
                lbase = (self.lpips_ws.data_ptr() + 255) // 256 * 256
                capi.check(L.lfp_attrib_set_lpips(self._h, eng.lpips._h, lbase, self.lpips_ws.data_ptr() + self.lpips_ws.numel() - lbase),
                           "attrib_set_lpips")
            self.stream = torch.cuda.Stream(self.device)
            self._dev_step = None

    def __del__(self):
        try:
            if self._h:
                torch.cuda.synchronize(self.device)
                capi.lib().lfp_attrib_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run(self, steps: int, graph: bool = True) -> None:
        L = capi.lib()
        st = self.st
        if st["step"] + steps > self.max_steps:
            raise RuntimeError(f"stepper was created for {self.max_steps} steps; asked to run to {st['step'] + steps}")
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.device(self.device):
            if self._dev_step != st["step"]:     # the host-side step index was changed behind the stepper (e.g. step_host)
                capi.check(L.lfp_attrib_set_step(self._h, st["step"], self.stream.cuda_stream), "attrib_set_step")
            capi.check(L.lfp_attrib_run(self._h, steps, 1 if graph else 0, self.stream.cuda_stream), "attrib_run")
        cur.wait_stream(self.stream)
        st["step"] += steps
        self._dev_step = st["step"]
        st["loss"] = self.loss
