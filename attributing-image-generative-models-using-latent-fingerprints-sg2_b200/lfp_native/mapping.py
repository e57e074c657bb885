"""Torch-facing wrapper of the native mapping network and PCA covariance (C-ABI group 8 of include/lfp_sg2.h):
``Generator.style`` (src/model.py:407-416) on many latents and the fp64 mean / covariance the PCA needs (src/PCA.py:62-108)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import capi
from .torch_glue import ptr, require_cuda, stream_ptr


class MappingPlan:
    def __init__(self, dim: int = 512, n_mlp: int = 8, lr_mul: float = 0.01, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("MappingPlan needs a CUDA device (no CPU fallback)")
        self.dim, self.n_mlp = dim, n_mlp
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_mapping_create(C.byref(self._h), dim, n_mlp, lr_mul), "mapping_create")

    def __del__(self):
        try:
            if self._h:
                capi.lib().lfp_mapping_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load(self, params: Dict[str, torch.Tensor]) -> None:
        """``params``: generator state_dict entries ``style.<i>.weight`` / ``style.<i>.bias``."""
        L = capi.lib()
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            keep = []
            for i in range(1, self.n_mlp + 1):
                for suffix in ("weight", "bias"):
                    name = f"style.{i}.{suffix}"
                    t = params[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
                    keep.append(t)
                    capi.check(L.lfp_mapping_set_param(self._h, name.encode(), ptr(t), t.numel(), st), "mapping_set_param")
            capi.check(L.lfp_mapping_finalize(self._h, st), "mapping_finalize")
            torch.cuda.current_stream(self.device).synchronize()

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        require_cuda(z, "z")
        z = z.to(torch.float32).contiguous()
        n = z.shape[0]
        out = torch.empty_like(z)
        scratch = torch.empty(2 * n * self.dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(capi.lib().lfp_mapping_forward(self._h, ptr(z), n, ptr(out), ptr(scratch), scratch.numel() * 4,
                                                      stream_ptr(self.device)), "mapping_forward")
        return out


def covariance(w: torch.Tensor):
    """fp64 (mean [dim], unbiased covariance [dim, dim]) of ``w [n, dim]`` with a fixed summation order."""
    require_cuda(w, "w")
    w = w.to(torch.float32).contiguous()
    n, dim = w.shape
    mean = torch.empty(dim, dtype=torch.float64, device=w.device)
    cov = torch.empty(dim, dim, dtype=torch.float64, device=w.device)
    with torch.cuda.device(w.device):
        capi.check(capi.lib().lfp_pca_covariance(ptr(w), n, dim, ptr(mean), ptr(cov), stream_ptr(w.device)), "pca_covariance")
    return mean, cov
