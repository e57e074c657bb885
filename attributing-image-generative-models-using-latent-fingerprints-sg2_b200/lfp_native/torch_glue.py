"""Small helpers to hand torch CUDA tensors to the C ABI (pointers, stream, dtype codes)."""
from __future__ import annotations

import torch

from . import capi

_DTYPE = {torch.float32: capi.F32, torch.float64: capi.F64, torch.float16: capi.F16}


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {t.dtype}: the native ops take float32, float64, float16")


def require_cuda(t: torch.Tensor, name: str) -> None:
    # same wording as the reference shim's CHECK_CUDA (src/op/upfirdn2d.cpp:9-10)
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
