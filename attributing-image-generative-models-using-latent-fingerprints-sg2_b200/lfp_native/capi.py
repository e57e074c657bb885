"""ctypes bindings for include/lfp_sg2.h (one prototype per exported symbol)."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LOCK = threading.Lock()

F32, F64, F16 = 0, 1, 2
PREC_FP32, PREC_TF32 = 0, 1
KINDS = ("conv_fwd", "conv_dgrad", "fir", "torgb", "act_bwd")


class LfpError(RuntimeError):
    """A liblfp_sg2 entry point returned non-zero."""


def lib_path() -> str:
    return os.path.join(_HERE, "liblfp_sg2.so")


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_PROTOS = {
    "lfp_last_error": (C.c_char_p, []),
    "lfp_version": (_i, []),
    "lfp_launch_count": (C.c_uint64, []),
    "lfp_upfirdn2d_out_size": (_i, [_i] * 12 + [C.POINTER(_i), C.POINTER(_i)]),
    "lfp_upfirdn2d": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _i64] + [_i] * 10 + [_vp]),
    "lfp_upfirdn2d_host": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _i64] + [_i] * 10),
    "lfp_fused_bias_act": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i, _i, _f, _f, _vp]),
    "lfp_fused_bias_act_host": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i, _i, _f, _f]),
    "lfp_bias_grad_reduce": (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "lfp_bias_grad_reduce_scratch": (_sz, [_i64, _i64, _i64]),
    "lfp_synth_create": (_i, [C.POINTER(_vp), _i, _i, _i, _vp, _i]),
    "lfp_synth_destroy": (None, [_vp]),
    "lfp_synth_n_latent": (_i, [_vp]),
    "lfp_synth_num_noise": (_i, [_vp]),
    "lfp_synth_set_param": (_i, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "lfp_synth_finalize": (_i, [_vp, _vp]),
    "lfp_synth_workspace_bytes": (_sz, [_vp, _i]),
    "lfp_synth_forward": (_i, [_vp, _i, _vp, C.POINTER(_vp), C.POINTER(_i), _vp, _vp, _sz, _i, _vp]),
    "lfp_synth_generate_workspace_bytes": (_sz, [_vp, _i]),
    "lfp_synth_generate": (_i, [_vp, _i, _vp, C.POINTER(_vp), C.POINTER(_i), _vp, _vp, _sz, _i, _vp]),
    "lfp_synth_backward": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "lfp_synth_num_convs": (_i, [_vp]),
    "lfp_synth_read_activation": (_i, [_vp, _i, _i, _vp, _vp, C.POINTER(_i), C.POINTER(_i), _vp]),
    "lfp_synth_forward_backward_host": (_i, [_vp, _i, _vp, C.POINTER(_vp), C.POINTER(_i), _vp, _vp, _vp, _i]),
    "lfp_synth_profile_begin": (_i, [_vp, _i]),
    "lfp_synth_profile_end": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]),
    "lfp_synth_profile_launches": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(C.c_float), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]),
    "lfp_attrib_create": (_i, [C.POINTER(_vp), _vp, _i, _i, _i, _i, _i] + [_vp] * 6 + [_f, _f, C.c_double, _i, _i]),
    "lfp_attrib_destroy": (None, [_vp]),
    "lfp_attrib_workspace_bytes": (_sz, [_vp]),
    "lfp_attrib_bind": (_i, [_vp, C.POINTER(_vp), C.POINTER(_i), _vp, _i] + [_vp] * 7 + [_i, _vp, _sz]),
    "lfp_attrib_set_lpips": (_i, [_vp, _vp, _vp, _sz]),
    "lfp_attrib_set_step": (_i, [_vp, _i, _vp]),
    "lfp_attrib_get_w0": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "lfp_attrib_run": (_i, [_vp, _i, _i, _vp]),
    "lfp_modconv_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, _i]),
    "lfp_modconv_destroy": (None, [_vp]),
    "lfp_modconv_set_param": (_i, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "lfp_modconv_finalize": (_i, [_vp, _vp]),
    "lfp_modconv_out_size": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "lfp_modconv_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "lfp_modconv_forward": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "lfp_modconv_backward": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "lfp_lpips_create": (_i, [C.POINTER(_vp), _i, _i]),
    "lfp_lpips_destroy": (None, [_vp]),
    "lfp_lpips_set_param": (_i, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "lfp_lpips_finalize": (_i, [_vp, _vp]),
    "lfp_lpips_workspace_bytes": (_sz, [_vp, _i]),
    "lfp_lpips_set_target": (_i, [_vp, _i, _vp, _vp, _sz, _i, _vp]),
    "lfp_lpips_loss_grad": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "lfp_mapping_create": (_i, [C.POINTER(_vp), _i, _i, _f]),
    "lfp_mapping_destroy": (None, [_vp]),
    "lfp_mapping_set_param": (_i, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "lfp_mapping_finalize": (_i, [_vp, _vp]),
    "lfp_mapping_forward": (_i, [_vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "lfp_pca_covariance": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "lfp_embed_forward": (_i, [_vp] * 6 + [_f, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lfp_embed_backward": (_i, [_vp] * 5 + [_f, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lfp_attrib_bound_loss": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "lfp_attrib_adam_update": (_i, [_vp] * 8 + [_f, _f] + [_vp] * 4 + [_i, _i, _i, _i] + [_f] * 7 + [_i, _vp]),
    "lfp_mse_loss_grad": (_i, [_vp, _vp, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "lfp_mse_scratch_bytes": (_sz, [_i, _i64]),
}


def lib() -> C.CDLL:
    """Load liblfp_sg2.so once; raise loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is None:
            path = lib_path()
            if not os.path.exists(path):
                raise LfpError(
                    f"{path} not found: build it with csrc/build.sh (or __graft_entry__.build()). "
                    "There is no CPU fallback for this path.")
            handle = C.CDLL(path)
            for name, (res, args) in _PROTOS.items():
                fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            _LIB = handle
    return _LIB


def exported_symbols():
    return sorted(_PROTOS)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().lfp_last_error().decode("utf-8", "replace")
        raise LfpError(f"{what or 'liblfp_sg2'} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().lfp_launch_count())
