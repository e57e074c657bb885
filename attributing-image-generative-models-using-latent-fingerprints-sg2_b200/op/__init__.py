"""Drop-in replacement for the reference's ``op`` package (src/op/__init__.py:1-2): same names,
argument meaning and error behaviour, backed by liblfp_sg2.so instead of JIT-built extensions."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d
from . import conv2d_gradfix  # noqa: F401  (model code does `from op import conv2d_gradfix`)
