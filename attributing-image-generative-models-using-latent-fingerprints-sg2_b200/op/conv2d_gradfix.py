"""``op.conv2d_gradfix`` - API of src/op/conv2d_gradfix.py:22-75 kept for callers of the module.

On torch >= 1.9 the reference's custom op is disabled (``could_use_op``, :78-92) and both entry
points forward to ``F.conv2d`` / ``F.conv_transpose2d`` after a warning; that pass-through
behaviour is what is preserved here (without the warning).  The fused synthesis path does not
come through here: its convolutions are the library's own kernels (lfp_native.synthesis).
"""
import contextlib

from torch.nn import functional as F

enabled = True
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation,
                    groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return F.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                              output_padding=output_padding, dilation=dilation, groups=groups)
