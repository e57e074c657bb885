"""Python host side of liblfp_sg2.so (ctypes; the library itself has no torch dependency).

The library is built in-tree by ``csrc/build.sh`` (or ``__graft_entry__.build()``) for sm_100a.
There is no fallback: if the shared object is missing or a CUDA call fails, an exception is
raised - the product path never routes through a CPU implementation.
"""
from .capi import LfpError, lib, lib_path, launch_count, check  # noqa: F401
