"""fastore_b200 -- B200-native fastore_bin categorise + scatter path.

The product is the C ABI in include/fastore_b200.h (libfastore_b200.so: hand-written sm_100a CUDA
kernels) plus the host C++ around it (libfastore_host.so, fastore_bin_b200).  This Python package
only builds and loads them; it contains no implementation of the path.
"""
from . import _native  # noqa: F401
from .binner import GpuBinner, BinBlock  # noqa: F401

__all__ = ["GpuBinner", "BinBlock"]
