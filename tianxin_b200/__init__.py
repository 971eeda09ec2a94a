"""tianxin_b200 -- B200-native finite-element assembly path behind Tianxin's Panzer API.

The product is `libtxasm.so` (hand-written sm_100a CUDA kernels behind the C ABI declared in
include/txasm.h).  This package is the thin host-side mirror of the reference interface used by
the tests and the bench: ctypes bindings (`capi`) and the AssemblyEngine-shaped driver
(`assembly_engine`).  There is no CPU fallback: importing works without a GPU (so the symbol
table can be checked), every compute call fails loudly without one.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
