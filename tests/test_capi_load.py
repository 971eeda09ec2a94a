"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol that
include/txasm.h declares, and refuses to compute without a device (no CPU fallback)."""
import os
import re

import pytest

from tianxin_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "txasm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(txasm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in txasm.h but not exported by libtxasm.so"
    assert sorted(capi.EXPORTS) == declared


def test_version_and_struct_sizes():
    import ctypes as C
    ma, mi = C.c_int(), C.c_int()
    assert capi.lib().txasm_version(C.byref(ma), C.byref(mi)) == 0
    assert (ma.value, mi.value) == (0, 1)
    assert C.sizeof(capi.Term) == 32 and C.sizeof(capi.InArgs) == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.TxasmError) as ei:
        capi.Handle()
    assert ei.value.code == capi.ECUDA and "no CPU fallback" in str(ei.value)


def test_no_synchronous_copies_in_the_library():
    """The handle may run on a caller's non-blocking stream: a plain cudaMemcpy / cudaMemset is not ordered with it
    (and an H2D copy from pageable memory may return before the data has landed).  Every copy in the library goes
    through copy_to_device_sync() or cudaMemcpyAsync on the handle's stream."""
    import glob, os, re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tianxin_b200", "csrc")
    for path in glob.glob(os.path.join(root, "*.cu")):
        src = open(path).read()
        assert not re.search(r"\bcudaMemcpy\(", src), path
        assert not re.search(r"\bcudaMemset\(", src), path
