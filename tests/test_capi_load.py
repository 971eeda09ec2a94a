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
