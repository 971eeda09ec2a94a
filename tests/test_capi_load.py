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
    assert (ma.value, mi.value) == (0, 2)
    # sizes of the structs as include/txasm.h lays them out on LP64 (the binding must follow the header)
    assert C.sizeof(capi.Term) == 48 and C.sizeof(capi.InArgs) == 72 and C.sizeof(capi.Info) == 120


def test_struct_sizes_match_the_header():
    """Compile a tiny C program against include/txasm.h and compare sizeof() with the ctypes mirrors."""
    import ctypes as C, subprocess, tempfile
    src = '#include <stdio.h>\n#include "txasm.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(txasm_term), sizeof(txasm_inargs), sizeof(txasm_info), sizeof(txasm_config), sizeof(txasm_timers));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(capi.Term), C.sizeof(capi.InArgs), C.sizeof(capi.Info), C.sizeof(capi.Config), C.sizeof(capi.Timers)]


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
