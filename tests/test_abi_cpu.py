"""The C-ABI library loads without a GPU and exports every symbol include/salun.h declares."""
import ctypes
import os
import re

from unlearn_saliency_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(salun_[a-z0-9_]+)\s*\(", src))
    return names


import pytest


@pytest.mark.parametrize("precision", ["bf16", "split"])
def test_library_exports_every_declared_symbol(precision):
    """both builds (libsalun.so, libsalun_split.so: csrc/Makefile) export the whole C ABI"""
    lib = _lib.lib(precision)
    decl = declared_symbols()
    assert len(decl) >= 15
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by {_lib.LIB_PATHS[precision]}"
    assert lib.salun_version() >= 1000


def test_the_two_builds_keep_their_own_kernels():
    """RTLD_LOCAL + -Bsymbolic: the same symbol resolves to a different address in each library"""
    a, b = _lib.lib("bf16"), _lib.lib("split")
    pa = ctypes.cast(a.salun_resnet_forward_backward, ctypes.c_void_p).value
    pb = ctypes.cast(b.salun_resnet_forward_backward, ctypes.c_void_p).value
    assert pa != pb


def test_bound_symbols_are_declared():
    decl = declared_symbols()
    for name in _lib.exported_symbols():
        assert name in decl, f"{name} bound in _lib.py but missing from include/*.h"


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = _lib.lib().salun_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and _lib.lib().salun_last_error()
    import pytest
    from unlearn_saliency_b200.tail import SalunContext
    with pytest.raises(RuntimeError):
        SalunContext(0)
