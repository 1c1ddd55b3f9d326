"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/gmql_cuda.h declares.
No compute calls are made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gmql_b200", "libgmqlcuda.so")
HDR = os.path.join(ROOT, "include", "gmql_cuda.h")


def _header_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmql_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_header_and_binding_agree():
    from gmql_b200 import _lib
    assert sorted(_lib.EXPORTS) == _header_functions()


def test_library_exports_every_declared_symbol(lib):
    for name in _header_functions():
        assert hasattr(lib, name), name


def test_abi_version(lib):
    lib.gmql_abi_version.restype = ctypes.c_int
    assert lib.gmql_abi_version() == 1


def test_no_device_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    rc = lib.gmql_ctx_create(0, None, ctypes.byref(ctx))
    assert rc < 0 and not ctx.value
    from gmql_b200 import GMQLCudaExecutor, GmqlCudaError
    with pytest.raises(GmqlCudaError):
        GMQLCudaExecutor()
