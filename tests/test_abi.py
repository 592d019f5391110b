"""The C-ABI library loads on a CPU-only box, exports every symbol of include/lumol_cuda.h, and refuses to
compute without a device (no fallback)."""

import ctypes
import os
import re

import pytest

from lumol_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "lumol_cuda.h")) as fd:
        header = fd.read()
    return sorted(set(re.findall(r"\b(lumol_cuda_[a-z0-9_]+)\s*\(", header)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_ffi.SIGNATURES)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(_ffi.LIBRARY_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.lumol_cuda_abi_version() == 1


def test_struct_layouts_match_the_header():
    # sizes implied by include/lumol_cuda.h (no implicit padding surprises)
    assert ctypes.sizeof(_ffi.Potential) == 48
    assert ctypes.sizeof(_ffi.Pair) == 16 + 8 * 10
    assert ctypes.sizeof(_ffi.Energy) == 64
    assert ctypes.sizeof(_ffi.Stats) == 8 * 21


def test_no_cpu_fallback_without_device():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a device is present")
    except ImportError:
        pass
    lib = _ffi.library()
    ctx = ctypes.c_void_p()
    status = lib.lumol_cuda_create(0, ctypes.byref(ctx))
    assert status == _ffi.ERROR_NO_DEVICE
    assert "no CPU fallback" in _ffi.last_error(None)
    import lumol_b200 as lumol
    import systems

    with pytest.raises(lumol.LumolCudaError):
        systems.argon().forces()
