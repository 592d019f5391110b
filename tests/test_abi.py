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


def test_rust_binding_declares_every_symbol_and_constant():
    """rust/ffi.rs (the binding a lumol maintainer adds) cannot be compiled here: keep it symbol-for-symbol
    with the header, and its enum constants equal to the header's values."""
    with open(os.path.join(ROOT, "rust", "ffi.rs")) as fd:
        binding = fd.read()
    assert sorted(set(re.findall(r"\bpub fn (lumol_cuda_[a-z0-9_]+)\s*\(", binding))) == declared_symbols()
    with open(os.path.join(ROOT, "include", "lumol_cuda.h")) as fd:
        header = fd.read()
    values = {name: int(value) for name, value in re.findall(r"\b(LUMOL_CUDA_[A-Z0-9_]+)\s*=\s*(-?\d+)", header)}
    constants = re.findall(r"pub const (LUMOL_CUDA_[A-Z0-9_]+): [iu]32 = (-?\d+);", binding)
    assert len(constants) > 30
    for name, value in constants:
        assert values[name] == int(value), name
    # the stats mirror has one field per header field
    stats_header = re.search(r"typedef struct \{([^}]*)\} lumol_cuda_stats;", header).group(1)
    stats_rust = re.search(r"pub struct lumol_cuda_stats \{([^}]*)\}", binding).group(1)
    assert re.findall(r"\b([a-z_0-9]+)(?:\[3\])?;", stats_header) == re.findall(r"pub ([a-z_0-9]+):", stats_rust)


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
    # the multi-device constructor refuses the same way, and checks its arguments before touching a device
    devices = (ctypes.c_int32 * 2)(0, 1)
    assert lib.lumol_cuda_create_multi(devices, 2, ctypes.byref(ctx)) == _ffi.ERROR_NO_DEVICE
    assert "no CPU fallback" in _ffi.last_error(None)
    twice = (ctypes.c_int32 * 2)(0, 0)
    assert lib.lumol_cuda_create_multi(twice, 2, ctypes.byref(ctx)) == _ffi.ERROR_INVALID_ARGUMENT
    assert "listed twice" in _ffi.last_error(None)
    assert lib.lumol_cuda_create_multi(devices, 0, ctypes.byref(ctx)) == _ffi.ERROR_INVALID_ARGUMENT
    import lumol_b200 as lumol
    import systems

    with pytest.raises(lumol.LumolCudaError):
        systems.argon().forces()
