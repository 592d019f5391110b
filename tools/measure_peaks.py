"""Measure the FP64 FMA peak and the device copy bandwidth through the C ABI (roofline denominators)."""
import ctypes, json, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumol_b200 import _ffi
lib = _ffi.library()
ctx = ctypes.c_void_p()
_ffi.check(None, lib.lumol_cuda_create(0, ctypes.byref(ctx)))
t, g = ctypes.c_double(), ctypes.c_double()
_ffi.check(ctx, lib.lumol_cuda_measure_fp64_peak(ctx, ctypes.byref(t)))
_ffi.check(ctx, lib.lumol_cuda_measure_copy_bandwidth(ctx, ctypes.byref(g)))
print(json.dumps({"fp64_tflops": t.value, "copy_gbs": g.value}))
