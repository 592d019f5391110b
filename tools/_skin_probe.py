import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lumol_b200 import _ffi, synthetic
from lumol_b200.device import DeviceSystem
system = synthetic.lj_box((128, 128, 64), seed=20240 + 20)
synthetic.maxwell_boltzmann(system, 120.0, seed=7)
for skin in (1.0, 1.3, 1.6, 2.0):
    device = DeviceSystem(0)
    device.sync(system, velocities=True)
    lib, ctx = device.lib, device.ctx
    _ffi.check(ctx, lib.lumol_cuda_set_neighbor_skin(ctx, skin))
    _ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, _ffi.INTEGRATOR_VELOCITY_VERLET, 1.0))
    _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 300))
    r0 = device.stats().neighbor_rebuilds
    lib.lumol_cuda_synchronize(ctx)
    t = time.perf_counter()
    _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 2000))
    lib.lumol_cuda_synchronize(ctx)
    dt = time.perf_counter() - t
    print(f"skin {skin}: {dt / 2000 * 1e3:.4f} ms/step, rebuilds {device.stats().neighbor_rebuilds - r0} in 2000 steps", flush=True)
    device.close()
