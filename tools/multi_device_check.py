"""One process, several devices: a context made by ``lumol_cuda_create_multi`` against a single-device context.

    python tools/multi_device_check.py 2        # devices 0, 1
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lumol_b200 as lumol
from lumol_b200 import _ffi, synthetic
from lumol_b200.device import DeviceSystem

TERMS = ("pairs", "pairs_tail", "bonds", "angles", "dihedrals", "coulomb_real", "coulomb_self", "coulomb_kspace")


def state(device, n):
    lib, ctx = device.lib, device.ctx
    x, v, f = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    _ffi.check(ctx, lib.lumol_cuda_get_positions(ctx, _ffi.as_double_pointer(x)))
    _ffi.check(ctx, lib.lumol_cuda_get_velocities(ctx, _ffi.as_double_pointer(v)))
    _ffi.check(ctx, lib.lumol_cuda_get_forces(ctx, _ffi.as_double_pointer(f)))
    return x, v, f


def check(name, system, devices, steps):
    results = []
    for target in (devices[0], devices):
        device = DeviceSystem(target)
        device.sync(system, velocities=True)
        evaluation = device.compute(forces=True, energy=True, virial=True)
        kinetic = device.kinetic_energy()
        lib, ctx = device.lib, device.ctx
        _ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, _ffi.INTEGRATOR_VELOCITY_VERLET, 1.0))
        for chunk in (steps // 3, steps - steps // 3):
            _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, chunk))
        x, v, f = state(device, system.size())
        after = device.compute(energy=True).energy
        results.append((evaluation, kinetic, x, v, f, after, device.stats().neighbor_rebuilds))
        device.close()
    (e1, k1, x1, v1, f1, a1, r1), (es, ks, xs, vs, fs, as_, rs) = results
    scale = max(np.abs(e1.forces).max(), 1e-300)
    force_error = np.abs(es.forces - e1.forces).max() / scale
    magnitude = sum(abs(getattr(e1.energy, t)) for t in TERMS)
    energy_error = max(abs(getattr(es.energy, t) - getattr(e1.energy, t)) for t in TERMS) / magnitude
    virial_error = np.abs(es.virial - e1.virial).max() / np.abs(e1.virial).max()
    assert force_error < 1e-12 and energy_error < 1e-12 and virial_error < 1e-12, (name, force_error, energy_error, virial_error)
    assert abs(ks - k1) <= 1e-12 * abs(k1), (name, ks, k1)
    assert np.abs(xs - x1).max() < 1e-9, (name, np.abs(xs - x1).max())
    assert np.abs(vs - v1).max() < 1e-9 * max(np.abs(v1).max(), 1e-3), name
    assert np.abs(fs - f1).max() < 1e-8 * np.abs(f1).max(), name
    after_magnitude = sum(abs(getattr(a1, t)) for t in TERMS)
    assert max(abs(getattr(as_, t) - getattr(a1, t)) for t in TERMS) < 1e-10 * after_magnitude, name
    print(f"{name}: {len(devices)} devices of one process vs 1: forces {force_error:.1e} energy {energy_error:.1e} virial {virial_error:.1e}; "
          f"after {steps} MD steps ({rs} rebuilds) positions {np.abs(xs - x1).max():.1e} A", flush=True)


def main():
    ndevices = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    devices = list(range(ndevices))
    import systems

    big = synthetic.lj_box(65, seed=11)  # sorted-resident engine, halo exchange between the devices
    synthetic.maxwell_boltzmann(big, 300.0, seed=3)
    check("lj-274625", big, devices, 70)

    lj = synthetic.lj_box(24, seed=3)
    synthetic.maxwell_boltzmann(lj, 120.0, seed=1)
    check("lj-13824", lj, devices, 12)

    water = synthetic.spce_box(10, flexible=True)
    ewald = lumol.SharedEwald(lumol.Ewald(9.0, 6, 0.32))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    water.set_coulomb_potential(ewald)
    synthetic.maxwell_boltzmann(water, 300.0, seed=5)
    check("spce-3000 (Ewald, bonded)", water, devices, 12)

    nacl = systems.md_nacl("wolf")
    nacl.positions += np.random.Generator(np.random.PCG64(9)).uniform(-0.2, 0.2, nacl.positions.shape)
    synthetic.maxwell_boltzmann(nacl, 300.0, seed=6)
    check("nacl-64 (all-pairs, Wolf)", nacl, devices, 12)

    # errors travel from the device threads to the caller with the reference's text
    device = DeviceSystem(devices)
    device.sync(lj, velocities=True)
    try:
        _ffi.check(device.ctx, device.lib.lumol_cuda_comm_init(device.ctx, 2, 0, (_ffi._c.c_uint8 * 128)()))
        raise AssertionError("comm_init on a multi-device context must fail")
    except lumol.LumolCudaError as error:
        assert "already shards" in str(error), str(error)
    device.close()
    print("multi-device check ok")


if __name__ == "__main__":
    main()
