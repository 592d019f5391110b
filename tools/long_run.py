"""Long NVE run of a synthetic LJ box: energy conservation and throughput at sizes beyond the bench default.

    python tools/long_run.py [--lattice 128x128x64] [--steps 20000] [--chunks 10]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from lumol_b200 import _ffi, md, synthetic  # noqa: E402
from lumol_b200.device import device_for  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--lattice", default="128x128x64")
    parser.add_argument("--steps", type=int, default=20000)
    parser.add_argument("--chunks", type=int, default=10)
    args = parser.parse_args()
    system = synthetic.lj_box(bench.lattice_of(args.lattice), seed=20240 + 20)
    synthetic.maxwell_boltzmann(system, 120.0, seed=7)
    n = system.size()
    device = device_for(system, velocities=True)
    lib, ctx = device.lib, device.ctx
    propagator = md.MolecularDynamics(1.0)
    propagator.setup(system)
    _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 200))  # the jittered lattice relaxes first

    def total_energy():
        result = device.compute(energy=True)
        kinetic = device.kinetic_energy()
        e = result.energy
        return e.pairs + e.pairs_tail + kinetic, kinetic

    e0, k0 = total_energy()
    out = {"atoms": n, "steps": args.steps, "energies": [e0]}
    start = time.perf_counter()
    for _ in range(args.chunks):
        _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, args.steps // args.chunks))
        out["energies"].append(total_energy()[0])
    seconds = time.perf_counter() - start
    e1, k1 = total_energy()
    out["relative_drift"] = abs(e1 - e0) / abs(e0)
    out["max_relative_excursion"] = max(abs(e - e0) for e in out["energies"]) / abs(e0)
    out["drift_per_kinetic"] = abs(e1 - e0) / k0
    out["atom_steps_per_s_wall_including_energy_queries"] = n * args.steps / seconds
    out["rebuilds"] = int(device.stats().neighbor_rebuilds)
    out["temperature_K_start_end"] = [2 * k0 / (3 * n * 8.31446284161522e-7), 2 * k1 / (3 * n * 8.31446284161522e-7)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
