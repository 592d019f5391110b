"""A short device-resident MD run of a bench workload, meant to be wrapped in ncu (see profiles/README.md).

    python tools/profile_step.py [--workload lj|spce] [--lattice 128x128x64] [--steps 4] [--skin 1.0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from lumol_b200 import _ffi, md  # noqa: E402
from lumol_b200.device import device_for  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--workload", default="lj")
    parser.add_argument("--lattice", default="128x128x64")
    parser.add_argument("--steps", type=int, default=4)
    parser.add_argument("--skin", type=float, default=None)
    args = parser.parse_args()
    system, description = bench.build_workload(args)
    device = device_for(system, velocities=True)
    if args.skin is not None:
        _ffi.check(device.ctx, device.lib.lumol_cuda_set_neighbor_skin(device.ctx, args.skin))
    propagator = md.MolecularDynamics(bench.TIMESTEP_FS)
    propagator.setup(system)
    _ffi.check(device.ctx, device.lib.lumol_cuda_md_run(device.ctx, args.steps))
    device.lib.lumol_cuda_synchronize(device.ctx)
    print(description["workload"], "steps", args.steps, "launches", device.stats().kernel_launches)


if __name__ == "__main__":
    main()
