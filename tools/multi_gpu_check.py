"""Run under torchrun: compares the sharded evaluation (N ranks, NCCL) with a single-GPU evaluation of the same system.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import lumol_b200 as lumol
from lumol_b200 import _ffi, md, parallel, synthetic
from lumol_b200.device import DeviceSystem


def evaluate(system, rank, world, local_rank, sharded, kspace=None):
    device = DeviceSystem(local_rank)
    if sharded:
        parallel.init_communicator(device, rank, world)
    if kspace is not None:
        device.set_kspace_algorithm(kspace)
    device.sync(system, velocities=True)
    result = device.compute(forces=True, energy=True, virial=True)
    return device, result


def check(name, system, rank, world, local_rank, kspace=None):
    single_device, single = evaluate(system, rank, world, local_rank, sharded=False, kspace=kspace)
    sharded_device, sharded = evaluate(system, rank, world, local_rank, sharded=True, kspace=kspace)
    scale = max(np.abs(single.forces).max(), 1e-300)
    force_error = np.abs(sharded.forces - single.forces).max() / scale
    terms = ("pairs", "pairs_tail", "bonds", "angles", "dihedrals", "coulomb_real", "coulomb_self", "coulomb_kspace")
    magnitude = sum(abs(getattr(single.energy, t)) for t in terms)
    energy_error = max(abs(getattr(sharded.energy, t) - getattr(single.energy, t)) for t in terms) / magnitude
    virial_error = np.abs(sharded.virial - single.virial).max() / np.abs(single.virial).max()
    assert force_error < 1e-12 and energy_error < 1e-12 and virial_error < 1e-12, (name, force_error, energy_error, virial_error)

    # host-driven step of a sharded run: every rank uploads only its block of positions (the blocks travel between the
    # devices) and downloads only its block of forces
    import ctypes

    moved = np.ascontiguousarray(system.positions + 0.01 * np.sin(np.arange(system.size() * 3).reshape(-1, 3)))
    lib = sharded_device.lib
    first, count = ctypes.c_int64(0), ctypes.c_int64(0)
    _ffi.check(sharded_device.ctx, lib.lumol_cuda_owned_range(sharded_device.ctx, ctypes.byref(first), ctypes.byref(count)))
    lo, hi = first.value, first.value + count.value
    block = np.ascontiguousarray(moved[lo:hi])
    _ffi.check(sharded_device.ctx, lib.lumol_cuda_set_owned_positions(sharded_device.ctx, _ffi.as_double_pointer(block)))
    owned_forces = np.zeros((max(hi - lo, 1), 3))
    _ffi.check(sharded_device.ctx, lib.lumol_cuda_compute(sharded_device.ctx, _ffi.FORCES | _ffi.OWNED_FORCES, _ffi.PART_ALL,
                                                          _ffi.as_double_pointer(owned_forces), None, None))
    _ffi.check(single_device.ctx, lib.lumol_cuda_set_positions(single_device.ctx, _ffi.as_double_pointer(moved)))
    reference = single_device.compute(forces=True).forces
    owned_error = np.abs(owned_forces[:hi - lo] - reference[lo:hi]).max() / max(np.abs(reference).max(), 1e-300) if hi > lo else 0.0
    assert owned_error < 1e-12, (name, "owned block", owned_error)
    original = np.ascontiguousarray(system.positions, dtype=np.float64)
    for device in (single_device, sharded_device):
        _ffi.check(device.ctx, lib.lumol_cuda_set_positions(device.ctx, _ffi.as_double_pointer(original)))

    # ten device-resident MD steps: sharded trajectory equals the single-GPU one to summation-order noise
    def run(device):
        lib, ctx = device.lib, device.ctx
        _ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, _ffi.INTEGRATOR_VELOCITY_VERLET, 1.0))
        _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, 10))
        x = np.zeros((system.size(), 3))
        v = np.zeros((system.size(), 3))
        _ffi.check(ctx, lib.lumol_cuda_get_positions(ctx, _ffi.as_double_pointer(x)))
        _ffi.check(ctx, lib.lumol_cuda_get_velocities(ctx, _ffi.as_double_pointer(v)))
        return x, v

    x1, v1 = run(single_device)
    xs, vs = run(sharded_device)
    assert np.abs(xs - x1).max() < 1e-10, np.abs(xs - x1).max()
    assert np.abs(vs - v1).max() < 1e-10 * max(np.abs(v1).max(), 1e-3)
    if rank == 0:
        print(f"{name}: {world} ranks vs 1 rank: forces {force_error:.1e} energy {energy_error:.1e} virial {virial_error:.1e}, "
              f"path {sharded_device.stats().neighbor_path}, MD positions {np.abs(xs - x1).max():.1e}")
    single_device.close()
    sharded_device.close()


def check_long_md(name, system, steps, rank, world, local_rank):
    """Device-resident velocity-Verlet MD long enough to cross neighbour-list rebuilds: the sharded run (units of atoms
    owned by ranks, halo frames pushed over NVLink, device-side all-gather at rebuilds) against the single-GPU run."""
    results = []
    for sharded in (False, True):
        device = DeviceSystem(local_rank)
        if sharded:
            parallel.init_communicator(device, rank, world)
        device.sync(system, velocities=True)
        lib, ctx = device.lib, device.ctx
        _ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, _ffi.INTEGRATOR_VELOCITY_VERLET, 1.0))
        for chunk in (steps // 3, steps - steps // 3):  # two calls: the state leaves and re-enters cell order
            _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, chunk))
        n = system.size()
        x, v, f = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        _ffi.check(ctx, lib.lumol_cuda_get_positions(ctx, _ffi.as_double_pointer(x)))
        _ffi.check(ctx, lib.lumol_cuda_get_velocities(ctx, _ffi.as_double_pointer(v)))
        _ffi.check(ctx, lib.lumol_cuda_get_forces(ctx, _ffi.as_double_pointer(f)))
        rebuilds = device.stats().neighbor_rebuilds
        # the sharded state is usable by the estimators afterwards
        energy = device.compute(energy=True).energy.pairs
        results.append((x, v, f, rebuilds, energy))
        device.close()
    (x1, v1, f1, r1, e1), (xs, vs, fs, rs, es) = results
    assert rs >= 2, f"{name}: the run crossed no rebuild ({rs})"
    assert np.abs(xs - x1).max() < 1e-9, (name, np.abs(xs - x1).max())
    assert np.abs(vs - v1).max() < 1e-9 * np.abs(v1).max(), (name, np.abs(vs - v1).max())
    assert np.abs(fs - f1).max() < 1e-8 * np.abs(f1).max(), (name, np.abs(fs - f1).max())
    assert abs(es - e1) < 1e-10 * abs(e1), (name, es, e1)
    if rank == 0:
        print(f"{name}: {world} ranks vs 1 rank after {steps} MD steps ({rs} rebuilds): positions {np.abs(xs - x1).max():.1e} A, "
              f"forces {np.abs(fs - f1).max() / np.abs(f1).max():.1e}")


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import systems

    # 274 625 atoms (odd count), 26 x 26 x 26 cells: every rank owns whole units; hot enough to rebuild within 70 steps
    big = synthetic.lj_box(65, seed=11)
    synthetic.maxwell_boltzmann(big, 300.0, seed=3)
    check_long_md("lj-274625 (sorted-resident MD, halo exchange)", big, 70, rank, world, local_rank)
    # fewer units than would fill every rank evenly, odd atom count
    small_box = synthetic.lj_box((17, 17, 19), seed=12)
    synthetic.maxwell_boltzmann(small_box, 300.0, seed=4)
    check_long_md("lj-5491 (sorted-resident MD, few units)", small_box, 60, rank, world, local_rank)

    lj = synthetic.lj_box(24, seed=3)  # 13824 atoms, cell list
    synthetic.maxwell_boltzmann(lj, 120.0, seed=1)
    check("lj-13824 (neighbour list)", lj, rank, world, local_rank)

    small = synthetic.lj_box(8, seed=4)  # 512 atoms, all-pairs
    synthetic.maxwell_boltzmann(small, 120.0, seed=2)
    check("lj-512 (all-pairs)", small, rank, world, local_rank)

    water = synthetic.spce_box(10, flexible=True)  # 3000 atoms: LJ + Ewald + bonds/angles
    ewald = lumol.SharedEwald(lumol.Ewald(9.0, 6, 0.32))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    water.set_coulomb_potential(ewald)
    synthetic.maxwell_boltzmann(water, 300.0, seed=5)
    check("spce-3000 (list + Ewald + bonded)", water, rank, world, local_rank)
    # the same with the register-tiled reciprocal-space kernels, atoms of rho and of the k-space forces sharded
    tiled = synthetic.spce_box(10, flexible=True)
    ewald = lumol.SharedEwald(lumol.Ewald(9.0, 11, 0.34))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    tiled.set_coulomb_potential(ewald)
    synthetic.maxwell_boltzmann(tiled, 300.0, seed=5)
    check("spce-3000 (tiled k-space kernels)", tiled, rank, world, local_rank, kspace=1)

    # kmax above 26: the tiled force kernel keeps e_x, e_y in its global scratch and walks atom tiles
    large = synthetic.spce_box(10, flexible=True)
    ewald = lumol.SharedEwald(lumol.Ewald(9.0, 30, 0.45))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    large.set_coulomb_potential(ewald)
    synthetic.maxwell_boltzmann(large, 300.0, seed=5)
    check("spce-3000 (tiled k-space kernels, kmax 30)", large, rank, world, local_rank, kspace=1)

    odd = synthetic.spce_box(9, flexible=True)  # 2187 atoms: odd count, peer-push inbox alignment
    ewald = lumol.SharedEwald(lumol.Ewald(8.5, 6, 0.33))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    odd.set_coulomb_potential(ewald)
    synthetic.maxwell_boltzmann(odd, 300.0, seed=8)
    check("spce-2187 (odd atom count)", odd, rank, world, local_rank)

    nacl = systems.md_nacl("wolf")
    nacl.positions += np.random.Generator(np.random.PCG64(9)).uniform(-0.2, 0.2, nacl.positions.shape)  # perfect lattice: zero forces
    synthetic.maxwell_boltzmann(nacl, 300.0, seed=6)
    check("nacl-64 (all-pairs, Wolf)", nacl, rank, world, local_rank)
    dist.barrier()
    if rank == 0:
        print("multi-GPU check ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
