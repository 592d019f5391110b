"""Monte Carlo cost calls on a bench system and on the 98k-atom SPC/E box, meant to be wrapped in ncu
(per-launch times of move_pairs_kernel / move_kspace_kernel / move_finish_kernel / move_accept_kernel), and timed
on the host without a profiler.

    python tools/profile_mc.py [--lattice 32] [--batch 64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import lumol_b200 as lumol  # noqa: E402


def time_costs(system, batch, repeats, rng):
    cache = lumol.EnergyCache()
    start = time.perf_counter()
    cache.init(system)
    init_s = time.perf_counter() - start
    nmol = len(system.molecules())

    def trial():
        molecule = int(rng.integers(0, nmol))
        bonding = system.molecule(molecule)
        return molecule, system.positions[bonding.start:bonding.end] + rng.uniform(-0.5, 0.5, 3)

    trials = [trial() for _ in range(batch)]
    cache.move_molecule_cost(system, *trials[0])
    start = time.perf_counter()
    for molecule, positions in trials[:repeats]:
        cache.move_molecule_cost(system, molecule, positions)
    single_us = (time.perf_counter() - start) / repeats * 1e6
    ids, news = [t[0] for t in trials], [t[1] for t in trials]
    cache.move_molecules_cost(system, ids, news)
    start = time.perf_counter()
    costs = cache.move_molecules_cost(system, ids, news)
    batch_us = (time.perf_counter() - start) / batch * 1e6
    # accept one and check the running energy against a full evaluation of the resident state
    cache.accept(1)
    bonding = system.molecule(ids[1])
    system.positions[bonding.start:bonding.end] = news[1]
    start = time.perf_counter()
    cache.update(system)
    accept_us = (time.perf_counter() - start) * 1e6
    full = system.potential_energy()
    return {
        "atoms": system.size(), "molecules": nmol, "full_energy_evaluation_ms": init_s * 1e3, "single_cost_us": single_us,
        "batch": batch, "batched_cost_us_per_trial": batch_us, "accept_us": accept_us,
        "cache_energy_minus_full_energy": cache.energy() - full, "energy": full, "cost_of_accepted": costs[1],
    }


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--lattice", default="32")
    parser.add_argument("--batch", type=int, default=64)
    args = parser.parse_args()
    rng = np.random.Generator(np.random.PCG64(3))
    import systems

    out = {"water_bench_ewald": time_costs(systems.water("ewald"), args.batch, 20, rng)}
    args.workload, args.steps, args.warmup = "spce", 1, 0
    system, description = bench.build_workload(args)
    out["spce_box"] = time_costs(system, args.batch, 10, rng)
    out["spce_box"]["workload"] = description["workload"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
