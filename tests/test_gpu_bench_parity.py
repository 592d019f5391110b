"""The bench configurations themselves against the CPU oracle (``-m gpu``), through the C ABI.

The boxes ``bench.py`` times are too large for the oracle's full O(N^2) loops of sys/compute.rs:37-55 (5.5e11 pair
visits at 1 048 576 atoms), so they are pinned on sampled atoms: ``orc_pair_forces_rows`` / ``orc_ewald_real_forces_rows``
evaluate the TOTAL force on a chosen atom over every other atom with the reference's roles and arithmetic.  The
98 304-atom SPC/E box (32 630 k-vectors) is small enough for the whole oracle.

Tolerances are north_star's: forces <= 1e-10, energies <= 1e-9.  Two force norms are asserted and recorded
(``gpurun_out/parity_errors.jsonl``): the error relative to the largest force component of the system, and the worst
PER-ATOM relative error |dF_i| / |F_i| over the compared atoms.
"""

import json
import os

import numpy as np
import pytest

import lumol_b200 as lumol
from lumol_b200 import _ffi, md, synthetic, units
from lumol_b200.device import device_for
from oracle import oracle
import systems

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORCE_TOL = 1e-10
ENERGY_TOL = 1e-9


def record(name, **values):
    """Append the measured errors of one comparison to gpurun_out/parity_errors.jsonl (read back after a GPU call)."""
    directory = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(directory, exist_ok=True)
        with open(os.path.join(directory, "parity_errors.jsonl"), "a") as fd:
            fd.write(json.dumps({"test": name, **{k: float(v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass


def force_errors(actual, expected, scale=None):
    """(error relative to the largest force component, worst per-atom relative error)."""
    scale = max(np.abs(expected).max() if scale is None else scale, 1e-300)
    norm = np.abs(actual - expected).max() / scale
    magnitude = np.linalg.norm(expected, axis=1)
    per_atom = (np.linalg.norm(actual - expected, axis=1) / np.maximum(magnitude, 1e-300)).max()
    return norm, per_atom


def assert_forces(name, actual, expected, scale=None, tol=FORCE_TOL):
    norm, per_atom = force_errors(actual, expected, scale)
    record(name, force_error_vs_max=norm, force_error_per_atom=per_atom, atoms=len(expected))
    print(f"{name}: force error {norm:.2e} of the largest component, worst per-atom relative error {per_atom:.2e}")
    assert norm <= tol, f"{name}: force error {norm:.3e} > {tol:.1e}"
    # per atom, as north_star words it; atoms whose net force is a near-complete cancellation of pair forces 1e3 times
    # larger cannot meet a bound relative to their own net force, hence the looser factor
    assert per_atom <= 1e3 * tol, f"{name}: per-atom relative force error {per_atom:.3e}"


def sampled_rows(n, count, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.unique(np.concatenate([np.linspace(0, n - 1, count // 2).astype(np.int64), rng.integers(0, n, count // 2)]))
    return rows


# ---- Lennard-Jones bench box ---------------------------------------------------------------------------------

@pytest.mark.parametrize("lattice", [(128, 128, 64), (256, 256, 128)])
def test_lj_bench_box_sampled_atoms_vs_oracle(lattice):
    """bench.py's workloads: the 8 388 608-atom argon box of the headline line and the 1 048 576-atom box beside it.
    512 (256 for the large box) sampled atoms, total force over all other atoms, forces-only and forces + energy + virial
    kernel instances; then thirty device-resident velocity-Verlet steps (sorted-resident engine, the neighbour list is
    reused with displaced atoms) and the same comparison at the final positions."""
    system = synthetic.lj_box(lattice, seed=20240 + 20)
    synthetic.maxwell_boltzmann(system, 120.0, seed=7)
    n = system.size()
    rows = sampled_rows(n, 512 if n < 2_000_000 else 256, seed=1)
    assert len(rows) >= 200
    label = f"lj-{n // 1048576}M"
    device = device_for(system, velocities=True)
    reference = oracle.OracleSystem(system)
    expected = reference.pair_forces_rows(rows)
    forces_only = device.compute(forces=True).forces
    assert device.stats().neighbor_path == 1
    scale = np.abs(forces_only).max()
    assert_forces(f"{label} forces-only kernel", forces_only[rows], expected, scale)
    full = device.compute(forces=True, energy=True, virial=True)
    assert_forces(f"{label} forces+energy+virial kernel", full.forces[rows], expected, scale)
    # Newton's third law over the whole box, and the virial against the force field it was computed with
    assert np.abs(full.forces.sum(axis=0)).max() < 1e-9 * scale * np.sqrt(n)

    # device-resident MD: lists reused between rebuilds
    propagator = md.MolecularDynamics(1.0)
    propagator.setup(system)
    lib, ctx = device.lib, device.ctx
    for steps in (7, 23):
        _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, steps))
        positions = np.zeros((n, 3))
        forces = np.zeros((n, 3))
        _ffi.check(ctx, lib.lumol_cuda_get_positions(ctx, _ffi.as_double_pointer(positions)))
        _ffi.check(ctx, lib.lumol_cuda_get_forces(ctx, _ffi.as_double_pointer(forces)))
        moved = synthetic.lj_box(lattice, seed=20240 + 20)
        moved.positions[:] = positions
        expected = oracle.OracleSystem(moved).pair_forces_rows(rows)
        assert_forces(f"{label} after {steps} more MD steps", forces[rows], expected, np.abs(forces).max())
    device.close()


# ---- SPC/E + Ewald bench box -----------------------------------------------------------------------------------

def spce_bench_system():
    """bench.py --workload spce: 32^3 flexible waters, Ewald::with_accuracy(9 A, 1e-5) -> kmax 25, 32 630 k-vectors."""
    system = synthetic.spce_box(32, flexible=True)
    ewald = lumol.Ewald.with_accuracy(9.0, 1e-5, system)
    shared = lumol.SharedEwald(ewald)
    shared.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(shared)
    return system, ewald


def test_spce_bench_box_vs_full_oracle():
    """The 98 304-atom SPC/E bench box against the WHOLE oracle: forces (pairs + Ewald real + k-space + bonded),
    every energy term, atomic virial.  About 1e10 pair visits and 6e9 atom-k products on the host cores."""
    system, ewald = spce_bench_system()
    assert ewald.kmax == 25
    # displaced from the lattice so that no net force is a pure cancellation
    rng = np.random.Generator(np.random.PCG64(5))
    system.positions += rng.uniform(-0.05, 0.05, system.positions.shape)
    device = device_for(system)
    result = device.compute(forces=True, energy=True, virial=True)
    stats = device.stats()
    assert stats.neighbor_path == 1 and stats.nkvectors == 32630
    forces_only = device.compute(forces=True).forces
    reference = oracle.OracleSystem(system)
    reference.lib.orc_set_threads(os.cpu_count() or 1)
    expected = reference.forces()
    assert_forces("spce-98k forces+energy+virial kernels", result.forces, expected)
    assert_forces("spce-98k forces-only kernels", forces_only, expected)
    terms = reference.energy_terms()
    names = ("pairs", "pairs_tail", "bonds", "angles", "dihedrals", "coulomb_real", "coulomb_self", "coulomb_kspace")
    magnitude = sum(abs(getattr(terms, name)) for name in names)
    worst = 0.0
    for name in names:
        error = abs(getattr(result.energy, name) - getattr(terms, name)) / magnitude
        worst = max(worst, error)
        assert error <= ENERGY_TOL, f"{name}: {getattr(result.energy, name)!r} vs {getattr(terms, name)!r}"
    virial = reference.atomic_virial()
    virial_error = np.abs(result.virial - virial).max() / np.abs(virial).max()
    record("spce-98k energy and virial", energy_error=worst, virial_error=virial_error)
    assert virial_error <= 1e-10
    device.close()


# ---- configs[1]: NaCl crystal with Ewald + Born-Mayer-Huggins pairs ---------------------------------------------

def bmh_pairs(system, cutoff):
    """Tosi-Fumi Born-Mayer-Huggins parameters for NaCl (additional to the reference's own inputs)."""
    kj = units.from_(1.0, "kJ/mol")
    rho = 0.317
    table = {
        ("Na", "Na"): (25.4435 * 1.25 * kj * 60.2214, 2.340, 101.17 * kj, 48.18 * kj),
        ("Na", "Cl"): (20.3548 * 1.00 * kj * 60.2214, 2.755, 674.48 * kj, 837.08 * kj),
        ("Cl", "Cl"): (15.2661 * 0.75 * kj * 60.2214, 3.170, 6985.70 * kj, 14031.0 * kj),
    }
    for pair, (a, sigma, c, d) in table.items():
        interaction = lumol.PairInteraction(lumol.BornMayerHuggins(a=a, c=c, d=d, sigma=sigma, rho=rho), cutoff)
        interaction.enable_tail_corrections()
        system.set_pair_potential(pair, interaction)


def test_nacl_ewald_born_mayer_huggins():
    """BASELINE.json configs[1] as worded: NaCl with Ewald electrostatics AND Born-Mayer-Huggins pairs.  The bench
    crystal (benches/data/nacl.pdb, 128 ions, all-pairs path) and a 13 824-ion rock-salt supercell on the cell-list
    path (general list kernel: exp + erfc in one pair)."""
    from test_gpu_parity import check_system

    system = systems.nacl("ewald")
    bmh_pairs(system, 9.0)
    check_system(system)

    a = 5.6402 / 2.0  # tests/data/md-nacl/small.xyz spacing
    side = 24
    grid = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), axis=-1).reshape(-1, 3)
    rng = np.random.Generator(np.random.PCG64(3))
    positions = (grid + 0.25) * a + rng.uniform(-0.15, 0.15, (len(grid), 3))
    names = ["Na" if (g.sum() % 2 == 0) else "Cl" for g in grid]
    charges = np.array([1.0 if name == "Na" else -1.0 for name in names])
    crystal = lumol.System(lumol.UnitCell.cubic(side * a))
    crystal.add_particles(names, positions, charges=charges)
    bmh_pairs(crystal, 9.0)
    crystal.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(9.0, 8)))
    device = check_system(crystal, molecular=False, path=1)
    assert device.stats().neighbor_path == 1


# ---- pairs sitting exactly on the cut-off, list path ------------------------------------------------------------

@pytest.mark.parametrize("charged", [False, True])
def test_pair_exactly_on_the_cutoff_list_path(charged):
    """pairs.rs:186 drops a pair at ``r >= rc``, ewald.rs:390 / wolf.rs:91 at ``r > rc``.  Atoms placed so that some
    separations are exactly the cut-off (representable coordinates), on the neighbour-list path: the interacting
    pair set must be the reference's, not a rounding accident of the list kernels' own distance arithmetic."""
    cutoff = 8.0
    length = 64.0  # 7 cells of 9.14 A per edge
    rng = np.random.Generator(np.random.PCG64(21))
    n_side = 12
    grid = np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), np.arange(n_side), indexing="ij"), axis=-1).reshape(-1, 3)
    positions = (grid + 0.5) * (length / n_side) + rng.uniform(-0.4, 0.4, (len(grid), 3))
    # partners at exactly rc along the axes and along a 3-4-5 / 2-3-6-7 style direction with an exact norm
    base = np.array([16.0, 24.0, 40.0])
    extra = [base, base + [cutoff, 0.0, 0.0], base + [0.0, -cutoff, 0.0], base + [0.0, 4.8, 6.4][::-1],
             base + [-6.0, 3.0, -2.0] * np.array(8.0 / 7.0)]
    # 4.8^2 + 6.4^2 = 64 and (6, 3, 2) * 8/7 has norm 8 only up to rounding: the oracle decides those
    positions = np.concatenate([positions, np.array(extra)])
    # pairs across the periodic boundary, exactly rc apart after the minimum image
    positions = np.concatenate([positions, np.array([[1.0, 32.0, 32.0], [length - 7.0, 32.0, 32.0]])])
    names = ["Ar"] * len(positions)
    system = lumol.System(lumol.UnitCell.cubic(length))
    charges = None
    if charged:
        charges = np.where(np.arange(len(positions)) % 2 == 0, 0.5, -0.5)
    system.add_particles(names, positions, charges=charges, masses=np.full(len(positions), 39.948))
    system.set_pair_potential(("Ar", "Ar"), lumol.PairInteraction(lumol.LennardJones(sigma=3.4, epsilon=units.from_(1.0, "kJ/mol")), cutoff))
    variants = [None]
    if charged:
        variants = [lumol.SharedEwald(lumol.Ewald(cutoff, 6, 0.35)), lumol.Wolf(cutoff)]
    from test_gpu_parity import check_system

    for coulomb in variants:
        if coulomb is not None:
            system.set_coulomb_potential(coulomb)
        device = check_system(system, molecular=False, path=1)
        assert device.stats().neighbor_path == 1
        reference = oracle.OracleSystem(system)
        # the pair count itself (pairs evaluated inside the cut-off) must agree with the all-pairs path, whose distance
        # arithmetic is the reference's
        listed = device.compute(energy=True)
        pairs_list, coulomb_list = device.stats().pair_count, device.stats().coulomb_pair_count
        device.set_neighbor_path(0)
        device.compute(energy=True)
        assert (device.stats().pair_count, device.stats().coulomb_pair_count) == (pairs_list, coulomb_list)
        terms = reference.energy_terms()
        assert abs(listed.energy.pairs - terms.pairs) <= ENERGY_TOL * abs(terms.pairs)
