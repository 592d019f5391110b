"""CUDA path vs the CPU oracle on the same inputs, through the C ABI (``-m gpu``).

Tolerances are the ones BASELINE.json's north_star states: per-atom forces and virials <= 1e-10 relative,
energies <= 1e-9 relative.  "Relative" is taken against the largest force component of the system (forces),
the largest virial component (virials) and the sum of the absolute energy terms (energies; the Ewald total
is a cancellation of much larger terms).
"""

import ctypes
import math

import numpy as np
import pytest

import lumol_b200 as lumol
from lumol_b200 import _ffi, units
from lumol_b200.compute import (
    AtomicVirial, EnergyEvaluator, Forces, KineticEnergy, MolecularVirial, PotentialEnergy, Pressure, Stress,
    Temperature, Virial,
)
from lumol_b200.consts import K_BOLTZMANN
from lumol_b200.device import device_for
from oracle import oracle
from reference_values import EWALD_NIST_VIRIAL, LAMMPS_FORCES, NIST_LJ, NIST_SPCE, round_at
import systems

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10
VIRIAL_TOL = 1e-10
ENERGY_TOL = 1e-9


def assert_forces(actual, expected, tol=FORCE_TOL):
    scale = max(np.abs(expected).max(), 1e-300)
    error = np.abs(actual - expected).max() / scale
    assert error <= tol, f"force error {error:.3e} > {tol:.1e}"


def assert_virial(actual, expected, tol=VIRIAL_TOL):
    scale = max(np.abs(expected).max(), 1e-300)
    error = np.abs(actual - expected).max() / scale
    assert error <= tol, f"virial error {error:.3e} > {tol:.1e}"


def assert_energy_terms(actual, expected, tol=ENERGY_TOL):
    names = ("pairs", "pairs_tail", "bonds", "angles", "dihedrals", "coulomb_real", "coulomb_self", "coulomb_kspace")
    magnitude = max(sum(abs(getattr(expected, name)) for name in names), 1e-300)
    # Wolf: the reference (and the oracle) return pair sum and self term as one number (wolf.rs:177-207);
    # the C ABI reports them as coulomb_real and coulomb_self, so only their sum is comparable.
    wolf = expected.coulomb_self == 0.0 and expected.coulomb_kspace == 0.0
    for name in names:
        if wolf and name.startswith("coulomb"):
            continue
        a, e = getattr(actual, name), getattr(expected, name)
        assert abs(a - e) <= tol * magnitude, f"{name}: {a!r} vs {e!r}"
    coulomb_a = actual.coulomb_real + actual.coulomb_self + actual.coulomb_kspace
    coulomb_e = expected.coulomb_real + expected.coulomb_self + expected.coulomb_kspace
    wolf_magnitude = abs(actual.coulomb_real) + abs(actual.coulomb_self)
    assert abs(coulomb_a - coulomb_e) <= tol * max(magnitude, wolf_magnitude), f"coulomb: {coulomb_a!r} vs {coulomb_e!r}"
    total_a = sum(getattr(actual, name) for name in names)
    total_e = sum(getattr(expected, name) for name in names)
    assert abs(total_a - total_e) <= tol * max(magnitude, wolf_magnitude)


def check_system(system, molecular=True, path=None):
    """Forces, energy terms, atomic (and molecular) virial of ``system`` against the oracle."""
    device = device_for(system)
    if path is not None:
        device.set_neighbor_path(path)
    reference = oracle.OracleSystem(system)
    result = device.compute(forces=True, energy=True, virial=True)
    assert_forces(result.forces, reference.forces())
    assert_energy_terms(result.energy, reference.energy_terms())
    assert_virial(result.virial, reference.atomic_virial())
    # force-only kernels are separate template instances: check them too
    assert_forces(device.compute(forces=True).forces, reference.forces())
    if molecular:
        actual = device.compute(molecular_virial=True, parts=_ffi.PART_PAIRS | _ffi.PART_COULOMB).virial
        assert_virial(actual, reference.molecular_virial())
    return device


# ---- the reference's bench systems (benches/*.rs) ---------------------------------------------------------

def test_argon_bench_system():
    device = check_system(systems.argon())
    assert device.stats().neighbor_path == 0  # L / rc = 2.5: all-pairs minimum image


@pytest.mark.parametrize("coulomb", ["ewald", "wolf"])
def test_nacl_bench_system(coulomb):
    check_system(systems.nacl(coulomb))


@pytest.mark.parametrize("coulomb", ["ewald", "wolf"])
def test_water_bench_system(coulomb):
    check_system(systems.water(coulomb))


def test_propane_bench_system():
    check_system(systems.propane())


def test_global_potentials_standalone():
    """GlobalPotential::{energy, forces, atomic_virial, molecular_virial} on their own (benches/nacl.rs:20-45);
    forces() accumulates into the caller's array (ewald.rs:897-905)."""
    for builder, potential in ((systems.nacl, lumol.SharedEwald(lumol.Ewald(9.5, 7))), (systems.nacl, lumol.Wolf(12.0)),
                               (systems.water, lumol.SharedEwald(lumol.Ewald(8.0, 7)))):
        system = builder()
        if builder is systems.water:
            potential.set_restriction(lumol.PairRestriction.InterMolecular)
        reference = oracle.OracleSystem(system, coulomb=potential)
        terms = reference.energy_terms()
        expected = terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace
        magnitude = abs(terms.coulomb_real) + abs(terms.coulomb_self) + abs(terms.coulomb_kspace)
        assert abs(potential.energy(system) - expected) <= ENERGY_TOL * magnitude
        forces = np.ones((system.size(), 3))
        potential.forces(system, forces)
        assert_forces(forces - 1.0, reference.coulomb_forces())
        assert_virial(potential.atomic_virial(system), reference.coulomb_atomic_virial())
        assert_virial(potential.molecular_virial(system), reference.coulomb_molecular_virial())


# ---- every pair potential, restriction and computation mode ------------------------------------------------

def random_molecular_system(seed, cell, natoms=96):
    """Chains of four atoms of two kinds in a box: exercises bond paths up to three bonds."""
    rng = np.random.Generator(np.random.PCG64(seed))
    system = lumol.System(cell)
    nmol = natoms // 4
    matrix = cell.matrix() if not cell.is_infinite() else np.eye(3) * 12.0
    for _ in range(nmol):
        origin = matrix @ rng.uniform(0.0, 1.0, 3)
        molecule = None
        for k, name in enumerate(("A", "B", "B", "A")):
            particle = lumol.Particle(name, origin + np.array([1.1 * k, 0.0, 0.0]) + rng.uniform(-0.25, 0.25, 3))
            particle.mass = 12.0 + k
            particle.charge = (0.4 if name == "A" else -0.4) * (1.0 + 0.1 * k) - (0.02 if k == 3 else 0.0)
            if molecule is None:
                molecule = lumol.Molecule(particle)
            else:
                molecule.add_particle_bonded_to(k - 1, particle)
        system.add_molecule(molecule)
    return system


POTENTIALS = {
    "lj": lambda: lumol.LennardJones(sigma=2.2, epsilon=2e-4),
    "harmonic": lambda: lumol.Harmonic(k=3e-4, x0=2.5),
    "buckingham": lambda: lumol.Buckingham(a=0.5, c=0.02, rho=0.35),
    "born": lambda: lumol.BornMayerHuggins(a=2e-3, c=0.05, d=0.08, sigma=2.4, rho=0.32),
    "morse": lambda: lumol.Morse(a=1.3, x0=2.6, depth=3e-4),
    "gaussian": lambda: lumol.Gaussian(a=4e-4, b=0.3),
    "mie": lambda: lumol.Mie(sigma=2.3, epsilon=2e-4, n=11.0, m=5.5),
    "table": lambda: lumol.TableComputation(lumol.LennardJones(sigma=2.2, epsilon=2e-4), 800, 5.5),
}

RESTRICTIONS = [
    lumol.PairRestriction.NONE, lumol.PairRestriction.IntraMolecular, lumol.PairRestriction.InterMolecular,
    lumol.PairRestriction.Exclude12, lumol.PairRestriction.Exclude13, lumol.PairRestriction.Exclude14,
    lumol.PairRestriction.Scale14(0.45),
]


@pytest.mark.parametrize("name", sorted(POTENTIALS))
def test_every_pair_potential(name):
    cell = lumol.UnitCell.ortho(11.0, 12.0, 13.0)
    for k, restriction in enumerate(RESTRICTIONS):
        system = random_molecular_system(seed=17 + k, cell=cell)
        for n, pair in enumerate((("A", "A"), ("A", "B"), ("B", "B"))):
            potential = POTENTIALS[name]()
            cutoff = 5.0 - 0.4 * n
            interaction = lumol.PairInteraction.shifted(potential, cutoff) if (k + n) % 2 else lumol.PairInteraction(potential, cutoff)
            interaction.set_restriction(restriction)
            if n != 1:
                interaction.enable_tail_corrections()
            system.set_pair_potential(pair, interaction)
        check_system(system)


def test_missing_and_null_pair_entries():
    """Kind pairs without an entry contribute nothing (interactions.rs:142-145); Null potentials too."""
    system = random_molecular_system(seed=3, cell=lumol.UnitCell.cubic(12.0))
    system.set_pair_potential(("A", "A"), lumol.PairInteraction(lumol.LennardJones(sigma=2.2, epsilon=2e-4), 5.0))
    system.set_pair_potential(("A", "B"), lumol.PairInteraction(lumol.NullPotential(), 5.0))
    check_system(system)


@pytest.mark.parametrize("shape", ["triclinic", "infinite"])
def test_other_cell_shapes(shape):
    if shape == "triclinic":
        cell = lumol.UnitCell.triclinic(12.0, 13.0, 14.0, 80.0, 95.0, 105.0)
    else:
        cell = lumol.UnitCell.infinite()
    system = random_molecular_system(seed=9, cell=cell)
    system.set_pair_potential(("A", "A"), lumol.PairInteraction(lumol.LennardJones(sigma=2.2, epsilon=2e-4), 5.0))
    system.set_pair_potential(("A", "B"), lumol.PairInteraction(lumol.Buckingham(a=0.5, c=0.02, rho=0.35), 4.5))
    system.set_pair_potential(("B", "B"), lumol.PairInteraction(lumol.Harmonic(k=3e-4, x0=2.5), 4.0))
    system.set_bond_potential(("A", "B"), lumol.Harmonic(k=1e-2, x0=1.1))
    system.set_bond_potential(("B", "B"), lumol.Morse(a=1.5, x0=1.1, depth=5e-3))
    system.set_angle_potential(("A", "B", "B"), lumol.CosineHarmonic(k=4e-3, x0=math.radians(170.0)))
    system.set_dihedral_potential(("A", "B", "B", "A"), lumol.Torsion(k=1e-3, delta=0.3, n=2))
    if shape == "triclinic":
        wolf = lumol.Wolf(5.0)
        wolf.set_restriction(lumol.PairRestriction.Exclude13)
        system.set_coulomb_potential(wolf)
        check_system(system)
        ewald = lumol.SharedEwald(lumol.Ewald(5.0, 6))
        ewald.set_restriction(lumol.PairRestriction.InterMolecular)
        system.set_coulomb_potential(ewald)
        check_system(system)
    else:
        device = device_for(system)
        reference = oracle.OracleSystem(system)
        result = device.compute(forces=True, energy=True)
        assert_forces(result.forces, reference.forces())
        assert_energy_terms(result.energy, reference.energy_terms())
        with pytest.raises(lumol.LumolCudaError, match="Can not compute virial for infinite cell"):
            device.compute(virial=True)
        system.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(5.0, 6)))
        with pytest.raises(lumol.LumolCudaError, match="Ewald is not defined with infinite unit cell"):
            system.potential_energy()


def test_wolf_with_scaled_14_and_ewald_exclusions():
    """Wolf honours Scale14 (wolf.rs:196-201); Ewald subtracts the erf correction for excluded pairs (ewald.rs:394-421)."""
    cell = lumol.UnitCell.ortho(12.0, 12.5, 13.0)
    for restriction in RESTRICTIONS:
        system = random_molecular_system(seed=21, cell=cell)
        wolf = lumol.Wolf(5.5)
        wolf.set_restriction(restriction)
        system.set_coulomb_potential(wolf)
        check_system(system)
        if restriction.kind != _ffi.RESTRICTION_SCALE14 and restriction.kind != _ffi.RESTRICTION_INTRA_MOLECULAR:
            ewald = lumol.SharedEwald(lumol.Ewald(5.5, 6, 0.55))
            ewald.set_restriction(restriction)
            system.set_coulomb_potential(ewald)
            check_system(system)
    ewald = lumol.SharedEwald(lumol.Ewald(5.5, 6))
    ewald.set_restriction(lumol.PairRestriction.Scale14(0.5))
    system.set_coulomb_potential(ewald)
    with pytest.raises((ValueError, lumol.LumolCudaError), match="Scaling restriction scheme using Ewald"):
        system.forces()


def test_edge_cases():
    # empty system
    empty = lumol.System(lumol.UnitCell.cubic(10.0))
    assert Forces().compute(empty).shape == (0, 3)
    assert PotentialEnergy().compute(empty) == 0.0
    # one atom; two atoms further apart than the cut-off
    system = lumol.system_from_xyz("""2
    cell: 20.0
    Ar 0.0 0.0 0.0
    Ar 9.0 0.0 0.0
    """)
    lj = lumol.PairInteraction(lumol.LennardJones(sigma=3.4, epsilon=1e-4), 8.0)
    system.set_pair_potential(("Ar", "Ar"), lj)
    np.testing.assert_array_equal(system.forces(), 0.0)
    assert system.potential_energy() == 0.0
    # exactly at the cut-off: pair energy/force are zero for r >= rc (pairs.rs:186, 213); coulomb keeps r == rc (ewald.rs:390)
    system.positions[1][0] = 8.0
    np.testing.assert_array_equal(system.forces(), 0.0)
    system.charges[:] = [1.0, -1.0]
    system.invalidate()
    system.set_coulomb_potential(lumol.Wolf(8.0))
    check_system(system, molecular=False)
    # cut-off larger than half the cell is refused on the host like the reference (system.rs:122-131)
    with pytest.raises(ValueError, match="cutoff bigger than half"):
        system.set_pair_potential(("Ar", "Ar"), lumol.PairInteraction(lumol.NullPotential(), 10.5))


# ---- known answers of the reference's estimator tests, through the public API ----------------------------------

def test_compute_known_answers():
    from test_oracle_kat import molecular_test_system, pairs_test_system, ulps_eq

    system = pairs_test_system()  # compute.rs:567-662
    forces = Forces().compute(system)
    force = units.from_(30.0, "kJ/mol/A")
    assert ulps_eq(forces[0][0], force) and ulps_eq(forces[1][0], -force)
    np.testing.assert_array_equal(forces[0] + forces[1], 0.0)
    assert ulps_eq(KineticEnergy().compute(system), 0.0007483016557453698)
    assert ulps_eq(Temperature().compute(system), 300.0)
    virial = Virial().compute(system)
    assert ulps_eq(virial[0][0], -force * 1.3)
    expected = 2.0 * K_BOLTZMANN * 300.0 / 1000.0 + (-force * 1.3) / (3.0 * 1000.0)
    assert ulps_eq(Pressure().compute(system), expected, max_ulps=8)
    assert ulps_eq(np.trace(Stress().compute(system)) / 3.0, Pressure().compute(system), max_ulps=8)
    assert system.total_energy() == system.kinetic_energy() + system.potential_energy()

    system = molecular_test_system()  # compute.rs:609-613; energy.rs:228-263
    pair = lumol.PairInteraction(lumol.LennardJones(epsilon=units.from_(100.0, "kJ/mol/A^2"), sigma=units.from_(0.8, "A")), 5.0)
    pair.enable_tail_corrections()
    system.set_pair_potential(("F", "F"), pair)
    evaluator = EnergyEvaluator(system)
    assert ulps_eq(evaluator.pairs(), units.from_(-258.3019360389957, "kJ/mol"), max_ulps=8)
    assert ulps_eq(evaluator.pairs_tail(), -0.0000028110338032153973)
    assert ulps_eq(evaluator.bonds(), units.from_(150.0, "kJ/mol"))
    assert ulps_eq(evaluator.angles(), units.from_(400.0, "kJ/mol"))
    assert ulps_eq(evaluator.dihedrals(), units.from_(1250.0, "kJ/mol"), max_ulps=15)


def test_wolf_and_ewald_known_answers():
    from test_oracle_kat import nacl_pair, single_water

    system = lumol.System(lumol.UnitCell.cubic(30.0))  # wolf.rs:29-48
    na = lumol.Particle("Na", (0.0, 0.0, 0.0))
    na.charge = 1.0
    cl = lumol.Particle("Cl", (2.0, 0.0, 0.0))
    cl.charge = -1.0
    system.add_molecule(lumol.Molecule(na))
    system.add_molecule(lumol.Molecule(cl))
    system.set_coulomb_potential(lumol.Wolf(12.0))
    assert abs(system.potential_energy() - -0.0729290269539354) <= 4 * np.finfo(float).eps * 0.073

    pair = nacl_pair()  # ewald.rs:1036-1050
    ewald = lumol.SharedEwald(lumol.Ewald(8.0, 10))
    assert abs(ewald.energy(pair) - -0.09262397663346732) < 1e-4
    lumol.SharedEwald(lumol.Ewald(8.0, 1)).energy(pair)  # "just checking that this does not crash"
    water = single_water()  # ewald.rs:1122-1133
    ewald = lumol.SharedEwald(lumol.Ewald(8.0, 10))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    assert abs(ewald.energy(water) - -0.000009243868813825495) < 1e-14 * 0.1
    # virial is energy (ewald.rs:1216-1231)
    for system in (single_water(), nacl_pair()):
        ewald = lumol.SharedEwald(lumol.Ewald(8.0, 10))
        assert math.isclose(ewald.energy(system), np.trace(ewald.atomic_virial(system)), rel_tol=1e-3)


# ---- NIST reference calculations and LAMMPS forces on the device -------------------------------------------------

@pytest.mark.parametrize("index,cutoff", sorted(NIST_LJ))
@pytest.mark.parametrize("path", [0, 1])
def test_nist_lennard_jones(index, cutoff, path):
    (energy_ref, e_dec), (virial_ref, v_dec), (tail_ref, t_dec) = NIST_LJ[(index, cutoff)]
    system = systems.nist_lj(index, cutoff, tail=False)
    if path == 1 and min(system.cell.lengths()) / cutoff < 3.0:
        pytest.skip("fewer than three cells per edge: the cell list does not apply")
    device_for(system).set_neighbor_path(path)
    energy = system.potential_energy()
    assert round_at(energy, e_dec) == energy_ref
    assert round_at(np.trace(system.virial()), v_dec) == virial_ref
    with_tail = systems.nist_lj(index, cutoff, tail=True)
    device_for(with_tail).set_neighbor_path(path)
    assert round_at(with_tail.potential_energy() - energy, t_dec) == tail_ref
    assert device_for(system).stats().neighbor_path == path


@pytest.mark.parametrize("index,cutoff", sorted(NIST_SPCE))
def test_nist_spce_energies(index, cutoff):
    total, pairs, tail, coulomb = NIST_SPCE[(index, cutoff)]
    system = systems.nist_spce(index)
    systems.set_nist_interactions(system, cutoff)
    assert abs((system.potential_energy() / K_BOLTZMANN - total) / total) < 1e-3
    evaluator = system.energy_evaluator()
    assert abs((evaluator.pairs() / K_BOLTZMANN - pairs) / pairs) < 1e-3
    assert abs((evaluator.pairs_tail() / K_BOLTZMANN - tail) / tail) < 1e-3
    assert abs((evaluator.coulomb() / K_BOLTZMANN - coulomb) / coulomb) < 1e-3


@pytest.mark.parametrize("index,cutoff", sorted(LAMMPS_FORCES))
def test_lammps_forces(index, cutoff, golden):
    kmax, alpha, dynamic = LAMMPS_FORCES[(index, cutoff)]
    system = systems.nist_spce(index)
    systems.set_lammps_interactions(system, float(cutoff), kmax, alpha)
    forces = system.forces() / units.from_(1.0, "kcal/mol/A")
    expected = golden[f"lammps-forces-{cutoff}-{index}/forces"]
    relative = np.abs((forces - expected) / expected)
    if dynamic:
        tolerance = np.where(np.abs(expected) < 1e-1, 1e-1, np.where(np.abs(expected) < 1.0, 5e-2, 1e-2))
    else:
        tolerance = 5e-3
    assert np.all(relative < tolerance)


@pytest.mark.parametrize("index,cutoff", sorted(EWALD_NIST_VIRIAL))
def test_nist_spce_virials(index, cutoff):
    kmax, alpha, real, real_tol, kspace, k_tol = EWALD_NIST_VIRIAL[(index, cutoff)]
    system = systems.nist_spce(index)
    ewald = lumol.SharedEwald(lumol.Ewald(cutoff, kmax, alpha))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    convert = units.from_(1.0, "atm") * system.volume()
    total = ewald.atomic_virial(system) / convert
    # the reference compares the real and k-space parts separately (relative, element-wise); for their sum that is
    # an absolute bound of tol * (|real| + |k-space|) per element
    bound = real_tol * np.abs(np.array(real)) + k_tol * np.abs(np.array(kspace))
    assert np.all(np.abs(total - (np.array(real) + np.array(kspace))) <= bound)
    assert math.isclose(ewald.energy(system), np.trace(ewald.atomic_virial(system)), rel_tol=1e-3)


# ---- cell list against all-pairs and the oracle -----------------------------------------------------------------

def test_nist_spce_full_parity_both_paths():
    """2250 atoms, L = 30, rc = 10: exactly three cells per edge; LJ + Ewald with molecular exclusions."""
    system = systems.nist_spce(4)
    systems.set_nist_interactions(system, 10.0)
    reference = oracle.OracleSystem(system)
    expected_forces, expected_terms, expected_virial = reference.forces(), reference.energy_terms(), reference.atomic_virial()
    for path in (0, 1):
        device = device_for(system)
        device.set_neighbor_path(path)
        result = device.compute(forces=True, energy=True, virial=True)
        assert device.stats().neighbor_path == path
        assert_forces(result.forces, expected_forces)
        assert_energy_terms(result.energy, expected_terms)
        assert_virial(result.virial, expected_virial)
        assert_forces(device.compute(forces=True).forces, expected_forces)


def test_lj_box_cell_list_vs_oracle():
    """4096-atom synthetic argon box: L / rc = 5.77, LJ fast path of the cell kernel."""
    system = systems.lj_box(16, seed=20240 + 12)
    device = check_system(system, molecular=False, path=1)
    stats = device.stats()
    assert stats.neighbor_path == 1 and tuple(stats.ncells) == (5, 5, 5)
    # unwrapped coordinates (lumol never wraps unless asked, controls.rs:78-87): shift atoms by whole cells
    shifted = systems.lj_box(16, seed=20240 + 12)
    rng = np.random.Generator(np.random.PCG64(7))
    shifted.positions += rng.integers(-3, 4, shifted.positions.shape) * shifted.cell.a()
    check_system(shifted, molecular=False, path=1)
    # shifted cut-off and a non-cubic cell go through the general kernel
    general = systems.lj_box(12, seed=5)
    general.cell = lumol.UnitCell.ortho(general.cell.a(), general.cell.a() * 1.13, general.cell.a() * 0.97)
    potential = lumol.LennardJones(sigma=3.4, epsilon=units.from_(1.0, "kJ/mol"))
    general.set_pair_potential(("Ar", "Ar"), lumol.PairInteraction.shifted(potential, 9.0))
    check_system(general, molecular=False, path=1)


def test_staged_lj_kernel_vs_oracle_and_global_format():
    """32 768-atom argon box (10 x 10 x 10 cells): the Lennard-Jones kernel that stages each block's neighbourhood in
    shared memory (16-bit list entries) against the oracle and against the same kernel with every block in the
    global list format (neighbour path 2)."""
    system = systems.lj_box(32, seed=20240 + 15)
    device = check_system(system, molecular=False, path=1)
    assert tuple(device.stats().ncells) == (10, 10, 10)
    staged = device.compute(forces=True, energy=True, virial=True)
    device.set_neighbor_path(2)
    plain = device.compute(forces=True, energy=True, virial=True)
    scale = np.abs(plain.forces).max()
    assert np.abs(staged.forces - plain.forces).max() < 1e-12 * scale
    assert abs(staged.energy.pairs - plain.energy.pairs) < 1e-12 * abs(plain.energy.pairs)
    assert np.abs(staged.virial - plain.virial).max() < 1e-12 * np.abs(plain.virial).max()
    # a flat box: blocks span several rows of cells and wrap around the periodic boundary inside a block
    flat = lumol.System(lumol.UnitCell.ortho(34.0, 35.0, 140.0))
    rng = np.random.Generator(np.random.PCG64(11))
    points = np.stack(np.meshgrid(np.arange(9), np.arange(9), np.arange(36), indexing="ij"), axis=-1).reshape(-1, 3)
    positions = (points + 0.5) * np.array([34.0 / 9, 35.0 / 9, 140.0 / 36]) + rng.uniform(-0.3, 0.3, (len(points), 3))
    flat.add_particles(["Ar"] * len(positions), positions)
    potential = lumol.LennardJones(sigma=3.4, epsilon=units.from_(1.0, "kJ/mol"))
    flat.set_pair_potential(("Ar", "Ar"), lumol.PairInteraction(potential, 10.0))
    device = check_system(flat, molecular=False, path=1)
    assert tuple(device.stats().ncells) == (3, 3, 12)


def test_staged_lj_kernel_full_size():
    """The bench box (1 048 576 atoms): staged kernel against the global-format kernel, Newton's third law."""
    from lumol_b200 import synthetic

    system = synthetic.lj_box((128, 128, 64), seed=20240 + 20)
    device = device_for(system)
    staged = device.compute(forces=True, energy=True, virial=True)
    assert device.stats().neighbor_path == 1
    pairs_staged = device.stats().pair_count
    device.set_neighbor_path(2)
    plain = device.compute(forces=True, energy=True, virial=True)
    assert device.stats().pair_count == pairs_staged
    scale = np.abs(plain.forces).max()
    assert np.abs(staged.forces - plain.forces).max() < 1e-12 * scale
    assert np.abs(staged.forces.sum(axis=0)).max() < 1e-9 * scale * np.sqrt(system.size())
    assert abs(staged.energy.pairs - plain.energy.pairs) < 1e-12 * abs(plain.energy.pairs)
    assert np.abs(staged.virial - plain.virial).max() < 1e-11 * np.abs(plain.virial).max()


def test_non_finite_positions_on_the_cell_path():
    """An exploded simulation must give an error, not a hang: the rebuild stops when a position is not finite, the
    force kernels return, and the context works again once the positions are sane."""
    system = systems.lj_box(16, seed=3)
    good = system.positions.copy()
    expected = device_for(system).compute(forces=True).forces
    system.positions[17, 1] = np.nan
    with pytest.raises(lumol.LumolCudaError, match="not finite"):
        device_for(system).compute(forces=True)
    system.positions[:] = good
    again = device_for(system).compute(forces=True).forces
    assert np.abs(again - expected).max() <= 1e-12 * np.abs(expected).max()


def test_water_box_cell_list_vs_oracle():
    """5184-atom synthetic SPC/E box: LJ + Ewald real space with intra-molecular exclusions on the cell path."""
    system = systems.spce_box(12)
    ewald = lumol.SharedEwald(lumol.Ewald(9.0, 6, 0.32))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    device = check_system(system, molecular=False, path=1)
    assert device.stats().neighbor_path == 1
    wolf = lumol.Wolf(9.0)
    wolf.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(wolf)
    check_system(system, molecular=False, path=1)


def test_multi_kind_cell_list_with_restrictions():
    """Two kinds, tables, Scale14 and bonded terms through the general cell kernel."""
    cell = lumol.UnitCell.ortho(16.0, 17.0, 18.0)
    system = random_molecular_system(seed=31, cell=cell, natoms=400)
    table = lumol.TableComputation(lumol.Buckingham(a=0.5, c=0.02, rho=0.35), 900, 5.2)
    first = lumol.PairInteraction(lumol.LennardJones(sigma=2.2, epsilon=2e-4), 5.0)
    first.set_restriction(lumol.PairRestriction.Scale14(0.5))
    second = lumol.PairInteraction(table, 4.8)
    second.set_restriction(lumol.PairRestriction.Exclude12)
    system.set_pair_potential(("A", "A"), first)
    system.set_pair_potential(("A", "B"), second)
    system.set_pair_potential(("B", "B"), lumol.PairInteraction.shifted(lumol.Mie(sigma=2.3, epsilon=2e-4, n=11.0, m=5.5), 4.5))
    system.set_bond_potential(("A", "B"), lumol.Harmonic(k=1e-2, x0=1.1))
    system.set_angle_potential(("A", "B", "B"), lumol.Harmonic(k=4e-3, x0=math.radians(170.0)))
    system.set_dihedral_potential(("A", "B", "B", "A"), lumol.Torsion(k=1e-3, delta=0.3, n=2))
    wolf = lumol.Wolf(5.2)
    wolf.set_restriction(lumol.PairRestriction.Exclude13)
    system.set_coulomb_potential(wolf)
    check_system(system, path=1)


# ---- Ewald reciprocal space --------------------------------------------------------------------------------------

def test_ewald_structure_factor():
    system = systems.nacl("ewald")
    device = device_for(system)
    device.compute(energy=True, parts=_ffi.PART_COULOMB)
    count = ctypes.c_int64()
    lib, ctx = device.lib, device.ctx
    _ffi.check(ctx, lib.lumol_cuda_ewald_kvectors(ctx, 0, ctypes.byref(count), None, None, None))
    nk = count.value
    assert nk == 706  # Ewald(9.5, 7) in a 25 A cube
    index = np.zeros((nk, 3), dtype=np.int32)
    factor = np.zeros(nk)
    rho = np.zeros((nk, 2))
    _ffi.check(ctx, lib.lumol_cuda_ewald_kvectors(ctx, nk, ctypes.byref(count), index.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                  _ffi.as_double_pointer(factor), _ffi.as_double_pointer(rho)))
    reference = oracle.OracleSystem(system)
    _, ref_index, ref_energy, _, _ = reference.ewald_factors()
    np.testing.assert_array_equal(index, ref_index)
    np.testing.assert_allclose(factor, ref_energy, rtol=1e-14)
    ref_rho = reference.ewald_rho(nk)
    assert np.abs(rho - ref_rho).max() <= 1e-11 * np.abs(ref_rho).max()
    # triclinic cell and large kmax
    cell = lumol.UnitCell.triclinic(12.0, 13.0, 14.0, 80.0, 95.0, 105.0)
    system = random_molecular_system(seed=4, cell=cell)
    ewald = lumol.SharedEwald(lumol.Ewald(5.0, 14, 0.7))
    system.set_coulomb_potential(ewald)
    check_system(system)


def test_tiled_kspace_kernels_vs_oracle_and_direct():
    """The register-tiled reciprocal-space kernels (sharing the +l / -l products) against the oracle and against the
    direct kernels: 5184-atom SPC/E box, kmax 10 and an odd kmax with fewer |l| values than a thread tile, and a
    triclinic cell whose rows of the k list are not symmetric in l."""
    for kmax, alpha in ((10, 0.32), (13, 0.36)):
        system = systems.spce_box(12)
        ewald = lumol.SharedEwald(lumol.Ewald(9.0, kmax, alpha))
        ewald.set_restriction(lumol.PairRestriction.InterMolecular)
        system.set_coulomb_potential(ewald)
        device = device_for(system)
        device.set_kspace_algorithm(1)
        device = check_system(system, molecular=False, path=1)
        tiled = device.compute(forces=True, energy=True, virial=True, parts=_ffi.PART_COULOMB)
        device.set_kspace_algorithm(0)
        direct = device.compute(forces=True, energy=True, virial=True, parts=_ffi.PART_COULOMB)
        assert np.abs(tiled.forces - direct.forces).max() < 1e-11 * np.abs(direct.forces).max()
        assert abs(tiled.energy.coulomb_kspace - direct.energy.coulomb_kspace) < 1e-11 * abs(direct.energy.coulomb_kspace)
        assert np.abs(tiled.virial - direct.virial).max() < 1e-10 * np.abs(direct.virial).max()
    cell = lumol.UnitCell.triclinic(22.0, 23.0, 24.0, 80.0, 95.0, 105.0)
    system = random_molecular_system(seed=8, cell=cell, natoms=480)
    system.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(6.0, 12, 0.5)))
    device_for(system).set_kspace_algorithm(1)
    check_system(system)


def test_tiled_kspace_kernels_large_kmax():
    """kmax above 26: the tiled force kernel keeps only the e_z phase table in shared memory and e_x, e_y in its
    global scratch (kmax = 54 is what Ewald::with_accuracy(9 A, 1e-5) gives a 1M-atom SPC/E box).  kmax 30 on a box
    where every block walks several atom tiles and several row splits; kmax 60 on a small one."""
    for n_side, kmax, alpha in ((10, 30, 0.45), (6, 60, 0.6)):
        # kmax 60 exceeds what the direct force kernel holds in shared memory: the automatic choice is the tiled one
        system = systems.spce_box(n_side)
        ewald = lumol.SharedEwald(lumol.Ewald(9.0, kmax, alpha))
        ewald.set_restriction(lumol.PairRestriction.InterMolecular)
        system.set_coulomb_potential(ewald)
        device = device_for(system)
        device.set_kspace_algorithm(1 if kmax == 30 else -1)
        reference = oracle.OracleSystem(system)
        tiled = device.compute(forces=True, energy=True, virial=True, parts=_ffi.PART_COULOMB)
        assert_forces(tiled.forces, reference.coulomb_forces())
        assert_energy_terms(tiled.energy, ewald_only_terms(reference))
        assert_virial(tiled.virial, reference.coulomb_atomic_virial())
        if kmax == 30:
            device.set_kspace_algorithm(0)
            direct = device.compute(forces=True, parts=_ffi.PART_COULOMB)
            assert np.abs(tiled.forces - direct.forces).max() < 1e-11 * np.abs(direct.forces).max()


def ewald_only_terms(reference):
    terms = reference.energy_terms()
    for name in ("pairs", "pairs_tail", "bonds", "angles", "dihedrals"):
        setattr(terms, name, 0.0)
    return terms


# ---- kinetic estimators ----------------------------------------------------------------------------------------------

def test_kinetic_estimators():
    system = systems.propane()
    systems.random_velocities(system, 250.0, seed=2)
    reference = oracle.OracleSystem(system)
    assert abs(system.kinetic_energy() - reference.kinetic_energy()) <= 1e-13 * reference.kinetic_energy()
    assert abs(system.temperature() - reference.temperature()) <= 1e-13 * reference.temperature()
    assert abs(system.pressure() - reference.pressure()) <= 1e-10 * abs(reference.pressure())
    assert_virial(system.stress(), reference.stress())
    system.simulated_degrees_of_freedom = ("molecules", 0)
    system.invalidate()
    reference = oracle.OracleSystem(system)
    assert abs(system.temperature() - reference.temperature()) <= 1e-13 * reference.temperature()
    assert_virial(system.virial(), reference.molecular_virial())
    assert abs(system.pressure() - reference.pressure()) <= 1e-9 * abs(reference.pressure())
