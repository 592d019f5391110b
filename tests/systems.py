"""Builders of the systems the reference's own tests and benches use, through the host-side API.

Every builder cites the reference input it reproduces (paths relative to the reference checkout).
"""

import os

import numpy as np

import lumol_b200 as lumol
from lumol_b200 import units
from lumol_b200.consts import K_BOLTZMANN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = np.load(os.path.join(ROOT, "tests", "golden", "lumol_fixtures.npz"))
    return _golden


def from_fixture(name, charges=None):
    """System from one configuration of the fixture archive (tests/golden/make_fixtures.py)."""
    data = golden()
    names = [str(n) for n in data[name + "/names"]]
    cell = data[name + "/cell"]
    system = lumol.System(lumol.UnitCell.ortho(cell[0], cell[1], cell[2]))
    q = None if charges is None else np.array([charges[n] for n in names])
    system.add_particles(names, data[name + "/positions"], charges=q)
    key = name + "/bonds"
    if key in data.files and len(data[key]):
        system.add_bonds(data[key])
    return system


def argon():
    """benches/data/argon.{pdb,toml}: 300 Ar, L = 25, LJ sigma 3.4 A, eps 1 kJ/mol, rc 10 A, tail corrections."""
    system = from_fixture("bench-argon")
    lj = lumol.PairInteraction(lumol.LennardJones(sigma=units.from_(3.4, "A"), epsilon=units.from_(1.0, "kJ/mol")), 10.0)
    lj.enable_tail_corrections()
    system.set_pair_potential(("Ar", "Ar"), lj)
    return system


def nacl(coulomb="ewald"):
    """benches/data/nacl.{pdb,toml} + benches/nacl.rs:10-16: null pairs, Ewald(9.5, 7) or Wolf(12)."""
    system = from_fixture("bench-nacl", charges={"Na": 1.0, "Cl": -1.0})
    for pair in (("Na", "Na"), ("Cl", "Cl"), ("Na", "Cl")):
        system.set_pair_potential(pair, lumol.PairInteraction(lumol.NullPotential(), 8.0))
    if coulomb == "ewald":
        system.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(9.5, 7)))
    elif coulomb == "wolf":
        system.set_coulomb_potential(lumol.Wolf(12.0))
    return system


def water(coulomb="ewald"):
    """benches/data/water.{pdb,toml} + benches/water.rs:10-20: 50 rigid waters, null pairs/bonds/angles,
    Ewald(8, 7) or Wolf(9), InterMolecular."""
    system = from_fixture("bench-water", charges={"O": -0.82, "H": 0.41})
    for pair in (("O", "O"), ("H", "H"), ("O", "H")):
        system.set_pair_potential(pair, lumol.PairInteraction(lumol.NullPotential(), 8.0))
    system.set_bond_potential(("O", "H"), lumol.NullPotential())
    system.set_angle_potential(("H", "O", "H"), lumol.NullPotential())
    if coulomb == "ewald":
        potential = lumol.SharedEwald(lumol.Ewald(8.0, 7))
    else:
        potential = lumol.Wolf(9.0)
    potential.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(potential)
    return system


def propane():
    """benches/data/propane.{pdb,toml}: 20 propane, LJ rc 10 + harmonic bonds/angles + torsions."""
    system = from_fixture("bench-propane")
    kcal = units.from_(1.0, "kcal/mol")
    for pair, (epsilon, sigma) in {
        ("C", "C"): (0.1078, 3.8138), ("H", "C"): (0.04735, 3.3662), ("H", "H"): (0.0208, 2.9186),
    }.items():
        system.set_pair_potential(pair, lumol.PairInteraction(lumol.LennardJones(sigma=sigma, epsilon=epsilon * kcal), 10.0))
    system.set_bond_potential(("C", "C"), lumol.Harmonic(k=232.52 * kcal, x0=1.538))
    system.set_bond_potential(("C", "H"), lumol.Harmonic(k=375.92 * kcal, x0=1.097))
    deg = units.from_(1.0, "deg")
    system.set_angle_potential(("C", "C", "C"), lumol.Harmonic(k=64.888 * kcal, x0=111.510 * deg))
    system.set_angle_potential(("C", "C", "H"), lumol.Harmonic(k=46.816 * kcal, x0=109.800 * deg))
    system.set_angle_potential(("H", "C", "H"), lumol.Harmonic(k=38.960 * kcal, x0=107.580 * deg))
    system.set_dihedral_potential(("H", "C", "C", "H"), lumol.Torsion(k=0.240 * kcal, delta=0.0, n=3))
    system.set_dihedral_potential(("H", "C", "C", "C"), lumol.Torsion(k=0.080 * kcal, delta=0.0, n=3))
    return system


def nist_lj(index, cutoff, tail):
    """tests/nist-lj.rs:41-55: reduced-unit LJ (sigma = epsilon = 1) on tests/data/nist-lj/lj-N.xyz."""
    system = from_fixture(f"nist-lj-{index}")
    lj = lumol.PairInteraction(lumol.LennardJones(sigma=1.0, epsilon=1.0), cutoff)
    if tail:
        lj.enable_tail_corrections()
    system.set_pair_potential(("H", "H"), lj)
    return system


def nist_spce(index):
    """tests/nist-spce.rs:24-63: SPC/E geometry, O-H bonds for every third atom, charges +-0.4238."""
    system = from_fixture(f"nist-spce-{index}", charges={"H": 0.42380, "O": -2.0 * 0.42380})
    n = system.size()
    bonds = [(i, i + 1) for i in range(0, n, 3)] + [(i, i + 2) for i in range(0, n, 3)]
    system.add_bonds(np.array(bonds))
    return system


def set_nist_interactions(system, cutoff):
    """tests/nist-spce.rs:65-83"""
    lj = lumol.PairInteraction(lumol.LennardJones(epsilon=78.19743111 * K_BOLTZMANN, sigma=3.16555789), cutoff)
    lj.enable_tail_corrections()
    system.set_pair_potential(("O", "O"), lj)
    system.set_pair_potential(("O", "H"), lumol.PairInteraction(lumol.NullPotential(), cutoff))
    system.set_pair_potential(("H", "H"), lumol.PairInteraction(lumol.NullPotential(), cutoff))
    alpha = 5.6 / min(min(system.cell.a(), system.cell.b()), system.cell.c())
    ewald = lumol.SharedEwald(lumol.Ewald(cutoff, 5, alpha))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)


def set_lammps_interactions(system, cutoff, kmax, alpha):
    """tests/nist-spce.rs:88-103"""
    lj = lumol.PairInteraction(lumol.LennardJones(epsilon=78.19743111 * K_BOLTZMANN, sigma=3.16555789), cutoff)
    system.set_pair_potential(("O", "O"), lj)
    system.set_pair_potential(("O", "H"), lumol.PairInteraction(lumol.NullPotential(), cutoff))
    system.set_pair_potential(("H", "H"), lumol.PairInteraction(lumol.NullPotential(), cutoff))
    ewald = lumol.SharedEwald(lumol.Ewald(cutoff, kmax, alpha))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)


def md_helium(shifted=False, table=False):
    """tests/data/md-helium/nve-{velocity-verlet,shifted,table}.toml: 125 He, L = 10, LJ sigma 2 eps 0.2 kJ/mol rc 4.5."""
    system = from_fixture("md-helium")
    potential = lumol.LennardJones(sigma=2.0, epsilon=units.from_(0.2, "kJ/mol"))
    if table:
        potential = lumol.TableComputation(potential, 1000, 5.0)  # computation = {table = {max = "5 A", n = 1000}}
    if shifted:
        interaction = lumol.PairInteraction.shifted(potential, 4.5)
    else:
        interaction = lumol.PairInteraction(potential, 4.5)
    system.set_pair_potential(("He", "He"), interaction)
    return system


def md_nacl(coulomb="ewald"):
    """tests/data/md-nacl/{ewald,wolf}.toml on small.xyz (64 ions, L = 11.2804)."""
    system = from_fixture("md-nacl-small", charges={"Na": 1.0, "Cl": -1.0})
    kcal = units.from_(1.0, "kcal/mol")
    for pair, (sigma, epsilon) in {("Na", "Na"): (2.497, 0.07826), ("Cl", "Cl"): (4.612, 0.02502), ("Na", "Cl"): (3.5545, 0.04425)}.items():
        system.set_pair_potential(pair, lumol.PairInteraction(lumol.LennardJones(sigma=sigma, epsilon=epsilon * kcal), 5.5))
    if coulomb == "ewald":
        system.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(5.5, 10)))
    else:
        system.set_coulomb_potential(lumol.Wolf(5.5))
    return system


def md_water():
    """tests/data/md-water/ewald.toml on small.pdb: flexible f-SPC water, Ewald(8.5, 8) inter-molecular."""
    system = from_fixture("md-water-small", charges={"O": -0.82, "H": 0.41})
    kcal = units.from_(1.0, "kcal/mol")
    system.set_pair_potential(("O", "O"), lumol.PairInteraction(lumol.LennardJones(sigma=3.16, epsilon=0.155 * kcal), 14.0))
    hh = lumol.PairInteraction(lumol.Harmonic(k=79.8 * kcal, x0=1.633), 14.0)
    hh.set_restriction(lumol.PairRestriction.IntraMolecular)
    system.set_pair_potential(("H", "H"), hh)
    system.set_pair_potential(("H", "O"), lumol.PairInteraction(lumol.NullPotential(), 14.0))
    system.set_bond_potential(("O", "H"), lumol.Harmonic(k=1054.2 * kcal, x0=1.0))
    system.set_angle_potential(("H", "O", "H"), lumol.Harmonic(k=75.9 * kcal, x0=units.from_(109.5, "deg")))
    ewald = lumol.SharedEwald(lumol.Ewald(8.5, 8))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    return system


def random_velocities(system, temperature, seed):
    """Maxwell-Boltzmann velocities from a seeded numpy generator (not parity relevant)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sigma = np.sqrt(K_BOLTZMANN * temperature / system.masses)[:, None]
    system.velocities = rng.standard_normal((system.size(), 3)) * sigma


def lj_box(n_side, density=0.0213, sigma=3.4, epsilon_kjmol=1.0, cutoff=10.0, jitter=0.3, seed=0, tail=True, name="Ar"):
    """Synthetic LJ argon box of SURVEY section 8d: simple-cubic lattice at the given number density,
    positions jittered by U(-jitter, jitter) per axis."""
    n = n_side ** 3
    length = (n / density) ** (1.0 / 3.0)
    spacing = length / n_side
    grid = np.arange(n_side) * spacing
    z, y, x = np.meshgrid(grid, grid, grid, indexing="ij")
    positions = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1) + 0.5 * spacing
    rng = np.random.Generator(np.random.PCG64(seed))
    positions += rng.uniform(-jitter, jitter, positions.shape)
    system = lumol.System(lumol.UnitCell.cubic(length))
    system.add_particles([name] * n, positions, masses=np.full(n, 39.948))
    lj = lumol.PairInteraction(lumol.LennardJones(sigma=sigma, epsilon=units.from_(epsilon_kjmol, "kJ/mol")), cutoff)
    if tail:
        lj.enable_tail_corrections()
    system.set_pair_potential((name, name), lj)
    return system


def spce_box(n_side, cutoff=9.0, seed=777, density=0.0334, flexible=False):
    """Synthetic SPC/E box of SURVEY section 8d: n_side^3 waters on a cubic lattice with random
    orientations; charges and O-O LJ of tests/nist-spce.rs:56-73."""
    nmol = n_side ** 3
    length = (nmol / density) ** (1.0 / 3.0)
    spacing = length / n_side
    rng = np.random.Generator(np.random.PCG64(seed))
    grid = (np.arange(n_side) + 0.5) * spacing
    z, y, x = np.meshgrid(grid, grid, grid, indexing="ij")
    centers = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    half = np.radians(109.47) / 2.0
    h1 = np.array([np.sin(half), np.cos(half), 0.0])
    h2 = np.array([-np.sin(half), np.cos(half), 0.0])
    # random rotations from normalised quaternions
    q = rng.standard_normal((nmol, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, a, b, c = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot = np.empty((nmol, 3, 3))
    rot[:, 0, 0] = 1 - 2 * (b * b + c * c); rot[:, 0, 1] = 2 * (a * b - c * w); rot[:, 0, 2] = 2 * (a * c + b * w)
    rot[:, 1, 0] = 2 * (a * b + c * w); rot[:, 1, 1] = 1 - 2 * (a * a + c * c); rot[:, 1, 2] = 2 * (b * c - a * w)
    rot[:, 2, 0] = 2 * (a * c - b * w); rot[:, 2, 1] = 2 * (b * c + a * w); rot[:, 2, 2] = 1 - 2 * (a * a + b * b)
    positions = np.empty((nmol, 3, 3))
    positions[:, 0] = centers
    positions[:, 1] = centers + rot @ h1
    positions[:, 2] = centers + rot @ h2
    names = ["O", "H", "H"] * nmol
    charges = np.tile([-0.8476, 0.4238, 0.4238], nmol)
    masses = np.tile([15.999, 1.008, 1.008], nmol)
    system = lumol.System(lumol.UnitCell.cubic(length))
    system.add_particles(names, positions.reshape(-1, 3), charges=charges, masses=masses)
    first = np.arange(0, 3 * nmol, 3)
    system.add_bonds(np.concatenate([np.stack([first, first + 1], axis=1), np.stack([first, first + 2], axis=1)]))
    lj = lumol.PairInteraction(lumol.LennardJones(epsilon=78.19743111 * K_BOLTZMANN, sigma=3.16555789), cutoff)
    lj.enable_tail_corrections()
    system.set_pair_potential(("O", "O"), lj)
    system.set_pair_potential(("O", "H"), lumol.PairInteraction(lumol.NullPotential(), cutoff))
    system.set_pair_potential(("H", "H"), lumol.PairInteraction(lumol.NullPotential(), cutoff))
    if flexible:
        kcal = units.from_(1.0, "kcal/mol")
        system.set_bond_potential(("O", "H"), lumol.Harmonic(k=1054.2 * kcal, x0=1.0))
        system.set_angle_potential(("H", "O", "H"), lumol.Harmonic(k=75.9 * kcal, x0=units.from_(109.5, "deg")))
    return system
