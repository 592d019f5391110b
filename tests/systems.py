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


from lumol_b200.synthetic import lj_box, spce_box  # noqa: E402,F401  (synthetic boxes of SURVEY section 8d)
