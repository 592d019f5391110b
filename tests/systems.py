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


# ---- Monte Carlo energy cache: the systems and trial positions of the reference's own cache tests --------------

def cache_testing_system():
    """sys/cache.rs:306-394 ``testing_system``: two H-O-O-H molecules, LJ / null pairs, harmonic bonded terms, Wolf(5)."""
    system = lumol.system_from_xyz("""8
    cell: 10.0
    H     0.895669     0.000000    -0.316667
    O     0.000000     0.000000     0.000000
    O     0.000000     0.000000     1.480000
    H    -0.895669     0.000000     1.796667
    O     3.000000     0.000000     0.000000
    O     3.000000     0.000000     1.480000
    H     3.895669     0.000000    -0.316667
    H     2.104330     0.000000     1.796667
    """)
    for (i, j) in ((0, 1), (0, 2), (1, 3), (4, 5), (4, 6), (5, 7)):
        assert system.add_bond(i, j) == []
    assert len(system.molecules()) == 2
    system.set_pair_potential(("H", "H"), lumol.PairInteraction(lumol.LennardJones(sigma=3.0, epsilon=units.from_(0.5, "kJ/mol")), 3.0))
    system.set_pair_potential(("O", "O"), lumol.PairInteraction(lumol.NullPotential(), 3.0))
    system.set_pair_potential(("O", "H"), lumol.PairInteraction(lumol.LennardJones(sigma=1.0, epsilon=units.from_(0.3, "kJ/mol")), 3.0))
    system.set_bond_potential(("O", "O"), lumol.Harmonic(x0=2.4, k=units.from_(522.0, "kJ/mol/A^2")))
    system.set_bond_potential(("O", "H"), lumol.Harmonic(x0=1.4, k=units.from_(122.0, "kJ/mol/A^2")))
    system.set_angle_potential(("O", "O", "H"), lumol.Harmonic(x0=np.radians(120.0), k=units.from_(150.0, "kJ/mol/deg^2")))
    system.set_dihedral_potential(("H", "O", "O", "H"), lumol.Harmonic(x0=np.radians(180.0), k=units.from_(800.0, "kJ/mol/deg^2")))
    system.set_coulomb_potential(lumol.Wolf(5.0))
    system.charges[:] = [-0.5 if name == "O" else 0.5 for name in system.names]
    system.invalidate()
    return system


# cache.rs:413-418 and 434-439: two successive trial positions of molecule 0
CACHE_MOVES = (
    np.array([
        [-0.987061, 0.59401, 0.427533],
        [-1.0744137409578138, 1.2111820514074991, -0.2893833856814936],
        [-1.4352068561309008, 2.5425486908430286, 0.24698514382209652],
        [-1.5225595970887147, 3.159720742250528, -0.46993124185939705],
    ]),
    np.array([
        [-0.49138099999999996, 1.08969, 0.923213],
        [-1.0139188773839494, 1.7555242433058806, 1.3546291257885033],
        [-2.4014453359903767, 1.2663193861239206, 1.5154051637316108],
        [-2.923983213374326, 1.9321536294298012, 1.946821289520114],
    ]),
)
# cache.rs:457 and 474: rigid translations of molecule 0, then of every molecule
CACHE_TRANSLATIONS = (np.array([1.0, 0.5, -0.5]), np.array([-0.9, 0.0, 1.8]))


def wolf_cache_system():
    """energy/global/wolf.rs:447-474: two SPC/E-charged waters in a 20 A box, O first."""
    system = lumol.system_from_xyz("""6
    cell: 20.0
    O  0.0  0.0  0.0
    H -0.7 -0.7  0.3
    H  0.3 -0.3 -0.8
    O  2.0  2.0  0.0
    H  1.3  1.3  0.3
    H  2.3  1.7 -0.8
    """)
    for (i, j) in ((0, 1), (0, 2), (3, 4), (3, 5)):
        assert system.add_bond(i, j) == []
    assert len(system.molecules()) == 2
    system.charges[:] = [-0.8476 if name == "O" else 0.4238 for name in system.names]
    system.invalidate()
    return system


# wolf.rs:486-490
WOLF_CACHE_MOVE = np.array([
    [4.0, 0.0, -2.0],
    [3.010010191494968, 0.19045656166589708, -2.1166435218719863],
    [4.0761078062722484, -0.8995901989882638, -2.0703212322750546],
])


def ewald_cache_system():
    """energy/global/ewald.rs:1291-1312: the same two waters with the oxygen in the middle of each molecule."""
    system = lumol.system_from_xyz("""6
    cell: 20.0
    H  0.3 -0.3 -0.8
    O  0.0  0.0  0.0
    H -0.7 -0.7  0.3
    H  2.3  1.7 -0.8
    O  2.0  2.0  0.0
    H  1.3  1.3  0.3
    """)
    for (i, j) in ((0, 1), (1, 2), (3, 4), (4, 5)):
        assert system.add_bond(i, j) == []
    assert len(system.molecules()) == 2
    system.charges[:] = [-0.8476 if name == "O" else 0.4238 for name in system.names]
    system.invalidate()
    return system


# ewald.rs:1329-1333
EWALD_CACHE_MOVE = np.array([
    [0.41727, 2.29401, -0.0558],
    [0.5097743599026461, 3.194114034722624, -0.020364564697826326],
    [-0.2501317777731211, 3.562366060753896, -0.6178033542374419],
])
