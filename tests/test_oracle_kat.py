"""Pins the CPU oracle against the known answers of the reference's own tests.

The reference cannot be built here (no Rust toolchain), so the oracle is a C restatement; this file is
what makes it trustworthy: every exact-equality literal, NIST value and LAMMPS force the reference asserts
for the hot path is re-asserted against the oracle, with the reference file:line next to each.
"""

import ctypes
import math

import numpy as np
import pytest

import lumol_b200 as lumol
from lumol_b200 import units
from lumol_b200.consts import K_BOLTZMANN
from oracle import oracle
from reference_values import (
    EWALD_NIST_ENERGY, EWALD_NIST_VIRIAL, EWALD_WITH_ACCURACY, LAMMPS_FORCES, NIST_LJ, NIST_SPCE, round_at,
)
import systems

lib = oracle.library()


def pot(kind, *params):
    record = oracle.OrcPotential()
    record.pot = kind
    for k, value in enumerate(params):
        record.p[k] = value
    return record


def energy(p, r):
    return lib.orc_potential_energy(ctypes.byref(p), r)


def force(p, r):
    return lib.orc_potential_force(ctypes.byref(p), r)


def tail_energy(p, rc):
    return lib.orc_potential_tail_energy(ctypes.byref(p), rc)


def tail_virial(p, rc):
    return lib.orc_potential_tail_virial(ctypes.byref(p), rc)


def ulps_eq(a, b, max_ulps=4):
    """approx::assert_ulps_eq with its defaults: |a - b| <= f64::EPSILON or within 4 ulps."""
    if a == b or abs(a - b) <= np.finfo(np.float64).eps:
        return True
    ia = np.float64(a).view(np.int64)
    ib = np.float64(b).view(np.int64)
    return (a < 0) == (b < 0) and abs(int(ia) - int(ib)) <= max_ulps


# ---- potentials: lumol-core/src/energy/functions.rs tests -------------------------------------------------

def test_lennard_jones_known_answers():
    lj = pot(oracle.POT_LJ, 2.0, 0.8)  # functions.rs:621-643
    assert energy(lj, 2.0) == 0.0
    assert energy(lj, 2.5) == -0.6189584744448002
    assert tail_energy(lj, 1.0) == 1388.0888888888887
    assert tail_energy(lj, 2.0) == -5.688888888888889
    assert tail_energy(lj, 14.42) == -0.022767318648783084
    assert tail_virial(lj, 1.0) == 17066.666666666668
    assert tail_virial(lj, 2.0) == -17.06666666666667
    assert tail_virial(lj, 14.42) == -0.1366035877536718
    assert abs(force(lj, 2.0 ** (1.0 / 6.0) * 2.0)) < 1e-15
    assert ulps_eq(force(lj, 2.5), -0.95773475733504)
    eps = 1e-9
    assert math.isclose((energy(lj, 4.0) - energy(lj, 4.0 + eps)) / eps, force(lj, 4.0), rel_tol=1e-6, abs_tol=1e-6)


def test_harmonic_known_answers():
    harmonic = pot(oracle.POT_HARMONIC, 50.0, 2.0)  # functions.rs:646-660
    assert energy(harmonic, 2.0) == 0.0
    assert energy(harmonic, 2.5) == 6.25
    assert force(harmonic, 2.0) == 0.0
    assert force(harmonic, 2.5) == -25.0
    assert tail_energy(harmonic, 1.0) == 0.0
    assert tail_virial(harmonic, 1.0) == 0.0


def test_cosine_harmonic_and_torsion():
    ch = pot(oracle.POT_COSINE_HARMONIC, 50.0, math.cos(2.0))  # functions.rs:663-676
    assert energy(ch, 2.0) == 0.0
    dcos = math.cos(2.5) - math.cos(2.0)
    assert energy(ch, 2.5) == 0.5 * 50.0 * dcos * dcos
    assert force(ch, 2.0) == 0.0
    assert force(ch, 2.5) == 50.0 * dcos * math.sin(2.5)
    torsion = pot(oracle.POT_TORSION, 5.0, 3.0, 3.0)  # functions.rs:679-694
    assert energy(torsion, 1.0) == 10.0
    assert energy(torsion, 1.1) == 5.0 * (1.0 + math.cos(3.0 * 1.1 - 3.0))
    assert force(torsion, 1.0) == 0.0


def test_buckingham_known_answers():
    buckingham = pot(oracle.POT_BUCKINGHAM, 2.0, 1.0, 2.0)  # functions.rs:697-715
    assert energy(buckingham, 2.0) == 0.7201338823428847
    assert force(buckingham, 2.0) == 0.32100444117144233
    assert tail_energy(buckingham, 10.0) == 1.8323882504179136
    assert tail_virial(buckingham, 10.0) == 33.422487868546725


def test_born_mayer_huggins_known_answers():
    born = pot(oracle.POT_BMH, 2.0, 1.0, 0.5, 2.0, 2.0)  # functions.rs:718-737
    assert energy(born, 2.0) == 1.986328125
    assert force(born, 2.0) == 0.9609375
    assert tail_energy(born, 10.0) == 4.981521444402363
    assert tail_virial(born, 10.0) == 69.13986044386026


def test_morse_known_answers():
    morse = pot(oracle.POT_MORSE, 2.0, 1.3, 4.0)  # functions.rs:740-757
    assert energy(morse, 1.0) == 2.703517287822119
    assert force(morse, 1.0) == -37.12187076378477
    assert tail_energy(morse, 1.0) == 0.0
    assert tail_virial(morse, 1.0) == 0.0


def test_gaussian_known_answers():
    gaussian = pot(oracle.POT_GAUSSIAN, 8.0, 2.0)  # functions.rs:760-771
    assert energy(gaussian, 0.0) == -8.0
    assert force(gaussian, 0.0) == 0.0
    assert abs(tail_energy(gaussian, 2.5) - -1.93518e-5) < 1e-10
    assert abs(tail_virial(gaussian, 2.5) - 5.23887e-4) < 1e-10


def test_mie_known_answers():
    prefac = lib.orc_mie_prefactor(0.8, 12.0, 6.0)
    mie = pot(oracle.POT_MIE, 2.0, 12.0, 6.0, prefac)  # functions.rs:781-801
    assert energy(mie, 2.0) == 0.0
    assert energy(mie, 2.5) == -0.6189584744448002
    assert math.isclose(tail_energy(mie, 1.0), 1388.0888889, rel_tol=1e-6)
    assert math.isclose(tail_energy(mie, 14.42), -0.022767318648783084, rel_tol=1e-6)
    assert math.isclose(tail_virial(mie, 2.0), -17.06666666666667, rel_tol=1e-6)
    assert abs(force(mie, 2.0 ** (1.0 / 6.0) * 2.0)) < 1e-15
    assert ulps_eq(force(mie, 2.5), -0.95773475733504)
    diverging = pot(oracle.POT_MIE, 2.0, 12.0, 2.0, lib.orc_mie_prefactor(0.8, 12.0, 2.0))  # functions.rs:810-814
    assert tail_energy(diverging, 2.0) == 0.0
    assert tail_virial(diverging, 2.0) == 0.0


# ---- PairInteraction and TableComputation -----------------------------------------------------------------

def pair_record(potential, cutoff, shifted=False, tail=False):
    record = oracle.OrcPair()
    record.potential = potential
    record.cutoff = cutoff
    record.shifted = int(shifted)
    record.tail = int(tail)
    return record


def test_pair_interaction_known_answers():
    harmonic = pot(oracle.POT_HARMONIC, 4.2, 0.5)  # pairs.rs:46-56, 72-85
    plain = pair_record(harmonic, 2.0)
    assert lib.orc_pair_energy(ctypes.byref(plain), 1.0) == 0.525
    assert lib.orc_pair_energy(ctypes.byref(plain), 2.0) == 0.0
    assert lib.orc_pair_force(ctypes.byref(plain), 1.0) == -2.1
    assert lib.orc_pair_force(ctypes.byref(plain), 2.0) == 0.0
    shifted = pair_record(harmonic, 2.0, shifted=True)
    assert lib.orc_pair_energy(ctypes.byref(shifted), 1.0) == -4.2
    assert abs(lib.orc_pair_energy(ctypes.byref(shifted), 1.999)) < 0.01
    assert lib.orc_pair_energy(ctypes.byref(shifted), 2.0) == 0.0

    lj = pot(oracle.POT_LJ, 1.0, 2.0)  # pairs.rs:318-360
    cut = pair_record(lj, 4.0)
    assert lib.orc_pair_force(ctypes.byref(cut), 2.5) == force(lj, 2.5)
    assert lib.orc_pair_energy(ctypes.byref(cut), 2.5) == energy(lj, 2.5)
    assert lib.orc_pair_force(ctypes.byref(cut), 4.1) == 0.0
    assert lib.orc_pair_energy(ctypes.byref(cut), 4.1) == 0.0
    shifted = pair_record(lj, 4.0, shifted=True, tail=True)
    assert ulps_eq(lib.orc_pair_energy(ctypes.byref(shifted), 2.5), -0.030681134109158216)
    assert lib.orc_pair_tail_energy(ctypes.byref(shifted)) == -0.041663275824652776
    assert ulps_eq(3.0 * (lib.orc_pair_tail_virial(ctypes.byref(shifted)) * (1.0 / 3.0)), -0.24995930989583334)

    small = pair_record(pot(oracle.POT_LJ, 0.5, 4.2), 2.0, tail=True)  # pairs.rs:250-258, 267-288
    assert lib.orc_pair_tail_energy(ctypes.byref(small)) == -0.010936609903971353
    assert lib.orc_pair_tail_virial(ctypes.byref(small)) * (1.0 / 3.0) == -0.02187143961588542


def test_table_computation_known_answers():
    harmonic = pot(oracle.POT_HARMONIC, 50.0, 2.0)  # computations.rs:184-211
    size, maximum = 1000, 4.0
    e = np.zeros(size)
    f = np.zeros(size)
    lib.orc_table_build(ctypes.byref(harmonic), size, maximum, oracle.dptr(e), oracle.dptr(f))

    def table_e(r):
        return lib.orc_table_energy(oracle.dptr(e), size, maximum, r)

    def table_f(r):
        return lib.orc_table_energy(oracle.dptr(f), size, maximum, r)

    assert table_e(2.5) == 6.25
    assert table_f(2.5) == -25.0
    delta = 4.0 / 1000.0
    assert table_e(4.0 - 2.0 * delta) == 99.2016
    assert table_f(4.0 - 2.0 * delta) == -99.6
    for r in (4.0 - delta, 4.0, 4.1):
        assert table_e(r) == 0.0
        assert table_f(r) == 0.0


# ---- cell geometry: lumol-core/src/sys/config/cells.rs tests -------------------------------------------------

def test_cell_geometry():
    def image(matrix, shape, v):
        v = np.array(v, dtype=np.float64)
        lib.orc_vector_image(oracle.dptr(np.ascontiguousarray(matrix)), shape, oracle.dptr(v))
        return v

    def wrap(matrix, shape, v):
        v = np.array(v, dtype=np.float64)
        lib.orc_wrap_vector(oracle.dptr(np.ascontiguousarray(matrix)), shape, oracle.dptr(v))
        return v

    cubic, ortho = np.diag([10.0, 10.0, 10.0]), np.diag([3.0, 4.0, 5.0])
    # cells.rs:652-679 vector_image
    np.testing.assert_array_equal(image(cubic, 1, [9.0, 18.0, -6.0]), [-1.0, -2.0, 4.0])
    np.testing.assert_array_equal(image(ortho, 1, [1.0, 1.5, 6.0]), [1.0, 1.5, 1.0])
    np.testing.assert_array_equal(image(np.zeros((3, 3)), 0, [1.0, 1.5, 6.0]), [1.0, 1.5, 6.0])
    tric90 = lumol.UnitCell.triclinic(3.0, 4.0, 5.0, 90.0, 90.0, 90.0).matrix()
    np.testing.assert_allclose(image(tric90, 2, [1.0, 1.5, 6.0]), [1.0, 1.5, 1.0], rtol=1e-15)
    # cells.rs:619-649 wrap_vector
    np.testing.assert_array_equal(wrap(cubic, 1, [9.0, 18.0, -6.0]), [9.0, 8.0, 4.0])
    np.testing.assert_array_equal(wrap(ortho, 1, [1.0, 1.5, 6.0]), [1.0, 1.5, 1.0])
    np.testing.assert_allclose(wrap(tric90, 2, [1.0, 1.5, 6.0]), [1.0, 1.5, 1.0], rtol=1e-15)
    # cells.rs:500-542 volume and lengths
    assert lib.orc_cell_volume(oracle.dptr(ortho), 1) == 60.0
    tric = np.ascontiguousarray(lumol.UnitCell.triclinic(3.0, 4.0, 5.0, 80.0, 90.0, 110.0).matrix())
    assert abs(lib.orc_cell_volume(oracle.dptr(tric), 2) - 55.410529) < 1e-6
    lengths = np.zeros(3)
    tric = np.ascontiguousarray(lumol.UnitCell.triclinic(3.0, 4.0, 5.0, 90.0, 80.0, 100.0).matrix())
    lib.orc_cell_lengths(oracle.dptr(tric), 2, oracle.dptr(lengths))
    np.testing.assert_array_equal(lengths, [2.908132319388713, 3.9373265973230853, 4.921658246653857])
    lib.orc_cell_lengths(oracle.dptr(ortho), 1, oracle.dptr(lengths))
    np.testing.assert_array_equal(lengths, [3.0, 4.0, 5.0])
    # cells.rs:577-601 k_vector
    two_pi_vol = 2.0 * math.pi / 60.0
    k = np.zeros(3)
    lib.orc_k_vector(oracle.dptr(ortho), oracle.dptr(np.ones(3)), oracle.dptr(k))
    np.testing.assert_array_equal(k, [4.0 * 5.0 * two_pi_vol, 3.0 * 5.0 * two_pi_vol, 3.0 * 4.0 * two_pi_vol])
    # angle and dihedral derivatives against finite differences (cells.rs:693-800 test the same way)
    rng = np.random.Generator(np.random.PCG64(5))
    r = rng.uniform(-1, 1, (4, 3))
    d = [np.zeros(3) for _ in range(4)]
    box = np.diag([10.0, 10.0, 10.0])
    theta = lib.orc_angle_and_derivatives(oracle.dptr(box), 1, oracle.dptr(r[0]), oracle.dptr(r[1]), oracle.dptr(r[2]),
                                          oracle.dptr(d[0]), oracle.dptr(d[1]), oracle.dptr(d[2]))
    eps = 1e-7
    for atom in range(3):
        for axis in range(3):
            moved = r.copy()
            moved[atom, axis] += eps
            t = lib.orc_angle_and_derivatives(oracle.dptr(box), 1, oracle.dptr(moved[0]), oracle.dptr(moved[1]), oracle.dptr(moved[2]),
                                              oracle.dptr(np.zeros(3)), oracle.dptr(np.zeros(3)), oracle.dptr(np.zeros(3)))
            assert abs((t - theta) / eps - d[atom][axis]) < 1e-5
    phi = lib.orc_dihedral_and_derivatives(oracle.dptr(box), 1, *[oracle.dptr(r[k]) for k in range(4)], *[oracle.dptr(d[k]) for k in range(4)])
    for atom in range(4):
        for axis in range(3):
            moved = r.copy()
            moved[atom, axis] += eps
            scratch = [np.zeros(3) for _ in range(4)]
            p = lib.orc_dihedral_and_derivatives(oracle.dptr(box), 1, *[oracle.dptr(moved[k]) for k in range(4)], *[oracle.dptr(s) for s in scratch])
            assert abs((p - phi) / eps - d[atom][axis]) < 1e-5


# ---- estimators: sys/energy.rs and sys/compute.rs tests ---------------------------------------------------------

def molecular_test_system():
    system = lumol.system_from_xyz("""4
    cell: 10.0
    F 0.0 0.0 0.0
    F 1.0 0.0 0.0
    F 1.0 1.0 0.0
    F 2.0 1.0 0.0
    """)
    assert system.add_bond(0, 1) == []
    assert system.add_bond(1, 2) == []
    assert system.add_bond(2, 3) == []
    assert len(system.molecules()) == 1
    system.set_bond_potential(("F", "F"), lumol.Harmonic(k=units.from_(100.0, "kJ/mol/A^2"), x0=units.from_(2.0, "A")))
    system.set_angle_potential(("F", "F", "F"), lumol.Harmonic(k=units.from_(100.0, "kJ/mol/deg^2"), x0=units.from_(88.0, "deg")))
    system.set_dihedral_potential(("F", "F", "F", "F"), lumol.Harmonic(k=units.from_(100.0, "kJ/mol/deg^2"), x0=units.from_(185.0, "deg")))
    system.set_pair_potential(("H", "O"), lumol.PairInteraction(lumol.NullPotential(), 0.0))
    return system


def test_energy_evaluator_known_answers():
    system = molecular_test_system()  # energy.rs:179-263
    pair = lumol.PairInteraction(lumol.LennardJones(epsilon=units.from_(100.0, "kJ/mol/A^2"), sigma=units.from_(0.8, "A")), 5.0)
    pair.enable_tail_corrections()
    system.set_pair_potential(("F", "F"), pair)
    terms = oracle.OracleSystem(system).energy_terms()
    assert ulps_eq(terms.pairs, units.from_(-258.3019360389957, "kJ/mol"))
    assert ulps_eq(terms.pairs_tail, -0.0000028110338032153973)
    assert ulps_eq(terms.bonds, units.from_(150.0, "kJ/mol"))
    assert ulps_eq(terms.angles, units.from_(400.0, "kJ/mol"))
    assert ulps_eq(terms.dihedrals, units.from_(1250.0, "kJ/mol"), max_ulps=15)


def pairs_test_system():
    system = lumol.system_from_xyz("""2
    cell: 10.0
    F 0.0 0.0 0.0 -0.007225222699367925 -0.002405756495275919  0.0026065109398392215
    F 1.3 0.0 0.0  0.001179633958023287  0.003525262341736351 -0.0004132774783154952
    """)
    interaction = lumol.PairInteraction(lumol.Harmonic(k=units.from_(300.0, "kJ/mol/A^2"), x0=units.from_(1.2, "A")), 5.0)
    interaction.enable_tail_corrections()
    system.set_pair_potential(("F", "F"), interaction)
    system.set_pair_potential(("H", "O"), lumol.PairInteraction(lumol.NullPotential(), 0.0))
    return system


def test_compute_known_answers():
    orc = oracle.OracleSystem(pairs_test_system())  # compute.rs:493-786
    forces = orc.forces()
    np.testing.assert_array_equal(forces[0] + forces[1], 0.0)
    force = units.from_(30.0, "kJ/mol/A")
    assert ulps_eq(forces[0][0], force) and ulps_eq(forces[1][0], -force)
    assert ulps_eq(orc.kinetic_energy(), 0.0007483016557453698)
    assert ulps_eq(orc.temperature(), 300.0)
    virial = orc.atomic_virial()
    assert ulps_eq(virial[0][0], -force * 1.3)
    assert np.count_nonzero(virial) == 1
    expected = 2.0 * K_BOLTZMANN * 300.0 / 1000.0 + (-force * 1.3) / (3.0 * 1000.0)
    assert ulps_eq(orc.pressure(), expected)
    assert ulps_eq(np.trace(orc.stress()) / 3.0, orc.pressure())

    molecular = oracle.OracleSystem(molecular_test_system())  # compute.rs:609-613, 650-662
    assert ulps_eq(molecular.potential_energy(), units.from_(1800.0, "kJ/mol"))
    total = molecular.forces().sum(axis=0)
    assert float(total @ total) < 1e-30
    w = units.from_(100.0, "kJ/mol/A")
    virial = molecular.atomic_virial()
    assert ulps_eq(virial[0][0], 2.0 * w) and ulps_eq(virial[1][1], 1.0 * w)


# ---- Wolf and Ewald ------------------------------------------------------------------------------------------

def nacl_pair(cell=20.0):
    system = lumol.system_from_xyz(f"""2
    cell: {cell}
    Cl 0.0 0.0 0.0
    Na 1.5 0.0 0.0
    """)
    system.charges[0] = -1.0
    system.charges[1] = 1.0
    return system


def single_water():
    system = lumol.system_from_xyz("""3
    cell: 20.0
    O  0.0  0.0  0.0
    H -0.7 -0.7  0.3
    H  0.3 -0.3 -0.8
    """)
    system.add_bond(0, 1)
    system.add_bond(0, 2)
    system.charges[:] = [-0.8476, 0.4238, 0.4238]
    return system


def test_wolf_known_answers():
    # doc-test wolf.rs:29-48: exact equality through erfc
    system = lumol.System(lumol.UnitCell.cubic(30.0))
    na = lumol.Particle("Na", (0.0, 0.0, 0.0))
    na.charge = 1.0
    cl = lumol.Particle("Cl", (2.0, 0.0, 0.0))
    cl.charge = -1.0
    system.add_molecule(lumol.Molecule(na))
    system.add_molecule(lumol.Molecule(cl))
    system.set_coulomb_potential(lumol.Wolf(12.0))
    assert oracle.OracleSystem(system).potential_energy() == -0.0729290269539354

    pair = nacl_pair()  # wolf.rs:355-401
    pair.set_coulomb_potential(lumol.Wolf(8.0))
    orc = oracle.OracleSystem(pair)
    assert abs(orc.energy_terms().coulomb_real - -0.09262397663346732) < 1e-2
    forces = orc.coulomb_forces()
    assert np.linalg.norm(forces[0] + forces[1]) <= np.finfo(float).eps
    expected = np.zeros((3, 3))
    expected[0][0] = -forces[0][0] * 1.5
    np.testing.assert_array_equal(orc.coulomb_atomic_virial(), expected)
    eps = 1e-9
    e0 = orc.potential_energy()
    pair.positions[0][0] += eps
    e1 = oracle.OracleSystem(pair).potential_energy()
    assert math.isclose((e0 - e1) / eps, oracle.OracleSystem(pair).coulomb_forces()[0][0], rel_tol=1e-6, abs_tol=1e-6)


def test_ewald_known_answers():
    pair = nacl_pair()  # ewald.rs:1036-1050
    pair.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(8.0, 10)))
    assert abs(oracle.OracleSystem(pair).potential_energy() - -0.09262397663346732) < 1e-4

    water = single_water()  # ewald.rs:1122-1133: assert_ulps_eq on a 3e-4 cancellation of +-0.03 terms
    ewald = lumol.SharedEwald(lumol.Ewald(8.0, 10))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    water.set_coulomb_potential(ewald)
    orc = oracle.OracleSystem(water)
    terms = orc.energy_terms()
    total = terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace
    magnitude = abs(terms.coulomb_real) + abs(terms.coulomb_self) + abs(terms.coulomb_kspace)
    # the reference asserts 4 ulps of the net with rayon's summation order; here: 1e-14 of the term magnitudes
    assert abs(total - -0.000009243868813825495) < 1e-14 * magnitude

    # virial is energy for point charges (ewald.rs:1216-1231; no restriction in that test)
    for system in (single_water(), nacl_pair()):
        system.set_coulomb_potential(lumol.SharedEwald(lumol.Ewald(8.0, 10)))
        o = oracle.OracleSystem(system)
        t = o.energy_terms()
        e = t.coulomb_real + t.coulomb_self + t.coulomb_kspace
        assert math.isclose(e, np.trace(o.coulomb_atomic_virial()), rel_tol=1e-3)

    # total force by finite differences (ewald.rs:1096-1114, 1177-1193)
    eps = 1e-9
    for system in (pair, water):
        o = oracle.OracleSystem(system)
        f = o.coulomb_forces()[0][0]
        e0 = o.potential_energy()
        system.positions[0][0] += eps
        e1 = oracle.OracleSystem(system).potential_energy()
        system.positions[0][0] -= eps
        assert math.isclose((e0 - e1) / eps, f, rel_tol=1e-5, abs_tol=1e-6)


def test_ewald_with_accuracy():
    alpha, kmax = ctypes.c_double(), ctypes.c_int32()
    orc = oracle.OracleSystem(single_water())
    lib.orc_ewald_with_accuracy(orc.ref, 8.5, 1e-6, ctypes.byref(alpha), ctypes.byref(kmax))  # ewald.rs:997-1001
    assert abs(alpha.value - 0.2998) < 1e-4 and kmax.value == 5
    for index, (expected_alpha, expected_kmax) in EWALD_WITH_ACCURACY.items():  # ewald.rs:1446-1462
        system = systems.nist_spce(index)
        orc = oracle.OracleSystem(system)
        lib.orc_ewald_with_accuracy(orc.ref, 9.0, 1e-5, ctypes.byref(alpha), ctypes.byref(kmax))
        assert abs(alpha.value - expected_alpha) < 1e-4 and kmax.value == expected_kmax
        host = lumol.Ewald.with_accuracy(9.0, 1e-5, system)
        assert host.kmax == expected_kmax and abs(host.alpha - alpha.value) < 1e-15


# ---- NIST reference calculations ----------------------------------------------------------------------------------

@pytest.mark.parametrize("index,cutoff", sorted(NIST_LJ))
def test_nist_lennard_jones(index, cutoff):
    (energy_ref, e_dec), (virial_ref, v_dec), (tail_ref, t_dec) = NIST_LJ[(index, cutoff)]
    plain = oracle.OracleSystem(systems.nist_lj(index, cutoff, tail=False))
    e = plain.potential_energy()
    assert round_at(e, e_dec) == energy_ref
    assert round_at(np.trace(plain.atomic_virial()), v_dec) == virial_ref
    with_tail = oracle.OracleSystem(systems.nist_lj(index, cutoff, tail=True))
    assert round_at(with_tail.potential_energy() - e, t_dec) == tail_ref


@pytest.mark.parametrize("index,cutoff", sorted(NIST_SPCE))
def test_nist_spce_energies(index, cutoff):
    total, pairs, tail, coulomb = NIST_SPCE[(index, cutoff)]
    system = systems.nist_spce(index)
    systems.set_nist_interactions(system, cutoff)
    terms = oracle.OracleSystem(system).energy_terms()
    e_coulomb = terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace
    e_total = terms.pairs + terms.pairs_tail + e_coulomb
    assert abs((e_total / K_BOLTZMANN - total) / total) < 1e-3
    assert abs((terms.pairs / K_BOLTZMANN - pairs) / pairs) < 1e-3
    assert abs((terms.pairs_tail / K_BOLTZMANN - tail) / tail) < 1e-3
    assert abs((e_coulomb / K_BOLTZMANN - coulomb) / coulomb) < 1e-3
    (real, real_tol), (kspace, k_tol), (self_, self_tol) = EWALD_NIST_ENERGY[(index, cutoff)]
    assert abs(terms.coulomb_real / K_BOLTZMANN - real) <= real_tol * abs(real)
    assert abs(terms.coulomb_kspace / K_BOLTZMANN - kspace) <= k_tol * abs(kspace)
    assert abs(terms.coulomb_self / K_BOLTZMANN - self_) <= self_tol * abs(self_)


@pytest.mark.parametrize("index,cutoff", [(1, 9.0), (2, 10.0), (3, 9.0)])
def test_nist_spce_virials(index, cutoff):
    kmax, alpha, real, real_tol, kspace, k_tol = EWALD_NIST_VIRIAL[(index, cutoff)]
    system = systems.nist_spce(index)
    ewald = lumol.SharedEwald(lumol.Ewald(cutoff, kmax, alpha))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    orc = oracle.OracleSystem(system)
    convert = units.from_(1.0, "atm") * system.volume()
    w_real = orc._matrix(lib.orc_ewald_real_atomic_virial) / convert
    w_k = orc._matrix(lib.orc_ewald_kspace_atomic_virial) / convert
    np.testing.assert_allclose(w_real, np.array(real), rtol=real_tol)
    np.testing.assert_allclose(w_k, np.array(kspace), rtol=k_tol)


@pytest.mark.parametrize("index,cutoff", [(1, 9), (2, 10), (3, 10)])
def test_lammps_forces(index, cutoff, golden):
    kmax, alpha, dynamic = LAMMPS_FORCES[(index, cutoff)]
    system = systems.nist_spce(index)
    systems.set_lammps_interactions(system, float(cutoff), kmax, alpha)
    forces = oracle.OracleSystem(system).forces() / units.from_(1.0, "kcal/mol/A")
    expected = golden[f"lammps-forces-{cutoff}-{index}/forces"]
    relative = np.abs((forces - expected) / expected)
    if dynamic:
        tolerance = np.where(np.abs(expected) < 1e-1, 1e-1, np.where(np.abs(expected) < 1.0, 5e-2, 1e-2))
    else:
        tolerance = 5e-3
    assert np.all(relative < tolerance)


# ---- topology -------------------------------------------------------------------------------------------------------

def test_bonding_matches_host_topology():
    """The oracle rebuilds angles/dihedrals/bond distances itself; the product's host topology must agree."""
    for system in (systems.propane(), systems.water(), molecular_test_system()):
        orc = oracle.OracleSystem(system)
        angles = {tuple(a) for a in orc.angles.tolist()}
        dihedrals = {tuple(d) for d in orc.dihedrals.tolist()}
        host_angles, host_dihedrals = set(), set()
        for bonding in system.bondings:
            host_angles |= bonding.angles
            host_dihedrals |= bonding.dihedrals
        assert angles == host_angles
        assert dihedrals == host_dihedrals
        n = system.size()
        rng = np.random.Generator(np.random.PCG64(3))
        for _ in range(300):
            i, j = int(rng.integers(n)), int(rng.integers(n))
            path = lib.orc_bond_path(orc.ref, i, j)
            host = system.bond_path(i, j)
            assert path == {-1: 0, 0: 1, 1: 2, 2: 3, 3: 4, 4: 5}[host]


def test_restriction_information():
    # restrictions.rs:85-114 and its tests (:140-296): pentane, paths 1..4 bonds
    excluded, scaling = ctypes.c_int32(), ctypes.c_double()
    table = {
        0: [0, 0, 0, 0, 0, 0],  # None
        1: [1, 0, 0, 0, 0, 0],  # IntraMolecular excludes pairs in different molecules
        2: [0, 1, 1, 1, 1, 1],  # InterMolecular
        3: [0, 0, 1, 0, 0, 0],  # Exclude12
        4: [0, 0, 1, 1, 0, 0],  # Exclude13
        5: [0, 0, 1, 1, 1, 0],  # Exclude14
        6: [0, 0, 1, 1, 0, 0],  # Scale14
    }
    for restriction, row in table.items():
        for path, expected in enumerate(row):
            lib.orc_restriction_information(restriction, 0.8, path, ctypes.byref(excluded), ctypes.byref(scaling))
            assert excluded.value == expected
            assert scaling.value == (0.8 if restriction == 6 and path == 4 else 1.0)


# ---- controls and barostats (lumol-sim/src/md/controls.rs, integrators.rs:176-342) ------------------------------------

def two_silver_atoms(second):
    system = lumol.System(lumol.UnitCell.cubic(10.0))
    system.add_particles(["Ag", "Ag"], np.array([[0.0, 0.0, 0.0], second]), masses=np.array([107.8682, 107.8682]))
    return system


def test_remove_rotation_known_answer():
    # controls.rs:108-120
    system = two_silver_atoms([1.0, 0.0, 0.0])
    velocity = np.array([[0.0, 1.0, 0.0], [0.0, -1.0, 2.0]])
    oracle.library().orc_remove_rotation(2, oracle.dptr(np.ascontiguousarray(system.masses)), oracle.dptr(np.ascontiguousarray(system.positions)),
                                         oracle.dptr(velocity))
    np.testing.assert_array_equal(velocity, [[0.0, 0.0, 1.0], [0.0, 0.0, 1.0]])


def test_rewrap_known_answers():
    # controls.rs:122-131
    system = two_silver_atoms([15.0, 0.0, 0.0])
    reference = oracle.OracleSystem(system)
    position = system.positions.copy()
    reference.lib.orc_rewrap(reference.ref, oracle.dptr(position))
    np.testing.assert_array_equal(position, [[0.0, 0.0, 0.0], [5.0, 0.0, 0.0]])
    # molecules.rs:334-345: a two-atom molecule moves as a whole
    molecule = lumol.System(lumol.UnitCell.cubic(5.0))
    molecule.add_particles(["O", "O"], np.array([[-2.0, 0.0, 0.0], [0.0, 0.0, 0.0]]), masses=np.array([15.999, 15.999]))
    molecule.add_bonds(np.array([[0, 1]]))
    reference = oracle.OracleSystem(molecule)
    position = molecule.positions.copy()
    reference.lib.orc_rewrap(reference.ref, oracle.dptr(position))
    np.testing.assert_array_equal(position, [[3.0, 0.0, 0.0], [5.0, 0.0, 0.0]])


def test_berendsen_barostats_reduce_to_velocity_verlet_and_scale_the_cell():
    """With eta = 1 and the target equal to the instantaneous pressure the barostat step IS a velocity-Verlet step
    (integrators.rs:211-255 against :44-69); a lower target pressure makes the next eta larger than one, and the cell
    is scaled by eta^3 (sic, integrators.rs:225) on the following step."""
    system = systems.md_helium()
    systems.random_velocities(system, 300.0, seed=3)
    n = system.size()
    plain = oracle.OracleSystem(system)
    acc = np.zeros((n, 3))
    plain.lib.orc_velocity_verlet_step(plain.ref, oracle.dptr(plain.position), oracle.dptr(plain.velocity), oracle.dptr(acc), 1.0)

    npt = oracle.OracleSystem(system)
    acc2 = np.zeros((n, 3))
    eta = ctypes.c_double(1.0)
    # pressure after the first half of the step is what the barostat compares with: take it from the plain run
    target = plain.pressure()
    status = npt.lib.orc_berendsen_barostat_step(npt.ref, oracle.dptr(npt.position), oracle.dptr(npt.velocity), oracle.dptr(acc2), 1.0,
                                                 target, 1000.0, ctypes.byref(eta), 0.0)
    assert status == 0
    np.testing.assert_array_equal(npt.position, plain.position)
    np.testing.assert_allclose(npt.velocity, plain.velocity, rtol=0, atol=1e-18)
    before = np.array(npt.s.cell[:])
    np.testing.assert_array_equal(before, system.cell.matrix().reshape(-1))
    eta_after_first = eta.value  # close to one: the target was the pressure at the end of the plain step
    assert abs(eta_after_first - 1.0) < 1e-4
    # second step with a much lower target: the box must grow
    status = npt.lib.orc_berendsen_barostat_step(npt.ref, oracle.dptr(npt.position), oracle.dptr(npt.velocity), oracle.dptr(acc2), 1.0,
                                                 target - 1e-6, 1000.0, ctypes.byref(eta), 0.0)
    assert status == 0 and eta.value > 1.0
    first_eta = eta.value
    status = npt.lib.orc_berendsen_barostat_step(npt.ref, oracle.dptr(npt.position), oracle.dptr(npt.velocity), oracle.dptr(acc2), 1.0,
                                                 target - 1e-6, 1000.0, ctypes.byref(eta), 0.0)
    after = np.array(npt.s.cell[:])
    np.testing.assert_allclose(after[0], before[0] * eta_after_first ** 3 * first_eta ** 3, rtol=1e-14)
    # the anisotropic barostat with a hydrostatic target and eta = 1 is velocity Verlet as well
    aniso = oracle.OracleSystem(system)
    acc3 = np.zeros((n, 3))
    eta9 = np.eye(3).reshape(-1).copy()
    stress = np.ascontiguousarray(plain.stress().reshape(-1))
    status = aniso.lib.orc_aniso_berendsen_barostat_step(aniso.ref, oracle.dptr(aniso.position), oracle.dptr(aniso.velocity), oracle.dptr(acc3),
                                                         1.0, oracle.dptr(stress), 1000.0, oracle.dptr(eta9), 0.0)
    assert status == 0
    np.testing.assert_array_equal(aniso.position, plain.position)
    np.testing.assert_allclose(eta9.reshape(3, 3), eta9.reshape(3, 3).T, rtol=0, atol=0)
    # the cell shrinking below twice the cut-off is the reference's panic (integrators.rs:227-236)
    small = oracle.OracleSystem(system)
    eta = ctypes.c_double(0.5)
    status = small.lib.orc_berendsen_barostat_step(small.ref, oracle.dptr(small.position), oracle.dptr(small.velocity), oracle.dptr(np.zeros((n, 3))),
                                                   1.0, target, 1000.0, ctypes.byref(eta), 12.0)
    assert status == 1


def test_forces_on_selected_rows_equal_the_full_loops():
    """``orc_pair_forces_rows`` / ``orc_ewald_real_forces_rows`` (total force on chosen atoms over every other atom, used to
    pin the 1M-atom bench box) against the full i < j loops of compute.rs:37-55 and ewald.rs:461-500."""
    import lumol_b200 as lumol

    system = systems.lj_box(9, seed=5)
    reference = oracle.OracleSystem(system)
    rows = np.arange(0, system.size(), 5)
    full = reference.pair_forces()
    assert np.abs(reference.pair_forces_rows(rows) - full[rows]).max() <= 1e-13 * np.abs(full).max()

    water = systems.nist_spce(1)
    systems.set_nist_interactions(water, 9.0)
    reference = oracle.OracleSystem(water)
    rows = np.arange(1, water.size(), 7)
    full = reference.pair_forces()
    assert np.abs(reference.pair_forces_rows(rows) - full[rows]).max() <= 1e-13 * np.abs(full).max()
    real = np.zeros((water.size(), 3))
    reference.lib.orc_ewald_real_forces(reference.ref, oracle.dptr(real))
    assert np.abs(reference.ewald_real_forces_rows(rows) - real[rows]).max() <= 1e-13 * np.abs(real).max()
