"""Topology behind ``PairRestriction``: the host mirror of ``Bonding`` / ``Configuration`` and
the oracle's own ``orc_bonding_rebuild`` against the known answers of the reference's tests
(lumol-core/src/sys/config/bonding.rs:325-438, configuration.rs:755-834)."""

import numpy as np

import lumol_b200 as lumol
from oracle import oracle


def chain(names):
    system = lumol.System(lumol.UnitCell.cubic(30.0))
    for k, name in enumerate(names):
        system.add_molecule(lumol.Molecule(lumol.Particle(name, (1.5 * k, 0.0, 0.0))))
    return system


def normalized(items):
    return {tuple(item) if item[0] < item[-1] else tuple(reversed(item)) for item in items}


def test_ethane_bonding():
    # bonding.rs:325-416: 3 - 0 -- 1 - 6 with 2, 4 on atom 0 and 5, 7 on atom 1
    system = chain(["C", "C", "H", "H", "H", "H", "H", "H"])
    for (i, j) in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 5), (1, 6), (1, 7)):
        system.add_bond(i, j)
    assert len(system.molecules()) == 1
    # add_bond may move particles to keep molecules contiguous: here the order is already contiguous
    bonding = system.molecule(0)
    assert (bonding.start, bonding.end, bonding.size()) == (0, 8, 8)
    assert bonding.bonds == {(0, 1), (0, 2), (0, 3), (0, 4), (1, 5), (1, 6), (1, 7)}
    angles = normalized([(0, 1, 5), (0, 1, 6), (0, 1, 7), (1, 0, 2), (1, 0, 3), (1, 0, 4), (2, 0, 3), (3, 0, 4), (2, 0, 4),
                         (5, 1, 6), (6, 1, 7), (5, 1, 7)])
    dihedrals = normalized([(a, 0, 1, b) for a in (2, 3, 4) for b in (5, 6, 7)])
    assert normalized(bonding.angles) == angles
    assert normalized(bonding.dihedrals) == dihedrals
    for (i, j), expected in (((0, 1), 1), ((1, 0), 1), ((0, 7), 2), ((7, 0), 2), ((3, 5), 3), ((5, 3), 3)):
        assert system.bond_path(i, j) == expected
    # the oracle's own rebuild from the bond list
    orc = oracle.OracleSystem(system)
    assert normalized(orc.angles.tolist()) == angles
    assert normalized(orc.dihedrals.tolist()) == dihedrals
    lib = oracle.library()
    for (i, j), expected in (((0, 1), 2), ((0, 7), 3), ((3, 5), 4), ((0, 0), 1)):
        assert lib.orc_bond_path(orc.ref, i, j) == expected


def test_cyclic_molecule():
    # bonding.rs:419-438: a four-membered ring; 0 and 3 are one bond AND three bonds apart
    system = chain(["C", "C", "C", "C"])
    for (i, j) in ((0, 1), (1, 2), (2, 3), (3, 0)):
        system.add_bond(i, j)
    bonding = system.molecule(0)
    bits = int(bonding.distances[0, 3])
    assert bits & 1 and bits & 4
    assert (0, 3, 2) in normalized(bonding.angles) or (2, 3, 0) in normalized(bonding.angles)
    assert (0, 1, 2) in normalized(bonding.angles)
    assert system.bond_path(0, 3) == 1  # the shortest path wins (configuration.rs:133-143)


def test_bond_path_of_pentane():
    # configuration.rs:755-773
    system = lumol.System(lumol.UnitCell.cubic(30.0))
    pentane = lumol.Molecule(lumol.Particle("CH3", (0.0, 0.0, 0.0)))
    pentane.add_particle_bonded_to(0, lumol.Particle("CH2", (1.5, 0.0, 0.0)))
    pentane.add_particle_bonded_to(1, lumol.Particle("CH2", (3.0, 0.0, 0.0)))
    pentane.add_particle_bonded_to(2, lumol.Particle("CH2", (4.5, 0.0, 0.0)))
    pentane.add_particle_bonded_to(3, lumol.Particle("CH3", (6.0, 0.0, 0.0)))
    system.add_molecule(pentane)
    system.add_molecule(lumol.Molecule(lumol.Particle("Zn", (9.0, 0.0, 0.0))))
    assert [system.bond_path(0, j) for j in range(6)] == [0, 1, 2, 3, 4, -1]
    orc = oracle.OracleSystem(system)
    lib = oracle.library()
    assert [lib.orc_bond_path(orc.ref, 0, j) for j in range(6)] == [1, 2, 3, 4, 5, 0]


def test_add_bond_permutations():
    # configuration.rs:776-805
    system = chain(["C", "H", "H", "H", "C", "H", "H", "H"])
    assert system.add_bond(0, 3) == [(3, 1), (1, 2), (2, 3)]
    assert system.add_bond(0, 3) == [(3, 2), (2, 3)]
    assert system.add_bond(0, 3) == []
    assert system.add_bond(4, 5) == []
    assert system.add_bond(4, 7) == [(7, 6), (6, 7)]
    assert system.add_bond(4, 7) == []
    # regression test of the reference's issue #76
    system = chain(["H", "H", "O"])
    assert system.add_bond(0, 2) == [(2, 1), (1, 2)]
    assert system.add_bond(2, 1) == []
    assert len(system.molecules()) == 1


def test_distances_with_and_without_a_cell():
    # configuration.rs:821-834 through the oracle's minimum image
    lib = oracle.library()
    cell = np.ascontiguousarray(np.diag([5.0, 5.0, 5.0]).reshape(-1))
    d = np.array([9.0, 0.0, 0.0])
    lib.orc_vector_image(oracle.dptr(cell), oracle.ORC_CELL_ORTHO if hasattr(oracle, "ORC_CELL_ORTHO") else 1, oracle.dptr(d))
    assert np.linalg.norm(d) == 1.0
    d = np.array([9.0, 0.0, 0.0])
    lib.orc_vector_image(oracle.dptr(cell), 0, oracle.dptr(d))
    assert np.linalg.norm(d) == 9.0


def test_bulk_added_atoms_bond_like_molecules_added_one_by_one():
    """`System.add_particles` makes one-atom molecules that share an immutable empty connection set (millions of free
    atoms in the bench boxes); bonding them afterwards, one bond at a time or in bulk, gives the molecules of
    `Configuration::add_bond` (configuration.rs:243-303)."""
    import lumol_b200 as lumol

    def fresh():
        system = lumol.System(lumol.UnitCell.cubic(20.0))
        system.add_particles(["O", "H", "H", "Ar"], np.arange(12.0).reshape(4, 3))
        return system

    system = fresh()
    assert all(bonding.size() == 1 and not bonding.bonds for bonding in system.bondings)
    system.add_bond(0, 1)
    system.add_bond(0, 2)
    bulk = fresh()
    bulk.add_bonds([(0, 1), (0, 2)])
    for built in (system, bulk):
        water, argon = built.bondings
        assert (water.start, water.end, argon.start, argon.end) == (0, 3, 3, 4)
        assert sorted(water.bonds) == [(0, 1), (0, 2)] and sorted(water.angles) == [(1, 0, 2)] and not water.dihedrals
        assert not argon.bonds and not argon.angles
    # the shared empty set was not touched
    assert not fresh().bondings[0].bonds
