"""The oracle's restatement of the Monte Carlo energy cache (sys/cache.rs, GlobalCache of Ewald and Wolf) against
the reference's own tests of it: every one asserts ``cost == E(after) - E(before)`` on a fixed system and fixed
trial positions, with the tolerance quoted next to each assertion below."""

import copy
import math

import numpy as np

import lumol_b200 as lumol
from oracle import oracle
import systems
from test_oracle_kat import ulps_eq


def moved(system, molecule, positions):
    after = copy.deepcopy(system)
    after._device = None
    bonding = after.molecule(molecule)
    after.positions[bonding.start:bonding.end] = positions
    return after


def relative_eq(a, b, max_relative):
    """approx::assert_relative_eq with its default epsilon."""
    return a == b or abs(a - b) <= np.finfo(float).eps or abs(a - b) <= max_relative * max(abs(a), abs(b))


def test_cache_move_molecule():
    # cache.rs:404-446: two successive moves of molecule 0, max_relative = 1e-9
    system = systems.cache_testing_system()
    for positions in systems.CACHE_MOVES:
        before = oracle.OracleSystem(system)
        old_energy = before.potential_energy()
        cost = before.move_molecule_cost(0, positions).sum()
        system = moved(system, 0, positions)
        new_energy = oracle.OracleSystem(system).potential_energy()
        assert relative_eq(cost, new_energy - old_energy, 1e-9)
        assert cost != 0.0


def test_cache_move_all_molecules():
    # cache.rs:449-482: assert_ulps_eq!(cost, new - old, epsilon = 1e-12)
    system = systems.cache_testing_system()
    first = copy.deepcopy(system)
    first._device = None
    bonding = first.molecule(0)
    first.positions[bonding.start:bonding.end] += systems.CACHE_TRANSLATIONS[0]
    before, after = oracle.OracleSystem(system), oracle.OracleSystem(first)
    cost = before.move_all_molecules_cost(after).sum()
    assert abs(cost - (after.potential_energy() - before.potential_energy())) <= 1e-12
    assert cost != 0.0

    # the cache now describes `first`; move every molecule of the original system
    second = copy.deepcopy(system)
    second._device = None
    second.positions += systems.CACHE_TRANSLATIONS[1]
    cached, after = oracle.OracleSystem(first), oracle.OracleSystem(second)
    cost = cached.move_all_molecules_cost(after).sum()
    assert abs(cost - (after.potential_energy() - cached.potential_energy())) <= 1e-12


def test_wolf_move_rigid_molecule():
    # wolf.rs:476-497: assert_ulps_eq!(cost, new_energy - old_energy)
    system = systems.wolf_cache_system()
    wolf = lumol.Wolf(8.0)
    wolf.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(wolf)
    before = oracle.OracleSystem(system)
    terms = before.move_molecule_cost(0, systems.WOLF_CACHE_MOVE)
    assert terms[0] == 0.0 and terms[2] == 0.0
    after = oracle.OracleSystem(moved(system, 0, systems.WOLF_CACHE_MOVE))
    difference = after.energy_terms().coulomb_real - before.energy_terms().coulomb_real
    assert ulps_eq(terms[1], difference) or abs(terms[1] - difference) < 4e-16 * abs(before.energy_terms().coulomb_real)


def test_ewald_move_molecule():
    # ewald.rs:1289-1377: real space (Ewald::new(8, 10)), k-space (Ewald::new(2, 10)), everything; max_relative = 1e-12
    for cutoff, pick in ((8.0, "real"), (2.0, "kspace"), (8.0, "all")):
        system = systems.ewald_cache_system()
        ewald = lumol.SharedEwald(lumol.Ewald(cutoff, 10))
        ewald.set_restriction(lumol.PairRestriction.InterMolecular)
        system.set_coulomb_potential(ewald)
        before = oracle.OracleSystem(system)
        after = oracle.OracleSystem(moved(system, 0, systems.EWALD_CACHE_MOVE))
        old, new = before.energy_terms(), after.energy_terms()
        terms = before.move_molecule_cost(0, systems.EWALD_CACHE_MOVE)
        if pick == "real":
            assert relative_eq(terms[1], new.coulomb_real - old.coulomb_real, 1e-12)
        elif pick == "kspace":
            assert relative_eq(terms[2], new.coulomb_kspace - old.coulomb_kspace, 1e-12)
        else:
            total_old = old.coulomb_real + old.coulomb_self + old.coulomb_kspace
            total_new = new.coulomb_real + new.coulomb_self + new.coulomb_kspace
            assert relative_eq(terms[1] + terms[2], total_new - total_old, 1e-12)
        assert terms[1] != 0.0 and terms[2] != 0.0


def test_ewald_delta_rho_is_the_change_of_the_structure_factor():
    # ewald.rs:833-837: the updater adds delta_rho to rho; rho + delta must be the structure factor of the moved system
    system = systems.ewald_cache_system()
    ewald = lumol.SharedEwald(lumol.Ewald(8.0, 10))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    before = oracle.OracleSystem(system)
    nk = len(before.ewald_factors()[2])
    _, delta = before.ewald_kspace_move_molecule_cost(0, systems.EWALD_CACHE_MOVE, nk)
    after = oracle.OracleSystem(moved(system, 0, systems.EWALD_CACHE_MOVE))
    np.testing.assert_allclose(before.ewald_rho(nk) + delta, after.ewald_rho(nk), rtol=0, atol=1e-14)
    assert math.isfinite(delta.sum())
