"""Monte Carlo energy cache on the device (``lumol_cuda_move_molecule*``) against the CPU oracle and against the
reference's own cache tests (sys/cache.rs:397-483, energy/global/ewald.rs:1289-1377, wolf.rs:476-497), ``-m gpu``.

Tolerances: a cost is a difference of two energies, so it is compared relative to the magnitude of the energy
terms it is the difference of (1e-9, north_star's energy tolerance), and the reference's own assertions
(``cost == E(after) - E(before)`` to 1e-9 relative for the cache, 1e-12 for Ewald) are re-asserted on the device
numbers with the same magnitudes.
"""

import ctypes

import numpy as np
import pytest

import lumol_b200 as lumol
from lumol_b200 import _ffi, synthetic
from lumol_b200.device import device_for
from oracle import oracle
import systems

pytestmark = pytest.mark.gpu

ENERGY_TOL = 1e-9


def magnitude(terms):
    names = ("pairs", "pairs_tail", "bonds", "angles", "dihedrals", "coulomb_real", "coulomb_self", "coulomb_kspace")
    return max(sum(abs(getattr(terms, name)) for name in names), 1e-300)


def device_costs(system, molecules, new_positions):
    """(pairs, coulomb real, coulomb k-space) per trial, straight from the C ABI."""
    device = device_for(system)
    ids = np.ascontiguousarray(molecules, dtype=np.int64)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in new_positions]))
    costs = (_ffi.Energy * len(ids))()
    _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecules_cost(
        device.ctx, len(ids), ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ffi.as_double_pointer(flat), costs))
    return np.array([[c.pairs, c.coulomb_real, c.coulomb_kspace] for c in costs])


def moved(system, molecule, positions):
    after = system.clone()
    bonding = after.molecule(molecule)
    after.positions[bonding.start:bonding.end] = positions
    return after


def rigid_trial(system, molecule, rng, delta=1.5):
    """A random rotation about the first atom plus a translation: what Rotate / Translate propose."""
    bonding = system.molecule(molecule)
    positions = system.positions[bonding.start:bonding.end]
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, a, b, c = q
    rot = np.array([
        [1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w)],
        [2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w)],
        [2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)],
    ])
    return positions[0] + (positions - positions[0]) @ rot.T + rng.uniform(-delta, delta, 3)


def check_against_oracle(system, molecules, trials, tol=ENERGY_TOL):
    reference = oracle.OracleSystem(system)
    scale = magnitude(reference.energy_terms())
    costs = device_costs(system, molecules, trials)
    for molecule, positions, cost in zip(molecules, trials, costs):
        expected = reference.move_molecule_cost(int(molecule), positions)
        assert np.abs(cost - expected).max() <= tol * scale, f"molecule {molecule}: {cost!r} vs {expected!r} (scale {scale:.3e})"
    return costs, scale


# ---- the reference's own cache tests, run on the device ---------------------------------------------------

def test_cache_energy():
    system = systems.cache_testing_system()  # cache.rs:397-402
    cache = lumol.EnergyCache()
    cache.init(system)
    energy = system.potential_energy()
    assert abs(cache.energy() - energy) <= 4e-16 * magnitude(device_for(system).compute(energy=True).energy)


def test_cache_move_molecule():
    system = systems.cache_testing_system()  # cache.rs:404-446
    cache = lumol.EnergyCache()
    old_energy = system.potential_energy()
    cache.init(system)
    for positions in systems.CACHE_MOVES:
        _, scale = check_against_oracle(system, [0], [positions])
        cost = cache.move_molecule_cost(system, 0, positions)
        system.positions[0:4] = positions
        cache.update(system)
        # the accepted positions are resident: this evaluation must not need the host copy
        resident = device_for(system, positions=False).compute(energy=True).energy.total()
        new_energy = system.potential_energy()
        assert abs(resident - new_energy) <= 1e-14 * scale
        assert abs(cost - (new_energy - old_energy)) <= 1e-9 * max(abs(cost), abs(new_energy - old_energy)) + 1e-13 * scale
        assert abs(cache.energy() - new_energy) <= 1e-12 * scale
        old_energy = new_energy


def test_cache_move_all_molecules():
    system = systems.cache_testing_system()  # cache.rs:449-482
    cache = lumol.EnergyCache()
    old_energy = system.potential_energy()
    cache.init(system)
    scale = magnitude(device_for(system).compute(energy=True).energy)

    new_system = system.clone()
    new_system.positions[0:4] += systems.CACHE_TRANSLATIONS[0]
    cost = cache.move_all_molecules_cost(new_system)
    new_energy = new_system.potential_energy()
    assert abs(cost - (new_energy - old_energy)) <= 1e-12 * scale
    expected = oracle.OracleSystem(system).move_all_molecules_cost(oracle.OracleSystem(new_system)).sum()
    assert abs(cost - expected) <= ENERGY_TOL * scale
    cache.update(new_system)
    assert abs(cache.energy() - new_energy) <= 1e-12 * scale

    old_energy = new_energy
    new_system = system.clone()
    new_system.positions += systems.CACHE_TRANSLATIONS[1]
    cost = cache.move_all_molecules_cost(new_system)
    assert abs(cost - (new_system.potential_energy() - old_energy)) <= 1e-12 * scale


def test_wolf_move_rigid_molecule():
    system = systems.wolf_cache_system()  # wolf.rs:476-497
    wolf = lumol.Wolf(8.0)
    wolf.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(wolf)
    old_energy = wolf.energy(system)
    costs, scale = check_against_oracle(system, [0], [systems.WOLF_CACHE_MOVE])
    assert costs[0][0] == 0.0 and costs[0][2] == 0.0
    new_energy = wolf.energy(moved(system, 0, systems.WOLF_CACHE_MOVE))
    assert abs(costs[0][1] - (new_energy - old_energy)) <= 1e-12 * scale


@pytest.mark.parametrize("cutoff", [8.0, 2.0])
def test_ewald_move_molecule(cutoff):
    system = systems.ewald_cache_system()  # ewald.rs:1289-1377
    ewald = lumol.SharedEwald(lumol.Ewald(cutoff, 10))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    device = device_for(system)
    old = device.compute(energy=True, parts=_ffi.PART_COULOMB).energy
    costs, scale = check_against_oracle(system, [0], [systems.EWALD_CACHE_MOVE])
    after = moved(system, 0, systems.EWALD_CACHE_MOVE)
    new = device_for(after).compute(energy=True, parts=_ffi.PART_COULOMB).energy
    # the reference asserts max_relative = 1e-12 on each part; here relative to the magnitude of the sums involved
    assert abs(costs[0][1] - (new.coulomb_real - old.coulomb_real)) <= 1e-12 * scale
    assert abs(costs[0][2] - (new.coulomb_kspace - old.coulomb_kspace)) <= 1e-12 * scale
    assert abs(new.coulomb_self - old.coulomb_self) == 0.0  # "No self cost", ewald.rs:942


def test_ewald_accept_updates_positions_and_structure_factor():
    system = systems.ewald_cache_system()  # ewald.rs:833-837
    ewald = lumol.SharedEwald(lumol.Ewald(8.0, 10))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    device = device_for(system)
    device_costs(system, [0], [systems.EWALD_CACHE_MOVE])
    _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecule_accept(device.ctx, 0))

    after = moved(system, 0, systems.EWALD_CACHE_MOVE)
    reference = oracle.OracleSystem(after)
    nk = len(reference.ewald_factors()[2])
    count = ctypes.c_int64()
    rho = np.zeros((nk, 2))
    index = np.zeros((nk, 3), dtype=np.int32)
    factor = np.zeros(nk)
    _ffi.check(device.ctx, device.lib.lumol_cuda_ewald_kvectors(
        device.ctx, nk, ctypes.byref(count), index.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _ffi.as_double_pointer(factor),
        _ffi.as_double_pointer(rho)))
    assert count.value == nk
    np.testing.assert_allclose(rho, reference.ewald_rho(nk), rtol=0, atol=1e-12)
    positions = np.zeros((system.size(), 3))
    _ffi.check(device.ctx, device.lib.lumol_cuda_get_positions(device.ctx, _ffi.as_double_pointer(positions)))
    np.testing.assert_array_equal(positions, after.positions)
    # a second move of the same molecule starts from the accepted positions and the updated rho(k)
    rng = np.random.Generator(np.random.PCG64(11))
    trial = rigid_trial(after, 0, rng)
    costs = (_ffi.Energy * 1)()
    ids = np.array([0], dtype=np.int64)
    _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecules_cost(
        device.ctx, 1, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ffi.as_double_pointer(np.ascontiguousarray(trial)), costs))
    expected = reference.move_molecule_cost(0, trial)
    scale = magnitude(reference.energy_terms())
    assert abs(costs[0].coulomb_real - expected[1]) <= ENERGY_TOL * scale
    assert abs(costs[0].coulomb_kspace - expected[2]) <= ENERGY_TOL * scale


# ---- bench systems and synthetic boxes: batches of trial moves --------------------------------------------

@pytest.mark.parametrize("name", ["argon", "nacl_ewald", "nacl_wolf", "water_ewald", "water_wolf", "propane"])
def test_bench_systems_batched_trials(name):
    """benches/*.rs ``move_molecule_cost``: random molecules, random rigid trial positions, one batch."""
    builders = {
        "argon": systems.argon, "nacl_ewald": lambda: systems.nacl("ewald"), "nacl_wolf": lambda: systems.nacl("wolf"),
        "water_ewald": lambda: systems.water("ewald"), "water_wolf": lambda: systems.water("wolf"), "propane": systems.propane,
    }
    system = builders[name]()
    rng = np.random.Generator(np.random.PCG64(5))
    molecules = rng.integers(0, len(system.molecules()), 6)
    trials = [rigid_trial(system, int(m), rng) for m in molecules]
    costs, scale = check_against_oracle(system, molecules, trials)
    # and against the definition: full energy after minus full energy before, on the device
    before = device_for(system).compute(energy=True).energy.total()
    for molecule, positions, cost in list(zip(molecules, trials, costs))[:3]:
        after = device_for(moved(system, int(molecule), positions)).compute(energy=True).energy.total()
        assert abs(cost.sum() - (after - before)) <= ENERGY_TOL * scale


def test_lennard_jones_box_on_the_cell_list():
    system = synthetic.lj_box(16, seed=3)  # 4096 atoms, cell-list path for the full evaluations
    rng = np.random.Generator(np.random.PCG64(8))
    molecules = rng.integers(0, system.size(), 16)
    trials = [system.positions[m:m + 1] + rng.uniform(-1.0, 1.0, (1, 3)) for m in molecules]
    costs, scale = check_against_oracle(system, molecules, trials)
    cache = lumol.EnergyCache()
    cache.init(system)
    assert device_for(system).stats().neighbor_path == 1
    batch = cache.move_molecules_cost(system, molecules, trials)
    np.testing.assert_array_equal(batch, costs.sum(axis=1))
    cache.accept(5)
    m = int(molecules[5])
    system.positions[m] = trials[5][0]
    cache.update(system)
    assert abs(cache.energy() - system.potential_energy()) <= ENERGY_TOL * scale
    with pytest.raises(RuntimeError, match="without call a"):
        cache.update(system)


def test_spce_box_with_ewald_on_the_cell_list():
    system = synthetic.spce_box(10)  # 3000 atoms, L = 31 A: cell list, 1000 molecules
    ewald = lumol.SharedEwald(lumol.Ewald(9.0, 7, 0.3))
    ewald.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_coulomb_potential(ewald)
    rng = np.random.Generator(np.random.PCG64(21))
    molecules = rng.integers(0, 1000, 8)
    trials = [rigid_trial(system, int(m), rng, delta=0.5) for m in molecules]
    costs, scale = check_against_oracle(system, molecules, trials)

    cache = lumol.EnergyCache()
    cache.init(system)
    start = cache.energy()
    accepted = 0
    for step in range(4):
        molecule = int(molecules[step])
        positions = rigid_trial(system, molecule, rng, delta=0.5)
        cost = cache.move_molecule_cost(system, molecule, positions)
        bonding = system.molecule(molecule)
        system.positions[bonding.start:bonding.end] = positions
        cache.update(system)
        accepted += cost
    # four accepted moves later the running energy is still the energy of the system (rho(k) updated in place)
    assert abs(cache.energy() - (start + accepted)) <= 1e-12 * scale
    assert abs(cache.energy() - system.potential_energy()) <= ENERGY_TOL * scale
    resident = device_for(system, positions=False).compute(energy=True).energy.total()
    assert abs(resident - system.potential_energy()) <= 1e-12 * scale


def test_accept_needs_a_pending_cost():
    system = systems.wolf_cache_system()
    system.set_coulomb_potential(lumol.Wolf(8.0))
    device = device_for(system)
    with pytest.raises(lumol.LumolCudaError, match="without call a"):
        _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecule_accept(device.ctx, 0))
    device_costs(system, [0], [systems.WOLF_CACHE_MOVE])
    device.sync(system)  # uploads the positions again: the pending trial is stale
    with pytest.raises(lumol.LumolCudaError, match="without call a"):
        _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecule_accept(device.ctx, 0))
    device_costs(system, [0], [systems.WOLF_CACHE_MOVE])
    with pytest.raises(lumol.LumolCudaError, match="not one of the last cost call"):
        _ffi.check(device.ctx, device.lib.lumol_cuda_move_molecule_accept(device.ctx, 3))
    with pytest.raises(lumol.LumolCudaError, match="out of range"):
        device_costs(system, [7], [systems.WOLF_CACHE_MOVE])
