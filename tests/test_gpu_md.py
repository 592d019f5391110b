"""Device-resident molecular dynamics vs the oracle's integrators, and the reference's MD test criteria (``-m gpu``).

The reference's MD tests (tests/md-helium.rs, md-nacl.rs, md-water.rs) run short trajectories and assert
energy conservation / thermostat targets; the same criteria are applied to the CUDA path here, next to a
step-by-step comparison with the oracle restating lumol-sim/src/md/integrators.rs.
"""

import ctypes

import numpy as np
import pytest

import lumol_b200 as lumol
from lumol_b200 import _ffi, md
from lumol_b200.consts import K_BOLTZMANN
from lumol_b200.device import device_for
from oracle import oracle
import systems

pytestmark = pytest.mark.gpu


def oracle_trajectory(system, integrator, dt, nsteps):
    """Step ``system`` on the CPU with the oracle's restatement of the reference integrators."""
    orc = oracle.OracleSystem(system)
    lib = orc.lib
    position, velocity = orc.position, orc.velocity
    n = system.size()
    aux = np.zeros((n, 3))
    if integrator == "verlet":
        lib.orc_verlet_setup(orc.ref, oracle.dptr(aux), dt)
    for _ in range(nsteps):
        if integrator == "velocity_verlet":
            lib.orc_velocity_verlet_step(orc.ref, oracle.dptr(position), oracle.dptr(velocity), oracle.dptr(aux), dt)
        elif integrator == "verlet":
            lib.orc_verlet_step(orc.ref, oracle.dptr(position), oracle.dptr(velocity), oracle.dptr(aux), dt)
        else:
            lib.orc_leapfrog_step(orc.ref, oracle.dptr(position), oracle.dptr(velocity), oracle.dptr(aux), dt)
    return position.copy(), velocity.copy()


INTEGRATORS = {"velocity_verlet": md.VelocityVerlet, "verlet": md.Verlet, "leap_frog": md.LeapFrog}


@pytest.mark.parametrize("integrator", sorted(INTEGRATORS))
@pytest.mark.parametrize("builder", ["helium", "water", "nacl"])
def test_integrators_follow_the_oracle(integrator, builder):
    # (the bench configurations argon.pdb / propane.pdb hold overlapping atoms, E ~ 1e9: fine for timing a single
    # evaluation as benches/*.rs do, useless for dynamics; the MD test inputs of tests/data/md-* are used instead)
    system = {"helium": systems.md_helium, "water": systems.md_water, "nacl": lambda: systems.md_nacl("ewald")}[builder]()
    systems.random_velocities(system, 300.0, seed=11)
    nsteps, dt = 20, 1.0
    expected_x, expected_v = oracle_trajectory(system, integrator, dt, nsteps)
    propagator = md.MolecularDynamics(INTEGRATORS[integrator](dt))
    propagator.propagate(system, nsteps)
    # 20 steps: the trajectories only differ by force summation order, amplified by the dynamics
    assert np.abs(system.positions - expected_x).max() < 1e-9
    assert np.abs(system.velocities - expected_v).max() < 1e-9 * max(np.abs(expected_v).max(), 1e-3)


def test_first_velocity_verlet_step_uses_zero_accelerations():
    """VelocityVerlet::setup zeroes the accelerations (integrators.rs:40-42): after one step x = x0 + v0 dt exactly."""
    system = systems.md_helium()
    systems.random_velocities(system, 300.0, seed=5)
    x0, v0 = system.positions.copy(), system.velocities.copy()
    system.forces()  # leaves forces on the device; setup must ignore them
    md.MolecularDynamics(2.0).propagate(system, 1)
    np.testing.assert_array_equal(system.positions, x0 + v0 * 2.0)


def relative_drift(system, propagator, nsteps, chunks=10):
    energies = [system.total_energy()]
    for _ in range(chunks):
        propagator.propagate(system, nsteps // chunks)
        energies.append(system.total_energy())
    energies = np.array(energies)
    return np.abs((energies - energies[0]) / energies[0]).max()


def helium_with_velocities(**kwargs):
    system = systems.md_helium(**kwargs)
    systems.random_velocities(system, 300.0, seed=3)
    md.RemoveTranslation().control(system)
    md.scale(system, 300.0)
    return system


def test_md_helium_energy_conservation():
    # tests/md-helium.rs:19-66: 1000 steps, |dE / E| < 5e-3 (velocity-Verlet), 1e-2 (Verlet), 5e-3 (leap-frog)
    for integrator, threshold in ((md.VelocityVerlet, 5e-3), (md.Verlet, 1e-2), (md.LeapFrog, 5e-3)):
        system = helium_with_velocities()
        assert relative_drift(system, md.MolecularDynamics(integrator(1.0)), 1000) < threshold
    # tests/md-helium.rs:114-143: shifted cut-off 2e-3, tabulated potential 5e-3
    assert relative_drift(helium_with_velocities(shifted=True), md.MolecularDynamics(1.0), 1000) < 2e-3
    assert relative_drift(helium_with_velocities(table=True), md.MolecularDynamics(1.0), 1000) < 5e-3


def test_md_nacl_and_water_energy_conservation():
    # tests/md-nacl.rs:16-30 (Wolf NVE, 1e-4) and :60-94 (Ewald NVE, 5e-3), 100 steps of 1 fs
    for coulomb, threshold in (("wolf", 1e-4), ("ewald", 5e-3)):
        system = systems.md_nacl(coulomb)
        systems.random_velocities(system, 300.0, seed=8)
        assert relative_drift(system, md.MolecularDynamics(1.0), 100) < threshold
    # tests/md-water.rs:13-27: flexible water with Ewald, 1e-1
    system = systems.md_water()
    systems.random_velocities(system, 300.0, seed=9)
    assert relative_drift(system, md.MolecularDynamics(1.0), 100) < 1e-1


def ideal_gas(n_side=10):
    """lumol-sim/tests/thermostats.rs:14-33: 1000 non-interacting atoms on a lattice."""
    n = n_side ** 3
    grid = np.arange(n_side, dtype=np.float64)
    z, y, x = np.meshgrid(grid, grid, grid, indexing="ij")
    system = lumol.System(lumol.UnitCell.cubic(float(n_side)))
    system.add_particles(["He"] * n, np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1))
    systems.random_velocities(system, 300.0, seed=4)
    return system


def test_thermostats():
    # Rescale: exact target when outside the tolerance (lumol-sim/tests/thermostats.rs:35-52)
    system = ideal_gas()
    propagator = md.MolecularDynamics(1.0)
    propagator.set_thermostat(md.RescaleThermostat(250.0))
    propagator.propagate(system, 1)
    assert abs(system.temperature() - 250.0) < 1e-9
    # Berendsen: converges to the target (thermostats.rs:54-76 checks the mean)
    system = ideal_gas()
    propagator = md.MolecularDynamics(1.0)
    propagator.set_thermostat(md.BerendsenThermostat(250.0, 20.0))
    propagator.propagate(system, 500)
    assert abs(system.temperature() - 250.0) < 1e-3 * 250.0
    # one Berendsen step against the oracle's factor
    system = ideal_gas()
    instant = system.temperature()
    v0 = system.velocities.copy()
    propagator = md.MolecularDynamics(1.0)
    propagator.set_thermostat(md.BerendsenThermostat(250.0, 50.0))
    propagator.propagate(system, 1)
    factor = oracle.library().orc_berendsen_thermostat_factor(250.0, instant, 50.0)
    np.testing.assert_allclose(system.velocities, v0 * factor, rtol=1e-13)
    # CSVR: mean and variance of the kinetic energy (thermostats.rs:78-118, tolerance 1e-2 on both here)
    system = ideal_gas()
    propagator = md.MolecularDynamics(1.0)
    propagator.set_thermostat(md.CSVRThermostat(250.0, 10.0))
    propagator.propagate(system, 200)
    temperatures = []
    for _ in range(400):
        propagator.propagate(system, 5, download=False)
        temperatures.append(system._device.kinetic_energy() * 2.0 / (system.degrees_of_freedom() * K_BOLTZMANN))
    temperatures = np.array(temperatures)
    dof = system.degrees_of_freedom()
    assert abs(temperatures.mean() - 250.0) < 1e-2 * 250.0
    expected_variance = 2.0 * 250.0 ** 2 / dof
    assert abs(temperatures.var() - expected_variance) < 0.3 * expected_variance


def test_remove_translation_and_scale():
    system = systems.md_water()
    systems.random_velocities(system, 300.0, seed=6)
    system.velocities += np.array([1e-3, -2e-3, 5e-4])
    expected = system.velocities.copy()
    oracle.library().orc_remove_translation(system.size(), oracle.dptr(np.ascontiguousarray(system.masses)), oracle.dptr(expected))
    md.RemoveTranslation().control(system)
    np.testing.assert_allclose(system.velocities, expected, rtol=0, atol=1e-15)
    md.scale(system, 123.0)
    assert abs(system.temperature() - 123.0) < 1e-10


def test_controls_inside_the_device_loop():
    system = systems.md_helium()
    systems.random_velocities(system, 300.0, seed=12)
    system.velocities += 1e-3
    propagator = md.MolecularDynamics(1.0)
    propagator.add_control(md.RemoveTranslation())
    propagator.propagate(system, 5)
    momentum = (system.masses[:, None] * system.velocities).sum(axis=0)
    assert np.abs(momentum).max() < 1e-12 * system.masses.sum()


def test_large_box_properties():
    """Size-independent checks at a size the O(N^2) oracle cannot reach (262 144 atoms): Newton's third law,
    the cell path against shifted periodic copies, and NVE drift."""
    system = systems.lj_box(64, seed=20240 + 18)
    device = device_for(system)
    result = device.compute(forces=True, energy=True, virial=True)
    assert device.stats().neighbor_path == 1
    forces = result.forces
    assert np.abs(forces.sum(axis=0)).max() < 1e-9 * np.abs(forces).max() * np.sqrt(system.size())
    # translating every atom by a lattice vector (or any vector) leaves forces and energy unchanged
    moved = systems.lj_box(64, seed=20240 + 18)
    moved.positions += np.array([3.0 * moved.cell.a(), -1.2345, 17.5])
    again = device_for(moved).compute(forces=True, energy=True, virial=True)
    assert np.abs(again.forces - forces).max() < 1e-9 * np.abs(forces).max()
    assert abs(again.energy.pairs - result.energy.pairs) < 1e-10 * abs(result.energy.pairs)
    assert np.abs(again.virial - result.virial).max() < 1e-9 * np.abs(result.virial).max()
    # a random sample of atoms against a direct minimum-image sum on the host
    rng = np.random.Generator(np.random.PCG64(1))
    length = system.cell.a()
    sigma, epsilon, cutoff = 3.4, lumol.units.from_(1.0, "kJ/mol"), 10.0
    for i in rng.integers(0, system.size(), 24):
        d = system.positions[i] - system.positions
        d -= np.round(d / length) * length
        r2 = (d * d).sum(axis=1)
        mask = (r2 < cutoff * cutoff) & (r2 > 0)
        s6 = (sigma * sigma / r2[mask]) ** 3
        fr = 24.0 * epsilon * (2.0 * s6 * s6 - s6) / r2[mask]
        expected = (fr[:, None] * d[mask]).sum(axis=0)
        assert np.abs(forces[i] - expected).max() < 1e-10 * np.abs(forces).max()
    systems.random_velocities(system, 120.0, seed=2)
    assert relative_drift(system, md.MolecularDynamics(1.0), 100, chunks=2) < 5e-4


# ---- barostats and the remaining controls ------------------------------------------------------------------------------

def oracle_barostat_trajectory(system, nsteps, dt, tau, pressure=None, stress=None):
    orc = oracle.OracleSystem(system)
    n = system.size()
    acc = np.zeros((n, 3))
    if stress is None:
        eta = ctypes.c_double(1.0)
        for _ in range(nsteps):
            assert orc.lib.orc_berendsen_barostat_step(orc.ref, oracle.dptr(orc.position), oracle.dptr(orc.velocity), oracle.dptr(acc), dt,
                                                       pressure, tau, ctypes.byref(eta), 0.0) == 0
    else:
        eta = np.eye(3).reshape(-1).copy()
        target = np.ascontiguousarray(np.array(stress, dtype=np.float64).reshape(-1))
        for _ in range(nsteps):
            assert orc.lib.orc_aniso_berendsen_barostat_step(orc.ref, oracle.dptr(orc.position), oracle.dptr(orc.velocity), oracle.dptr(acc),
                                                             dt, oracle.dptr(target), tau, oracle.dptr(eta), 0.0) == 0
    return orc.position.copy(), orc.velocity.copy(), np.array(orc.s.cell[:]).reshape(3, 3)


@pytest.mark.parametrize("builder", ["helium", "nacl"])
def test_berendsen_barostats_follow_the_oracle(builder):
    """BerendsenBarostat and AnisoBerendsenBarostat (integrators.rs:176-342): ten device steps against the oracle's
    restatement, positions, velocities and the scaled cell."""
    build = {"helium": systems.md_helium, "nacl": lambda: systems.md_nacl("wolf")}[builder]
    pressure = lumol.units.from_(500.0, "bar")
    for kind in ("isotropic", "anisotropic"):
        system = build()
        systems.random_velocities(system, 300.0, seed=5)
        reference = build()
        reference.velocities = system.velocities.copy()
        if kind == "isotropic":
            integrator = md.BerendsenBarostat(1.0, pressure, 100.0)
            x, v, cell = oracle_barostat_trajectory(reference, 10, 1.0, 100.0, pressure=pressure)
        else:
            stress = pressure * np.eye(3)
            stress[0, 1] = stress[1, 0] = 0.1 * pressure
            integrator = md.AnisoBerendsenBarostat(1.0, stress, 100.0)
            x, v, cell = oracle_barostat_trajectory(reference, 10, 1.0, 100.0, stress=stress)
        md.MolecularDynamics(integrator).propagate(system, 10)
        assert np.abs(system.positions - x).max() < 1e-9
        assert np.abs(system.velocities - v).max() < 1e-10 * max(np.abs(v).max(), 1e-3)
        assert np.abs(system.cell.matrix() - cell).max() < 1e-10 * np.abs(cell).max()
        assert not np.allclose(cell, reference.cell.matrix(), rtol=1e-9, atol=0)  # the box did change


def test_berendsen_barostat_reaches_the_target_pressure_and_refuses_small_cells():
    # tests/md-helium.rs:96-111 (npt-berendsen-barostat.toml): pressure within 5e-2 relative after equilibration
    system = systems.md_helium()
    systems.random_velocities(system, 300.0, seed=3)
    target = lumol.units.from_(5000.0, "bar")
    propagator = md.MolecularDynamics(md.BerendsenBarostat(1.0, target, 1000.0))
    propagator.set_thermostat(md.BerendsenThermostat(300.0, 100.0))
    propagator.propagate(system, 3000)
    pressures = []
    for _ in range(20):
        propagator.propagate(system, 50)
        pressures.append(system.pressure())
    assert abs(np.mean(pressures) - target) / target < 5e-2
    # integrators.rs:227-236
    small = systems.md_helium()
    systems.random_velocities(small, 300.0, seed=3)
    squeeze = md.MolecularDynamics(md.BerendsenBarostat(1.0, lumol.units.from_(1e7, "bar"), 10.0))
    with pytest.raises(lumol.LumolCudaError, match="Tried to decrease the cell size in Berendesen barostat"):
        squeeze.propagate(small, 2000)


def test_remove_rotation_and_rewrap():
    # controls.rs:108-131 on the device
    system = lumol.System(lumol.UnitCell.cubic(10.0))
    system.add_particles(["Ag", "Ag"], np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]), masses=np.array([107.8682, 107.8682]))
    system.velocities = np.array([[0.0, 1.0, 0.0], [0.0, -1.0, 2.0]])
    md.RemoveRotation().control(system)
    np.testing.assert_allclose(system.velocities, [[0.0, 0.0, 1.0], [0.0, 0.0, 1.0]], rtol=0, atol=1e-15)
    system.positions = np.array([[0.0, 0.0, 0.0], [15.0, 0.0, 0.0]])
    md.Rewrap().control(system)
    np.testing.assert_array_equal(system.positions, [[0.0, 0.0, 0.0], [5.0, 0.0, 0.0]])
    # a molecular system against the oracle, then both controls inside the device loop
    water = systems.md_water()
    systems.random_velocities(water, 300.0, seed=6)
    rng = np.random.Generator(np.random.PCG64(2))
    water.positions += rng.integers(-2, 3, (water.size() // 3, 1, 3)).repeat(3, axis=1).reshape(-1, 3) * water.cell.a()
    expected_v = water.velocities.copy()
    oracle.library().orc_remove_rotation(water.size(), oracle.dptr(np.ascontiguousarray(water.masses)),
                                         oracle.dptr(np.ascontiguousarray(water.positions)), oracle.dptr(expected_v))
    reference = oracle.OracleSystem(water)
    expected_x = water.positions.copy()
    reference.lib.orc_rewrap(reference.ref, oracle.dptr(expected_x))
    md.RemoveRotation().control(water)
    np.testing.assert_allclose(water.velocities, expected_v, rtol=0, atol=1e-12 * np.abs(expected_v).max())
    md.Rewrap().control(water)
    np.testing.assert_allclose(water.positions, expected_x, rtol=0, atol=1e-12)
    propagator = md.MolecularDynamics(1.0)
    propagator.add_control(md.RemoveTranslation())
    propagator.add_control(md.RemoveRotation())
    propagator.add_control(md.Rewrap())
    propagator.propagate(water, 5)
    centers = water.positions.reshape(-1, 3, 3)
    weights = water.masses.reshape(-1, 3, 1)
    com = (centers * weights).sum(axis=1) / weights.sum(axis=1)
    assert (com >= -1e-9).all() and (com < water.cell.a() + 1e-9).all()


def test_sorted_resident_md_equals_the_step_by_step_path(monkeypatch):
    """NVE velocity-Verlet of a single-LJ system on the neighbour-list path runs with the state in cell order for the
    duration of lumol_cuda_md_run (pairs_lj2.cu): same trajectory as the path that keeps the caller's atom order
    (LUMOL_CUDA_SORTED_MD=0), across neighbour-list rebuilds and across two calls."""
    import os

    from lumol_b200 import synthetic
    from lumol_b200.device import DeviceSystem

    results = []
    for knob in ("0", "1"):
        monkeypatch.setenv("LUMOL_CUDA_SORTED_MD", knob)
        system = synthetic.lj_box(32, seed=20240 + 15)
        synthetic.maxwell_boltzmann(system, 300.0, seed=3)
        device = DeviceSystem(0)
        device.sync(system, velocities=True)
        lib, ctx = device.lib, device.ctx
        _ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, _ffi.INTEGRATOR_VELOCITY_VERLET, 1.0))
        for steps in (25, 45):
            _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, steps))
        n = system.size()
        x, v, f = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        _ffi.check(ctx, lib.lumol_cuda_get_positions(ctx, _ffi.as_double_pointer(x)))
        _ffi.check(ctx, lib.lumol_cuda_get_velocities(ctx, _ffi.as_double_pointer(v)))
        _ffi.check(ctx, lib.lumol_cuda_get_forces(ctx, _ffi.as_double_pointer(f)))
        results.append((x, v, f, device.stats().neighbor_rebuilds))
        # forces of the final state against an evaluation from scratch
        fresh = device.compute(forces=True).forces
        assert np.abs(fresh - f).max() < 1e-11 * np.abs(f).max()
        device.close()
    (x0, v0, f0, r0), (x1, v1, f1, r1) = results
    assert r0 >= 2 and r1 >= 2
    assert np.abs(x1 - x0).max() < 1e-10
    assert np.abs(v1 - v0).max() < 1e-10 * np.abs(v0).max()
    assert np.abs(f1 - f0).max() < 1e-9 * np.abs(f0).max()
