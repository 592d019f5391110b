"""The oracle's restatement of the reference integrators against the reference's own MD tests (CPU): the energy
conservation thresholds of tests/md-helium.rs and tests/md-nacl.rs hold for trajectories stepped by the oracle."""

import numpy as np
import pytest

from lumol_b200.consts import K_BOLTZMANN
from oracle import oracle
import systems


def prepared(system, temperature, seed):
    """BoltzmannVelocities + RemoveTranslation + scale (velocities.rs:16-22, controls.rs:30-41) on the host."""
    systems.random_velocities(system, temperature, seed)
    momentum = (system.masses[:, None] * system.velocities).sum(axis=0)
    system.velocities -= momentum / system.masses.sum()
    instant = oracle.OracleSystem(system).temperature()
    system.velocities *= np.sqrt(temperature / instant)
    return system


def total_energy(system, position, velocity):
    system.positions, system.velocities = position.copy(), velocity.copy()
    reference = oracle.OracleSystem(system)
    return reference.potential_energy() + reference.kinetic_energy()


def relative_drift(system, integrator, nsteps, dt=1.0):
    orc = oracle.OracleSystem(system)
    lib = orc.lib
    position, velocity = orc.position, orc.velocity
    aux = np.zeros((system.size(), 3))
    before = total_energy(system, position, velocity)
    if integrator == "verlet":
        lib.orc_verlet_setup(orc.ref, oracle.dptr(aux), dt)
    step = {"velocity_verlet": lib.orc_velocity_verlet_step, "verlet": lib.orc_verlet_step, "leap_frog": lib.orc_leapfrog_step}[integrator]
    for _ in range(nsteps):
        step(orc.ref, oracle.dptr(position), oracle.dptr(velocity), oracle.dptr(aux), dt)
    after = total_energy(system, position, velocity)
    assert abs(np.abs(position).max()) < 1e6
    return abs((after - before) / before)


@pytest.mark.parametrize("integrator,threshold", [("velocity_verlet", 5e-3), ("verlet", 1e-2), ("leap_frog", 5e-3)])
def test_md_helium_conservation(integrator, threshold):
    # tests/md-helium.rs:19-66: 1000 steps of 1 fs at 300 K
    system = prepared(systems.md_helium(), 300.0, seed=3)
    assert relative_drift(system, integrator, 1000) < threshold


def test_md_helium_shifted_and_tabulated():
    # tests/md-helium.rs:114-143
    assert relative_drift(prepared(systems.md_helium(shifted=True), 300.0, seed=3), "velocity_verlet", 1000) < 2e-3


def test_md_nacl_wolf_conservation():
    # tests/md-nacl.rs:16-30: Wolf, 100 steps, 1e-4
    system = prepared(systems.md_nacl("wolf"), 300.0, seed=8)
    assert relative_drift(system, "velocity_verlet", 100) < 1e-4


def test_temperature_after_scaling_is_the_target():
    system = prepared(systems.md_helium(), 300.0, seed=5)
    reference = oracle.OracleSystem(system)
    assert abs(reference.temperature() - 300.0) < 1e-9
    kinetic = 0.5 * (system.masses[:, None] * system.velocities ** 2).sum()
    assert abs(reference.kinetic_energy() - kinetic) < 1e-12 * kinetic
    dof = 3 * system.size()
    assert abs(reference.temperature() - 2.0 * kinetic / (dof * K_BOLTZMANN)) < 1e-9
