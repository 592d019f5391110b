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


# ---- flexible molecules: tests/md-butane.rs, tests/md-methane.rs ------------------------------------------------

def molecules_fixture(name):
    import os

    import lumol_b200 as lumol

    data = np.load(os.path.join(systems.ROOT, "tests", "golden", "md_molecules.npz"))
    cell = data[name + "/cell"]
    system = lumol.System(lumol.UnitCell.ortho(cell[0], cell[1], cell[2]))
    system.add_particles([str(n) for n in data[name + "/names"]], data[name + "/positions"])
    system.add_bonds(data[name + "/bonds"])
    return system


def butane():
    """tests/data/md-butane/{butane.toml,nve.toml}: 50 united-atom butanes, shifted LJ rc 10 A inter-molecular,
    harmonic bonds and angles, torsion."""
    import lumol_b200 as lumol
    from lumol_b200 import units

    system = molecules_fixture("md-butane")
    kcal = units.from_(1.0, "kcal/mol")
    pair = lumol.PairInteraction.shifted(lumol.LennardJones(sigma=3.4, epsilon=0.7 * kcal), 10.0)
    pair.set_restriction(lumol.PairRestriction.InterMolecular)
    system.set_pair_potential(("C", "C"), pair)
    system.set_bond_potential(("C", "C"), lumol.Harmonic(x0=1.53, k=units.from_(225.0, "kcal/mol/A^2")))
    system.set_angle_potential(("C", "C", "C"), lumol.Harmonic(x0=units.from_(115.0, "deg"), k=units.from_(58.0, "kcal/mol/rad^2")))
    system.set_dihedral_potential(("C", "C", "C", "C"), lumol.Torsion(n=3, delta=units.from_(180.0, "deg"), k=1.5 * kcal))
    return system


def methane():
    """tests/data/md-methane/{methane.toml,nve.toml}: 150 methanes, shifted LJ on C-C, harmonic C-H bonds and H-C-H angles."""
    import lumol_b200 as lumol
    from lumol_b200 import units

    system = molecules_fixture("md-methane")
    kcal = units.from_(1.0, "kcal/mol")
    system.set_pair_potential(("C", "C"), lumol.PairInteraction.shifted(lumol.LennardJones(sigma=3.7, epsilon=0.2981 * kcal), 10.0))
    system.set_pair_potential(("C", "H"), lumol.PairInteraction.shifted(lumol.NullPotential(), 10.0))
    system.set_pair_potential(("H", "H"), lumol.PairInteraction.shifted(lumol.NullPotential(), 10.0))
    system.set_bond_potential(("C", "H"), lumol.Harmonic(x0=1.09, k=units.from_(390.0, "kcal/mol/A^2")))
    system.set_angle_potential(("H", "C", "H"), lumol.Harmonic(x0=units.from_(109.5, "deg"), k=units.from_(70.0, "kcal/mol/rad^2")))
    return system


def test_md_butane_topology_and_conservation():
    # tests/md-butane.rs:13-29: 50 molecules of 3 bonds, 2 angles, 1 dihedral; :32-47: 1000 steps, |dE / E| < 1e-3
    system = butane()
    assert len(system.molecules()) == 50
    for molecule in system.molecules():
        assert (len(molecule.bonds), len(molecule.angles), len(molecule.dihedrals)) == (3, 2, 1)
    orc = oracle.OracleSystem(system)
    assert (len(orc.angles), len(orc.dihedrals)) == (100, 50)
    prepared(system, 300.0, seed=12)
    assert relative_drift(system, "velocity_verlet", 1000) < 1e-3


def test_md_methane_topology_and_conservation():
    # tests/md-methane.rs:13-28: 150 molecules of 4 bonds, 6 angles, no dihedral; :31-45: 500 steps, |dE / E| < 1e-2
    system = methane()
    assert len(system.molecules()) == 150
    for molecule in system.molecules():
        assert (len(molecule.bonds), len(molecule.angles), len(molecule.dihedrals)) == (4, 6, 0)
    prepared(system, 300.0, seed=13)
    assert relative_drift(system, "velocity_verlet", 500) < 1e-2
