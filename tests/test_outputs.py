"""``lumol_sim::output`` mirror: text formats, output frequencies, and -- on the device --
the reference's own golden lines (lumol-sim/src/output/*.rs tests, system of output/tests.rs:35-50)."""

import numpy as np
import pytest

import lumol_b200 as lumol
from lumol_b200 import md, output, units
from oracle import oracle


def output_system():
    """output/tests.rs:35-50"""
    system = lumol.System(lumol.UnitCell.cubic(10.0))
    system.add_molecule(lumol.Molecule(lumol.Particle("F", (0.0, 0.0, 0.0))))
    system.add_molecule(lumol.Molecule(lumol.Particle("F", (1.3, 0.0, 0.0))))
    system.velocities[0] = [0.1, 0.0, 0.0]
    system.velocities[1] = [0.0, 0.0, 0.0]
    harmonic = lumol.Harmonic(k=units.from_(300.0, "kJ/mol/A^2"), x0=units.from_(1.2, "A"))
    system.set_pair_potential(("F", "F"), lumol.PairInteraction(harmonic, 5.0))
    system.step = 42
    return system


def written(factory, system, tmp_path):
    path = tmp_path / "out.dat"
    out = factory(str(path))
    out.setup(system)
    out.write(system)
    out.finish(system)
    out.close()
    return path.read_text(encoding="utf-8").splitlines()


def test_display_is_rust_display():
    for value, text in ((0.0, "0"), (1000.0, "1000"), (30.000000000000025, "30.000000000000025"), (-0.5, "-0.5"),
                        (1e-7, "0.0000001"), (1e21, "1000000000000000000000"), (90.0, "90"), (float("inf"), "inf")):
        assert output.display(value) == text


def test_cell_output_golden(tmp_path):
    # output/cell.rs:57-66
    assert written(output.CellOutput, output_system(), tmp_path) == [
        "# Unit cell of the simulation", "# Step A/Å B/Å C/Å α/deg β/deg γ/deg", "42 10 10 10 90 90 90"]


def test_oracle_reproduces_the_output_goldens():
    """The numbers of the reference's output tests, from the oracle: energy.rs:56-62, forces.rs:60-66,
    stress.rs:66-71, properties.rs:60-66."""
    system = output_system()
    reference = oracle.OracleSystem(system)
    potential = units.to(reference.potential_energy(), "kJ/mol")
    kinetic = units.to(reference.kinetic_energy(), "kJ/mol")
    total = units.to(reference.potential_energy() + reference.kinetic_energy(), "kJ/mol")
    assert [output.display(v) for v in (potential, kinetic, total)] == ["1.5000000000000027", "949.9201593348566", "951.4201593348566"]
    forces = reference.forces() * units.to(1.0, "kJ/mol/A")
    assert [output.display(v) for v in forces[0]] == ["30.000000000000025", "0", "0"]
    assert [output.display(v) for v in forces[1]] == ["-30.000000000000025", "0", "0"]
    stress = reference.stress() * units.to(1.0, "bar")
    assert output.display(stress[0][0]) == "30899.975184239443"
    assert output.display(units.to(reference.temperature(), "K")) == "38083.04389172312"
    assert output.display(units.to(reference.pressure(), "bar")) == "10299.991728079816"


class RecordingPropagator:
    def __init__(self):
        self.blocks = []

    def setup(self, system):
        pass

    def propagate(self, system, nsteps, download=True):
        self.blocks.append(nsteps)
        system.step += nsteps


class RecordingOutput(output.Output):
    def __init__(self):
        self.steps = []

    def write(self, system):
        self.steps.append(system.step)


def test_simulation_runs_whole_blocks_between_outputs():
    """simulations.rs:88-95: every output sees exactly the steps that are multiples of its frequency, and the
    propagator is entered once per block of steps, not once per step."""
    system = lumol.System(lumol.UnitCell.cubic(10.0))
    propagator = RecordingPropagator()
    simulation = md.Simulation(propagator)
    every3, every5 = RecordingOutput(), RecordingOutput()
    simulation.add_output_with_frequency(every3, 3)
    simulation.add_output_with_frequency(every5, 5)
    simulation.run(system, 11)
    assert every3.steps == [3, 6, 9] and every5.steps == [5, 10]
    assert propagator.blocks == [3, 2, 1, 3, 1, 1] and system.step == 11
    # continuing from step 11
    simulation.run(system, 4)
    assert every3.steps == [3, 6, 9, 12, 15] and every5.steps == [5, 10, 15]
    plain = md.Simulation(RecordingPropagator())
    plain.run(system, 1000)
    assert plain.propagator.blocks == [1000]
    with pytest.raises(ValueError):
        simulation.add_output_with_frequency(RecordingOutput(), 0)


# ---- on the device ---------------------------------------------------------------------------------------------

def numbers(line):
    return np.array([float(v) for v in line.split()[1:]])


@pytest.mark.gpu
def test_output_goldens_on_the_device(tmp_path):
    system = output_system()
    lines = written(output.EnergyOutput, system, tmp_path)
    assert lines[:2] == ["# Energy of the simulation (kJ/mol)", "# Step Potential Kinetic Total"]
    assert lines[2].startswith("42 ")
    np.testing.assert_allclose(numbers(lines[2]), [1.5000000000000027, 949.9201593348566, 951.4201593348566], rtol=1e-13)
    lines = written(output.ForcesOutput, system, tmp_path)
    assert lines[:2] == ["2", "forces in kJ/mol/A at step 42"]
    assert lines[2].split()[0] == "F" and lines[3].split()[0] == "F"
    np.testing.assert_allclose(numbers(lines[2]), [30.000000000000025, 0.0, 0.0], rtol=1e-13)
    np.testing.assert_allclose(numbers(lines[3]), [-30.000000000000025, 0.0, 0.0], rtol=1e-13)
    lines = written(output.StressOutput, system, tmp_path)
    assert lines[1] == "# step stress.xx stress.yy stress.zz stress.xy stress.xz stress.yz"
    np.testing.assert_allclose(numbers(lines[2]), [30899.975184239443, 0, 0, 0, 0, 0], rtol=1e-13, atol=1e-9)
    lines = written(output.PropertiesOutput, system, tmp_path)
    assert lines[1] == "# Step Volume/A^3 Temperature/K Pressure/bar"
    np.testing.assert_allclose(numbers(lines[2]), [1000.0, 38083.04389172312, 10299.991728079816], rtol=1e-13)


@pytest.mark.gpu
def test_outputs_read_the_resident_state(tmp_path):
    """Between two output steps nothing is downloaded: the energies an output writes are those of the device state,
    equal to what a run that downloads after every step reports, and the host arrays are refreshed when the run ends."""
    import systems

    def build():
        system = systems.lj_box(6, seed=2)
        systems.random_velocities(system, 120.0, seed=4)
        return system

    system = build()
    simulation = md.Simulation(md.MolecularDynamics(1.0))
    path = tmp_path / "energy.dat"
    energy = output.EnergyOutput(str(path))
    trajectory = output.TrajectoryOutput(str(tmp_path / "traj.xyz"))
    simulation.add_output_with_frequency(energy, 10)
    simulation.add_output_with_frequency(trajectory, 20)
    start = system.positions.copy()
    simulation.run(system, 40)
    energy.close()
    trajectory.close()
    lines = path.read_text().splitlines()[2:]
    assert [int(line.split()[0]) for line in lines] == [10, 20, 30, 40]

    stepwise = build()
    propagator = md.MolecularDynamics(1.0)
    expected = []
    for block in range(4):
        propagator.propagate(stepwise, 10)
        expected.append([units.to(stepwise.potential_energy(), "kJ/mol"), units.to(stepwise.kinetic_energy(), "kJ/mol")])
    got = np.array([[float(v) for v in line.split()[1:3]] for line in lines])
    np.testing.assert_allclose(got, np.array(expected), rtol=1e-9)
    # the run ended: host arrays are those of the device
    assert not system._resident
    np.testing.assert_allclose(system.positions, stepwise.positions, rtol=0, atol=1e-9)
    assert np.abs(system.positions - start).max() > 1e-3
    frames = (tmp_path / "traj.xyz").read_text().splitlines()
    assert frames[0] == str(system.size()) and frames[1] == "step 20"
    assert len(frames) == 2 * (system.size() + 2)
