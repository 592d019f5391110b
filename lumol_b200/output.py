"""Host-side mirror of ``lumol_sim::output`` for the quantities that come out of the hot path (SURVEY section 8f,
N4): energy, forces, stress, properties, cell, plus a plain XYZ trajectory.

The point of this module is *when* data leaves the device: ``Simulation.run`` (md.py) runs whole blocks of steps on
the device between two output steps; an output evaluates what it needs from the resident state (energies, forces,
stress: one evaluation, a few scalars or one n x 3 array back) and only ``TrajectoryOutput`` downloads positions.
Text formats and headers are the reference's (lumol-sim/src/output/*.rs); numbers are formatted like Rust's
``Display`` for ``f64`` (shortest round-trip digits, positional, no trailing ``.0``).  chemfiles trajectory
formats stay on the Rust side.
"""

import numpy as np

from . import units


def display(value):
    """Rust ``format!("{}", f64)``."""
    value = float(value)
    if value != value:
        return "NaN"
    if value in (float("inf"), float("-inf")):
        return "inf" if value > 0 else "-inf"
    return np.format_float_positional(value, unique=True, trim="-")


class Output:
    """``Output`` trait (output/mod.rs:19-30)."""

    def setup(self, system):
        pass

    def write(self, system):
        raise NotImplementedError

    def finish(self, system):
        pass


class _FileOutput(Output):
    def __init__(self, filename):
        self.path = filename
        self.file = open(filename, "w", encoding="utf-8")

    def _line(self, text):
        self.file.write(text + "\n")

    def finish(self, system):
        self.file.flush()

    def close(self):
        self.file.close()


class EnergyOutput(_FileOutput):
    """output/energy.rs:34-46: ``step potential kinetic total`` in kJ/mol."""

    def setup(self, system):
        self._line("# Energy of the simulation (kJ/mol)")
        self._line("# Step Potential Kinetic Total")

    def write(self, system):
        potential = units.to(system.potential_energy(), "kJ/mol")
        kinetic = units.to(system.kinetic_energy(), "kJ/mol")
        total = units.to(system.total_energy(), "kJ/mol")
        self._line(f"{system.step} {display(potential)} {display(kinetic)} {display(total)}")


class ForcesOutput(_FileOutput):
    """output/forces.rs:32-49: XYZ-like frames of forces in kJ/mol/A."""

    def write(self, system):
        forces = system.forces()
        conversion = units.to(1.0, "kJ/mol/A")
        self._line(f"{len(forces)}")
        self._line(f"forces in kJ/mol/A at step {system.step}")
        for name, force in zip(system.names, forces):
            x, y, z = conversion * force[0], conversion * force[1], conversion * force[2]
            self._line(f"{name} {display(x)} {display(y)} {display(z)}")


class StressOutput(_FileOutput):
    """output/stress.rs:32-56: ``step xx yy zz xy xz yz`` in bar."""

    def setup(self, system):
        self._line("# Stress tensor of the simulation (bar)")
        self._line("# step stress.xx stress.yy stress.zz stress.xy stress.xz stress.yz")

    def write(self, system):
        conversion = units.to(1.0, "bar")
        stress = system.stress()
        values = [stress[0][0], stress[1][1], stress[2][2], stress[0][1], stress[0][2], stress[1][2]]
        self._line(f"{system.step} " + " ".join(display(v * conversion) for v in values))


class PropertiesOutput(_FileOutput):
    """output/properties.rs:40-51: ``step volume temperature pressure``."""

    def setup(self, system):
        self._line("# Physical properties of the simulation")
        self._line("# Step Volume/A^3 Temperature/K Pressure/bar")

    def write(self, system):
        volume = units.to(system.volume(), "A^3")
        temperature = units.to(system.temperature(), "K")
        pressure = units.to(system.pressure(), "bar")
        self._line(f"{system.step} {display(volume)} {display(temperature)} {display(pressure)}")


class CellOutput(_FileOutput):
    """output/cell.rs:33-49"""

    def setup(self, system):
        self._line("# Unit cell of the simulation")
        self._line("# Step A/Å B/Å C/Å α/deg β/deg γ/deg")

    def write(self, system):
        cell = system.cell
        values = [cell.a(), cell.b(), cell.c(), cell.alpha(), cell.beta(), cell.gamma()]
        self._line(f"{system.step} " + " ".join(display(v) for v in values))


class TrajectoryOutput(_FileOutput):
    """Plain XYZ frames (the reference writes through chemfiles, output/trajectory.rs:58-67; only this output needs
    the positions on the host, so only it downloads them)."""

    def write(self, system):
        system.sync_from_device()
        self._line(f"{system.size()}")
        self._line(f"step {system.step}")
        for name, position in zip(system.names, system.positions):
            self._line(f"{name} {display(position[0])} {display(position[1])} {display(position[2])}")


__all__ = ["Output", "EnergyOutput", "ForcesOutput", "StressOutput", "PropertiesOutput", "CellOutput", "TrajectoryOutput", "display"]
