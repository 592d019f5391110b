"""Host-side mirror of ``lumol_core::sys::compute`` and ``sys::energy``: the ``Compute`` estimators.

Each ``compute(system)`` makes one call into the CUDA library through the C ABI; the small amount of
scalar post-processing (temperature, pressure, stress) follows the reference line by line
(lumol-core/src/sys/compute.rs:111-480, lumol-core/src/sys/energy.rs:16-158).
"""

import numpy as np

from . import _ffi
from .consts import K_BOLTZMANN
from .device import device_for


class Compute:
    """``Compute`` trait (compute.rs:21-26)."""

    def compute(self, system):
        raise NotImplementedError


class Forces(Compute):
    """compute.rs:30-108: pair + bonded + coulomb forces, an (n, 3) array."""

    def compute(self, system):
        return device_for(system).compute(forces=True).forces


class PotentialEnergy(Compute):
    """compute.rs:111-128"""

    def compute(self, system):
        return device_for(system).compute(energy=True).energy.total()


class KineticEnergy(Compute):
    """compute.rs:133-144"""

    def compute(self, system):
        return device_for(system, velocities=True).kinetic_energy()


class TotalEnergy(Compute):
    """compute.rs:147-155"""

    def compute(self, system):
        kinetic = KineticEnergy().compute(system)
        potential = PotentialEnergy().compute(system)
        return kinetic + potential


class Temperature(Compute):
    """compute.rs:164-172"""

    def compute(self, system):
        kinetic = KineticEnergy().compute(system)
        dof = float(system.degrees_of_freedom())
        return 2.0 * kinetic / (dof * K_BOLTZMANN)


class Volume(Compute):
    """compute.rs:175-182"""

    def compute(self, system):
        return system.cell.volume()


class AtomicVirial(Compute):
    """compute.rs:195-255"""

    def compute(self, system):
        if system.cell.is_infinite():
            raise ValueError("Can not compute virial for infinite cell")
        return device_for(system).compute(virial=True).virial


class MolecularVirial(Compute):
    """compute.rs:278-364"""

    def compute(self, system):
        if system.cell.is_infinite():
            raise ValueError("Can not compute virial for infinite cell")
        return device_for(system).compute(molecular_virial=True, parts=_ffi.PART_PAIRS | _ffi.PART_COULOMB).virial


class Virial(Compute):
    """compute.rs:372-381: molecular virial when molecules are the simulated degrees of freedom."""

    def compute(self, system):
        if system.simulated_degrees_of_freedom[0] == "molecules":
            return MolecularVirial().compute(system)
        return AtomicVirial().compute(system)


class PressureAtTemperature(Compute):
    """compute.rs:393-408"""

    def __init__(self, temperature):
        self.temperature = temperature

    def compute(self, system):
        if system.cell.is_infinite():
            raise ValueError("Can not compute pressure for infinite cell")
        assert self.temperature >= 0.0, "assertion failed: self.temperature >= 0.0"
        virial = float(np.trace(Virial().compute(system)))
        volume = system.volume()
        dof = float(system.degrees_of_freedom())
        return (dof * K_BOLTZMANN * self.temperature + virial) / (3.0 * volume)


class Pressure(Compute):
    """compute.rs:419-428"""

    def compute(self, system):
        if system.cell.is_infinite():
            raise ValueError("Can not compute pressure for infinite cell")
        return PressureAtTemperature(Temperature().compute(system)).compute(system)


class StressAtTemperature(Compute):
    """compute.rs:439-455"""

    def __init__(self, temperature):
        self.temperature = temperature

    def compute(self, system):
        assert self.temperature >= 0.0, "assertion failed: self.temperature >= 0.0"
        if system.cell.is_infinite():
            raise ValueError("Can not compute stress for infinite cell")
        virial = Virial().compute(system)
        volume = system.volume()
        dof = float(system.degrees_of_freedom())
        kinetic = dof / 3.0 * K_BOLTZMANN * self.temperature * np.eye(3)
        return (kinetic + virial) / volume


class Stress(Compute):
    """compute.rs:465-480"""

    def compute(self, system):
        if system.cell.is_infinite():
            raise ValueError("Can not compute stress for infinite cell")
        kinetic = device_for(system, velocities=True).kinetic_tensor()
        volume = system.volume()
        virial = Virial().compute(system)
        return (kinetic + virial) / volume


class EnergyEvaluator:
    """``EnergyEvaluator`` (energy.rs:16-158): separate components of the potential energy."""

    def __init__(self, system):
        self.system = system

    def _terms(self, parts):
        return device_for(self.system).compute(energy=True, parts=parts).energy

    def pairs(self):
        return self._terms(_ffi.PART_PAIRS).pairs

    def pairs_tail(self):
        if self.system.cell.is_infinite():
            return 0.0
        return self._terms(_ffi.PART_PAIRS).pairs_tail

    def bonds(self):
        return self._terms(_ffi.PART_BONDED).bonds

    def angles(self):
        return self._terms(_ffi.PART_BONDED).angles

    def dihedrals(self):
        return self._terms(_ffi.PART_BONDED).dihedrals

    def coulomb(self):
        terms = self._terms(_ffi.PART_COULOMB)
        return terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace

    def global_(self):
        return 0.0
