"""Host-side mirror of ``lumol_sim::md``: integrators, thermostats, controls and the MD propagator.

The reference steps positions and velocities in host loops and calls ``system.forces()`` once per step
(lumol-sim/src/md/integrators.rs:44-69, molecular_dynamics.rs:66-76).  Here the whole step runs on the
device: ``MolecularDynamics.run`` hands the step count to ``lumol_cuda_md_run`` and positions,
velocities and forces never leave HBM until the run ends.
"""

import ctypes
import math

import numpy as np

from . import _ffi
from .consts import K_BOLTZMANN
from .device import device_for


class Integrator:
    """``Integrator`` trait (integrators.rs:10-17)."""

    KIND = None

    def __init__(self, timestep):
        self.timestep = float(timestep)


class VelocityVerlet(Integrator):
    """integrators.rs:24-70"""

    KIND = _ffi.INTEGRATOR_VELOCITY_VERLET


class Verlet(Integrator):
    """integrators.rs:78-123"""

    KIND = _ffi.INTEGRATOR_VERLET


class LeapFrog(Integrator):
    """integrators.rs:131-169"""

    KIND = _ffi.INTEGRATOR_LEAP_FROG


class BerendsenBarostat(Integrator):
    """integrators.rs:176-255: velocity-Verlet with isotropic Berendsen scaling of positions and cell.  ``tau`` is the
    barostat time scale in units of the timestep."""

    KIND = _ffi.INTEGRATOR_BERENDSEN_BAROSTAT

    def __init__(self, timestep, pressure, tau):
        super().__init__(timestep)
        self.tau = float(tau)
        self.target = np.zeros((3, 3))
        self.target[0, 0] = float(pressure)


class AnisoBerendsenBarostat(Integrator):
    """integrators.rs:260-341: the same with a target stress matrix and a scaling matrix."""

    KIND = _ffi.INTEGRATOR_ANISO_BERENDSEN_BAROSTAT

    def __init__(self, timestep, stress, tau):
        super().__init__(timestep)
        self.tau = float(tau)
        self.target = np.ascontiguousarray(np.array(stress, dtype=np.float64).reshape(3, 3))

    @classmethod
    def hydrostatic(cls, timestep, pressure, tau):
        """integrators.rs:288-290"""
        return cls(timestep, float(pressure) * np.eye(3), tau)


class Thermostat:
    """``Thermostat`` trait (thermostats.rs:13-24)."""

    KIND = _ffi.THERMOSTAT_NONE
    temperature = 0.0
    parameter = 0.0


class RescaleThermostat(Thermostat):
    """thermostats.rs:36-73"""

    KIND = _ffi.THERMOSTAT_RESCALE

    def __init__(self, temperature, tolerance=None):
        if not temperature >= 0.0:
            raise ValueError("The temperature must be positive in thermostats.")
        self.temperature = float(temperature)
        self.parameter = 0.05 * temperature if tolerance is None else float(tolerance)

    @classmethod
    def with_tolerance(cls, temperature, tol):
        return cls(temperature, tol)


class BerendsenThermostat(Thermostat):
    """thermostats.rs:88-120"""

    KIND = _ffi.THERMOSTAT_BERENDSEN

    def __init__(self, temperature, tau):
        if not temperature >= 0.0:
            raise ValueError("The temperature must be positive in thermostats.")
        if not tau >= 1.0:
            raise ValueError("The timestep must be larger than 1 in berendsen thermostat.")
        self.temperature = float(temperature)
        self.parameter = float(tau)


class CSVRThermostat(Thermostat):
    """thermostats.rs:134-211.  The stochastic terms (one Gaussian and one chi-squared sum per step,
    thermostats.rs:175-193) come from a host RNG and are uploaded as a block of (gauss, wiener) pairs;
    the velocity scaling factor itself is computed on the device."""

    KIND = _ffi.THERMOSTAT_CSVR

    def __init__(self, temperature, tau, seed=0xEBA8E429):
        if not temperature >= 0.0:
            raise ValueError("The temperature must be positive in thermostats.")
        if not tau >= 1.0:
            raise ValueError("The timestep must be larger than 1 in CSVR thermostat.")
        self.temperature = float(temperature)
        self.parameter = float(tau)
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def sum_noises(self, n, steps):
        """``CSVRThermostat::sum_noises`` (thermostats.rs:175-193) for ``steps`` steps at once."""
        noise = np.zeros((steps, 2))
        if n == 0:
            return noise
        gauss = self.rng.standard_normal(steps)
        noise[:, 0] = gauss
        if n == 1:
            noise[:, 1] = gauss * gauss
        elif n % 2 == 0:
            noise[:, 1] = 2.0 * self.rng.gamma(n // 2, 1.0, steps)
        else:
            noise[:, 1] = 2.0 * self.rng.gamma((n - 1) // 2, 1.0, steps) + gauss * gauss
        return noise


class Control:
    """``Control`` trait (controls.rs:13-24)."""

    FLAG = 0


class RemoveTranslation(Control):
    """controls.rs:27-41"""

    FLAG = _ffi.CONTROL_REMOVE_TRANSLATION

    def control(self, system):
        device = device_for(system, velocities=True)
        _ffi.check(device.ctx, device.lib.lumol_cuda_remove_translation(device.ctx))
        device.download(system, positions=False)


class RemoveRotation(Control):
    """controls.rs:44-73"""

    FLAG = _ffi.CONTROL_REMOVE_ROTATION

    def control(self, system):
        device = device_for(system, velocities=True)
        _ffi.check(device.ctx, device.lib.lumol_cuda_remove_rotation(device.ctx))
        device.download(system, positions=False)


class Rewrap(Control):
    """controls.rs:76-87"""

    FLAG = _ffi.CONTROL_REWRAP

    def control(self, system):
        device = device_for(system, velocities=True)
        _ffi.check(device.ctx, device.lib.lumol_cuda_rewrap(device.ctx))
        device.download(system, velocities=False)


def scale(system, temperature):
    """``velocities::scale`` (velocities.rs:16-22)."""
    instant = system.temperature()
    factor = math.sqrt(temperature / instant)
    device = device_for(system, velocities=True)
    _ffi.check(device.ctx, device.lib.lumol_cuda_scale_velocities(device.ctx, factor))
    device.download(system, positions=False)


class MolecularDynamics:
    """``MolecularDynamics`` (molecular_dynamics.rs:14-76): integrator, optional thermostat, controls."""

    def __init__(self, timestep_or_integrator):
        if isinstance(timestep_or_integrator, Integrator):
            self.integrator = timestep_or_integrator
        else:
            self.integrator = VelocityVerlet(timestep_or_integrator)
        self.thermostat = None
        self.controls = []
        self._ready_for = None

    @classmethod
    def from_integrator(cls, integrator):
        return cls(integrator)

    def set_thermostat(self, thermostat):
        self.thermostat = thermostat

    def add_control(self, control):
        self.controls.append(control)

    def setup(self, system):
        """``Propagator::setup`` (molecular_dynamics.rs:59-64): uploads x, v and runs ``Integrator::setup``."""
        device = device_for(system, velocities=True)
        lib, ctx = device.lib, device.ctx
        _ffi.check(ctx, lib.lumol_cuda_md_setup(ctx, self.integrator.KIND, self.integrator.timestep))
        if isinstance(self.integrator, (BerendsenBarostat, AnisoBerendsenBarostat)):
            target = np.ascontiguousarray(self.integrator.target, dtype=np.float64)
            _ffi.check(ctx, lib.lumol_cuda_md_set_barostat(ctx, _ffi.as_double_pointer(target), self.integrator.tau))
        thermostat = self.thermostat if self.thermostat is not None else Thermostat()
        _ffi.check(ctx, lib.lumol_cuda_md_set_thermostat(ctx, thermostat.KIND, thermostat.temperature, thermostat.parameter))
        flags = 0
        for control in self.controls:
            flags |= control.FLAG
        _ffi.check(ctx, lib.lumol_cuda_md_set_controls(ctx, flags))
        self._ready_for = (id(system), system._version)
        return device

    def propagate(self, system, nsteps=1, download=True):
        """``nsteps`` calls of ``MolecularDynamics::propagate`` without leaving the device."""
        if self._ready_for != (id(system), system._version):
            device = self.setup(system)
        else:
            # arrays that are not resident on the device (the host edited them, or never ran) are uploaded again: like the
            # reference's propagate, the step works on the current system arrays
            device = device_for(system, velocities=True)
        lib, ctx = device.lib, device.ctx
        if isinstance(self.thermostat, CSVRThermostat):
            noise = np.ascontiguousarray(self.thermostat.sum_noises(system.degrees_of_freedom() - 1, nsteps))
            _ffi.check(ctx, lib.lumol_cuda_md_set_csvr_noise(ctx, nsteps, _ffi.as_double_pointer(noise)))
        _ffi.check(ctx, lib.lumol_cuda_md_run(ctx, nsteps))
        system.step += nsteps
        if isinstance(self.integrator, (BerendsenBarostat, AnisoBerendsenBarostat)):
            # the barostat scaled the cell on the device: bring it back (system.cell.scale_mut, integrators.rs:225, 309)
            matrix = np.zeros((3, 3))
            _ffi.check(ctx, lib.lumol_cuda_get_cell(ctx, _ffi.as_double_pointer(matrix)))
            system.cell = type(system.cell)(matrix, system.cell.shape())  # scale_mut keeps the shape (cells.rs:203-207)
            device._synced_cell = (system.cell.shape(), system.cell.matrix().tobytes())
        # positions and velocities are now newer on the device than in the host arrays
        system._resident = {"positions", "velocities"}
        if download:
            device.download(system)
            # the host arrays now equal the device state: keep the device copy authoritative
            self._ready_for = (id(system), system._version)


class Simulation:
    """``Simulation`` (lumol-sim/src/simulations.rs:52-125) for an MD propagator.  Between two output steps the
    propagator keeps positions, velocities and forces on the device; an output reads what it needs from the resident
    state (output.py), and the host arrays are refreshed once, when the run ends."""

    def __init__(self, propagator):
        self.propagator = propagator
        self.outputs = []

    def add_output(self, output):
        """simulations.rs:104-106"""
        self.add_output_with_frequency(output, 1)

    def add_output_with_frequency(self, output, frequency):
        """simulations.rs:111-113: ``output`` is written every time ``system.step`` is a multiple of ``frequency``."""
        if frequency < 1:
            raise ValueError("the output frequency must be at least 1")
        self.outputs.append((output, int(frequency)))

    def run(self, system, nsteps):
        """simulations.rs:69-101"""
        self.propagator.setup(system)
        for output, _ in self.outputs:
            output.setup(system)
        done = 0
        while done < nsteps:
            block = nsteps - done
            for _, frequency in self.outputs:
                block = min(block, frequency - system.step % frequency)
            self.propagator.propagate(system, block, download=False)
            done += block
            for output, frequency in self.outputs:
                if system.step % frequency == 0:
                    output.write(system)
        for output, _ in self.outputs:
            output.finish(system)
        system.sync_from_device()


def forces_to_host(device, n):
    forces = np.zeros((n, 3))
    _ffi.check(device.ctx, device.lib.lumol_cuda_get_forces(device.ctx, _ffi.as_double_pointer(forces)))
    return forces


__all__ = [
    "VelocityVerlet", "Verlet", "LeapFrog", "RescaleThermostat", "BerendsenThermostat", "CSVRThermostat",
    "RemoveTranslation", "RemoveRotation", "Rewrap", "BerendsenBarostat", "AnisoBerendsenBarostat", "MolecularDynamics", "Simulation", "scale", "K_BOLTZMANN",
]
