"""Host-side mirror of ``lumol_core::energy``: potentials, pair interactions, restrictions, coulomb.

Names, argument meaning and error messages follow the reference (paths relative to its checkout):
  potentials        lumol-core/src/energy/functions.rs
  PairInteraction   lumol-core/src/energy/pairs.rs:27-296
  TableComputation  lumol-core/src/energy/computations.rs:70-174
  PairRestriction   lumol-core/src/energy/restrictions.rs:13-114
  Ewald/SharedEwald lumol-core/src/energy/global/ewald.rs:228-351, 847-956
  Wolf              lumol-core/src/energy/global/wolf.rs:53-84, 172-331

What runs here is parameter bookkeeping only: the shift ``V(rc)``, the tail-correction scalars and
the tables of user potentials are evaluated once on the host (as the reference does at construction
time) and shipped through the C ABI; every per-pair evaluation happens in the CUDA kernels.  The
closed forms below are needed for exactly those host-side scalars and for tabulating.
"""

import math

from . import _ffi
from .consts import FOUR_PI_EPSILON_0


def _powi6(x):
    x2 = x * x
    return x2 * (x2 * x2)


class Potential:
    """``Potential`` trait (energy/mod.rs:98-103): ``energy(x)`` and ``force(x) = -dV/dx``."""

    #: lumol_cuda_potential_kind of the closed form, or None for user potentials
    KIND = None

    def energy(self, x):
        raise NotImplementedError

    def force(self, x):
        raise NotImplementedError

    def parameters(self):
        """The five parameter slots of ``lumol_cuda_potential`` / ``lumol_cuda_pair``."""
        return (0.0, 0.0, 0.0, 0.0, 0.0)


class PairPotential(Potential):
    """``PairPotential`` trait (energy/mod.rs:134-163).  Subclass it to define a custom potential: it is
    tabulated on the host with ``TableComputation`` and interpolated on the device."""

    def tail_energy(self, cutoff):
        raise NotImplementedError

    def tail_virial(self, cutoff):
        raise NotImplementedError


class NullPotential(PairPotential):
    """functions.rs:31-38"""

    KIND = _ffi.POTENTIAL_NULL

    def energy(self, x):
        return 0.0

    def force(self, x):
        return 0.0

    def tail_energy(self, cutoff):
        return 0.0

    def tail_virial(self, cutoff):
        return 0.0


class LennardJones(PairPotential):
    """functions.rs:79-107"""

    KIND = _ffi.POTENTIAL_LJ

    def __init__(self, sigma, epsilon):
        self.sigma = float(sigma)
        self.epsilon = float(epsilon)

    def parameters(self):
        return (self.sigma, self.epsilon, 0.0, 0.0, 0.0)

    def energy(self, r):
        s6 = _powi6(self.sigma / r)
        return 4.0 * self.epsilon * (s6 * s6 - s6)

    def force(self, r):
        s6 = _powi6(self.sigma / r)
        return -24.0 * self.epsilon * (s6 - 2.0 * (s6 * s6)) / r

    def tail_energy(self, cutoff):
        s3 = self.sigma * self.sigma * self.sigma
        rc3 = cutoff * cutoff * cutoff
        s9 = s3 * s3 * s3
        rc9 = rc3 * rc3 * rc3
        return 4.0 / 3.0 * self.epsilon * s3 * (1.0 / 3.0 * s9 / rc9 - s3 / rc3)

    def tail_virial(self, cutoff):
        s3 = self.sigma * self.sigma * self.sigma
        rc3 = cutoff * cutoff * cutoff
        s9 = s3 * s3 * s3
        rc9 = rc3 * rc3 * rc3
        return 8.0 * self.epsilon * s3 * (2.0 / 3.0 * s9 / rc9 - s3 / rc3)


class Harmonic(PairPotential):
    """functions.rs:135-156; also a bond, angle and dihedral potential."""

    KIND = _ffi.POTENTIAL_HARMONIC

    def __init__(self, k, x0):
        self.k = float(k)
        self.x0 = float(x0)

    def parameters(self):
        return (self.k, self.x0, 0.0, 0.0, 0.0)

    def energy(self, x):
        dx = x - self.x0
        return 0.5 * self.k * dx * dx

    def force(self, x):
        return self.k * (self.x0 - x)

    def tail_energy(self, cutoff):
        return 0.0

    def tail_virial(self, cutoff):
        return 0.0


class CosineHarmonic(Potential):
    """functions.rs:184-210 (angles and dihedrals)."""

    KIND = _ffi.POTENTIAL_COSINE_HARMONIC

    def __init__(self, k, x0):
        self.k = float(k)
        self.cos_x0 = math.cos(x0)

    def parameters(self):
        return (self.k, self.cos_x0, 0.0, 0.0, 0.0)

    def energy(self, x):
        dr = math.cos(x) - self.cos_x0
        return 0.5 * self.k * dr * dr

    def force(self, x):
        return self.k * (math.cos(x) - self.cos_x0) * math.sin(x)


class Torsion(Potential):
    """functions.rs:237-258 (dihedrals)."""

    KIND = _ffi.POTENTIAL_TORSION

    def __init__(self, k, delta, n):
        self.k = float(k)
        self.delta = float(delta)
        self.n = int(n)

    def parameters(self):
        return (self.k, self.delta, float(self.n), 0.0, 0.0)

    def energy(self, phi):
        return self.k * (1.0 + math.cos(float(self.n) * phi - self.delta))

    def force(self, phi):
        return self.k * float(self.n) * math.sin(float(self.n) * phi - self.delta)


class Buckingham(PairPotential):
    """functions.rs:286-319"""

    KIND = _ffi.POTENTIAL_BUCKINGHAM

    def __init__(self, a, c, rho):
        self.a = float(a)
        self.c = float(c)
        self.rho = float(rho)

    def parameters(self):
        return (self.a, self.c, self.rho, 0.0, 0.0)

    def energy(self, r):
        r3 = r * r * r
        r6 = r3 * r3
        return self.a * math.exp(-r / self.rho) - self.c / r6

    def force(self, r):
        r3 = r * r * r
        r7 = r3 * r3 * r
        return self.a / self.rho * math.exp(-r / self.rho) - 6.0 * self.c / r7

    def tail_energy(self, rc):
        rc2 = rc * rc
        rc3 = rc2 * rc
        exp = math.exp(-rc / self.rho)
        factor = rc2 - 2.0 * rc * self.rho + 2.0 * self.rho * self.rho
        return self.a * self.rho * exp * factor - self.c / (3.0 * rc3)

    def tail_virial(self, rc):
        rc2 = rc * rc
        rc3 = rc2 * rc
        exp = math.exp(-rc / self.rho)
        factor = rc3 + 3.0 * rc2 * self.rho + 6.0 * rc * self.rho * self.rho + 6.0 * self.rho * self.rho * self.rho
        # the "+ 8.0" is in the reference (functions.rs:317) and pinned by its test (functions.rs:710)
        return self.a * exp * factor - 20.0 * self.c / rc3 + 8.0


class BornMayerHuggins(PairPotential):
    """functions.rs:353-386"""

    KIND = _ffi.POTENTIAL_BMH

    def __init__(self, a, c, d, sigma, rho):
        self.a = float(a)
        self.c = float(c)
        self.d = float(d)
        self.sigma = float(sigma)
        self.rho = float(rho)

    def parameters(self):
        return (self.a, self.c, self.d, self.sigma, self.rho)

    def energy(self, r):
        r2 = r * r
        r6 = r2 * r2 * r2
        exp = math.exp((self.sigma - r) / self.rho)
        return self.a * exp - self.c / r6 + self.d / (r6 * r2)

    def force(self, r):
        r2 = r * r
        r7 = r2 * r2 * r2 * r
        exp = math.exp((self.sigma - r) / self.rho)
        return self.a / self.rho * exp - 6.0 * self.c / r7 + 8.0 * self.d / (r7 * r2)

    def tail_energy(self, rc):
        rc2 = rc * rc
        rc3 = rc2 * rc
        exp = math.exp((self.sigma - rc) / self.rho)
        factor = rc2 - 2.0 * rc * self.rho + 2.0 * self.rho * self.rho
        return self.a * self.rho * exp * factor - self.c / (3.0 * rc3) + self.d / (5.0 * rc2 * rc3)

    def tail_virial(self, rc):
        rc2 = rc * rc
        rc3 = rc2 * rc
        exp = math.exp((self.sigma - rc) / self.rho)
        factor = rc3 + 3.0 * rc2 * self.rho + 6.0 * rc * self.rho * self.rho + 6.0 * self.rho * self.rho * self.rho
        return self.a * exp * factor - 20.0 * self.c / rc3 + 8.0 * self.d / (5.0 * rc2 * rc3)


class Morse(PairPotential):
    """functions.rs:414-433.  ``force`` is the reference's formula ``2 D a (1 - e^2)``, which is not
    ``-dE/dr`` away from ``x0``; it is pinned by the reference's own test (functions.rs:749)."""

    KIND = _ffi.POTENTIAL_MORSE

    def __init__(self, a, x0, depth):
        self.a = float(a)
        self.x0 = float(x0)
        self.depth = float(depth)

    def parameters(self):
        return (self.a, self.x0, self.depth, 0.0, 0.0)

    def energy(self, r):
        rc = 1.0 - math.exp((self.x0 - r) * self.a)
        return self.depth * rc * rc

    def force(self, r):
        exp = math.exp((self.x0 - r) * self.a)
        return 2.0 * self.depth * (1.0 - exp * exp) * self.a

    def tail_energy(self, cutoff):
        return 0.0

    def tail_virial(self, cutoff):
        return 0.0


class Gaussian(PairPotential):
    """functions.rs:462-494"""

    KIND = _ffi.POTENTIAL_GAUSSIAN

    def __init__(self, a, b):
        if not b > 0.0:
            raise ValueError('"b" has to be positive in Gaussian potential')
        self.a = float(a)
        self.b = float(b)

    def parameters(self):
        return (self.a, self.b, 0.0, 0.0, 0.0)

    def energy(self, r):
        return -self.a * math.exp(-self.b * r * r)

    def force(self, r):
        return 2.0 * self.b * r * self.energy(r)

    def tail_energy(self, rc):
        return self.energy(rc) * rc / (2.0 * self.b) - self.a * math.sqrt(math.pi) * math.erfc(
            math.sqrt(self.b) * rc
        ) / (4.0 * math.pow(self.b, 3.0 / 2.0))

    def tail_virial(self, rc):
        return 3.0 * math.sqrt(math.pi) * self.a * math.erfc(math.sqrt(self.b) * rc) / (
            4.0 * math.pow(self.b, 3.0 / 2.0)
        ) - self.energy(rc) * rc * (2.0 * self.b * rc * rc + 3.0) / (2.0 * self.b)


class Mie(PairPotential):
    """functions.rs:538-592"""

    KIND = _ffi.POTENTIAL_MIE

    def __init__(self, sigma, epsilon, n, m):
        if not m < n:
            raise ValueError("The repulsive exponent n has to be larger than the attractive exponent m")
        self.sigma = float(sigma)
        self.n = float(n)
        self.m = float(m)
        self.prefac = n / (n - m) * math.pow(n / m, m / (n - m)) * epsilon

    def parameters(self):
        return (self.sigma, self.n, self.m, self.prefac, 0.0)

    def energy(self, r):
        sigma_r = self.sigma / r
        return self.prefac * (math.pow(sigma_r, self.n) - math.pow(sigma_r, self.m))

    def force(self, r):
        sigma_r = self.sigma / r
        return self.prefac * (self.n * math.pow(sigma_r, self.n) - self.m * math.pow(sigma_r, self.m)) / r

    def tail_energy(self, cutoff):
        if self.m <= 3.0:
            return 0.0
        sigma_rc = self.sigma / cutoff
        n_3, m_3 = self.n - 3.0, self.m - 3.0
        s3 = self.sigma * (self.sigma * self.sigma)
        return self.prefac * s3 * (math.pow(sigma_rc, n_3) / n_3 - math.pow(sigma_rc, m_3) / m_3)

    def tail_virial(self, cutoff):
        if self.m <= 3.0:
            return 0.0
        sigma_rc = self.sigma / cutoff
        n_3, m_3 = self.n - 3.0, self.m - 3.0
        s3 = self.sigma * (self.sigma * self.sigma)
        return self.prefac * s3 * (math.pow(sigma_rc, n_3) * self.n / n_3 - math.pow(sigma_rc, m_3) * self.m / m_3)


class TableComputation(PairPotential):
    """``TableComputation::new(potential, size, max)`` (computations.rs:102-119): energy and force
    sampled at ``r = i * max / size`` for ``i < size``, linearly interpolated (computations.rs:123-145)."""

    KIND = _ffi.POTENTIAL_TABLE

    def __init__(self, potential, size, max):
        self.potential = potential
        self.size = int(size)
        self.cutoff = float(max)
        self.delta = self.cutoff / float(self.size)
        self.energy_table = []
        self.force_table = []
        for i in range(self.size):
            r = float(i) * self.delta
            self.energy_table.append(_guarded(potential.energy, r))
            self.force_table.append(_guarded(potential.force, r))

    def _interpolate(self, table, r):
        quotient = r / self.delta
        if quotient != quotient or quotient < 0.0:
            bin_ = 0
        elif quotient == math.inf:
            return 0.0
        else:
            bin_ = int(math.floor(quotient))
        if bin_ < len(table) - 1:
            dx = r - float(bin_) * self.delta
            slope = (table[bin_ + 1] - table[bin_]) / self.delta
            return table[bin_] + dx * slope
        return 0.0

    def energy(self, r):
        return self._interpolate(self.energy_table, r)

    def force(self, r):
        return self._interpolate(self.force_table, r)

    def tail_energy(self, cutoff):
        return self.potential.tail_energy(cutoff)

    def tail_virial(self, cutoff):
        return self.potential.tail_virial(cutoff)


def _guarded(function, r):
    """Python raises where IEEE arithmetic gives inf/NaN (e.g. LJ at r = 0, the first table entry)."""
    try:
        return float(function(r))
    except ZeroDivisionError:
        return math.inf
    except OverflowError:
        return math.inf


class PairRestriction:
    """``PairRestriction`` (restrictions.rs:13-33)."""

    def __init__(self, kind, scaling=1.0):
        self.kind = kind
        self.scaling = float(scaling)

    def __eq__(self, other):
        return isinstance(other, PairRestriction) and self.kind == other.kind and self.scaling == other.scaling

    def __hash__(self):
        return hash((self.kind, self.scaling))

    def __repr__(self):
        return f"PairRestriction({self.kind}, {self.scaling})"

    @staticmethod
    def Scale14(scaling):
        return PairRestriction(_ffi.RESTRICTION_SCALE14, scaling)


PairRestriction.NONE = PairRestriction(_ffi.RESTRICTION_NONE)
PairRestriction.IntraMolecular = PairRestriction(_ffi.RESTRICTION_INTRA_MOLECULAR)
PairRestriction.InterMolecular = PairRestriction(_ffi.RESTRICTION_INTER_MOLECULAR)
PairRestriction.Exclude12 = PairRestriction(_ffi.RESTRICTION_EXCLUDE12)
PairRestriction.Exclude13 = PairRestriction(_ffi.RESTRICTION_EXCLUDE13)
PairRestriction.Exclude14 = PairRestriction(_ffi.RESTRICTION_EXCLUDE14)


class PairInteraction:
    """``PairInteraction`` (pairs.rs:27-38): potential + cutoff + optional shift + tail flag + restriction."""

    def __init__(self, potential, cutoff):
        """``PairInteraction::new`` (pairs.rs:57-65)."""
        self.potential = potential
        self.cutoff_ = float(cutoff)
        self.restriction_ = PairRestriction.NONE
        self.shift = None
        self.tail = False

    @classmethod
    def shifted(cls, potential, cutoff):
        """``PairInteraction::shifted`` (pairs.rs:86-95): ``shift = potential.energy(cutoff)``."""
        interaction = cls(potential, cutoff)
        interaction.shift = potential.energy(float(cutoff))
        return interaction

    def enable_tail_corrections(self):
        self.tail = True

    def restriction(self):
        return self.restriction_

    def set_restriction(self, restriction):
        self.restriction_ = restriction

    def cutoff(self):
        return self.cutoff_

    # host-side evaluation (pairs.rs:185-296); the device evaluates the same expressions per pair
    def energy(self, r):
        if r >= self.cutoff_:
            return 0.0
        energy = self.potential.energy(r)
        return energy if self.shift is None else energy - self.shift

    def force(self, r):
        if r >= self.cutoff_:
            return 0.0
        return self.potential.force(r)

    def tail_energy(self):
        return self.potential.tail_energy(self.cutoff_) if self.tail else 0.0

    def tail_virial(self):
        """Trace-less scalar: the tensor is ``tail_virial() * identity / 3`` (pairs.rs:289-296)."""
        return self.potential.tail_virial(self.cutoff_) if self.tail else 0.0


class Ewald:
    """``Ewald::new(cutoff, kmax, alpha)`` (ewald.rs:278-305)."""

    def __init__(self, cutoff, kmax, alpha=None):
        alpha = math.pi / cutoff if alpha is None else alpha
        if cutoff < 0.0:
            raise ValueError("the cutoff can not be negative in Ewald")
        if alpha < 0.0:
            raise ValueError("alpha can not be negative in Ewald")
        if kmax == 0:
            raise ValueError("kmax can not be 0 in Ewald")
        self.rc = float(cutoff)
        self.alpha = float(alpha)
        self.kmax = int(kmax)
        self.restriction = PairRestriction.NONE

    @classmethod
    def with_accuracy(cls, cutoff, accuracy, configuration):
        """``Ewald::with_accuracy`` (ewald.rs:312-351)."""
        if cutoff < 0.0:
            raise ValueError("the cutoff can not be negative in Ewald")
        if accuracy < 0.0:
            raise ValueError("accuracy can not be negative in Ewald")
        q2 = 0.0
        for charge in configuration.charges:
            q2 += charge * charge
        q2 /= FOUR_PI_EPSILON_0
        natoms = float(configuration.size())
        lengths = configuration.cell.lengths()
        alpha = accuracy * math.sqrt(natoms * cutoff * lengths[0] * lengths[1] * lengths[2]) / (2.0 * q2)
        if alpha >= 1.0:
            alpha = (1.35 - 0.15 * math.log(accuracy)) / cutoff
        else:
            alpha = math.sqrt(-math.log(alpha)) / cutoff
        min_length = min(min(lengths[0], lengths[1]), lengths[2])

        def error(kmax):
            arg = math.pi * kmax / (alpha * min_length)
            return 2.0 / math.sqrt(math.pi) * q2 * alpha / min_length / math.sqrt(kmax * natoms) * math.exp(-arg * arg)

        kmax = 1
        while error(float(kmax)) > accuracy:
            kmax += 1
        return cls(cutoff, kmax, alpha)


class _Coulomb:
    """Shared behaviour of the two ``CoulombicPotential`` implementations as ``GlobalPotential``s
    (energy/global/mod.rs:84-105): each method evaluates the coulomb part alone on the device."""

    def energy(self, configuration):
        from .device import device_for

        terms = device_for(configuration, coulomb=self).compute(energy=True, parts=_ffi.PART_COULOMB).energy
        return terms.coulomb_real + terms.coulomb_self + terms.coulomb_kspace

    def forces(self, configuration, forces):
        """Accumulates into ``forces`` like the reference (ewald.rs:897-905)."""
        from .device import device_for

        forces += device_for(configuration, coulomb=self).compute(forces=True, parts=_ffi.PART_COULOMB).forces

    def atomic_virial(self, configuration):
        from .device import device_for

        return device_for(configuration, coulomb=self).compute(virial=True, parts=_ffi.PART_COULOMB).virial

    def molecular_virial(self, configuration):
        from .device import device_for

        return device_for(configuration, coulomb=self).compute(molecular_virial=True, parts=_ffi.PART_COULOMB).virial


class SharedEwald(_Coulomb):
    """``SharedEwald`` (ewald.rs:847-956)."""

    def __init__(self, ewald):
        self.ewald = ewald

    def cutoff(self):
        return self.ewald.rc

    def set_restriction(self, restriction):
        self.ewald.restriction = restriction

    def _configure(self, lib, ctx):
        restriction = self.ewald.restriction
        if restriction.kind == _ffi.RESTRICTION_SCALE14:
            raise ValueError("Scaling restriction scheme using Ewald are not implemented")
        _ffi.check(ctx, lib.lumol_cuda_set_coulomb_ewald(ctx, self.ewald.rc, self.ewald.alpha, self.ewald.kmax, restriction.kind))


class Wolf(_Coulomb):
    """``Wolf::new(cutoff)`` (wolf.rs:68-84)."""

    def __init__(self, cutoff):
        if not cutoff > 0.0:
            raise ValueError("Got a negative cutoff in Wolf summation")
        self.cutoff_ = float(cutoff)
        self.restriction = PairRestriction.NONE

    def cutoff(self):
        return self.cutoff_

    def set_restriction(self, restriction):
        self.restriction = restriction

    def _configure(self, lib, ctx):
        _ffi.check(ctx, lib.lumol_cuda_set_coulomb_wolf(ctx, self.cutoff_, self.restriction.kind, self.restriction.scaling))
