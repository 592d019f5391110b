"""Synthetic benchmark boxes: jittered-lattice Lennard-Jones argon and SPC/E water.

Input generators only (numpy on the host); used by bench.py and by the parity tests.
"""

import numpy as np

from . import units
from .consts import K_BOLTZMANN
from .energy import Ewald, Harmonic, LennardJones, NullPotential, PairInteraction, PairRestriction, SharedEwald
from .sys import System, UnitCell


def lj_box(n_side, density=0.0213, sigma=3.4, epsilon_kjmol=1.0, cutoff=10.0, jitter=0.3, seed=0, tail=True, name="Ar"):
    """Synthetic LJ argon box: simple-cubic lattice at the given number density (atoms / A^3), positions
    jittered by U(-jitter, jitter) per axis.  ``n_side`` is an int (cubic box) or a (nx, ny, nz) tuple
    (orthorhombic box with the same lattice spacing on every axis)."""
    nx, ny, nz = (n_side, n_side, n_side) if np.isscalar(n_side) else n_side
    n = nx * ny * nz
    spacing = (1.0 / density) ** (1.0 / 3.0)
    z, y, x = np.meshgrid(np.arange(nz) * spacing, np.arange(ny) * spacing, np.arange(nx) * spacing, indexing="ij")
    positions = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1) + 0.5 * spacing
    rng = np.random.Generator(np.random.PCG64(seed))
    positions += rng.uniform(-jitter, jitter, positions.shape)
    system = System(UnitCell.ortho(nx * spacing, ny * spacing, nz * spacing))
    system.add_particles([name] * n, positions, masses=np.full(n, 39.948))
    lj = PairInteraction(LennardJones(sigma=sigma, epsilon=units.from_(epsilon_kjmol, "kJ/mol")), cutoff)
    if tail:
        lj.enable_tail_corrections()
    system.set_pair_potential((name, name), lj)
    return system


def maxwell_boltzmann(system, temperature, seed):
    """Velocities from a seeded numpy generator (any host RNG will do: not parity relevant)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sigma = np.sqrt(K_BOLTZMANN * temperature / system.masses)[:, None]
    system.velocities = rng.standard_normal((system.size(), 3)) * sigma


def spce_box(n_side, cutoff=9.0, seed=777, density=0.0334, flexible=False):
    """Synthetic SPC/E box of SURVEY section 8d: n_side^3 waters on a cubic lattice with random
    orientations; charges and O-O LJ of tests/nist-spce.rs:56-73."""
    nmol = n_side ** 3
    length = (nmol / density) ** (1.0 / 3.0)
    spacing = length / n_side
    rng = np.random.Generator(np.random.PCG64(seed))
    grid = (np.arange(n_side) + 0.5) * spacing
    z, y, x = np.meshgrid(grid, grid, grid, indexing="ij")
    centers = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    half = np.radians(109.47) / 2.0
    h1 = np.array([np.sin(half), np.cos(half), 0.0])
    h2 = np.array([-np.sin(half), np.cos(half), 0.0])
    # random rotations from normalised quaternions
    q = rng.standard_normal((nmol, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, a, b, c = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot = np.empty((nmol, 3, 3))
    rot[:, 0, 0] = 1 - 2 * (b * b + c * c); rot[:, 0, 1] = 2 * (a * b - c * w); rot[:, 0, 2] = 2 * (a * c + b * w)
    rot[:, 1, 0] = 2 * (a * b + c * w); rot[:, 1, 1] = 1 - 2 * (a * a + c * c); rot[:, 1, 2] = 2 * (b * c - a * w)
    rot[:, 2, 0] = 2 * (a * c - b * w); rot[:, 2, 1] = 2 * (b * c + a * w); rot[:, 2, 2] = 1 - 2 * (a * a + b * b)
    positions = np.empty((nmol, 3, 3))
    positions[:, 0] = centers
    positions[:, 1] = centers + rot @ h1
    positions[:, 2] = centers + rot @ h2
    names = ["O", "H", "H"] * nmol
    charges = np.tile([-0.8476, 0.4238, 0.4238], nmol)
    masses = np.tile([15.999, 1.008, 1.008], nmol)
    system = System(UnitCell.cubic(length))
    system.add_particles(names, positions.reshape(-1, 3), charges=charges, masses=masses)
    first = np.arange(0, 3 * nmol, 3)
    system.add_bonds(np.concatenate([np.stack([first, first + 1], axis=1), np.stack([first, first + 2], axis=1)]))
    lj = PairInteraction(LennardJones(epsilon=78.19743111 * K_BOLTZMANN, sigma=3.16555789), cutoff)
    lj.enable_tail_corrections()
    system.set_pair_potential(("O", "O"), lj)
    system.set_pair_potential(("O", "H"), PairInteraction(NullPotential(), cutoff))
    system.set_pair_potential(("H", "H"), PairInteraction(NullPotential(), cutoff))
    if flexible:
        kcal = units.from_(1.0, "kcal/mol")
        system.set_bond_potential(("O", "H"), Harmonic(k=1054.2 * kcal, x0=1.0))
        system.set_angle_potential(("H", "O", "H"), Harmonic(k=75.9 * kcal, x0=units.from_(109.5, "deg")))
    return system
