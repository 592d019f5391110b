"""Host-side mirror of ``lumol_core::sys``: ``UnitCell``, ``Particle``, ``Molecule``, ``System``.

Only what the force-evaluation path needs: the SoA particle vectors (particles.rs:31-46), the cell
(cells.rs:31-39), contiguous molecules with their bond-distance matrices (bonding.rs:16-30,
connect.rs:142-165), and the interactions map keyed by particle kinds (interactions.rs:62-77).
The property facade (``forces()``, ``potential_energy()``, ``virial()``, ...) of system.rs:249-318
routes to the CUDA library through ``lumol_b200.compute``.
"""

import math

import numpy as np

from . import _ffi
from .consts import K_BOLTZMANN

# BondDistances bits (connect.rs:142-150)
BOND_ONE, BOND_TWO, BOND_THREE, BOND_FAR = 1, 2, 4, 8

# periodic-table masses for the elements the reference's fixtures use (sys/config/mass.rs)
_MASSES = {
    "H": 1.008, "He": 4.002602, "C": 12.011, "N": 14.007, "O": 15.999, "F": 18.9984032, "Li": 6.94, "Mg": 24.305, "S": 32.06, "K": 39.0983, "Ca": 40.078, "Zn": 65.38, "Br": 79.904, "Ne": 20.1797,
    "Na": 22.98976928, "Cl": 35.45, "Ar": 39.948, "Ag": 107.8682, "Kr": 83.798, "Xe": 131.293,
}


class UnitCell:
    """``UnitCell`` (cells.rs:31-92): row-major matrix whose columns are the lattice vectors."""

    def __init__(self, matrix, shape):
        self._matrix = np.array(matrix, dtype=np.float64).reshape(3, 3)
        self._shape = shape

    @classmethod
    def infinite(cls):
        return cls(np.zeros((3, 3)), _ffi.CELL_INFINITE)

    @classmethod
    def ortho(cls, a, b, c):
        if not (a > 0.0 and b > 0.0 and c > 0.0):
            raise ValueError("Cell lengths must be positive")
        return cls(np.diag([float(a), float(b), float(c)]), _ffi.CELL_ORTHORHOMBIC)

    @classmethod
    def cubic(cls, length):
        if not length > 0.0:
            raise ValueError("Cell lengths must be positive")
        return cls.ortho(length, length, length)

    @classmethod
    def triclinic(cls, a, b, c, alpha, beta, gamma):
        """cells.rs:72-92 (angles in degrees)."""
        if not (a > 0.0 and b > 0.0 and c > 0.0):
            raise ValueError("Cell lengths must be positive")
        cos_alpha = math.cos(math.radians(alpha))
        cos_beta = math.cos(math.radians(beta))
        sin_gamma, cos_gamma = math.sin(math.radians(gamma)), math.cos(math.radians(gamma))
        b_x, b_y = b * cos_gamma, b * sin_gamma
        c_x = c * cos_beta
        c_y = c * (cos_alpha - cos_beta * cos_gamma) / sin_gamma
        c_z = math.sqrt(c * c - c_y * c_y - c_x * c_x)
        return cls([[a, b_x, c_x], [0.0, b_y, c_y], [0.0, 0.0, c_z]], _ffi.CELL_TRICLINIC)

    def shape(self):
        return self._shape

    def is_infinite(self):
        return self._shape == _ffi.CELL_INFINITE

    def matrix(self):
        return self._matrix.copy()

    def _vectors(self):
        return self._matrix[:, 0], self._matrix[:, 1], self._matrix[:, 2]

    def a(self):
        return float(np.linalg.norm(self._matrix[:, 0])) if self._shape == _ffi.CELL_TRICLINIC else float(self._matrix[0, 0])

    def b(self):
        return float(np.linalg.norm(self._matrix[:, 1])) if self._shape == _ffi.CELL_TRICLINIC else float(self._matrix[1, 1])

    def c(self):
        return float(np.linalg.norm(self._matrix[:, 2])) if self._shape == _ffi.CELL_TRICLINIC else float(self._matrix[2, 2])

    def _angle(self, u, v):
        """cells.rs:149-182: 90 for orthorhombic and infinite cells, else the angle of two cell vectors in degrees."""
        if self._shape != _ffi.CELL_TRICLINIC:
            return 90.0
        cosine = float(np.dot(u, v) / (np.linalg.norm(u) * np.linalg.norm(v)))
        return math.degrees(math.acos(max(-1.0, min(1.0, cosine))))

    def alpha(self):
        return self._angle(self._matrix[:, 1], self._matrix[:, 2])

    def beta(self):
        return self._angle(self._matrix[:, 0], self._matrix[:, 2])

    def gamma(self):
        return self._angle(self._matrix[:, 0], self._matrix[:, 1])

    def lengths(self):
        """Distances between opposite faces (cells.rs:134-146)."""
        if self.is_infinite():
            return np.array([math.inf, math.inf, math.inf])
        a, b, c = self._vectors()
        na, nb, nc = np.cross(b, c), np.cross(c, a), np.cross(a, b)
        na, nb, nc = na / np.linalg.norm(na), nb / np.linalg.norm(nb), nc / np.linalg.norm(nc)
        return np.array([abs(np.dot(na, a)), abs(np.dot(nb, b)), abs(np.dot(nc, c))])

    def volume(self):
        """cells.rs:185-199"""
        if self.is_infinite():
            return 0.0
        if self._shape == _ffi.CELL_ORTHORHOMBIC:
            return self.a() * self.b() * self.c()
        a, b, c = self._vectors()
        return float(np.dot(a, np.cross(b, c)))

    def scale(self, s):
        """``UnitCell::scale`` (cells.rs:212-220): new cell ``s * cell``."""
        if self.is_infinite():
            raise ValueError("can not scale infinite cells")
        return UnitCell(np.asarray(s, dtype=np.float64).reshape(3, 3) @ self._matrix, self._shape)

    def fractional(self, vector):
        return np.linalg.inv(self._matrix) @ np.asarray(vector, dtype=np.float64)

    def cartesian(self, fractional):
        return self._matrix @ np.asarray(fractional, dtype=np.float64)

    def __eq__(self, other):
        return isinstance(other, UnitCell) and self._shape == other._shape and np.array_equal(self._matrix, other._matrix)


class Particle:
    """``Particle`` (particles.rs:31-46)."""

    def __init__(self, name, position=(0.0, 0.0, 0.0)):
        self.name = name
        self.charge = 0.0
        self.mass = _MASSES.get(name, 0.0)
        self.position = np.array(position, dtype=np.float64)
        self.velocity = np.zeros(3)

    @classmethod
    def with_position(cls, name, position):
        return cls(name, position)


class Molecule:
    """A molecule under construction: particles plus bonds between local indices (molecules.rs)."""

    def __init__(self, particle):
        self.particles = [particle]
        self.bonds = []

    def add_particle_bonded_to(self, other, particle):
        self.particles.append(particle)
        self.bonds.append((other, len(self.particles) - 1))

    def add_bond(self, i, j):
        self.bonds.append((i, j))

    def size(self):
        return len(self.particles)


_SINGLE_FAR = np.full((1, 1), BOND_FAR, dtype=np.uint8)
_SINGLE_FAR.setflags(write=False)
_NO_CONNECTIONS = frozenset()


class Bonding:
    """``Bonding`` (bonding.rs:16-30) for one molecule spanning atoms ``[start, end)``."""

    __slots__ = ("start", "end", "bonds", "angles", "dihedrals", "distances")

    def __init__(self, start, end):
        self.start = start
        self.end = end
        self.bonds = set()  # (i, j) global indices, i < j
        self.angles = set()
        self.dihedrals = set()
        # single atoms share one read-only matrix (systems with millions of free atoms)
        self.distances = _SINGLE_FAR if end - start == 1 else np.full((end - start, end - start), BOND_FAR, dtype=np.uint8)

    @classmethod
    def single(cls, start):
        """A one-atom molecule without the three per-instance sets: millions of free atoms (liquid argon) share one
        immutable empty set; `bonds_for_update` swaps in a real set before the first bond is added."""
        self = object.__new__(cls)
        self.start = start
        self.end = start + 1
        self.bonds = self.angles = self.dihedrals = _NO_CONNECTIONS
        self.distances = _SINGLE_FAR
        return self

    def bonds_for_update(self):
        if not isinstance(self.bonds, set):
            self.bonds = set(self.bonds)
        return self.bonds

    def size(self):
        return self.end - self.start

    def translate_by(self, delta):
        self.start += delta
        self.end += delta
        self.bonds = {(i + delta, j + delta) for (i, j) in self.bonds}
        self.angles = {(i + delta, j + delta, k + delta) for (i, j, k) in self.angles}
        self.dihedrals = {(i + delta, j + delta, k + delta, m + delta) for (i, j, k, m) in self.dihedrals}

    @staticmethod
    def _angle(first, second, third):
        # Angle::new (connect.rs:50-63)
        return (min(first, third), second, max(first, third))

    @staticmethod
    def _dihedral(first, second, third, fourth):
        # Dihedral::new (connect.rs:88-103)
        if max(first, second) < max(third, fourth):
            return (first, second, third, fourth)
        return (fourth, third, second, first)

    def rebuild(self):
        """``Bonding::rebuild`` (bonding.rs:78-127): angles and dihedrals from the bond list."""
        self.angles, self.dihedrals = set(), set()
        bonds = sorted(self.bonds)
        for b1 in bonds:
            for b2 in bonds:
                if b1 == b2:
                    continue
                if b1[0] == b2[1]:
                    angle = self._angle(b2[0], b2[1], b1[1])
                elif b1[1] == b2[0]:
                    angle = self._angle(b1[0], b1[1], b2[1])
                elif b1[1] == b2[1]:
                    angle = self._angle(b1[0], b1[1], b2[0])
                elif b1[0] == b2[0]:
                    angle = self._angle(b1[1], b1[0], b2[1])
                else:
                    continue
                self.angles.add(angle)
                for b3 in bonds:
                    if b2 == b3:
                        continue
                    if angle[2] == b3[0] and angle[1] != b3[1]:
                        dihedral = self._dihedral(angle[0], angle[1], angle[2], b3[1])
                    elif angle[2] == b3[1] and angle[1] != b3[0]:
                        dihedral = self._dihedral(angle[0], angle[1], angle[2], b3[0])
                    elif angle[0] == b3[1] and angle[1] != b3[0]:
                        dihedral = self._dihedral(b3[0], angle[0], angle[1], angle[2])
                    elif angle[0] == b3[0] and angle[1] != b3[1]:
                        dihedral = self._dihedral(b3[1], angle[0], angle[1], angle[2])
                    else:
                        continue
                    self.dihedrals.add(dihedral)
        self.rebuild_connections()

    def rebuild_connections(self):
        """``Bonding::rebuild_connections`` (bonding.rs:130-155)."""
        n, first = self.size(), self.start
        self.distances = np.full((n, n), BOND_FAR, dtype=np.uint8)
        for (i, j) in self.bonds:
            self.distances[i - first, j - first] |= BOND_ONE
            self.distances[j - first, i - first] |= BOND_ONE
        for (i, _, k) in self.angles:
            self.distances[i - first, k - first] |= BOND_TWO
            self.distances[k - first, i - first] |= BOND_TWO
        for (i, _, _, m) in self.dihedrals:
            self.distances[i - first, m - first] |= BOND_THREE
            self.distances[m - first, i - first] |= BOND_THREE


def _normalize_pair(a, b):
    return (a, b) if a <= b else (b, a)


def _normalize_angle(a, b, c):
    return (a, b, c) if a <= c else (c, b, a)


def _normalize_dihedral(a, b, c, d):
    # interactions.rs:39-52: order by the outer kinds, then the inner ones
    if (a, b) <= (d, c):
        return (a, b, c, d)
    return (d, c, b, a)


class System:
    """``System`` (system.rs:41-53) with ``Configuration`` flattened in (system.rs:434-446 ``Deref``)."""

    def __init__(self, cell=None):
        self.cell = cell if cell is not None else UnitCell.infinite()
        self.names = []
        self.kinds = np.zeros(0, dtype=np.uint32)
        self.charges = np.zeros(0)
        self.masses = np.zeros(0)
        self.positions = np.zeros((0, 3))
        self.velocities = np.zeros((0, 3))
        self.bondings = []  # one Bonding per molecule, contiguous and ordered
        self.molecule_ids = np.zeros(0, dtype=np.int64)
        # Interactions (interactions.rs:62-77)
        self._kind_names = {}
        self.pairs = {}
        self.bond_potentials = {}
        self.angle_potentials = {}
        self.dihedral_potentials = {}
        self.coulomb = None
        self.simulated_degrees_of_freedom = ("particles", 0)
        self.external_temperature = None
        self.step = 0
        self._device = None
        self._version = 0  # bumped by every structural change, so the device state is rebuilt
        self._resident = set()  # "positions" / "velocities" while the device copy is newer than the host arrays

    @classmethod
    def with_cell(cls, cell):
        return cls(cell)

    # ---- particles ---------------------------------------------------------------------------------
    def size(self):
        return len(self.names)

    def is_empty(self):
        return self.size() == 0

    def get_kind(self, name):
        """interactions.rs:94-102"""
        if name not in self._kind_names:
            self._kind_names[name] = len(self._kind_names)
        return self._kind_names[name]

    def add_molecule(self, molecule):
        """configuration.rs:240-258: particles go to the end of the list."""
        if isinstance(molecule, Particle):
            molecule = Molecule(molecule)
        start = self.size()
        count = molecule.size()
        self.names += [p.name for p in molecule.particles]
        self.kinds = np.concatenate([self.kinds, np.array([self.get_kind(p.name) for p in molecule.particles], dtype=np.uint32)])
        self.charges = np.concatenate([self.charges, [p.charge for p in molecule.particles]])
        self.masses = np.concatenate([self.masses, [p.mass for p in molecule.particles]])
        self.positions = np.concatenate([self.positions, np.array([p.position for p in molecule.particles]).reshape(-1, 3)])
        self.velocities = np.concatenate([self.velocities, np.array([p.velocity for p in molecule.particles]).reshape(-1, 3)])
        bonding = Bonding(start, start + count)
        bonding.bonds = {_normalize_pair(start + i, start + j) for (i, j) in molecule.bonds}
        bonding.rebuild()
        self.molecule_ids = np.concatenate([self.molecule_ids, np.full(count, len(self.bondings), dtype=np.int64)])
        self.bondings.append(bonding)
        self._version += 1

    def add_particles(self, names, positions, charges=None, masses=None, velocities=None):
        """Bulk version of ``add_molecule(Molecule::new(particle))`` for large synthetic systems."""
        count = len(names)
        start = self.size()
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(count, 3)
        self.names += list(names)
        kind_of = {name: self.get_kind(name) for name in dict.fromkeys(names)}
        self.kinds = np.concatenate([self.kinds, np.array([kind_of[n] for n in names], dtype=np.uint32)])
        self.charges = np.concatenate([self.charges, np.zeros(count) if charges is None else np.asarray(charges, dtype=np.float64)])
        default_mass = np.array([_MASSES.get(n, 0.0) for n in names])
        self.masses = np.concatenate([self.masses, default_mass if masses is None else np.asarray(masses, dtype=np.float64)])
        self.positions = np.concatenate([self.positions, positions])
        self.velocities = np.concatenate(
            [self.velocities, np.zeros((count, 3)) if velocities is None else np.asarray(velocities, dtype=np.float64).reshape(count, 3)]
        )
        first_id = len(self.bondings)
        self.molecule_ids = np.concatenate([self.molecule_ids, np.arange(first_id, first_id + count, dtype=np.int64)])
        self.bondings += [Bonding.single(start + i) for i in range(count)]
        self._version += 1

    def molecules(self):
        return list(self.bondings)

    def molecule(self, index):
        return self.bondings[index]

    def molecule_id(self, i):
        return int(self.molecule_ids[i])

    def are_in_same_molecule(self, i, j):
        return self.molecule_ids[i] == self.molecule_ids[j]

    def _permute(self, permutation):
        """new[k] = old[permutation[k]] for every per-particle array."""
        permutation = np.asarray(permutation)
        self.names = [self.names[k] for k in permutation]
        self.kinds = self.kinds[permutation]
        self.charges = self.charges[permutation]
        self.masses = self.masses[permutation]
        self.positions = np.ascontiguousarray(self.positions[permutation])
        self.velocities = np.ascontiguousarray(self.velocities[permutation])

    def add_bond(self, particle_i, particle_j):
        """``Configuration::add_bond`` (configuration.rs:178-234): merges the two molecules, moving the
        particles of the higher-numbered one right after the other so molecules stay contiguous.
        Returns the list of ``(old, new)`` index permutations that were applied."""
        assert particle_i != particle_j
        molid_i, molid_j = int(self.molecule_ids[particle_i]), int(self.molecule_ids[particle_j])
        permutations = []
        if molid_i != molid_j:
            new_molid, old_molid = min(molid_i, molid_j), max(molid_i, molid_j)
            new_mol, old_mol = self.bondings[new_molid], self.bondings[old_molid]
            size, first, second = old_mol.size(), old_mol.start, new_mol.end
            n = self.size()
            if new_mol.end != old_mol.start:
                order = list(range(0, second)) + list(range(first, first + size)) + list(range(second, first)) + list(
                    range(first + size, n)
                )
                self._permute(order)
                for i in range(size):
                    permutations.append((first + i, second + i))
                for bonding in self.bondings[new_molid + 1 : old_molid]:
                    for i in range(bonding.start, bonding.end):
                        permutations.append((i, i + size))
                    bonding.translate_by(size)
                delta = first - second
                old_mol.translate_by(-delta)
                if molid_i == new_molid:
                    particle_j -= delta
                else:
                    particle_i -= delta
            # merge_with (bonding.rs:157-175)
            new_mol.end = old_mol.end if new_mol.end == old_mol.start else new_mol.end + size
            new_mol.bonds |= old_mol.bonds
            new_mol.angles |= old_mol.angles
            new_mol.dihedrals |= old_mol.dihedrals
            del self.bondings[old_molid]
            self.molecule_ids = np.zeros(n, dtype=np.int64)
            for index, bonding in enumerate(self.bondings):
                self.molecule_ids[bonding.start : bonding.end] = index
        bonding = self.bondings[int(self.molecule_ids[particle_i])]
        bonding.bonds_for_update().add(_normalize_pair(particle_i, particle_j))
        bonding.rebuild()
        self._version += 1
        return permutations

    def add_bonds(self, bonds):
        """Bulk ``add_bond`` for molecules that are already contiguous: ``bonds`` is an (nb, 2) array and every
        bond must join atoms that end up in one contiguous range.  Avoids the O(n) permutation per bond."""
        bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        n = self.size()
        # union-find over atoms
        parent = np.arange(n)

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for bonding in self.bondings:
            for k in range(bonding.start + 1, bonding.end):
                parent[find(k)] = find(bonding.start)
        for (i, j) in bonds:
            ri, rj = find(int(i)), find(int(j))
            if ri != rj:
                parent[max(ri, rj)] = min(ri, rj)
        roots = np.array([find(k) for k in range(n)])
        if np.any(np.diff(roots) < 0):
            raise ValueError("add_bonds needs molecules that are contiguous in the particle list")
        old_bonds = set()
        for bonding in self.bondings:
            old_bonds |= bonding.bonds
        all_bonds = old_bonds | {_normalize_pair(int(i), int(j)) for (i, j) in bonds}
        starts = np.flatnonzero(np.concatenate([[True], np.diff(roots) != 0]))
        ends = np.concatenate([starts[1:], [n]])
        self.bondings = [Bonding(int(s), int(e)) for s, e in zip(starts, ends)]
        self.molecule_ids = np.repeat(np.arange(len(starts), dtype=np.int64), ends - starts)
        for (i, j) in all_bonds:
            self.bondings[int(self.molecule_ids[i])].bonds_for_update().add((i, j))
        # molecules of the same shape share one rebuild
        cache = {}
        for bonding in self.bondings:
            key = (bonding.size(), tuple(sorted((i - bonding.start, j - bonding.start) for (i, j) in bonding.bonds)))
            if key not in cache:
                bonding.rebuild()
                cache[key] = (
                    {tuple(x - bonding.start for x in a) for a in bonding.angles},
                    {tuple(x - bonding.start for x in d) for d in bonding.dihedrals},
                    bonding.distances,
                )
            else:
                angles, dihedrals, distances = cache[key]
                bonding.angles = {tuple(x + bonding.start for x in a) for a in angles}
                bonding.dihedrals = {tuple(x + bonding.start for x in d) for d in dihedrals}
                bonding.distances = distances
        self._version += 1

    def bond_path(self, i, j):
        """``Configuration::bond_path`` (configuration.rs:124-144) as the BondDistances-style code:
        -1 none, 0 same particle, 1/2/3 bonds, 4 far."""
        if self.molecule_ids[i] != self.molecule_ids[j]:
            return -1
        if i == j:
            return 0
        bonding = self.bondings[int(self.molecule_ids[i])]
        bits = int(bonding.distances[i - bonding.start, j - bonding.start])
        if bits & BOND_ONE:
            return 1
        if bits & BOND_TWO:
            return 2
        if bits & BOND_THREE:
            return 3
        return 4

    def center_of_mass(self):
        total = float(np.sum(self.masses))
        return (self.masses[:, None] * self.positions).sum(axis=0) / total

    def clone(self):
        """``System: Clone``: per-particle arrays and topology are copied, potentials are shared, and the copy gets
        its own device state on first use."""
        import copy

        self.sync_from_device()
        device, self._device = self._device, None
        try:
            other = copy.deepcopy(self)
        finally:
            self._device = device
        return other

    def sync_from_device(self):
        """Refresh ``positions`` / ``velocities`` after device-resident steps (``propagate(..., download=False)``);
        a no-op when the host arrays are current."""
        if self._device is not None and self._resident:
            self._device.download(self, positions="positions" in self._resident, velocities="velocities" in self._resident)

    def invalidate(self):
        """Call after writing into ``charges``, ``masses`` or ``kinds`` in place (positions and velocities are
        re-read at every evaluation; the other per-particle vectors are cached on the device)."""
        self._version += 1

    # ---- interactions ----------------------------------------------------------------------------------
    def _check_cutoff(self, cutoff):
        if np.any(0.5 * self.cell.lengths() < cutoff):
            raise ValueError(
                "Can not add a potential with a cutoff bigger than half of the smallest cell length. "
                "Try increasing the cell size or decreasing the cutoff."
            )

    def set_pair_potential(self, kinds, potential):
        """system.rs:122-131"""
        self._check_cutoff(potential.cutoff())
        i, j = kinds
        self.pairs[_normalize_pair(self.get_kind(i), self.get_kind(j))] = potential
        self._version += 1

    def set_bond_potential(self, kinds, potential):
        i, j = kinds
        self.bond_potentials[_normalize_pair(self.get_kind(i), self.get_kind(j))] = potential
        self._version += 1

    def set_angle_potential(self, kinds, potential):
        i, j, k = kinds
        self.angle_potentials[_normalize_angle(self.get_kind(i), self.get_kind(j), self.get_kind(k))] = potential
        self._version += 1

    def set_dihedral_potential(self, kinds, potential):
        i, j, k, m = kinds
        key = _normalize_dihedral(self.get_kind(i), self.get_kind(j), self.get_kind(k), self.get_kind(m))
        self.dihedral_potentials[key] = potential
        self._version += 1

    def set_coulomb_potential(self, potential):
        """system.rs:159-170"""
        cutoff = potential.cutoff()
        if cutoff is not None:
            self._check_cutoff(cutoff)
        self.coulomb = potential
        self._version += 1

    def pair_potential(self, i, j):
        """system.rs:178-182"""
        return self.pairs.get(_normalize_pair(int(self.kinds[i]), int(self.kinds[j])))

    def bond_potential(self, i, j):
        return self.bond_potentials.get(_normalize_pair(int(self.kinds[i]), int(self.kinds[j])))

    def angle_potential(self, i, j, k):
        return self.angle_potentials.get(_normalize_angle(int(self.kinds[i]), int(self.kinds[j]), int(self.kinds[k])))

    def dihedral_potential(self, i, j, k, m):
        key = _normalize_dihedral(int(self.kinds[i]), int(self.kinds[j]), int(self.kinds[k]), int(self.kinds[m]))
        return self.dihedral_potentials.get(key)

    def coulomb_potential(self):
        return self.coulomb

    def maximum_cutoff(self):
        """interactions.rs:166-196"""
        cutoffs = [pair.cutoff() for pair in self.pairs.values()]
        if self.coulomb is not None and self.coulomb.cutoff() is not None:
            cutoffs.append(self.coulomb.cutoff())
        return max(cutoffs) if cutoffs else None

    # ---- properties (system.rs:249-318) ---------------------------------------------------------------------
    def simulated_temperature(self, temperature):
        self.external_temperature = temperature

    def degrees_of_freedom(self):
        mode, frozen = self.simulated_degrees_of_freedom
        if mode == "molecules":
            return 3 * len(self.bondings)
        return 3 * self.size() - frozen

    def volume(self):
        return self.cell.volume()

    def forces(self):
        from .compute import Forces

        return Forces().compute(self)

    def potential_energy(self):
        from .compute import PotentialEnergy

        return PotentialEnergy().compute(self)

    def kinetic_energy(self):
        from .compute import KineticEnergy

        return KineticEnergy().compute(self)

    def total_energy(self):
        from .compute import TotalEnergy

        return TotalEnergy().compute(self)

    def temperature(self):
        """system.rs:276-281: the external temperature wins when one is set."""
        from .compute import Temperature

        if self.external_temperature is not None:
            return self.external_temperature
        return Temperature().compute(self)

    def virial(self):
        from .compute import Virial

        return Virial().compute(self)

    def pressure(self):
        from .compute import Pressure, PressureAtTemperature

        if self.external_temperature is not None:
            return PressureAtTemperature(self.external_temperature).compute(self)
        return Pressure().compute(self)

    def stress(self):
        from .compute import Stress, StressAtTemperature

        if self.external_temperature is not None:
            return StressAtTemperature(self.external_temperature).compute(self)
        return Stress().compute(self)

    def energy_evaluator(self):
        from .compute import EnergyEvaluator

        return EnergyEvaluator(self)


def system_from_xyz(content):
    """``utils::system_from_xyz`` of the reference's test helpers (lumol-core/src/utils/mod.rs): XYZ text whose
    comment line may hold ``cell: a [b c]``, optional velocities after the positions."""
    lines = [line.strip() for line in content.strip().splitlines()]
    natoms = int(lines[0])
    comment = lines[1]
    cell = UnitCell.infinite()
    if "cell:" in comment:
        values = [float(v) for v in comment.split("cell:")[1].split()]
        cell = UnitCell.cubic(values[0]) if len(values) == 1 else UnitCell.ortho(values[0], values[1], values[2])
    system = System(cell)
    names, positions, velocities = [], [], []
    for line in lines[2 : 2 + natoms]:
        fields = line.split()
        names.append(fields[0])
        positions.append([float(v) for v in fields[1:4]])
        velocities.append([float(v) for v in fields[4:7]] if len(fields) >= 7 else [0.0, 0.0, 0.0])
    system.add_particles(names, np.array(positions), velocities=np.array(velocities))
    return system


__all__ = ["UnitCell", "Particle", "Molecule", "Bonding", "System", "system_from_xyz", "K_BOLTZMANN"]
