"""ctypes binding of the CPU oracle (``oracle/lumol_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  ``OracleSystem`` flattens a host-side ``lumol_b200.System`` into the
oracle's own ``orc_system``; the topology (angles, dihedrals, bond-distance matrices) is rebuilt by the
oracle's own ``orc_bonding_rebuild`` from the bond list, not taken from the product's tables.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.path.join(_HERE, "_build", "liblumol_oracle.so")

POT_NULL, POT_LJ, POT_HARMONIC, POT_BUCKINGHAM, POT_BMH, POT_MORSE, POT_GAUSSIAN, POT_MIE = range(8)
POT_COSINE_HARMONIC, POT_TORSION, POT_ABSENT = 8, 9, -1
COULOMB_NONE, COULOMB_EWALD, COULOMB_WOLF = 0, 1, 2

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


class OrcPotential(ctypes.Structure):
    _fields_ = [("pot", ctypes.c_int32), ("_pad", ctypes.c_int32), ("p", ctypes.c_double * 5)]


class OrcPair(ctypes.Structure):
    _fields_ = [
        ("potential", OrcPotential),
        ("cutoff", ctypes.c_double),
        ("shifted", ctypes.c_int32),
        ("tail", ctypes.c_int32),
        ("restriction", ctypes.c_int32),
        ("table_n", ctypes.c_int32),
        ("scale14", ctypes.c_double),
        ("table_max", ctypes.c_double),
        ("table_energy", _dp),
        ("table_force", _dp),
    ]


class OrcSystem(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int64),
        ("position", _dp),
        ("velocity", _dp),
        ("mass", _dp),
        ("charge", _dp),
        ("kind", ctypes.POINTER(ctypes.c_uint32)),
        ("cell", ctypes.c_double * 9),
        ("shape", ctypes.c_int32),
        ("nkinds", ctypes.c_int32),
        ("nmol", ctypes.c_int64),
        ("mol_start", _ip),
        ("molid", _ip),
        ("bond_dist", ctypes.POINTER(ctypes.c_uint8)),
        ("bond_dist_off", _ip),
        ("pairs", ctypes.POINTER(OrcPair)),
        ("nbonds", ctypes.c_int64),
        ("bonds", _ip),
        ("bond_pot", ctypes.POINTER(OrcPotential)),
        ("nangles", ctypes.c_int64),
        ("angles", _ip),
        ("angle_pot", ctypes.POINTER(OrcPotential)),
        ("ndihedrals", ctypes.c_int64),
        ("dihedrals", _ip),
        ("dihedral_pot", ctypes.POINTER(OrcPotential)),
        ("coulomb", ctypes.c_int32),
        ("coulomb_restriction", ctypes.c_int32),
        ("coulomb_scale14", ctypes.c_double),
        ("rc", ctypes.c_double),
        ("alpha", ctypes.c_double),
        ("kmax", ctypes.c_int32),
        ("dof_mode", ctypes.c_int32),
        ("dof_frozen", ctypes.c_int64),
    ]


class OrcEnergyTerms(ctypes.Structure):
    _fields_ = [(name, ctypes.c_double) for name in
                ("pairs", "pairs_tail", "bonds", "angles", "dihedrals", "coulomb_real", "coulomb_self", "coulomb_kspace")]


def build():
    """Compile the oracle with the committed Makefile (gcc -O2 -fopenmp -ffp-contract=off)."""
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


_library = None


def library():
    global _library
    if _library is None:
        if not os.path.exists(LIBRARY_PATH):
            build()
        lib = ctypes.CDLL(LIBRARY_PATH)
        sp = ctypes.POINTER(OrcSystem)
        pp = ctypes.POINTER(OrcPotential)
        prp = ctypes.POINTER(OrcPair)
        d = ctypes.c_double
        signatures = {
            "orc_potential_energy": (d, [pp, d]),
            "orc_potential_force": (d, [pp, d]),
            "orc_potential_tail_energy": (d, [pp, d]),
            "orc_potential_tail_virial": (d, [pp, d]),
            "orc_mie_prefactor": (d, [d, d, d]),
            "orc_potential_virial": (None, [pp, _dp, _dp]),
            "orc_table_build": (None, [pp, ctypes.c_int32, d, _dp, _dp]),
            "orc_table_energy": (d, [_dp, ctypes.c_int32, d, d]),
            "orc_pair_energy": (d, [prp, d]),
            "orc_pair_force": (d, [prp, d]),
            "orc_pair_virial": (None, [prp, _dp, _dp]),
            "orc_pair_tail_energy": (d, [prp]),
            "orc_pair_tail_virial": (d, [prp]),
            "orc_restriction_information": (None, [ctypes.c_int32, d, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), _dp]),
            "orc_bond_path": (ctypes.c_int32, [sp, ctypes.c_int64, ctypes.c_int64]),
            "orc_bonding_rebuild": (None, [ctypes.c_int64, ctypes.c_int64, _ip, _ip, _ip, _ip, _ip, ctypes.POINTER(ctypes.c_uint8)]),
            "orc_matrix_inverse": (None, [_dp, _dp]),
            "orc_vector_image": (None, [_dp, ctypes.c_int32, _dp]),
            "orc_wrap_vector": (None, [_dp, ctypes.c_int32, _dp]),
            "orc_cell_volume": (d, [_dp, ctypes.c_int32]),
            "orc_cell_lengths": (None, [_dp, ctypes.c_int32, _dp]),
            "orc_k_vector": (None, [_dp, _dp, _dp]),
            "orc_angle_and_derivatives": (d, [_dp, ctypes.c_int32, _dp, _dp, _dp, _dp, _dp, _dp]),
            "orc_dihedral_and_derivatives": (d, [_dp, ctypes.c_int32, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
            "orc_set_threads": (None, [ctypes.c_int32]),
            "orc_get_threads": (ctypes.c_int32, []),
            "orc_pair_forces": (None, [sp, _dp]),
            "orc_bonded_forces": (None, [sp, _dp]),
            "orc_pair_forces_sample": (ctypes.c_int64, [sp, ctypes.c_int64, _ip, _dp]),
            "orc_pair_forces_rows": (None, [sp, ctypes.c_int64, _ip, _dp]),
            "orc_ewald_real_forces_rows": (None, [sp, ctypes.c_int64, _ip, _dp]),
            "orc_coulomb_forces": (None, [sp, _dp]),
            "orc_forces": (None, [sp, _dp]),
            "orc_pairs_energy": (d, [sp]),
            "orc_pairs_tail_energy": (d, [sp]),
            "orc_energy_terms_compute": (None, [sp, ctypes.POINTER(OrcEnergyTerms)]),
            "orc_potential_energy_total": (d, [sp]),
            "orc_pair_atomic_virial": (None, [sp, _dp]),
            "orc_tail_virial": (None, [sp, _dp]),
            "orc_bond_virial": (None, [sp, _dp]),
            "orc_coulomb_atomic_virial": (None, [sp, _dp]),
            "orc_coulomb_molecular_virial": (None, [sp, _dp]),
            "orc_atomic_virial": (None, [sp, _dp]),
            "orc_molecular_virial": (None, [sp, _dp]),
            "orc_kinetic_energy": (d, [sp]),
            "orc_degrees_of_freedom": (ctypes.c_int64, [sp]),
            "orc_temperature": (d, [sp]),
            "orc_pressure_at_temperature": (d, [sp, d]),
            "orc_pressure": (d, [sp]),
            "orc_stress_at_temperature": (None, [sp, d, _dp]),
            "orc_stress": (None, [sp, _dp]),
            "orc_ewald_factors": (ctypes.c_int64, [sp, _dp, ctypes.c_int64, _ip, _dp, _dp, _dp]),
            "orc_ewald_real_energy": (d, [sp]),
            "orc_ewald_self_energy": (d, [sp]),
            "orc_ewald_kspace_energy": (d, [sp]),
            "orc_ewald_real_forces": (None, [sp, _dp]),
            "orc_ewald_kspace_forces": (None, [sp, _dp]),
            "orc_ewald_real_atomic_virial": (None, [sp, _dp]),
            "orc_ewald_kspace_atomic_virial": (None, [sp, _dp]),
            "orc_ewald_rho": (None, [sp, ctypes.c_int64, _dp]),
            "orc_ewald_with_accuracy": (None, [sp, d, d, _dp, ctypes.POINTER(ctypes.c_int32)]),
            "orc_wolf_energy": (d, [sp]),
            "orc_wolf_forces": (None, [sp, _dp]),
            "orc_wolf_atomic_virial": (None, [sp, _dp]),
            "orc_velocity_verlet_step": (None, [sp, _dp, _dp, _dp, d]),
            "orc_verlet_setup": (None, [sp, _dp, d]),
            "orc_verlet_step": (None, [sp, _dp, _dp, _dp, d]),
            "orc_leapfrog_step": (None, [sp, _dp, _dp, _dp, d]),
            "orc_scale_velocities": (None, [ctypes.c_int64, _dp, d]),
            "orc_berendsen_thermostat_factor": (d, [d, d, d]),
            "orc_rescale_thermostat_factor": (d, [d, d]),
            "orc_remove_translation": (None, [ctypes.c_int64, _dp, _dp]),
            "orc_remove_rotation": (None, [ctypes.c_int64, _dp, _dp, _dp]),
            "orc_rewrap": (None, [sp, _dp]),
            "orc_berendsen_barostat_step": (ctypes.c_int, [sp, _dp, _dp, _dp, d, d, d, _dp, d]),
            "orc_aniso_berendsen_barostat_step": (ctypes.c_int, [sp, _dp, _dp, _dp, d, _dp, d, _dp, d]),
            "orc_move_molecule_pairs_cost": (d, [sp, ctypes.c_int64, _dp]),
            "orc_ewald_real_move_molecule_cost": (d, [sp, ctypes.c_int64, _dp]),
            "orc_ewald_kspace_move_molecule_cost": (d, [sp, ctypes.c_int64, _dp, ctypes.c_int64, _dp]),
            "orc_wolf_move_molecule_cost": (d, [sp, ctypes.c_int64, _dp]),
            "orc_move_molecule_cost": (None, [sp, ctypes.c_int64, _dp, _dp]),
            "orc_move_all_molecules_cost": (None, [sp, sp, _dp]),
        }
        for name, (restype, argtypes) in signatures.items():
            function = getattr(lib, name)
            function.restype = restype
            function.argtypes = argtypes
        _library = lib
    return _library


def dptr(array):
    return array.ctypes.data_as(_dp)


def iptr(array):
    return array.ctypes.data_as(_ip)


def potential_record(potential):
    """OrcPotential from a host-side potential object (matching on the class name, not on product enums)."""
    record = OrcPotential()
    if potential is None:
        record.pot = POT_ABSENT
        return record
    name = type(potential).__name__
    table = {
        "NullPotential": (POT_NULL, ()),
        "LennardJones": (POT_LJ, ("sigma", "epsilon")),
        "Harmonic": (POT_HARMONIC, ("k", "x0")),
        "Buckingham": (POT_BUCKINGHAM, ("a", "c", "rho")),
        "BornMayerHuggins": (POT_BMH, ("a", "c", "d", "sigma", "rho")),
        "Morse": (POT_MORSE, ("a", "x0", "depth")),
        "Gaussian": (POT_GAUSSIAN, ("a", "b")),
        "Mie": (POT_MIE, ("sigma", "n", "m", "prefac")),
        "CosineHarmonic": (POT_COSINE_HARMONIC, ("k", "cos_x0")),
        "Torsion": (POT_TORSION, ("k", "delta", "n")),
    }
    if name not in table:
        raise TypeError(f"the oracle has no closed form for {name}")
    record.pot, fields = table[name]
    for k, field in enumerate(fields):
        record.p[k] = float(getattr(potential, field))
    return record


_RESTRICTION = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6}


class OracleSystem:
    """Flattens a ``lumol_b200.System`` for the oracle and keeps the backing arrays alive."""

    def __init__(self, system, coulomb="system"):
        lib = library()
        self.lib = lib
        n = system.size()
        self.n = n
        self.position = np.ascontiguousarray(system.positions, dtype=np.float64).reshape(n, 3).copy()
        self.velocity = np.ascontiguousarray(system.velocities, dtype=np.float64).reshape(n, 3).copy()
        self.mass = np.ascontiguousarray(system.masses, dtype=np.float64).copy()
        self.charge = np.ascontiguousarray(system.charges, dtype=np.float64).copy()
        self.kind = np.ascontiguousarray(system.kinds, dtype=np.uint32).copy()
        s = OrcSystem()
        s.n = n
        s.position, s.velocity, s.mass, s.charge = dptr(self.position), dptr(self.velocity), dptr(self.mass), dptr(self.charge)
        s.kind = self.kind.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        matrix = system.cell.matrix().reshape(-1)
        for k in range(9):
            s.cell[k] = matrix[k]
        s.shape = system.cell.shape()
        nkinds = max(len(system._kind_names), int(self.kind.max()) + 1 if n else 0)
        s.nkinds = nkinds

        # topology, rebuilt by the oracle from the bond lists
        bondings = system.bondings
        nmol = len(bondings)
        self.mol_start = np.zeros(nmol + 1, dtype=np.int64)
        self.molid = np.zeros(max(n, 1), dtype=np.int64)
        self.bd_off = np.zeros(max(nmol, 1), dtype=np.int64)
        chunks, bonds, angles, dihedrals = [], [], [], []
        cursor = 0
        cache = {}
        if nmol == n:
            # only free atoms: no topology to rebuild
            self.mol_start[:n] = np.arange(n)
            self.molid[:n] = np.arange(n)
            chunks.append(np.full(1, 8, dtype=np.uint8))
            bondings = []
        for m, bonding in enumerate(bondings):
            self.mol_start[m] = bonding.start
            self.molid[bonding.start:bonding.end] = m
            size = bonding.end - bonding.start
            local = np.array(sorted((i - bonding.start, j - bonding.start) for (i, j) in bonding.bonds), dtype=np.int64).reshape(-1, 2)
            key = (size, local.tobytes())
            if key not in cache:
                na, nd = ctypes.c_int64(), ctypes.c_int64()
                distances = np.zeros(size * size, dtype=np.uint8)
                count = ctypes.c_int64(len(local))
                lib.orc_bonding_rebuild(size, count, iptr(local), ctypes.byref(na), None, ctypes.byref(nd), None, None)
                a = np.zeros((max(na.value, 1), 3), dtype=np.int64)
                d = np.zeros((max(nd.value, 1), 4), dtype=np.int64)
                lib.orc_bonding_rebuild(size, count, iptr(local), ctypes.byref(na), iptr(a), ctypes.byref(nd), iptr(d),
                                        distances.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
                cache[key] = (cursor, local, a[: na.value], d[: nd.value])
                chunks.append(distances)
                cursor += size * size
            offset, local, a, d = cache[key]
            self.bd_off[m] = offset
            bonds.append(local + bonding.start)
            angles.append(a + bonding.start)
            dihedrals.append(d + bonding.start)
        self.mol_start[nmol] = n
        self.bond_dist = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.uint8)
        s.nmol = nmol
        s.mol_start, s.molid, s.bond_dist_off = iptr(self.mol_start), iptr(self.molid), iptr(self.bd_off)
        s.bond_dist = self.bond_dist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))

        # pair table
        self.pairs = (OrcPair * max(nkinds * nkinds, 1))()
        self._tables = []
        for k in range(nkinds * nkinds):
            self.pairs[k].potential.pot = POT_ABSENT
        for (a, b), interaction in system.pairs.items():
            record = OrcPair()
            potential = interaction.potential
            if type(potential).__name__ == "TableComputation":
                inner = potential_record(potential.potential)
                record.potential = inner
                record.table_n = potential.size
                record.table_max = potential.cutoff
                energy = np.zeros(potential.size)
                force = np.zeros(potential.size)
                lib.orc_table_build(ctypes.byref(inner), potential.size, potential.cutoff, dptr(energy), dptr(force))
                self._tables.append((energy, force))
                record.table_energy, record.table_force = dptr(energy), dptr(force)
            else:
                record.potential = potential_record(potential)
            record.cutoff = interaction.cutoff()
            record.shifted = 0 if interaction.shift is None else 1
            record.tail = 1 if interaction.tail else 0
            record.restriction = _RESTRICTION[interaction.restriction().kind]
            record.scale14 = interaction.restriction().scaling
            self.pairs[a * nkinds + b] = record
            self.pairs[b * nkinds + a] = record
        s.pairs = self.pairs

        # bonded terms (only when the system has bonded potentials, like the reference's loops over
        # molecule.bonds() which find no potential otherwise)
        def flatten(groups, width):
            if not groups:
                return np.zeros((0, width), dtype=np.int64)
            return np.ascontiguousarray(np.concatenate([g.reshape(-1, width) for g in groups]), dtype=np.int64)

        self.bonds = flatten(bonds, 2)
        self.angles = flatten(angles, 3)
        self.dihedrals = flatten(dihedrals, 4)
        self.bond_pot = (OrcPotential * max(len(self.bonds), 1))(*[potential_record(system.bond_potential(int(i), int(j))) for (i, j) in self.bonds])
        self.angle_pot = (OrcPotential * max(len(self.angles), 1))(*[potential_record(system.angle_potential(int(i), int(j), int(k))) for (i, j, k) in self.angles])
        self.dihedral_pot = (OrcPotential * max(len(self.dihedrals), 1))(
            *[potential_record(system.dihedral_potential(int(i), int(j), int(k), int(m))) for (i, j, k, m) in self.dihedrals]
        )
        s.nbonds, s.bonds, s.bond_pot = len(self.bonds), iptr(self.bonds), self.bond_pot
        s.nangles, s.angles, s.angle_pot = len(self.angles), iptr(self.angles), self.angle_pot
        s.ndihedrals, s.dihedrals, s.dihedral_pot = len(self.dihedrals), iptr(self.dihedrals), self.dihedral_pot

        potential = system.coulomb if isinstance(coulomb, str) else coulomb
        s.coulomb = COULOMB_NONE
        s.coulomb_scale14 = 1.0
        if potential is not None:
            if type(potential).__name__ == "SharedEwald":
                s.coulomb = COULOMB_EWALD
                s.rc, s.alpha, s.kmax = potential.ewald.rc, potential.ewald.alpha, potential.ewald.kmax
                s.coulomb_restriction = _RESTRICTION[potential.ewald.restriction.kind]
                s.coulomb_scale14 = potential.ewald.restriction.scaling
            else:
                s.coulomb = COULOMB_WOLF
                s.rc = potential.cutoff()
                s.coulomb_restriction = _RESTRICTION[potential.restriction.kind]
                s.coulomb_scale14 = potential.restriction.scaling
        mode, frozen = system.simulated_degrees_of_freedom
        s.dof_mode = 1 if mode == "molecules" else 0
        s.dof_frozen = frozen
        self.s = s
        self.ref = ctypes.byref(s)

    # ---- estimators ------------------------------------------------------------------------------------
    def _vector(self, function):
        out = np.zeros((self.n, 3))
        function(self.ref, dptr(out))
        return out

    def _matrix(self, function):
        out = np.zeros((3, 3))
        function(self.ref, dptr(out))
        return out

    def forces(self):
        return self._vector(self.lib.orc_forces)

    def pair_forces(self):
        return self._vector(self.lib.orc_pair_forces)

    def pair_forces_rows(self, rows):
        """Total pair force on the atoms ``rows`` (every j != i): boxes too large for the full O(N^2) loop."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros((len(rows), 3))
        self.lib.orc_pair_forces_rows(self.ref, len(rows), iptr(rows), dptr(out))
        return out

    def ewald_real_forces_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros((len(rows), 3))
        self.lib.orc_ewald_real_forces_rows(self.ref, len(rows), iptr(rows), dptr(out))
        return out

    def coulomb_forces(self):
        return self._vector(self.lib.orc_coulomb_forces)

    def energy_terms(self):
        terms = OrcEnergyTerms()
        self.lib.orc_energy_terms_compute(self.ref, ctypes.byref(terms))
        return terms

    def potential_energy(self):
        return self.lib.orc_potential_energy_total(self.ref)

    def atomic_virial(self):
        return self._matrix(self.lib.orc_atomic_virial)

    def molecular_virial(self):
        return self._matrix(self.lib.orc_molecular_virial)

    def coulomb_atomic_virial(self):
        return self._matrix(self.lib.orc_coulomb_atomic_virial)

    def coulomb_molecular_virial(self):
        return self._matrix(self.lib.orc_coulomb_molecular_virial)

    def kinetic_energy(self):
        return self.lib.orc_kinetic_energy(self.ref)

    def temperature(self):
        return self.lib.orc_temperature(self.ref)

    def pressure(self):
        return self.lib.orc_pressure(self.ref)

    def stress(self):
        return self._matrix(self.lib.orc_stress)

    def ewald_factors(self):
        kmax2 = ctypes.c_double()
        nk = self.lib.orc_ewald_factors(self.ref, ctypes.byref(kmax2), 0, None, None, None, None)
        index = np.zeros((nk, 3), dtype=np.int64)
        energy = np.zeros(nk)
        field = np.zeros((nk, 3))
        virial = np.zeros((nk, 9))
        self.lib.orc_ewald_factors(self.ref, ctypes.byref(kmax2), nk, iptr(index), dptr(energy), dptr(field), dptr(virial))
        return kmax2.value, index, energy, field, virial

    def ewald_rho(self, nk):
        rho = np.zeros((nk, 2))
        self.lib.orc_ewald_rho(self.ref, nk, dptr(rho))
        return rho

    # ---- Monte Carlo energy cache (sys/cache.rs) ---------------------------------------------------------
    def move_molecule_cost(self, molecule, new_positions):
        """(pairs, coulomb real space or Wolf, coulomb k-space) terms of ``EnergyCache::move_molecule_cost``."""
        positions = np.ascontiguousarray(new_positions, dtype=np.float64)
        out = np.zeros(3)
        self.lib.orc_move_molecule_cost(self.ref, molecule, dptr(positions), dptr(out))
        return out

    def ewald_kspace_move_molecule_cost(self, molecule, new_positions, nk):
        positions = np.ascontiguousarray(new_positions, dtype=np.float64)
        delta = np.zeros((nk, 2))
        cost = self.lib.orc_ewald_kspace_move_molecule_cost(self.ref, molecule, dptr(positions), nk, dptr(delta))
        return cost, delta

    def move_all_molecules_cost(self, after):
        """(pairs, tail, coulomb) terms of ``EnergyCache::move_all_molecules_cost``; ``self`` is the cached system."""
        out = np.zeros(3)
        self.lib.orc_move_all_molecules_cost(self.ref, after.ref, dptr(out))
        return out
