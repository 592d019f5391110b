"""Bridge between the host-side ``System`` and one ``lumol_cuda_context``.

This is what the Rust shim of INTEGRATION.md does inside lumol: flatten ``Configuration`` and
``Interactions`` into the arrays of the C ABI, keep the device copy in sync with a version counter, and
turn status codes back into errors.  All arithmetic happens in ``liblumol_cuda.so``.
"""

import ctypes

import numpy as np

from . import _ffi


class ComputeResult:
    def __init__(self):
        self.forces = None
        self.energy = None
        self.virial = None


def _pair_record(interaction, table_id):
    record = _ffi.Pair()
    potential = interaction.potential
    kind = potential.KIND
    if kind is None or kind == _ffi.POTENTIAL_TABLE:
        kind = _ffi.POTENTIAL_TABLE
        record.table = table_id
    else:
        for k, value in enumerate(potential.parameters()):
            record.p[k] = value
    record.potential = kind
    restriction = interaction.restriction()
    record.restriction = restriction.kind
    record.scale14 = restriction.scaling
    record.cutoff = interaction.cutoff()
    record.shift = 0.0 if interaction.shift is None else interaction.shift
    record.tail_energy = interaction.tail_energy()
    record.tail_virial = interaction.tail_virial()
    return record


class DeviceSystem:
    """Owns a ``lumol_cuda_context`` and mirrors one ``System`` into it.  ``device`` is a CUDA ordinal, or a sequence of
    ordinals for one context sharded over several devices of this process (``lumol_cuda_create_multi``)."""

    def __init__(self, device=0):
        self.lib = _ffi.library()
        self.ctx = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            devices = (ctypes.c_int32 * len(device))(*device)
            status = self.lib.lumol_cuda_create_multi(devices, len(device), ctypes.byref(self.ctx))
        else:
            status = self.lib.lumol_cuda_create(device, ctypes.byref(self.ctx))
        if status < 0:
            raise _ffi.LumolCudaError(status, _ffi.last_error(None))
        self._synced_version = None
        self._synced_cell = None
        self._n = 0

    def close(self):
        if self.ctx:
            self.lib.lumol_cuda_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        return _ffi.check(self.ctx, status)

    # ---- synchronisation -------------------------------------------------------------------------------
    def sync(self, system, coulomb="system", positions=True, velocities=False):
        """Upload what changed since the last call.  ``coulomb`` overrides the system's coulomb potential
        (used when a ``GlobalPotential`` is evaluated on its own, e.g. benches/nacl.rs:20-45)."""
        lib, ctx = self.lib, self.ctx
        cell = system.cell
        cell_key = (cell.shape(), cell.matrix().tobytes())
        if cell_key != self._synced_cell:
            matrix = np.ascontiguousarray(cell.matrix(), dtype=np.float64)
            self._check(lib.lumol_cuda_set_cell(ctx, _ffi.as_double_pointer(matrix), cell.shape()))
            self._synced_cell = cell_key

        # arrays a device-resident MD run left newer on the device than on the host (md.py): never overwritten
        # by the stale host copy; the host copy is refreshed by download()
        resident = getattr(system, "_resident", None) or set()
        full = system._version != self._synced_version
        if full:
            if resident and self._n == system.size():
                self.download(system, positions="positions" in resident, velocities="velocities" in resident)
            self._upload_structure(system)
            self._synced_version = system._version
        else:
            if positions and "positions" not in resident:
                array = np.ascontiguousarray(system.positions, dtype=np.float64)
                self._check(lib.lumol_cuda_set_positions(ctx, _ffi.as_double_pointer(array)))
            if velocities and "velocities" not in resident:
                array = np.ascontiguousarray(system.velocities, dtype=np.float64)
                self._check(lib.lumol_cuda_set_velocities(ctx, _ffi.as_double_pointer(array)))

        # the coulomb potential is cheap to describe: configure it at every call (the library keeps its
        # k-vector table as long as alpha, kmax and the cell are unchanged)
        potential = system.coulomb if isinstance(coulomb, str) else coulomb
        if potential is None:
            self._check(lib.lumol_cuda_set_coulomb_none(ctx))
        else:
            potential._configure(lib, ctx)

    def _upload_structure(self, system):
        lib, ctx = self.lib, self.ctx
        n = system.size()
        self._n = n
        position = np.ascontiguousarray(system.positions, dtype=np.float64).reshape(n, 3)
        velocity = np.ascontiguousarray(system.velocities, dtype=np.float64).reshape(n, 3)
        mass = np.ascontiguousarray(system.masses, dtype=np.float64)
        charge = np.ascontiguousarray(system.charges, dtype=np.float64)
        kind = np.ascontiguousarray(system.kinds, dtype=np.uint32)
        self._check(
            lib.lumol_cuda_set_particles(
                ctx, n, _ffi.as_double_pointer(position), _ffi.as_double_pointer(velocity), _ffi.as_double_pointer(mass),
                _ffi.as_double_pointer(charge), kind.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
            )
        )

        # molecules: contiguous ranges + bond-distance matrices, identical matrices stored once
        bondings = system.bondings
        if any(b.size() > 1 for b in bondings):
            nmol = len(bondings)
            start = np.zeros(nmol + 1, dtype=np.uint64)
            offsets = np.zeros(nmol, dtype=np.uint64)
            chunks, seen, cursor = [], {}, 0
            for m, bonding in enumerate(bondings):
                start[m] = bonding.start
                matrix = np.ascontiguousarray(bonding.distances, dtype=np.uint8)
                key = (matrix.shape[0], matrix.tobytes())
                if key not in seen:
                    seen[key] = cursor
                    chunks.append(matrix.reshape(-1))
                    cursor += matrix.size
                offsets[m] = seen[key]
            start[nmol] = n
            data = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.uint8)
            self._check(
                lib.lumol_cuda_set_molecules(
                    ctx, nmol, start.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                    offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                    data.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), data.size,
                )
            )
        else:
            self._check(lib.lumol_cuda_set_molecules(ctx, 0, None, None, None, 0))

        # pair table over every kind known to the system (interactions.rs:94-102)
        nkinds = max(len(system._kind_names), int(system.kinds.max()) + 1 if n else 0)
        self._check(lib.lumol_cuda_clear_tables(ctx))
        records = (_ffi.Pair * max(nkinds * nkinds, 1))()
        for k in range(nkinds * nkinds):
            records[k].potential = _ffi.POTENTIAL_ABSENT
        for (a, b), interaction in system.pairs.items():
            table_id = -1
            potential = interaction.potential
            if potential.KIND is None:
                raise TypeError(
                    "custom PairPotential objects must be wrapped in TableComputation(potential, size, max) "
                    "to be evaluated on the device"
                )
            if potential.KIND == _ffi.POTENTIAL_TABLE:
                energy = np.ascontiguousarray(potential.energy_table, dtype=np.float64)
                force = np.ascontiguousarray(potential.force_table, dtype=np.float64)
                table_id = self._check(
                    lib.lumol_cuda_add_table(ctx, potential.size, potential.cutoff, _ffi.as_double_pointer(energy), _ffi.as_double_pointer(force))
                )
            record = _pair_record(interaction, table_id)
            records[a * nkinds + b] = record
            records[b * nkinds + a] = record
        self._check(lib.lumol_cuda_set_pairs(ctx, nkinds, records))

        # bonded terms: explicit lists with a potential id per entry
        potentials, index_of = [], {}

        def potential_id(potential):
            if potential is None:
                return -1
            if potential.KIND is None or potential.KIND == _ffi.POTENTIAL_TABLE:
                raise TypeError("bonded potentials must be one of the built-in closed forms")
            key = (potential.KIND, tuple(potential.parameters()))
            if key not in index_of:
                index_of[key] = len(potentials)
                record = _ffi.Potential()
                record.potential = potential.KIND
                for k, value in enumerate(potential.parameters()):
                    record.p[k] = value
                potentials.append(record)
            return index_of[key]

        bonds, bond_ids, angles, angle_ids, dihedrals, dihedral_ids = [], [], [], [], [], []
        have_bonded = bool(system.bond_potentials or system.angle_potentials or system.dihedral_potentials)
        if have_bonded:
            for bonding in bondings:
                for (i, j) in sorted(bonding.bonds):
                    bonds.append((i, j))
                    bond_ids.append(potential_id(system.bond_potential(i, j)))
                for (i, j, k) in sorted(bonding.angles):
                    angles.append((i, j, k))
                    angle_ids.append(potential_id(system.angle_potential(i, j, k)))
                for (i, j, k, m) in sorted(bonding.dihedrals):
                    dihedrals.append((i, j, k, m))
                    dihedral_ids.append(potential_id(system.dihedral_potential(i, j, k, m)))
        array = (_ffi.Potential * max(len(potentials), 1))(*potentials)
        self._check(lib.lumol_cuda_set_bonded_potentials(ctx, len(potentials), array))
        for setter, terms, ids in (
            (lib.lumol_cuda_set_bonds, bonds, bond_ids),
            (lib.lumol_cuda_set_angles, angles, angle_ids),
            (lib.lumol_cuda_set_dihedrals, dihedrals, dihedral_ids),
        ):
            atoms = np.ascontiguousarray(terms, dtype=np.int64).reshape(-1)
            pots = np.ascontiguousarray(ids, dtype=np.int32)
            self._check(
                setter(ctx, len(ids), atoms.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), pots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
            )

        mode, frozen = system.simulated_degrees_of_freedom
        self._check(lib.lumol_cuda_md_set_degrees_of_freedom(ctx, _ffi.DOF_MOLECULES if mode == "molecules" else _ffi.DOF_PARTICLES, frozen))

    # ---- evaluation --------------------------------------------------------------------------------------
    def compute(self, forces=False, energy=False, virial=False, molecular_virial=False, parts=_ffi.PART_ALL):
        what = (
            (_ffi.FORCES if forces else 0) | (_ffi.ENERGY if energy else 0) | (_ffi.ATOMIC_VIRIAL if virial else 0)
            | (_ffi.MOLECULAR_VIRIAL if molecular_virial else 0)
        )
        result = ComputeResult()
        force_array = np.zeros((self._n, 3)) if forces else None
        energy_record = _ffi.Energy()
        virial_array = np.zeros((3, 3))
        self._check(
            self.lib.lumol_cuda_compute(
                self.ctx, what, parts, _ffi.as_double_pointer(force_array) if forces else None,
                ctypes.byref(energy_record), _ffi.as_double_pointer(virial_array),
            )
        )
        result.forces = force_array
        result.energy = energy_record
        result.virial = virial_array
        return result

    def kinetic_energy(self):
        value = ctypes.c_double()
        self._check(self.lib.lumol_cuda_kinetic_energy(self.ctx, ctypes.byref(value)))
        return value.value

    def kinetic_tensor(self):
        tensor = np.zeros((3, 3))
        self._check(self.lib.lumol_cuda_kinetic_tensor(self.ctx, _ffi.as_double_pointer(tensor)))
        return tensor

    def download(self, system, positions=True, velocities=True):
        resident = getattr(system, "_resident", None)
        if positions:
            array = np.zeros((self._n, 3))
            self._check(self.lib.lumol_cuda_get_positions(self.ctx, _ffi.as_double_pointer(array)))
            system.positions = array
            if resident:
                resident.discard("positions")
        if velocities:
            array = np.zeros((self._n, 3))
            self._check(self.lib.lumol_cuda_get_velocities(self.ctx, _ffi.as_double_pointer(array)))
            system.velocities = array
            if resident:
                resident.discard("velocities")

    def stats(self):
        stats = _ffi.Stats()
        self._check(self.lib.lumol_cuda_get_stats(self.ctx, ctypes.byref(stats)))
        return stats

    def set_neighbor_path(self, path):
        self._check(self.lib.lumol_cuda_set_neighbor_path(self.ctx, path))

    def set_kspace_algorithm(self, algorithm):
        self._check(self.lib.lumol_cuda_set_kspace_algorithm(self.ctx, algorithm))


def device_for(system, coulomb="system", velocities=False, positions=True):
    """The ``DeviceSystem`` of ``system``, created on first use and synchronised with the host arrays.
    ``positions=False`` keeps the resident positions (after device-side Monte Carlo moves) unless the structure of
    the system changed."""
    if system._device is None:
        system._device = DeviceSystem(getattr(system, "device_ordinal", 0))
    system._device.sync(system, coulomb=coulomb, positions=positions, velocities=velocities)
    return system._device
