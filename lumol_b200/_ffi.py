"""ctypes binding of ``include/lumol_cuda.h`` (the C ABI of ``liblumol_cuda.so``).

There is no fallback: if the shared library is missing, or the machine has no CUDA device, every
product path raises ``LumolCudaError``.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.path.join(_HERE, "liblumol_cuda.so")

# enums of include/lumol_cuda.h
POTENTIAL_NULL, POTENTIAL_LJ, POTENTIAL_HARMONIC, POTENTIAL_BUCKINGHAM, POTENTIAL_BMH = 0, 1, 2, 3, 4
POTENTIAL_MORSE, POTENTIAL_GAUSSIAN, POTENTIAL_MIE, POTENTIAL_COSINE_HARMONIC, POTENTIAL_TORSION = 5, 6, 7, 8, 9
POTENTIAL_TABLE, POTENTIAL_ABSENT = 10, -1
RESTRICTION_NONE, RESTRICTION_INTRA_MOLECULAR, RESTRICTION_INTER_MOLECULAR = 0, 1, 2
RESTRICTION_EXCLUDE12, RESTRICTION_EXCLUDE13, RESTRICTION_EXCLUDE14, RESTRICTION_SCALE14 = 3, 4, 5, 6
CELL_INFINITE, CELL_ORTHORHOMBIC, CELL_TRICLINIC = 0, 1, 2
FORCES, ENERGY, ATOMIC_VIRIAL, MOLECULAR_VIRIAL = 1, 2, 4, 8
OWNED_FORCES = 16  # sharded contexts: lumol_cuda_compute returns only this rank's block of forces (lumol_cuda_owned_range)
PART_PAIRS, PART_BONDED, PART_COULOMB, PART_ALL = 1, 2, 4, 7
INTEGRATOR_VELOCITY_VERLET, INTEGRATOR_VERLET, INTEGRATOR_LEAP_FROG = 0, 1, 2
INTEGRATOR_BERENDSEN_BAROSTAT, INTEGRATOR_ANISO_BERENDSEN_BAROSTAT = 3, 4
THERMOSTAT_NONE, THERMOSTAT_RESCALE, THERMOSTAT_BERENDSEN, THERMOSTAT_CSVR = 0, 1, 2, 3
CONTROL_REMOVE_TRANSLATION, CONTROL_REMOVE_ROTATION, CONTROL_REWRAP = 1, 2, 4
DOF_PARTICLES, DOF_MOLECULES = 0, 1

SUCCESS = 0
ERROR_INVALID_ARGUMENT, ERROR_NO_DEVICE, ERROR_CUDA, ERROR_STATE = -1, -2, -3, -4
ERROR_INFINITE_CELL, ERROR_NOT_FINITE, ERROR_UNSUPPORTED, ERROR_COMM = -5, -6, -7, -8


class LumolCudaError(RuntimeError):
    """A non-zero status from the C ABI; the Rust shim turns these into the reference's panics."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


class Potential(ctypes.Structure):
    _fields_ = [("potential", ctypes.c_int32), ("reserved", ctypes.c_int32), ("p", ctypes.c_double * 5)]


class Pair(ctypes.Structure):
    _fields_ = [
        ("potential", ctypes.c_int32),
        ("restriction", ctypes.c_int32),
        ("table", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("p", ctypes.c_double * 5),
        ("cutoff", ctypes.c_double),
        ("shift", ctypes.c_double),
        ("scale14", ctypes.c_double),
        ("tail_energy", ctypes.c_double),
        ("tail_virial", ctypes.c_double),
    ]


class Energy(ctypes.Structure):
    _fields_ = [
        ("pairs", ctypes.c_double),
        ("pairs_tail", ctypes.c_double),
        ("bonds", ctypes.c_double),
        ("angles", ctypes.c_double),
        ("dihedrals", ctypes.c_double),
        ("coulomb_real", ctypes.c_double),
        ("coulomb_self", ctypes.c_double),
        ("coulomb_kspace", ctypes.c_double),
    ]

    def total(self):
        # PotentialEnergy::compute (compute.rs:114-127); coulomb = real + self + k_space (ewald.rs:888-895)
        energy = self.pairs
        energy += self.pairs_tail
        energy += self.bonds
        energy += self.angles
        energy += self.dihedrals
        energy += self.coulomb_real + self.coulomb_self + self.coulomb_kspace
        return energy


class Stats(ctypes.Structure):
    _fields_ = [
        ("natoms", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("neighbor_path", ctypes.c_int64),
        ("ncells", ctypes.c_int64 * 3),
        ("nkvectors", ctypes.c_int64),
        ("pair_launches", ctypes.c_int64),
        ("pair_ms", ctypes.c_double),
        ("kspace_launches", ctypes.c_int64),
        ("kspace_ms", ctypes.c_double),
        ("integrate_launches", ctypes.c_int64),
        ("integrate_ms", ctypes.c_double),
        ("neighbor_launches", ctypes.c_int64),
        ("neighbor_ms", ctypes.c_double),
        ("comm_launches", ctypes.c_int64),
        ("comm_ms", ctypes.c_double),
        ("pair_count", ctypes.c_double),
        ("coulomb_pair_count", ctypes.c_double),
        ("neighbor_rebuilds", ctypes.c_int64),
        ("neighbor_skin", ctypes.c_double),
    ]


_c = ctypes
_ctx = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); this is the complete list of symbols include/lumol_cuda.h declares
SIGNATURES = {
    "lumol_cuda_abi_version": (_c.c_int32, []),
    "lumol_cuda_create": (_c.c_int32, [_c.c_int32, _c.POINTER(_ctx)]),
    "lumol_cuda_create_multi": (_c.c_int32, [_c.POINTER(_c.c_int32), _c.c_int32, _c.POINTER(_ctx)]),
    "lumol_cuda_destroy": (_c.c_int32, [_ctx]),
    "lumol_cuda_last_error": (_c.c_char_p, [_ctx]),
    "lumol_cuda_set_cell": (_c.c_int32, [_ctx, _dp, _c.c_int32]),
    "lumol_cuda_set_particles": (_c.c_int32, [_ctx, _c.c_int64, _dp, _dp, _dp, _dp, _c.POINTER(_c.c_uint32)]),
    "lumol_cuda_set_positions": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_set_owned_positions": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_set_velocities": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_get_positions": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_owned_range": (_c.c_int32, [_ctx, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
    "lumol_cuda_get_velocities": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_get_forces": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_set_molecules": (
        _c.c_int32,
        [_ctx, _c.c_int64, _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint64), _c.POINTER(_c.c_uint8), _c.c_uint64],
    ),
    "lumol_cuda_set_pairs": (_c.c_int32, [_ctx, _c.c_int32, _c.POINTER(Pair)]),
    "lumol_cuda_add_table": (_c.c_int32, [_ctx, _c.c_int32, _c.c_double, _dp, _dp]),
    "lumol_cuda_clear_tables": (_c.c_int32, [_ctx]),
    "lumol_cuda_set_bonded_potentials": (_c.c_int32, [_ctx, _c.c_int32, _c.POINTER(Potential)]),
    "lumol_cuda_set_bonds": (_c.c_int32, [_ctx, _c.c_int64, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int32)]),
    "lumol_cuda_set_angles": (_c.c_int32, [_ctx, _c.c_int64, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int32)]),
    "lumol_cuda_set_dihedrals": (_c.c_int32, [_ctx, _c.c_int64, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int32)]),
    "lumol_cuda_set_coulomb_none": (_c.c_int32, [_ctx]),
    "lumol_cuda_set_coulomb_ewald": (_c.c_int32, [_ctx, _c.c_double, _c.c_double, _c.c_int32, _c.c_int32]),
    "lumol_cuda_set_coulomb_wolf": (_c.c_int32, [_ctx, _c.c_double, _c.c_int32, _c.c_double]),
    "lumol_cuda_compute": (_c.c_int32, [_ctx, _c.c_uint32, _c.c_uint32, _dp, _c.POINTER(Energy), _dp]),
    "lumol_cuda_kinetic_energy": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_kinetic_tensor": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_ewald_kvectors": (
        _c.c_int32,
        [_ctx, _c.c_int64, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int32), _dp, _dp],
    ),
    "lumol_cuda_move_molecule_cost": (_c.c_int32, [_ctx, _c.c_int64, _dp, _c.POINTER(Energy)]),
    "lumol_cuda_move_molecules_cost": (_c.c_int32, [_ctx, _c.c_int64, _c.POINTER(_c.c_int64), _dp, _c.POINTER(Energy)]),
    "lumol_cuda_move_molecule_accept": (_c.c_int32, [_ctx, _c.c_int64]),
    "lumol_cuda_md_setup": (_c.c_int32, [_ctx, _c.c_int32, _c.c_double]),
    "lumol_cuda_md_set_degrees_of_freedom": (_c.c_int32, [_ctx, _c.c_int32, _c.c_int64]),
    "lumol_cuda_md_set_thermostat": (_c.c_int32, [_ctx, _c.c_int32, _c.c_double, _c.c_double]),
    "lumol_cuda_md_set_csvr_noise": (_c.c_int32, [_ctx, _c.c_int64, _dp]),
    "lumol_cuda_md_set_controls": (_c.c_int32, [_ctx, _c.c_uint32]),
    "lumol_cuda_md_run": (_c.c_int32, [_ctx, _c.c_int64]),
    "lumol_cuda_scale_velocities": (_c.c_int32, [_ctx, _c.c_double]),
    "lumol_cuda_remove_translation": (_c.c_int32, [_ctx]),
    "lumol_cuda_remove_rotation": (_c.c_int32, [_ctx]),
    "lumol_cuda_rewrap": (_c.c_int32, [_ctx]),
    "lumol_cuda_md_set_barostat": (_c.c_int32, [_ctx, _dp, _c.c_double]),
    "lumol_cuda_get_cell": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_comm_unique_id": (_c.c_int32, [_c.POINTER(_c.c_uint8)]),
    "lumol_cuda_comm_init": (_c.c_int32, [_ctx, _c.c_int32, _c.c_int32, _c.POINTER(_c.c_uint8)]),
    "lumol_cuda_set_profiling": (_c.c_int32, [_ctx, _c.c_int32]),
    "lumol_cuda_get_stats": (_c.c_int32, [_ctx, _c.POINTER(Stats)]),
    "lumol_cuda_reset_stats": (_c.c_int32, [_ctx]),
    "lumol_cuda_set_neighbor_path": (_c.c_int32, [_ctx, _c.c_int32]),
    "lumol_cuda_set_neighbor_skin": (_c.c_int32, [_ctx, _c.c_double]),
    "lumol_cuda_set_kspace_algorithm": (_c.c_int32, [_ctx, _c.c_int32]),
    "lumol_cuda_stream": (_c.c_void_p, [_ctx]),
    "lumol_cuda_synchronize": (_c.c_int32, [_ctx]),
    "lumol_cuda_measure_fp64_peak": (_c.c_int32, [_ctx, _dp]),
    "lumol_cuda_measure_copy_bandwidth": (_c.c_int32, [_ctx, _dp]),
}

_library = None


def library():
    """Load ``liblumol_cuda.so`` (built in-tree by ``__graft_entry__.build()``); raises when it is absent."""
    global _library
    if _library is None:
        if not os.path.exists(LIBRARY_PATH):
            raise LumolCudaError(
                ERROR_STATE,
                f"{LIBRARY_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "lumol_b200 has no CPU fallback.",
            )
        lib = ctypes.CDLL(LIBRARY_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            function = getattr(lib, name)
            function.restype = restype
            function.argtypes = argtypes
        _library = lib
    return _library


def last_error(ctx):
    message = library().lumol_cuda_last_error(ctx)
    return message.decode("utf-8", "replace") if message else ""


def check(ctx, status):
    """Raise ``LumolCudaError`` for a negative status."""
    if status < 0:
        raise LumolCudaError(status, last_error(ctx) or f"lumol_cuda error {status}")
    return status


def as_double_pointer(array):
    return array.ctypes.data_as(_dp)
