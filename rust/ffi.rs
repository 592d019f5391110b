//! `extern "C"` declarations of include/lumol_cuda.h for the lumol workspace (a `lumol-cuda` crate next to
//! lumol-core).  One-to-one with the header; see INTEGRATION.md for the trait impls built on it.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct lumol_cuda_context {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct lumol_cuda_potential {
    pub potential: i32,
    pub reserved: i32,
    pub p: [f64; 5],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct lumol_cuda_pair {
    pub potential: i32,
    pub restriction: i32,
    pub table: i32,
    pub reserved: i32,
    pub p: [f64; 5],
    pub cutoff: f64,
    pub shift: f64,
    pub scale14: f64,
    pub tail_energy: f64,
    pub tail_virial: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct lumol_cuda_energy {
    pub pairs: f64,
    pub pairs_tail: f64,
    pub bonds: f64,
    pub angles: f64,
    pub dihedrals: f64,
    pub coulomb_real: f64,
    pub coulomb_self: f64,
    pub coulomb_kspace: f64,
}

/// Counters of the measurement harness (`lumol_cuda_stats`).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct lumol_cuda_stats {
    pub natoms: i64,
    pub kernel_launches: i64,
    pub neighbor_path: i64,
    pub ncells: [i64; 3],
    pub nkvectors: i64,
    pub pair_launches: i64,
    pub pair_ms: f64,
    pub kspace_launches: i64,
    pub kspace_ms: f64,
    pub integrate_launches: i64,
    pub integrate_ms: f64,
    pub neighbor_launches: i64,
    pub neighbor_ms: f64,
    pub comm_launches: i64,
    pub comm_ms: f64,
    pub pair_count: f64,
    pub coulomb_pair_count: f64,
    pub neighbor_rebuilds: i64,
    pub neighbor_skin: f64,
}

pub const LUMOL_CUDA_FORCES: u32 = 1;
pub const LUMOL_CUDA_ENERGY: u32 = 2;
pub const LUMOL_CUDA_ATOMIC_VIRIAL: u32 = 4;
pub const LUMOL_CUDA_MOLECULAR_VIRIAL: u32 = 8;
pub const LUMOL_CUDA_OWNED_FORCES: u32 = 16;
pub const LUMOL_CUDA_PART_PAIRS: u32 = 1;
pub const LUMOL_CUDA_PART_BONDED: u32 = 2;
pub const LUMOL_CUDA_PART_COULOMB: u32 = 4;
pub const LUMOL_CUDA_PART_ALL: u32 = 7;

pub const LUMOL_CUDA_POTENTIAL_ABSENT: i32 = -1;
pub const LUMOL_CUDA_POTENTIAL_NULL: i32 = 0;
pub const LUMOL_CUDA_POTENTIAL_LJ: i32 = 1;
pub const LUMOL_CUDA_POTENTIAL_HARMONIC: i32 = 2;
pub const LUMOL_CUDA_POTENTIAL_BUCKINGHAM: i32 = 3;
pub const LUMOL_CUDA_POTENTIAL_BMH: i32 = 4;
pub const LUMOL_CUDA_POTENTIAL_MORSE: i32 = 5;
pub const LUMOL_CUDA_POTENTIAL_GAUSSIAN: i32 = 6;
pub const LUMOL_CUDA_POTENTIAL_MIE: i32 = 7;
pub const LUMOL_CUDA_POTENTIAL_COSINE_HARMONIC: i32 = 8;
pub const LUMOL_CUDA_POTENTIAL_TORSION: i32 = 9;
pub const LUMOL_CUDA_POTENTIAL_TABLE: i32 = 10;
pub const LUMOL_CUDA_RESTRICTION_NONE: i32 = 0;
pub const LUMOL_CUDA_RESTRICTION_INTRA_MOLECULAR: i32 = 1;
pub const LUMOL_CUDA_RESTRICTION_INTER_MOLECULAR: i32 = 2;
pub const LUMOL_CUDA_RESTRICTION_EXCLUDE12: i32 = 3;
pub const LUMOL_CUDA_RESTRICTION_EXCLUDE13: i32 = 4;
pub const LUMOL_CUDA_RESTRICTION_EXCLUDE14: i32 = 5;
pub const LUMOL_CUDA_RESTRICTION_SCALE14: i32 = 6;
pub const LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET: i32 = 0;
pub const LUMOL_CUDA_INTEGRATOR_VERLET: i32 = 1;
pub const LUMOL_CUDA_INTEGRATOR_LEAP_FROG: i32 = 2;
pub const LUMOL_CUDA_INTEGRATOR_BERENDSEN_BAROSTAT: i32 = 3;
pub const LUMOL_CUDA_INTEGRATOR_ANISO_BERENDSEN_BAROSTAT: i32 = 4;

extern "C" {
    pub fn lumol_cuda_abi_version() -> i32;
    pub fn lumol_cuda_create(device: i32, ctx: *mut *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_create_multi(devices: *const i32, ndevices: i32, ctx: *mut *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_destroy(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_last_error(ctx: *const lumol_cuda_context) -> *const c_char;
    pub fn lumol_cuda_set_cell(ctx: *mut lumol_cuda_context, cell: *const f64, shape: i32) -> i32;
    pub fn lumol_cuda_set_particles(
        ctx: *mut lumol_cuda_context, n: i64, position: *const f64, velocity: *const f64, mass: *const f64,
        charge: *const f64, kind: *const u32,
    ) -> i32;
    pub fn lumol_cuda_set_positions(ctx: *mut lumol_cuda_context, position: *const f64) -> i32;
    pub fn lumol_cuda_set_owned_positions(ctx: *mut lumol_cuda_context, owned_position: *const f64) -> i32;
    pub fn lumol_cuda_set_velocities(ctx: *mut lumol_cuda_context, velocity: *const f64) -> i32;
    pub fn lumol_cuda_get_positions(ctx: *mut lumol_cuda_context, position: *mut f64) -> i32;
    pub fn lumol_cuda_get_velocities(ctx: *mut lumol_cuda_context, velocity: *mut f64) -> i32;
    pub fn lumol_cuda_get_forces(ctx: *mut lumol_cuda_context, forces: *mut f64) -> i32;
    pub fn lumol_cuda_set_molecules(
        ctx: *mut lumol_cuda_context, nmol: i64, start: *const u64, bond_distances_offset: *const u64,
        bond_distances: *const u8, bond_distances_size: u64,
    ) -> i32;
    pub fn lumol_cuda_set_pairs(ctx: *mut lumol_cuda_context, nkinds: i32, pairs: *const lumol_cuda_pair) -> i32;
    pub fn lumol_cuda_add_table(ctx: *mut lumol_cuda_context, size: i32, max: f64, energy: *const f64, force: *const f64) -> i32;
    pub fn lumol_cuda_clear_tables(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_set_bonded_potentials(ctx: *mut lumol_cuda_context, n: i32, potentials: *const lumol_cuda_potential) -> i32;
    pub fn lumol_cuda_set_bonds(ctx: *mut lumol_cuda_context, n: i64, atoms: *const i64, potential: *const i32) -> i32;
    pub fn lumol_cuda_set_angles(ctx: *mut lumol_cuda_context, n: i64, atoms: *const i64, potential: *const i32) -> i32;
    pub fn lumol_cuda_set_dihedrals(ctx: *mut lumol_cuda_context, n: i64, atoms: *const i64, potential: *const i32) -> i32;
    pub fn lumol_cuda_set_coulomb_none(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_set_coulomb_ewald(ctx: *mut lumol_cuda_context, cutoff: f64, alpha: f64, kmax: i32, restriction: i32) -> i32;
    pub fn lumol_cuda_set_coulomb_wolf(ctx: *mut lumol_cuda_context, cutoff: f64, restriction: i32, scale14: f64) -> i32;
    pub fn lumol_cuda_compute(
        ctx: *mut lumol_cuda_context, what: u32, parts: u32, forces: *mut f64, energy: *mut lumol_cuda_energy,
        virial: *mut f64,
    ) -> i32;
    pub fn lumol_cuda_kinetic_energy(ctx: *mut lumol_cuda_context, kinetic: *mut f64) -> i32;
    pub fn lumol_cuda_kinetic_tensor(ctx: *mut lumol_cuda_context, tensor: *mut f64) -> i32;
    pub fn lumol_cuda_move_molecule_cost(
        ctx: *mut lumol_cuda_context, molecule: i64, new_positions: *const f64, cost: *mut lumol_cuda_energy,
    ) -> i32;
    pub fn lumol_cuda_move_molecules_cost(
        ctx: *mut lumol_cuda_context, ntrials: i64, molecules: *const i64, new_positions: *const f64,
        costs: *mut lumol_cuda_energy,
    ) -> i32;
    pub fn lumol_cuda_move_molecule_accept(ctx: *mut lumol_cuda_context, trial: i64) -> i32;
    pub fn lumol_cuda_md_setup(ctx: *mut lumol_cuda_context, integrator: i32, timestep: f64) -> i32;
    pub fn lumol_cuda_md_set_degrees_of_freedom(ctx: *mut lumol_cuda_context, mode: i32, frozen: i64) -> i32;
    pub fn lumol_cuda_md_set_thermostat(ctx: *mut lumol_cuda_context, thermostat: i32, temperature: f64, parameter: f64) -> i32;
    pub fn lumol_cuda_md_set_csvr_noise(ctx: *mut lumol_cuda_context, nsteps: i64, noise: *const f64) -> i32;
    pub fn lumol_cuda_md_set_controls(ctx: *mut lumol_cuda_context, controls: u32) -> i32;
    pub fn lumol_cuda_md_run(ctx: *mut lumol_cuda_context, nsteps: i64) -> i32;
    pub fn lumol_cuda_scale_velocities(ctx: *mut lumol_cuda_context, factor: f64) -> i32;
    pub fn lumol_cuda_remove_translation(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_remove_rotation(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_rewrap(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_md_set_barostat(ctx: *mut lumol_cuda_context, target: *const f64, tau: f64) -> i32;
    pub fn lumol_cuda_get_cell(ctx: *mut lumol_cuda_context, cell: *mut f64) -> i32;
    pub fn lumol_cuda_comm_unique_id(id: *mut u8) -> i32;
    pub fn lumol_cuda_comm_init(ctx: *mut lumol_cuda_context, nranks: i32, rank: i32, id: *const u8) -> i32;
    pub fn lumol_cuda_owned_range(ctx: *mut lumol_cuda_context, first: *mut i64, count: *mut i64) -> i32;
    pub fn lumol_cuda_set_neighbor_skin(ctx: *mut lumol_cuda_context, skin: f64) -> i32;
    pub fn lumol_cuda_set_neighbor_path(ctx: *mut lumol_cuda_context, path: i32) -> i32;
    pub fn lumol_cuda_set_kspace_algorithm(ctx: *mut lumol_cuda_context, algorithm: i32) -> i32;
    pub fn lumol_cuda_ewald_kvectors(
        ctx: *mut lumol_cuda_context, capacity: i64, count: *mut i64, index: *mut i32, energy_factor: *mut f64, rho: *mut f64,
    ) -> i32;
    pub fn lumol_cuda_set_profiling(ctx: *mut lumol_cuda_context, enabled: i32) -> i32;
    pub fn lumol_cuda_get_stats(ctx: *mut lumol_cuda_context, stats: *mut lumol_cuda_stats) -> i32;
    pub fn lumol_cuda_reset_stats(ctx: *mut lumol_cuda_context) -> i32;
    pub fn lumol_cuda_measure_fp64_peak(ctx: *mut lumol_cuda_context, tflops: *mut f64) -> i32;
    pub fn lumol_cuda_measure_copy_bandwidth(ctx: *mut lumol_cuda_context, gbs: *mut f64) -> i32;
    pub fn lumol_cuda_stream(ctx: *mut lumol_cuda_context) -> *mut c_void;
    pub fn lumol_cuda_synchronize(ctx: *mut lumol_cuda_context) -> i32;
}
