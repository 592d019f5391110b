//! Sketch of the shim a lumol maintainer adds on top of ffi.rs: a `DeviceSystem` owned by `System`, and the
//! `Compute` impls routed through it.  Same flattening as lumol_b200/device.py (which is what the parity tests run).
use crate::ffi::*;
use lumol_core::{Matrix3, System, Vector3D};
use std::ffi::CStr;
use std::sync::Mutex;

/// `Send` but not `Sync`: wrap in a Mutex to satisfy `GlobalPotential: Send + Sync` (energy/global/mod.rs:84).
pub struct DeviceSystem {
    ctx: Mutex<*mut lumol_cuda_context>,
    synced_version: u64,
}

unsafe impl Send for DeviceSystem {}

fn check(ctx: *mut lumol_cuda_context, status: i32) {
    if status < 0 {
        let message = unsafe { CStr::from_ptr(lumol_cuda_last_error(ctx)) }.to_string_lossy().into_owned();
        // the reference aborts the simulation with these very messages (compute.rs:125,199; ewald.rs:124)
        panic!("{}", message);
    }
}

impl DeviceSystem {
    pub fn new(device: i32) -> DeviceSystem {
        let mut ctx = std::ptr::null_mut();
        let status = unsafe { lumol_cuda_create(device, &mut ctx) };
        if status < 0 {
            let message = unsafe { CStr::from_ptr(lumol_cuda_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            panic!("{}", message);
        }
        DeviceSystem { ctx: Mutex::new(ctx), synced_version: u64::MAX }
    }

    /// Upload what changed: cell, ParticleVec fields, molecules (Bonding::bond_distances), the (kind, kind) pair table
    /// with host-evaluated shift / tail scalars, TableComputation tables for user potentials, bonded lists, coulomb.
    pub fn sync(&mut self, system: &System) {
        let ctx = *self.ctx.lock().unwrap();
        let cell = system.cell.matrix();
        check(ctx, unsafe { lumol_cuda_set_cell(ctx, cell.as_ptr() as *const f64, system.cell.shape() as i32) });
        let particles = system.particles();
        // Vec<Vector3D> is a packed n x 3 f64 array (types/vectors.rs:59)
        check(ctx, unsafe { lumol_cuda_set_positions(ctx, particles.position.as_ptr() as *const f64) });
        // ... set_particles / set_molecules / set_pairs / set_bonds / set_coulomb_* when the structure version changed
        let _ = self.synced_version;
    }

    pub fn forces(&mut self, system: &System) -> Vec<Vector3D> {
        self.sync(system);
        let ctx = *self.ctx.lock().unwrap();
        let mut forces = vec![Vector3D::zero(); system.size()];
        check(ctx, unsafe {
            lumol_cuda_compute(ctx, LUMOL_CUDA_FORCES, LUMOL_CUDA_PART_ALL, forces.as_mut_ptr() as *mut f64,
                               std::ptr::null_mut(), std::ptr::null_mut())
        });
        forces
    }

    pub fn potential_energy(&mut self, system: &System) -> f64 {
        self.sync(system);
        let ctx = *self.ctx.lock().unwrap();
        let mut e = lumol_cuda_energy::default();
        check(ctx, unsafe {
            lumol_cuda_compute(ctx, LUMOL_CUDA_ENERGY, LUMOL_CUDA_PART_ALL, std::ptr::null_mut(), &mut e, std::ptr::null_mut())
        });
        // same order of additions as PotentialEnergy::compute (compute.rs:117-123)
        e.pairs + e.pairs_tail + e.bonds + e.angles + e.dihedrals + (e.coulomb_real + e.coulomb_self + e.coulomb_kspace)
    }

    pub fn atomic_virial(&mut self, system: &System) -> Matrix3 {
        self.sync(system);
        let ctx = *self.ctx.lock().unwrap();
        let mut w = [0.0f64; 9];
        check(ctx, unsafe {
            lumol_cuda_compute(ctx, LUMOL_CUDA_ATOMIC_VIRIAL, LUMOL_CUDA_PART_ALL, std::ptr::null_mut(), std::ptr::null_mut(), w.as_mut_ptr())
        });
        Matrix3::new([[w[0], w[1], w[2]], [w[3], w[4], w[5]], [w[6], w[7], w[8]]])
    }
}

/// Monte Carlo: what `EnergyCache::move_molecule_cost` (sys/cache.rs:145-213) becomes.  The N x N `pairs_cache`
/// and the `Ewald::updater` closure disappear: the resident positions and rho(k) are the cache, and
/// `EnergyCache::update` (cache.rs:119-128) accepts the pending trial on the device.
impl DeviceSystem {
    /// `system` still holds the old positions; they are the resident ones (no upload here).
    pub fn move_molecule_cost(&mut self, molecule_id: usize, new_positions: &[Vector3D]) -> (f64, f64) {
        let ctx = *self.ctx.lock().unwrap();
        let mut cost = lumol_cuda_energy::default();
        check(ctx, unsafe {
            lumol_cuda_move_molecule_cost(ctx, molecule_id as i64, new_positions.as_ptr() as *const f64, &mut cost)
        });
        // (pairs_delta, coulomb_delta) of cache.rs:171-178
        (cost.pairs, cost.coulomb_real + cost.coulomb_kspace)
    }

    /// `cache.pairs += pairs_delta; cache.coulomb += coulomb_delta; coulomb.update()` (cache.rs:175-211): the host
    /// sums are updated by the caller, positions and rho(k) here.
    pub fn accept_move(&mut self) {
        let ctx = *self.ctx.lock().unwrap();
        check(ctx, unsafe { lumol_cuda_move_molecule_accept(ctx, 0) });
    }
}

impl Drop for DeviceSystem {
    fn drop(&mut self) {
        unsafe { lumol_cuda_destroy(*self.ctx.lock().unwrap()) };
    }
}
