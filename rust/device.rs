//! The shim a lumol maintainer adds on top of ffi.rs: a `DeviceSystem` owned by `System`, and the `Compute` impls
//! routed through it; describe.rs lists the small accessors it needs inside lumol-core.
//! Not compiled in this image (no rustc).  Same flattening as lumol_b200/device.py (which is what the parity tests run).
use crate::ffi::*;
use lumol_core::{CellShape, Matrix3, System, Vector3D};
use std::ffi::CStr;
use std::sync::Mutex;

/// `Send` but not `Sync`: wrap in a Mutex to satisfy `GlobalPotential: Send + Sync` (energy/global/mod.rs:84).
pub struct DeviceSystem {
    ctx: Mutex<*mut lumol_cuda_context>,
    synced_version: u64,
    synced_cell: Option<(Matrix3, CellShape)>,
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
        DeviceSystem { ctx: Mutex::new(ctx), synced_version: u64::MAX, synced_cell: None }
    }

    /// One context over several devices of this process (`lumol_cuda_create_multi`): same calls afterwards, the library
    /// shards the system over the devices behind them.  `LUMOL_CUDA_DEVICES=0,1,2,3` in lumol's environment is the
    /// intended switch (`System::device()` builds the context lazily on first use).
    pub fn new_multi(devices: &[i32]) -> DeviceSystem {
        let mut ctx = std::ptr::null_mut();
        let status = unsafe { lumol_cuda_create_multi(devices.as_ptr(), devices.len() as i32, &mut ctx) };
        if status < 0 {
            let message = unsafe { CStr::from_ptr(lumol_cuda_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            panic!("{}", message);
        }
        DeviceSystem { ctx: Mutex::new(ctx), synced_version: u64::MAX, synced_cell: None }
    }

    /// Upload what changed since the last call (the flattening of lumol_b200/device.py, which the parity tests run):
    /// cell when it differs, positions always (the host arrays are the truth in host-driven mode), everything else
    /// when `System::structure_version()` moved -- a counter lumol-core bumps in `add_molecule`, `add_bond`,
    /// `particles_mut()` on kind / charge / mass, and every `set_*_potential` (system.rs:80-175).
    pub fn sync(&mut self, system: &System) {
        let ctx = *self.ctx.lock().unwrap();
        let cell = system.cell.matrix();
        if self.synced_cell != Some((cell, system.cell.shape())) {
            // Matrix3 is [[f64; 3]; 3] row-major (types/matrix.rs:73); CellShape discriminants match the header
            check(ctx, unsafe { lumol_cuda_set_cell(ctx, cell.as_ptr() as *const f64, system.cell.shape() as i32) });
            self.synced_cell = Some((cell, system.cell.shape()));
        }
        if self.synced_version == system.structure_version() {
            // Vec<Vector3D> is a packed n x 3 f64 array (types/vectors.rs:59)
            check(ctx, unsafe { lumol_cuda_set_positions(ctx, system.particles().position.as_ptr() as *const f64) });
        } else {
            self.upload_structure(ctx, system);
            self.synced_version = system.structure_version();
        }
        match system.coulomb_potential() {
            None => check(ctx, unsafe { lumol_cuda_set_coulomb_none(ctx) }),
            Some(coulomb) => {
                if !coulomb.device_configure(ctx) {
                    panic!("this coulombic potential has no device implementation");
                }
            }
        }
        if !system.global_potentials().is_empty() {
            panic!("user-defined global potentials are evaluated on the host only");
        }
    }

    fn upload_structure(&mut self, ctx: *mut lumol_cuda_context, system: &System) {
        let particles = system.particles();
        let n = system.size();
        // Kind is a newtype over u32 (particles.rs:18-22)
        let kinds: Vec<u32> = particles.kind.iter().map(|kind| kind.0).collect();
        check(ctx, unsafe {
            lumol_cuda_set_particles(ctx, n as i64, particles.position.as_ptr() as *const f64,
                particles.velocity.as_ptr() as *const f64, particles.mass.as_ptr(), particles.charge.as_ptr(), kinds.as_ptr())
        });

        // molecules: contiguous ranges (bonding.rs:50-57) and the bond-distance bit masks (bonding.rs:130-155,
        // 276-279), one matrix per distinct molecule hash
        let mut start = Vec::<u64>::new();
        let mut offset = Vec::<u64>::new();
        let mut bytes = Vec::<u8>::new();
        let mut known = std::collections::BTreeMap::new(); // MoleculeHash is Ord, not Hash (molecules.rs:17)
        let mut any_bonded = false;
        for molecule in system.molecules() {
            start.push(molecule.start() as u64);
            any_bonded |= molecule.size() > 1;
            let first = *known.entry(molecule.hash()).or_insert_with(|| {
                let first = bytes.len() as u64;
                for i in molecule.indexes() {
                    for j in molecule.indexes() {
                        bytes.push(molecule.bond_distances(i, j).bits());
                    }
                }
                first
            });
            offset.push(first);
        }
        start.push(n as u64);
        if any_bonded {
            check(ctx, unsafe {
                lumol_cuda_set_molecules(ctx, offset.len() as i64, start.as_ptr(), offset.as_ptr(), bytes.as_ptr(), bytes.len() as u64)
            });
        } else {
            check(ctx, unsafe { lumol_cuda_set_molecules(ctx, 0, std::ptr::null(), std::ptr::null(), std::ptr::null(), 0) });
        }

        // (kind, kind) table; a pair without an entry keeps POTENTIAL_ABSENT (interactions.rs:142-145)
        let nkinds = kinds.iter().max().map_or(0, |&k| k as usize + 1);
        let mut absent = lumol_cuda_pair::default();
        absent.potential = LUMOL_CUDA_POTENTIAL_ABSENT;
        let mut records = vec![absent; std::cmp::max(nkinds * nkinds, 1)];
        let mut tables = Vec::new();
        let mut representative = vec![usize::MAX; nkinds];
        for (i, &kind) in kinds.iter().enumerate() {
            if representative[kind as usize] == usize::MAX {
                representative[kind as usize] = i;
            }
        }
        for a in 0..nkinds {
            for b in a..nkinds {
                if representative[a] == usize::MAX || representative[b] == usize::MAX {
                    continue;
                }
                if let Some(interaction) = system.pair_potential(representative[a], representative[b]) {
                    let record = interaction.device_record(&mut tables);
                    records[a * nkinds + b] = record;
                    records[b * nkinds + a] = record;
                }
            }
        }
        check(ctx, unsafe { lumol_cuda_clear_tables(ctx) });
        for (size, max, energy, force) in &tables {
            // ids are handed out in call order, which is the order device_record numbered them in
            check(ctx, unsafe { lumol_cuda_add_table(ctx, *size as i32, *max, energy.as_ptr(), force.as_ptr()) });
        }
        check(ctx, unsafe { lumol_cuda_set_pairs(ctx, nkinds as i32, records.as_ptr()) });

        // bonded terms: explicit lists, one potential id per entry; identical closed forms share an id.  A bond with no
        // potential gets id -1 and contributes nothing (the reference warns once, compute.rs:40-62).
        let mut potentials = Vec::<lumol_cuda_potential>::new();
        let mut id_of = |form: Option<(i32, [f64; 5])>| -> i32 {
            let (potential, p) = match form {
                Some(form) => form,
                None => panic!("bonded potentials need a closed form known to the device library"),
            };
            if let Some(id) = potentials.iter().position(|known| known.potential == potential && known.p == p) {
                return id as i32;
            }
            potentials.push(lumol_cuda_potential { potential, reserved: 0, p });
            potentials.len() as i32 - 1
        };
        let (mut bonds, mut bond_ids) = (Vec::<i64>::new(), Vec::<i32>::new());
        let (mut angles, mut angle_ids) = (Vec::<i64>::new(), Vec::<i32>::new());
        let (mut dihedrals, mut dihedral_ids) = (Vec::<i64>::new(), Vec::<i32>::new());
        for molecule in system.molecules() {
            // HashSet iteration order is arbitrary: sort, so that the summation order on the device is reproducible
            let mut sorted: Vec<_> = molecule.bonds().iter().collect();
            sorted.sort();
            for bond in sorted {
                bonds.extend_from_slice(&[bond.i() as i64, bond.j() as i64]);
                bond_ids.push(system.bond_potential(bond.i(), bond.j()).map_or(-1, |p| id_of(p.device_form())));
            }
            let mut sorted: Vec<_> = molecule.angles().iter().collect();
            sorted.sort_by_key(|a| (a.i(), a.j(), a.k()));
            for angle in sorted {
                angles.extend_from_slice(&[angle.i() as i64, angle.j() as i64, angle.k() as i64]);
                angle_ids.push(system.angle_potential(angle.i(), angle.j(), angle.k()).map_or(-1, |p| id_of(p.device_form())));
            }
            let mut sorted: Vec<_> = molecule.dihedrals().iter().collect();
            sorted.sort_by_key(|d| (d.i(), d.j(), d.k(), d.m()));
            for d in sorted {
                dihedrals.extend_from_slice(&[d.i() as i64, d.j() as i64, d.k() as i64, d.m() as i64]);
                dihedral_ids.push(system.dihedral_potential(d.i(), d.j(), d.k(), d.m()).map_or(-1, |p| id_of(p.device_form())));
            }
        }
        check(ctx, unsafe { lumol_cuda_set_bonded_potentials(ctx, potentials.len() as i32, potentials.as_ptr()) });
        check(ctx, unsafe { lumol_cuda_set_bonds(ctx, bond_ids.len() as i64, bonds.as_ptr(), bond_ids.as_ptr()) });
        check(ctx, unsafe { lumol_cuda_set_angles(ctx, angle_ids.len() as i64, angles.as_ptr(), angle_ids.as_ptr()) });
        check(ctx, unsafe { lumol_cuda_set_dihedrals(ctx, dihedral_ids.len() as i64, dihedrals.as_ptr(), dihedral_ids.as_ptr()) });
    }

    /// Sharded contexts: the block of atoms whose forces this rank reduces (`lumol_cuda_owned_range`); with
    /// `LUMOL_CUDA_OWNED_FORCES` a host-driven step downloads 24 B per owned atom instead of 24 B per atom.
    pub fn owned_forces(&mut self, system: &System) -> (std::ops::Range<usize>, Vec<Vector3D>) {
        self.sync(system);
        let ctx = *self.ctx.lock().unwrap();
        let (mut first, mut count) = (0i64, 0i64);
        check(ctx, unsafe { lumol_cuda_owned_range(ctx, &mut first, &mut count) });
        let mut forces = vec![Vector3D::zero(); count as usize];
        check(ctx, unsafe {
            lumol_cuda_compute(ctx, LUMOL_CUDA_FORCES | LUMOL_CUDA_OWNED_FORCES, LUMOL_CUDA_PART_ALL,
                               forces.as_mut_ptr() as *mut f64, std::ptr::null_mut(), std::ptr::null_mut())
        });
        (first as usize..(first + count) as usize, forces)
    }

    /// `MolecularDynamics::propagate` with a velocity-Verlet integrator (md/integrators.rs:39-69): `nsteps` whole steps
    /// on the device, then the host arrays are refreshed for outputs.
    pub fn md_run(&mut self, system: &mut System, timestep: f64, nsteps: usize) {
        self.sync(system);
        let ctx = *self.ctx.lock().unwrap();
        check(ctx, unsafe { lumol_cuda_set_velocities(ctx, system.particles().velocity.as_ptr() as *const f64) });
        check(ctx, unsafe { lumol_cuda_md_setup(ctx, LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET, timestep) });
        check(ctx, unsafe { lumol_cuda_md_run(ctx, nsteps as i64) });
        let particles = system.particles_mut();
        check(ctx, unsafe { lumol_cuda_get_positions(ctx, particles.position.as_mut_ptr() as *mut f64) });
        check(ctx, unsafe { lumol_cuda_get_velocities(ctx, particles.velocity.as_mut_ptr() as *mut f64) });
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
