//! Additions to lumol-core that let the shim see through `Box<dyn PairPotential>` and the private fields of
//! `PairInteraction`, `Ewald` and `Wolf`.  Every trait method added here has a default, so user code that implements
//! `Potential` / `CoulombicPotential` keeps compiling; a user potential simply answers `None` and is shipped as a
//! table (energy/computations.rs:102-119).  Not compiled in this image (no rustc): see INTEGRATION.md.
//!
//! Where each block goes is named above it.
use crate::ffi::*;

// ---- lumol-core/src/energy/mod.rs:98-103, inside `pub trait Potential` --------------------------------------------
//
//     /// Closed form known to the device library: (LUMOL_CUDA_POTENTIAL_*, parameters in the header's order).
//     fn device_form(&self) -> Option<(i32, [f64; 5])> { None }
//
// ---- lumol-core/src/energy/functions.rs, one line inside each `impl Potential for ...` -----------------------------
// (parameter orders of include/lumol_cuda.h:45-63; Mie ships the prefactor computed by Mie::new, functions.rs:540-550;
//  CosineHarmonic ships cos(x0), functions.rs:190-195)

pub mod forms {
    use super::*;
    use lumol_core::energy::*;

    pub fn null(_: &NullPotential) -> Option<(i32, [f64; 5])> {
        Some((LUMOL_CUDA_POTENTIAL_NULL, [0.0; 5]))
    }
    pub fn lennard_jones(p: &LennardJones) -> Option<(i32, [f64; 5])> {
        Some((LUMOL_CUDA_POTENTIAL_LJ, [p.sigma, p.epsilon, 0.0, 0.0, 0.0]))
    }
    pub fn harmonic(p: &Harmonic) -> Option<(i32, [f64; 5])> {
        Some((LUMOL_CUDA_POTENTIAL_HARMONIC, [p.k, p.x0, 0.0, 0.0, 0.0]))
    }
    pub fn buckingham(p: &Buckingham) -> Option<(i32, [f64; 5])> {
        Some((LUMOL_CUDA_POTENTIAL_BUCKINGHAM, [p.a, p.c, p.rho, 0.0, 0.0]))
    }
    pub fn born_mayer_huggins(p: &BornMayerHuggins) -> Option<(i32, [f64; 5])> {
        Some((LUMOL_CUDA_POTENTIAL_BMH, [p.a, p.c, p.d, p.sigma, p.rho]))
    }
    pub fn morse(p: &Morse) -> Option<(i32, [f64; 5])> {
        Some((LUMOL_CUDA_POTENTIAL_MORSE, [p.a, p.x0, p.depth, 0.0, 0.0]))
    }
    // the four below read private fields: they live in functions.rs itself
    //   Gaussian        -> (LUMOL_CUDA_POTENTIAL_GAUSSIAN,        [self.a, self.b, 0, 0, 0])
    //   Mie             -> (LUMOL_CUDA_POTENTIAL_MIE,             [self.sigma, self.n, self.m, self.prefac, 0])
    //   CosineHarmonic  -> (LUMOL_CUDA_POTENTIAL_COSINE_HARMONIC, [self.k, self.cos_x0, 0, 0, 0])
    //   Torsion         -> (LUMOL_CUDA_POTENTIAL_TORSION,         [self.k, self.delta, self.n as f64, 0, 0])
}

// ---- lumol-core/src/energy/restrictions.rs, `impl PairRestriction` -------------------------------------------------
pub fn restriction_code(restriction: lumol_core::energy::PairRestriction) -> (i32, f64) {
    use lumol_core::energy::PairRestriction::*;
    match restriction {
        None => (LUMOL_CUDA_RESTRICTION_NONE, 1.0),
        IntraMolecular => (LUMOL_CUDA_RESTRICTION_INTRA_MOLECULAR, 1.0),
        InterMolecular => (LUMOL_CUDA_RESTRICTION_INTER_MOLECULAR, 1.0),
        Exclude12 => (LUMOL_CUDA_RESTRICTION_EXCLUDE12, 1.0),
        Exclude13 => (LUMOL_CUDA_RESTRICTION_EXCLUDE13, 1.0),
        Exclude14 => (LUMOL_CUDA_RESTRICTION_EXCLUDE14, 1.0),
        Scale14(scaling) => (LUMOL_CUDA_RESTRICTION_SCALE14, scaling),
    }
}

// ---- lumol-core/src/energy/pairs.rs, `impl PairInteraction` (reads the private `computation`) ----------------------
//
//     /// Everything the device needs for this (kind, kind) entry.  `tables` receives a tabulated copy of a
//     /// potential without a closed form; its index goes in `record.table`.
//     pub(crate) fn device_record(&self, tables: &mut Vec<(usize, f64, Vec<f64>, Vec<f64>)>) -> lumol_cuda_pair {
//         let mut record = lumol_cuda_pair::default();
//         match self.potential.device_form() {
//             Some((potential, p)) => { record.potential = potential; record.p = p; record.table = -1; }
//             None => {
//                 // same grid as TableComputation::new (computations.rs:102-119): r_i = i * delta, delta = max / size
//                 let size = DEVICE_TABLE_SIZE; let max = self.cutoff; let delta = max / size as f64;
//                 let energy = (0..size).map(|i| self.potential.energy(i as f64 * delta)).collect();
//                 let force  = (0..size).map(|i| self.potential.force(i as f64 * delta)).collect();
//                 record.potential = LUMOL_CUDA_POTENTIAL_TABLE; record.table = tables.len() as i32;
//                 tables.push((size, max, energy, force));
//             }
//         }
//         let (restriction, scale14) = restriction_code(self.restriction);
//         record.restriction = restriction; record.scale14 = scale14;
//         record.cutoff = self.cutoff;
//         record.shift = match self.computation {                       // pairs.rs:9-22, 86-95
//             PairComputation::Cutoff => 0.0,
//             PairComputation::Shifted(shift) => shift,
//         };
//         record.tail_energy = self.tail_energy();                        // pairs.rs:259-274, host-evaluated once
//         record.tail_virial = self.tail_virial()[0][0];                  // pairs.rs:289-296: isotropic, w * identity
//         record
//     }
//
// A `TableComputation` configured by the user (TOML `computation = {table = ...}`) answers `device_form() == None`
// through the default and carries its own (size, max, energy, force): `TableComputation::device_table()` returns
// clones of those four private fields and is preferred over re-tabulating.

// ---- lumol-core/src/energy/global/mod.rs:197-202, inside `pub trait CoulombicPotential` ----------------------------
//
//     /// Describe this solver to the device library; `false` when it has no device implementation.
//     fn device_configure(&self, _ctx: *mut lumol_cuda_context) -> bool { false }
//
// ewald.rs (`impl CoulombicPotential for SharedEwald`, reading `parameters` and `restriction` through the RwLock):
//
//     fn device_configure(&self, ctx: *mut lumol_cuda_context) -> bool {
//         let ewald = self.read();
//         let (restriction, _) = restriction_code(ewald.restriction);
//         let p = &ewald.parameters;                                      // ewald.rs:82-96: alpha, rc, kmax
//         check(ctx, unsafe { lumol_cuda_set_coulomb_ewald(ctx, p.rc, p.alpha, p.kmax as i32, restriction) });
//         true
//     }
//
// `Ewald::with_accuracy` (ewald.rs:312-350) resolves alpha and kmax on the host in `precompute`; the shim calls
// `device_configure` after `SharedEwald::precompute`, so the device only ever sees explicit numbers.
//
// wolf.rs (`impl CoulombicPotential for Wolf`):
//
//     fn device_configure(&self, ctx: *mut lumol_cuda_context) -> bool {
//         let (restriction, scale14) = restriction_code(self.restriction);
//         check(ctx, unsafe { lumol_cuda_set_coulomb_wolf(ctx, self.cutoff, restriction, scale14) });
//         true                                                            // alpha = pi / cutoff inside, wolf.rs:68-84
//     }

// ---- lumol-core/src/sys/system.rs ------------------------------------------------------------------------------------
//
//     /// Bumped by everything that changes what the device mirrors apart from positions, velocities and the cell:
//     /// `add_molecule` (system.rs:80), `Configuration::add_bond` / `remove_molecule`, `particles_mut()` when kinds,
//     /// masses or charges are written, and every `set_*_potential` / `add_global_potential` (system.rs:122-175).
//     pub fn structure_version(&self) -> u64 { self.structure_version }
//
// `device.rs::sync` compares it with the version of its last upload; positions are uploaded at every evaluation (the
// host arrays are the truth in host-driven mode), everything else only when the counter moved.  The Python mirror does
// the same with `System._version` (lumol_b200/sys.py, lumol_b200/device.py).
