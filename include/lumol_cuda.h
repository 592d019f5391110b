/*
 * lumol_cuda.h -- C ABI of the B200 (sm_100a, FP64) force-evaluation library.
 *
 * This is the drop-in boundary for lumol's force/energy/virial hot path.  The
 * reference (lumol-org/lumol, Rust) has no FFI of its own: its stable surface
 * is a set of traits.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference checkout); INTEGRATION.md
 * shows the `extern "C"` block and the trait impls a lumol maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success and a negative lumol_cuda_status on
 *     error; lumol_cuda_last_error() gives the message.  Nothing aborts or
 *     throws across the ABI.
 *   - host pointers are borrowed for the duration of the call only; outputs go
 *     to caller-allocated host buffers; the library owns all device memory.
 *   - Vec<Vector3D> is passed as packed n x 3 doubles (types/vectors.rs:59),
 *     Matrix3 as 9 row-major doubles (types/matrix.rs:73).
 *   - a context is Send but not Sync: one host thread drives it at a time.
 *   - all arithmetic is FP64.  There is no CPU fallback: without a CUDA device
 *     lumol_cuda_create fails with LUMOL_CUDA_ERROR_NO_DEVICE.
 */
#ifndef LUMOL_CUDA_H
#define LUMOL_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUMOL_CUDA_ABI_VERSION 1

typedef struct lumol_cuda_context lumol_cuda_context;

typedef enum {
    LUMOL_CUDA_SUCCESS = 0,
    LUMOL_CUDA_ERROR_INVALID_ARGUMENT = -1,
    LUMOL_CUDA_ERROR_NO_DEVICE = -2,
    LUMOL_CUDA_ERROR_CUDA = -3,
    LUMOL_CUDA_ERROR_STATE = -4,        /* call order: e.g. compute before set_particles */
    LUMOL_CUDA_ERROR_INFINITE_CELL = -5, /* "Ewald is not defined with infinite unit cell" ewald.rs:124;
                                            "Can not compute virial for infinite cell" compute.rs:199 */
    LUMOL_CUDA_ERROR_NOT_FINITE = -6,    /* "Potential energy is infinite!" compute.rs:125 */
    LUMOL_CUDA_ERROR_UNSUPPORTED = -7,
    LUMOL_CUDA_ERROR_COMM = -8
} lumol_cuda_status;

/* CellShape, sys/config/cells.rs:15-22 */
typedef enum { LUMOL_CUDA_CELL_INFINITE = 0, LUMOL_CUDA_CELL_ORTHORHOMBIC = 1, LUMOL_CUDA_CELL_TRICLINIC = 2 } lumol_cuda_cell_shape;

/* Built-in potentials, energy/functions.rs; parameter slots p[0..5]:
 *   NULL        -                                   functions.rs:31-38
 *   LJ          sigma, epsilon                      functions.rs:79-107
 *   HARMONIC    k, x0                               functions.rs:135-156
 *   BUCKINGHAM  a, c, rho                           functions.rs:286-319
 *   BMH         a, c, d, sigma, rho                 functions.rs:353-386
 *   MORSE       a, x0, depth                        functions.rs:414-433
 *   GAUSSIAN    a, b                                functions.rs:474-494
 *   MIE         sigma, n, m, prefactor              functions.rs:538-592 (prefactor from Mie::new)
 *   COSINE_HARMONIC k, cos(x0)                      functions.rs:199-206 (angles/dihedrals only)
 *   TORSION     k, delta, n                         functions.rs:245-255 (dihedrals only)
 *   TABLE       table id in `table`                 energy/computations.rs:70-146
 *   ABSENT      no entry in the interactions map    sys/interactions.rs:142-145
 */
typedef enum {
    LUMOL_CUDA_POTENTIAL_NULL = 0,
    LUMOL_CUDA_POTENTIAL_LJ = 1,
    LUMOL_CUDA_POTENTIAL_HARMONIC = 2,
    LUMOL_CUDA_POTENTIAL_BUCKINGHAM = 3,
    LUMOL_CUDA_POTENTIAL_BMH = 4,
    LUMOL_CUDA_POTENTIAL_MORSE = 5,
    LUMOL_CUDA_POTENTIAL_GAUSSIAN = 6,
    LUMOL_CUDA_POTENTIAL_MIE = 7,
    LUMOL_CUDA_POTENTIAL_COSINE_HARMONIC = 8,
    LUMOL_CUDA_POTENTIAL_TORSION = 9,
    LUMOL_CUDA_POTENTIAL_TABLE = 10,
    LUMOL_CUDA_POTENTIAL_ABSENT = -1
} lumol_cuda_potential_kind;

/* PairRestriction, energy/restrictions.rs:13-33 */
typedef enum {
    LUMOL_CUDA_RESTRICTION_NONE = 0,
    LUMOL_CUDA_RESTRICTION_INTRA_MOLECULAR = 1,
    LUMOL_CUDA_RESTRICTION_INTER_MOLECULAR = 2,
    LUMOL_CUDA_RESTRICTION_EXCLUDE12 = 3,
    LUMOL_CUDA_RESTRICTION_EXCLUDE13 = 4,
    LUMOL_CUDA_RESTRICTION_EXCLUDE14 = 5,
    LUMOL_CUDA_RESTRICTION_SCALE14 = 6
} lumol_cuda_restriction;

/* A Box<dyn Potential> with closed-form parameters (bonds, angles, dihedrals). */
typedef struct {
    int32_t potential; /* lumol_cuda_potential_kind */
    int32_t reserved;
    double p[5];
} lumol_cuda_potential;

/* One PairInteraction (energy/pairs.rs:27-38) for a (kind_i, kind_j) entry. */
typedef struct {
    int32_t potential;   /* lumol_cuda_potential_kind */
    int32_t restriction; /* lumol_cuda_restriction */
    int32_t table;       /* table id from lumol_cuda_add_table when potential == TABLE */
    int32_t reserved;
    double p[5];
    double cutoff;      /* energy/force are 0 for r >= cutoff, pairs.rs:185-218 */
    double shift;       /* PairComputation::Shifted(shift), 0 for plain cutoff; pairs.rs:86-95 */
    double scale14;     /* PairRestriction::Scale14 factor */
    double tail_energy; /* PairInteraction::tail_energy(), host-evaluated, 0 when tails are off; pairs.rs:259-265 */
    double tail_virial; /* scalar potential.tail_virial(cutoff), 0 when off; pairs.rs:289-296 */
} lumol_cuda_pair;

/* Energy terms in the order PotentialEnergy::compute adds them, sys/compute.rs:114-127 */
typedef struct {
    double pairs;          /* EnergyEvaluator::pairs, sys/energy.rs:47-59 */
    double pairs_tail;     /* EnergyEvaluator::pairs_tail, sys/energy.rs:63-79 */
    double bonds;          /* sys/energy.rs:90-99 */
    double angles;         /* sys/energy.rs:109-118 */
    double dihedrals;      /* sys/energy.rs:128-137 */
    double coulomb_real;   /* Ewald::real_space_energy ewald.rs:430-457, or the Wolf pair sum wolf.rs:177-207 */
    double coulomb_self;   /* Ewald::self_energy ewald.rs:619-626, or the Wolf self term wolf.rs:99-101 */
    double coulomb_kspace; /* Ewald::k_space_energy ewald.rs:677-687 */
} lumol_cuda_energy;

/* What lumol_cuda_compute evaluates (bit mask). */
enum {
    LUMOL_CUDA_FORCES = 1,          /* Forces::compute, sys/compute.rs:33-107 */
    LUMOL_CUDA_ENERGY = 2,          /* PotentialEnergy::compute, sys/compute.rs:114-127 */
    LUMOL_CUDA_ATOMIC_VIRIAL = 4,   /* AtomicVirial::compute, sys/compute.rs:198-254 */
    LUMOL_CUDA_MOLECULAR_VIRIAL = 8, /* MolecularVirial::compute, sys/compute.rs:281-363 */
    LUMOL_CUDA_OWNED_FORCES = 16     /* with LUMOL_CUDA_FORCES on a sharded context: only this rank's block of forces (lumol_cuda_owned_range) */
};

/* Which interaction families take part (bit mask); lets the host implement
 * EnergyEvaluator::{pairs,bonds,...,coulomb} and GlobalPotential::{energy,forces,atomic_virial,molecular_virial}
 * for SharedEwald / Wolf alone (energy/global/mod.rs:84-105; benches/nacl.rs:20-45). */
enum {
    LUMOL_CUDA_PART_PAIRS = 1,   /* pair potentials + tail corrections */
    LUMOL_CUDA_PART_BONDED = 2,  /* bonds, angles, dihedrals */
    LUMOL_CUDA_PART_COULOMB = 4, /* Ewald (real + self + k-space) or Wolf */
    LUMOL_CUDA_PART_ALL = 7
};

/* Integrators, lumol-sim/src/md/integrators.rs */
typedef enum {
    LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET = 0, /* integrators.rs:39-70 */
    LUMOL_CUDA_INTEGRATOR_VERLET = 1,          /* integrators.rs:92-123 */
    LUMOL_CUDA_INTEGRATOR_LEAP_FROG = 2,       /* integrators.rs:145-169 */
    LUMOL_CUDA_INTEGRATOR_BERENDSEN_BAROSTAT = 3,       /* integrators.rs:176-255 */
    LUMOL_CUDA_INTEGRATOR_ANISO_BERENDSEN_BAROSTAT = 4  /* integrators.rs:260-341 */
} lumol_cuda_integrator;

/* Thermostats, lumol-sim/src/md/thermostats.rs */
typedef enum {
    LUMOL_CUDA_THERMOSTAT_NONE = 0,
    LUMOL_CUDA_THERMOSTAT_RESCALE = 1,   /* thermostats.rs:66-73, parameter = tolerance */
    LUMOL_CUDA_THERMOSTAT_BERENDSEN = 2, /* thermostats.rs:112-120, parameter = tau */
    LUMOL_CUDA_THERMOSTAT_CSVR = 3       /* thermostats.rs:195-211, parameter = tau; noise from the host RNG */
} lumol_cuda_thermostat;

/* Controls, lumol-sim/src/md/controls.rs (bit mask) */
enum {
    LUMOL_CUDA_CONTROL_REMOVE_TRANSLATION = 1, /* controls.rs:30-41 */
    LUMOL_CUDA_CONTROL_REMOVE_ROTATION = 2,    /* controls.rs:47-73 */
    LUMOL_CUDA_CONTROL_REWRAP = 4              /* controls.rs:80-87 */
};

/* DegreesOfFreedom, sys/system.rs:249-255 */
typedef enum { LUMOL_CUDA_DOF_PARTICLES = 0, LUMOL_CUDA_DOF_MOLECULES = 1 } lumol_cuda_dof;

/* Counters for the measurement harness ("metrics"). */
typedef struct {
    int64_t natoms;
    int64_t kernel_launches;  /* kernels launched by this context since creation / last reset */
    int64_t neighbor_path;    /* 0: tiled all-pairs minimum image, 1: cell list */
    int64_t ncells[3];
    int64_t nkvectors;
    int64_t pair_launches;    /* launches of the dominant pair kernel */
    double pair_ms;           /* accumulated device time of that kernel (CUDA events), when profiling is on */
    int64_t kspace_launches;
    double kspace_ms;
    int64_t integrate_launches;
    double integrate_ms;
    int64_t neighbor_launches;
    double neighbor_ms;
    int64_t comm_launches;
    double comm_ms;
    double pair_count;         /* pairs inside their cut-off at the last energy/virial evaluation (each once) */
    double coulomb_pair_count; /* same for the coulomb real-space / Wolf term */
    int64_t neighbor_rebuilds; /* neighbour-list rebuilds since the context was created */
    double neighbor_skin;      /* skin in use (may be smaller than requested in small boxes) */
} lumol_cuda_stats;

/* ---- lifetime ----------------------------------------------------------------------------------- */
int32_t lumol_cuda_abi_version(void);
/* device: CUDA ordinal.  Fails (no fallback) when there is no usable device. */
int32_t lumol_cuda_create(int32_t device, lumol_cuda_context** ctx);
/* One context over several devices of THIS process (a single lumol process drives all the GPUs of the node; SURVEY
 * section 8b: lumol has one `System` and one thread of control).  The library keeps one sharded context and one host
 * thread per device and connects them with NCCL and NVLink peer memory, exactly as lumol_cuda_comm_init does between
 * processes; every other entry point accepts the returned context unchanged, takes whole-system arrays and returns
 * whole-system results (forces are downloaded block by block from the device that owns them).  Not available on such a
 * context: lumol_cuda_comm_init, the barostats and the molecular virial (as for any sharded context).
 * ndevices == 1 is lumol_cuda_create(devices[0]). */
int32_t lumol_cuda_create_multi(const int32_t* devices, int32_t ndevices, lumol_cuda_context** ctx);
int32_t lumol_cuda_destroy(lumol_cuda_context* ctx);
/* Message of the last error on this context (or of the last failed create when ctx is NULL). */
const char* lumol_cuda_last_error(const lumol_cuda_context* ctx);

/* ---- state upload (Configuration, sys/config/configuration.rs:42-51) ------------------------------ */
/* UnitCell: row-major cell matrix whose columns are the lattice vectors (cells.rs:31-39, 234-255). */
int32_t lumol_cuda_set_cell(lumol_cuda_context* ctx, const double cell[9], int32_t shape);
/* ParticleVec field vectors (particles.rs:31-46).  velocity may be NULL (zeros). */
int32_t lumol_cuda_set_particles(lumol_cuda_context* ctx, int64_t n, const double* position, const double* velocity,
                                 const double* mass, const double* charge, const uint32_t* kind);
int32_t lumol_cuda_set_positions(lumol_cuda_context* ctx, const double* position);
/* Sharded contexts (lumol_cuda_comm_init): upload only the positions of this rank's block of atoms (count x 3 doubles,
 * lumol_cuda_owned_range); the ranks exchange their blocks on the device over NVLink.  Collective: every rank calls it.
 * Host-driven steps then move 24 B per owned atom over each PCIe link instead of 24 B per atom.  On an unsharded context
 * it is lumol_cuda_set_positions. */
int32_t lumol_cuda_set_owned_positions(lumol_cuda_context* ctx, const double* owned_position);
int32_t lumol_cuda_set_velocities(lumol_cuda_context* ctx, const double* velocity);
int32_t lumol_cuda_get_positions(lumol_cuda_context* ctx, double* position);
int32_t lumol_cuda_get_velocities(lumol_cuda_context* ctx, double* velocity);
/* Forces of the last compute / MD step, n x 3. */
int32_t lumol_cuda_get_forces(lumol_cuda_context* ctx, double* forces);
/* Molecules are contiguous atom ranges [start[m], start[m+1]) (configuration.rs:178-234);
 * bond_distances holds, for molecule m, a size x size matrix of BondDistances bytes
 * (connect.rs:142-165, bonding.rs:130-155) at byte offset bond_distances_offset[m] (molecules of the
 * same type may share one matrix).  nmol == 0 resets to one molecule per atom. */
int32_t lumol_cuda_set_molecules(lumol_cuda_context* ctx, int64_t nmol, const uint64_t* start,
                                 const uint64_t* bond_distances_offset, const uint8_t* bond_distances,
                                 uint64_t bond_distances_size);

/* ---- interactions (Interactions, sys/interactions.rs:62-77) ---------------------------------------- */
/* nkinds x nkinds symmetric table indexed by ParticleKind (system.rs:178-182). */
int32_t lumol_cuda_set_pairs(lumol_cuda_context* ctx, int32_t nkinds, const lumol_cuda_pair* pairs);
/* TableComputation::new(potential, size, max) tables built on the host (computations.rs:102-119).
 * Returns the table id (>= 0) or a negative status. */
int32_t lumol_cuda_add_table(lumol_cuda_context* ctx, int32_t size, double max, const double* energy,
                             const double* force);
int32_t lumol_cuda_clear_tables(lumol_cuda_context* ctx);
/* Bonded terms: explicit index lists (Bonding::bonds/angles/dihedrals, bonding.rs:254-266) and, per
 * entry, an index into `potentials` (-1: no potential for these kinds, interactions.rs:147-160). */
int32_t lumol_cuda_set_bonded_potentials(lumol_cuda_context* ctx, int32_t npotentials,
                                         const lumol_cuda_potential* potentials);
int32_t lumol_cuda_set_bonds(lumol_cuda_context* ctx, int64_t n, const int64_t* atoms /* n x 2 */,
                             const int32_t* potential);
int32_t lumol_cuda_set_angles(lumol_cuda_context* ctx, int64_t n, const int64_t* atoms /* n x 3 */,
                              const int32_t* potential);
int32_t lumol_cuda_set_dihedrals(lumol_cuda_context* ctx, int64_t n, const int64_t* atoms /* n x 4 */,
                                 const int32_t* potential);
/* CoulombicPotential (energy/global/mod.rs:197-202) */
int32_t lumol_cuda_set_coulomb_none(lumol_cuda_context* ctx);
/* Ewald::new(cutoff, kmax, alpha) + set_restriction (ewald.rs:278-305, 952-956).  The k-vector factor
 * table is rebuilt inside the library whenever the cell changes (Ewald::prepare, ewald.rs:353-378). */
int32_t lumol_cuda_set_coulomb_ewald(lumol_cuda_context* ctx, double cutoff, double alpha, int32_t kmax,
                                     int32_t restriction);
/* Wolf::new(cutoff) + set_restriction (wolf.rs:68-84, 327-331) */
int32_t lumol_cuda_set_coulomb_wolf(lumol_cuda_context* ctx, double cutoff, int32_t restriction, double scale14);

/* ---- evaluation (Compute, sys/compute.rs:21-26) ---------------------------------------------------- */
/* One fused pass over the device-resident positions.  forces (n x 3) is OVERWRITTEN with the sum of
 * the requested parts; energy and virial likewise.  Any of forces/energy/virial may be NULL.  When
 * `what` has both virial bits the atomic virial is returned. */
int32_t lumol_cuda_compute(lumol_cuda_context* ctx, uint32_t what, uint32_t parts, double* forces,
                           lumol_cuda_energy* energy, double virial[9]);
/* Sharded contexts (lumol_cuda_comm_init): the block of atoms, in the caller's order, whose forces this rank evaluates in
 * lumol_cuda_compute.  With LUMOL_CUDA_OWNED_FORCES in `what`, `forces` receives only those count x 3 values (no
 * all-gather between the ranks, count x 24 bytes to the host instead of n x 24): what a host-side integrator that is
 * itself sharded over processes needs (the reference has no such notion: its Forces::compute fills one Vec, compute.rs:33). */
int32_t lumol_cuda_owned_range(lumol_cuda_context* ctx, int64_t* first, int64_t* count);
/* KineticEnergy (compute.rs:134-144), and sum_i m_i v_i (x) v_i for Stress (compute.rs:471-474) */
int32_t lumol_cuda_kinetic_energy(lumol_cuda_context* ctx, double* kinetic);
int32_t lumol_cuda_kinetic_tensor(lumol_cuda_context* ctx, double tensor[9]);
/* rho(k) and the k-vector table of the last Ewald evaluation (ewald.rs:633-674), for tests. */
int32_t lumol_cuda_ewald_kvectors(lumol_cuda_context* ctx, int64_t capacity, int64_t* count, int32_t* index /* 3 per k */,
                                  double* energy_factor, double* rho /* 2 per k */);

/* ---- Monte Carlo energy cache (EnergyCache, sys/cache.rs; GlobalCache, energy/global/mod.rs:169-189) --------- */
/* EnergyCache::move_molecule_cost (cache.rs:145-213) with the coulomb part of SharedEwald / Wolf
 * (ewald.rs:572-613, 758-839, 933-946; wolf.rs:121-165): energy change of moving the rigid molecule `molecule`
 * to new_positions (its size x 3), by term: cost->pairs, cost->coulomb_real, cost->coulomb_kspace are new minus old,
 * the other fields 0 (tail, bonded and self terms do not change); the cost is their sum.  The resident state is
 * NOT changed.  The old pair energies the reference reads from its N x N pairs_cache are re-evaluated from the
 * resident positions; rho(k) is the one of the resident positions (recomputed only when they changed since the last
 * Ewald evaluation). */
int32_t lumol_cuda_move_molecule_cost(lumol_cuda_context* ctx, int64_t molecule, const double* new_positions,
                                      lumol_cuda_energy* cost);
/* The same for `ntrials` independent trial moves, each against the same resident state, in one batch of launches:
 * new_positions holds the new positions of molecules[0], molecules[1], ... one after the other. */
int32_t lumol_cuda_move_molecules_cost(lumol_cuda_context* ctx, int64_t ntrials, const int64_t* molecules,
                                       const double* new_positions, lumol_cuda_energy* costs);
/* EnergyCache::update after an accepted move (cache.rs:115-128, 175-211) plus what the reference's move does to the
 * system (mc/moves/translate.rs:115-120): the molecule of trial `trial` of the last cost call takes its new
 * positions on the device and rho(k) += delta rho(k) (ewald.rs:833-837).  Fails with LUMOL_CUDA_ERROR_STATE and the
 * reference's message when no cost call is pending or the resident positions changed since. */
int32_t lumol_cuda_move_molecule_accept(lumol_cuda_context* ctx, int64_t trial);

/* ---- device-resident molecular dynamics (lumol-sim/src/md) ------------------------------------------ */
/* Integrator::setup (integrators.rs:40-42, 92-101, 146-148) */
int32_t lumol_cuda_md_setup(lumol_cuda_context* ctx, int32_t integrator, double timestep);
int32_t lumol_cuda_md_set_degrees_of_freedom(lumol_cuda_context* ctx, int32_t mode, int64_t frozen);
int32_t lumol_cuda_md_set_thermostat(lumol_cuda_context* ctx, int32_t thermostat, double temperature, double parameter);
/* CSVR noise for the next `nsteps` steps: (gauss, wiener) pairs from the host RNG (thermostats.rs:175-193) */
int32_t lumol_cuda_md_set_csvr_noise(lumol_cuda_context* ctx, int64_t nsteps, const double* noise /* 2 per step */);
int32_t lumol_cuda_md_set_controls(lumol_cuda_context* ctx, uint32_t controls);
/* MolecularDynamics::propagate x nsteps (molecular_dynamics.rs:66-76) with no host round trip. */
int32_t lumol_cuda_md_run(lumol_cuda_context* ctx, int64_t nsteps);
/* velocities::scale / thermostat scaling (velocities.rs:16-22), RemoveTranslation (controls.rs:30-41) */
int32_t lumol_cuda_scale_velocities(lumol_cuda_context* ctx, double factor);
int32_t lumol_cuda_remove_translation(lumol_cuda_context* ctx);
/* RemoveRotation::control and Rewrap::control (controls.rs:47-87) on the resident state. */
int32_t lumol_cuda_remove_rotation(lumol_cuda_context* ctx);
int32_t lumol_cuda_rewrap(lumol_cuda_context* ctx);
/* BerendsenBarostat / AnisoBerendsenBarostat (integrators.rs:176-342), selected by lumol_cuda_md_setup: target is
 * the stress matrix (row-major; the isotropic barostat reads target[0] as its pressure), tau the time scale in units
 * of the timestep.  Every step scales positions and cell by eta, evaluates pressure (or stress) and forces in one
 * pass at the new positions, and updates eta; positions and velocities stay on the device, the host sees nine sums per
 * step.  lumol_cuda_md_run fails with the reference's message when the cell shrinks below twice the cut-off. */
int32_t lumol_cuda_md_set_barostat(lumol_cuda_context* ctx, const double target[9], double tau);
/* The cell matrix as it is now (the barostats change it). */
int32_t lumol_cuda_get_cell(lumol_cuda_context* ctx, double cell[9]);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ----------------------------------------------- */
/* Rank 0 creates the id, the host launcher broadcasts it, every rank calls comm_init.  Afterwards each
 * rank owns a contiguous block of atoms for pair forces / integration, and the Ewald structure factor,
 * energies and virials are all-reduced. */
int32_t lumol_cuda_comm_unique_id(uint8_t id[128]);
int32_t lumol_cuda_comm_init(lumol_cuda_context* ctx, int32_t nranks, int32_t rank, const uint8_t id[128]);

/* ---- measurement ---------------------------------------------------------------------------------- */
int32_t lumol_cuda_set_profiling(lumol_cuda_context* ctx, int32_t enabled);
int32_t lumol_cuda_get_stats(lumol_cuda_context* ctx, lumol_cuda_stats* stats);
int32_t lumol_cuda_reset_stats(lumol_cuda_context* ctx);
/* Force the neighbour search path: -1 automatic, 0 all-pairs, 1 cell list (error when a cell edge has < 3 cells),
 * 2 cell list whose Lennard-Jones blocks all keep the global list format (no shared-memory staging; used to
 * cross-check the staged kernel). */
int32_t lumol_cuda_set_neighbor_path(lumol_cuda_context* ctx, int32_t path);
/* Verlet skin of the neighbour list (default 1 A): the list is rebuilt, on the device, when an atom has moved more
 * than skin / 2 since the last build.  Results do not depend on it: every listed pair is re-tested against its
 * cut-off in FP64 at every evaluation. */
int32_t lumol_cuda_set_neighbor_skin(lumol_cuda_context* ctx, double skin);
/* Kernels of the Ewald reciprocal-space sums (eik_dot_r and k_space_*, ewald.rs:633-733): -1 automatic (tiled for
 * large N x Nk), 0 direct (one thread per k-vector / per atom), 1 tiled (register-tiled contractions sharing the
 * +l / -l products).  Same sums, different summation order. */
int32_t lumol_cuda_set_kspace_algorithm(lumol_cuda_context* ctx, int32_t algorithm);
/* cudaStream_t the context launches on, for event timing by the harness. */
void* lumol_cuda_stream(lumol_cuda_context* ctx);
int32_t lumol_cuda_synchronize(lumol_cuda_context* ctx);
/* Measured FP64 FMA throughput of the device in TFLOP/s (dependent-free DFMA chains on every SM), the
 * denominator of the FP64 roofline; and a device-to-device copy bandwidth in GB/s. */
int32_t lumol_cuda_measure_fp64_peak(lumol_cuda_context* ctx, double* tflops);
int32_t lumol_cuda_measure_copy_bandwidth(lumol_cuda_context* ctx, double* gbs);

#ifdef __cplusplus
}
#endif

#endif /* LUMOL_CUDA_H */
