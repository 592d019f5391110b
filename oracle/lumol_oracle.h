/*
 * lumol_oracle.h -- CPU restatement of lumol's force-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product (lumol_b200/, include/); only tests/, the smoke test
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py use it, as the checker and the CPU baseline.
 *
 * Parity status: PINNED.  The reference (Rust) cannot be compiled in this
 * image (no cargo/rustc), so this file restates the reference's arithmetic in
 * plain C, function by function, in the reference's own evaluation order
 * (every function cites the file:line it follows, paths relative to the
 * reference checkout), and tests/test_oracle_kat.py checks it against every
 * exact-equality known answer, NIST value and LAMMPS force dump the
 * reference's own tests hold for this path.
 *
 * Third-party arithmetic: erf/erfc come from the un-vendored crate
 * `special = "0.10"` in the reference (lumol-core/src/math.rs:8-18); glibc's
 * erf/erfc are used here and reproduce the reference's exact-equality tests
 * at that boundary (wolf.rs:29-48).
 *
 * Compile with -ffp-contract=off: Rust never contracts a*b+c into an FMA.
 */
#ifndef LUMOL_ORACLE_H
#define LUMOL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* consts.rs:9-15 */
#define ORC_K_BOLTZMANN 8.31446284161522e-7
#define ORC_FOUR_PI_EPSILON_0 7.197589831304046

/* Potential identifiers (energy/functions.rs) and their parameter slots. */
enum {
    ORC_POT_NULL = 0,            /* functions.rs:31-38                                */
    ORC_POT_LJ = 1,              /* p = {sigma, epsilon}          functions.rs:79-107 */
    ORC_POT_HARMONIC = 2,        /* p = {k, x0}                   functions.rs:135-156 */
    ORC_POT_BUCKINGHAM = 3,      /* p = {a, c, rho}               functions.rs:286-319 */
    ORC_POT_BMH = 4,             /* p = {a, c, d, sigma, rho}     functions.rs:353-386 */
    ORC_POT_MORSE = 5,           /* p = {a, x0, depth}            functions.rs:414-433 */
    ORC_POT_GAUSSIAN = 6,        /* p = {a, b}                    functions.rs:474-494 */
    ORC_POT_MIE = 7,             /* p = {sigma, n, m, prefac}     functions.rs:538-592 */
    ORC_POT_COSINE_HARMONIC = 8, /* p = {k, cos_x0}               functions.rs:199-206 */
    ORC_POT_TORSION = 9,         /* p = {k, delta, n}             functions.rs:245-255 */
    ORC_POT_ABSENT = -1          /* no entry in the interactions map (interactions.rs:142-145) */
};

/* energy/restrictions.rs:13-33 */
enum {
    ORC_RESTRICT_NONE = 0,
    ORC_RESTRICT_INTRA = 1,
    ORC_RESTRICT_INTER = 2,
    ORC_RESTRICT_EXCLUDE12 = 3,
    ORC_RESTRICT_EXCLUDE13 = 4,
    ORC_RESTRICT_EXCLUDE14 = 5,
    ORC_RESTRICT_SCALE14 = 6
};

/* energy/restrictions.rs:37-50 */
enum {
    ORC_PATH_NONE = 0,
    ORC_PATH_SAME = 1,
    ORC_PATH_ONE = 2,
    ORC_PATH_TWO = 3,
    ORC_PATH_THREE = 4,
    ORC_PATH_FAR = 5
};

/* sys/config/cells.rs:15-22 */
enum { ORC_CELL_INFINITE = 0, ORC_CELL_ORTHO = 1, ORC_CELL_TRICLINIC = 2 };

/* sys/config/connect.rs:142-165 */
enum { ORC_BOND_ONE = 1, ORC_BOND_TWO = 2, ORC_BOND_THREE = 4, ORC_BOND_FAR = 8 };

enum { ORC_COULOMB_NONE = 0, ORC_COULOMB_EWALD = 1, ORC_COULOMB_WOLF = 2 };

typedef struct {
    int32_t pot;
    int32_t _pad;
    double p[5];
} orc_potential;

/* energy/pairs.rs:27-38 PairInteraction (+ energy/computations.rs:70-80 when table_n > 0) */
typedef struct {
    orc_potential potential; /* pot == ORC_POT_ABSENT: no interaction for this kind pair   */
    double cutoff;
    int32_t shifted;         /* PairComputation::Shifted; shift = energy(cutoff), pairs.rs:86-95 */
    int32_t tail;            /* enable_tail_corrections, pairs.rs:112-114 */
    int32_t restriction;
    int32_t table_n;         /* > 0: wrapped in TableComputation::new(potential, table_n, table_max) */
    double scale14;
    double table_max;
    const double* table_energy; /* table_n entries, from orc_table_build */
    const double* table_force;
} orc_pair;

typedef struct {
    int64_t n;
    const double* position; /* n x 3, Vec<Vector3D> layout (types/vectors.rs:59) */
    const double* velocity; /* n x 3 */
    const double* mass;
    const double* charge;
    const uint32_t* kind;
    double cell[9]; /* row-major Matrix3, columns are the lattice vectors (cells.rs:234-255) */
    int32_t shape;
    int32_t nkinds;
    /* molecules: contiguous atom ranges (configuration.rs:42-51) */
    int64_t nmol;
    const int64_t* mol_start;     /* nmol + 1 */
    const int64_t* molid;         /* n */
    const uint8_t* bond_dist;     /* concatenated k x k BondDistances bytes */
    const int64_t* bond_dist_off; /* nmol offsets into bond_dist */
    const orc_pair* pairs;        /* nkinds x nkinds, symmetric */
    /* bonded terms, in the order the caller enumerates them */
    int64_t nbonds;
    const int64_t* bonds; /* nbonds x 2 */
    const orc_potential* bond_pot;
    int64_t nangles;
    const int64_t* angles; /* nangles x 3 */
    const orc_potential* angle_pot;
    int64_t ndihedrals;
    const int64_t* dihedrals; /* ndihedrals x 4 */
    const orc_potential* dihedral_pot;
    /* coulomb */
    int32_t coulomb;
    int32_t coulomb_restriction;
    double coulomb_scale14;
    double rc;
    double alpha;
    int32_t kmax;
    int32_t dof_mode; /* 0 particles, 1 molecules; frozen in dof_frozen (system.rs:249-255) */
    int64_t dof_frozen;
} orc_system;

/* energy terms in the order of PotentialEnergy::compute (compute.rs:114-127) */
typedef struct {
    double pairs;
    double pairs_tail;
    double bonds;
    double angles;
    double dihedrals;
    double coulomb_real; /* Ewald real space, or the whole Wolf sum */
    double coulomb_self;
    double coulomb_kspace;
} orc_energy_terms;

/* ---- potentials -------------------------------------------------------- */
double orc_potential_energy(const orc_potential* pot, double r);
double orc_potential_force(const orc_potential* pot, double r);
double orc_potential_tail_energy(const orc_potential* pot, double rc);
double orc_potential_tail_virial(const orc_potential* pot, double rc);
double orc_mie_prefactor(double epsilon, double n, double m);
void orc_potential_virial(const orc_potential* pot, const double r[3], double w[9]);

/* ---- TableComputation -------------------------------------------------- */
void orc_table_build(const orc_potential* pot, int32_t size, double max, double* energy, double* force);
double orc_table_energy(const double* table, int32_t size, double max, double r);

/* ---- PairInteraction --------------------------------------------------- */
double orc_pair_energy(const orc_pair* pair, double r);
double orc_pair_force(const orc_pair* pair, double r);
void orc_pair_virial(const orc_pair* pair, const double r[3], double w[9]);
double orc_pair_tail_energy(const orc_pair* pair);
double orc_pair_tail_virial(const orc_pair* pair); /* scalar; the tensor is this * identity / 3 */

/* ---- restrictions / topology -------------------------------------------- */
void orc_restriction_information(int32_t restriction, double scale14, int32_t path, int32_t* excluded, double* scaling);
int32_t orc_bond_path(const orc_system* s, int64_t i, int64_t j);
/* Bonding::rebuild + rebuild_connections for one molecule of `natoms` atoms with local bond indices.
 * Returns counts through nangles/ndihedrals; angles/dihedrals buffers may be NULL to only count. */
void orc_bonding_rebuild(int64_t natoms, int64_t nbonds, const int64_t* bonds, int64_t* nangles, int64_t* angles,
                         int64_t* ndihedrals, int64_t* dihedrals, uint8_t* distances);

/* ---- cell geometry ----------------------------------------------------- */
void orc_matrix_inverse(const double m[9], double inv[9]);
void orc_vector_image(const double cell[9], int32_t shape, double v[3]);
void orc_wrap_vector(const double cell[9], int32_t shape, double v[3]);
double orc_cell_volume(const double cell[9], int32_t shape);
void orc_cell_lengths(const double cell[9], int32_t shape, double lengths[3]);
void orc_k_vector(const double cell[9], const double index[3], double k[3]);
double orc_angle_and_derivatives(const double cell[9], int32_t shape, const double* r1, const double* r2,
                                 const double* r3, double d1[3], double d2[3], double d3[3]);
double orc_dihedral_and_derivatives(const double cell[9], int32_t shape, const double* r1, const double* r2,
                                    const double* r3, const double* r4, double d1[3], double d2[3], double d3[3],
                                    double d4[3]);

/* ---- estimators (sys/compute.rs, sys/energy.rs) -------------------------- */
void orc_set_threads(int32_t nthreads);
int32_t orc_get_threads(void);
void orc_pair_forces(const orc_system* s, double* forces);   /* compute.rs:37-60, zero-initialised output */
void orc_bonded_forces(const orc_system* s, double* forces); /* compute.rs:62-97, accumulates */
int64_t orc_pair_forces_sample(const orc_system* s, int64_t nrows, const int64_t* rows, double* checksum);
/* total pair force on selected atoms (every j != i), for boxes too large for the full loop; out: nrows x 3 */
void orc_pair_forces_rows(const orc_system* s, int64_t nrows, const int64_t* rows, double* out);
void orc_coulomb_forces(const orc_system* s, double* forces); /* accumulates, like GlobalPotential::forces */
void orc_forces(const orc_system* s, double* forces);         /* Forces::compute */
double orc_pairs_energy(const orc_system* s);
double orc_pairs_tail_energy(const orc_system* s);
void orc_energy_terms_compute(const orc_system* s, orc_energy_terms* out);
double orc_potential_energy_total(const orc_system* s);
void orc_pair_atomic_virial(const orc_system* s, double w[9]);
void orc_tail_virial(const orc_system* s, double w[9]);
void orc_bond_virial(const orc_system* s, double w[9]);
void orc_coulomb_atomic_virial(const orc_system* s, double w[9]);
void orc_coulomb_molecular_virial(const orc_system* s, double w[9]);
void orc_atomic_virial(const orc_system* s, double w[9]);
void orc_molecular_virial(const orc_system* s, double w[9]);
double orc_kinetic_energy(const orc_system* s);
int64_t orc_degrees_of_freedom(const orc_system* s);
double orc_temperature(const orc_system* s);
double orc_pressure_at_temperature(const orc_system* s, double temperature);
double orc_pressure(const orc_system* s);
void orc_stress_at_temperature(const orc_system* s, double temperature, double out[9]);
void orc_stress(const orc_system* s, double out[9]);

/* ---- Ewald / Wolf components (energy/global/ewald.rs, wolf.rs) ------------ */
int64_t orc_ewald_factors(const orc_system* s, double* kmax2_out, int64_t capacity, int64_t* index /*3 per k*/,
                          double* energy, double* field /*3 per k*/, double* virial /*9 per k*/);
double orc_ewald_real_energy(const orc_system* s);
double orc_ewald_self_energy(const orc_system* s);
double orc_ewald_kspace_energy(const orc_system* s);
void orc_ewald_real_forces(const orc_system* s, double* forces);
void orc_ewald_real_forces_rows(const orc_system* s, int64_t nrows, const int64_t* rows, double* out);
void orc_ewald_kspace_forces(const orc_system* s, double* forces);
void orc_ewald_real_atomic_virial(const orc_system* s, double w[9]);
void orc_ewald_kspace_atomic_virial(const orc_system* s, double w[9]);
void orc_ewald_rho(const orc_system* s, int64_t capacity, double* rho /*2 per k*/);
void orc_ewald_with_accuracy(const orc_system* s, double cutoff, double accuracy, double* alpha, int32_t* kmax);
double orc_wolf_energy(const orc_system* s);
void orc_wolf_forces(const orc_system* s, double* forces);
void orc_wolf_atomic_virial(const orc_system* s, double w[9]);

/* ---- integrators / thermostats / controls (lumol-sim/src/md) --------------- */
/* All of these mutate position/velocity arrays passed separately so that orc_system stays const. */
void orc_velocity_verlet_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt);
void orc_verlet_setup(const orc_system* s, double* prevpos, double dt);
void orc_verlet_step(orc_system* s, double* position, double* velocity, double* prevpos, double dt);
void orc_leapfrog_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt);
void orc_scale_velocities(int64_t n, double* velocity, double factor);
double orc_berendsen_thermostat_factor(double temperature, double instant, double tau);
double orc_rescale_thermostat_factor(double temperature, double instant);
void orc_remove_translation(int64_t n, const double* mass, double* velocity);
/* controls.rs:47-87 */
void orc_remove_rotation(int64_t n, const double* mass, const double* position, double* velocity);
void orc_rewrap(const orc_system* s, double* position);
/* integrators.rs:211-255, 295-341: one step; return 1 where the reference panics (cell smaller than the cut-off) */
int orc_berendsen_barostat_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt,
                                double pressure, double tau, double* eta, double maximum_cutoff);
int orc_aniso_berendsen_barostat_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt,
                                      const double stress[9], double tau, double eta[9], double maximum_cutoff);

/* ---- Monte Carlo energy cache (sys/cache.rs:145-283; ewald.rs:572-613, 758-839; wolf.rs:121-165) -------- */
/* new_positions: size-of-molecule x 3.  The old pair energies / phases / rho(k) the reference reads from its caches
 * are re-evaluated on the spot from the system's positions (the caches are taken to be current). */
double orc_move_molecule_pairs_cost(const orc_system* s, int64_t molecule, const double* new_positions);
double orc_ewald_real_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions);
double orc_ewald_kspace_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions,
                                           int64_t capacity, double* delta_rho /* 2 per k, may be NULL */);
double orc_wolf_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions);
/* out = {pairs, coulomb real space (or Wolf), coulomb k-space}; EnergyCache::move_molecule_cost is their sum */
void orc_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions, double out[3]);
/* out = {inter-molecular pairs, pairs_tail, coulomb} differences; EnergyCache::move_all_molecules_cost is their sum */
void orc_move_all_molecules_cost(const orc_system* before, const orc_system* after, double out[3]);

#ifdef __cplusplus
}
#endif

#endif /* LUMOL_ORACLE_H */
