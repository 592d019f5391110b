/*
 * lumol_oracle.c -- CPU restatement of lumol's force-evaluation hot path.
 * TEST INFRASTRUCTURE ONLY; see lumol_oracle.h for the rules and the parity status.
 *
 * Every function follows the reference's evaluation order (cited file:line,
 * relative to the reference checkout).  Loops that the reference runs under
 * rayon are run under OpenMP here with the same per-thread accumulation
 * scheme (utils/thread_vec.rs:13-56), so the sum order differs from a given
 * rayon run only in the way two rayon runs differ from each other.
 */
#include "lumol_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846264338327950288
/* std::f64::consts::FRAC_2_SQRT_PI */
#define FRAC_2_SQRT_PI 1.12837916709551257389615890312154517

static int32_t g_threads = 0;

void orc_set_threads(int32_t nthreads) {
    g_threads = nthreads;
#ifdef _OPENMP
    if (nthreads > 0) {
        omp_set_num_threads(nthreads);
    }
#endif
}

int32_t orc_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

static int max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int thread_id(void) {
#ifdef _OPENMP
    return omp_get_thread_num();
#else
    return 0;
#endif
}

/* ------------------------------------------------------------------------ */
/* small vector / matrix helpers (types/vectors.rs, types/matrix.rs)          */
/* ------------------------------------------------------------------------ */

/* vectors.rs:215-218 dot product, left to right */
static inline double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* vectors.rs:102-117 */
static inline double norm2_3(const double a[3]) { return dot3(a, a); }
static inline double norm3(const double a[3]) { return sqrt(norm2_3(a)); }

/* vectors.rs:221-228 */
static inline void cross3(const double a[3], const double b[3], double out[3]) {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    out[0] = x;
    out[1] = y;
    out[2] = z;
}

/* matrix.rs:391-399 */
static inline void matvec(const double m[9], const double v[3], double out[3]) {
    double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    out[0] = x;
    out[1] = y;
    out[2] = z;
}

/* matrix.rs:245-249 */
static double determinant(const double m[9]) {
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* matrix.rs:212-227 */
void orc_matrix_inverse(const double m[9], double res[9]) {
    double inverse_determinant = 1.0 / determinant(m);
    res[0] = (m[4] * m[8] - m[7] * m[5]) * inverse_determinant;
    res[1] = (m[2] * m[7] - m[1] * m[8]) * inverse_determinant;
    res[2] = (m[1] * m[5] - m[2] * m[4]) * inverse_determinant;
    res[3] = (m[5] * m[6] - m[3] * m[8]) * inverse_determinant;
    res[4] = (m[0] * m[8] - m[2] * m[6]) * inverse_determinant;
    res[5] = (m[3] * m[2] - m[0] * m[5]) * inverse_determinant;
    res[6] = (m[3] * m[7] - m[6] * m[4]) * inverse_determinant;
    res[7] = (m[6] * m[1] - m[0] * m[7]) * inverse_determinant;
    res[8] = (m[0] * m[4] - m[3] * m[1]) * inverse_determinant;
}

/* ------------------------------------------------------------------------ */
/* potentials (energy/functions.rs)                                          */
/* ------------------------------------------------------------------------ */

/* f64::powi(x, 6) as LLVM expands it: square-and-multiply, x^2 * x^4 */
static inline double powi6(double x) {
    double x2 = x * x;
    double x4 = x2 * x2;
    return x2 * x4;
}

/* functions.rs:540-551 Mie::new */
double orc_mie_prefactor(double epsilon, double n, double m) { return n / (n - m) * pow(n / m, m / (n - m)) * epsilon; }

double orc_potential_energy(const orc_potential* pot, double r) {
    const double* p = pot->p;
    switch (pot->pot) {
    case ORC_POT_LJ: { /* functions.rs:80-83 */
        double s6 = powi6(p[0] / r);
        return 4.0 * p[1] * (s6 * s6 - s6);
    }
    case ORC_POT_HARMONIC: { /* functions.rs:136-139 */
        double dx = r - p[1];
        return 0.5 * p[0] * dx * dx;
    }
    case ORC_POT_BUCKINGHAM: { /* functions.rs:287-292 */
        double r3 = r * r * r;
        double r6 = r3 * r3;
        double e = exp(-r / p[2]);
        return p[0] * e - p[1] / r6;
    }
    case ORC_POT_BMH: { /* functions.rs:354-359 */
        double r2 = r * r;
        double r6 = r2 * r2 * r2;
        double e = exp((p[3] - r) / p[4]);
        return p[0] * e - p[1] / r6 + p[2] / (r6 * r2);
    }
    case ORC_POT_MORSE: { /* functions.rs:415-418 */
        double rc = 1.0 - exp((p[1] - r) * p[0]);
        return p[2] * rc * rc;
    }
    case ORC_POT_GAUSSIAN: /* functions.rs:475-477 */
        return -p[0] * exp(-p[1] * r * r);
    case ORC_POT_MIE: { /* functions.rs:553-558 */
        double sigma_r = p[0] / r;
        double repulsive = pow(sigma_r, p[1]);
        double attractive = pow(sigma_r, p[2]);
        return p[3] * (repulsive - attractive);
    }
    case ORC_POT_COSINE_HARMONIC: { /* functions.rs:199-202 */
        double dr = cos(r) - p[1];
        return 0.5 * p[0] * dr * dr;
    }
    case ORC_POT_TORSION: { /* functions.rs:245-249 */
        double c = cos(p[2] * r - p[1]);
        return p[0] * (1.0 + c);
    }
    default: /* NullPotential, functions.rs:31-38 */
        return 0.0;
    }
}

double orc_potential_force(const orc_potential* pot, double r) {
    const double* p = pot->p;
    switch (pot->pot) {
    case ORC_POT_LJ: { /* functions.rs:85-88 */
        double s6 = powi6(p[0] / r);
        return -24.0 * p[1] * (s6 - 2.0 * (s6 * s6)) / r;
    }
    case ORC_POT_HARMONIC: /* functions.rs:141-143 */
        return p[0] * (p[1] - r);
    case ORC_POT_BUCKINGHAM: { /* functions.rs:294-299 */
        double r3 = r * r * r;
        double r7 = r3 * r3 * r;
        double e = exp(-r / p[2]);
        return p[0] / p[2] * e - 6.0 * p[1] / r7;
    }
    case ORC_POT_BMH: { /* functions.rs:361-366 */
        double r2 = r * r;
        double r7 = r2 * r2 * r2 * r;
        double e = exp((p[3] - r) / p[4]);
        return p[0] / p[4] * e - 6.0 * p[1] / r7 + 8.0 * p[2] / (r7 * r2);
    }
    case ORC_POT_MORSE: { /* functions.rs:420-423; as written, not -dE/dr */
        double e = exp((p[1] - r) * p[0]);
        return 2.0 * p[2] * (1.0 - e * e) * p[0];
    }
    case ORC_POT_GAUSSIAN: /* functions.rs:479-481 */
        return 2.0 * p[1] * r * orc_potential_energy(pot, r);
    case ORC_POT_MIE: { /* functions.rs:560-565 */
        double sigma_r = p[0] / r;
        double repulsive = pow(sigma_r, p[1]);
        double attractive = pow(sigma_r, p[2]);
        return p[3] * (p[1] * repulsive - p[2] * attractive) / r;
    }
    case ORC_POT_COSINE_HARMONIC: /* functions.rs:204-206 */
        return p[0] * (cos(r) - p[1]) * sin(r);
    case ORC_POT_TORSION: { /* functions.rs:251-255 */
        double sn = sin(p[2] * r - p[1]);
        return p[0] * p[2] * sn;
    }
    default:
        return 0.0;
    }
}

double orc_potential_tail_energy(const orc_potential* pot, double rc) {
    const double* p = pot->p;
    switch (pot->pot) {
    case ORC_POT_LJ: { /* functions.rs:92-98 */
        double s3 = p[0] * p[0] * p[0];
        double rc3 = rc * rc * rc;
        double s9 = s3 * s3 * s3;
        double rc9 = rc3 * rc3 * rc3;
        return 4.0 / 3.0 * p[1] * s3 * (1.0 / 3.0 * s9 / rc9 - s3 / rc3);
    }
    case ORC_POT_BUCKINGHAM: { /* functions.rs:303-309 */
        double rc2 = rc * rc;
        double rc3 = rc2 * rc;
        double e = exp(-rc / p[2]);
        double factor = rc2 - 2.0 * rc * p[2] + 2.0 * p[2] * p[2];
        return p[0] * p[2] * e * factor - p[1] / (3.0 * rc3);
    }
    case ORC_POT_BMH: { /* functions.rs:370-376 */
        double rc2 = rc * rc;
        double rc3 = rc2 * rc;
        double e = exp((p[3] - rc) / p[4]);
        double factor = rc2 - 2.0 * rc * p[4] + 2.0 * p[4] * p[4];
        return p[0] * p[4] * e * factor - p[1] / (3.0 * rc3) + p[2] / (5.0 * rc2 * rc3);
    }
    case ORC_POT_GAUSSIAN: /* functions.rs:485-488 */
        return orc_potential_energy(pot, rc) * rc / (2.0 * p[1]) -
               p[0] * sqrt(PI) * erfc(sqrt(p[1]) * rc) / (4.0 * pow(p[1], 3.0 / 2.0));
    case ORC_POT_MIE: { /* functions.rs:569-579 */
        if (p[2] <= 3.0) {
            return 0.0;
        }
        double sigma_rc = p[0] / rc;
        double n_3 = p[1] - 3.0;
        double m_3 = p[2] - 3.0;
        double repulsive = pow(sigma_rc, n_3);
        double attractive = pow(sigma_rc, m_3);
        /* sigma.powi(3): square-and-multiply gives sigma * sigma^2 */
        double s3 = p[0] * (p[0] * p[0]);
        return p[3] * s3 * (repulsive / n_3 - attractive / m_3);
    }
    default: /* Null, Harmonic (functions.rs:150-155), Morse (functions.rs:427-432) */
        return 0.0;
    }
}

double orc_potential_tail_virial(const orc_potential* pot, double rc) {
    const double* p = pot->p;
    switch (pot->pot) {
    case ORC_POT_LJ: { /* functions.rs:100-106 */
        double s3 = p[0] * p[0] * p[0];
        double rc3 = rc * rc * rc;
        double s9 = s3 * s3 * s3;
        double rc9 = rc3 * rc3 * rc3;
        return 8.0 * p[1] * s3 * (2.0 / 3.0 * s9 / rc9 - s3 / rc3);
    }
    case ORC_POT_BUCKINGHAM: { /* functions.rs:311-318, including the stray + 8.0 pinned by functions.rs:710 */
        double rc2 = rc * rc;
        double rc3 = rc2 * rc;
        double e = exp(-rc / p[2]);
        double factor = rc3 + 3.0 * rc2 * p[2] + 6.0 * rc * p[2] * p[2] + 6.0 * p[2] * p[2] * p[2];
        return p[0] * e * factor - 20.0 * p[1] / rc3 + 8.0;
    }
    case ORC_POT_BMH: { /* functions.rs:378-385 */
        double rc2 = rc * rc;
        double rc3 = rc2 * rc;
        double e = exp((p[3] - rc) / p[4]);
        double factor = rc3 + 3.0 * rc2 * p[4] + 6.0 * rc * p[4] * p[4] + 6.0 * p[4] * p[4] * p[4];
        return p[0] * e * factor - 20.0 * p[1] / rc3 + 8.0 * p[2] / (5.0 * rc2 * rc3);
    }
    case ORC_POT_GAUSSIAN: /* functions.rs:490-493 */
        return 3.0 * sqrt(PI) * p[0] * erfc(sqrt(p[1]) * rc) / (4.0 * pow(p[1], 3.0 / 2.0)) -
               orc_potential_energy(pot, rc) * rc * (2.0 * p[1] * rc * rc + 3.0) / (2.0 * p[1]);
    case ORC_POT_MIE: { /* functions.rs:581-591 */
        if (p[2] <= 3.0) {
            return 0.0;
        }
        double sigma_rc = p[0] / rc;
        double n_3 = p[1] - 3.0;
        double m_3 = p[2] - 3.0;
        double repulsive = pow(sigma_rc, n_3);
        double attractive = pow(sigma_rc, m_3);
        double s3 = p[0] * (p[0] * p[0]);
        return p[3] * s3 * (repulsive * p[1] / n_3 - attractive * p[2] / m_3);
    }
    default:
        return 0.0;
    }
}

/* energy/mod.rs:137-142 default PairPotential::virial given force(r) */
static void virial_from_force(double fact, const double r[3], double w[9]) {
    double norm = norm3(r);
    double rn[3] = {r[0] / norm, r[1] / norm, r[2] / norm};
    double force[3] = {fact * rn[0], fact * rn[1], fact * rn[2]};
    /* vectors.rs:150-156 tensorial: w[i][j] = force[i] * r[j] */
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) {
            w[3 * a + b] = force[a] * r[b];
        }
    }
}

void orc_potential_virial(const orc_potential* pot, const double r[3], double w[9]) {
    virial_from_force(orc_potential_force(pot, norm3(r)), r, w);
}

/* ------------------------------------------------------------------------ */
/* TableComputation (energy/computations.rs)                                 */
/* ------------------------------------------------------------------------ */

/* computations.rs:102-119 */
void orc_table_build(const orc_potential* pot, int32_t size, double max, double* energy, double* force) {
    double delta = max / (double)size;
    for (int32_t i = 0; i < size; i++) {
        double r = (double)i * delta;
        energy[i] = orc_potential_energy(pot, r);
        force[i] = orc_potential_force(pot, r);
    }
}

/* computations.rs:123-145 (the same code serves the energy and the force tables) */
double orc_table_energy(const double* table, int32_t size, double max, double r) {
    double delta = max / (double)size;
    double q = floor(r / delta);
    /* `as usize` saturates: negative and NaN -> 0, huge -> usize::MAX */
    uint64_t bin;
    if (!(q > 0.0)) {
        bin = 0;
    } else if (q >= 18446744073709551615.0) {
        bin = UINT64_MAX;
    } else {
        bin = (uint64_t)q;
    }
    if (bin < (uint64_t)(size - 1)) {
        double dx = r - (double)bin * delta;
        double slope = (table[bin + 1] - table[bin]) / delta;
        return table[bin] + dx * slope;
    }
    return 0.0;
}

/* ------------------------------------------------------------------------ */
/* PairInteraction (energy/pairs.rs)                                         */
/* ------------------------------------------------------------------------ */

static inline double inner_energy(const orc_pair* pair, double r) {
    if (pair->table_n > 0) {
        return orc_table_energy(pair->table_energy, pair->table_n, pair->table_max, r);
    }
    return orc_potential_energy(&pair->potential, r);
}

static inline double inner_force(const orc_pair* pair, double r) {
    if (pair->table_n > 0) {
        return orc_table_energy(pair->table_force, pair->table_n, pair->table_max, r);
    }
    return orc_potential_force(&pair->potential, r);
}

/* pairs.rs:185-195 */
double orc_pair_energy(const orc_pair* pair, double r) {
    if (r >= pair->cutoff) {
        return 0.0;
    }
    double energy = inner_energy(pair, r);
    if (pair->shifted) {
        /* pairs.rs:86-95: shift = potential.energy(cutoff) */
        return energy - inner_energy(pair, pair->cutoff);
    }
    return energy;
}

/* pairs.rs:212-218 */
double orc_pair_force(const orc_pair* pair, double r) {
    if (r >= pair->cutoff) {
        return 0.0;
    }
    return inner_force(pair, r);
}

/* pairs.rs:237-243 */
void orc_pair_virial(const orc_pair* pair, const double r[3], double w[9]) {
    if (norm3(r) >= pair->cutoff) {
        memset(w, 0, 9 * sizeof(double));
        return;
    }
    virial_from_force(inner_force(pair, norm3(r)), r, w);
}

/* pairs.rs:259-265; TableComputation delegates to the wrapped potential, computations.rs:148-174 */
double orc_pair_tail_energy(const orc_pair* pair) {
    if (pair->tail) {
        return orc_potential_tail_energy(&pair->potential, pair->cutoff);
    }
    return 0.0;
}

/* pairs.rs:289-296 */
double orc_pair_tail_virial(const orc_pair* pair) {
    if (pair->tail) {
        return orc_potential_tail_virial(&pair->potential, pair->cutoff);
    }
    return 0.0;
}

/* ------------------------------------------------------------------------ */
/* restrictions and topology                                                 */
/* ------------------------------------------------------------------------ */

/* restrictions.rs:85-114 */
void orc_restriction_information(int32_t restriction, double scale14, int32_t path, int32_t* excluded, double* scaling) {
    int same = path != ORC_PATH_NONE;
    int ex = 0;
    switch (restriction) {
    case ORC_RESTRICT_NONE:
        ex = 0;
        break;
    case ORC_RESTRICT_INTER:
        ex = same;
        break;
    case ORC_RESTRICT_INTRA:
        ex = !same;
        break;
    case ORC_RESTRICT_EXCLUDE12:
        ex = path == ORC_PATH_ONE;
        break;
    case ORC_RESTRICT_EXCLUDE13:
    case ORC_RESTRICT_SCALE14:
        ex = path == ORC_PATH_ONE || path == ORC_PATH_TWO;
        break;
    case ORC_RESTRICT_EXCLUDE14:
        ex = path == ORC_PATH_ONE || path == ORC_PATH_TWO || path == ORC_PATH_THREE;
        break;
    }
    *excluded = ex;
    *scaling = (restriction == ORC_RESTRICT_SCALE14 && path == ORC_PATH_THREE) ? scale14 : 1.0;
}

/* configuration.rs:124-144 */
int32_t orc_bond_path(const orc_system* s, int64_t i, int64_t j) {
    if (s->molid[i] != s->molid[j]) {
        return ORC_PATH_NONE;
    }
    if (i == j) {
        return ORC_PATH_SAME;
    }
    int64_t mol = s->molid[i];
    int64_t first = s->mol_start[mol];
    int64_t size = s->mol_start[mol + 1] - first;
    uint8_t connect = s->bond_dist[s->bond_dist_off[mol] + (i - first) * size + (j - first)];
    if (connect & ORC_BOND_ONE) {
        return ORC_PATH_ONE;
    } else if (connect & ORC_BOND_TWO) {
        return ORC_PATH_TWO;
    } else if (connect & ORC_BOND_THREE) {
        return ORC_PATH_THREE;
    }
    return ORC_PATH_FAR;
}

static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
static int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

/* sys/config/connect.rs:50-63 Angle::new; :88-103 Dihedral::new */
static void angle_new(int64_t a, int64_t b, int64_t c, int64_t out[3]) {
    out[0] = imin(a, c);
    out[1] = b;
    out[2] = imax(a, c);
}

static void dihedral_new(int64_t a, int64_t b, int64_t c, int64_t d, int64_t out[4]) {
    if (imax(a, b) < imax(c, d)) {
        out[0] = a;
        out[1] = b;
        out[2] = c;
        out[3] = d;
    } else {
        out[0] = d;
        out[1] = c;
        out[2] = b;
        out[3] = a;
    }
}

static int find_tuple(const int64_t* set, int64_t count, const int64_t* item, int width) {
    for (int64_t i = 0; i < count; i++) {
        if (memcmp(set + i * width, item, (size_t)width * sizeof(int64_t)) == 0) {
            return 1;
        }
    }
    return 0;
}

/* bonding.rs:78-155 Bonding::rebuild + rebuild_connections.  The reference keeps HashSets, so the
 * enumeration order of the output is unspecified there; here it is the loop order. */
void orc_bonding_rebuild(int64_t natoms, int64_t nbonds, const int64_t* bonds_in, int64_t* nangles_out, int64_t* angles,
                         int64_t* ndihedrals_out, int64_t* dihedrals, uint8_t* distances) {
    /* Bond::new sorts (i, j) (connect.rs:17-22) */
    int64_t* bonds = (int64_t*)malloc((size_t)(nbonds > 0 ? nbonds : 1) * 2 * sizeof(int64_t));
    for (int64_t b = 0; b < nbonds; b++) {
        bonds[2 * b] = imin(bonds_in[2 * b], bonds_in[2 * b + 1]);
        bonds[2 * b + 1] = imax(bonds_in[2 * b], bonds_in[2 * b + 1]);
    }
    int64_t cap_a = 16, cap_d = 16;
    int64_t na = 0, nd = 0;
    int64_t* a_set = (int64_t*)malloc((size_t)cap_a * 3 * sizeof(int64_t));
    int64_t* d_set = (int64_t*)malloc((size_t)cap_d * 4 * sizeof(int64_t));

    for (int64_t b1 = 0; b1 < nbonds; b1++) {
        int64_t b1i = bonds[2 * b1], b1j = bonds[2 * b1 + 1];
        for (int64_t b2 = 0; b2 < nbonds; b2++) {
            int64_t b2i = bonds[2 * b2], b2j = bonds[2 * b2 + 1];
            if (b1i == b2i && b1j == b2j) {
                continue;
            }
            int64_t angle[3];
            if (b1i == b2j) {
                angle_new(b2i, b2j, b1j, angle);
            } else if (b1j == b2i) {
                angle_new(b1i, b1j, b2j, angle);
            } else if (b1j == b2j) {
                angle_new(b1i, b1j, b2i, angle);
            } else if (b1i == b2i) {
                angle_new(b1j, b1i, b2j, angle);
            } else {
                continue;
            }
            if (!find_tuple(a_set, na, angle, 3)) {
                if (na == cap_a) {
                    cap_a *= 2;
                    a_set = (int64_t*)realloc(a_set, (size_t)cap_a * 3 * sizeof(int64_t));
                }
                memcpy(a_set + 3 * na, angle, 3 * sizeof(int64_t));
                na++;
            }
            for (int64_t b3 = 0; b3 < nbonds; b3++) {
                int64_t b3i = bonds[2 * b3], b3j = bonds[2 * b3 + 1];
                if (b2i == b3i && b2j == b3j) {
                    continue;
                }
                int64_t dihedral[4];
                if (angle[2] == b3i && angle[1] != b3j) {
                    dihedral_new(angle[0], angle[1], angle[2], b3j, dihedral);
                } else if (angle[2] == b3j && angle[1] != b3i) {
                    dihedral_new(angle[0], angle[1], angle[2], b3i, dihedral);
                } else if (angle[0] == b3j && angle[1] != b3i) {
                    dihedral_new(b3i, angle[0], angle[1], angle[2], dihedral);
                } else if (angle[0] == b3i && angle[1] != b3j) {
                    dihedral_new(b3j, angle[0], angle[1], angle[2], dihedral);
                } else {
                    continue;
                }
                if (!find_tuple(d_set, nd, dihedral, 4)) {
                    if (nd == cap_d) {
                        cap_d *= 2;
                        d_set = (int64_t*)realloc(d_set, (size_t)cap_d * 4 * sizeof(int64_t));
                    }
                    memcpy(d_set + 4 * nd, dihedral, 4 * sizeof(int64_t));
                    nd++;
                }
            }
        }
    }

    if (distances != NULL) {
        /* Array2::default -> BondDistances::FAR everywhere (connect.rs:161-165) */
        memset(distances, ORC_BOND_FAR, (size_t)(natoms * natoms));
        for (int64_t b = 0; b < nbonds; b++) {
            distances[bonds[2 * b] * natoms + bonds[2 * b + 1]] |= ORC_BOND_ONE;
            distances[bonds[2 * b + 1] * natoms + bonds[2 * b]] |= ORC_BOND_ONE;
        }
        for (int64_t a = 0; a < na; a++) {
            distances[a_set[3 * a] * natoms + a_set[3 * a + 2]] |= ORC_BOND_TWO;
            distances[a_set[3 * a + 2] * natoms + a_set[3 * a]] |= ORC_BOND_TWO;
        }
        for (int64_t d = 0; d < nd; d++) {
            distances[d_set[4 * d] * natoms + d_set[4 * d + 3]] |= ORC_BOND_THREE;
            distances[d_set[4 * d + 3] * natoms + d_set[4 * d]] |= ORC_BOND_THREE;
        }
    }
    if (angles != NULL) {
        memcpy(angles, a_set, (size_t)na * 3 * sizeof(int64_t));
    }
    if (dihedrals != NULL) {
        memcpy(dihedrals, d_set, (size_t)nd * 4 * sizeof(int64_t));
    }
    *nangles_out = na;
    *ndihedrals_out = nd;
    free(bonds);
    free(a_set);
    free(d_set);
}

/* ------------------------------------------------------------------------ */
/* UnitCell (sys/config/cells.rs)                                            */
/* ------------------------------------------------------------------------ */

typedef struct {
    double cell[9];
    double inv[9];
    int32_t shape;
} geom_t;

static void geom_init(geom_t* g, const double cell[9], int32_t shape) {
    memcpy(g->cell, cell, 9 * sizeof(double));
    g->shape = shape;
    if (shape == ORC_CELL_INFINITE) {
        memset(g->inv, 0, 9 * sizeof(double));
    } else {
        orc_matrix_inverse(cell, g->inv);
    }
}

/* cells.rs:284-300 */
static inline void geom_image(const geom_t* g, double v[3]) {
    if (g->shape == ORC_CELL_ORTHO) {
        /* a(), b(), c() are cell[0][0], cell[1][1], cell[2][2] for orthorhombic cells (cells.rs:94-116) */
        v[0] -= round(v[0] / g->cell[0]) * g->cell[0];
        v[1] -= round(v[1] / g->cell[4]) * g->cell[4];
        v[2] -= round(v[2] / g->cell[8]) * g->cell[8];
    } else if (g->shape == ORC_CELL_TRICLINIC) {
        double f[3];
        matvec(g->inv, v, f);
        f[0] -= round(f[0]);
        f[1] -= round(f[1]);
        f[2] -= round(f[2]);
        matvec(g->cell, f, v);
    }
}

void orc_vector_image(const double cell[9], int32_t shape, double v[3]) {
    geom_t g;
    geom_init(&g, cell, shape);
    geom_image(&g, v);
}

/* cells.rs:263-279 */
void orc_wrap_vector(const double cell[9], int32_t shape, double v[3]) {
    geom_t g;
    geom_init(&g, cell, shape);
    if (shape == ORC_CELL_ORTHO) {
        v[0] -= floor(v[0] / g.cell[0]) * g.cell[0];
        v[1] -= floor(v[1] / g.cell[4]) * g.cell[4];
        v[2] -= floor(v[2] / g.cell[8]) * g.cell[8];
    } else if (shape == ORC_CELL_TRICLINIC) {
        double f[3];
        matvec(g.inv, v, f);
        f[0] -= floor(f[0]);
        f[1] -= floor(f[1]);
        f[2] -= floor(f[2]);
        matvec(g.cell, f, v);
    }
}

static void cell_vectors(const double cell[9], double a[3], double b[3], double c[3]) {
    /* cells.rs:234-255: lattice vectors are the matrix columns */
    a[0] = cell[0];
    a[1] = cell[3];
    a[2] = cell[6];
    b[0] = cell[1];
    b[1] = cell[4];
    b[2] = cell[7];
    c[0] = cell[2];
    c[1] = cell[5];
    c[2] = cell[8];
}

/* cells.rs:185-199 */
double orc_cell_volume(const double cell[9], int32_t shape) {
    if (shape == ORC_CELL_INFINITE) {
        return 0.0;
    } else if (shape == ORC_CELL_ORTHO) {
        return cell[0] * cell[4] * cell[8];
    }
    double a[3], b[3], c[3], bc[3];
    cell_vectors(cell, a, b, c);
    cross3(b, c, bc);
    return dot3(a, bc);
}

/* cells.rs:134-146 */
void orc_cell_lengths(const double cell[9], int32_t shape, double lengths[3]) {
    if (shape == ORC_CELL_INFINITE) {
        lengths[0] = lengths[1] = lengths[2] = INFINITY;
        return;
    }
    double a[3], b[3], c[3], na[3], nb[3], nc[3];
    cell_vectors(cell, a, b, c);
    cross3(b, c, na);
    cross3(c, a, nb);
    cross3(a, b, nc);
    double n;
    n = norm3(na);
    for (int k = 0; k < 3; k++) na[k] /= n;
    n = norm3(nb);
    for (int k = 0; k < 3; k++) nb[k] /= n;
    n = norm3(nc);
    for (int k = 0; k < 3; k++) nc[k] /= n;
    lengths[0] = fabs(dot3(na, a));
    lengths[1] = fabs(dot3(nb, b));
    lengths[2] = fabs(dot3(nc, c));
}

/* cells.rs:224-226: 2.0 * PI * self.inv * Vector3D::from(index); (2.0 * PI) is an f64, then f64 * Matrix3, then * vector */
static void geom_k_vector(const geom_t* g, const double index[3], double k[3]) {
    double two_pi = 2.0 * PI;
    double m[9];
    for (int a = 0; a < 9; a++) {
        m[a] = two_pi * g->inv[a];
    }
    matvec(m, index, k);
}

void orc_k_vector(const double cell[9], const double index[3], double k[3]) {
    geom_t g;
    geom_init(&g, cell, ORC_CELL_TRICLINIC);
    geom_k_vector(&g, index, k);
}

/* cells.rs:335-359 */
static double geom_angle_and_derivatives(const geom_t* g, const double* r1, const double* r2, const double* r3,
                                         double d1[3], double d2[3], double d3[3]) {
    double r12[3] = {r1[0] - r2[0], r1[1] - r2[1], r1[2] - r2[2]};
    geom_image(g, r12);
    double r23[3] = {r3[0] - r2[0], r3[1] - r2[1], r3[2] - r2[2]};
    geom_image(g, r23);

    double r12_norm = norm3(r12);
    double r23_norm = norm3(r23);
    double r12n[3] = {r12[0] / r12_norm, r12[1] / r12_norm, r12[2] / r12_norm};
    double r23n[3] = {r23[0] / r23_norm, r23[1] / r23_norm, r23[2] / r23_norm};

    double c = dot3(r12n, r23n);
    double sin_inv = 1.0 / sqrt(1.0 - c * c);

    for (int k = 0; k < 3; k++) {
        d1[k] = sin_inv * (c * r12n[k] - r23n[k]) / r12_norm;
        d3[k] = sin_inv * (c * r23n[k] - r12n[k]) / r23_norm;
        d2[k] = -(d1[k] + d3[k]);
    }
    return acos(c);
}

double orc_angle_and_derivatives(const double cell[9], int32_t shape, const double* r1, const double* r2,
                                 const double* r3, double d1[3], double d2[3], double d3[3]) {
    geom_t g;
    geom_init(&g, cell, shape);
    return geom_angle_and_derivatives(&g, r1, r2, r3, d1, d2, d3);
}

/* cells.rs:322-333 */
static double geom_angle(const geom_t* g, const double* r1, const double* r2, const double* r3) {
    double r12[3] = {r1[0] - r2[0], r1[1] - r2[1], r1[2] - r2[2]};
    geom_image(g, r12);
    double r23[3] = {r3[0] - r2[0], r3[1] - r2[1], r3[2] - r2[2]};
    geom_image(g, r23);
    return acos(dot3(r12, r23) / (norm3(r12) * norm3(r23)));
}

/* cells.rs:379-411 */
static double geom_dihedral_and_derivatives(const geom_t* g, const double* r1, const double* r2, const double* r3,
                                            const double* r4, double d1[3], double d2[3], double d3[3], double d4[3]) {
    double r12[3] = {r2[0] - r1[0], r2[1] - r1[1], r2[2] - r1[2]};
    geom_image(g, r12);
    double r23[3] = {r3[0] - r2[0], r3[1] - r2[1], r3[2] - r2[2]};
    geom_image(g, r23);
    double r34[3] = {r4[0] - r3[0], r4[1] - r3[1], r4[2] - r3[2]};
    geom_image(g, r34);

    double u[3], v[3];
    cross3(r12, r23, u);
    cross3(r23, r34, v);
    double u_norm2 = norm2_3(u);
    double v_norm2 = norm2_3(v);
    double r23_norm2 = norm2_3(r23);
    double r23_norm = sqrt(r23_norm2);

    double f1 = -r23_norm / u_norm2;
    double f4 = r23_norm / v_norm2;
    for (int k = 0; k < 3; k++) {
        d1[k] = f1 * u[k];
        d4[k] = f4 * v[k];
    }
    double r23_r34 = dot3(r23, r34);
    double r12_r23 = dot3(r12, r23);

    double c21 = -r12_r23 / r23_norm2 - 1.0;
    double c24 = r23_r34 / r23_norm2;
    double c34 = -r23_r34 / r23_norm2 - 1.0;
    double c31 = r12_r23 / r23_norm2;
    for (int k = 0; k < 3; k++) {
        d2[k] = c21 * d1[k] + c24 * d4[k];
        d3[k] = c34 * d4[k] + c31 * d1[k];
    }
    /* r23_norm * v * r12: (f64 * Vector3D) dot Vector3D */
    double sv[3] = {r23_norm * v[0], r23_norm * v[1], r23_norm * v[2]};
    return atan2(dot3(sv, r12), dot3(u, v));
}

double orc_dihedral_and_derivatives(const double cell[9], int32_t shape, const double* r1, const double* r2,
                                    const double* r3, const double* r4, double d1[3], double d2[3], double d3[3],
                                    double d4[3]) {
    geom_t g;
    geom_init(&g, cell, shape);
    return geom_dihedral_and_derivatives(&g, r1, r2, r3, r4, d1, d2, d3, d4);
}

/* cells.rs:361-377 */
static double geom_dihedral(const geom_t* g, const double* r1, const double* r2, const double* r3, const double* r4) {
    double r12[3] = {r2[0] - r1[0], r2[1] - r1[1], r2[2] - r1[2]};
    geom_image(g, r12);
    double r23[3] = {r3[0] - r2[0], r3[1] - r2[1], r3[2] - r2[2]};
    geom_image(g, r23);
    double r34[3] = {r4[0] - r3[0], r4[1] - r3[1], r4[2] - r3[2]};
    geom_image(g, r34);
    double u[3], v[3];
    cross3(r12, r23, u);
    cross3(r23, r34, v);
    double r23_norm = norm3(r23);
    double sv[3] = {r23_norm * v[0], r23_norm * v[1], r23_norm * v[2]};
    return atan2(dot3(sv, r12), dot3(u, v));
}

/* configuration.rs:399-403 nearest_image(i, j) = image(r_i - r_j) */
static inline void nearest_image(const orc_system* s, const geom_t* g, int64_t i, int64_t j, double d[3]) {
    d[0] = s->position[3 * i] - s->position[3 * j];
    d[1] = s->position[3 * i + 1] - s->position[3 * j + 1];
    d[2] = s->position[3 * i + 2] - s->position[3 * j + 2];
    geom_image(g, d);
}

/* configuration.rs:393-397 distance(i, j) -> cells.rs:316-320: image(r_j - r_i).norm() */
static inline double distance(const orc_system* s, const geom_t* g, int64_t i, int64_t j) {
    double d[3] = {s->position[3 * j] - s->position[3 * i], s->position[3 * j + 1] - s->position[3 * i + 1],
                   s->position[3 * j + 2] - s->position[3 * i + 2]};
    geom_image(g, d);
    return norm3(d);
}

/* system.rs:178-182 + interactions.rs:142-145 */
static inline const orc_pair* pair_potential(const orc_system* s, int64_t i, int64_t j) {
    const orc_pair* pair = &s->pairs[(int64_t)s->kind[i] * s->nkinds + s->kind[j]];
    return pair->potential.pot == ORC_POT_ABSENT ? NULL : pair;
}

/* ------------------------------------------------------------------------ */
/* per-thread force buffers (utils/thread_vec.rs:13-56)                      */
/* ------------------------------------------------------------------------ */

typedef struct {
    int nthreads;
    int64_t n3;
    double* data;
} thread_vec_t;

static void thread_vec_init(thread_vec_t* tv, int64_t n3) {
    tv->nthreads = max_threads();
    tv->n3 = n3;
    tv->data = (double*)calloc((size_t)tv->nthreads * (size_t)(n3 > 0 ? n3 : 1), sizeof(double));
}

/* thread_vec.rs:48-55 sum_into: buffers added one after the other */
static void thread_vec_sum_into(thread_vec_t* tv, double* out) {
    for (int t = 0; t < tv->nthreads; t++) {
        const double* local = tv->data + (size_t)t * (size_t)tv->n3;
        for (int64_t k = 0; k < tv->n3; k++) {
            out[k] += local[k];
        }
    }
    free(tv->data);
}

/* ------------------------------------------------------------------------ */
/* Forces (sys/compute.rs:33-107)                                            */
/* ------------------------------------------------------------------------ */

/* compute.rs:37-60 */
void orc_pair_forces(const orc_system* s, double* out) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    thread_vec_t tv;
    thread_vec_init(&tv, 3 * n);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
        double* forces = tv.data + (size_t)thread_id() * (size_t)(3 * n);
        double force_i[3] = {0.0, 0.0, 0.0};
        for (int64_t j = i + 1; j < n; j++) {
            int32_t path = orc_bond_path(s, i, j);
            double d[3];
            nearest_image(s, &g, i, j, d);
            double r = norm3(d);
            double dn[3] = {d[0] / r, d[1] / r, d[2] / r};
            const orc_pair* potential = pair_potential(s, i, j);
            if (potential != NULL) {
                int32_t excluded;
                double scaling;
                orc_restriction_information(potential->restriction, potential->scale14, path, &excluded, &scaling);
                if (!excluded) {
                    double f = scaling * orc_pair_force(potential, r);
                    for (int k = 0; k < 3; k++) {
                        double force = f * dn[k];
                        force_i[k] += force;
                        forces[3 * j + k] -= force;
                    }
                }
            }
        }
        for (int k = 0; k < 3; k++) {
            forces[3 * i + k] += force_i[k];
        }
    }
    memset(out, 0, (size_t)(3 * n) * sizeof(double));
    thread_vec_sum_into(&tv, out);
}

/* Bounded sample of compute.rs:37-55 for the CPU baseline: the inner loop `for j in (i + 1)..natoms` of the
 * rows listed in `rows`, spread over the OpenMP threads exactly like the full loop, forces accumulated in
 * per-thread buffers.  Rows drawn uniformly from [0, n) cost on average what a row of the full O(N^2) loop
 * costs, so (time * n / nrows) estimates one full evaluation.  Returns the number of pairs inside the cut-off
 * among the visited ones; `checksum` receives the sum of |force_i| over the sampled rows. */
int64_t orc_pair_forces_sample(const orc_system* s, int64_t nrows, const int64_t* rows, double* checksum) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    /* The per-thread buffers (ThreadLocalVec, threads x 3N doubles) are set up once per full evaluation in the
     * reference; a sample of a few rows must not pay for them at every call (400 MB of page faults against a few
     * milliseconds of pair visits), so they are kept between calls.  Their content is never read back here. */
    static thread_vec_t tv = {0, 0, NULL};
    if (tv.data == NULL || tv.nthreads != max_threads() || tv.n3 != 3 * n) {
        free(tv.data);
        thread_vec_init(&tv, 3 * n);
    }
    int64_t inside = 0;
    double sum = 0.0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : inside, sum)
    for (int64_t row = 0; row < nrows; row++) {
        int64_t i = rows[row];
        double* forces = tv.data + (size_t)thread_id() * (size_t)(3 * n);
        double force_i[3] = {0.0, 0.0, 0.0};
        for (int64_t j = i + 1; j < n; j++) {
            int32_t path = orc_bond_path(s, i, j);
            double d[3];
            nearest_image(s, &g, i, j, d);
            double r = norm3(d);
            double dn[3] = {d[0] / r, d[1] / r, d[2] / r};
            const orc_pair* potential = pair_potential(s, i, j);
            if (potential != NULL) {
                int32_t excluded;
                double scaling;
                orc_restriction_information(potential->restriction, potential->scale14, path, &excluded, &scaling);
                if (!excluded) {
                    double f = scaling * orc_pair_force(potential, r);
                    if (r < potential->cutoff) inside++;
                    for (int k = 0; k < 3; k++) {
                        double force = f * dn[k];
                        force_i[k] += force;
                        forces[3 * j + k] -= force;
                    }
                }
            }
        }
        for (int k = 0; k < 3; k++) {
            forces[3 * i + k] += force_i[k];
        }
        sum += fabs(force_i[0]) + fabs(force_i[1]) + fabs(force_i[2]);
    }
    if (checksum) *checksum = sum;
    return inside;
}

/* Total pair force on the atoms listed in `rows`, for boxes too large for the full O(N^2) loop: every pair (i, j),
 * j != i, is evaluated exactly as compute.rs:40-53 evaluates it when its turn comes in the i < j loop: the atom with
 * the smaller index is the first argument of nearest_image and receives +f, the other one -f.  out: nrows x 3. */
void orc_pair_forces_rows(const orc_system* s, int64_t nrows, const int64_t* rows, double* out) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t row = 0; row < nrows; row++) {
        int64_t t = rows[row];
        double total[3] = {0.0, 0.0, 0.0};
        for (int64_t o = 0; o < n; o++) {
            if (o == t) continue;
            int64_t i = t < o ? t : o, j = t < o ? o : t;
            int32_t path = orc_bond_path(s, i, j);
            double d[3];
            nearest_image(s, &g, i, j, d);
            double r = norm3(d);
            double dn[3] = {d[0] / r, d[1] / r, d[2] / r};
            const orc_pair* potential = pair_potential(s, i, j);
            if (potential == NULL) continue;
            int32_t excluded;
            double scaling;
            orc_restriction_information(potential->restriction, potential->scale14, path, &excluded, &scaling);
            if (excluded) continue;
            double f = scaling * orc_pair_force(potential, r);
            for (int k = 0; k < 3; k++) {
                double force = f * dn[k];
                if (t == i) {
                    total[k] += force;
                } else {
                    total[k] -= force;
                }
            }
        }
        for (int k = 0; k < 3; k++) out[3 * row + k] = total[k];
    }
}

/* compute.rs:62-97 */
void orc_bonded_forces(const orc_system* s, double* forces) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    for (int64_t b = 0; b < s->nbonds; b++) {
        int64_t i = s->bonds[2 * b], j = s->bonds[2 * b + 1];
        double d[3];
        nearest_image(s, &g, i, j, d);
        double r = norm3(d);
        double dn[3] = {d[0] / r, d[1] / r, d[2] / r};
        if (s->bond_pot[b].pot != ORC_POT_ABSENT) {
            double f = orc_potential_force(&s->bond_pot[b], r);
            for (int k = 0; k < 3; k++) {
                double force = f * dn[k];
                forces[3 * i + k] += force;
                forces[3 * j + k] -= force;
            }
        }
    }
    for (int64_t a = 0; a < s->nangles; a++) {
        int64_t i = s->angles[3 * a], j = s->angles[3 * a + 1], k = s->angles[3 * a + 2];
        double d1[3], d2[3], d3[3];
        double theta = geom_angle_and_derivatives(&g, s->position + 3 * i, s->position + 3 * j, s->position + 3 * k, d1,
                                                  d2, d3);
        if (s->angle_pot[a].pot != ORC_POT_ABSENT) {
            double f = orc_potential_force(&s->angle_pot[a], theta);
            for (int c = 0; c < 3; c++) {
                forces[3 * i + c] += f * d1[c];
                forces[3 * j + c] += f * d2[c];
                forces[3 * k + c] += f * d3[c];
            }
        }
    }
    for (int64_t q = 0; q < s->ndihedrals; q++) {
        int64_t i = s->dihedrals[4 * q], j = s->dihedrals[4 * q + 1], k = s->dihedrals[4 * q + 2],
                m = s->dihedrals[4 * q + 3];
        double d1[3], d2[3], d3[3], d4[3];
        double phi = geom_dihedral_and_derivatives(&g, s->position + 3 * i, s->position + 3 * j, s->position + 3 * k,
                                                   s->position + 3 * m, d1, d2, d3, d4);
        if (s->dihedral_pot[q].pot != ORC_POT_ABSENT) {
            double f = orc_potential_force(&s->dihedral_pot[q], phi);
            for (int c = 0; c < 3; c++) {
                forces[3 * i + c] += f * d1[c];
                forces[3 * j + c] += f * d2[c];
                forces[3 * k + c] += f * d3[c];
                forces[3 * m + c] += f * d4[c];
            }
        }
    }
}

void orc_coulomb_forces(const orc_system* s, double* forces) {
    if (s->coulomb == ORC_COULOMB_EWALD) {
        /* ewald.rs:897-905 */
        orc_ewald_real_forces(s, forces);
        orc_ewald_kspace_forces(s, forces);
    } else if (s->coulomb == ORC_COULOMB_WOLF) {
        orc_wolf_forces(s, forces);
    }
}

void orc_forces(const orc_system* s, double* forces) {
    orc_pair_forces(s, forces);
    orc_bonded_forces(s, forces);
    orc_coulomb_forces(s, forces);
}

/* ------------------------------------------------------------------------ */
/* EnergyEvaluator (sys/energy.rs)                                           */
/* ------------------------------------------------------------------------ */

/* energy.rs:47-59 with energy.rs:32-44 inlined */
double orc_pairs_energy(const orc_system* s) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    double total = 0.0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
    for (int64_t i = 0; i < n; i++) {
        double local_energy = 0.0;
        for (int64_t j = i + 1; j < n; j++) {
            double d[3];
            nearest_image(s, &g, i, j, d);
            double r = norm3(d);
            int32_t path = orc_bond_path(s, i, j);
            const orc_pair* potential = pair_potential(s, i, j);
            if (potential != NULL) {
                int32_t excluded;
                double scaling;
                orc_restriction_information(potential->restriction, potential->scale14, path, &excluded, &scaling);
                if (!excluded) {
                    local_energy += scaling * orc_pair_energy(potential, r);
                } else {
                    local_energy += 0.0;
                }
            } else {
                local_energy += 0.0;
            }
        }
        total += local_energy;
    }
    return total;
}

/* Composition::all_particles (composition.rs:123-129): kinds with a non-zero count, ascending */
static int64_t* composition(const orc_system* s) {
    int64_t* counts = (int64_t*)calloc((size_t)(s->nkinds > 0 ? s->nkinds : 1), sizeof(int64_t));
    for (int64_t i = 0; i < s->n; i++) {
        counts[s->kind[i]]++;
    }
    return counts;
}

/* energy.rs:63-79 */
double orc_pairs_tail_energy(const orc_system* s) {
    if (s->shape == ORC_CELL_INFINITE) {
        return 0.0;
    }
    double energy = 0.0;
    double volume = orc_cell_volume(s->cell, s->shape);
    int64_t* counts = composition(s);
    for (int32_t i = 0; i < s->nkinds; i++) {
        if (counts[i] == 0) continue;
        for (int32_t j = 0; j < s->nkinds; j++) {
            if (counts[j] == 0) continue;
            double two_pi_density = 2.0 * PI * (double)counts[i] * (double)counts[j] / volume;
            const orc_pair* potential = &s->pairs[(int64_t)i * s->nkinds + j];
            if (potential->potential.pot != ORC_POT_ABSENT) {
                energy += two_pi_density * orc_pair_tail_energy(potential);
            }
        }
    }
    free(counts);
    return energy;
}

/* energy.rs:90-141 */
static double bonds_energy(const orc_system* s, const geom_t* g) {
    double energy = 0.0;
    for (int64_t b = 0; b < s->nbonds; b++) {
        int64_t i = s->bonds[2 * b], j = s->bonds[2 * b + 1];
        double d[3];
        nearest_image(s, g, i, j, d);
        double r = norm3(d);
        energy += s->bond_pot[b].pot != ORC_POT_ABSENT ? orc_potential_energy(&s->bond_pot[b], r) : 0.0;
    }
    return energy;
}

static double angles_energy(const orc_system* s, const geom_t* g) {
    double energy = 0.0;
    for (int64_t a = 0; a < s->nangles; a++) {
        int64_t i = s->angles[3 * a], j = s->angles[3 * a + 1], k = s->angles[3 * a + 2];
        double theta = geom_angle(g, s->position + 3 * i, s->position + 3 * j, s->position + 3 * k);
        energy += s->angle_pot[a].pot != ORC_POT_ABSENT ? orc_potential_energy(&s->angle_pot[a], theta) : 0.0;
    }
    return energy;
}

static double dihedrals_energy(const orc_system* s, const geom_t* g) {
    double energy = 0.0;
    for (int64_t q = 0; q < s->ndihedrals; q++) {
        int64_t i = s->dihedrals[4 * q], j = s->dihedrals[4 * q + 1], k = s->dihedrals[4 * q + 2],
                m = s->dihedrals[4 * q + 3];
        double phi = geom_dihedral(g, s->position + 3 * i, s->position + 3 * j, s->position + 3 * k, s->position + 3 * m);
        energy += s->dihedral_pot[q].pot != ORC_POT_ABSENT ? orc_potential_energy(&s->dihedral_pot[q], phi) : 0.0;
    }
    return energy;
}

void orc_energy_terms_compute(const orc_system* s, orc_energy_terms* out) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    memset(out, 0, sizeof(*out));
    out->pairs = orc_pairs_energy(s);
    out->pairs_tail = orc_pairs_tail_energy(s);
    out->bonds = bonds_energy(s, &g);
    out->angles = angles_energy(s, &g);
    out->dihedrals = dihedrals_energy(s, &g);
    if (s->coulomb == ORC_COULOMB_EWALD) {
        out->coulomb_real = orc_ewald_real_energy(s);
        out->coulomb_self = orc_ewald_self_energy(s);
        out->coulomb_kspace = orc_ewald_kspace_energy(s);
    } else if (s->coulomb == ORC_COULOMB_WOLF) {
        out->coulomb_real = orc_wolf_energy(s);
    }
}

/* compute.rs:114-127 */
double orc_potential_energy_total(const orc_system* s) {
    orc_energy_terms t;
    orc_energy_terms_compute(s, &t);
    double energy = t.pairs;
    energy += t.pairs_tail;
    energy += t.bonds;
    energy += t.angles;
    energy += t.dihedrals;
    /* ewald.rs:888-895: real + self + k_space */
    energy += t.coulomb_real + t.coulomb_self + t.coulomb_kspace;
    return energy;
}

/* ------------------------------------------------------------------------ */
/* Virials (sys/compute.rs:198-364)                                          */
/* ------------------------------------------------------------------------ */

static inline void mat_add(double a[9], const double b[9]) {
    for (int k = 0; k < 9; k++) a[k] += b[k];
}

/* compute.rs:202-216 */
void orc_pair_atomic_virial(const orc_system* s, double w[9]) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    int nt = max_threads();
    double* partial = (double*)calloc((size_t)nt * 9, sizeof(double));
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
        double local_virial[9] = {0};
        for (int64_t j = i + 1; j < n; j++) {
            int32_t path = orc_bond_path(s, i, j);
            const orc_pair* potential = pair_potential(s, i, j);
            if (potential != NULL) {
                int32_t excluded;
                double scaling;
                orc_restriction_information(potential->restriction, potential->scale14, path, &excluded, &scaling);
                if (!excluded) {
                    double d[3], wp[9];
                    nearest_image(s, &g, i, j, d);
                    orc_pair_virial(potential, d, wp);
                    for (int k = 0; k < 9; k++) {
                        local_virial[k] += scaling * wp[k];
                    }
                }
            }
        }
        mat_add(partial + 9 * thread_id(), local_virial);
    }
    memset(w, 0, 9 * sizeof(double));
    for (int t = 0; t < nt; t++) {
        mat_add(w, partial + 9 * t);
    }
    free(partial);
}

/* compute.rs:219-228 with pairs.rs:289-296 */
void orc_tail_virial(const orc_system* s, double w[9]) {
    memset(w, 0, 9 * sizeof(double));
    double volume = orc_cell_volume(s->cell, s->shape);
    int64_t* counts = composition(s);
    for (int32_t i = 0; i < s->nkinds; i++) {
        if (counts[i] == 0) continue;
        for (int32_t j = 0; j < s->nkinds; j++) {
            if (counts[j] == 0) continue;
            double two_pi_density = 2.0 * PI * (double)counts[i] * (double)counts[j] / volume;
            const orc_pair* potential = &s->pairs[(int64_t)i * s->nkinds + j];
            if (potential->potential.pot != ORC_POT_ABSENT && potential->tail) {
                /* tensor = Matrix3::one() / 3.0; tail_virial * tensor */
                double t = orc_pair_tail_virial(potential) * (1.0 / 3.0);
                w[0] += two_pi_density * t;
                w[4] += two_pi_density * t;
                w[8] += two_pi_density * t;
            }
        }
    }
    free(counts);
}

/* compute.rs:231-239; BondPotential::virial has the same default body as PairPotential::virial (energy/mod.rs:170-176) */
void orc_bond_virial(const orc_system* s, double w[9]) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    memset(w, 0, 9 * sizeof(double));
    for (int64_t b = 0; b < s->nbonds; b++) {
        int64_t i = s->bonds[2 * b], j = s->bonds[2 * b + 1];
        double r[3], wb[9];
        nearest_image(s, &g, i, j, r);
        if (s->bond_pot[b].pot != ORC_POT_ABSENT) {
            orc_potential_virial(&s->bond_pot[b], r, wb);
            mat_add(w, wb);
        }
    }
}

void orc_coulomb_atomic_virial(const orc_system* s, double w[9]) {
    memset(w, 0, 9 * sizeof(double));
    if (s->coulomb == ORC_COULOMB_EWALD) {
        /* ewald.rs:907-914 */
        double real[9], kspace[9];
        orc_ewald_real_atomic_virial(s, real);
        orc_ewald_kspace_atomic_virial(s, kspace);
        for (int k = 0; k < 9; k++) w[k] = real[k] + kspace[k];
    } else if (s->coulomb == ORC_COULOMB_WOLF) {
        orc_wolf_atomic_virial(s, w);
    }
}

/* compute.rs:198-254 */
void orc_atomic_virial(const orc_system* s, double w[9]) {
    double t[9];
    orc_pair_atomic_virial(s, w);
    orc_tail_virial(s, t);
    mat_add(w, t);
    orc_bond_virial(s, t);
    mat_add(w, t);
    if (s->coulomb != ORC_COULOMB_NONE) {
        orc_coulomb_atomic_virial(s, t);
        mat_add(w, t);
    }
}

/* molecules.rs:256-264 */
static void center_of_mass(const orc_system* s, int64_t mol, double com[3]) {
    double total_mass = 0.0;
    com[0] = com[1] = com[2] = 0.0;
    for (int64_t i = s->mol_start[mol]; i < s->mol_start[mol + 1]; i++) {
        total_mass += s->mass[i];
        for (int k = 0; k < 3; k++) {
            com[k] += s->mass[i] * s->position[3 * i + k];
        }
    }
    for (int k = 0; k < 3; k++) {
        com[k] /= total_mass;
    }
}

/* Kind of pair term summed by the molecular virial loops */
enum { MV_PAIRS = 0, MV_EWALD = 1, MV_WOLF = 2 };

static double ewald_real_force_pair(const orc_system* s, int32_t excluded, double qiqj, double r);
static double wolf_force_pair(const orc_system* s, double qiqj, double rij, double force_constant);
static void wolf_constants(const orc_system* s, double* alpha, double* energy_constant, double* force_constant);

/* compute.rs:286-311, ewald.rs:507-545 and wolf.rs:287-325 share this molecule-pair loop */
static void molecular_pair_virial(const orc_system* s, int mode, double w[9]) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int nt = max_threads();
    double* partial = (double*)calloc((size_t)nt * 9, sizeof(double));
    double walpha = 0, wec = 0, wfc = 0;
    if (mode == MV_WOLF) {
        wolf_constants(s, &walpha, &wec, &wfc);
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t mi = 0; mi < s->nmol; mi++) {
        double local_virial[9] = {0};
        double ri[3];
        center_of_mass(s, mi, ri);
        for (int64_t mj = mi + 1; mj < s->nmol; mj++) {
            double rj[3];
            center_of_mass(s, mj, rj);
            double r_ij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
            geom_image(&g, r_ij);
            for (int64_t a = s->mol_start[mi]; a < s->mol_start[mi + 1]; a++) {
                if (mode != MV_PAIRS && s->charge[a] == 0.0) continue;
                for (int64_t b = s->mol_start[mj]; b < s->mol_start[mj + 1]; b++) {
                    if (mode != MV_PAIRS && s->charge[b] == 0.0) continue;
                    int32_t path = orc_bond_path(s, a, b);
                    double r_ab[3];
                    nearest_image(s, &g, a, b, r_ab);
                    double w_ab[9];
                    if (mode == MV_PAIRS) {
                        const orc_pair* potential = pair_potential(s, a, b);
                        if (potential == NULL) continue;
                        int32_t excluded;
                        double scaling;
                        orc_restriction_information(potential->restriction, potential->scale14, path, &excluded,
                                                    &scaling);
                        if (excluded) continue;
                        orc_pair_virial(potential, r_ab, w_ab);
                        for (int k = 0; k < 9; k++) w_ab[k] = scaling * w_ab[k];
                    } else {
                        int32_t excluded;
                        double scaling;
                        orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded,
                                                    &scaling);
                        double fr;
                        if (mode == MV_EWALD) {
                            fr = ewald_real_force_pair(s, excluded, s->charge[a] * s->charge[b], norm3(r_ab));
                        } else {
                            if (excluded) continue;
                            fr = scaling * wolf_force_pair(s, s->charge[a] * s->charge[b], norm3(r_ab), wfc);
                        }
                        double force[3] = {fr * r_ab[0], fr * r_ab[1], fr * r_ab[2]};
                        for (int p = 0; p < 3; p++)
                            for (int q = 0; q < 3; q++) w_ab[3 * p + q] = force[p] * r_ab[q];
                    }
                    /* w_ab * (r_ab * r_ij) / r_ab.norm2(): (Matrix3 * f64) / f64 */
                    double proj = dot3(r_ab, r_ij);
                    double n2 = norm2_3(r_ab);
                    for (int k = 0; k < 9; k++) {
                        local_virial[k] += w_ab[k] * proj / n2;
                    }
                }
            }
        }
        mat_add(partial + 9 * thread_id(), local_virial);
    }
    memset(w, 0, 9 * sizeof(double));
    for (int t = 0; t < nt; t++) {
        mat_add(w, partial + 9 * t);
    }
    free(partial);
}

void orc_coulomb_molecular_virial(const orc_system* s, double w[9]) {
    memset(w, 0, 9 * sizeof(double));
    if (s->coulomb == ORC_COULOMB_EWALD) {
        /* ewald.rs:916-923 and :736-753 */
        double real[9], atomic[9];
        molecular_pair_virial(s, MV_EWALD, real);
        orc_ewald_kspace_atomic_virial(s, atomic);
        double* forces = (double*)calloc((size_t)(3 * s->n > 0 ? 3 * s->n : 1), sizeof(double));
        orc_ewald_kspace_forces(s, forces);
        double correction[9] = {0};
        for (int64_t mol = 0; mol < s->nmol; mol++) {
            double com[3];
            center_of_mass(s, mol, com);
            for (int64_t i = s->mol_start[mol]; i < s->mol_start[mol + 1]; i++) {
                double di[3] = {s->position[3 * i] - com[0], s->position[3 * i + 1] - com[1],
                                s->position[3 * i + 2] - com[2]};
                for (int p = 0; p < 3; p++)
                    for (int q = 0; q < 3; q++) correction[3 * p + q] += forces[3 * i + p] * di[q];
            }
        }
        free(forces);
        for (int k = 0; k < 9; k++) w[k] = real[k] + (atomic[k] - correction[k]);
    } else if (s->coulomb == ORC_COULOMB_WOLF) {
        molecular_pair_virial(s, MV_WOLF, w);
    }
}

/* compute.rs:281-363 (bond potentials are ignored with a warning there) */
void orc_molecular_virial(const orc_system* s, double w[9]) {
    double t[9];
    molecular_pair_virial(s, MV_PAIRS, w);
    orc_tail_virial(s, t);
    mat_add(w, t);
    if (s->coulomb != ORC_COULOMB_NONE) {
        orc_coulomb_molecular_virial(s, t);
        mat_add(w, t);
    }
}

/* ------------------------------------------------------------------------ */
/* kinetic side (sys/compute.rs:134-171, 393-480; system.rs:249-255)          */
/* ------------------------------------------------------------------------ */

double orc_kinetic_energy(const orc_system* s) {
    double energy = 0.0;
    for (int64_t i = 0; i < s->n; i++) {
        energy += 0.5 * s->mass[i] * norm2_3(s->velocity + 3 * i);
    }
    return energy;
}

int64_t orc_degrees_of_freedom(const orc_system* s) {
    if (s->dof_mode == 1) {
        return 3 * s->nmol;
    }
    return 3 * s->n - s->dof_frozen;
}

double orc_temperature(const orc_system* s) {
    double kinetic = orc_kinetic_energy(s);
    double dof = (double)orc_degrees_of_freedom(s);
    return 2.0 * kinetic / (dof * ORC_K_BOLTZMANN);
}

static void system_virial(const orc_system* s, double w[9]) {
    /* compute.rs:375-380 */
    if (s->dof_mode == 1) {
        orc_molecular_virial(s, w);
    } else {
        orc_atomic_virial(s, w);
    }
}

double orc_pressure_at_temperature(const orc_system* s, double temperature) {
    double w[9];
    system_virial(s, w);
    double virial = w[0] + w[4] + w[8];
    double volume = orc_cell_volume(s->cell, s->shape);
    double dof = (double)orc_degrees_of_freedom(s);
    return (dof * ORC_K_BOLTZMANN * temperature + virial) / (3.0 * volume);
}

double orc_pressure(const orc_system* s) { return orc_pressure_at_temperature(s, orc_temperature(s)); }

void orc_stress_at_temperature(const orc_system* s, double temperature, double out[9]) {
    double w[9];
    system_virial(s, w);
    double volume = orc_cell_volume(s->cell, s->shape);
    double dof = (double)orc_degrees_of_freedom(s);
    double kin = dof / 3.0 * ORC_K_BOLTZMANN * temperature;
    for (int k = 0; k < 9; k++) {
        double kinetic = (k == 0 || k == 4 || k == 8) ? kin * 1.0 : kin * 0.0;
        out[k] = (kinetic + w[k]) / volume;
    }
}

void orc_stress(const orc_system* s, double out[9]) {
    double kinetic[9] = {0};
    for (int64_t i = 0; i < s->n; i++) {
        const double* v = s->velocity + 3 * i;
        for (int p = 0; p < 3; p++)
            for (int q = 0; q < 3; q++) kinetic[3 * p + q] += s->mass[i] * (v[p] * v[q]);
    }
    double volume = orc_cell_volume(s->cell, s->shape);
    double w[9];
    system_virial(s, w);
    for (int k = 0; k < 9; k++) {
        out[k] = (kinetic[k] + w[k]) / volume;
    }
}

/* ------------------------------------------------------------------------ */
/* Ewald (energy/global/ewald.rs)                                            */
/* ------------------------------------------------------------------------ */

typedef struct {
    int64_t nk;
    int64_t* index;  /* 3 per k */
    double* energy;  /* 1 per k */
    double* field;   /* 3 per k */
    double* virial;  /* 9 per k */
    double kmax2;
} ewald_factors_t;

/* ewald.rs:134-141 */
static void factor_from_k_vector(ewald_factors_t* f, const double k_vector[3], double k2, int64_t ikx, int64_t iky,
                                 int64_t ikz, double alpha_sq_inv_fourth, double four_pi_v) {
    int64_t n = f->nk;
    double energy = four_pi_v * exp(-k2 * alpha_sq_inv_fourth) / k2;
    double two_e = 2.0 * energy;
    double virial_factor = -2.0 * (1.0 / k2 + alpha_sq_inv_fourth);
    f->index[3 * n] = ikx;
    f->index[3 * n + 1] = iky;
    f->index[3 * n + 2] = ikz;
    f->energy[n] = energy;
    for (int a = 0; a < 3; a++) {
        f->field[3 * n + a] = two_e * k_vector[a];
    }
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) {
            /* Matrix3::one() + virial_factor * k.tensorial(k), then energy * that */
            double one = a == b ? 1.0 : 0.0;
            double v = one + virial_factor * (k_vector[a] * k_vector[b]);
            f->virial[9 * n + 3 * a + b] = energy * v;
        }
    }
    f->nk = n + 1;
}

/* ewald.rs:353-378 prepare + ewald.rs:115-183 compute_ewald_factors */
static void ewald_factors_init(ewald_factors_t* f, const orc_system* s, const geom_t* g) {
    int64_t kmax = s->kmax;
    int64_t kmax3d = 4 * kmax * kmax * kmax + 6 * kmax * kmax + 3 * kmax;
    f->nk = 0;
    f->index = (int64_t*)malloc((size_t)kmax3d * 3 * sizeof(int64_t));
    f->energy = (double*)malloc((size_t)kmax3d * sizeof(double));
    f->field = (double*)malloc((size_t)kmax3d * 3 * sizeof(double));
    f->virial = (double*)malloc((size_t)kmax3d * 9 * sizeof(double));

    double ones[3] = {1.0, 1.0, 1.0};
    double k111[3];
    geom_k_vector(g, ones, k111);
    double max = fmax(fmax(k111[0], k111[1]), k111[2]) * (double)kmax;
    f->kmax2 = 1.0001 * max * max;

    double alpha_sq_inv_fourth = 0.25 / (s->alpha * s->alpha);
    double four_pi_v = 4.0 * PI / orc_cell_volume(s->cell, s->shape);

    for (int64_t ikx = 1; ikx < kmax; ikx++) {
        for (int64_t iky = -kmax; iky < kmax; iky++) {
            for (int64_t ikz = -kmax; ikz < kmax; ikz++) {
                double idx[3] = {(double)ikx, (double)iky, (double)ikz};
                double kv[3];
                geom_k_vector(g, idx, kv);
                double k2 = norm2_3(kv);
                if (k2 > f->kmax2) continue;
                factor_from_k_vector(f, kv, k2, ikx, iky, ikz, alpha_sq_inv_fourth, four_pi_v);
            }
        }
    }
    for (int64_t iky = 1; iky < kmax; iky++) {
        for (int64_t ikz = -kmax; ikz < kmax; ikz++) {
            double idx[3] = {0.0, (double)iky, (double)ikz};
            double kv[3];
            geom_k_vector(g, idx, kv);
            double k2 = norm2_3(kv);
            if (k2 > f->kmax2) continue;
            factor_from_k_vector(f, kv, k2, 0, iky, ikz, alpha_sq_inv_fourth, four_pi_v);
        }
    }
    for (int64_t ikz = 1; ikz < kmax; ikz++) {
        double idx[3] = {0.0, 0.0, (double)ikz};
        double kv[3];
        geom_k_vector(g, idx, kv);
        double k2 = norm2_3(kv);
        if (k2 > f->kmax2) continue;
        factor_from_k_vector(f, kv, k2, 0, 0, ikz, alpha_sq_inv_fourth, four_pi_v);
    }
}

static void ewald_factors_free(ewald_factors_t* f) {
    free(f->index);
    free(f->energy);
    free(f->field);
    free(f->virial);
}

int64_t orc_ewald_factors(const orc_system* s, double* kmax2_out, int64_t capacity, int64_t* index, double* energy,
                          double* field, double* virial) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    ewald_factors_t f;
    ewald_factors_init(&f, s, &g);
    int64_t nk = f.nk;
    if (kmax2_out) *kmax2_out = f.kmax2;
    if (capacity >= nk) {
        if (index) memcpy(index, f.index, (size_t)nk * 3 * sizeof(int64_t));
        if (energy) memcpy(energy, f.energy, (size_t)nk * sizeof(double));
        if (field) memcpy(field, f.field, (size_t)nk * 3 * sizeof(double));
        if (virial) memcpy(virial, f.virial, (size_t)nk * 9 * sizeof(double));
    }
    ewald_factors_free(&f);
    return nk;
}

/* ewald.rs:387-400 */
static double ewald_real_energy_pair(const orc_system* s, int32_t excluded, double qiqj, double r) {
    if (r > s->rc) {
        return 0.0;
    }
    if (excluded) {
        return -qiqj / ORC_FOUR_PI_EPSILON_0 * erf(s->alpha * r) / r;
    }
    return qiqj / ORC_FOUR_PI_EPSILON_0 * erfc(s->alpha * r) / r;
}

/* ewald.rs:408-428 */
static double ewald_real_force_pair(const orc_system* s, int32_t excluded, double qiqj, double r) {
    if (r > s->rc) {
        return 0.0;
    }
    if (excluded) {
        return qiqj / (ORC_FOUR_PI_EPSILON_0 * r * r) *
               (s->alpha * FRAC_2_SQRT_PI * exp(-s->alpha * s->alpha * r * r) - erf(s->alpha * r) / r);
    }
    return qiqj / (ORC_FOUR_PI_EPSILON_0 * r * r) *
           (s->alpha * FRAC_2_SQRT_PI * exp(-s->alpha * s->alpha * r * r) + erfc(s->alpha * r) / r);
}

/* ewald.rs:430-457 */
double orc_ewald_real_energy(const orc_system* s) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    double total = 0.0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
    for (int64_t i = 0; i < n; i++) {
        double local_energy = 0.0;
        double qi = s->charge[i];
        if (qi == 0.0) continue;
        for (int64_t j = i + 1; j < n; j++) {
            double qj = s->charge[j];
            if (qj == 0.0) continue;
            int32_t path = orc_bond_path(s, i, j);
            int32_t excluded;
            double scaling;
            orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
            double r = distance(s, &g, i, j);
            local_energy += ewald_real_energy_pair(s, excluded, qi * qj, r);
        }
        total += local_energy;
    }
    return total;
}

/* ewald.rs:461-500 */
void orc_ewald_real_forces(const orc_system* s, double* out) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    thread_vec_t tv;
    thread_vec_init(&tv, 3 * n);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
        double* forces = tv.data + (size_t)thread_id() * (size_t)(3 * n);
        double force_i[3] = {0.0, 0.0, 0.0};
        double qi = s->charge[i];
        if (qi == 0.0) continue;
        for (int64_t j = i + 1; j < n; j++) {
            double qj = s->charge[j];
            if (qj == 0.0) continue;
            int32_t path = orc_bond_path(s, i, j);
            int32_t excluded;
            double scaling;
            orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
            double rij[3];
            nearest_image(s, &g, i, j, rij);
            double fr = ewald_real_force_pair(s, excluded, qi * qj, norm3(rij));
            for (int k = 0; k < 3; k++) {
                double force = fr * rij[k];
                force_i[k] += force;
                forces[3 * j + k] -= force;
            }
        }
        for (int k = 0; k < 3; k++) {
            forces[3 * i + k] += force_i[k];
        }
    }
    thread_vec_sum_into(&tv, out);
}

/* The same loop body (ewald.rs:461-500) for the atoms listed in `rows` against every other atom.  out: nrows x 3. */
void orc_ewald_real_forces_rows(const orc_system* s, int64_t nrows, const int64_t* rows, double* out) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t row = 0; row < nrows; row++) {
        int64_t t = rows[row];
        double total[3] = {0.0, 0.0, 0.0};
        if (s->charge[t] != 0.0) {
            for (int64_t o = 0; o < n; o++) {
                if (o == t || s->charge[o] == 0.0) continue;
                int64_t i = t < o ? t : o, j = t < o ? o : t;
                int32_t path = orc_bond_path(s, i, j);
                int32_t excluded;
                double scaling;
                orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
                double rij[3];
                nearest_image(s, &g, i, j, rij);
                double fr = ewald_real_force_pair(s, excluded, s->charge[i] * s->charge[j], norm3(rij));
                for (int k = 0; k < 3; k++) {
                    double force = fr * rij[k];
                    if (t == i) {
                        total[k] += force;
                    } else {
                        total[k] -= force;
                    }
                }
            }
        }
        for (int k = 0; k < 3; k++) out[3 * row + k] = total[k];
    }
}

/* ewald.rs:502-530 */
void orc_ewald_real_atomic_virial(const orc_system* s, double w[9]) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    int64_t n = s->n;
    int nt = max_threads();
    double* partial = (double*)calloc((size_t)nt * 9, sizeof(double));
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
        double qi = s->charge[i];
        if (qi == 0.0) continue;
        double local_virial[9] = {0};
        for (int64_t j = i + 1; j < n; j++) {
            double qj = s->charge[j];
            if (qj == 0.0) continue;
            int32_t path = orc_bond_path(s, i, j);
            int32_t excluded;
            double scaling;
            orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
            double rij[3];
            nearest_image(s, &g, i, j, rij);
            double fr = ewald_real_force_pair(s, excluded, qi * qj, norm3(rij));
            double force[3] = {fr * rij[0], fr * rij[1], fr * rij[2]};
            for (int p = 0; p < 3; p++)
                for (int q = 0; q < 3; q++) local_virial[3 * p + q] += force[p] * rij[q];
        }
        mat_add(partial + 9 * thread_id(), local_virial);
    }
    memset(w, 0, 9 * sizeof(double));
    for (int t = 0; t < nt; t++) {
        mat_add(w, partial + 9 * t);
    }
    free(partial);
}

/* ewald.rs:619-626 */
double orc_ewald_self_energy(const orc_system* s) {
    double q2 = 0.0;
    for (int64_t i = 0; i < s->n; i++) {
        q2 += s->charge[i] * s->charge[i];
    }
    return -s->alpha / sqrt(PI) * q2 / ORC_FOUR_PI_EPSILON_0;
}

typedef struct {
    double re, im;
} cplx;

/* complex.rs:219-228 */
static inline cplx cmul(cplx a, cplx b) {
    cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
    return r;
}

typedef struct {
    int64_t kmax;
    int64_t n;
    cplx* eikr; /* [(k + kmax) * 3 + spatial] * n + i */
    cplx* rho;
} ewald_kspace_t;

static inline cplx* eikr_at(const ewald_kspace_t* ks, int64_t k, int spatial, int64_t i) {
    return &ks->eikr[((k + ks->kmax) * 3 + spatial) * ks->n + i];
}

/* ewald.rs:633-674 eik_dot_r */
static void eik_dot_r(ewald_kspace_t* ks, const orc_system* s, const geom_t* g, const ewald_factors_t* f) {
    int64_t n = s->n;
    int64_t kmax = s->kmax;
    ks->kmax = kmax;
    ks->n = n;
    ks->eikr = (cplx*)malloc((size_t)((2 * kmax + 1) * 3 * (n > 0 ? n : 1)) * sizeof(cplx));
    ks->rho = (cplx*)malloc((size_t)(f->nk > 0 ? f->nk : 1) * sizeof(cplx));

    for (int spatial = 0; spatial < 3; spatial++) {
        double k_idx[3] = {0.0, 0.0, 0.0};
        k_idx[spatial] = 1.0;
        double k_vector[3];
        geom_k_vector(g, k_idx, k_vector);
        for (int64_t i = 0; i < n; i++) {
            double phi = dot3(k_vector, s->position + 3 * i);
            cplx one = {1.0, 0.0};
            /* Complex::polar(1.0, phi) = (1.0 * cos, 1.0 * sin), complex.rs:58-63 */
            cplx e1 = {1.0 * cos(phi), 1.0 * sin(phi)};
            cplx em1 = {e1.re, -e1.im};
            *eikr_at(ks, 0, spatial, i) = one;
            *eikr_at(ks, 1, spatial, i) = e1;
            *eikr_at(ks, -1, spatial, i) = em1;
        }
    }
    for (int spatial = 0; spatial < 3; spatial++) {
        for (int64_t k = 2; k < kmax + 1; k++) {
            for (int64_t i = 0; i < n; i++) {
                cplx v = cmul(*eikr_at(ks, k - 1, spatial, i), *eikr_at(ks, 1, spatial, i));
                cplx c = {v.re, -v.im};
                *eikr_at(ks, k, spatial, i) = v;
                *eikr_at(ks, -k, spatial, i) = c;
            }
        }
    }
    /* The reference runs this loop serially; each rho(k) is an independent serial sum over atoms, so
     * distributing k-vectors over threads changes no rounding. */
#pragma omp parallel for schedule(static)
    for (int64_t ik = 0; ik < f->nk; ik++) {
        int64_t ikx = f->index[3 * ik], iky = f->index[3 * ik + 1], ikz = f->index[3 * ik + 2];
        cplx partial = {0.0, 0.0};
        for (int64_t i = 0; i < n; i++) {
            cplx phi = cmul(cmul(*eikr_at(ks, ikx, 0, i), *eikr_at(ks, iky, 1, i)), *eikr_at(ks, ikz, 2, i));
            partial.re += s->charge[i] * phi.re;
            partial.im += s->charge[i] * phi.im;
        }
        ks->rho[ik] = partial;
    }
}

static void kspace_free(ewald_kspace_t* ks) {
    free(ks->eikr);
    free(ks->rho);
}

void orc_ewald_rho(const orc_system* s, int64_t capacity, double* rho) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    ewald_factors_t f;
    ewald_factors_init(&f, s, &g);
    ewald_kspace_t ks;
    eik_dot_r(&ks, s, &g, &f);
    for (int64_t ik = 0; ik < f.nk && ik < capacity; ik++) {
        rho[2 * ik] = ks.rho[ik].re;
        rho[2 * ik + 1] = ks.rho[ik].im;
    }
    kspace_free(&ks);
    ewald_factors_free(&f);
}

/* ewald.rs:677-687 */
double orc_ewald_kspace_energy(const orc_system* s) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    ewald_factors_t f;
    ewald_factors_init(&f, s, &g);
    ewald_kspace_t ks;
    eik_dot_r(&ks, s, &g, &f);
    double energy = 0.0;
    for (int64_t ik = 0; ik < f.nk; ik++) {
        /* complex.rs norm2 = real^2 + imag^2 */
        double n2 = ks.rho[ik].re * ks.rho[ik].re + ks.rho[ik].im * ks.rho[ik].im;
        energy += f.energy[ik] * n2;
    }
    kspace_free(&ks);
    ewald_factors_free(&f);
    return energy / ORC_FOUR_PI_EPSILON_0;
}

/* ewald.rs:690-720 */
void orc_ewald_kspace_forces(const orc_system* s, double* forces) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    ewald_factors_t f;
    ewald_factors_init(&f, s, &g);
    ewald_kspace_t ks;
    eik_dot_r(&ks, s, &g, &f);
    int64_t n = s->n;
    thread_vec_t tv;
    thread_vec_init(&tv, 3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t ik = 0; ik < f.nk; ik++) {
        double* field = tv.data + (size_t)thread_id() * (size_t)(3 * n);
        int64_t ikx = f.index[3 * ik], iky = f.index[3 * ik + 1], ikz = f.index[3 * ik + 2];
        cplx rho_conj = {ks.rho[ik].re, -ks.rho[ik].im};
        for (int64_t i = 0; i < n; i++) {
            cplx e = cmul(cmul(*eikr_at(&ks, ikx, 0, i), *eikr_at(&ks, iky, 1, i)), *eikr_at(&ks, ikz, 2, i));
            cplx partial = cmul(e, rho_conj);
            for (int c = 0; c < 3; c++) {
                field[3 * i + c] += partial.im * f.field[3 * ik + c];
            }
        }
    }
    double* field = (double*)calloc((size_t)(3 * n > 0 ? 3 * n : 1), sizeof(double));
    thread_vec_sum_into(&tv, field);
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            forces[3 * i + c] += s->charge[i] * field[3 * i + c] / ORC_FOUR_PI_EPSILON_0;
        }
    }
    free(field);
    kspace_free(&ks);
    ewald_factors_free(&f);
}

/* ewald.rs:722-733 */
void orc_ewald_kspace_atomic_virial(const orc_system* s, double w[9]) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    ewald_factors_t f;
    ewald_factors_init(&f, s, &g);
    ewald_kspace_t ks;
    eik_dot_r(&ks, s, &g, &f);
    memset(w, 0, 9 * sizeof(double));
    for (int64_t ik = 0; ik < f.nk; ik++) {
        double n2 = ks.rho[ik].re * ks.rho[ik].re + ks.rho[ik].im * ks.rho[ik].im;
        for (int k = 0; k < 9; k++) {
            w[k] += n2 * f.virial[9 * ik + k];
        }
    }
    for (int k = 0; k < 9; k++) {
        w[k] = w[k] / ORC_FOUR_PI_EPSILON_0;
    }
    kspace_free(&ks);
    ewald_factors_free(&f);
}

/* ewald.rs:312-351 */
void orc_ewald_with_accuracy(const orc_system* s, double cutoff, double accuracy, double* alpha_out, int32_t* kmax_out) {
    double q2 = 0.0;
    for (int64_t i = 0; i < s->n; i++) {
        q2 += s->charge[i] * s->charge[i];
    }
    q2 /= ORC_FOUR_PI_EPSILON_0;
    double natoms = (double)s->n;
    double lengths[3];
    orc_cell_lengths(s->cell, s->shape, lengths);
    double alpha = accuracy * sqrt(natoms * cutoff * lengths[0] * lengths[1] * lengths[2]) / (2.0 * q2);
    if (alpha >= 1.0) {
        alpha = (1.35 - 0.15 * log(accuracy)) / cutoff;
    } else {
        alpha = sqrt(-log(alpha)) / cutoff;
    }
    double min_length = fmin(fmin(lengths[0], lengths[1]), lengths[2]);
    int32_t kmax = 1;
    for (;;) {
        double k = (double)kmax;
        double arg = PI * k / (alpha * min_length);
        double error = FRAC_2_SQRT_PI * q2 * alpha / min_length / sqrt(k * natoms) * exp(-arg * arg);
        if (!(error > accuracy)) break;
        kmax += 1;
    }
    *alpha_out = alpha;
    *kmax_out = kmax;
}

/* ------------------------------------------------------------------------ */
/* Wolf (energy/global/wolf.rs)                                              */
/* ------------------------------------------------------------------------ */

/* wolf.rs:68-84; alpha = PI / cutoff */
static void wolf_constants(const orc_system* s, double* alpha, double* energy_constant, double* force_constant) {
    double cutoff = s->rc;
    double a = PI / cutoff;
    double alpha_cutoff = a * cutoff;
    double alpha_cutoff_2 = alpha_cutoff * alpha_cutoff;
    *alpha = a;
    *energy_constant = erfc(alpha_cutoff) / cutoff;
    *force_constant = erfc(alpha_cutoff) / (cutoff * cutoff) + FRAC_2_SQRT_PI * a * exp(-alpha_cutoff_2) / cutoff;
}

/* wolf.rs:106-117 */
static double wolf_force_pair(const orc_system* s, double qiqj, double rij, double force_constant) {
    if (rij > s->rc) {
        return 0.0;
    }
    double alpha = PI / s->rc;
    double rij2 = rij * rij;
    double alpha_rij = alpha * rij;
    double exp_alpha_rij = exp(-alpha_rij * alpha_rij);
    double factor = erfc(alpha_rij) / rij2 + alpha * FRAC_2_SQRT_PI * exp_alpha_rij / rij;
    return qiqj * (factor - force_constant) / (rij * ORC_FOUR_PI_EPSILON_0);
}

/* wolf.rs:177-207 */
double orc_wolf_energy(const orc_system* s) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    double alpha, ec, fc;
    wolf_constants(s, &alpha, &ec, &fc);
    int64_t n = s->n;
    double total = 0.0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
    for (int64_t i = 0; i < n; i++) {
        double energy = 0.0;
        double qi = s->charge[i];
        if (qi == 0.0) continue;
        for (int64_t j = i + 1; j < n; j++) {
            double qj = s->charge[j];
            if (qj == 0.0) continue;
            int32_t path = orc_bond_path(s, i, j);
            int32_t excluded;
            double scaling;
            orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
            if (excluded) continue;
            double rij = distance(s, &g, i, j);
            /* wolf.rs:90-96 energy_pair */
            double e = rij > s->rc ? 0.0 : (qi * qj) * (erfc(alpha * rij) / rij - ec) / ORC_FOUR_PI_EPSILON_0;
            energy += scaling * e;
        }
        /* wolf.rs:99-101 energy_self */
        double self = qi * qi * 0.5 * (ec + alpha * FRAC_2_SQRT_PI) / ORC_FOUR_PI_EPSILON_0;
        total += energy - self;
    }
    return total;
}

/* wolf.rs:209-249 */
void orc_wolf_forces(const orc_system* s, double* out) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    double alpha, ec, fc;
    wolf_constants(s, &alpha, &ec, &fc);
    int64_t n = s->n;
    thread_vec_t tv;
    thread_vec_init(&tv, 3 * n);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
        double* forces = tv.data + (size_t)thread_id() * (size_t)(3 * n);
        double force_i[3] = {0.0, 0.0, 0.0};
        double qi = s->charge[i];
        if (qi == 0.0) continue;
        for (int64_t j = i + 1; j < n; j++) {
            double qj = s->charge[j];
            if (qj == 0.0) continue;
            int32_t path = orc_bond_path(s, i, j);
            int32_t excluded;
            double scaling;
            orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
            if (excluded) continue;
            double rij[3];
            nearest_image(s, &g, i, j, rij);
            double fr = scaling * wolf_force_pair(s, qi * qj, norm3(rij), fc);
            for (int k = 0; k < 3; k++) {
                double force = fr * rij[k];
                force_i[k] += force;
                forces[3 * j + k] -= force;
            }
        }
        for (int k = 0; k < 3; k++) {
            forces[3 * i + k] += force_i[k];
        }
    }
    thread_vec_sum_into(&tv, out);
}

/* wolf.rs:251-283 */
void orc_wolf_atomic_virial(const orc_system* s, double w[9]) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    double alpha, ec, fc;
    wolf_constants(s, &alpha, &ec, &fc);
    int64_t n = s->n;
    int nt = max_threads();
    double* partial = (double*)calloc((size_t)nt * 9, sizeof(double));
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
        double qi = s->charge[i];
        if (qi == 0.0) continue;
        double local_virial[9] = {0};
        for (int64_t j = i + 1; j < n; j++) {
            double qj = s->charge[j];
            if (qj == 0.0) continue;
            int32_t path = orc_bond_path(s, i, j);
            int32_t excluded;
            double scaling;
            orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
            if (excluded) continue;
            double rij[3];
            nearest_image(s, &g, i, j, rij);
            double fr = scaling * wolf_force_pair(s, qi * qj, norm3(rij), fc);
            double force[3] = {fr * rij[0], fr * rij[1], fr * rij[2]};
            for (int p = 0; p < 3; p++)
                for (int q = 0; q < 3; q++) local_virial[3 * p + q] += force[p] * rij[q];
        }
        mat_add(partial + 9 * thread_id(), local_virial);
    }
    memset(w, 0, 9 * sizeof(double));
    for (int t = 0; t < nt; t++) {
        mat_add(w, partial + 9 * t);
    }
    free(partial);
}

/* ------------------------------------------------------------------------ */
/* Integrators / thermostats / controls (lumol-sim/src/md)                   */
/* ------------------------------------------------------------------------ */

/* integrators.rs:44-69.  `s->position`/`s->velocity` must alias `position`/`velocity`. */
void orc_velocity_verlet_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt) {
    int64_t n = s->n;
    for (int64_t k = 0; k < 3 * n; k++) {
        /* 0.5 * dt * acceleration: (0.5 * dt) * a */
        velocity[k] += 0.5 * dt * accelerations[k];
        position[k] += velocity[k] * dt;
    }
    s->position = position;
    s->velocity = velocity;
    double* forces = (double*)malloc((size_t)(3 * n > 0 ? 3 * n : 1) * sizeof(double));
    orc_forces(s, forces);
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            accelerations[3 * i + c] = forces[3 * i + c] / s->mass[i];
        }
    }
    for (int64_t k = 0; k < 3 * n; k++) {
        velocity[k] += 0.5 * dt * accelerations[k];
    }
    free(forces);
}

/* integrators.rs:92-101 */
void orc_verlet_setup(const orc_system* s, double* prevpos, double dt) {
    for (int64_t k = 0; k < 3 * s->n; k++) {
        prevpos[k] = s->position[k] - s->velocity[k] * dt;
    }
}

/* integrators.rs:103-122 */
void orc_verlet_step(orc_system* s, double* position, double* velocity, double* prevpos, double dt) {
    int64_t n = s->n;
    s->position = position;
    s->velocity = velocity;
    double* forces = (double*)malloc((size_t)(3 * n > 0 ? 3 * n : 1) * sizeof(double));
    orc_forces(s, forces);
    double dt2 = dt * dt;
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            int64_t k = 3 * i + c;
            double tmp = position[k];
            /* 2.0 * position - prevpos + dt2 / mass * force */
            position[k] = 2.0 * position[k] - prevpos[k] + dt2 / s->mass[i] * forces[k];
            velocity[k] = (position[k] - prevpos[k]) / (2.0 * dt);
            prevpos[k] = tmp;
        }
    }
    free(forces);
}

/* integrators.rs:145-169 */
void orc_leapfrog_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt) {
    int64_t n = s->n;
    double dt2 = dt * dt;
    for (int64_t k = 0; k < 3 * n; k++) {
        /* velocity * dt + 0.5 * acceleration * dt2 */
        position[k] += velocity[k] * dt + 0.5 * accelerations[k] * dt2;
    }
    s->position = position;
    s->velocity = velocity;
    double* forces = (double*)malloc((size_t)(3 * n > 0 ? 3 * n : 1) * sizeof(double));
    orc_forces(s, forces);
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            int64_t k = 3 * i + c;
            double new_acceleration = forces[k] / s->mass[i];
            velocity[k] += 0.5 * (accelerations[k] + new_acceleration) * dt;
            accelerations[k] = new_acceleration;
        }
    }
    free(forces);
}

/* velocities.rs:16-22 and thermostats.rs:112-120 apply `v *= factor` */
void orc_scale_velocities(int64_t n, double* velocity, double factor) {
    for (int64_t k = 0; k < 3 * n; k++) {
        velocity[k] *= factor;
    }
}

/* thermostats.rs:113-115 */
double orc_berendsen_thermostat_factor(double temperature, double instant, double tau) {
    return sqrt(1.0 + (temperature / instant - 1.0) / tau);
}

/* velocities.rs:17-18 */
double orc_rescale_thermostat_factor(double temperature, double instant) { return sqrt(temperature / instant); }

/* controls.rs:30-41 */
void orc_remove_translation(int64_t n, const double* mass, double* velocity) {
    double total_mass = 0.0;
    for (int64_t i = 0; i < n; i++) {
        total_mass += mass[i];
    }
    double com_velocity[3] = {0.0, 0.0, 0.0};
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            com_velocity[c] += velocity[3 * i + c] * mass[i] / total_mass;
        }
    }
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            velocity[3 * i + c] -= com_velocity[c];
        }
    }
}

/* ---- barostats and the remaining controls -------------------------------------------------------- */

static const double ORC_WATER_COMPRESSIBILITY = 7372.0; /* integrators.rs:172 */

static void mat3_mul(const double a[9], const double b[9], double out[9]) {
    double r[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    memcpy(out, r, sizeof(r));
}

/* Second half of both barostat steps: forces, accelerations, half kick (integrators.rs:246-254, 332-340). */
static void barostat_finish(orc_system* s, double* velocity, double* accelerations, double dt) {
    int64_t n = s->n;
    double* forces = (double*)malloc((size_t)(3 * n > 0 ? 3 * n : 1) * sizeof(double));
    orc_forces(s, forces);
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            accelerations[3 * i + c] = forces[3 * i + c] / s->mass[i];
        }
    }
    for (int64_t k = 0; k < 3 * n; k++) {
        velocity[k] += 0.5 * dt * accelerations[k];
    }
    free(forces);
}

/* BerendsenBarostat::integrate (integrators.rs:211-255).  Returns 1 when the reference would panic because the
 * scaled cell is smaller than twice `maximum_cutoff` (pass 0 when the system has no cut-off). */
int orc_berendsen_barostat_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt,
                                double pressure, double tau, double* eta, double maximum_cutoff) {
    int64_t n = s->n;
    for (int64_t k = 0; k < 3 * n; k++) {
        velocity[k] += 0.5 * dt * accelerations[k];
        position[k] *= *eta;
        position[k] += velocity[k] * dt;
    }
    s->position = position;
    s->velocity = velocity;
    /* cell.scale_mut(eta * eta * eta * Matrix3::one()): self.cell *= factor (cells.rs:203-207) */
    double factor[9] = {0};
    factor[0] = factor[4] = factor[8] = (*eta) * (*eta) * (*eta) * 1.0;
    mat3_mul(s->cell, factor, s->cell);
    if (maximum_cutoff > 0.0) {
        double lengths[3];
        orc_cell_lengths(s->cell, s->shape, lengths);
        for (int c = 0; c < 3; c++) {
            if (0.5 * lengths[c] <= maximum_cutoff) return 1;
        }
    }
    double eta3 = 1.0 - ORC_WATER_COMPRESSIBILITY / tau * (pressure - orc_pressure(s));
    *eta = cbrt(eta3);
    barostat_finish(s, velocity, accelerations, dt);
    return 0;
}

/* AnisoBerendsenBarostat::integrate (integrators.rs:295-341); eta and stress are row-major Matrix3. */
int orc_aniso_berendsen_barostat_step(orc_system* s, double* position, double* velocity, double* accelerations, double dt,
                                      const double stress[9], double tau, double eta[9], double maximum_cutoff) {
    int64_t n = s->n;
    for (int64_t i = 0; i < n; i++) {
        double* x = position + 3 * i;
        double* v = velocity + 3 * i;
        for (int c = 0; c < 3; c++) v[c] += 0.5 * dt * accelerations[3 * i + c];
        double scaled[3];
        for (int c = 0; c < 3; c++) scaled[c] = eta[3 * c] * x[0] + eta[3 * c + 1] * x[1] + eta[3 * c + 2] * x[2];
        for (int c = 0; c < 3; c++) x[c] = scaled[c] + v[c] * dt;
    }
    s->position = position;
    s->velocity = velocity;
    mat3_mul(s->cell, eta, s->cell); /* cell.scale_mut(eta) */
    if (maximum_cutoff > 0.0) {
        double lengths[3];
        orc_cell_lengths(s->cell, s->shape, lengths);
        for (int c = 0; c < 3; c++) {
            if (0.5 * lengths[c] <= maximum_cutoff) return 1;
        }
    }
    double factor = dt * ORC_WATER_COMPRESSIBILITY / tau;
    double current[9];
    orc_stress(s, current);
    for (int k = 0; k < 9; k++) {
        double one = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;
        eta[k] = one - factor * (stress[k] - current[k]);
    }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < i; j++) {
            eta[3 * i + j] = 0.5 * (eta[3 * i + j] + eta[3 * j + i]);
            eta[3 * j + i] = eta[3 * i + j];
        }
    }
    barostat_finish(s, velocity, accelerations, dt);
    return 0;
}

/* RemoveRotation::control (controls.rs:47-73) */
void orc_remove_rotation(int64_t n, const double* mass, const double* position, double* velocity) {
    double total_mass = 0.0, com[3] = {0.0, 0.0, 0.0};
    for (int64_t i = 0; i < n; i++) {
        total_mass += mass[i];
        for (int c = 0; c < 3; c++) com[c] += mass[i] * position[3 * i + c];
    }
    for (int c = 0; c < 3; c++) com[c] /= total_mass;
    double moment[3] = {0.0, 0.0, 0.0}, inertia[9] = {0};
    for (int64_t i = 0; i < n; i++) {
        double d[3], v[3];
        for (int c = 0; c < 3; c++) {
            d[c] = position[3 * i + c] - com[c];
            v[c] = velocity[3 * i + c];
        }
        const double cross[3] = {d[1] * v[2] - d[2] * v[1], d[2] * v[0] - d[0] * v[2], d[0] * v[1] - d[1] * v[0]};
        for (int c = 0; c < 3; c++) moment[c] += mass[i] * cross[c];
        for (int p = 0; p < 3; p++)
            for (int q = 0; q < 3; q++) inertia[3 * p + q] += -mass[i] * (d[p] * d[q]);
    }
    double trace = inertia[0] + inertia[4] + inertia[8];
    inertia[0] += trace;
    inertia[4] += trace;
    inertia[8] += trace;
    double inverse[9];
    orc_matrix_inverse(inertia, inverse);
    double angular[3];
    for (int c = 0; c < 3; c++) angular[c] = inverse[3 * c] * moment[0] + inverse[3 * c + 1] * moment[1] + inverse[3 * c + 2] * moment[2];
    for (int64_t i = 0; i < n; i++) {
        double d[3];
        for (int c = 0; c < 3; c++) d[c] = position[3 * i + c] - com[c];
        velocity[3 * i] -= d[1] * angular[2] - d[2] * angular[1];
        velocity[3 * i + 1] -= d[2] * angular[0] - d[0] * angular[2];
        velocity[3 * i + 2] -= d[0] * angular[1] - d[1] * angular[0];
    }
}

/* Rewrap::control (controls.rs:80-87) with Molecule::wrap (molecules.rs:287-296) */
void orc_rewrap(const orc_system* s, double* position) {
    for (int64_t m = 0; m < s->nmol; m++) {
        double total_mass = 0.0, com[3] = {0.0, 0.0, 0.0};
        for (int64_t i = s->mol_start[m]; i < s->mol_start[m + 1]; i++) {
            total_mass += s->mass[i];
            for (int c = 0; c < 3; c++) com[c] += s->mass[i] * position[3 * i + c];
        }
        double wrapped[3];
        for (int c = 0; c < 3; c++) {
            com[c] /= total_mass;
            wrapped[c] = com[c];
        }
        orc_wrap_vector(s->cell, s->shape, wrapped);
        for (int64_t i = s->mol_start[m]; i < s->mol_start[m + 1]; i++) {
            for (int c = 0; c < 3; c++) position[3 * i + c] += wrapped[c] - com[c];
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Monte Carlo energy cache (sys/cache.rs) and the GlobalCache of Ewald/Wolf  */
/* ------------------------------------------------------------------------ */

/* cells.rs:316-320 distance(u, v) = image(v - u).norm() */
static inline double distance_points(const geom_t* g, const double* u, const double* v) {
    double d[3] = {v[0] - u[0], v[1] - u[1], v[2] - u[2]};
    geom_image(g, d);
    return norm3(d);
}

/* energy.rs:32-45 EnergyEvaluator::pair */
static double evaluator_pair(const orc_system* s, int32_t path, double r, int64_t i, int64_t j) {
    const orc_pair* potential = pair_potential(s, i, j);
    if (potential == NULL) {
        return 0.0;
    }
    int32_t excluded;
    double scaling;
    orc_restriction_information(potential->restriction, potential->scale14, path, &excluded, &scaling);
    if (!excluded) {
        return scaling * orc_pair_energy(potential, r);
    }
    return 0.0;
}

/* The entry pairs_cache[(i, j)] as EnergyCache::init fills it (cache.rs:81-90): evaluated for i < j from
 * nearest_image(i, j).norm() and mirrored. */
static double cached_pair(const orc_system* s, const geom_t* g, int64_t a, int64_t b) {
    int64_t i = imin(a, b), j = imax(a, b);
    double d[3];
    nearest_image(s, g, i, j, d);
    return evaluator_pair(s, orc_bond_path(s, i, j), norm3(d), i, j);
}

/* cache.rs:145-172: pair part of EnergyCache::move_molecule_cost.  The reference reads the old pair
 * energies from pairs_cache; they are re-evaluated here exactly as EnergyCache::init stored them. */
double orc_move_molecule_pairs_cost(const orc_system* s, int64_t molecule, const double* new_positions) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    double pairs_delta = 0.0;
    int64_t first = s->mol_start[molecule], last = s->mol_start[molecule + 1];
    for (int64_t part_i = first; part_i < last; part_i++) {
        int64_t i = part_i - first;
        for (int64_t other = 0; other < s->nmol; other++) {
            if (other == molecule) continue;
            for (int64_t part_j = s->mol_start[other]; part_j < s->mol_start[other + 1]; part_j++) {
                double r = distance_points(&g, s->position + 3 * part_j, new_positions + 3 * i);
                int32_t path = orc_bond_path(s, part_i, part_j);
                double energy = evaluator_pair(s, path, r, part_i, part_j);
                pairs_delta += energy;
                pairs_delta -= cached_pair(s, &g, part_i, part_j);
            }
        }
    }
    return pairs_delta;
}

/* ewald.rs:572-613 real_space_move_molecule_cost */
double orc_ewald_real_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    double old_energy = 0.0, new_energy = 0.0;
    int64_t first = s->mol_start[molecule], last = s->mol_start[molecule + 1];
    for (int64_t part_i = first; part_i < last; part_i++) {
        int64_t i = part_i - first;
        double qi = s->charge[part_i];
        if (qi == 0.0) continue;
        for (int64_t other = 0; other < s->nmol; other++) {
            if (other == molecule) continue;
            for (int64_t part_j = s->mol_start[other]; part_j < s->mol_start[other + 1]; part_j++) {
                double qj = s->charge[part_j];
                if (qj == 0.0) continue;
                double old_r = distance(s, &g, part_i, part_j);
                double new_r = distance_points(&g, new_positions + 3 * i, s->position + 3 * part_j);
                int32_t path = orc_bond_path(s, part_i, part_j);
                int32_t excluded;
                double scaling;
                orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
                old_energy += ewald_real_energy_pair(s, excluded, qi * qj, old_r);
                new_energy += ewald_real_energy_pair(s, excluded, qi * qj, new_r);
            }
        }
    }
    return new_energy - old_energy;
}

/* ewald.rs:758-839 delta_rho_move_rigid_molecules + k_space_move_molecule_cost.  The reference reads the old
 * phases and rho(k) from the cache left by its last eik_dot_r; here that call is made on the spot, i.e. the cache
 * is taken to be current (it is in the reference's own test, ewald.rs:1326-1327).  delta_rho (2 per k) may be NULL. */
double orc_ewald_kspace_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions,
                                           int64_t capacity, double* delta_rho) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    ewald_factors_t f;
    ewald_factors_init(&f, s, &g);
    ewald_kspace_t ks;
    eik_dot_r(&ks, s, &g, &f);

    double old_energy = 0.0;
    for (int64_t ik = 0; ik < f.nk; ik++) {
        old_energy += f.energy[ik] * (ks.rho[ik].re * ks.rho[ik].re + ks.rho[ik].im * ks.rho[ik].im);
    }
    old_energy /= ORC_FOUR_PI_EPSILON_0;

    int64_t first = s->mol_start[molecule], size = s->mol_start[molecule + 1] - first;
    int64_t kmax = s->kmax;
    ewald_kspace_t fresh;
    fresh.kmax = kmax;
    fresh.n = size;
    fresh.eikr = (cplx*)malloc((size_t)((2 * kmax + 1) * 3 * (size > 0 ? size : 1)) * sizeof(cplx));
    fresh.rho = NULL;
    for (int spatial = 0; spatial < 3; spatial++) {
        double k_idx[3] = {0.0, 0.0, 0.0};
        k_idx[spatial] = 1.0;
        double k_vector[3];
        geom_k_vector(&g, k_idx, k_vector);
        for (int64_t i = 0; i < size; i++) {
            double phi = dot3(k_vector, new_positions + 3 * i);
            cplx one = {1.0, 0.0};
            cplx e1 = {1.0 * cos(phi), 1.0 * sin(phi)};
            cplx em1 = {e1.re, -e1.im};
            *eikr_at(&fresh, 0, spatial, i) = one;
            *eikr_at(&fresh, 1, spatial, i) = e1;
            *eikr_at(&fresh, -1, spatial, i) = em1;
        }
    }
    for (int spatial = 0; spatial < 3; spatial++) {
        for (int64_t k = 2; k < kmax + 1; k++) {
            for (int64_t i = 0; i < size; i++) {
                cplx v = cmul(*eikr_at(&fresh, k - 1, spatial, i), *eikr_at(&fresh, 1, spatial, i));
                cplx c = {v.re, -v.im};
                *eikr_at(&fresh, k, spatial, i) = v;
                *eikr_at(&fresh, -k, spatial, i) = c;
            }
        }
    }

    double new_energy = 0.0;
    for (int64_t ik = 0; ik < f.nk; ik++) {
        int64_t ikx = f.index[3 * ik], iky = f.index[3 * ik + 1], ikz = f.index[3 * ik + 2];
        cplx partial = {0.0, 0.0};
        for (int64_t i = 0; i < size; i++) {
            int64_t part_i = first + i;
            cplx old_phi = cmul(cmul(*eikr_at(&ks, ikx, 0, part_i), *eikr_at(&ks, iky, 1, part_i)), *eikr_at(&ks, ikz, 2, part_i));
            cplx new_phi = cmul(cmul(*eikr_at(&fresh, ikx, 0, i), *eikr_at(&fresh, iky, 1, i)), *eikr_at(&fresh, ikz, 2, i));
            partial.re += s->charge[part_i] * (new_phi.re - old_phi.re);
            partial.im += s->charge[part_i] * (new_phi.im - old_phi.im);
        }
        if (delta_rho != NULL && ik < capacity) {
            delta_rho[2 * ik] = partial.re;
            delta_rho[2 * ik + 1] = partial.im;
        }
        double re = ks.rho[ik].re + partial.re, im = ks.rho[ik].im + partial.im;
        new_energy += f.energy[ik] * (re * re + im * im);
    }
    new_energy /= ORC_FOUR_PI_EPSILON_0;

    free(fresh.eikr);
    kspace_free(&ks);
    ewald_factors_free(&f);
    return new_energy - old_energy;
}

/* wolf.rs:121-165 GlobalCache::move_molecule_cost */
double orc_wolf_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions) {
    geom_t g;
    geom_init(&g, s->cell, s->shape);
    double alpha, ec, fc;
    wolf_constants(s, &alpha, &ec, &fc);
    double old_energy = 0.0, new_energy = 0.0;
    int64_t first = s->mol_start[molecule], last = s->mol_start[molecule + 1];
    for (int64_t part_i = first; part_i < last; part_i++) {
        int64_t i = part_i - first;
        double qi = s->charge[part_i];
        if (qi == 0.0) continue;
        for (int64_t other = 0; other < s->nmol; other++) {
            if (other == molecule) continue;
            for (int64_t part_j = s->mol_start[other]; part_j < s->mol_start[other + 1]; part_j++) {
                double qj = s->charge[part_j];
                if (qj == 0.0) continue;
                int32_t path = orc_bond_path(s, part_i, part_j);
                int32_t excluded;
                double scaling;
                orc_restriction_information(s->coulomb_restriction, s->coulomb_scale14, path, &excluded, &scaling);
                if (excluded) continue;
                double old_r = distance(s, &g, part_i, part_j);
                double new_r = distance_points(&g, new_positions + 3 * i, s->position + 3 * part_j);
                /* wolf.rs:90-96 energy_pair */
                double old_e = old_r > s->rc ? 0.0 : (qi * qj) * (erfc(alpha * old_r) / old_r - ec) / ORC_FOUR_PI_EPSILON_0;
                double new_e = new_r > s->rc ? 0.0 : (qi * qj) * (erfc(alpha * new_r) / new_r - ec) / ORC_FOUR_PI_EPSILON_0;
                old_energy += scaling * old_e;
                new_energy += scaling * new_e;
            }
        }
    }
    return new_energy - old_energy;
}

/* EnergyCache::move_molecule_cost (cache.rs:145-174) by term: out = {pairs_delta, coulomb real space (or the whole
 * Wolf sum), coulomb k-space}; the cost is their sum (global potentials other than coulomb are out of scope). */
void orc_move_molecule_cost(const orc_system* s, int64_t molecule, const double* new_positions, double out[3]) {
    out[0] = orc_move_molecule_pairs_cost(s, molecule, new_positions);
    out[1] = 0.0;
    out[2] = 0.0;
    if (s->coulomb == ORC_COULOMB_EWALD) {
        /* SharedEwald::move_molecule_cost, ewald.rs:933-946: real + k_space, no self cost */
        out[1] = orc_ewald_real_move_molecule_cost(s, molecule, new_positions);
        out[2] = orc_ewald_kspace_move_molecule_cost(s, molecule, new_positions, 0, NULL);
    } else if (s->coulomb == ORC_COULOMB_WOLF) {
        out[1] = orc_wolf_move_molecule_cost(s, molecule, new_positions);
    }
}

/* EnergyCache::move_all_molecules_cost (cache.rs:230-283): `before` is the system the cache was initialised with,
 * `after` the system with every molecule moved rigidly (possibly in a new cell).  out = {pairs_delta over the
 * inter-molecular pairs, pairs_tail - cache.pairs_tail, new_coulomb - cache.coulomb}. */
void orc_move_all_molecules_cost(const orc_system* before, const orc_system* after, double out[3]) {
    geom_t g_before, g_after;
    geom_init(&g_before, before->cell, before->shape);
    geom_init(&g_after, after->cell, after->shape);
    double pairs_delta = 0.0;
    for (int64_t mi = 0; mi < after->nmol; mi++) {
        for (int64_t mj = mi + 1; mj < after->nmol; mj++) {
            for (int64_t part_i = after->mol_start[mi]; part_i < after->mol_start[mi + 1]; part_i++) {
                for (int64_t part_j = after->mol_start[mj]; part_j < after->mol_start[mj + 1]; part_j++) {
                    double r = distance(after, &g_after, part_i, part_j);
                    int32_t path = orc_bond_path(after, part_i, part_j);
                    double energy = evaluator_pair(after, path, r, part_i, part_j);
                    pairs_delta += energy;
                    pairs_delta -= cached_pair(before, &g_before, part_i, part_j);
                }
            }
        }
    }
    orc_energy_terms terms_before, terms_after;
    orc_energy_terms_compute(before, &terms_before);
    orc_energy_terms_compute(after, &terms_after);
    out[0] = pairs_delta;
    out[1] = terms_after.pairs_tail - terms_before.pairs_tail;
    double coulomb_before = terms_before.coulomb_real + terms_before.coulomb_self + terms_before.coulomb_kspace;
    double coulomb_after = terms_after.coulomb_real + terms_after.coulomb_self + terms_after.coulomb_kspace;
    out[2] = coulomb_after - coulomb_before;
}
