// Device-side arithmetic shared by every kernel: potentials, restrictions, periodic images and
// block reductions.  All FP64.  Formulas follow the reference's evaluation order where it matters
// for parity (citations are file:line relative to the reference checkout).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lumol_cuda.h"

namespace lumol {

constexpr double FOUR_PI_EPSILON_0 = 7.197589831304046;  // consts.rs:15
constexpr double K_BOLTZMANN = 8.31446284161522e-7;       // consts.rs:9
constexpr double FRAC_2_SQRT_PI = 1.12837916709551257389615890312154517;
constexpr double PI = 3.14159265358979323846264338327950288;

// BondDistances bits (connect.rs:142-165)
constexpr unsigned BOND_ONE = 1, BOND_TWO = 2, BOND_THREE = 4, BOND_FAR = 8;

// One PairInteraction as the kernels see it (from lumol_cuda_pair).
struct PairParams {
    int potential;
    int restriction;
    int table;
    int pad;
    double p[5];
    double cutoff;
    double shift;
    double scale14;
};

struct TableDesc {
    int size;
    int offset;  // into the concatenated energy / force arrays
    double delta;
};

struct CellView {
    double h[9];    // cell matrix, row-major
    double inv[9];  // its inverse (matrix.rs:212-227)
    int shape;
};

struct CoulombView {
    int kind;  // 0 none, 1 Ewald, 2 Wolf
    int restriction;
    double scale14;
    double rc;
    double alpha;
    double wolf_energy_constant;  // wolf.rs:75
    double wolf_force_constant;   // wolf.rs:76
};

// ------------------------------------------------------------------------------------------------
// periodic images
// ------------------------------------------------------------------------------------------------

// UnitCell::vector_image (cells.rs:284-300): round() is half-away-from-zero like f64::round, and the
// division is kept (not a multiplication by 1/L) so images flip at exactly the same separations.
__device__ __forceinline__ void vector_image(const CellView& c, double& x, double& y, double& z) {
    if (c.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC) {
        x -= round(x / c.h[0]) * c.h[0];
        y -= round(y / c.h[4]) * c.h[4];
        z -= round(z / c.h[8]) * c.h[8];
    } else if (c.shape == LUMOL_CUDA_CELL_TRICLINIC) {
        double fx = c.inv[0] * x + c.inv[1] * y + c.inv[2] * z;
        double fy = c.inv[3] * x + c.inv[4] * y + c.inv[5] * z;
        double fz = c.inv[6] * x + c.inv[7] * y + c.inv[8] * z;
        fx -= round(fx);
        fy -= round(fy);
        fz -= round(fz);
        x = c.h[0] * fx + c.h[1] * fy + c.h[2] * fz;
        y = c.h[3] * fx + c.h[4] * fy + c.h[5] * fz;
        z = c.h[6] * fx + c.h[7] * fy + c.h[8] * fz;
    }
}

// The same image and the squared norm with the reference's roundings: Rust never contracts a * b + c, so neither do
// these (explicit __dmul_rn / __dadd_rn).  Used where a pair may sit exactly on a cut-off (`r >= rc`, pairs.rs:186;
// `r > rc`, ewald.rs:390): the all-pairs kernel, the Monte Carlo cost kernels and the fix-up of the list kernels.
__device__ __forceinline__ double dot3_exact(double a0, double a1, double a2, double b0, double b1, double b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

__device__ __forceinline__ void vector_image_exact(const CellView& c, double& x, double& y, double& z) {
    if (c.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC) {
        x = __dadd_rn(x, -__dmul_rn(round(__ddiv_rn(x, c.h[0])), c.h[0]));
        y = __dadd_rn(y, -__dmul_rn(round(__ddiv_rn(y, c.h[4])), c.h[4]));
        z = __dadd_rn(z, -__dmul_rn(round(__ddiv_rn(z, c.h[8])), c.h[8]));
    } else if (c.shape == LUMOL_CUDA_CELL_TRICLINIC) {
        double fx = dot3_exact(c.inv[0], c.inv[1], c.inv[2], x, y, z);
        double fy = dot3_exact(c.inv[3], c.inv[4], c.inv[5], x, y, z);
        double fz = dot3_exact(c.inv[6], c.inv[7], c.inv[8], x, y, z);
        fx = __dadd_rn(fx, -round(fx));
        fy = __dadd_rn(fy, -round(fy));
        fz = __dadd_rn(fz, -round(fz));
        x = dot3_exact(c.h[0], c.h[1], c.h[2], fx, fy, fz);
        y = dot3_exact(c.h[3], c.h[4], c.h[5], fx, fy, fz);
        z = dot3_exact(c.h[6], c.h[7], c.h[8], fx, fy, fz);
    }
}

// ------------------------------------------------------------------------------------------------
// potentials (energy/functions.rs).  Returns energy and force(r) = -dV/dr as the reference defines
// it (Morse keeps the reference's formula, functions.rs:420-423).
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void potential_eval(int kind, const double* __restrict__ p, double r, double& energy,
                                               double& force) {
    switch (kind) {
    case LUMOL_CUDA_POTENTIAL_LJ: {  // functions.rs:80-88
        double s = p[0] / r;
        double s2 = s * s;
        double s6 = s2 * (s2 * s2);
        energy = 4.0 * p[1] * (s6 * s6 - s6);
        force = -24.0 * p[1] * (s6 - 2.0 * (s6 * s6)) / r;
        break;
    }
    case LUMOL_CUDA_POTENTIAL_HARMONIC: {  // functions.rs:136-143
        double dx = r - p[1];
        energy = 0.5 * p[0] * dx * dx;
        force = p[0] * (p[1] - r);
        break;
    }
    case LUMOL_CUDA_POTENTIAL_BUCKINGHAM: {  // functions.rs:287-299
        double r3 = r * r * r;
        double r6 = r3 * r3;
        double e = exp(-r / p[2]);
        energy = p[0] * e - p[1] / r6;
        force = p[0] / p[2] * e - 6.0 * p[1] / (r6 * r);
        break;
    }
    case LUMOL_CUDA_POTENTIAL_BMH: {  // functions.rs:354-366
        double r2 = r * r;
        double r6 = r2 * r2 * r2;
        double e = exp((p[3] - r) / p[4]);
        double r7 = r6 * r;
        energy = p[0] * e - p[1] / r6 + p[2] / (r6 * r2);
        force = p[0] / p[4] * e - 6.0 * p[1] / r7 + 8.0 * p[2] / (r7 * r2);
        break;
    }
    case LUMOL_CUDA_POTENTIAL_MORSE: {  // functions.rs:415-423
        double e = exp((p[1] - r) * p[0]);
        double rc = 1.0 - e;
        energy = p[2] * rc * rc;
        force = 2.0 * p[2] * (1.0 - e * e) * p[0];
        break;
    }
    case LUMOL_CUDA_POTENTIAL_GAUSSIAN: {  // functions.rs:475-481
        energy = -p[0] * exp(-p[1] * r * r);
        force = 2.0 * p[1] * r * energy;
        break;
    }
    case LUMOL_CUDA_POTENTIAL_MIE: {  // functions.rs:553-565
        double sr = p[0] / r;
        double rep = pow(sr, p[1]);
        double att = pow(sr, p[2]);
        energy = p[3] * (rep - att);
        force = p[3] * (p[1] * rep - p[2] * att) / r;
        break;
    }
    case LUMOL_CUDA_POTENTIAL_COSINE_HARMONIC: {  // functions.rs:199-206
        double sn, cs;
        sincos(r, &sn, &cs);
        double dr = cs - p[1];
        energy = 0.5 * p[0] * dr * dr;
        force = p[0] * dr * sn;
        break;
    }
    case LUMOL_CUDA_POTENTIAL_TORSION: {  // functions.rs:245-255
        double sn, cs;
        sincos(p[2] * r - p[1], &sn, &cs);
        energy = p[0] * (1.0 + cs);
        force = p[0] * p[2] * sn;
        break;
    }
    default:  // NullPotential (functions.rs:31-38) and absent entries
        energy = 0.0;
        force = 0.0;
        break;
    }
}

// TableComputation::compute_energy / compute_force (computations.rs:123-145)
__device__ __forceinline__ double table_lookup(const double* __restrict__ table, int size, double delta, double r) {
    double q = floor(r / delta);
    if (!(q < (double)(size - 1))) {
        return 0.0;
    }
    int bin = q > 0.0 ? (int)q : 0;  // `as usize` saturates negatives and NaN to 0
    double dx = r - (double)bin * delta;
    double t0 = table[bin];
    double slope = (table[bin + 1] - t0) / delta;
    return t0 + dx * slope;
}

// PairInteraction::{energy, force} (pairs.rs:185-218) for r < cutoff (the caller tests the cutoff).
__device__ __forceinline__ void pair_eval(const PairParams& pp, const TableDesc* __restrict__ tables,
                                          const double* __restrict__ table_energy,
                                          const double* __restrict__ table_force, double r, double& energy,
                                          double& force) {
    if (pp.potential == LUMOL_CUDA_POTENTIAL_TABLE) {
        TableDesc t = tables[pp.table];
        energy = table_lookup(table_energy + t.offset, t.size, t.delta, r);
        force = table_lookup(table_force + t.offset, t.size, t.delta, r);
    } else {
        potential_eval(pp.potential, pp.p, r, energy, force);
    }
    energy -= pp.shift;
}

// The same for pair tables that only hold Lennard-Jones, harmonic and null entries (water models, ionic crystals
// with LJ pairs): without the exp / pow / table branches the neighbour-list kernel is a third smaller, and that kernel
// waits on instruction fetch (profiles/r1y_list_force_kernel_spce98k_summary.csv).
__device__ __forceinline__ void pair_eval_simple(const PairParams& pp, double r, double rinv, double& energy, double& force) {
    if (pp.potential == LUMOL_CUDA_POTENTIAL_LJ) {  // functions.rs:80-88 with 1 / r supplied
        const double s = pp.p[0] * rinv;
        const double s2 = s * s;
        const double s6 = s2 * (s2 * s2);
        energy = 4.0 * pp.p[1] * (s6 * s6 - s6);
        force = -24.0 * pp.p[1] * (s6 - 2.0 * (s6 * s6)) * rinv;
    } else if (pp.potential == LUMOL_CUDA_POTENTIAL_HARMONIC) {  // functions.rs:136-143
        const double dx = r - pp.p[1];
        energy = 0.5 * pp.p[0] * dx * dx;
        force = pp.p[0] * (pp.p[1] - r);
    } else {
        energy = 0.0;
        force = 0.0;
    }
    energy -= pp.shift;
}

// ------------------------------------------------------------------------------------------------
// restrictions (restrictions.rs:85-114).  `bits` is the BondDistances byte of the pair, or 0 when
// the two atoms are in different molecules (BondPath::None).
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ bool restriction_excluded(int restriction, unsigned bits, double scale14, double& scaling) {
    const bool same = bits != 0;
    const bool one = (bits & BOND_ONE) != 0;
    const bool two = !one && (bits & BOND_TWO) != 0;
    const bool three = !one && !two && (bits & BOND_THREE) != 0;
    scaling = 1.0;
    switch (restriction) {
    case LUMOL_CUDA_RESTRICTION_INTER_MOLECULAR:
        return same;
    case LUMOL_CUDA_RESTRICTION_INTRA_MOLECULAR:
        return !same;
    case LUMOL_CUDA_RESTRICTION_EXCLUDE12:
        return one;
    case LUMOL_CUDA_RESTRICTION_EXCLUDE13:
        return one || two;
    case LUMOL_CUDA_RESTRICTION_EXCLUDE14:
        return one || two || three;
    case LUMOL_CUDA_RESTRICTION_SCALE14:
        if (three) {
            scaling = scale14;
        }
        return one || two;
    default:
        return false;
    }
}

// ------------------------------------------------------------------------------------------------
// Coulomb pair terms
// ------------------------------------------------------------------------------------------------

// Ewald::real_space_{energy,force}_pair (ewald.rs:387-428); returns force / r.
__device__ __forceinline__ void ewald_real_pair(const CoulombView& c, bool excluded, double qiqj, double r,
                                                double& energy, double& force_over_r) {
    double ar = c.alpha * r;
    double gauss = c.alpha * FRAC_2_SQRT_PI * exp(-ar * ar);
    if (excluded) {
        double e = erf(ar) / r;
        energy = -qiqj / FOUR_PI_EPSILON_0 * e;
        force_over_r = qiqj / (FOUR_PI_EPSILON_0 * r * r) * (gauss - e);
    } else {
        double e = erfc(ar) / r;
        energy = qiqj / FOUR_PI_EPSILON_0 * e;
        force_over_r = qiqj / (FOUR_PI_EPSILON_0 * r * r) * (gauss + e);
    }
}

// Wolf::energy_pair / force_pair (wolf.rs:90-117); returns force / r.
__device__ __forceinline__ void wolf_pair(const CoulombView& c, double qiqj, double r, double& energy,
                                          double& force_over_r) {
    double ar = c.alpha * r;
    double ec = erfc(ar);
    energy = qiqj * (ec / r - c.wolf_energy_constant) / FOUR_PI_EPSILON_0;
    double factor = ec / (r * r) + c.alpha * FRAC_2_SQRT_PI * exp(-ar * ar) / r;
    force_over_r = qiqj * (factor - c.wolf_force_constant) / (r * FOUR_PI_EPSILON_0);
}

// The same two terms with 1 / r supplied by the caller (one rsqrt of r^2 instead of a square root and three
// divisions per pair; constants folded).  Used by the neighbour-list kernel, where the pair rate matters; results
// differ from the divisions above in the last bit or two.
constexpr double INV_FOUR_PI_EPSILON_0 = 1.0 / FOUR_PI_EPSILON_0;

__device__ __forceinline__ void ewald_real_pair_rinv(const CoulombView& c, bool excluded, double qiqj, double r, double rinv,
                                                     double& energy, double& force_over_r) {
    const double ar = c.alpha * r;
    const double gauss = c.alpha * FRAC_2_SQRT_PI * exp(-ar * ar);
    const double q = qiqj * INV_FOUR_PI_EPSILON_0;
    // excluded pairs: -erf(ar) / r = (erfc(ar) - 1) / r (ewald.rs:395-399, 419-425)
    const double e = (excluded ? -erf(ar) : erfc(ar)) * rinv;
    energy = q * e;
    force_over_r = q * (rinv * rinv) * (gauss + e);
}

__device__ __forceinline__ void wolf_pair_rinv(const CoulombView& c, double qiqj, double r, double rinv, double& energy,
                                               double& force_over_r) {
    const double ar = c.alpha * r;
    const double ec = erfc(ar);
    const double q = qiqj * INV_FOUR_PI_EPSILON_0;
    energy = q * (ec * rinv - c.wolf_energy_constant);
    const double factor = ec * (rinv * rinv) + c.alpha * FRAC_2_SQRT_PI * exp(-ar * ar) * rinv;
    force_over_r = q * (factor - c.wolf_force_constant) * rinv;
}

// ------------------------------------------------------------------------------------------------
// reductions: warp shuffle, then one shared-memory pass, then per-block partials that a second
// kernel sums in a fixed order, so results do not depend on scheduling.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, offset);
    }
    return v;
}

// Sum NV values over the block; thread 0 ends up with the totals in v[].  `scratch` holds NV * 32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        v[k] = warp_sum(v[k]);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            scratch[k * 32 + warp] = v[k];
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = lane < nwarps ? scratch[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
    __syncthreads();
}

}  // namespace lumol
