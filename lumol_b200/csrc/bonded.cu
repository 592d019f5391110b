// Bonded terms: bonds, angles, dihedrals over explicit index lists.
//   Forces::compute            sys/compute.rs:62-97
//   EnergyEvaluator::{bonds,angles,dihedrals}   sys/energy.rs:90-141
//   AtomicVirial bond part     sys/compute.rs:231-239
// Geometry follows UnitCell::{angle_and_derivatives, dihedral_and_derivatives} (cells.rs:335-411).
// One thread per term; forces are scattered with FP64 atomics (the lists are small next to the pair
// work: a few terms per atom).
#include "context.hpp"

namespace lumol {

struct BondedArgs {
    const double* __restrict__ pos;
    const lumol_cuda_potential* __restrict__ potentials;
    const int* __restrict__ terms;  // (arity + 1) ints per term: atoms..., potential id
    int count;
    CellView cell;
    int do_forces;
    int do_sums;
    int o_lo, o_hi;  // atoms owned by this rank: forces go to those only, sums are taken by the owner of the first atom
    double* __restrict__ force;
    double* __restrict__ partials;
};

constexpr int BONDED_THREADS = 128;
constexpr int BONDED_NV = 7;  // energy, W[6]

__device__ __forceinline__ void add_force(const BondedArgs& a, int i, double x, double y, double z) {
    if (i < a.o_lo || i >= a.o_hi) return;
    double* force = a.force;
    atomicAdd(force + 3 * i, x);
    atomicAdd(force + 3 * i + 1, y);
    atomicAdd(force + 3 * i + 2, z);
}

__global__ void __launch_bounds__(BONDED_THREADS) bonds_kernel(BondedArgs a) {
    __shared__ double scratch[32 * BONDED_NV];
    double acc[BONDED_NV];
#pragma unroll
    for (int k = 0; k < BONDED_NV; k++) acc[k] = 0.0;

    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.count) {
        const int i = a.terms[3 * t], j = a.terms[3 * t + 1], pid = a.terms[3 * t + 2];
        if (pid >= 0) {
            double dx = a.pos[3 * i] - a.pos[3 * j];
            double dy = a.pos[3 * i + 1] - a.pos[3 * j + 1];
            double dz = a.pos[3 * i + 2] - a.pos[3 * j + 2];
            vector_image(a.cell, dx, dy, dz);
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            const lumol_cuda_potential pot = a.potentials[pid];
            double e, f;
            potential_eval(pot.potential, pot.p, r, e, f);
            const double fr = f / r;
            if (a.do_forces) {
                add_force(a, i, fr * dx, fr * dy, fr * dz);
                add_force(a, j, -fr * dx, -fr * dy, -fr * dz);
            }
            if (i >= a.o_lo && i < a.o_hi) {
                acc[0] = e;
                acc[1] = fr * dx * dx;
                acc[2] = fr * dx * dy;
                acc[3] = fr * dx * dz;
                acc[4] = fr * dy * dy;
                acc[5] = fr * dy * dz;
                acc[6] = fr * dz * dz;
            }
        }
    }
    if (a.do_sums) {
        block_sum<BONDED_NV>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < BONDED_NV; k++) a.partials[(size_t)blockIdx.x * BONDED_NV + k] = acc[k];
        }
    }
}

__global__ void __launch_bounds__(BONDED_THREADS) angles_kernel(BondedArgs a) {
    __shared__ double scratch[32];
    double acc[1] = {0.0};
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.count) {
        const int i = a.terms[4 * t], j = a.terms[4 * t + 1], k = a.terms[4 * t + 2], pid = a.terms[4 * t + 3];
        if (pid >= 0) {
            // cells.rs:335-359
            double ax = a.pos[3 * i] - a.pos[3 * j], ay = a.pos[3 * i + 1] - a.pos[3 * j + 1],
                   az = a.pos[3 * i + 2] - a.pos[3 * j + 2];
            vector_image(a.cell, ax, ay, az);
            double bx = a.pos[3 * k] - a.pos[3 * j], by = a.pos[3 * k + 1] - a.pos[3 * j + 1],
                   bz = a.pos[3 * k + 2] - a.pos[3 * j + 2];
            vector_image(a.cell, bx, by, bz);
            const double an = sqrt(ax * ax + ay * ay + az * az);
            const double bn = sqrt(bx * bx + by * by + bz * bz);
            const double ux = ax / an, uy = ay / an, uz = az / an;
            const double vx = bx / bn, vy = by / bn, vz = bz / bn;
            const double c = ux * vx + uy * vy + uz * vz;
            const double sin_inv = 1.0 / sqrt(1.0 - c * c);
            const double theta = acos(c);
            const lumol_cuda_potential pot = a.potentials[pid];
            double e, f;
            potential_eval(pot.potential, pot.p, theta, e, f);
            if (a.do_forces) {
                const double d1x = sin_inv * (c * ux - vx) / an, d1y = sin_inv * (c * uy - vy) / an,
                             d1z = sin_inv * (c * uz - vz) / an;
                const double d3x = sin_inv * (c * vx - ux) / bn, d3y = sin_inv * (c * vy - uy) / bn,
                             d3z = sin_inv * (c * vz - uz) / bn;
                add_force(a, i, f * d1x, f * d1y, f * d1z);
                add_force(a, j, -f * (d1x + d3x), -f * (d1y + d3y), -f * (d1z + d3z));
                add_force(a, k, f * d3x, f * d3y, f * d3z);
            }
            if (i >= a.o_lo && i < a.o_hi) acc[0] = e;
        }
    }
    if (a.do_sums) {
        block_sum<1>(acc, scratch);
        if (threadIdx.x == 0) a.partials[blockIdx.x] = acc[0];
    }
}

__global__ void __launch_bounds__(BONDED_THREADS) dihedrals_kernel(BondedArgs a) {
    __shared__ double scratch[32];
    double acc[1] = {0.0};
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.count) {
        const int i = a.terms[5 * t], j = a.terms[5 * t + 1], k = a.terms[5 * t + 2], m = a.terms[5 * t + 3],
                  pid = a.terms[5 * t + 4];
        if (pid >= 0) {
            // cells.rs:379-411
            double r12x = a.pos[3 * j] - a.pos[3 * i], r12y = a.pos[3 * j + 1] - a.pos[3 * i + 1],
                   r12z = a.pos[3 * j + 2] - a.pos[3 * i + 2];
            vector_image(a.cell, r12x, r12y, r12z);
            double r23x = a.pos[3 * k] - a.pos[3 * j], r23y = a.pos[3 * k + 1] - a.pos[3 * j + 1],
                   r23z = a.pos[3 * k + 2] - a.pos[3 * j + 2];
            vector_image(a.cell, r23x, r23y, r23z);
            double r34x = a.pos[3 * m] - a.pos[3 * k], r34y = a.pos[3 * m + 1] - a.pos[3 * k + 1],
                   r34z = a.pos[3 * m + 2] - a.pos[3 * k + 2];
            vector_image(a.cell, r34x, r34y, r34z);

            const double ux = r12y * r23z - r12z * r23y, uy = r12z * r23x - r12x * r23z,
                         uz = r12x * r23y - r12y * r23x;
            const double vx = r23y * r34z - r23z * r34y, vy = r23z * r34x - r23x * r34z,
                         vz = r23x * r34y - r23y * r34x;
            const double u2 = ux * ux + uy * uy + uz * uz;
            const double v2 = vx * vx + vy * vy + vz * vz;
            const double r23_2 = r23x * r23x + r23y * r23y + r23z * r23z;
            const double r23n = sqrt(r23_2);
            const double phi = atan2(r23n * (vx * r12x + vy * r12y + vz * r12z), ux * vx + uy * vy + uz * vz);
            const lumol_cuda_potential pot = a.potentials[pid];
            double e, f;
            potential_eval(pot.potential, pot.p, phi, e, f);
            if (a.do_forces) {
                const double f1 = -r23n / u2, f4 = r23n / v2;
                const double d1x = f1 * ux, d1y = f1 * uy, d1z = f1 * uz;
                const double d4x = f4 * vx, d4y = f4 * vy, d4z = f4 * vz;
                const double r23_r34 = r23x * r34x + r23y * r34y + r23z * r34z;
                const double r12_r23 = r12x * r23x + r12y * r23y + r12z * r23z;
                const double c21 = -r12_r23 / r23_2 - 1.0, c24 = r23_r34 / r23_2;
                const double c34 = -r23_r34 / r23_2 - 1.0, c31 = r12_r23 / r23_2;
                add_force(a, i, f * d1x, f * d1y, f * d1z);
                add_force(a, j, f * (c21 * d1x + c24 * d4x), f * (c21 * d1y + c24 * d4y),
                          f * (c21 * d1z + c24 * d4z));
                add_force(a, k, f * (c34 * d4x + c31 * d1x), f * (c34 * d4y + c31 * d1y),
                          f * (c34 * d4z + c31 * d1z));
                add_force(a, m, f * d4x, f * d4y, f * d4z);
            }
            if (i >= a.o_lo && i < a.o_hi) acc[0] = e;
        }
    }
    if (a.do_sums) {
        block_sum<1>(acc, scratch);
        if (threadIdx.x == 0) a.partials[blockIdx.x] = acc[0];
    }
}

__global__ void zero_results_kernel(double* results, int first, int count) {
    const int k = threadIdx.x;
    if (k < count) results[first + k] = 0.0;
}

int launch_bonded(Context* ctx, const ComputeRequest& req) {
    const bool sums = req.energy || req.virial;
    // energies/virial default to zero when a list is empty (nobody reads them after a forces-only evaluation)
    if (sums) {
        zero_results_kernel<<<1, 32, 0, ctx->stream>>>(ctx->results.ptr, RES_E_BONDS, 9);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }
    if (!req.forces && !sums) return 0;

    BondedArgs a;
    a.pos = ctx->position.ptr;
    a.potentials = ctx->bonded.ptr;
    a.cell = ctx->cell;
    a.do_forces = req.forces;
    a.do_sums = sums;
    a.force = ctx->force.ptr;
    int64_t o_lo, o_hi;
    ctx->owned_range(ctx->n, o_lo, o_hi);
    a.o_lo = (int)o_lo;
    a.o_hi = (int)o_hi;

    auto run = [&](int64_t count, const int* terms, int which) -> int {
        if (count == 0) return 0;
        a.terms = terms;
        a.count = (int)count;
        const int blocks = (int)((count + BONDED_THREADS - 1) / BONDED_THREADS);
        const int nv = which == 0 ? BONDED_NV : 1;
        LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)blocks * nv));
        a.partials = ctx->partials.ptr;
        if (which == 0) {
            bonds_kernel<<<blocks, BONDED_THREADS, 0, ctx->stream>>>(a);
        } else if (which == 1) {
            angles_kernel<<<blocks, BONDED_THREADS, 0, ctx->stream>>>(a);
        } else {
            dihedrals_kernel<<<blocks, BONDED_THREADS, 0, ctx->stream>>>(a);
        }
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        if (sums) {
            if (which == 0) {
                // RES_E_BONDS and RES_W_BONDS[6] are contiguous
                int status = launch_reduce(ctx, blocks, BONDED_NV, RES_E_BONDS);
                if (status != 0) return status;
            } else {
                int status = launch_reduce(ctx, blocks, 1, which == 1 ? RES_E_ANGLES : RES_E_DIHEDRALS);
                if (status != 0) return status;
            }
        }
        return 0;
    };

    int status = run(ctx->nbonds, ctx->bonds.ptr, 0);
    if (status != 0) return status;
    status = run(ctx->nangles, ctx->angles.ptr, 1);
    if (status != 0) return status;
    return run(ctx->ndihedrals, ctx->dihedrals.ptr, 2);
}

}  // namespace lumol
