// Device-resident molecular dynamics: integrators, kinetic sums, thermostat scaling, controls.
//   VelocityVerlet / Verlet / LeapFrog   lumol-sim/src/md/integrators.rs:39-169
//   KineticEnergy, Temperature           lumol-core/src/sys/compute.rs:134-171
//   Rescale / Berendsen / CSVR           lumol-sim/src/md/thermostats.rs:66-211
//   velocities::scale                    lumol-sim/src/velocities.rs:16-22
//   RemoveTranslation                    lumol-sim/src/md/controls.rs:30-41
//   MolecularDynamics::propagate         lumol-sim/src/md/molecular_dynamics.rs:66-76
//
// These kernels are HBM-bound element-wise passes over the packed n x 3 arrays (viewed as 3n doubles so
// every access is coalesced).  The arithmetic uses explicit __dmul_rn / __dadd_rn so that no FMA is
// contracted: each update rounds exactly like the reference's `v += 0.5 * dt * a; x += v * dt`.
#include "context.hpp"

namespace lumol {

constexpr int INT_THREADS = 256;

static inline int grid_for(int64_t count, int sm_count) {
    int64_t blocks = (count + INT_THREADS - 1) / INT_THREADS;
    int64_t cap = (int64_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Sharded runs: the drift kernel is also the position exchange.  Every new position is stored into the inbox of every
// other GPU (peer stores over NVLink, comm.cu); the last block to finish raises this rank's arrival flag on the peers.
__device__ __forceinline__ void peer_store(const PeerPush& push, int64_t k, double x) {
    for (int p = 0; p < push.nranks; p++) {
        if (p != push.rank) push.inbox[p][k] = x;
    }
}

__device__ __forceinline__ void peer_publish(const PeerPush& push) {
    if (push.nranks <= 1) return;
    __shared__ bool last;
    __threadfence_system();  // this thread's peer stores are visible before the block's ticket is taken
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(push.counter, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *push.counter = 0;
        if (threadIdx.x < push.nranks && (int)threadIdx.x != push.rank) {
            __threadfence_system();
            *(reinterpret_cast<volatile int*>(push.flags[threadIdx.x]) + push.rank) = push.epoch;
        }
    }
}

// VelocityVerlet first half (integrators.rs:47-53): a = f / m; v += (0.5 dt) a; x += v dt.
// The reference stores accelerations; a = force / mass is recomputed here from the resident forces,
// which gives the identical value and saves one n x 3 array (208 B/atom/step in total).
// MERGED: the second half of the previous step and the first half of this one in one pass (no thermostat or control
// runs between them): the same three roundings as the two kernels, 128 instead of 208 bytes per atom.
template <bool MERGED>
__device__ __forceinline__ double vv_advance(double half_dt, double dt, double f, double m, double& v, double x) {
    const double kick = __dmul_rn(half_dt, __ddiv_rn(f, m));
    v = __dadd_rn(v, kick);
    if (MERGED) v = __dadd_rn(v, kick);
    return __dadd_rn(x, __dmul_rn(v, dt));
}

// Two consecutive components per thread: 16-byte loads and stores, also towards the peers' inboxes.
template <bool MERGED>
__global__ void __launch_bounds__(INT_THREADS)
    vv_kick_drift_kernel(int64_t lo3, int64_t hi3, double half_dt, double dt, const double* __restrict__ force,
                         const double* __restrict__ mass, double* __restrict__ velocity, double* __restrict__ position,
                         PeerPush push) {
    const int64_t first = lo3 & ~(int64_t)1;
    for (int64_t k = first + 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); k < hi3; k += 2 * (int64_t)gridDim.x * blockDim.x) {
        if (k >= lo3 && k + 1 < hi3) {
            const double2 f = *reinterpret_cast<const double2*>(force + k);
            double2 v = *reinterpret_cast<const double2*>(velocity + k);
            double2 x = *reinterpret_cast<const double2*>(position + k);
            x.x = vv_advance<MERGED>(half_dt, dt, f.x, mass[k / 3], v.x, x.x);
            x.y = vv_advance<MERGED>(half_dt, dt, f.y, mass[(k + 1) / 3], v.y, x.y);
            *reinterpret_cast<double2*>(velocity + k) = v;
            *reinterpret_cast<double2*>(position + k) = x;
            for (int p = 0; p < push.nranks; p++) {
                if (p != push.rank) *reinterpret_cast<double2*>(push.inbox[p] + k) = x;
            }
        } else {
            for (int64_t e = max(k, lo3); e < min(k + 2, hi3); e++) {
                double v = velocity[e];
                const double x = vv_advance<MERGED>(half_dt, dt, force[e], mass[e / 3], v, position[e]);
                velocity[e] = v;
                position[e] = x;
                peer_store(push, e, x);
            }
        }
    }
    peer_publish(push);
}

// VelocityVerlet second half (integrators.rs:55-68): a = f / m; v += (0.5 dt) a.
__global__ void __launch_bounds__(INT_THREADS)
    vv_kick_kernel(int64_t lo3, int64_t hi3, double half_dt, const double* __restrict__ force,
                   const double* __restrict__ mass, double* __restrict__ velocity) {
    for (int64_t k = lo3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (int64_t)gridDim.x * blockDim.x) {
        const double a = __ddiv_rn(force[k], mass[k / 3]);
        velocity[k] = __dadd_rn(velocity[k], __dmul_rn(half_dt, a));
    }
}

// Verlet::setup (integrators.rs:92-101): prevpos = x - v dt
__global__ void __launch_bounds__(INT_THREADS)
    verlet_setup_kernel(int64_t n3, double dt, const double* __restrict__ position, const double* __restrict__ velocity,
                        double* __restrict__ prevpos) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += (int64_t)gridDim.x * blockDim.x) {
        prevpos[k] = __dadd_rn(position[k], -__dmul_rn(velocity[k], dt));
    }
}

// Verlet::integrate (integrators.rs:110-121)
__global__ void __launch_bounds__(INT_THREADS)
    verlet_kernel(int64_t lo3, int64_t hi3, double dt, const double* __restrict__ force, const double* __restrict__ mass,
                  double* __restrict__ position, double* __restrict__ velocity, double* __restrict__ prevpos) {
    const double dt2 = __dmul_rn(dt, dt);
    const double two_dt = __dmul_rn(2.0, dt);
    for (int64_t k = lo3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (int64_t)gridDim.x * blockDim.x) {
        const double x = position[k];
        const double p = prevpos[k];
        // 2.0 * position - prevpos + dt2 / mass * force
        const double xn = __dadd_rn(__dadd_rn(__dmul_rn(2.0, x), -p), __dmul_rn(__ddiv_rn(dt2, mass[k / 3]), force[k]));
        position[k] = xn;
        velocity[k] = __ddiv_rn(__dadd_rn(xn, -p), two_dt);
        prevpos[k] = x;
    }
}

// LeapFrog first loop (integrators.rs:154-158): x += v dt + 0.5 a dt^2
__global__ void __launch_bounds__(INT_THREADS)
    leapfrog_drift_kernel(int64_t lo3, int64_t hi3, double dt, const double* __restrict__ velocity,
                          const double* __restrict__ accel, double* __restrict__ position) {
    const double dt2 = __dmul_rn(dt, dt);
    for (int64_t k = lo3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (int64_t)gridDim.x * blockDim.x) {
        const double step = __dadd_rn(__dmul_rn(velocity[k], dt), __dmul_rn(__dmul_rn(0.5, accel[k]), dt2));
        position[k] = __dadd_rn(position[k], step);
    }
}

// LeapFrog second loop (integrators.rs:160-167)
__global__ void __launch_bounds__(INT_THREADS)
    leapfrog_kick_kernel(int64_t lo3, int64_t hi3, double dt, const double* __restrict__ force,
                         const double* __restrict__ mass, double* __restrict__ velocity, double* __restrict__ accel) {
    for (int64_t k = lo3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (int64_t)gridDim.x * blockDim.x) {
        const double a_new = __ddiv_rn(force[k], mass[k / 3]);
        const double mean = __dmul_rn(0.5, __dadd_rn(accel[k], a_new));
        velocity[k] = __dadd_rn(velocity[k], __dmul_rn(mean, dt));
        accel[k] = a_new;
    }
}

// ------------------------------------------------------------------------------------------------
// kinetic sums
// ------------------------------------------------------------------------------------------------

constexpr int KIN_NV = 11;  // K, sum m v(x)v [6], sum m v [3], sum m

__global__ void __launch_bounds__(INT_THREADS)
    kinetic_kernel(int lo, int hi, const double* __restrict__ velocity, const double* __restrict__ mass,
                   double* __restrict__ partials) {
    __shared__ double scratch[32 * KIN_NV];
    double acc[KIN_NV];
#pragma unroll
    for (int k = 0; k < KIN_NV; k++) acc[k] = 0.0;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        const double m = mass[i];
        const double vx = velocity[3 * i], vy = velocity[3 * i + 1], vz = velocity[3 * i + 2];
        // compute.rs:139: 0.5 * mass * velocity.norm2()
        acc[0] += 0.5 * m * (vx * vx + vy * vy + vz * vz);
        acc[1] += m * (vx * vx);
        acc[2] += m * (vx * vy);
        acc[3] += m * (vx * vz);
        acc[4] += m * (vy * vy);
        acc[5] += m * (vy * vz);
        acc[6] += m * (vz * vz);
        acc[7] += m * vx;
        acc[8] += m * vy;
        acc[9] += m * vz;
        acc[10] += m;
    }
    block_sum<KIN_NV>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < KIN_NV; k++) partials[(size_t)blockIdx.x * KIN_NV + k] = acc[k];
    }
}

int launch_kinetic(Context* ctx, bool tensor) {
    (void)tensor;
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    int blocks = grid_for(hi - lo, ctx->sm_count);
    if (blocks > 1024) blocks = 1024;
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)blocks * KIN_NV));
    {
        ScopedClock clock(ctx, &ctx->clk_integrate);
        kinetic_kernel<<<blocks, INT_THREADS, 0, ctx->stream>>>((int)lo, (int)hi, ctx->velocity.ptr, ctx->mass.ptr,
                                                                ctx->partials.ptr);
        ctx->launches++;
        ctx->clk_integrate.launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }
    // RES_KINETIC, RES_KINETIC_TENSOR[6], RES_MOMENTUM[4] are contiguous
    int status = launch_reduce(ctx, blocks, KIN_NV, RES_KINETIC);
    if (status != 0) return status;
    if (ctx->nranks > 1) {
        status = comm_allreduce(ctx, ctx->results.ptr + RES_KINETIC, KIN_NV);
    }
    return status;
}

// ------------------------------------------------------------------------------------------------
// thermostats and controls
// ------------------------------------------------------------------------------------------------

struct ThermostatArgs {
    int kind;
    double temperature;
    double parameter;
    double dof;
    const double* noise;  // CSVR: gauss, wiener for this step
};

// One thread turns the reduced kinetic energy into the velocity scaling factor.
__global__ void thermostat_factor_kernel(ThermostatArgs a, double* __restrict__ results) {
    const double kinetic = results[RES_KINETIC];
    double factor = 1.0;
    if (a.kind == LUMOL_CUDA_THERMOSTAT_RESCALE) {
        // thermostats.rs:67-72 + velocities.rs:16-22
        const double instant = 2.0 * kinetic / (a.dof * K_BOLTZMANN);
        if (fabs(instant - a.temperature) > a.parameter) {
            factor = sqrt(a.temperature / instant);
        }
    } else if (a.kind == LUMOL_CUDA_THERMOSTAT_BERENDSEN) {
        // thermostats.rs:113-115
        const double instant = 2.0 * kinetic / (a.dof * K_BOLTZMANN);
        factor = sqrt(1.0 + (a.temperature / instant - 1.0) / a.parameter);
    } else if (a.kind == LUMOL_CUDA_THERMOSTAT_CSVR) {
        // thermostats.rs:196-207
        const double target_per_dof = K_BOLTZMANN * a.temperature / 2.0;
        const double kinetic_factor = target_per_dof / kinetic;
        const double exp_1 = exp(-1.0 / a.parameter);
        const double exp_2 = (1.0 - exp_1) * kinetic_factor;
        const double gauss = a.noise[0], wiener = a.noise[1];
        const double scale = exp_1 + exp_2 * (gauss * gauss + wiener) + 2.0 * gauss * sqrt(exp_1 * exp_2);
        factor = sqrt(scale);
    }
    results[RES_SCALE_FACTOR] = factor;
}

__global__ void __launch_bounds__(INT_THREADS)
    scale_kernel(int64_t lo3, int64_t hi3, double factor, const double* __restrict__ device_factor,
                 double* __restrict__ velocity) {
    const double f = device_factor != nullptr ? *device_factor : factor;
    for (int64_t k = lo3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (int64_t)gridDim.x * blockDim.x) {
        velocity[k] = __dmul_rn(velocity[k], f);
    }
}

int launch_scale_velocities(Context* ctx, double factor, bool from_device) {
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    ScopedClock clock(ctx, &ctx->clk_integrate);
    scale_kernel<<<grid_for(3 * (hi - lo), ctx->sm_count), INT_THREADS, 0, ctx->stream>>>(
        3 * lo, 3 * hi, factor, from_device ? ctx->results.ptr + RES_SCALE_FACTOR : nullptr, ctx->velocity.ptr);
    ctx->launches++;
    ctx->clk_integrate.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// RemoveTranslation (controls.rs:30-41): v -= sum_i v_i m_i / M
__global__ void __launch_bounds__(INT_THREADS)
    remove_translation_kernel(int64_t lo3, int64_t hi3, const double* __restrict__ results, double* __restrict__ velocity) {
    const double total = results[RES_MOMENTUM + 3];
    const double c[3] = {results[RES_MOMENTUM] / total, results[RES_MOMENTUM + 1] / total,
                         results[RES_MOMENTUM + 2] / total};
    for (int64_t k = lo3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (int64_t)gridDim.x * blockDim.x) {
        velocity[k] -= c[k % 3];
    }
}

int launch_remove_translation(Context* ctx) {
    int status = launch_kinetic(ctx, false);
    if (status != 0) return status;
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    ScopedClock clock(ctx, &ctx->clk_integrate);
    remove_translation_kernel<<<grid_for(3 * (hi - lo), ctx->sm_count), INT_THREADS, 0, ctx->stream>>>(
        3 * lo, 3 * hi, ctx->results.ptr, ctx->velocity.ptr);
    ctx->launches++;
    ctx->clk_integrate.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// RemoveRotation (controls.rs:47-73): center of mass, then angular momentum L and inertia I about it, then
// v -= (x - com) ^ (I^-1 L).  Two reductions and one element-wise pass, nothing leaves the device.
constexpr int ROT_NV = 9;

__global__ void __launch_bounds__(INT_THREADS)
    center_kernel(int n, const double* __restrict__ position, const double* __restrict__ mass, double* __restrict__ partials) {
    __shared__ double scratch[32 * 4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double m = mass[i];
        acc[0] += m * position[3 * i];
        acc[1] += m * position[3 * i + 1];
        acc[2] += m * position[3 * i + 2];
        acc[3] += m;
    }
    block_sum<4>(acc, scratch);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; k++) partials[(size_t)blockIdx.x * 4 + k] = acc[k];
    }
}

__global__ void __launch_bounds__(INT_THREADS)
    rotation_kernel(int n, const double* __restrict__ position, const double* __restrict__ velocity,
                    const double* __restrict__ mass, const double* __restrict__ results, double* __restrict__ partials) {
    __shared__ double scratch[32 * ROT_NV];
    const double total = results[RES_CENTER + 3];
    const double cx = results[RES_CENTER] / total, cy = results[RES_CENTER + 1] / total, cz = results[RES_CENTER + 2] / total;
    double acc[ROT_NV];
#pragma unroll
    for (int k = 0; k < ROT_NV; k++) acc[k] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double m = mass[i];
        const double dx = position[3 * i] - cx, dy = position[3 * i + 1] - cy, dz = position[3 * i + 2] - cz;
        const double vx = velocity[3 * i], vy = velocity[3 * i + 1], vz = velocity[3 * i + 2];
        acc[0] += m * (dy * vz - dz * vy);
        acc[1] += m * (dz * vx - dx * vz);
        acc[2] += m * (dx * vy - dy * vx);
        acc[3] += -m * (dx * dx);
        acc[4] += -m * (dx * dy);
        acc[5] += -m * (dx * dz);
        acc[6] += -m * (dy * dy);
        acc[7] += -m * (dy * dz);
        acc[8] += -m * (dz * dz);
    }
    block_sum<ROT_NV>(acc, scratch);
    if (threadIdx.x == 0) {
        for (int k = 0; k < ROT_NV; k++) partials[(size_t)blockIdx.x * ROT_NV + k] = acc[k];
    }
}

__global__ void __launch_bounds__(INT_THREADS)
    remove_rotation_kernel(int n, const double* __restrict__ position, const double* __restrict__ results,
                           double* __restrict__ velocity) {
    const double total = results[RES_CENTER + 3];
    const double cx = results[RES_CENTER] / total, cy = results[RES_CENTER + 1] / total, cz = results[RES_CENTER + 2] / total;
    const double* r = results + RES_ROTATION;
    // inertia += trace on the diagonal (controls.rs:62-65), then the adjugate inverse of matrix.rs:212-227
    const double trace = r[3] + r[6] + r[8];
    const double m00 = r[3] + trace, m01 = r[4], m02 = r[5], m11 = r[6] + trace, m12 = r[7], m22 = r[8] + trace;
    const double det = m00 * (m11 * m22 - m12 * m12) - m01 * (m01 * m22 - m12 * m02) + m02 * (m01 * m12 - m11 * m02);
    const double i00 = (m11 * m22 - m12 * m12) / det, i01 = (m02 * m12 - m01 * m22) / det, i02 = (m01 * m12 - m02 * m11) / det;
    const double i11 = (m00 * m22 - m02 * m02) / det, i12 = (m01 * m02 - m00 * m12) / det, i22 = (m00 * m11 - m01 * m01) / det;
    const double wx = i00 * r[0] + i01 * r[1] + i02 * r[2];
    const double wy = i01 * r[0] + i11 * r[1] + i12 * r[2];
    const double wz = i02 * r[0] + i12 * r[1] + i22 * r[2];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double dx = position[3 * i] - cx, dy = position[3 * i + 1] - cy, dz = position[3 * i + 2] - cz;
        velocity[3 * i] -= dy * wz - dz * wy;
        velocity[3 * i + 1] -= dz * wx - dx * wz;
        velocity[3 * i + 2] -= dx * wy - dy * wx;
    }
}

int launch_remove_rotation(Context* ctx) {
    if (ctx->nranks > 1) return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "RemoveRotation is not available in sharded runs");
    const int n = (int)ctx->n;
    int blocks = grid_for(n, ctx->sm_count);
    if (blocks > 1024) blocks = 1024;
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)blocks * ROT_NV));
    ScopedClock clock(ctx, &ctx->clk_integrate);
    center_kernel<<<blocks, INT_THREADS, 0, ctx->stream>>>(n, ctx->position.ptr, ctx->mass.ptr, ctx->partials.ptr);
    int status = launch_reduce(ctx, blocks, 4, RES_CENTER);
    if (status != 0) return status;
    rotation_kernel<<<blocks, INT_THREADS, 0, ctx->stream>>>(n, ctx->position.ptr, ctx->velocity.ptr, ctx->mass.ptr,
                                                            ctx->results.ptr, ctx->partials.ptr);
    status = launch_reduce(ctx, blocks, ROT_NV, RES_ROTATION);
    if (status != 0) return status;
    remove_rotation_kernel<<<blocks, INT_THREADS, 0, ctx->stream>>>(n, ctx->position.ptr, ctx->results.ptr, ctx->velocity.ptr);
    ctx->launches += 3;
    ctx->clk_integrate.launches += 3;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// Rewrap (controls.rs:80-87): every molecule is translated so that its center of mass lies in the cell
// (Molecule::wrap, molecules.rs:287-296; UnitCell::wrap_vector, cells.rs:263-279).  One thread per molecule.
__global__ void __launch_bounds__(INT_THREADS)
    rewrap_kernel(int nmol, const int* __restrict__ mol_start, CellView cell, const double* __restrict__ mass,
                  double* __restrict__ position) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmol || cell.shape == LUMOL_CUDA_CELL_INFINITE) return;
    const int lo = mol_start[m], hi = mol_start[m + 1];
    double total = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
    // no FMA contraction: a center of mass that sits on a cell face must fall on the same side as in the reference
    for (int i = lo; i < hi; i++) {
        total = __dadd_rn(total, mass[i]);
        cx = __dadd_rn(cx, __dmul_rn(mass[i], position[3 * i]));
        cy = __dadd_rn(cy, __dmul_rn(mass[i], position[3 * i + 1]));
        cz = __dadd_rn(cz, __dmul_rn(mass[i], position[3 * i + 2]));
    }
    cx = __ddiv_rn(cx, total);
    cy = __ddiv_rn(cy, total);
    cz = __ddiv_rn(cz, total);
    double wx = cx, wy = cy, wz = cz;
    if (cell.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC) {
        wx = __dadd_rn(wx, -__dmul_rn(floor(__ddiv_rn(wx, cell.h[0])), cell.h[0]));
        wy = __dadd_rn(wy, -__dmul_rn(floor(__ddiv_rn(wy, cell.h[4])), cell.h[4]));
        wz = __dadd_rn(wz, -__dmul_rn(floor(__ddiv_rn(wz, cell.h[8])), cell.h[8]));
    } else {
        double fx = cell.inv[0] * wx + cell.inv[1] * wy + cell.inv[2] * wz;
        double fy = cell.inv[3] * wx + cell.inv[4] * wy + cell.inv[5] * wz;
        double fz = cell.inv[6] * wx + cell.inv[7] * wy + cell.inv[8] * wz;
        fx -= floor(fx);
        fy -= floor(fy);
        fz -= floor(fz);
        wx = cell.h[0] * fx + cell.h[1] * fy + cell.h[2] * fz;
        wy = cell.h[3] * fx + cell.h[4] * fy + cell.h[5] * fz;
        wz = cell.h[6] * fx + cell.h[7] * fy + cell.h[8] * fz;
    }
    const double dx = wx - cx, dy = wy - cy, dz = wz - cz;
    for (int i = lo; i < hi; i++) {
        position[3 * i] += dx;
        position[3 * i + 1] += dy;
        position[3 * i + 2] += dz;
    }
}

int launch_rewrap(Context* ctx) {
    if (ctx->nranks > 1) return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "Rewrap is not available in sharded runs");
    if (ctx->nmol == 0) return 0;  // without lumol_cuda_set_molecules every atom is its own molecule (nmol == n)
    ScopedClock clock(ctx, &ctx->clk_integrate);
    const int nmol = (int)ctx->nmol;
    rewrap_kernel<<<(nmol + INT_THREADS - 1) / INT_THREADS, INT_THREADS, 0, ctx->stream>>>(nmol, ctx->mol_start.ptr, ctx->cell, ctx->mass.ptr,
                                                                                          ctx->position.ptr);
    ctx->launches++;
    ctx->clk_integrate.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// First half of the Berendsen barostat steps (integrators.rs:215-223, 299-307): v += (0.5 dt) a; x = eta x; x += v dt.
__global__ void __launch_bounds__(INT_THREADS)
    barostat_drift_kernel(int n, double half_dt, double dt, const double* __restrict__ force, const double* __restrict__ mass,
                          double* __restrict__ velocity, double* __restrict__ position, const double e00, const double e01,
                          const double e02, const double e10, const double e11, const double e12, const double e20, const double e21,
                          const double e22, int isotropic) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double m = mass[i];
        double v[3], x[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            v[c] = __dadd_rn(velocity[3 * i + c], __dmul_rn(half_dt, __ddiv_rn(force[3 * i + c], m)));
            x[c] = position[3 * i + c];
            velocity[3 * i + c] = v[c];
        }
        double s[3];
        if (isotropic) {
            s[0] = __dmul_rn(x[0], e00);
            s[1] = __dmul_rn(x[1], e00);
            s[2] = __dmul_rn(x[2], e00);
        } else {
            s[0] = __dadd_rn(__dadd_rn(__dmul_rn(e00, x[0]), __dmul_rn(e01, x[1])), __dmul_rn(e02, x[2]));
            s[1] = __dadd_rn(__dadd_rn(__dmul_rn(e10, x[0]), __dmul_rn(e11, x[1])), __dmul_rn(e12, x[2]));
            s[2] = __dadd_rn(__dadd_rn(__dmul_rn(e20, x[0]), __dmul_rn(e21, x[1])), __dmul_rn(e22, x[2]));
        }
#pragma unroll
        for (int c = 0; c < 3; c++) position[3 * i + c] = __dadd_rn(s[c], __dmul_rn(v[c], dt));
    }
}

int launch_barostat_drift(Context* ctx, const double eta[9], bool isotropic) {
    const int n = (int)ctx->n;
    ScopedClock clock(ctx, &ctx->clk_integrate);
    barostat_drift_kernel<<<grid_for(n, ctx->sm_count), INT_THREADS, 0, ctx->stream>>>(
        n, 0.5 * ctx->dt, ctx->dt, ctx->force.ptr, ctx->mass.ptr, ctx->velocity.ptr, ctx->position.ptr, eta[0], eta[1], eta[2], eta[3],
        eta[4], eta[5], eta[6], eta[7], eta[8], isotropic ? 1 : 0);
    ctx->launches++;
    ctx->clk_integrate.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

int launch_second_kick(Context* ctx) {
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    ScopedClock clock(ctx, &ctx->clk_integrate);
    vv_kick_kernel<<<grid_for(3 * (hi - lo), ctx->sm_count), INT_THREADS, 0, ctx->stream>>>(3 * lo, 3 * hi, 0.5 * ctx->dt, ctx->force.ptr,
                                                                                         ctx->mass.ptr, ctx->velocity.ptr);
    ctx->launches++;
    ctx->clk_integrate.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// MD driver
// ------------------------------------------------------------------------------------------------

int md_setup(Context* ctx) {
    const int64_t n3 = 3 * ctx->n;
    LUMOL_CUDA_CHECK(ctx, ctx->force.reserve((size_t)n3));
    if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET || ctx->integrator == LUMOL_CUDA_INTEGRATOR_BERENDSEN_BAROSTAT ||
        ctx->integrator == LUMOL_CUDA_INTEGRATOR_ANISO_BERENDSEN_BAROSTAT) {
        // VelocityVerlet::setup zeroes the accelerations (integrators.rs:40-42): the first half kick
        // of the first step is a no-op whatever the forces are.
        LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->force.ptr, 0, (size_t)n3 * sizeof(double), ctx->stream));
    } else if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_VERLET) {
        LUMOL_CUDA_CHECK(ctx, ctx->aux.reserve((size_t)n3));
        verlet_setup_kernel<<<grid_for(n3, ctx->sm_count), INT_THREADS, 0, ctx->stream>>>(
            n3, ctx->dt, ctx->position.ptr, ctx->velocity.ptr, ctx->aux.ptr);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    } else {
        LUMOL_CUDA_CHECK(ctx, ctx->aux.reserve((size_t)n3));
        LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->aux.ptr, 0, (size_t)n3 * sizeof(double), ctx->stream));
    }
    ctx->md_step = 0;
    return 0;
}

// `first` / `last`: position of the step inside one lumol_cuda_md_run call.  Without thermostat and controls the
// second half kick of a step is merged with the first half of the next one (same arithmetic, one pass less).
int md_step(Context* ctx, bool first, bool last) {
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    const int64_t lo3 = 3 * lo, hi3 = 3 * hi;
    const int grid = grid_for(hi3 - lo3, ctx->sm_count);
    ComputeRequest req;
    req.forces = true;
    req.pairs = req.bonded = req.coulomb = true;
    int status = 0;

    if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET) {
        const bool merged = ctx->thermostat == LUMOL_CUDA_THERMOSTAT_NONE && ctx->controls == 0;
        // sharded: the drift stores the new positions straight into the other GPUs' inboxes (nranks stays 0 when
        // peer memory is not available: NCCL all-gather below)
        PeerPush push;
        if ((status = comm_peer_push_begin(ctx, &push)) != 0) return status;
        {
            ScopedClock clock(ctx, &ctx->clk_integrate);
            if (merged && !first) {
                vv_kick_drift_kernel<true><<<grid, INT_THREADS, 0, ctx->stream>>>(lo3, hi3, 0.5 * ctx->dt, ctx->dt, ctx->force.ptr,
                                                                                  ctx->mass.ptr, ctx->velocity.ptr,
                                                                                  ctx->position.ptr, push);
            } else {
                vv_kick_drift_kernel<false><<<grid, INT_THREADS, 0, ctx->stream>>>(lo3, hi3, 0.5 * ctx->dt, ctx->dt, ctx->force.ptr,
                                                                                   ctx->mass.ptr, ctx->velocity.ptr,
                                                                                   ctx->position.ptr, push);
            }
            ctx->launches++;
            ctx->clk_integrate.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        if (push.nranks > 1) {
            if ((status = comm_peer_gather(ctx, push)) != 0) return status;
        } else if (ctx->nranks > 1 && (status = comm_allgather_positions(ctx)) != 0) {
            return status;
        }
        if ((status = evaluate_forces_device(ctx, req)) != 0) return status;
        if (!merged || last) {
            ScopedClock clock(ctx, &ctx->clk_integrate);
            vv_kick_kernel<<<grid, INT_THREADS, 0, ctx->stream>>>(lo3, hi3, 0.5 * ctx->dt, ctx->force.ptr, ctx->mass.ptr,
                                                                  ctx->velocity.ptr);
            ctx->launches++;
            ctx->clk_integrate.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
    } else if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_VERLET) {
        if ((status = evaluate_forces_device(ctx, req)) != 0) return status;
        {
            ScopedClock clock(ctx, &ctx->clk_integrate);
            verlet_kernel<<<grid, INT_THREADS, 0, ctx->stream>>>(lo3, hi3, ctx->dt, ctx->force.ptr, ctx->mass.ptr,
                                                                 ctx->position.ptr, ctx->velocity.ptr, ctx->aux.ptr);
            ctx->launches++;
            ctx->clk_integrate.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        if (ctx->nranks > 1 && (status = comm_allgather_positions(ctx)) != 0) return status;
    } else if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_LEAP_FROG) {
        {
            ScopedClock clock(ctx, &ctx->clk_integrate);
            leapfrog_drift_kernel<<<grid, INT_THREADS, 0, ctx->stream>>>(lo3, hi3, ctx->dt, ctx->velocity.ptr,
                                                                         ctx->aux.ptr, ctx->position.ptr);
            ctx->launches++;
            ctx->clk_integrate.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        if (ctx->nranks > 1 && (status = comm_allgather_positions(ctx)) != 0) return status;
        if ((status = evaluate_forces_device(ctx, req)) != 0) return status;
        {
            ScopedClock clock(ctx, &ctx->clk_integrate);
            leapfrog_kick_kernel<<<grid, INT_THREADS, 0, ctx->stream>>>(lo3, hi3, ctx->dt, ctx->force.ptr, ctx->mass.ptr,
                                                                        ctx->velocity.ptr, ctx->aux.ptr);
            ctx->launches++;
            ctx->clk_integrate.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
    } else if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_BERENDSEN_BAROSTAT ||
               ctx->integrator == LUMOL_CUDA_INTEGRATOR_ANISO_BERENDSEN_BAROSTAT) {
        if ((status = barostat_step(ctx)) != 0) return status;
    } else {
        return ctx->fail(LUMOL_CUDA_ERROR_STATE, "lumol_cuda_md_setup was not called");
    }

    // thermostat (molecular_dynamics.rs:68-70)
    if (ctx->thermostat != LUMOL_CUDA_THERMOSTAT_NONE) {
        if ((status = launch_kinetic(ctx, false)) != 0) return status;
        ThermostatArgs t;
        t.kind = ctx->thermostat;
        t.temperature = ctx->thermostat_temperature;
        t.parameter = ctx->thermostat_parameter;
        t.dof = ctx->dof_mode == LUMOL_CUDA_DOF_MOLECULES ? 3.0 * (double)ctx->nmol
                                                          : (double)(3 * ctx->n - ctx->dof_frozen);
        t.noise = nullptr;
        if (ctx->thermostat == LUMOL_CUDA_THERMOSTAT_CSVR) {
            if (ctx->csvr_cursor >= ctx->csvr_count) {
                return ctx->fail(LUMOL_CUDA_ERROR_STATE, "CSVR thermostat ran out of host-provided noise");
            }
            t.noise = ctx->csvr_noise_dev.ptr + 2 * ctx->csvr_cursor;
            ctx->csvr_cursor++;
        }
        thermostat_factor_kernel<<<1, 1, 0, ctx->stream>>>(t, ctx->results.ptr);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        if ((status = launch_scale_velocities(ctx, 1.0, true)) != 0) return status;
    }
    // controls (molecular_dynamics.rs:72-74)
    if (ctx->controls & LUMOL_CUDA_CONTROL_REMOVE_TRANSLATION) {
        if ((status = launch_remove_translation(ctx)) != 0) return status;
    }
    if (ctx->controls & LUMOL_CUDA_CONTROL_REMOVE_ROTATION) {
        if ((status = launch_remove_rotation(ctx)) != 0) return status;
    }
    if (ctx->controls & LUMOL_CUDA_CONTROL_REWRAP) {
        if ((status = launch_rewrap(ctx)) != 0) return status;
    }
    ctx->md_step++;
    return 0;
}

}  // namespace lumol
