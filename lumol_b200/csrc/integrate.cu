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

// ------------------------------------------------------------------------------------------------
// MD driver
// ------------------------------------------------------------------------------------------------

int md_setup(Context* ctx) {
    const int64_t n3 = 3 * ctx->n;
    LUMOL_CUDA_CHECK(ctx, ctx->force.reserve((size_t)n3));
    if (ctx->integrator == LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET) {
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
    ctx->md_step++;
    return 0;
}

}  // namespace lumol
