// Tiled all-pairs minimum-image kernel: the production path whenever a cell edge holds fewer than
// three cut-off lengths (every system in the reference's own tests and benches) and
// for infinite cells.  It is the direct analogue of the reference's N^2 loops
//   Forces::compute          sys/compute.rs:37-55
//   EnergyEvaluator::pairs   sys/energy.rs:47-59
//   AtomicVirial::compute    sys/compute.rs:202-216
//   MolecularVirial::compute sys/compute.rs:286-311
//   Ewald real space         energy/global/ewald.rs:430-545
//   Wolf                     energy/global/wolf.rs:177-325
// with one warp per atom i, the 32 lanes striding over j.  Every unordered pair is visited from both
// sides (no atomics, no Newton-3 scatter); energies and virials are accumulated from the j > i visit
// only, so each pair contributes exactly once, with d = x_i - x_j as in the reference.
#include "context.hpp"

namespace lumol {

struct AllPairsArgs {
    int n;
    int i_lo, i_hi;
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    const unsigned* __restrict__ kind;
    const int* __restrict__ mol_first;
    const int* __restrict__ bd_row;
    const int* __restrict__ mol_of;
    const unsigned char* __restrict__ bond_dist;
    const double* __restrict__ mol_com;
    int nkinds;
    const PairParams* __restrict__ pairs;
    const TableDesc* __restrict__ tables;
    const double* __restrict__ table_energy;
    const double* __restrict__ table_force;
    CellView cell;
    CoulombView coulomb;
    int do_pairs;
    int do_coulomb;
    int write_forces;  // energy-only queries must not clobber the forces the integrator holds
    double* __restrict__ force;
    double* __restrict__ partials;
};

constexpr int MODE_FORCES = 0;     // forces only
constexpr int MODE_FULL = 1;       // forces + energies + atomic virial
constexpr int MODE_MOLECULAR = 2;  // molecular virial only

constexpr int ALLPAIRS_THREADS = 256;
constexpr int ALLPAIRS_WARPS = ALLPAIRS_THREADS / 32;
constexpr int ALLPAIRS_NV = 16;  // e_pairs, e_coulomb, W_pairs[6], W_coulomb[6], pair count, coulomb pair count

template <int MODE>
__global__ void __launch_bounds__(ALLPAIRS_THREADS) allpairs_kernel(AllPairsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairParams* sp = reinterpret_cast<PairParams*>(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw + sizeof(PairParams) * a.nkinds * a.nkinds);

    {
        // stage the (kind, kind) table in shared memory
        const int words = (int)(sizeof(PairParams) / sizeof(double)) * a.nkinds * a.nkinds;
        const double* src = reinterpret_cast<const double*>(a.pairs);
        double* dst = reinterpret_cast<double*>(sp);
        for (int w = threadIdx.x; w < words; w += blockDim.x) {
            dst[w] = src[w];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int i = a.i_lo + blockIdx.x * ALLPAIRS_WARPS + warp;

    double acc[ALLPAIRS_NV];
#pragma unroll
    for (int k = 0; k < ALLPAIRS_NV; k++) acc[k] = 0.0;

    if (i < a.i_hi) {
        const double xi = a.pos[3 * i], yi = a.pos[3 * i + 1], zi = a.pos[3 * i + 2];
        const double qi = a.charge[i];
        const unsigned ki = a.kind[i];
        const int mfi = a.mol_first[i];
        const int rowi = a.bd_row[i];
        const int moli = a.mol_of[i];
        double cx = 0.0, cy = 0.0, cz = 0.0;
        if (MODE == MODE_MOLECULAR) {
            cx = a.mol_com[3 * moli];
            cy = a.mol_com[3 * moli + 1];
            cz = a.mol_com[3 * moli + 2];
        }
        double fx = 0.0, fy = 0.0, fz = 0.0;

        for (int j = lane; j < a.n; j += 32) {
            if (j == i) continue;
            const int mfj = a.mol_first[j];
            const bool same_molecule = mfj == mfi;
            if (MODE == MODE_MOLECULAR && same_molecule) continue;
            // configuration.rs:399-403 nearest_image(i, j) = image(r_i - r_j)
            double dx = xi - a.pos[3 * j];
            double dy = yi - a.pos[3 * j + 1];
            double dz = zi - a.pos[3 * j + 2];
            // image and norm with the reference's roundings: a pair on the cut-off falls on the reference's side
            vector_image_exact(a.cell, dx, dy, dz);
            const double r2 = dot3_exact(dx, dy, dz, dx, dy, dz);
            const double r = sqrt(r2);
            const unsigned bits = same_molecule ? a.bond_dist[rowi + (j - mfj)] : 0u;
            // which side of the pair accumulates the scalar sums
            const bool count = MODE == MODE_MOLECULAR ? (moli < a.mol_of[j]) : (j > i);

            double proj = 1.0;
            if (MODE == MODE_MOLECULAR) {
                // compute.rs:288-303: r_ij = image(com_i - com_j); factor (r_ab . r_ij) / |r_ab|^2
                const int molj = a.mol_of[j];
                double rx = cx - a.mol_com[3 * molj];
                double ry = cy - a.mol_com[3 * molj + 1];
                double rz = cz - a.mol_com[3 * molj + 2];
                vector_image(a.cell, rx, ry, rz);
                proj = (dx * rx + dy * ry + dz * rz) / r2;
            }

            if (a.do_pairs) {
                const PairParams& pp = sp[ki * a.nkinds + a.kind[j]];
                if (pp.potential > LUMOL_CUDA_POTENTIAL_NULL && r < pp.cutoff) {
                    double scaling;
                    if (!restriction_excluded(pp.restriction, bits, pp.scale14, scaling)) {
                        double e, f;
                        pair_eval(pp, a.tables, a.table_energy, a.table_force, r, e, f);
                        const double fr = scaling * f / r;
                        if (MODE != MODE_MOLECULAR) {
                            fx += fr * dx;
                            fy += fr * dy;
                            fz += fr * dz;
                        }
                        if (MODE != MODE_FORCES && count) {
                            const double w = fr * proj;
                            acc[0] += scaling * e;
                            acc[14] += 1.0;
                            acc[2] += w * dx * dx;
                            acc[3] += w * dx * dy;
                            acc[4] += w * dx * dz;
                            acc[5] += w * dy * dy;
                            acc[6] += w * dy * dz;
                            acc[7] += w * dz * dz;
                        }
                    }
                }
            }

            if (a.do_coulomb && a.coulomb.kind != 0 && r <= a.coulomb.rc) {
                const double qj = a.charge[j];
                if (qi != 0.0 && qj != 0.0) {
                    double scaling;
                    const bool excluded =
                        restriction_excluded(a.coulomb.restriction, bits, a.coulomb.scale14, scaling);
                    double e = 0.0, fr = 0.0;
                    bool active = true;
                    if (a.coulomb.kind == 1) {
                        ewald_real_pair(a.coulomb, excluded, qi * qj, r, e, fr);
                    } else if (!excluded) {
                        wolf_pair(a.coulomb, qi * qj, r, e, fr);
                        e *= scaling;
                        fr *= scaling;
                    } else {
                        active = false;
                    }
                    if (active) {
                        if (MODE != MODE_MOLECULAR) {
                            fx += fr * dx;
                            fy += fr * dy;
                            fz += fr * dz;
                        }
                        if (MODE != MODE_FORCES && count) {
                            const double w = fr * proj;
                            acc[1] += e;
                            acc[15] += 1.0;
                            acc[8] += w * dx * dx;
                            acc[9] += w * dx * dy;
                            acc[10] += w * dx * dz;
                            acc[11] += w * dy * dy;
                            acc[12] += w * dy * dz;
                            acc[13] += w * dz * dz;
                        }
                    }
                }
            }
        }

        if (MODE != MODE_MOLECULAR && a.write_forces) {
            fx = warp_sum(fx);
            fy = warp_sum(fy);
            fz = warp_sum(fz);
            if (lane == 0) {
                a.force[3 * i] = fx;
                a.force[3 * i + 1] = fy;
                a.force[3 * i + 2] = fz;
            }
        }
    }

    if (MODE != MODE_FORCES) {
        block_sum<ALLPAIRS_NV>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < ALLPAIRS_NV; k++) {
                a.partials[(size_t)blockIdx.x * ALLPAIRS_NV + k] = acc[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// second-stage reduction: one block adds the per-block partials in a fixed order
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partials, int nblocks,
                                                              int nvalues, double* __restrict__ out, int out_stride) {
    // block b sums the slice [b * per, (b + 1) * per) of the rows; with one block that is everything
    __shared__ double scratch[32];
    const int per = (nblocks + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per;
    const int hi = min(nblocks, lo + per);
    for (int k = 0; k < nvalues; k++) {
        double v = 0.0;
        for (int b = lo + threadIdx.x; b < hi; b += blockDim.x) {
            v += partials[(size_t)b * nvalues + k];
        }
        double w[1] = {v};
        block_sum<1>(w, scratch);
        if (threadIdx.x == 0) {
            out[(size_t)blockIdx.x * out_stride + k] = w[0];
        }
    }
}

constexpr int REDUCE_FANOUT = 128;

int launch_reduce(Context* ctx, int nblocks, int nvalues, int first_slot) {
    if (nblocks > 4 * REDUCE_FANOUT) {
        // two stages: REDUCE_FANOUT blocks fold the rows into the tail of the partials buffer
        LUMOL_CUDA_CHECK(ctx, ctx->reduce_scratch.reserve((size_t)REDUCE_FANOUT * nvalues));
        reduce_partials_kernel<<<REDUCE_FANOUT, 256, 0, ctx->stream>>>(ctx->partials.ptr, nblocks, nvalues,
                                                                       ctx->reduce_scratch.ptr, nvalues);
        reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(ctx->reduce_scratch.ptr, REDUCE_FANOUT, nvalues,
                                                           ctx->results.ptr + first_slot, nvalues);
        ctx->launches += 2;
    } else {
        reduce_partials_kernel<<<1, 256, 0, ctx->stream>>>(ctx->partials.ptr, nblocks, nvalues,
                                                           ctx->results.ptr + first_slot, nvalues);
        ctx->launches++;
    }
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// self terms and molecule centres
// ------------------------------------------------------------------------------------------------

// Ewald::self_energy needs sum q^2 (ewald.rs:619-626); Wolf::energy subtracts energy_self(q_i) for
// every atom with q_i != 0 (wolf.rs:99-101, 204), which is also proportional to sum q^2.
__global__ void __launch_bounds__(256) charge2_kernel(const double* __restrict__ charge, int lo, int hi,
                                                      double* __restrict__ partials) {
    __shared__ double scratch[32];
    double v = 0.0;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        const double q = charge[i];
        v += q * q;
    }
    double w[1] = {v};
    block_sum<1>(w, scratch);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = w[0];
    }
}

int launch_coulomb_self(Context* ctx) {
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    int blocks = (int)((hi - lo + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    if (blocks < 1) blocks = 1;
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve(blocks));
    charge2_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->charge.ptr, (int)lo, (int)hi, ctx->partials.ptr);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return launch_reduce(ctx, blocks, 1, RES_CHARGE2);
}

// Molecule::center_of_mass (molecules.rs:256-264): serial sum over the molecule's atoms, unwrapped.
__global__ void molecule_com_kernel(int nmol, const int* __restrict__ mol_start, const double* __restrict__ pos,
                                    const double* __restrict__ mass, double* __restrict__ com) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    double total = 0.0, x = 0.0, y = 0.0, z = 0.0;
    for (int i = mol_start[m]; i < mol_start[m + 1]; i++) {
        const double w = mass[i];
        total += w;
        x += w * pos[3 * i];
        y += w * pos[3 * i + 1];
        z += w * pos[3 * i + 2];
    }
    com[3 * m] = x / total;
    com[3 * m + 1] = y / total;
    com[3 * m + 2] = z / total;
}

int launch_molecule_com(Context* ctx) {
    LUMOL_CUDA_CHECK(ctx, ctx->mol_com.reserve((size_t)ctx->nmol * 3));
    int blocks = (int)((ctx->nmol + 127) / 128);
    molecule_com_kernel<<<blocks, 128, 0, ctx->stream>>>((int)ctx->nmol, ctx->mol_start.ptr, ctx->position.ptr,
                                                         ctx->mass.ptr, ctx->mol_com.ptr);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------

int launch_pairs_allpairs(Context* ctx, const ComputeRequest& req) {
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    const int owned = (int)(hi - lo);

    AllPairsArgs a;
    a.n = (int)ctx->n;
    a.i_lo = (int)lo;
    a.i_hi = (int)hi;
    a.pos = ctx->position.ptr;
    a.charge = ctx->charge.ptr;
    a.kind = ctx->kind.ptr;
    a.mol_first = ctx->mol_first.ptr;
    a.bd_row = ctx->bd_row.ptr;
    a.mol_of = ctx->mol_of.ptr;
    a.bond_dist = ctx->bond_dist.ptr;
    a.mol_com = ctx->mol_com.ptr;
    a.nkinds = ctx->nkinds;
    a.pairs = ctx->pairs.ptr;
    a.tables = ctx->tables.ptr;
    a.table_energy = ctx->table_energy.ptr;
    a.table_force = ctx->table_force.ptr;
    a.cell = ctx->cell;
    a.coulomb = ctx->coulomb;
    a.do_pairs = req.pairs && ctx->any_pair;
    a.do_coulomb = req.coulomb && ctx->coulomb.kind != 0;
    a.force = ctx->force.ptr;
    a.write_forces = req.forces;

    const int blocks = owned > 0 ? (owned + ALLPAIRS_WARPS - 1) / ALLPAIRS_WARPS : 0;
    const size_t smem = sizeof(PairParams) * (size_t)ctx->nkinds * ctx->nkinds + 32 * ALLPAIRS_NV * sizeof(double);
    if (smem > 200 * 1024) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "too many particle kinds (%d) for the shared pair table",
                         ctx->nkinds);
    }
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)(blocks > 0 ? blocks : 1) * ALLPAIRS_NV));
    a.partials = ctx->partials.ptr;

    auto configure = [&](const void* kernel) -> cudaError_t {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    };

    const bool scalar_sums = req.energy || req.virial;
    if (blocks > 0 && (req.forces || scalar_sums)) {
        ScopedClock clock(ctx, &ctx->clk_pair);
        if (scalar_sums) {
            if (smem > 48 * 1024) LUMOL_CUDA_CHECK(ctx, configure((const void*)allpairs_kernel<MODE_FULL>));
            allpairs_kernel<MODE_FULL><<<blocks, ALLPAIRS_THREADS, smem, ctx->stream>>>(a);
        } else {
            if (smem > 48 * 1024) LUMOL_CUDA_CHECK(ctx, configure((const void*)allpairs_kernel<MODE_FORCES>));
            allpairs_kernel<MODE_FORCES><<<blocks, ALLPAIRS_THREADS, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
        ctx->clk_pair.launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }
    if (scalar_sums) {
        // slots RES_E_PAIRS, RES_E_COULOMB_REAL, RES_W_PAIRS[6], RES_W_COULOMB_REAL[6] are contiguous
        int status = launch_reduce(ctx, blocks, ALLPAIRS_NV, RES_E_PAIRS);
        if (status != 0) return status;
    }

    if (req.molecular_virial) {
        int status = launch_molecule_com(ctx);
        if (status != 0) return status;
        a.mol_com = ctx->mol_com.ptr;
        if (blocks > 0) {
            ScopedClock clock(ctx, &ctx->clk_pair);
            if (smem > 48 * 1024) LUMOL_CUDA_CHECK(ctx, configure((const void*)allpairs_kernel<MODE_MOLECULAR>));
            allpairs_kernel<MODE_MOLECULAR><<<blocks, ALLPAIRS_THREADS, smem, ctx->stream>>>(a);
            ctx->launches++;
            ctx->clk_pair.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        // the molecular kernel fills the same 14-value layout; keep it apart from the atomic sums
        status = launch_reduce(ctx, blocks, ALLPAIRS_NV, RES_MOLECULAR_BLOCK);
        if (status != 0) return status;
    }
    return 0;
}

}  // namespace lumol
