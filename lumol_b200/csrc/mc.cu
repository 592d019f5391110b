// Monte Carlo trial moves: the energy cost of moving one rigid molecule, evaluated against the resident state.
//   EnergyCache::move_molecule_cost          sys/cache.rs:145-213
//   Ewald::real_space_move_molecule_cost     energy/global/ewald.rs:572-613
//   Ewald::delta_rho_move_rigid_molecules    energy/global/ewald.rs:758-808
//   Ewald::k_space_move_molecule_cost        energy/global/ewald.rs:810-839
//   Wolf::move_molecule_cost (GlobalCache)   energy/global/wolf.rs:121-165
//   EnergyCache::update / Ewald updater      sys/cache.rs:175-211, ewald.rs:833-837
//
// The reference keeps an N x N table of pair energies on the host (pairs_cache) to know what the moved molecule
// contributed before the move.  Here nothing of size N^2 exists: the old and the new energy of every (moved atom,
// other atom) pair are evaluated side by side, one thread per other atom, the atoms of the moved molecule held in
// shared memory.  A batch of independent trial moves (grid.y) is evaluated against the same state in one launch;
// accepting one of them writes its positions into the resident arrays and adds its delta rho(k) to the structure
// factor, as the reference's updaters do.
#include "context.hpp"

namespace lumol {

constexpr int MOVE_THREADS = 256;
constexpr int MOVE_NV = 4;  // pairs new, pairs old, coulomb real space new, coulomb real space old

struct MovePairsArgs {
    int n;
    int max_size;  // atoms of the largest molecule: stride of the shared-memory arrays
    const int2* __restrict__ trials;      // (molecule, first row of its new positions)
    const double* __restrict__ new_pos;   // rows of 3
    const int* __restrict__ mol_start;
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    const unsigned* __restrict__ kind;
    int nkinds;
    const PairParams* __restrict__ pairs;
    const TableDesc* __restrict__ tables;
    const double* __restrict__ table_energy;
    const double* __restrict__ table_force;
    CellView cell;
    CoulombView coulomb;
    int do_pairs;
    int do_coulomb;
    double* __restrict__ partials;  // [trial][block][MOVE_NV]
};

// |image(a - b)|: cells.rs:316-320 with u = b, v = a
__device__ __forceinline__ double image_distance(const CellView& cell, double ax, double ay, double az, double bx, double by,
                                                 double bz) {
    double dx = ax - bx, dy = ay - by, dz = az - bz;
    vector_image_exact(cell, dx, dy, dz);
    return sqrt(dot3_exact(dx, dy, dz, dx, dy, dz));
}

// grid: (blocks over the atoms of the system, trials)
__global__ void __launch_bounds__(MOVE_THREADS) move_pairs_kernel(MovePairsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairParams* sp = reinterpret_cast<PairParams*>(smem_raw);
    double* old_pos = reinterpret_cast<double*>(smem_raw + sizeof(PairParams) * a.nkinds * a.nkinds);
    double* new_pos = old_pos + 3 * a.max_size;
    double* q = new_pos + 3 * a.max_size;
    double* scratch = q + a.max_size;                                   // 32 * MOVE_NV
    unsigned* k = reinterpret_cast<unsigned*>(scratch + 32 * MOVE_NV);  // max_size

    const int2 trial = a.trials[blockIdx.y];
    const int first = a.mol_start[trial.x], last = a.mol_start[trial.x + 1];
    const int size = last - first;
    {
        const int words = (int)(sizeof(PairParams) / sizeof(double)) * a.nkinds * a.nkinds;
        const double* src = reinterpret_cast<const double*>(a.pairs);
        double* dst = reinterpret_cast<double*>(sp);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
        for (int w = threadIdx.x; w < 3 * size; w += blockDim.x) {
            old_pos[w] = a.pos[3 * (size_t)first + w];
            new_pos[w] = a.new_pos[3 * (size_t)trial.y + w];
        }
        for (int w = threadIdx.x; w < size; w += blockDim.x) {
            q[w] = a.charge[first + w];
            k[w] = a.kind[first + w];
        }
    }
    __syncthreads();

    double acc[MOVE_NV];
#pragma unroll
    for (int v = 0; v < MOVE_NV; v++) acc[v] = 0.0;

    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += gridDim.x * blockDim.x) {
        if (j >= first && j < last) continue;  // cache.rs:159: every molecule but the moved one
        const double xj = a.pos[3 * (size_t)j], yj = a.pos[3 * (size_t)j + 1], zj = a.pos[3 * (size_t)j + 2];
        const double qj = a.charge[j];
        const unsigned kj = a.kind[j];
        for (int i = 0; i < size; i++) {
            const double r_new = image_distance(a.cell, new_pos[3 * i], new_pos[3 * i + 1], new_pos[3 * i + 2], xj, yj, zj);
            const double r_old = image_distance(a.cell, old_pos[3 * i], old_pos[3 * i + 1], old_pos[3 * i + 2], xj, yj, zj);
            if (a.do_pairs) {
                // EnergyEvaluator::pair (energy.rs:32-45) with BondPath::None: the atoms are in different molecules
                const PairParams& pp = sp[k[i] * a.nkinds + kj];
                double scaling;
                if (pp.potential > LUMOL_CUDA_POTENTIAL_NULL && !restriction_excluded(pp.restriction, 0u, pp.scale14, scaling)) {
                    double e, f;
                    if (r_new < pp.cutoff) {
                        pair_eval(pp, a.tables, a.table_energy, a.table_force, r_new, e, f);
                        acc[0] += scaling * e;
                    }
                    if (r_old < pp.cutoff) {
                        pair_eval(pp, a.tables, a.table_energy, a.table_force, r_old, e, f);
                        acc[1] += scaling * e;
                    }
                }
            }
            if (a.do_coulomb) {
                const double qi = q[i];
                if (qi != 0.0 && qj != 0.0) {
                    double scaling;
                    const bool excluded = restriction_excluded(a.coulomb.restriction, 0u, a.coulomb.scale14, scaling);
                    double e, fr;
                    if (a.coulomb.kind == 1) {
                        // ewald.rs:387-400: excluded pairs carry the -erf term
                        if (r_new <= a.coulomb.rc) {
                            ewald_real_pair(a.coulomb, excluded, qi * qj, r_new, e, fr);
                            acc[2] += e;
                        }
                        if (r_old <= a.coulomb.rc) {
                            ewald_real_pair(a.coulomb, excluded, qi * qj, r_old, e, fr);
                            acc[3] += e;
                        }
                    } else if (!excluded) {
                        // wolf.rs:149-158
                        if (r_new <= a.coulomb.rc) {
                            wolf_pair(a.coulomb, qi * qj, r_new, e, fr);
                            acc[2] += scaling * e;
                        }
                        if (r_old <= a.coulomb.rc) {
                            wolf_pair(a.coulomb, qi * qj, r_old, e, fr);
                            acc[3] += scaling * e;
                        }
                    }
                }
            }
        }
    }

    block_sum<MOVE_NV>(acc, scratch);
    if (threadIdx.x == 0) {
        double* out = a.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * MOVE_NV;
#pragma unroll
        for (int v = 0; v < MOVE_NV; v++) out[v] = acc[v];
    }
}

// ------------------------------------------------------------------------------------------------
// reciprocal space
// ------------------------------------------------------------------------------------------------

struct MoveKspaceArgs {
    int max_size;
    int kmax;
    int nk;
    double basis[9];  // rows: k_vector of the three unit indices
    const int2* __restrict__ trials;
    const double* __restrict__ new_pos;
    const int* __restrict__ mol_start;
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    const short4* __restrict__ kindex;
    const double* __restrict__ kenergy;
    const double2* __restrict__ rho;
    double2* __restrict__ delta_rho;  // [trial][nk]
    double* __restrict__ partials;    // [trial][block][2]: sum factor |rho + delta|^2, sum factor |rho|^2
};

__device__ __forceinline__ double2 move_cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);  // complex.rs:219-228
}

__device__ __forceinline__ double2 move_table_at(const double2* __restrict__ t, int idx) {
    double2 v = t[idx < 0 ? -idx : idx];
    if (idx < 0) v.y = -v.y;  // eikr[-k] = conj(eikr[k]), ewald.rs:771, 780
    return v;
}

// grid: (blocks over the k-vectors, trials).  Shared: phase tables of the molecule at its old and at its new
// positions, [old | new][atom][axis][0..kmax], built by the reference's recursion (ewald.rs:765-783).
__global__ void __launch_bounds__(MOVE_THREADS) move_kspace_kernel(MoveKspaceArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int stride = 3 * (a.kmax + 1);
    double2* table = reinterpret_cast<double2*>(smem_raw);
    double* q = reinterpret_cast<double*>(table + (size_t)2 * a.max_size * stride);
    double* scratch = q + a.max_size;  // 32 * 2

    const int2 trial = a.trials[blockIdx.y];
    const int first = a.mol_start[trial.x];
    const int size = a.mol_start[trial.x + 1] - first;

    for (int w = threadIdx.x; w < 2 * 3 * size; w += blockDim.x) {
        const int which = w / (3 * size);  // 0: old positions, 1: new positions
        const int rest = w - which * 3 * size;
        const int atom = rest / 3, axis = rest - 3 * atom;
        const double* x = which == 0 ? a.pos + 3 * (size_t)(first + atom) : a.new_pos + 3 * (size_t)(trial.y + atom);
        const double phase = a.basis[3 * axis] * x[0] + a.basis[3 * axis + 1] * x[1] + a.basis[3 * axis + 2] * x[2];
        double sn, cs;
        sincos(phase, &sn, &cs);
        double2* t = table + ((size_t)which * a.max_size + atom) * stride + axis * (a.kmax + 1);
        const double2 e1 = make_double2(cs, sn);
        double2 e = make_double2(1.0, 0.0);
        t[0] = e;
        if (a.kmax >= 1) {
            e = e1;
            t[1] = e;
        }
        for (int m = 2; m <= a.kmax; m++) {
            e = move_cmul(e, e1);
            t[m] = e;
        }
    }
    for (int w = threadIdx.x; w < size; w += blockDim.x) q[w] = a.charge[first + w];
    __syncthreads();

    double acc[2] = {0.0, 0.0};
    const int ik = blockIdx.x * blockDim.x + threadIdx.x;
    if (ik < a.nk) {
        const short4 idx = a.kindex[ik];
        double2 partial = make_double2(0.0, 0.0);
        for (int atom = 0; atom < size; atom++) {
            const double2* t_old = table + (size_t)atom * stride;
            const double2* t_new = table + ((size_t)a.max_size + atom) * stride;
            const double2 old_phi = move_cmul(move_cmul(move_table_at(t_old, idx.x), move_table_at(t_old + (a.kmax + 1), idx.y)),
                                              move_table_at(t_old + 2 * (a.kmax + 1), idx.z));
            const double2 new_phi = move_cmul(move_cmul(move_table_at(t_new, idx.x), move_table_at(t_new + (a.kmax + 1), idx.y)),
                                              move_table_at(t_new + 2 * (a.kmax + 1), idx.z));
            partial.x += q[atom] * (new_phi.x - old_phi.x);  // ewald.rs:802
            partial.y += q[atom] * (new_phi.y - old_phi.y);
        }
        a.delta_rho[(size_t)blockIdx.y * a.nk + ik] = partial;
        const double2 r = a.rho[ik];
        const double factor = a.kenergy[ik];
        const double nx = r.x + partial.x, ny = r.y + partial.y;
        acc[0] = factor * (nx * nx + ny * ny);  // ewald.rs:826-828
        acc[1] = factor * (r.x * r.x + r.y * r.y);  // ewald.rs:816-818
    }
    block_sum<2>(acc, scratch);
    if (threadIdx.x == 0) {
        double* out = a.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
        out[0] = acc[0];
        out[1] = acc[1];
    }
}

// One block per trial folds the per-block partial sums in a fixed order: results[trial][6] =
// pairs new, pairs old, coulomb real new, coulomb real old, k-space new, k-space old.
__global__ void __launch_bounds__(256) move_finish_kernel(int pair_blocks, const double* __restrict__ pair_partials, int k_blocks,
                                                          const double* __restrict__ k_partials, double* __restrict__ results) {
    __shared__ double scratch[32 * 6];
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int t = blockIdx.x;
    for (int b = threadIdx.x; b < pair_blocks; b += blockDim.x) {
        const double* p = pair_partials + ((size_t)t * pair_blocks + b) * MOVE_NV;
#pragma unroll
        for (int v = 0; v < MOVE_NV; v++) acc[v] += p[v];
    }
    for (int b = threadIdx.x; b < k_blocks; b += blockDim.x) {
        const double* p = k_partials + ((size_t)t * k_blocks + b) * 2;
        acc[4] += p[0];
        acc[5] += p[1];
    }
    block_sum<6>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int v = 0; v < 4; v++) results[(size_t)t * 6 + v] = acc[v];
        results[(size_t)t * 6 + 4] = acc[4] / FOUR_PI_EPSILON_0;  // ewald.rs:829
        results[(size_t)t * 6 + 5] = acc[5] / FOUR_PI_EPSILON_0;  // ewald.rs:819
    }
}

// EnergyCache::update after an accepted move: the molecule takes its new positions (the reference's caller does
// that itself, e.g. mc/moves/translate.rs:115-120), rho(k) += delta rho(k) (ewald.rs:833-837).
__global__ void __launch_bounds__(256) move_accept_kernel(int first, int size, const double* __restrict__ new_pos, double* __restrict__ pos,
                                                          int nk, const double2* __restrict__ delta_rho, double2* __restrict__ rho) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < 3 * size) pos[3 * (size_t)first + w] = new_pos[w];
    if (w < nk) {
        double2 r = rho[w];
        const double2 d = delta_rho[w];
        r.x += d.x;
        r.y += d.y;
        rho[w] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------

// Costs of the `ntrials` moves already uploaded to ctx->mc_trials / ctx->mc_new_pos; leaves 6 sums per trial in
// ctx->mc_results.  `max_size`: atoms of the largest molecule among the trials.
int launch_move_cost(Context* ctx, int ntrials, int max_size) {
    const bool do_pairs = ctx->any_pair;
    const bool do_coulomb = ctx->coulomb.kind != 0;
    const bool do_kspace = ctx->coulomb.kind == 1;
    const int n = (int)ctx->n;

    // few blocks per trial when there are many trials, enough to fill the device when there is one
    int pair_blocks = (n + MOVE_THREADS - 1) / MOVE_THREADS;
    int cap = 4 * ctx->sm_count / ntrials;
    if (cap < 1) cap = 1;
    if (pair_blocks > cap) pair_blocks = cap;
    if (pair_blocks < 1) pair_blocks = 1;
    LUMOL_CUDA_CHECK(ctx, ctx->mc_results.reserve((size_t)ntrials * 6));
    LUMOL_CUDA_CHECK(ctx, ctx->mc_pair_partials.reserve((size_t)ntrials * pair_blocks * MOVE_NV));

    {
        MovePairsArgs a;
        a.n = n;
        a.max_size = max_size;
        a.trials = ctx->mc_trials.ptr;
        a.new_pos = ctx->mc_new_pos.ptr;
        a.mol_start = ctx->mol_start.ptr;
        a.pos = ctx->position.ptr;
        a.charge = ctx->charge.ptr;
        a.kind = ctx->kind.ptr;
        a.nkinds = ctx->nkinds;
        a.pairs = ctx->pairs.ptr;
        a.tables = ctx->tables.ptr;
        a.table_energy = ctx->table_energy.ptr;
        a.table_force = ctx->table_force.ptr;
        a.cell = ctx->cell;
        a.coulomb = ctx->coulomb;
        a.do_pairs = do_pairs;
        a.do_coulomb = do_coulomb;
        a.partials = ctx->mc_pair_partials.ptr;
        const size_t smem = sizeof(PairParams) * (size_t)ctx->nkinds * ctx->nkinds +
                            ((size_t)7 * max_size + 32 * MOVE_NV) * sizeof(double) + (size_t)max_size * sizeof(unsigned);
        if (smem > 200 * 1024) {
            return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "a molecule of %d atoms with %d particle kinds does not fit in shared memory",
                             max_size, ctx->nkinds);
        }
        if (smem > 48 * 1024) {
            LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute((const void*)move_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        move_pairs_kernel<<<dim3(pair_blocks, ntrials), MOVE_THREADS, smem, ctx->stream>>>(a);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }

    int k_blocks = 0;
    if (do_kspace && ctx->nk > 0) {
        const int nk = (int)ctx->nk;
        k_blocks = (nk + MOVE_THREADS - 1) / MOVE_THREADS;
        LUMOL_CUDA_CHECK(ctx, ctx->mc_delta_rho.reserve((size_t)ntrials * nk));
        LUMOL_CUDA_CHECK(ctx, ctx->mc_k_partials.reserve((size_t)ntrials * k_blocks * 2));
        MoveKspaceArgs a;
        a.max_size = max_size;
        a.kmax = ctx->kmax;
        a.nk = nk;
        for (int k = 0; k < 9; k++) a.basis[k] = ctx->kbasis[k];
        a.trials = ctx->mc_trials.ptr;
        a.new_pos = ctx->mc_new_pos.ptr;
        a.mol_start = ctx->mol_start.ptr;
        a.pos = ctx->position.ptr;
        a.charge = ctx->charge.ptr;
        a.kindex = ctx->kindex.ptr;
        a.kenergy = ctx->kenergy.ptr;
        a.rho = ctx->rho.ptr;
        a.delta_rho = ctx->mc_delta_rho.ptr;
        a.partials = ctx->mc_k_partials.ptr;
        const size_t smem = (size_t)2 * max_size * 3 * (ctx->kmax + 1) * sizeof(double2) + ((size_t)max_size + 64) * sizeof(double);
        if (smem > 200 * 1024) {
            return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "the phase tables of a molecule of %d atoms with kmax = %d do not fit in shared memory",
                             max_size, ctx->kmax);
        }
        if (smem > 48 * 1024) {
            LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute((const void*)move_kspace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        move_kspace_kernel<<<dim3(k_blocks, ntrials), MOVE_THREADS, smem, ctx->stream>>>(a);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }

    move_finish_kernel<<<ntrials, 256, 0, ctx->stream>>>(pair_blocks, ctx->mc_pair_partials.ptr, k_blocks, ctx->mc_k_partials.ptr,
                                                         ctx->mc_results.ptr);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

int launch_move_accept(Context* ctx, int trial, int first, int size, int64_t row) {
    const bool with_rho = ctx->coulomb.kind == 1 && ctx->nk > 0;
    const int nk = with_rho ? (int)ctx->nk : 0;
    const int work = nk > 3 * size ? nk : 3 * size;
    move_accept_kernel<<<(work + 255) / 256, 256, 0, ctx->stream>>>(first, size, ctx->mc_new_pos.ptr + 3 * row, ctx->position.ptr, nk,
                                                                    with_rho ? ctx->mc_delta_rho.ptr + (size_t)trial * nk : nullptr,
                                                                    ctx->rho.ptr);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

}  // namespace lumol
