// Ewald reciprocal space.
//   compute_ewald_factors / Ewald::prepare      energy/global/ewald.rs:115-183, 353-378
//   Ewald::eik_dot_r (structure factor rho(k))  energy/global/ewald.rs:633-674
//   k_space_energy / forces / atomic_virial     energy/global/ewald.rs:677-733
//   k_space_molecular_virial                    energy/global/ewald.rs:736-753
//
// The reference caches e^{i k_c x_i} for every (k index, axis, atom) in a (2 kmax + 1) x 3 x N array and
// rebuilds it serially on every call.  Here nothing of size N x kmax touches HBM: each block rebuilds
// the per-axis phase tables of its own tile of atoms in shared memory by the same recursion
// e(m) = e(m - 1) * e(1) (ewald.rs:655-662), then
//   rho kernel   : one thread per k-vector, the tile's atoms streamed from shared memory, partial
//                  structure factors written per atom chunk and summed in a fixed order (no atomics);
//   force kernel : one thread per atom, looping over every k-vector with rho(k) and the factor table
//                  staged through shared memory: F_i += q_i / 4 pi eps0 * sum_k Im(e^{ik.r_i} conj rho_k) field_k.
#include "context.hpp"

#include <cmath>

namespace lumol {

// ------------------------------------------------------------------------------------------------
// factor table (host, FP64, same enumeration order and arithmetic as the reference)
// ------------------------------------------------------------------------------------------------

static void host_k_vector(const double inv[9], const double idx[3], double k[3]) {
    // cells.rs:224-226: (2 pi * inv) * index
    const double two_pi = 2.0 * PI;
    double m[9];
    for (int a = 0; a < 9; a++) m[a] = two_pi * inv[a];
    for (int a = 0; a < 3; a++) k[a] = m[3 * a] * idx[0] + m[3 * a + 1] * idx[1] + m[3 * a + 2] * idx[2];
}

static double host_volume(const CellView& c) {
    if (c.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC) {
        return c.h[0] * c.h[4] * c.h[8];
    }
    // a . (b ^ c) with the lattice vectors as matrix columns (cells.rs:185-199)
    const double a[3] = {c.h[0], c.h[3], c.h[6]}, b[3] = {c.h[1], c.h[4], c.h[7]}, cc[3] = {c.h[2], c.h[5], c.h[8]};
    const double bx = b[1] * cc[2] - b[2] * cc[1], by = b[2] * cc[0] - b[0] * cc[2], bz = b[0] * cc[1] - b[1] * cc[0];
    return a[0] * bx + a[1] * by + a[2] * bz;
}

int ewald_prepare(Context* ctx) {
    if (ctx->cell.shape == LUMOL_CUDA_CELL_INFINITE) {
        return ctx->fail(LUMOL_CUDA_ERROR_INFINITE_CELL, "Ewald is not defined with infinite unit cell");
    }
    if (ctx->ewald_generation == ctx->cell_generation) {
        return 0;  // Ewald::prepare: nothing to do while the cell is unchanged (ewald.rs:354-359)
    }
    const int kmax = ctx->kmax;
    if (kmax > 32767) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "kmax = %d is too large", kmax);
    }
    const double* inv = ctx->cell.inv;
    const double ones[3] = {1.0, 1.0, 1.0};
    double k111[3];
    host_k_vector(inv, ones, k111);
    const double maxc = std::fmax(std::fmax(k111[0], k111[1]), k111[2]) * (double)kmax;
    ctx->kmax2 = 1.0001 * maxc * maxc;  // ewald.rs:362-363

    const double alpha = ctx->coulomb.alpha;
    const double alpha_sq_inv_fourth = 0.25 / (alpha * alpha);
    const double four_pi_v = 4.0 * PI / host_volume(ctx->cell);

    std::vector<short4> index;
    std::vector<double> energy, virial;
    ctx->host_kindex.clear();
    ctx->host_kenergy.clear();
    auto push = [&](int ikx, int iky, int ikz) {
        const double idx[3] = {(double)ikx, (double)iky, (double)ikz};
        double kv[3];
        host_k_vector(inv, idx, kv);
        const double k2 = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
        if (k2 > ctx->kmax2) return;
        // ewald.rs:134-141
        const double e = four_pi_v * std::exp(-k2 * alpha_sq_inv_fourth) / k2;
        const double virial_factor = -2.0 * (1.0 / k2 + alpha_sq_inv_fourth);
        index.push_back(make_short4((short)ikx, (short)iky, (short)ikz, 0));
        energy.push_back(e);
        const int map[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
        for (int c = 0; c < 6; c++) {
            const int a = map[c][0], b = map[c][1];
            const double one = a == b ? 1.0 : 0.0;
            virial.push_back(e * (one + virial_factor * (kv[a] * kv[b])));
        }
        ctx->host_kindex.push_back(ikx);
        ctx->host_kindex.push_back(iky);
        ctx->host_kindex.push_back(ikz);
        ctx->host_kenergy.push_back(e);
    };
    // enumeration order of ewald.rs:144-182 (upper bounds exclusive)
    for (int ikx = 1; ikx < kmax; ikx++)
        for (int iky = -kmax; iky < kmax; iky++)
            for (int ikz = -kmax; ikz < kmax; ikz++) push(ikx, iky, ikz);
    for (int iky = 1; iky < kmax; iky++)
        for (int ikz = -kmax; ikz < kmax; ikz++) push(0, iky, ikz);
    for (int ikz = 1; ikz < kmax; ikz++) push(0, 0, ikz);

    ctx->nk = (int64_t)index.size();
    const double e0[3] = {1, 0, 0}, e1[3] = {0, 1, 0}, e2[3] = {0, 0, 1};
    host_k_vector(inv, e0, ctx->kbasis);
    host_k_vector(inv, e1, ctx->kbasis + 3);
    host_k_vector(inv, e2, ctx->kbasis + 6);

    const size_t nk = index.size() > 0 ? index.size() : 1;
    LUMOL_CUDA_CHECK(ctx, ctx->kindex.reserve(nk));
    LUMOL_CUDA_CHECK(ctx, ctx->kenergy.reserve(nk));
    LUMOL_CUDA_CHECK(ctx, ctx->kvirial.reserve(nk * 6));
    LUMOL_CUDA_CHECK(ctx, ctx->rho.reserve(nk));
    if (!index.empty()) {
        LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->kindex.ptr, index.data(), index.size() * sizeof(short4),
                                              cudaMemcpyHostToDevice, ctx->stream));
        LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->kenergy.ptr, energy.data(), energy.size() * sizeof(double),
                                              cudaMemcpyHostToDevice, ctx->stream));
        LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->kvirial.ptr, virial.data(), virial.size() * sizeof(double),
                                              cudaMemcpyHostToDevice, ctx->stream));
        LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // the vectors above go out of scope
    }
    ctx->ewald_generation = ctx->cell_generation;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------

struct KBasis {
    double b[9];  // rows: k_vector([1,0,0]), k_vector([0,1,0]), k_vector([0,0,1])
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    // complex.rs:219-228
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// phase table entry for a signed index: eikr[-k] = conj(eikr[k]) (ewald.rs:652, 660)
__device__ __forceinline__ double2 table_at(const double2* __restrict__ t, int idx) {
    double2 v = t[idx < 0 ? -idx : idx];
    if (idx < 0) v.y = -v.y;
    return v;
}

// ------------------------------------------------------------------------------------------------
// structure factor
// ------------------------------------------------------------------------------------------------

constexpr int RHO_THREADS = 256;

struct RhoArgs {
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    int a_lo, a_hi;  // atoms of this rank
    int nchunks;     // gridDim.y
    int tile;        // atoms per shared-memory tile
    int kmax;
    int nk;
    KBasis basis;
    const short4* __restrict__ kindex;
    double2* __restrict__ rho_partial;  // nchunks x nk
};

// grid: (ceil(nk / RHO_THREADS), nchunks).  Shared: tile x 3 x (kmax + 1) complex + tile charges.
__global__ void __launch_bounds__(RHO_THREADS) ewald_rho_kernel(RhoArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int stride = 3 * (a.kmax + 1);
    double2* table = reinterpret_cast<double2*>(smem_raw);                 // [atom][axis][m]
    double* q = reinterpret_cast<double*>(table + (size_t)a.tile * stride);  // [atom]

    const int ik = blockIdx.x * RHO_THREADS + threadIdx.x;
    short4 idx = make_short4(0, 0, 0, 0);
    if (ik < a.nk) idx = a.kindex[ik];

    // contiguous slice of this rank's atoms for this chunk
    const int owned = a.a_hi - a.a_lo;
    const int per_chunk = (owned + a.nchunks - 1) / a.nchunks;
    const int c_lo = a.a_lo + blockIdx.y * per_chunk;
    int c_hi = c_lo + per_chunk;
    if (c_hi > a.a_hi) c_hi = a.a_hi;

    double2 acc = make_double2(0.0, 0.0);
    for (int base = c_lo; base < c_hi; base += a.tile) {
        const int count = min(a.tile, c_hi - base);
        __syncthreads();
        // phase tables: 3 * count independent recursions
        for (int w = threadIdx.x; w < 3 * count; w += RHO_THREADS) {
            const int atom = w / 3, axis = w - 3 * atom;
            const int i = base + atom;
            const double phase = a.basis.b[3 * axis] * a.pos[3 * i] + a.basis.b[3 * axis + 1] * a.pos[3 * i + 1] +
                                 a.basis.b[3 * axis + 2] * a.pos[3 * i + 2];
            double sn, cs;
            sincos(phase, &sn, &cs);
            double2* t = table + (size_t)atom * stride + axis * (a.kmax + 1);
            const double2 e1 = make_double2(cs, sn);
            double2 e = make_double2(1.0, 0.0);
            t[0] = e;
            if (a.kmax >= 1) {
                e = e1;
                t[1] = e;
            }
            for (int m = 2; m <= a.kmax; m++) {
                e = cmul(e, e1);
                t[m] = e;
            }
        }
        for (int w = threadIdx.x; w < count; w += RHO_THREADS) {
            q[w] = a.charge[base + w];
        }
        __syncthreads();
        if (ik < a.nk) {
            for (int atom = 0; atom < count; atom++) {
                const double2* t = table + (size_t)atom * stride;
                const double2 ex = table_at(t, idx.x);
                const double2 ey = table_at(t + (a.kmax + 1), idx.y);
                const double2 ez = table_at(t + 2 * (a.kmax + 1), idx.z);
                const double2 phi = cmul(cmul(ex, ey), ez);  // ewald.rs:666-668
                const double qa = q[atom];
                acc.x += qa * phi.x;
                acc.y += qa * phi.y;
            }
        }
    }
    if (ik < a.nk) {
        a.rho_partial[(size_t)blockIdx.y * a.nk + ik] = acc;
    }
}

// rho(k) = sum over atom chunks, in chunk order; fused with the energy / virial sums over k.
struct RhoReduceArgs {
    int nk;
    int nchunks;
    const double2* __restrict__ rho_partial;
    double2* __restrict__ rho;
};

__global__ void __launch_bounds__(256) ewald_rho_reduce_kernel(RhoReduceArgs a) {
    const int ik = blockIdx.x * blockDim.x + threadIdx.x;
    if (ik >= a.nk) return;
    double2 acc = make_double2(0.0, 0.0);
    for (int c = 0; c < a.nchunks; c++) {
        const double2 v = a.rho_partial[(size_t)c * a.nk + ik];
        acc.x += v.x;
        acc.y += v.y;
    }
    a.rho[ik] = acc;
}

constexpr int KSUM_NV = 7;  // energy, W[6]

// k_space_energy (ewald.rs:677-687) and k_space_atomic_virial (ewald.rs:722-733)
__global__ void __launch_bounds__(256)
    ewald_ksum_kernel(int nk, const double2* __restrict__ rho, const double* __restrict__ kenergy,
                      const double* __restrict__ kvirial, double* __restrict__ partials) {
    __shared__ double scratch[32 * KSUM_NV];
    double acc[KSUM_NV];
#pragma unroll
    for (int k = 0; k < KSUM_NV; k++) acc[k] = 0.0;
    for (int ik = blockIdx.x * blockDim.x + threadIdx.x; ik < nk; ik += gridDim.x * blockDim.x) {
        const double2 r = rho[ik];
        const double n2 = r.x * r.x + r.y * r.y;
        acc[0] += kenergy[ik] * n2;
#pragma unroll
        for (int c = 0; c < 6; c++) acc[1 + c] += n2 * kvirial[(size_t)ik * 6 + c];
    }
    block_sum<KSUM_NV>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < KSUM_NV; k++) partials[(size_t)blockIdx.x * KSUM_NV + k] = acc[k] / FOUR_PI_EPSILON_0;
    }
}

// ------------------------------------------------------------------------------------------------
// forces
// ------------------------------------------------------------------------------------------------

constexpr int KFORCE_THREADS = 128;
constexpr int KFORCE_STAGE = 256;  // k-vectors staged through shared memory per step

struct KForceArgs {
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    int a_lo, a_hi;
    int kmax;
    int nk;
    KBasis basis;
    const short4* __restrict__ kindex;
    const double* __restrict__ kenergy;
    const double2* __restrict__ rho;
    double* __restrict__ force;  // accumulated
    // molecular virial correction (ewald.rs:736-753): sum_i f_i (x) (x_i - com(mol_i))
    int do_correction;
    const int* __restrict__ mol_of;
    const double* __restrict__ mol_com;
    double* __restrict__ partials;  // 9 per block
    int add_to_force;
};

// Shared: 3 x (kmax + 1) x KFORCE_THREADS complex, laid out [axis][m][thread] so a warp reading one m
// touches consecutive words; plus the staged k-vectors.
__global__ void __launch_bounds__(KFORCE_THREADS) ewald_force_kernel(KForceArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* table = reinterpret_cast<double2*>(smem_raw);
    const int rows = 3 * (a.kmax + 1);
    double2* s_rho = table + (size_t)rows * KFORCE_THREADS;
    double* s_energy = reinterpret_cast<double*>(s_rho + KFORCE_STAGE);
    short4* s_index = reinterpret_cast<short4*>(s_energy + KFORCE_STAGE);
    __shared__ double scratch[32 * 9];

    const int i = a.a_lo + blockIdx.x * KFORCE_THREADS + threadIdx.x;
    const bool active = i < a.a_hi;
    double2* mine = table + threadIdx.x;

    double px = 0.0, py = 0.0, pz = 0.0;
    if (active) {
        px = a.pos[3 * i];
        py = a.pos[3 * i + 1];
        pz = a.pos[3 * i + 2];
#pragma unroll
        for (int axis = 0; axis < 3; axis++) {
            const double phase = a.basis.b[3 * axis] * px + a.basis.b[3 * axis + 1] * py + a.basis.b[3 * axis + 2] * pz;
            double sn, cs;
            sincos(phase, &sn, &cs);
            const double2 e1 = make_double2(cs, sn);
            double2 e = make_double2(1.0, 0.0);
            double2* t = mine + (size_t)axis * (a.kmax + 1) * KFORCE_THREADS;
            t[0] = e;
            if (a.kmax >= 1) {
                e = e1;
                t[KFORCE_THREADS] = e;
            }
            for (int m = 2; m <= a.kmax; m++) {
                e = cmul(e, e1);
                t[(size_t)m * KFORCE_THREADS] = e;
            }
        }
    }

    // field_i = sum_k Im(phi conj rho) * 2 e_k k_vec; k_vec = h b0 + k b1 + l b2, so accumulate the three
    // index-weighted sums and apply the basis once at the end.
    double sh = 0.0, sk = 0.0, sl = 0.0;
    for (int base = 0; base < a.nk; base += KFORCE_STAGE) {
        const int count = min(KFORCE_STAGE, a.nk - base);
        __syncthreads();
        for (int w = threadIdx.x; w < count; w += KFORCE_THREADS) {
            s_rho[w] = a.rho[base + w];
            s_energy[w] = a.kenergy[base + w];
            s_index[w] = a.kindex[base + w];
        }
        __syncthreads();
        if (active) {
            for (int w = 0; w < count; w++) {
                const short4 idx = s_index[w];
                const int hx = idx.x < 0 ? -idx.x : idx.x;
                const int hy = idx.y < 0 ? -idx.y : idx.y;
                const int hz = idx.z < 0 ? -idx.z : idx.z;
                double2 ex = mine[(size_t)hx * KFORCE_THREADS];
                double2 ey = mine[(size_t)(a.kmax + 1 + hy) * KFORCE_THREADS];
                double2 ez = mine[(size_t)(2 * (a.kmax + 1) + hz) * KFORCE_THREADS];
                if (idx.x < 0) ex.y = -ex.y;
                if (idx.y < 0) ey.y = -ey.y;
                if (idx.z < 0) ez.y = -ez.y;
                const double2 phi = cmul(cmul(ex, ey), ez);
                const double2 r = s_rho[w];
                // Im(phi * conj(rho)) (ewald.rs:706-707)
                const double im = phi.y * r.x - phi.x * r.y;
                const double t = im * (2.0 * s_energy[w]);
                sh += t * (double)idx.x;
                sk += t * (double)idx.y;
                sl += t * (double)idx.z;
            }
        }
    }

    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) acc[k] = 0.0;
    if (active) {
        const double scale = a.charge[i] / FOUR_PI_EPSILON_0;  // ewald.rs:716-718
        const double fx = scale * (sh * a.basis.b[0] + sk * a.basis.b[3] + sl * a.basis.b[6]);
        const double fy = scale * (sh * a.basis.b[1] + sk * a.basis.b[4] + sl * a.basis.b[7]);
        const double fz = scale * (sh * a.basis.b[2] + sk * a.basis.b[5] + sl * a.basis.b[8]);
        if (a.add_to_force) {
            a.force[3 * i] += fx;
            a.force[3 * i + 1] += fy;
            a.force[3 * i + 2] += fz;
        }
        if (a.do_correction) {
            const int m = a.mol_of[i];
            const double dx = px - a.mol_com[3 * m], dy = py - a.mol_com[3 * m + 1], dz = pz - a.mol_com[3 * m + 2];
            acc[0] = fx * dx;
            acc[1] = fx * dy;
            acc[2] = fx * dz;
            acc[3] = fy * dx;
            acc[4] = fy * dy;
            acc[5] = fy * dz;
            acc[6] = fz * dx;
            acc[7] = fz * dy;
            acc[8] = fz * dz;
        }
    }
    if (a.do_correction) {
        block_sum<9>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) a.partials[(size_t)blockIdx.x * 9 + k] = acc[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------

int launch_ewald_kspace(Context* ctx, const ComputeRequest& req) {
    int status = ewald_prepare(ctx);
    if (status != 0) return status;
    const int nk = (int)ctx->nk;
    int64_t lo, hi;
    ctx->owned_range(ctx->n, lo, hi);
    const int owned = (int)(hi - lo);
    const int kmax = ctx->kmax;

    KBasis basis;
    for (int k = 0; k < 9; k++) basis.b[k] = ctx->kbasis[k];

    if (nk == 0) {
        LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->results.ptr + RES_E_KSPACE, 0, 7 * sizeof(double), ctx->stream));
        LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->results.ptr + RES_W_KSPACE_CORRECTION, 0, 9 * sizeof(double),
                                              ctx->stream));
        return 0;
    }

    // ---- rho(k) ---------------------------------------------------------------------------------
    {
        const int kblocks = (nk + RHO_THREADS - 1) / RHO_THREADS;
        // shared-memory tile: as many atoms as fit in ~96 KB, capped at 256
        const size_t per_atom = (size_t)3 * (kmax + 1) * sizeof(double2) + sizeof(double);
        int tile = (int)((96 * 1024) / per_atom);
        if (tile > 256) tile = 256;
        if (tile < 1) {
            return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "kmax = %d needs more shared memory than one atom tile", kmax);
        }
        // enough chunks to fill the GPU a few times over, never more than one chunk per tile
        int nchunks = (4 * ctx->sm_count + kblocks - 1) / kblocks;
        const int max_chunks = (owned + tile - 1) / tile;
        if (nchunks > max_chunks) nchunks = max_chunks;
        if (nchunks < 1) nchunks = 1;
        if (nchunks > 65535) nchunks = 65535;
        LUMOL_CUDA_CHECK(ctx, ctx->rho_partial.reserve((size_t)nchunks * nk));

        RhoArgs a;
        a.pos = ctx->position.ptr;
        a.charge = ctx->charge.ptr;
        a.a_lo = (int)lo;
        a.a_hi = (int)hi;
        a.nchunks = nchunks;
        a.tile = tile;
        a.kmax = kmax;
        a.nk = nk;
        a.basis = basis;
        a.kindex = ctx->kindex.ptr;
        a.rho_partial = ctx->rho_partial.ptr;
        const size_t smem = per_atom * tile;
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute((const void*)ewald_rho_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            ScopedClock clock(ctx, &ctx->clk_kspace);
            ewald_rho_kernel<<<dim3(kblocks, nchunks), RHO_THREADS, smem, ctx->stream>>>(a);
            ctx->launches++;
            ctx->clk_kspace.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        RhoReduceArgs r;
        r.nk = nk;
        r.nchunks = nchunks;
        r.rho_partial = ctx->rho_partial.ptr;
        r.rho = ctx->rho.ptr;
        ewald_rho_reduce_kernel<<<(nk + 255) / 256, 256, 0, ctx->stream>>>(r);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        // every rank needs the full structure factor
        if (ctx->nranks > 1) {
            status = comm_allreduce(ctx, reinterpret_cast<double*>(ctx->rho.ptr), 2 * (int64_t)nk);
            if (status != 0) return status;
        }
    }

    // ---- energy and virial ------------------------------------------------------------------------
    if (req.energy || req.virial || req.molecular_virial) {
        int blocks = (nk + 255) / 256;
        if (blocks > 256) blocks = 256;
        LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)blocks * KSUM_NV));
        ewald_ksum_kernel<<<blocks, 256, 0, ctx->stream>>>(nk, ctx->rho.ptr, ctx->kenergy.ptr, ctx->kvirial.ptr,
                                                           ctx->partials.ptr);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        status = launch_reduce(ctx, blocks, KSUM_NV, RES_E_KSPACE);
        if (status != 0) return status;
    }

    // ---- forces -------------------------------------------------------------------------------------
    if ((req.forces || req.molecular_virial) && owned > 0) {
        if (req.molecular_virial) {
            status = launch_molecule_com(ctx);
            if (status != 0) return status;
        }
        const int blocks = (owned + KFORCE_THREADS - 1) / KFORCE_THREADS;
        KForceArgs a;
        a.pos = ctx->position.ptr;
        a.charge = ctx->charge.ptr;
        a.a_lo = (int)lo;
        a.a_hi = (int)hi;
        a.kmax = kmax;
        a.nk = nk;
        a.basis = basis;
        a.kindex = ctx->kindex.ptr;
        a.kenergy = ctx->kenergy.ptr;
        a.rho = ctx->rho.ptr;
        a.force = ctx->force.ptr;
        a.do_correction = req.molecular_virial;
        a.mol_of = ctx->mol_of.ptr;
        a.mol_com = ctx->mol_com.ptr;
        a.add_to_force = req.forces;
        LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)blocks * 9));
        a.partials = ctx->partials.ptr;
        const size_t smem = (size_t)3 * (kmax + 1) * KFORCE_THREADS * sizeof(double2) +
                            KFORCE_STAGE * (sizeof(double2) + sizeof(double) + sizeof(short4));
        if (smem > 220 * 1024) {
            return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "kmax = %d needs %zu bytes of shared memory per block", kmax,
                             smem);
        }
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute((const void*)ewald_force_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            ScopedClock clock(ctx, &ctx->clk_kspace);
            ewald_force_kernel<<<blocks, KFORCE_THREADS, smem, ctx->stream>>>(a);
            ctx->launches++;
            ctx->clk_kspace.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        if (req.molecular_virial) {
            status = launch_reduce(ctx, blocks, 9, RES_W_KSPACE_CORRECTION);
            if (status != 0) return status;
        }
    }
    return 0;
}

}  // namespace lumol
