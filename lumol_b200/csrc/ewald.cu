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
#include <algorithm>
#include <cstdlib>

namespace lumol {

// One (h, k) row of the k list for the tiled kernels: its entries are contiguous, l in [l_lo, l_hi].
struct KRow {
    int h, k;
    int base;     // list index of (h, k, l_lo)
    int l_lo_hi;  // l_lo | l_hi << 16 (both as 16-bit signed)
};
static_assert(sizeof(KRow) == sizeof(int4), "KRow is stored in an int4 buffer");

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
    // rows of the tiled kernels: the entries of one (h, k) are contiguous with l increasing by one
    {
        std::vector<KRow> rows;
        bool regular = true;
        for (size_t e = 0; e < index.size(); e++) {
            const short4 v = index[e];
            if (!rows.empty() && rows.back().h == v.x && rows.back().k == v.y) {
                const int l_hi = rows.back().l_lo_hi >> 16;
                if (v.z != l_hi + 1) regular = false;
                rows.back().l_lo_hi = (rows.back().l_lo_hi & 0xffff) | ((int)v.z << 16);
            } else {
                KRow row;
                row.h = v.x;
                row.k = v.y;
                row.base = (int)e;
                row.l_lo_hi = ((int)v.z & 0xffff) | ((int)v.z << 16);
                if (v.x < 0) regular = false;
                rows.push_back(row);
            }
        }
        // The k list is a sphere: row (h, k) only has |l| <= sqrt(kmax^2 - h^2 - k^2).  Rows in order of decreasing reach make
        // the row tiles of the tiled kernels homogeneous, so that a tile's bound on |l| (largest reach among its rows) is
        // tight and never grows from one tile to the next.  Nothing else depends on the order of this table.
        auto reach_of = [](const KRow& row) {
            const int l_lo = (int)(short)(row.l_lo_hi & 0xffff), l_hi = row.l_lo_hi >> 16;
            return std::max(std::abs(l_lo), std::abs(l_hi));
        };
        std::stable_sort(rows.begin(), rows.end(), [&](const KRow& x, const KRow& y) { return reach_of(x) > reach_of(y); });
        ctx->krows_regular = regular && !rows.empty();
        ctx->nkrows = (int64_t)rows.size();
        LUMOL_CUDA_CHECK(ctx, ctx->krows.reserve(rows.size() + 1));
        if (!rows.empty()) {
            LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(reinterpret_cast<KRow*>(ctx->krows.ptr), rows.data(), rows.size() * sizeof(KRow), cudaMemcpyHostToDevice,
                                                  ctx->stream));
            LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
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
    ctx->ewald_table_version++;
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

// ================================================================================================
// Tiled reciprocal space (large N x Nk)
// ================================================================================================
//
// rho(h,k,l) = sum_i q_i e_x(h,i) e_y(k,i) e_z(l,i) is a contraction over atoms of A[(h,k), i] = q_i e_x e_y with
// B[i, l] = e_z(l, i), and the force needs W0[i,(h,k)] = sum_l e_z(l,i) G(h,k,l), W1 = sum_l l e_z(l,i) G(h,k,l),
// G = 2 e_k conj(rho_k): a contraction over l.  Both are register-tiled here like a GEMM, with the phase tables of a
// tile of atoms rebuilt in shared memory by the reference's recursion (ewald.rs:655-662).  e_z(-l) = conj e_z(l)
// makes +l and -l share their four real products, so a complex (atom, k) pair costs 2 DFMA for rho and 4 for the
// forces instead of the 10 + 15 FP64 instructions of the direct kernels above.
//
// The k list is enumerated with l fastest (ewald.rs:144-182): the entries of one (h, k) "row" are contiguous,
// l in [l_lo, l_hi]; rows carry the list index of their first entry.

constexpr int TK_THREADS = 256;
constexpr int TK_ATOMS = 32;  // atoms per shared-memory tile of the rho kernel
constexpr int TK_TR = 4;      // rows per thread: 4 x 4 complex tile = 64 DFMA per 8 LDS.128 (the shared-memory pipe moves
                              // four wavefronts per 128-bit load whatever the broadcast, so 2 x 4 tiles were bound by it)
constexpr int TK_TL = 4;      // |l| values per thread

__device__ __forceinline__ int row_index(const KRow& row, int l) {
    const int l_lo = (short)(row.l_lo_hi & 0xffff), l_hi = row.l_lo_hi >> 16;
    return (l >= l_lo && l <= l_hi) ? row.base + (l - l_lo) : -1;
}

struct TiledRhoArgs {
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    int a_lo, a_hi;
    int nchunks;
    int kmax;
    int nk, nrows;
    int lq;          // thread columns: ceil((kmax + 1) / TK_TL)
    int row_groups;  // thread rows: TK_THREADS / lq
    KBasis basis;
    const KRow* __restrict__ rows;
    double2* __restrict__ rho_partial;  // nchunks x nk
};

// Shared: tables [atom][axis][m] (TK_ATOMS x 3 x lpad complex, lpad = lq * TK_TL >= kmax + 1) and A [atom][row]
// (TK_ATOMS x rows per block): in the main loop every thread of a warp reads the same atom, so rows and |l|
// values must be the fastest index for the reads to be conflict-free.
// grid: (row tiles, atom chunks)
// barrier + maximum of `value` over the block (through one shared word the caller provides)
__device__ __forceinline__ int block_max_barrier(int value, int* shared_word) {
    if (threadIdx.x == 0) *shared_word = 0;
    __syncthreads();
    value = __reduce_max_sync(0xffffffffu, value);
    if ((threadIdx.x & 31) == 0 && value > 0) atomicMax(shared_word, value);
    __syncthreads();
    return *shared_word;
}

// The contraction of one atom tile for a thread whose block needs only the first NL of its TK_TL groups of |l| values
// (rows are sorted by reach: a block's rows all stop below NL * lq).
template <int NL>
__device__ __forceinline__ void rho_contract(double (&acc)[TK_TR][TK_TL][4], const double2* __restrict__ arow,
                                             const double2* __restrict__ bcol, int rows_per_block, int row_groups, int lpad, int lq_count) {
#pragma unroll 2
    for (int atom = 0; atom < TK_ATOMS; atom++) {
        double2 av[TK_TR], bv[NL];
#pragma unroll
        for (int r = 0; r < TK_TR; r++) av[r] = arow[(size_t)atom * rows_per_block + r * row_groups];
#pragma unroll
        for (int l = 0; l < NL; l++) bv[l] = bcol[(size_t)atom * 3 * lpad + l * lq_count];
#pragma unroll
        for (int r = 0; r < TK_TR; r++)
#pragma unroll
            for (int l = 0; l < NL; l++) {
                acc[r][l][0] = fma(av[r].x, bv[l].x, acc[r][l][0]);
                acc[r][l][1] = fma(av[r].y, bv[l].y, acc[r][l][1]);
                acc[r][l][2] = fma(av[r].x, bv[l].y, acc[r][l][2]);
                acc[r][l][3] = fma(av[r].y, bv[l].x, acc[r][l][3]);
            }
    }
}

__global__ void __launch_bounds__(TK_THREADS, 1) ewald_rho_tiled_kernel(TiledRhoArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lpad = a.lq * TK_TL;
    double2* table = reinterpret_cast<double2*>(smem_raw);
    double2* amat = table + (size_t)TK_ATOMS * 3 * lpad;
    const int rows_per_block = a.row_groups * TK_TR;
    __shared__ KRow s_rows[TK_THREADS * 2];  // rows_per_block <= 2 * TK_THREADS (kmax >= 8: at most 85 row groups)

    const int t = threadIdx.x;
    const int row0 = blockIdx.x * rows_per_block;
    __shared__ int s_reach;
    int reach = 0;
    for (int r = t; r < rows_per_block; r += TK_THREADS) {
        KRow row;
        row.h = row.k = 0;
        row.base = 0;
        row.l_lo_hi = (0 << 16) | 1;  // empty interval: l_lo = 1 > l_hi = 0
        if (row0 + r < a.nrows) row = a.rows[row0 + r];
        s_rows[r] = row;
        const int l_lo = (int)(short)(row.l_lo_hi & 0xffff), l_hi = row.l_lo_hi >> 16;
        if (l_lo <= l_hi) reach = max(reach, max(abs(l_lo), abs(l_hi)));
    }
    // groups of |l| values the rows of this block reach: group j holds |l| in [j * lq, (j + 1) * lq)
    const int block_reach = block_max_barrier(reach, &s_reach);  // (barrier: also for blocks whose atom chunk is empty)
    const int ngroups = min(TK_TL, block_reach / a.lq + 1);
    const int lq = t % a.lq, rg = t / a.lq;
    const bool worker = rg < a.row_groups;

    const int owned = a.a_hi - a.a_lo;
    const int per_chunk = (owned + a.nchunks - 1) / a.nchunks;
    const int c_lo = a.a_lo + blockIdx.y * per_chunk;
    const int c_hi = min(a.a_hi, c_lo + per_chunk);

    // P1 = sum ar br, P2 = sum ai bi, P3 = sum ar bi, P4 = sum ai br per (row, |l|)
    double acc[TK_TR][TK_TL][4];
#pragma unroll
    for (int r = 0; r < TK_TR; r++)
#pragma unroll
        for (int l = 0; l < TK_TL; l++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][l][c] = 0.0;

    for (int base = c_lo; base < c_hi; base += TK_ATOMS) {
        const int count = min(TK_ATOMS, c_hi - base);
        __syncthreads();
        // phase tables: 3 * TK_ATOMS independent recursions; missing atoms get zero charge below
        if (t < 3 * TK_ATOMS) {
            const int atom = t % TK_ATOMS, axis = t / TK_ATOMS;
            const int i = base + min(atom, count - 1);
            const double phase = a.basis.b[3 * axis] * a.pos[3 * i] + a.basis.b[3 * axis + 1] * a.pos[3 * i + 1] +
                                 a.basis.b[3 * axis + 2] * a.pos[3 * i + 2];
            double sn, cs;
            sincos(phase, &sn, &cs);
            double2* column = table + (size_t)(atom * 3 + axis) * lpad;
            const double2 e1 = make_double2(cs, sn);
            double2 e = make_double2(1.0, 0.0);
            column[0] = e;
            for (int m = 1; m < lpad; m++) {  // entries beyond kmax are computed but never stored in rho
                e = m == 1 ? e1 : cmul(e, e1);
                column[m] = e;
            }
        }
        __syncthreads();
        // A[atom][row] = q e_x(h) e_y(k); h >= 0 in the half space, e_y(-k) = conj e_y(k)
        int atom = t / rows_per_block, r = t - atom * rows_per_block;
        const int atom_step = TK_THREADS / rows_per_block, r_step = TK_THREADS - atom_step * rows_per_block;
        for (int w = t; w < rows_per_block * TK_ATOMS; w += TK_THREADS) {
            const KRow row = s_rows[r];
            const double2* tables = table + (size_t)atom * 3 * lpad;
            const double2 ex = tables[row.h];
            double2 ey = tables[lpad + abs(row.k)];
            if (row.k < 0) ey.y = -ey.y;
            const double q = atom < count ? a.charge[base + atom] : 0.0;
            const double2 e = cmul(ex, ey);
            amat[w] = make_double2(q * e.x, q * e.y);
            atom += atom_step;
            r += r_step;
            if (r >= rows_per_block) {
                r -= rows_per_block;
                atom++;
            }
        }
        __syncthreads();
        if (worker) {
            // thread (rg, lq) owns rows rg + r * row_groups and |l| = lq + l * a.lq: for a given r or l the threads of
            // a warp read consecutive 16-byte words
            const double2* arow = amat + rg;
            const double2* bcol = table + 2 * lpad + lq;
            switch (ngroups) {
                case 1: rho_contract<1>(acc, arow, bcol, rows_per_block, a.row_groups, lpad, a.lq); break;
                case 2: rho_contract<2>(acc, arow, bcol, rows_per_block, a.row_groups, lpad, a.lq); break;
                case 3: rho_contract<3>(acc, arow, bcol, rows_per_block, a.row_groups, lpad, a.lq); break;
                default: rho_contract<TK_TL>(acc, arow, bcol, rows_per_block, a.row_groups, lpad, a.lq); break;
            }
        }
    }
    if (worker) {
        double2* out = a.rho_partial + (size_t)blockIdx.y * a.nk;
#pragma unroll
        for (int r = 0; r < TK_TR; r++) {
            const KRow row = s_rows[rg + r * a.row_groups];
#pragma unroll
            for (int l = 0; l < TK_TL; l++) {
                const int m = lq + l * a.lq;
                if (m > a.kmax) continue;
                const int plus = row_index(row, m);
                if (plus >= 0) out[plus] = make_double2(acc[r][l][0] - acc[r][l][1], acc[r][l][2] + acc[r][l][3]);
                const int minus = m > 0 ? row_index(row, -m) : -1;
                if (minus >= 0) out[minus] = make_double2(acc[r][l][0] + acc[r][l][1], acc[r][l][3] - acc[r][l][2]);
            }
        }
    }
}

// G matrix of the force kernel: per (row, |l|) four doubles
//   S = g+ + g-, D = g+ - g-, g = 2 e_k conj(rho_k) (0 outside the list; for l = 0 only g+): {S.re, S.im, D.re, D.im}
// (the l-weighted sum W1 uses the same numbers with l e_z(l) instead of e_z(l))
__global__ void __launch_bounds__(256)
    ewald_gmat_kernel(int nrows, int kmax, const KRow* __restrict__ rows, const double2* __restrict__ rho,
                      const double* __restrict__ kenergy, double* __restrict__ gmat) {
    const int lp = kmax + 1;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nrows * lp) return;
    const int r = w / lp, m = w % lp;
    const KRow row = rows[r];
    double2 gp = make_double2(0.0, 0.0), gm = make_double2(0.0, 0.0);
    const int plus = row_index(row, m);
    if (plus >= 0) {
        const double2 v = rho[plus];
        const double f = 2.0 * kenergy[plus];
        gp = make_double2(f * v.x, -f * v.y);
    }
    const int minus = m > 0 ? row_index(row, -m) : -1;
    if (minus >= 0) {
        const double2 v = rho[minus];
        const double f = 2.0 * kenergy[minus];
        gm = make_double2(f * v.x, -f * v.y);
    }
    const double sr = gp.x + gm.x, si = gp.y + gm.y, dr = gp.x - gm.x, di = gp.y - gm.y;
    double* g = gmat + (size_t)w * 4;
    g[0] = sr;
    g[1] = si;
    g[2] = dr;
    g[3] = di;
}

constexpr int TF_TA = 4;          // atoms per thread
constexpr int TF_TR = 4;          // rows per thread: 12 LDS.128 per 136 FP64 instructions and |l| step
constexpr int TF_ROW_GROUPS = 8;  // thread rows
constexpr int TF_ROWS = TF_ROW_GROUPS * TF_TR;  // rows per staged tile: 32
static_assert(TF_ROWS == 32, "the first warp of the tiled force kernel loads one row per lane");

struct TiledForceArgs {
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    int a_lo, a_hi;
    int kmax;
    int nrows;
    int nsplit;  // gridDim.y: the rows are split between blocks of the same atoms
    KBasis basis;
    const KRow* __restrict__ rows;
    const double* __restrict__ gmat;
    double* __restrict__ force;    // nsplit == 1: accumulated directly
    double* __restrict__ partial;  // nsplit > 1: [split][owned atom][3] (sum_h, sum_k, sum_l), reduced afterwards
    double2* xy_scratch;           // XY_GLOBAL: per block, e_x | e_y tables of its atom tile ((kmax + 1) x TF_ATOMS each)
};

// Shared: tables [axis][m][atom] (3 x (kmax + 1) x TF_ATOMS complex) + G tile (TF_ROWS x (kmax + 1) x 4 doubles).
// TF_ATOMS atoms per tile, TF_ATOMS / TF_TA x TF_ROW_GROUPS threads.  XY_GLOBAL (large kmax): only the e_z table,
// which the inner loop streams, stays in shared memory; e_x and e_y, read once per row and atom in the epilogue of a
// row tile, live in a per-block scratch in global memory (L2-resident), so that two blocks of 64 atoms still fit on
// an SM at kmax = 54 (a 1M-atom SPC/E box) where three tables allowed one block of 32 atoms, i.e. two warps per SM.
// A block walks the atom tiles blockIdx.x, blockIdx.x + gridDim.x, ... (one tile per block unless XY_GLOBAL).
template <int TF_ATOMS, bool XY_GLOBAL>
__global__ void __launch_bounds__(TF_ATOMS / TF_TA * TF_ROW_GROUPS) ewald_force_tiled_kernel(TiledForceArgs a) {
    constexpr int TF_THREADS = TF_ATOMS / TF_TA * TF_ROW_GROUPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lp = a.kmax + 1;
    double2* table = reinterpret_cast<double2*>(smem_raw);
    double* gtile = reinterpret_cast<double*>(table + (size_t)(XY_GLOBAL ? 1 : 3) * lp * TF_ATOMS);
    const int gstride = lp * 4 + 4;  // doubles per row of the staged G tile: neighbouring rows in different banks
    __shared__ KRow s_rows[TF_ROWS];
    __shared__ int s_mend;
    // e_x | e_y of this block's atom tile: in shared memory, or in this block's slice of the global scratch (plain
    // loads and stores: written and read by the same block on either side of a barrier)
    double2* xy = XY_GLOBAL ? a.xy_scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * lp * TF_ATOMS : table;
    double2* ez_table = XY_GLOBAL ? table : table + (size_t)2 * lp * TF_ATOMS;

    const int t = threadIdx.x;
    // thread (ag, rg) owns atoms ag + x * AG and rows rg + r * TF_ROW_GROUPS: for a given x the threads of a warp
    // read consecutive 16-byte words of the phase tables
    constexpr int AG = TF_ATOMS / TF_TA;
    const int ag = t % AG, rg = t / AG;
    // rows of this block: [row_lo, row_hi), walked in tiles of TF_ROWS
    const int rows_per_split = ((a.nrows + a.nsplit - 1) / a.nsplit + TF_ROWS - 1) / TF_ROWS * TF_ROWS;
    const int row_lo = blockIdx.y * rows_per_split;
    const int row_hi = min(a.nrows, row_lo + rows_per_split);
    const int natiles = (a.a_hi - a.a_lo + TF_ATOMS - 1) / TF_ATOMS;

    for (int atile = blockIdx.x; atile < natiles; atile += gridDim.x) {
        const int base = a.a_lo + atile * TF_ATOMS;
        const int count = min(TF_ATOMS, a.a_hi - base);
        __syncthreads();  // the sums of the previous atom tile were read from the memory of the tables
        for (int w = t; w < 3 * TF_ATOMS; w += TF_THREADS) {
            const int atom = w % TF_ATOMS, axis = w / TF_ATOMS;
            const int i = base + min(atom, count - 1);
            const double phase = a.basis.b[3 * axis] * a.pos[3 * i] + a.basis.b[3 * axis + 1] * a.pos[3 * i + 1] +
                                 a.basis.b[3 * axis + 2] * a.pos[3 * i + 2];
            double sn, cs;
            sincos(phase, &sn, &cs);
            double2* column = (axis == 2 ? ez_table : xy + (size_t)axis * lp * TF_ATOMS) + atom;
            const double2 e1 = make_double2(cs, sn);
            double2 e = make_double2(1.0, 0.0);
            column[0] = e;
            if (a.kmax >= 1) {
                e = e1;
                column[TF_ATOMS] = e;
            }
            for (int m = 2; m <= a.kmax; m++) {
                e = cmul(e, e1);
                column[(size_t)m * TF_ATOMS] = e;
            }
        }

        double sh[TF_TA], sk[TF_TA], sl[TF_TA];
#pragma unroll
        for (int x = 0; x < TF_TA; x++) sh[x] = sk[x] = sl[x] = 0.0;

        const double2* ez = ez_table + ag;
        int stage_m = lp;  // |l| values of the G tile to stage
        for (int tile = row_lo; tile < row_hi; tile += TF_ROWS) {
            __syncthreads();
            const int nrows = min(TF_ROWS, row_hi - tile);
            {
                // rows come in order of decreasing reach: no row of this tile reaches beyond the bound of the previous tile
                const double* src = a.gmat + (size_t)tile * lp * 4;
                const int columns = stage_m * 4;
                for (int w = t; w < TF_ROWS * columns; w += TF_THREADS) {
                    const int r = w / columns, c = w - r * columns;
                    gtile[r * gstride + c] = r < nrows ? src[(size_t)r * lp * 4 + c] : 0.0;
                }
            }
            if (t < TF_ROWS) {  // exactly the first warp
                KRow row;
                row.h = row.k = row.base = 0;
                row.l_lo_hi = 1;
                int reach = 0;
                if (t < nrows) {
                    row = a.rows[tile + t];
                    const int l_lo = (int)(short)(row.l_lo_hi & 0xffff), l_hi = row.l_lo_hi >> 16;
                    reach = max(abs(l_lo), abs(l_hi));
                }
                s_rows[t] = row;
                // |l| beyond the reach of every row of the tile multiplies zeros of the G matrix: not visited
                reach = __reduce_max_sync(0xffffffffu, reach);
                if (t == 0) s_mend = min(lp, reach + 1);
            }
            __syncthreads();
            const int mend = s_mend;
            stage_m = mend;
            double w0[TF_TR][TF_TA][2], w1[TF_TR][TF_TA][2];
#pragma unroll
            for (int r = 0; r < TF_TR; r++)
#pragma unroll
                for (int x = 0; x < TF_TA; x++) w0[r][x][0] = w0[r][x][1] = w1[r][x][0] = w1[r][x][1] = 0.0;
            const double* g0 = gtile + (size_t)rg * gstride;
            for (int m = 0; m < mend; m++) {
                double2 z[TF_TA], zl[TF_TA];
                const double weight = (double)m;
#pragma unroll
                for (int x = 0; x < TF_TA; x++) {
                    z[x] = ez[(size_t)m * TF_ATOMS + x * AG];
                    zl[x] = make_double2(weight * z[x].x, weight * z[x].y);
                }
#pragma unroll
                for (int r = 0; r < TF_TR; r++) {
                    // S.re, S.im, D.re, D.im
                    const double4 g = *reinterpret_cast<const double4*>(g0 + (size_t)r * TF_ROW_GROUPS * gstride + m * 4);
#pragma unroll
                    for (int x = 0; x < TF_TA; x++) {
                        w0[r][x][0] = fma(z[x].x, g.x, fma(-z[x].y, g.w, w0[r][x][0]));
                        w0[r][x][1] = fma(z[x].x, g.y, fma(z[x].y, g.z, w0[r][x][1]));
                        w1[r][x][0] = fma(zl[x].x, g.z, fma(-zl[x].y, g.y, w1[r][x][0]));
                        w1[r][x][1] = fma(zl[x].x, g.w, fma(zl[x].y, g.x, w1[r][x][1]));
                    }
                }
            }
            // Im(u W) with u = e_x(h) e_y(k)
#pragma unroll
            for (int r = 0; r < TF_TR; r++) {
                const KRow row = s_rows[rg + r * TF_ROW_GROUPS];
#pragma unroll
                for (int x = 0; x < TF_TA; x++) {
                    const double2 ex = xy[(size_t)row.h * TF_ATOMS + ag + x * AG];
                    double2 ey = xy[(size_t)(lp + abs(row.k)) * TF_ATOMS + ag + x * AG];
                    if (row.k < 0) ey.y = -ey.y;
                    const double2 u = cmul(ex, ey);
                    const double t0 = u.x * w0[r][x][1] + u.y * w0[r][x][0];
                    const double t1 = u.x * w1[r][x][1] + u.y * w1[r][x][0];
                    sh[x] = fma(t0, (double)row.h, sh[x]);
                    sk[x] = fma(t0, (double)row.k, sk[x]);
                    sl[x] += t1;
                }
            }
        }
        // sum over the row groups of the block, in a fixed order (the phase tables are dead: reuse their memory)
        __syncthreads();
        double (*s_sum)[TF_ATOMS][3] = reinterpret_cast<double (*)[TF_ATOMS][3]>(smem_raw);
#pragma unroll
        for (int x = 0; x < TF_TA; x++) {
            s_sum[rg][ag + x * AG][0] = sh[x];
            s_sum[rg][ag + x * AG][1] = sk[x];
            s_sum[rg][ag + x * AG][2] = sl[x];
        }
        __syncthreads();
        if (t < count) {
            double th = 0.0, tk = 0.0, tl = 0.0;
#pragma unroll
            for (int g = 0; g < TF_ROW_GROUPS; g++) {
                th += s_sum[g][t][0];
                tk += s_sum[g][t][1];
                tl += s_sum[g][t][2];
            }
            const int i = base + t;
            if (a.nsplit == 1) {
                const double scale = a.charge[i] / FOUR_PI_EPSILON_0;  // ewald.rs:716-718
                a.force[3 * i] += scale * (th * a.basis.b[0] + tk * a.basis.b[3] + tl * a.basis.b[6]);
                a.force[3 * i + 1] += scale * (th * a.basis.b[1] + tk * a.basis.b[4] + tl * a.basis.b[7]);
                a.force[3 * i + 2] += scale * (th * a.basis.b[2] + tk * a.basis.b[5] + tl * a.basis.b[8]);
            } else {
                double* out = a.partial + ((size_t)blockIdx.y * (a.a_hi - a.a_lo) + (i - a.a_lo)) * 3;
                out[0] = th;
                out[1] = tk;
                out[2] = tl;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
    ewald_force_combine_kernel(int a_lo, int a_hi, int nsplit, KBasis basis, const double* __restrict__ charge,
                               const double* __restrict__ partial, double* __restrict__ force) {
    const int i = a_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a_hi) return;
    double th = 0.0, tk = 0.0, tl = 0.0;
    for (int s = 0; s < nsplit; s++) {
        const double* p = partial + ((size_t)s * (a_hi - a_lo) + (i - a_lo)) * 3;
        th += p[0];
        tk += p[1];
        tl += p[2];
    }
    const double scale = charge[i] / FOUR_PI_EPSILON_0;
    force[3 * i] += scale * (th * basis.b[0] + tk * basis.b[3] + tl * basis.b[6]);
    force[3 * i + 1] += scale * (th * basis.b[1] + tk * basis.b[4] + tl * basis.b[7]);
    force[3 * i + 2] += scale * (th * basis.b[2] + tk * basis.b[5] + tl * basis.b[8]);
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

    // ---- which kernels -----------------------------------------------------------------------------
    // tiled (GEMM-like) kernels for large N x Nk, direct ones for the small systems and the molecular virial
    const int lq = (kmax + 1 + TK_TL - 1) / TK_TL;
    const size_t rho_tables = (size_t)3 * (lq * TK_TL) * TK_ATOMS * sizeof(double2);
    const int row_groups = TK_THREADS / lq;
    const size_t rho_smem = rho_tables + (size_t)row_groups * TK_TR * TK_ATOMS * sizeof(double2);
    const size_t force_tables64 = (size_t)3 * (kmax + 1) * 64 * sizeof(double2), gtile = (size_t)TF_ROWS * ((kmax + 1) * 4 + 4) * sizeof(double);
    // force kernel: three phase tables in shared memory (small kmax) or the e_z table alone (XY_GLOBAL)
    const bool force_wide = force_tables64 + gtile <= 110 * 1024;
    const bool tiled_possible = ctx->krows_regular && kmax >= 8 && kmax <= 255 && rho_smem <= 200 * 1024 &&
                                (force_wide || force_tables64 / 3 + gtile + 1024 <= 226 * 1024);
    // the direct force kernel holds the phase tables of 128 atoms: beyond kmax = 34 only the tiled kernels fit
    const size_t direct_force_smem = (size_t)3 * (kmax + 1) * KFORCE_THREADS * sizeof(double2) +
                                     KFORCE_STAGE * (sizeof(double2) + sizeof(double) + sizeof(short4));
    bool tiled = tiled_possible && (ctx->kspace_algorithm == 1 ||
                                    (ctx->kspace_algorithm < 0 && ((int64_t)owned * nk >= (int64_t)1 << 24 || direct_force_smem > 220 * 1024)));
    const bool tiled_forces = tiled && !req.molecular_virial;

    // ---- rho(k) ---------------------------------------------------------------------------------
    int nchunks = 1;
    if (tiled) {
        const int nrows = (int)ctx->nkrows;
        const int rows_per_block = row_groups * TK_TR;
        const int row_tiles = (nrows + rows_per_block - 1) / rows_per_block;
        // one block per SM: a whole number of waves (four) so that no wave runs with a handful of blocks
        nchunks = 4 * ctx->sm_count / row_tiles;
        const int max_chunks = (owned + TK_ATOMS - 1) / TK_ATOMS;
        if (nchunks > max_chunks) nchunks = max_chunks;
        if (nchunks < 1) nchunks = 1;
        if (nchunks > 65535) nchunks = 65535;
        LUMOL_CUDA_CHECK(ctx, ctx->rho_partial.reserve((size_t)nchunks * nk));
        TiledRhoArgs a;
        a.pos = ctx->position.ptr;
        a.charge = ctx->charge.ptr;
        a.a_lo = (int)lo;
        a.a_hi = (int)hi;
        a.nchunks = nchunks;
        a.kmax = kmax;
        a.nk = nk;
        a.nrows = nrows;
        a.lq = lq;
        a.row_groups = row_groups;
        a.basis = basis;
        a.rows = reinterpret_cast<const KRow*>(ctx->krows.ptr);
        a.rho_partial = ctx->rho_partial.ptr;
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute((const void*)ewald_rho_tiled_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rho_smem));
        {
            ScopedClock clock(ctx, &ctx->clk_kspace);
            ewald_rho_tiled_kernel<<<dim3(row_tiles, nchunks), TK_THREADS, rho_smem, ctx->stream>>>(a);
            ctx->launches++;
            ctx->clk_kspace.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
    } else {
        const int kblocks = (nk + RHO_THREADS - 1) / RHO_THREADS;
        // shared-memory tile: as many atoms as fit in ~96 KB, capped at 256
        const size_t per_atom = (size_t)3 * (kmax + 1) * sizeof(double2) + sizeof(double);
        int tile = (int)((96 * 1024) / per_atom);
        if (tile > 256) tile = 256;
        if (tile < 1) {
            return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "kmax = %d needs more shared memory than one atom tile", kmax);
        }
        // enough chunks to fill the GPU a few times over, never more than one chunk per tile
        nchunks = (4 * ctx->sm_count + kblocks - 1) / kblocks;
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
    }
    {
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
        ctx->rho_positions_epoch = ctx->positions_epoch;
        ctx->rho_table_version = ctx->ewald_table_version;
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
    if (tiled_forces && req.forces && owned > 0) {
        const int nrows = (int)ctx->nkrows;
        const int lp = kmax + 1;
        LUMOL_CUDA_CHECK(ctx, ctx->kgmat.reserve((size_t)nrows * lp * 4 + 8));
        ewald_gmat_kernel<<<(nrows * lp + 255) / 256, 256, 0, ctx->stream>>>(nrows, kmax, reinterpret_cast<const KRow*>(ctx->krows.ptr),
                                                                           ctx->rho.ptr, ctx->kenergy.ptr, ctx->kgmat.ptr);
        ctx->launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        // small kmax: the three phase tables of 64 atoms in shared memory, two blocks per SM.  Large kmax: e_z only
        // (e_x, e_y in a global scratch); blocks then walk the atom tiles.
        const bool wide = force_wide;
        const bool xy_global = !wide;
        const int atoms = 64;
        const size_t smem = wide ? force_tables64 + gtile : force_tables64 / 3 + gtile;
        const int atiles = (owned + atoms - 1) / atoms;
        int nsplit = (2 * ctx->sm_count + atiles - 1) / atiles;
        const int max_split = (nrows + TF_ROWS - 1) / TF_ROWS;
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit > 64) nsplit = 64;
        if (nsplit < 1) nsplit = 1;
        int ablocks = atiles;
        if (xy_global && ablocks > 2 * ctx->sm_count) ablocks = 2 * ctx->sm_count;
        TiledForceArgs a;
        a.pos = ctx->position.ptr;
        a.charge = ctx->charge.ptr;
        a.a_lo = (int)lo;
        a.a_hi = (int)hi;
        a.kmax = kmax;
        a.nrows = nrows;
        a.nsplit = nsplit;
        a.basis = basis;
        a.rows = reinterpret_cast<const KRow*>(ctx->krows.ptr);
        a.gmat = ctx->kgmat.ptr;
        a.force = ctx->force.ptr;
        a.partial = nullptr;
        a.xy_scratch = nullptr;
        if (nsplit > 1) {
            LUMOL_CUDA_CHECK(ctx, ctx->kforce_partial.reserve((size_t)nsplit * owned * 3));
            a.partial = ctx->kforce_partial.ptr;
        }
        if (xy_global) {
            LUMOL_CUDA_CHECK(ctx, ctx->kxy_scratch.reserve((size_t)ablocks * nsplit * 2 * lp * atoms));
            a.xy_scratch = ctx->kxy_scratch.ptr;
        }
        const void* kernel = wide ? (const void*)ewald_force_tiled_kernel<64, false> : (const void*)ewald_force_tiled_kernel<64, true>;
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // two blocks of 57 KB tables + 57 KB G tile at kmax = 54 need the whole 228 KB of the SM
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        {
            ScopedClock clock(ctx, &ctx->clk_kspace);
            void* params[] = {&a};
            LUMOL_CUDA_CHECK(ctx, cudaLaunchKernel(kernel, dim3(ablocks, nsplit), dim3(atoms / TF_TA * TF_ROW_GROUPS), params, smem,
                                                   ctx->stream));
            ctx->launches++;
            ctx->clk_kspace.launches++;
        }
        if (nsplit > 1) {
            ewald_force_combine_kernel<<<(owned + 255) / 256, 256, 0, ctx->stream>>>((int)lo, (int)hi, nsplit, basis, ctx->charge.ptr,
                                                                                    ctx->kforce_partial.ptr, ctx->force.ptr);
            ctx->launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
    } else if ((req.forces || req.molecular_virial) && owned > 0) {
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
