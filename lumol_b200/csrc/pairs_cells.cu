// Cell-list pair path: counting sort of the atoms into cells of edge >= the largest cut-off, then one
// block per home cell sweeping its 27 neighbour cells.  The reference has no neighbour search at all
// (SURVEY F1): this replaces the O(N^2) loops of
//   Forces::compute        sys/compute.rs:37-55
//   EnergyEvaluator::pairs sys/energy.rs:47-59
//   AtomicVirial::compute  sys/compute.rs:202-216
//   Ewald real space       energy/global/ewald.rs:430-530
//   Wolf                   energy/global/wolf.rs:177-283
// by an O(N) sweep that visits exactly the same pairs (every pair with r below the cut-off lies in the
// 27-cell neighbourhood when each edge holds at least three cells).
//
// Data layout in HBM: `sorted_pos` holds (x, y, z, q) as one 32-byte double4 per atom, positions wrapped
// into the cell with the floor convention of UnitCell::wrap_vector (cells.rs:263-279), atoms of one cell
// contiguous, cells in z-major order; `sorted_info` holds (kind, first atom of the molecule, bond-distance
// row, original index).  The order inside a cell is by original index, so the sort is deterministic
// (every rank of a multi-GPU run derives the same order).
//
// Pair kernel: the block stages the neighbourhood (27 cells, periodic shifts already applied) in shared
// memory as FP64 double4 plus an FP32 copy relative to the home-cell centre.  One warp owns one atom i at
// a time: the 32 lanes test 32 candidates per iteration in FP32 against a slightly enlarged cut-off,
// compact the survivors into a per-warp queue with ballot/popc, and whenever 32 survivors are queued the
// warp evaluates them densely in FP64 (exact r < rc test included).  The FP64 pipe therefore only sees
// pairs that are (almost all) inside the cut-off, at full lane occupancy.  No atomics: each pair is
// evaluated from both sides, energies/virials are taken from the side with the larger sorted index.
#include "context.hpp"

namespace lumol {

// ------------------------------------------------------------------------------------------------
// neighbour-path choice
// ------------------------------------------------------------------------------------------------

constexpr int CELL_PATH_MIN_ATOMS = 3000;

// Returns 1 when the cell list can (and should) be used, 0 for the all-pairs kernel; fills ctx->ncell.
int choose_neighbor_path(Context* ctx, double cutoff) {
    ctx->ncell[0] = ctx->ncell[1] = ctx->ncell[2] = 0;
    bool possible = ctx->cell.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC && cutoff > 0.0;
    if (possible) {
        const double lengths[3] = {ctx->cell.h[0], ctx->cell.h[4], ctx->cell.h[8]};
        for (int d = 0; d < 3; d++) {
            double nc = floor(lengths[d] / cutoff);
            if (nc > 1024.0) nc = 1024.0;
            ctx->ncell[d] = (int)nc;
            if (ctx->ncell[d] < 3) possible = false;
        }
    }
    if (ctx->forced_path == 0) return 0;
    if (ctx->forced_path == 1) return possible ? 1 : -1;
    return possible && ctx->n >= CELL_PATH_MIN_ATOMS ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// counting sort
// ------------------------------------------------------------------------------------------------

struct GridView {
    int nc[3];
    double length[3];
    double edge[3];
};

__device__ __forceinline__ double wrap_coordinate(double x, double length) {
    // UnitCell::wrap_vector, orthorhombic branch (cells.rs:266-270)
    return x - floor(x / length) * length;
}

__device__ __forceinline__ int cell_coordinate(double wrapped, double length, int nc) {
    int c = (int)(wrapped / length * (double)nc);
    if (c >= nc) c = nc - 1;  // wrapped == length after rounding
    if (c < 0) c = 0;
    return c;
}

__global__ void __launch_bounds__(256)
    cell_assign_kernel(int n, GridView g, const double* __restrict__ pos, int* __restrict__ cell_of,
                       int* __restrict__ slot_of, int* __restrict__ cell_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_coordinate(wrap_coordinate(pos[3 * i], g.length[0]), g.length[0], g.nc[0]);
    const int cy = cell_coordinate(wrap_coordinate(pos[3 * i + 1], g.length[1]), g.length[1], g.nc[1]);
    const int cz = cell_coordinate(wrap_coordinate(pos[3 * i + 2], g.length[2]), g.length[2], g.nc[2]);
    const int c = (cz * g.nc[1] + cy) * g.nc[0] + cx;
    cell_of[i] = c;
    slot_of[i] = atomicAdd(cell_count + c, 1);
}

// exclusive scan of `count` ints in three passes (block scan, scan of block sums, add back)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int value, int* shared, int& total) {
    // inclusive warp scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v = value;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) shared[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (blockDim.x >> 5) ? shared[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        shared[lane] = w;  // inclusive scan of warp totals
    }
    __syncthreads();
    const int warp_offset = warp == 0 ? 0 : shared[warp - 1];
    total = shared[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warp_offset + v - value;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    scan_blocks_kernel(int count, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ block_sums) {
    __shared__ int shared[32];
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    int items[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        items[k] = base + k < count ? in[base + k] : 0;
        sum += items[k];
    }
    int total;
    int offset = block_exclusive_scan(sum, shared, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < count) out[base + k] = offset;
        offset += items[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int nblocks, int* __restrict__ block_sums) {
    __shared__ int shared[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int idx = base + threadIdx.x;
        const int value = idx < nblocks ? block_sums[idx] : 0;
        int total;
        const int offset = block_exclusive_scan(value, shared, total);
        if (idx < nblocks) block_sums[idx] = carry + offset;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
    scan_add_kernel(int count, int* __restrict__ out, const int* __restrict__ block_sums, int total_count) {
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    const int add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < count) out[base + k] += add;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[count] = total_count;
}

// first pass of the scatter: original indices grouped by cell, arrival order
__global__ void __launch_bounds__(256)
    cell_group_kernel(int n, const int* __restrict__ cell_of, const int* __restrict__ slot_of,
                      const int* __restrict__ cell_start, int* __restrict__ grouped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    grouped[cell_start[cell_of[i]] + slot_of[i]] = i;
}

struct ScatterArgs {
    int n;
    GridView g;
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    const unsigned* __restrict__ kind;
    const int* __restrict__ mol_first;
    const int* __restrict__ bd_row;
    const int* __restrict__ cell_of;
    const int* __restrict__ cell_start;
    const int* __restrict__ grouped;
    int* __restrict__ order;
    double4* __restrict__ sorted_pos;
    int4* __restrict__ sorted_info;
};

// second pass: rank inside the cell = number of cell mates with a smaller original index
__global__ void __launch_bounds__(256) cell_scatter_kernel(ScatterArgs a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    const int i = a.grouped[s];
    const int c = a.cell_of[i];
    const int lo = a.cell_start[c], hi = a.cell_start[c + 1];
    int rank = 0;
    for (int t = lo; t < hi; t++) {
        rank += a.grouped[t] < i ? 1 : 0;
    }
    const int dst = lo + rank;
    a.order[dst] = i;
    a.sorted_pos[dst] = make_double4(wrap_coordinate(a.pos[3 * i], a.g.length[0]),
                                     wrap_coordinate(a.pos[3 * i + 1], a.g.length[1]),
                                     wrap_coordinate(a.pos[3 * i + 2], a.g.length[2]), a.charge[i]);
    a.sorted_info[dst] = make_int4((int)a.kind[i], a.mol_first[i], a.bd_row[i], i);
}

static int build_cells(Context* ctx, const GridView& g) {
    const int n = (int)ctx->n;
    const int ncells = g.nc[0] * g.nc[1] * g.nc[2];
    LUMOL_CUDA_CHECK(ctx, ctx->cell_of.reserve((size_t)2 * n));  // cell_of + slot_of
    LUMOL_CUDA_CHECK(ctx, ctx->cell_count.reserve((size_t)ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_start.reserve((size_t)ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->order.reserve((size_t)2 * n));  // order + grouped
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_pos.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_info.reserve((size_t)n));
    const int scan_blocks = (ncells + SCAN_BLOCK - 1) / SCAN_BLOCK;
    LUMOL_CUDA_CHECK(ctx, ctx->scan_scratch.reserve((size_t)scan_blocks + 1));

    int* cell_of = ctx->cell_of.ptr;
    int* slot_of = ctx->cell_of.ptr + n;
    int* order = ctx->order.ptr;
    int* grouped = ctx->order.ptr + n;

    ScopedClock clock(ctx, &ctx->clk_neighbor);
    LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->cell_count.ptr, 0, ((size_t)ncells + 1) * sizeof(int), ctx->stream));
    const int blocks = (n + 255) / 256;
    cell_assign_kernel<<<blocks, 256, 0, ctx->stream>>>(n, g, ctx->position.ptr, cell_of, slot_of, ctx->cell_count.ptr);
    scan_blocks_kernel<<<scan_blocks, SCAN_THREADS, 0, ctx->stream>>>(ncells, ctx->cell_count.ptr, ctx->cell_start.ptr,
                                                                      ctx->scan_scratch.ptr);
    scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(scan_blocks, ctx->scan_scratch.ptr);
    scan_add_kernel<<<scan_blocks, SCAN_THREADS, 0, ctx->stream>>>(ncells, ctx->cell_start.ptr, ctx->scan_scratch.ptr, n);
    cell_group_kernel<<<blocks, 256, 0, ctx->stream>>>(n, cell_of, slot_of, ctx->cell_start.ptr, grouped);
    ScatterArgs s;
    s.n = n;
    s.g = g;
    s.pos = ctx->position.ptr;
    s.charge = ctx->charge.ptr;
    s.kind = ctx->kind.ptr;
    s.mol_first = ctx->mol_first.ptr;
    s.bd_row = ctx->bd_row.ptr;
    s.cell_of = cell_of;
    s.cell_start = ctx->cell_start.ptr;
    s.grouped = grouped;
    s.order = order;
    s.sorted_pos = ctx->sorted_pos.ptr;
    s.sorted_info = ctx->sorted_info.ptr;
    cell_scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(s);
    ctx->launches += 6;
    ctx->clk_neighbor.launches += 6;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pair kernel
// ------------------------------------------------------------------------------------------------

constexpr int CELL_THREADS = 128;
constexpr int CELL_WARPS = CELL_THREADS / 32;
constexpr int CELL_ICHUNK = 128;  // home atoms whose forces are accumulated in shared memory at a time
constexpr int CELL_NV = 16;       // same layout as the all-pairs kernel
constexpr int CELL_MODE_FORCES = 0;
constexpr int CELL_MODE_FULL = 1;

struct CellArgs {
    GridView g;
    int cell_lo, cell_hi;  // home cells handled by this launch
    int o_lo, o_hi;        // original-index range of the atoms this rank owns (forces are computed for those)
    int tile;              // candidates staged per pass
    const int* __restrict__ cell_start;
    const double4* __restrict__ sorted_pos;
    const int4* __restrict__ sorted_info;
    const unsigned char* __restrict__ bond_dist;
    int nkinds;
    const PairParams* __restrict__ pairs;
    const TableDesc* __restrict__ tables;
    const double* __restrict__ table_energy;
    const double* __restrict__ table_force;
    CoulombView coulomb;
    int do_pairs, do_coulomb;
    double cutoff2;     // (largest cut-off)^2, exact FP64 gate for the general path
    float cutoff2_f32;  // the same, enlarged by 1e-4 relative for the FP32 pre-filter
    // LJ fast path
    double lj_sigma2, lj_epsilon, lj_cutoff2, lj_shift;
    int write_forces;            // energy-only queries must not clobber the forces the integrator holds
    double* __restrict__ force;  // original order, n x 3
    double* __restrict__ partials;
};

template <bool LJ_ONLY, int MODE>
__device__ __forceinline__ void evaluate_candidate(const CellArgs& a, const PairParams* __restrict__ sp,
                                                   const double4& pi, const int4& info_i, int s_i, const double4& pj,
                                                   const int4& info_j, int s_j, double& fx, double& fy, double& fz,
                                                   double (&acc)[CELL_NV]) {
    const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const bool count = s_j > s_i;
    if (LJ_ONLY) {
        if (r2 < a.lj_cutoff2) {
            const double rinv2 = 1.0 / r2;
            const double s2 = a.lj_sigma2 * rinv2;
            const double s6 = s2 * s2 * s2;
            // force(r) / r = -24 eps (s6 - 2 s6^2) / r^2 (functions.rs:85-88)
            const double fr = 24.0 * a.lj_epsilon * (2.0 * s6 * s6 - s6) * rinv2;
            fx += fr * dx;
            fy += fr * dy;
            fz += fr * dz;
            if (MODE == CELL_MODE_FULL && count) {
                acc[0] += 4.0 * a.lj_epsilon * (s6 * s6 - s6) - a.lj_shift;
                acc[14] += 1.0;
                acc[2] += fr * dx * dx;
                acc[3] += fr * dx * dy;
                acc[4] += fr * dx * dz;
                acc[5] += fr * dy * dy;
                acc[6] += fr * dy * dz;
                acc[7] += fr * dz * dz;
            }
        }
        return;
    }
    if (!(r2 < a.cutoff2 * 1.0000000001)) return;
    const double r = sqrt(r2);
    const bool same_molecule = info_i.y == info_j.y;
    const unsigned bits = same_molecule ? a.bond_dist[info_i.z + (info_j.w - info_j.y)] : 0u;
    if (a.do_pairs) {
        const PairParams& pp = sp[info_i.x * a.nkinds + info_j.x];
        if (pp.potential > LUMOL_CUDA_POTENTIAL_NULL && r < pp.cutoff) {
            double scaling;
            if (!restriction_excluded(pp.restriction, bits, pp.scale14, scaling)) {
                double e, f;
                pair_eval(pp, a.tables, a.table_energy, a.table_force, r, e, f);
                const double fr = scaling * f / r;
                fx += fr * dx;
                fy += fr * dy;
                fz += fr * dz;
                if (MODE == CELL_MODE_FULL && count) {
                    acc[0] += scaling * e;
                    acc[14] += 1.0;
                    acc[2] += fr * dx * dx;
                    acc[3] += fr * dx * dy;
                    acc[4] += fr * dx * dz;
                    acc[5] += fr * dy * dy;
                    acc[6] += fr * dy * dz;
                    acc[7] += fr * dz * dz;
                }
            }
        }
    }
    if (a.do_coulomb && r <= a.coulomb.rc) {
        const double qi = pi.w, qj = pj.w;
        if (qi != 0.0 && qj != 0.0) {
            double scaling;
            const bool excluded = restriction_excluded(a.coulomb.restriction, bits, a.coulomb.scale14, scaling);
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
                fx += fr * dx;
                fy += fr * dy;
                fz += fr * dz;
                if (MODE == CELL_MODE_FULL && count) {
                    acc[1] += e;
                    acc[15] += 1.0;
                    acc[8] += fr * dx * dx;
                    acc[9] += fr * dx * dy;
                    acc[10] += fr * dx * dz;
                    acc[11] += fr * dy * dy;
                    acc[12] += fr * dy * dz;
                    acc[13] += fr * dz * dz;
                }
            }
        }
    }
}

template <bool LJ_ONLY, int MODE>
__global__ void __launch_bounds__(CELL_THREADS) cell_pairs_kernel(CellArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4* tpos = reinterpret_cast<double4*>(smem_raw);
    float4* tf32 = reinterpret_cast<float4*>(tpos + a.tile);
    int4* tinfo = reinterpret_cast<int4*>(tf32 + a.tile);
    PairParams* sp = reinterpret_cast<PairParams*>(tinfo + (LJ_ONLY ? 0 : a.tile));

    __shared__ int r_start[27];
    __shared__ int r_prefix[28];
    __shared__ double r_shift[27][3];
    __shared__ double sh_force[CELL_ICHUNK * 3];
    __shared__ unsigned short queue[CELL_WARPS][64];
    __shared__ double scratch[32 * CELL_NV];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c = a.cell_lo + blockIdx.x;
    const int cx = c % a.g.nc[0];
    const int cy = (c / a.g.nc[0]) % a.g.nc[1];
    const int cz = c / (a.g.nc[0] * a.g.nc[1]);

    if (!LJ_ONLY) {
        const int words = (int)(sizeof(PairParams) / sizeof(double)) * a.nkinds * a.nkinds;
        const double* src = reinterpret_cast<const double*>(a.pairs);
        double* dst = reinterpret_cast<double*>(sp);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    if (threadIdx.x < 27) {
        const int t = threadIdx.x;
        int nx = cx + (t % 3) - 1, ny = cy + ((t / 3) % 3) - 1, nz = cz + (t / 9) - 1;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        if (nx < 0) { nx += a.g.nc[0]; sx = -a.g.length[0]; }
        if (nx >= a.g.nc[0]) { nx -= a.g.nc[0]; sx = a.g.length[0]; }
        if (ny < 0) { ny += a.g.nc[1]; sy = -a.g.length[1]; }
        if (ny >= a.g.nc[1]) { ny -= a.g.nc[1]; sy = a.g.length[1]; }
        if (nz < 0) { nz += a.g.nc[2]; sz = -a.g.length[2]; }
        if (nz >= a.g.nc[2]) { nz -= a.g.nc[2]; sz = a.g.length[2]; }
        const int nb = (nz * a.g.nc[1] + ny) * a.g.nc[0] + nx;
        r_start[t] = a.cell_start[nb];
        r_prefix[t + 1] = a.cell_start[nb + 1] - a.cell_start[nb];  // length for now
        r_shift[t][0] = sx;
        r_shift[t][1] = sy;
        r_shift[t][2] = sz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        r_prefix[0] = 0;
        for (int t = 0; t < 27; t++) r_prefix[t + 1] += r_prefix[t];
    }
    __syncthreads();
    const int total = r_prefix[27];

    // home atoms of this block; only those this rank owns are evaluated
    const int hs = a.cell_start[c], he = a.cell_start[c + 1];
    {
        int owned = 0;
        for (int s = hs + threadIdx.x; s < he; s += blockDim.x) {
            const int orig = a.sorted_info[s].w;
            owned |= (orig >= a.o_lo && orig < a.o_hi) ? 1 : 0;
        }
        if (!__syncthreads_or(owned)) {
            if (MODE == CELL_MODE_FULL && threadIdx.x == 0) {
                for (int k = 0; k < CELL_NV; k++) a.partials[(size_t)blockIdx.x * CELL_NV + k] = 0.0;
            }
            return;
        }
    }

    // centre of the home cell: origin of the FP32 coordinates
    const double ox = ((double)cx + 0.5) * a.g.edge[0];
    const double oy = ((double)cy + 0.5) * a.g.edge[1];
    const double oz = ((double)cz + 0.5) * a.g.edge[2];

    double acc[CELL_NV];
#pragma unroll
    for (int k = 0; k < CELL_NV; k++) acc[k] = 0.0;

    for (int ic = hs; ic < he; ic += CELL_ICHUNK) {
        const int ni = min(CELL_ICHUNK, he - ic);
        for (int t = threadIdx.x; t < 3 * ni; t += blockDim.x) sh_force[t] = 0.0;

        for (int base = 0; base < total; base += a.tile) {
            const int cnt = min(a.tile, total - base);
            __syncthreads();
            // stage candidates [base, base + cnt) of the flattened neighbourhood
            for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
                const int gidx = base + t;
                int r = 0;
                while (gidx >= r_prefix[r + 1]) r++;
                const int s = r_start[r] + (gidx - r_prefix[r]);
                double4 p = a.sorted_pos[s];
                p.x += r_shift[r][0];
                p.y += r_shift[r][1];
                p.z += r_shift[r][2];
                tpos[t] = p;
                tf32[t] = make_float4((float)(p.x - ox), (float)(p.y - oy), (float)(p.z - oz), __int_as_float(s));
                if (!LJ_ONLY) tinfo[t] = a.sorted_info[s];
            }
            __syncthreads();

            for (int il = warp; il < ni; il += CELL_WARPS) {
                const int s_i = ic + il;
                const int4 info_i = a.sorted_info[s_i];
                if (info_i.w < a.o_lo || info_i.w >= a.o_hi) continue;  // warp-uniform
                const double4 pi = a.sorted_pos[s_i];
                const float fxi = (float)(pi.x - ox), fyi = (float)(pi.y - oy), fzi = (float)(pi.z - oz);
                double fx = 0.0, fy = 0.0, fz = 0.0;
                int queued = 0;
                unsigned short* q = queue[warp];

                for (int b = 0; b < cnt; b += 32) {
                    const int t = b + lane;
                    bool pass = false;
                    if (t < cnt) {
                        const float4 f = tf32[t];
                        const float ddx = fxi - f.x, ddy = fyi - f.y, ddz = fzi - f.z;
                        const float r2f = ddx * ddx + ddy * ddy + ddz * ddz;
                        pass = r2f < a.cutoff2_f32 && __float_as_int(f.w) != s_i;
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, pass);
                    if (pass) q[queued + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)t;
                    queued += __popc(mask);
                    __syncwarp();
                    if (queued >= 32) {
                        const int t2 = q[lane];
                        const int4 info_j = LJ_ONLY ? make_int4(0, 0, 0, 0) : tinfo[t2];
                        evaluate_candidate<LJ_ONLY, MODE>(a, sp, pi, info_i, s_i, tpos[t2], info_j,
                                                          __float_as_int(tf32[t2].w), fx, fy, fz, acc);
                        const int rest = queued - 32;
                        const unsigned short carry = lane < rest ? q[32 + lane] : (unsigned short)0;
                        __syncwarp();
                        if (lane < rest) q[lane] = carry;
                        queued = rest;
                        __syncwarp();
                    }
                }
                if (lane < queued) {
                    const int t2 = q[lane];
                    const int4 info_j = LJ_ONLY ? make_int4(0, 0, 0, 0) : tinfo[t2];
                    evaluate_candidate<LJ_ONLY, MODE>(a, sp, pi, info_i, s_i, tpos[t2], info_j,
                                                      __float_as_int(tf32[t2].w), fx, fy, fz, acc);
                }
                __syncwarp();
                fx = warp_sum(fx);
                fy = warp_sum(fy);
                fz = warp_sum(fz);
                if (lane == 0) {
                    sh_force[3 * il] += fx;
                    sh_force[3 * il + 1] += fy;
                    sh_force[3 * il + 2] += fz;
                }
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < ni; t += blockDim.x) {
            const int orig = a.sorted_info[ic + t].w;
            if (orig < a.o_lo || orig >= a.o_hi || !a.write_forces) continue;
            a.force[3 * orig] = sh_force[3 * t];
            a.force[3 * orig + 1] = sh_force[3 * t + 1];
            a.force[3 * orig + 2] = sh_force[3 * t + 2];
        }
        __syncthreads();
    }

    if (MODE == CELL_MODE_FULL) {
        block_sum<CELL_NV>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < CELL_NV; k++) a.partials[(size_t)blockIdx.x * CELL_NV + k] = acc[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------

int launch_pairs_cells(Context* ctx, const ComputeRequest& req) {
    GridView g;
    for (int d = 0; d < 3; d++) {
        g.nc[d] = ctx->ncell[d];
        g.length[d] = ctx->cell.h[4 * d];
        g.edge[d] = g.length[d] / (double)g.nc[d];
    }
    int status = build_cells(ctx, g);
    if (status != 0) return status;

    const int ncells = g.nc[0] * g.nc[1] * g.nc[2];
    int64_t o_lo, o_hi;
    ctx->owned_range(ctx->n, o_lo, o_hi);

    const bool do_pairs = req.pairs && ctx->any_pair;
    const bool do_coulomb = req.coulomb && ctx->coulomb.kind != 0;
    double cutoff = 0.0;
    if (do_pairs) cutoff = ctx->max_pair_cutoff;
    if (do_coulomb && ctx->coulomb.rc > cutoff) cutoff = ctx->coulomb.rc;

    CellArgs a;
    a.g = g;
    a.cell_lo = 0;
    a.cell_hi = ncells;
    a.o_lo = (int)o_lo;
    a.o_hi = (int)o_hi;
    a.cell_start = ctx->cell_start.ptr;
    a.sorted_pos = ctx->sorted_pos.ptr;
    a.sorted_info = ctx->sorted_info.ptr;
    a.bond_dist = ctx->bond_dist.ptr;
    a.nkinds = ctx->nkinds;
    a.pairs = ctx->pairs.ptr;
    a.tables = ctx->tables.ptr;
    a.table_energy = ctx->table_energy.ptr;
    a.table_force = ctx->table_force.ptr;
    a.coulomb = ctx->coulomb;
    a.do_pairs = do_pairs;
    a.do_coulomb = do_coulomb;
    a.cutoff2 = cutoff * cutoff;
    a.cutoff2_f32 = (float)(cutoff * cutoff * 1.0001);
    a.force = ctx->force.ptr;
    a.write_forces = req.forces;

    const bool lj_only = do_pairs && !do_coulomb && ctx->single_lj;
    if (lj_only) {
        const lumol_cuda_pair& p = ctx->host_pairs[0];
        a.lj_sigma2 = p.p[0] * p.p[0];
        a.lj_epsilon = p.p[1];
        a.lj_cutoff2 = p.cutoff * p.cutoff;
        a.lj_shift = p.shift;
    } else {
        a.lj_sigma2 = a.lj_epsilon = a.lj_cutoff2 = a.lj_shift = 0.0;
    }

    // candidates staged per pass: the mean neighbourhood plus a quarter, a multiple of 32
    const double mean = 27.0 * (double)ctx->n / (double)ncells;
    int tile = (int)(mean * 1.25) + 32;
    tile = (tile + 31) / 32 * 32;
    if (tile < 128) tile = 128;
    const size_t per_entry = sizeof(double4) + sizeof(float4) + (lj_only ? 0 : sizeof(int4));
    const size_t table_bytes = lj_only ? 0 : sizeof(PairParams) * (size_t)ctx->nkinds * ctx->nkinds;
    const size_t budget = 64 * 1024;
    if (table_bytes > 100 * 1024) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "too many particle kinds (%d) for the shared pair table",
                         ctx->nkinds);
    }
    if ((size_t)tile * per_entry > budget) tile = (int)(budget / per_entry) / 32 * 32;
    if (tile > 16384) tile = 16384;  // queue entries are 16-bit
    a.tile = tile;
    const size_t smem = (size_t)tile * per_entry + table_bytes;

    const bool full = req.energy || req.virial;
    const int blocks = ncells;
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)blocks * CELL_NV));
    a.partials = ctx->partials.ptr;

    const void* kernel;
    if (lj_only) {
        kernel = full ? (const void*)cell_pairs_kernel<true, CELL_MODE_FULL>
                      : (const void*)cell_pairs_kernel<true, CELL_MODE_FORCES>;
    } else {
        kernel = full ? (const void*)cell_pairs_kernel<false, CELL_MODE_FULL>
                      : (const void*)cell_pairs_kernel<false, CELL_MODE_FORCES>;
    }
    LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ScopedClock clock(ctx, &ctx->clk_pair);
        void* params[] = {&a};
        LUMOL_CUDA_CHECK(ctx, cudaLaunchKernel(kernel, dim3(blocks), dim3(CELL_THREADS), params, smem, ctx->stream));
        ctx->launches++;
        ctx->clk_pair.launches++;
    }
    if (full) {
        status = launch_reduce(ctx, blocks, CELL_NV, RES_E_PAIRS);
        if (status != 0) return status;
    }
    return 0;
}

}  // namespace lumol
