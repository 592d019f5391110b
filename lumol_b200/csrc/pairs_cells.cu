// Cell-list pair path: counting sort of the atoms into cells of edge >= the largest cut-off, then a sweep of
// the 27 neighbour cells of every home cell.  The reference has no neighbour search at all: this
// replaces the O(N^2) loops of
//   Forces::compute        sys/compute.rs:37-55
//   EnergyEvaluator::pairs sys/energy.rs:47-59
//   AtomicVirial::compute  sys/compute.rs:202-216
//   Ewald real space       energy/global/ewald.rs:430-530
//   Wolf                   energy/global/wolf.rs:177-283
// by an O(N) sweep that visits exactly the same pairs (every pair with r below the cut-off lies in the
// 27-cell neighbourhood when each edge holds at least three cells).
//
// Data layout in HBM (rebuilt every evaluation from the positions in original order):
//   sorted_pos  double4 (x, y, z, q): position RELATIVE TO THE CENTRE OF THE ATOM'S OWN CELL (the wrap into the
//               cell uses the floor convention of UnitCell::wrap_vector, cells.rs:263-279) and the charge;
//   sorted_f32  float4: the same relative position in FP32, for the pre-filter;
//   sorted_info int4 (kind, first atom of the molecule, bond-distance row, original index);
//   cell_start  exclusive scan of the cell populations; cells in z-major order, atoms of one cell contiguous and
//               ordered by original index, so the sort is deterministic (every rank derives the same order).
// With cell-relative coordinates the separation of atom i in home cell h and atom j in the neighbour cell
// h + (a, b, c) is (x_i - (a, b, c) * edge) - x_j whatever the periodic wrap of the cell index, so the kernel
// needs no image logic at all and keeps full relative precision in arbitrarily large boxes.
//
// Pair kernel: one lane per atom i, one warp per home cell (cells above 32 atoms take several passes), each warp
// walking its own contiguous range of cells so that successive home cells reuse their neighbourhoods from L1.
//   phase 1 (FP32 / integer pipes): the warp streams the candidates of the 27 neighbour cells; all lanes read
//     the same candidate (a broadcast load), each lane tests it against its own atom with an enlarged FP32
//     cut-off and appends survivors to its private queue in shared memory;
//   phase 2 (FP64 pipe): every lane walks its queue: gathers the FP64 position, applies the exact r < rc test
//     and evaluates the pair.  The queue is flushed whenever a lane could overflow.
// Forces are accumulated in registers by the lane that owns the atom: no atomics, no shuffles; each pair is
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
    float4* __restrict__ sorted_f32;
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
    // position relative to the centre of the cell the atom was binned into
    const int cx = c % a.g.nc[0], cy = (c / a.g.nc[0]) % a.g.nc[1], cz = c / (a.g.nc[0] * a.g.nc[1]);
    const double x = wrap_coordinate(a.pos[3 * i], a.g.length[0]) - ((double)cx + 0.5) * a.g.edge[0];
    const double y = wrap_coordinate(a.pos[3 * i + 1], a.g.length[1]) - ((double)cy + 0.5) * a.g.edge[1];
    const double z = wrap_coordinate(a.pos[3 * i + 2], a.g.length[2]) - ((double)cz + 0.5) * a.g.edge[2];
    a.sorted_pos[dst] = make_double4(x, y, z, a.charge[i]);
    a.sorted_f32[dst] = make_float4((float)x, (float)y, (float)z, 0.0f);
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
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_f32.reserve((size_t)n));
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
    s.sorted_f32 = ctx->sorted_f32.ptr;
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
constexpr int CELL_QUEUE = 96;   // queue slots per lane
constexpr int CELL_CHUNK = 16;   // candidates streamed between two overflow checks
constexpr int CELL_NV = 16;      // same layout as the all-pairs kernel
constexpr int CELL_MODE_FORCES = 0;
constexpr int CELL_MODE_FULL = 1;
constexpr unsigned QUEUE_INDEX_MASK = (1u << 26) - 1u;

struct CellArgs {
    GridView g;
    int ncells;
    int cells_per_warp;
    int o_lo, o_hi;  // original-index range of the atoms this rank owns (forces are computed for those)
    const int* __restrict__ cell_start;
    const double4* __restrict__ sorted_pos;
    const float4* __restrict__ sorted_f32;
    const int4* __restrict__ sorted_info;
    const unsigned char* __restrict__ bond_dist;
    int nkinds;
    const PairParams* __restrict__ pairs;
    const TableDesc* __restrict__ tables;
    const double* __restrict__ table_energy;
    const double* __restrict__ table_force;
    CoulombView coulomb;
    int do_pairs, do_coulomb;
    double cutoff2;     // (largest cut-off)^2: early-out of the general path
    float cutoff2_f32;  // the same, enlarged by 1e-4 relative for the FP32 pre-filter
    // LJ fast path
    double lj_sigma2, lj_epsilon24, lj_epsilon4, lj_cutoff2, lj_shift;
    int write_forces;            // energy-only queries must not clobber the forces the integrator holds
    double* __restrict__ force;  // original order, n x 3
    double* __restrict__ partials;
};

// One queued candidate, FP64: exact cut-off test and pair evaluation for the lane's atom.
template <bool LJ_ONLY, int MODE>
__device__ __forceinline__ void evaluate_candidate(const CellArgs& a, const PairParams* __restrict__ sp,
                                                   double xi, double yi, double zi, double qi, const int4& info_i,
                                                   int s_i, int s_j, double& fx, double& fy, double& fz,
                                                   double (&acc)[CELL_NV]) {
    const double4 pj = a.sorted_pos[s_j];
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const bool count = s_j > s_i;
    if (LJ_ONLY) {
        if (r2 < a.lj_cutoff2) {
            const double rinv2 = 1.0 / r2;
            const double s2 = a.lj_sigma2 * rinv2;
            const double s6 = s2 * s2 * s2;
            // force(r) / r = -24 eps (s6 - 2 s6^2) / r^2 (functions.rs:85-88)
            const double fr = a.lj_epsilon24 * s6 * (2.0 * s6 - 1.0) * rinv2;
            fx += fr * dx;
            fy += fr * dy;
            fz += fr * dz;
            if (MODE == CELL_MODE_FULL && count) {
                acc[0] += a.lj_epsilon4 * (s6 * s6 - s6) - a.lj_shift;
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
    const int4 info_j = a.sorted_info[s_j];
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
        const double qj = pj.w;
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
    unsigned* queues = reinterpret_cast<unsigned*>(smem_raw);  // [warp][slot][lane]
    PairParams* sp = reinterpret_cast<PairParams*>(queues + CELL_WARPS * CELL_QUEUE * 32);
    __shared__ double offset64[27][3];  // (a, b, c) * edge for the 27 neighbour offsets
    __shared__ float offset32[27][3];
    __shared__ double scratch[32 * CELL_NV];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    if (!LJ_ONLY) {
        const int words = (int)(sizeof(PairParams) / sizeof(double)) * a.nkinds * a.nkinds;
        const double* src = reinterpret_cast<const double*>(a.pairs);
        double* dst = reinterpret_cast<double*>(sp);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    if (threadIdx.x < 27) {
        const int t = threadIdx.x;
        const double ox = (double)((t % 3) - 1) * a.g.edge[0];
        const double oy = (double)(((t / 3) % 3) - 1) * a.g.edge[1];
        const double oz = (double)((t / 9) - 1) * a.g.edge[2];
        offset64[t][0] = ox;
        offset64[t][1] = oy;
        offset64[t][2] = oz;
        offset32[t][0] = (float)ox;
        offset32[t][1] = (float)oy;
        offset32[t][2] = (float)oz;
    }
    __syncthreads();

    unsigned* q = queues + warp * CELL_QUEUE * 32 + lane;  // slot k at q[k * 32]

    double acc[CELL_NV];
#pragma unroll
    for (int k = 0; k < CELL_NV; k++) acc[k] = 0.0;

    const int global_warp = blockIdx.x * CELL_WARPS + warp;
    const int cell_lo = global_warp * a.cells_per_warp;
    const int cell_hi = min(a.ncells, cell_lo + a.cells_per_warp);

    for (int c = cell_lo; c < cell_hi; c++) {
        const int hs = a.cell_start[c], he = a.cell_start[c + 1];
        if (hs == he) continue;
        const int cx = c % a.g.nc[0];
        const int cy = (c / a.g.nc[0]) % a.g.nc[1];
        const int cz = c / (a.g.nc[0] * a.g.nc[1]);

        for (int base = hs; base < he; base += 32) {
            // this lane's atom (lanes without an atom, or with an atom another rank owns, sit far away)
            const int s_i = base + lane;
            bool active = s_i < he;
            int4 info_i = make_int4(0, 0, 0, -1);
            double xi = 0.0, yi = 0.0, zi = 0.0, qi = 0.0;
            if (active) {
                info_i = a.sorted_info[s_i];
                active = info_i.w >= a.o_lo && info_i.w < a.o_hi;
                const double4 pi = a.sorted_pos[s_i];
                xi = pi.x;
                yi = pi.y;
                zi = pi.z;
                qi = pi.w;
            }
            if (!__any_sync(0xffffffffu, active)) continue;
            const float xf = active ? (float)xi : 1.0e18f, yf = (float)yi, zf = (float)zi;

            double fx = 0.0, fy = 0.0, fz = 0.0;
            // next free slot of this lane's queue as a 32-bit shared-memory address (slots are 128 bytes apart)
            const unsigned queue_base = (unsigned)__cvta_generic_to_shared(q);
            unsigned tail = queue_base;

            // phase 2: walk the private queues (FP64)
            auto flush = [&]() {
                const int queued = (int)((tail - queue_base) >> 7);
                const int longest = __reduce_max_sync(0xffffffffu, queued);
                for (int k = 0; k < longest; k++) {
                    if (k < queued) {
                        const unsigned entry = q[k * 32];
                        const int code = (int)(entry >> 26);
                        const int s_j = (int)(entry & QUEUE_INDEX_MASK);
                        if (s_j != s_i) {
                            evaluate_candidate<LJ_ONLY, MODE>(a, sp, xi - offset64[code][0], yi - offset64[code][1],
                                                              zi - offset64[code][2], qi, info_i, s_i, s_j, fx, fy, fz, acc);
                        }
                    }
                }
                tail = queue_base;
            };

            // phase 1: stream the 27 neighbour cells (FP32 pre-filter); rows of three cells along x
            for (int row = 0; row < 9; row++) {
                int ny = cy + (row % 3) - 1, nz = cz + (row / 3) - 1;
                ny += ny < 0 ? a.g.nc[1] : 0;
                ny -= ny >= a.g.nc[1] ? a.g.nc[1] : 0;
                nz += nz < 0 ? a.g.nc[2] : 0;
                nz -= nz >= a.g.nc[2] ? a.g.nc[2] : 0;
                const int row_base = (nz * a.g.nc[1] + ny) * a.g.nc[0];
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const int code = row * 3 + dx;
                    int nx = cx + dx - 1;
                    nx += nx < 0 ? a.g.nc[0] : 0;
                    nx -= nx >= a.g.nc[0] ? a.g.nc[0] : 0;
                    const int s0 = a.cell_start[row_base + nx], s1 = a.cell_start[row_base + nx + 1];
                    // atom i seen from the neighbour cell's centre
                    const float xr = xf - offset32[code][0], yr = yf - offset32[code][1], zr = zf - offset32[code][2];
                    const unsigned tag = (unsigned)code << 26;
                    for (int chunk = s0; chunk < s1; chunk += CELL_CHUNK) {
                        if (__any_sync(0xffffffffu, tail > queue_base + 128u * (CELL_QUEUE - CELL_CHUNK))) flush();
                        const int stop = min(s1, chunk + CELL_CHUNK);
#pragma unroll 4
                        for (int s_j = chunk; s_j < stop; s_j++) {
                            const float4 f = __ldg(a.sorted_f32 + s_j);  // same address in every lane: one broadcast
                            const float ddx = xr - f.x, ddy = yr - f.y, ddz = zr - f.z;
                            const float r2 = ddx * ddx + ddy * ddy + ddz * ddz;
                            if (r2 < a.cutoff2_f32) {  // the atom itself passes too; phase 2 drops it
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(tail), "r"(tag + (unsigned)s_j) : "memory");
                                tail += 128u;
                            }
                        }
                    }
                }
            }
            flush();

            if (active && a.write_forces) {
                a.force[3 * info_i.w] = fx;
                a.force[3 * info_i.w + 1] = fy;
                a.force[3 * info_i.w + 2] = fz;
            }
        }
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
    if (ctx->n >= (int64_t)QUEUE_INDEX_MASK) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "the cell list handles at most %u atoms per GPU", QUEUE_INDEX_MASK);
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
    a.ncells = ncells;
    a.o_lo = (int)o_lo;
    a.o_hi = (int)o_hi;
    a.cell_start = ctx->cell_start.ptr;
    a.sorted_pos = ctx->sorted_pos.ptr;
    a.sorted_f32 = ctx->sorted_f32.ptr;
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
    a.lj_sigma2 = a.lj_epsilon24 = a.lj_epsilon4 = a.lj_cutoff2 = a.lj_shift = 0.0;
    if (lj_only) {
        const lumol_cuda_pair& p = ctx->host_pairs[0];
        a.lj_sigma2 = p.p[0] * p.p[0];
        a.lj_epsilon24 = 24.0 * p.p[1];
        a.lj_epsilon4 = 4.0 * p.p[1];
        a.lj_cutoff2 = p.cutoff * p.cutoff;
        a.lj_shift = p.shift;
    }

    const size_t table_bytes = lj_only ? 0 : sizeof(PairParams) * (size_t)ctx->nkinds * ctx->nkinds;
    if (table_bytes > 100 * 1024) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "too many particle kinds (%d) for the shared pair table",
                         ctx->nkinds);
    }
    const size_t smem = (size_t)CELL_WARPS * CELL_QUEUE * 32 * sizeof(unsigned) + table_bytes;

    // persistent warps: each walks a contiguous run of cells (neighbouring cells share 2/3 of their
    // neighbourhood, which then comes from L1)
    int blocks = ctx->sm_count * 6;
    int warps = blocks * CELL_WARPS;
    int cells_per_warp = (ncells + warps - 1) / warps;
    if (cells_per_warp < 1) cells_per_warp = 1;
    blocks = (ncells + cells_per_warp * CELL_WARPS - 1) / (cells_per_warp * CELL_WARPS);
    a.cells_per_warp = cells_per_warp;

    const bool full = req.energy || req.virial;
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
