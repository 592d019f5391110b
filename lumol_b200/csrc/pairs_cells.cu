// Cell-list / Verlet-list pair path for boxes with at least three (cut-off + skin) lengths per edge.
// The reference has no neighbour search at all: this replaces the O(N^2) loops of
//   Forces::compute        sys/compute.rs:37-55
//   EnergyEvaluator::pairs sys/energy.rs:47-59
//   AtomicVirial::compute  sys/compute.rs:202-216
//   Ewald real space       energy/global/ewald.rs:430-530
//   Wolf                   energy/global/wolf.rs:177-283
// by an O(N) evaluation over exactly the same pairs: every listed pair is still tested against its cut-off in
// FP64 at every evaluation; the list only bounds which pairs can pass.
//
// Rebuild (only when some atom moved more than skin / 2 since the last one; decided on the device, no host
// round trip: ONE cooperative kernel, launched at every evaluation, that returns at once unless the refresh kernel
// wrote the evaluation's epoch into the rebuild flag):
//   1. counting sort of the atoms into cells of edge >= cut-off + skin (z-major cells, atoms of a cell
//      contiguous and ordered by original index, so every rank of a multi-GPU run derives the same order);
//   2. per atom, in cell order ("sorted index"):
//        sorted_pos  double4 (x, y, z, q): position RELATIVE TO THE CENTRE OF THE ATOM'S CELL + charge
//        frame       x | y | z planes: the same position in the frame of the box, in units of sigma, every cell
//                    starting at an even index (bulk copies of the Lennard-Jones kernel)
//        sorted_f32  float4: the cell-relative position in FP32 (pre-filter of the list build)
//        sorted_info int4 (kind, first atom of the molecule, bond-distance row, original index)
//      With cell-relative coordinates the separation of atom i in cell h and atom j in cell h + (a, b, c) is
//      (x_i - (a, b, c) * edge) - x_j whatever the periodic wrap of the cell index;
//   3. staging tables of the Lennard-Jones kernel: per block of TB consecutive atoms the cells of its neighbourhood,
//      their slots in the block's shared-memory copy, the runs of consecutive cells (one bulk copy each);
//   4. list build: one lane per atom, one warp per 32-atom chunk of a home cell; the warp streams the 27 neighbour
//      cells, every lane tests the broadcast candidate against its own atom in FP32 with an enlarged radius and
//      appends survivors to its column: (offset code << 26) | j in the general format, 16-bit slot numbers for a
//      staged block; the columns of 32 consecutive atoms are interleaved in one contiguous slab;
//   5. (separate kernel) bank-aware order of the staged columns.
// Every evaluation:
//   - positions in sorted order are refreshed as rel0 + (x - x_at_build), so an atom keeps the periodic image
//     it had at build time, and the displacement is compared with skin / 2;
//   - general kernel (any potential, Ewald real space, Wolf): up to four threads per atom walk its column, gather the
//     FP64 position of each neighbour (software-pipelined one entry ahead), apply the exact r < rc test and
//     accumulate the force in registers; energies / virials are taken from the side with the larger sorted index;
//   - Lennard-Jones kernel: one thread per atom, neighbourhood of the block staged in shared memory by TMA bulk copies,
//     17 FP64 instructions per listed pair; half of the energy and virial from each side of a pair.
// No atomics on the pair path: every pair is evaluated from both sides, sums are reduced in a fixed order.
#include "context.hpp"
#include "neighbor_common.cuh"

#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

namespace lumol {

// ------------------------------------------------------------------------------------------------
// neighbour-path choice
// ------------------------------------------------------------------------------------------------

constexpr int CELL_PATH_MIN_ATOMS = 3000;

// Returns 1 when the cell list can (and should) be used, 0 for the all-pairs kernel, -1 when it was forced
// but is impossible; fills ctx->ncell and ctx->skin_effective.
int choose_neighbor_path(Context* ctx, double cutoff) {
    ctx->ncell[0] = ctx->ncell[1] = ctx->ncell[2] = 0;
    bool possible = ctx->cell.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC && cutoff > 0.0;
    if (possible) {
        const double lengths[3] = {ctx->cell.h[0], ctx->cell.h[4], ctx->cell.h[8]};
        // the skin shrinks when the box cannot hold three (cut-off + skin) lengths per edge
        double skin = ctx->skin;
        for (int d = 0; d < 3; d++) {
            const double room = lengths[d] / 3.0 - cutoff;
            if (room < skin) skin = room * 0.999;
        }
        if (skin < 0.0) skin = 0.0;
        ctx->skin_effective = skin;
        for (int d = 0; d < 3; d++) {
            double nc = floor(lengths[d] / (cutoff + skin));
            if (nc > 1024.0) nc = 1024.0;
            ctx->ncell[d] = (int)nc;
            if (ctx->ncell[d] < 3) possible = false;
        }
    }
    if (ctx->forced_path == 0) return 0;
    if (ctx->forced_path >= 1) return possible ? 1 : -1;
    return possible && ctx->n >= CELL_PATH_MIN_ATOMS ? 1 : 0;
}

// Box-frame arrays (read by the bulk copies of the Lennard-Jones kernel, 16-byte granularity): the atoms of cell c
// start at an even index, one slot of padding per cell at most.
__host__ __device__ __forceinline__ int frame_start(int cell, int first_sorted) { return (first_sorted + cell + 1) & ~1; }

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
    double4* __restrict__ rel0;
    float4* __restrict__ sorted_f32;
    int4* __restrict__ sorted_info;
    int* __restrict__ sorted_cell;
    double* __restrict__ frame;  // x, y, z planes of `frame_stride` doubles: positions in the frame of the box
    size_t frame_stride;
    double frame_scale;  // 1 / sigma for a Lennard-Jones system: the staged kernel works in reduced lengths
    double* __restrict__ xref;
};

// second pass: rank inside the cell = number of cell mates with a smaller original index
__device__ __forceinline__ void cell_scatter_phase(int vb, const ScatterArgs& a) {
    const int s = vb * REBUILD_THREADS + threadIdx.x;
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
    const double px = a.pos[3 * i], py = a.pos[3 * i + 1], pz = a.pos[3 * i + 2];
    const double x = wrap_coordinate(px, a.g.length[0]) - ((double)cx + 0.5) * a.g.edge[0];
    const double y = wrap_coordinate(py, a.g.length[1]) - ((double)cy + 0.5) * a.g.edge[1];
    const double z = wrap_coordinate(pz, a.g.length[2]) - ((double)cz + 0.5) * a.g.edge[2];
    const double4 p = make_double4(x, y, z, a.charge[i]);
    a.sorted_pos[dst] = p;
    a.rel0[dst] = p;
    a.sorted_f32[dst] = make_float4((float)x, (float)y, (float)z, 0.0f);
    a.sorted_info[dst] = make_int4((int)a.kind[i], a.mol_first[i], a.bd_row[i], i);
    a.sorted_cell[dst] = c;
    // the same position in the frame of the box (cell centre + relative part), read by the staged kernel
    const int f = frame_start(c, lo) + rank;
    a.frame[f] = (x + ((double)cx + 0.5) * a.g.edge[0]) * a.frame_scale;
    a.frame[a.frame_stride + f] = (y + ((double)cy + 0.5) * a.g.edge[1]) * a.frame_scale;
    a.frame[2 * a.frame_stride + f] = (z + ((double)cz + 0.5) * a.g.edge[2]) * a.frame_scale;
    a.xref[3 * i] = px;
    a.xref[3 * i + 1] = py;
    a.xref[3 * i + 2] = pz;
}

// Between rebuilds: positions in sorted order follow the atoms with the image they had at build time, and any
// displacement above skin / 2 requests a rebuild.
__global__ void __launch_bounds__(256)
    list_update_kernel(int n, GridView g, const int* __restrict__ order, const double* __restrict__ pos,
                       const double* __restrict__ xref, const double4* __restrict__ rel0,
                       const int* __restrict__ sorted_cell, const int* __restrict__ cell_start,
                       double4* __restrict__ sorted_pos,
                       double* __restrict__ frame, size_t frame_stride, double frame_scale, double threshold2, int epoch,
                       const unsigned char* __restrict__ cell_needed, int* __restrict__ flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = order[s];
    const double dx = pos[3 * i] - xref[3 * i];
    const double dy = pos[3 * i + 1] - xref[3 * i + 1];
    const double dz = pos[3 * i + 2] - xref[3 * i + 2];
    if (!(dx * dx + dy * dy + dz * dz <= threshold2)) flags[FLAG_REBUILD] = epoch;  // also catches NaN
    // Sharded runs: a rank refreshes only the cells its own lists reach, but it watches the displacement of EVERY atom,
    // so that all ranks rebuild at the same step (a rank rebuilding alone stalls the others at the next exchange:
    // measured +60 us per step on eight GPUs when each rank decided from its own atoms only).
    if (cell_needed != nullptr && cell_needed[sorted_cell[s]] == 0) return;
    const double4 r = rel0[s];
    const double x = r.x + dx, y = r.y + dy, z = r.z + dz;
    // the cell-relative copy is read by the general kernel and by blocks that could not be staged only
    if (frame == nullptr || flags[FLAG_UNSTAGED] != 0) sorted_pos[s] = make_double4(x, y, z, r.w);
    if (frame != nullptr) {
        const int c = sorted_cell[s];
        const int cx = c % g.nc[0], cy = (c / g.nc[0]) % g.nc[1], cz = c / (g.nc[0] * g.nc[1]);
        const int first = cell_start[c];
        const int f = frame_start(c, first) + (s - first);
        frame[f] = (x + ((double)cx + 0.5) * g.edge[0]) * frame_scale;
        frame[frame_stride + f] = (y + ((double)cy + 0.5) * g.edge[1]) * frame_scale;
        frame[2 * frame_stride + f] = (z + ((double)cz + 0.5) * g.edge[2]) * frame_scale;
    }
}


// ------------------------------------------------------------------------------------------------
// staging tables
// ------------------------------------------------------------------------------------------------
//
// The Lennard-Jones force kernel works on blocks of TB consecutive atoms of the cell order.  The atoms any of
// them can interact with live in the 27-cell neighbourhoods of the block's home cells; the kernel copies those
// once per evaluation into shared memory, in coordinates relative to the block's first cell with the periodic
// shift already applied, and the list of a staged block holds 16-bit indices into that copy.  The inner loop
// then has no image arithmetic and no global gathers at all (the previous version was bound by the L1 tag
// stage: two scattered 32-byte global loads per neighbour and lane).
//
// Home cells of a block are the contiguous range [c0, c0 + K) of the x-fastest linear cell order, i.e. up to
// a few row segments.  Segment g covers x in [xa, xa + len) of row r0 + g and brings the (len + 2) x 3 x 3
// cells around it, enumerated plane by plane; entry index and staged offset are pure arithmetic of (c0, K),
// so nothing needs searching.  A block whose neighbourhood does not fit (dense cells, sparse gas, tiny grids)
// keeps the 32-bit global list format and takes the gather path inside the same kernel.

constexpr int TB = 256;  // atoms per block of the force kernel
constexpr int STAGE_MAX_ENTRIES = 160;
constexpr int STAGE_MAX_SEGMENTS = 6;
constexpr int STAGE_SLOTS = 4528;                   // atoms of the staged copy, dummy slot 0 included (a multiple of 16)
constexpr int STAGE_BYTES = 3 * STAGE_SLOTS * 8;    // x | y | z planes: 106 KiB, two blocks per SM
constexpr int STAGE_ATOMS_MAX = STAGE_SLOTS;

struct BlockRows {
    int c0, K, nx, r0, xa0, len0, nseg, nentries;

    __host__ __device__ void init(int first_cell, int cells, int cells_x) {
        c0 = first_cell;
        K = cells;
        nx = cells_x;
        r0 = c0 / nx;
        xa0 = c0 - r0 * nx;
        len0 = min(nx - xa0, K);
        const int rest = K - len0;
        const int more = (rest + nx - 1) / nx;
        nseg = 1 + more;
        nentries = 9 * (len0 + 2);
        if (more > 0) {
            const int last = rest - (more - 1) * nx;
            nentries += (more - 1) * 9 * (nx + 2) + 9 * (last + 2);
        }
    }
    __host__ __device__ void segment(int g, int& xa, int& len, int& base) const {
        if (g == 0) {
            xa = xa0;
            len = len0;
            base = 0;
        } else {
            xa = 0;
            len = min(nx, K - len0 - (g - 1) * nx);
            base = 9 * (len0 + 2) + (g - 1) * 9 * (nx + 2);
        }
    }
};

struct TableArgs {
    GridView g;
    int n, nblocks;
    int allow;            // 0: every block keeps the global format (general kernel)
    int stage_atoms_max;  // shared-memory capacity of the force kernel in atoms, dummy slot included
    const int* __restrict__ cell_start;
    const int* __restrict__ sorted_cell;
    int2* __restrict__ runs;     // first box-frame index, first slot | periodic image << 16 of every run
    int4* __restrict__ header;   // c0, K (negated when the rank owns no atom of the block), entries | runs << 16 (-1: not staged),
                                 // staged slots (negated when some cell is seen through a periodic boundary)
    int4* __restrict__ entries;  // first box-frame index, slot offset | count << 16, image x | image y << 16, image z
    const int4* __restrict__ sorted_info;
    int o_lo, o_hi;
    int* __restrict__ flags;
};

// one warp per block of TB atoms
__device__ __forceinline__ void block_table_phase(int vb, const TableArgs& a) {
    const int lane = threadIdx.x & 31;
    const int block = vb * REBUILD_WARPS + (threadIdx.x >> 5);
    if (block >= a.nblocks) return;
    const int s_first = block * TB, s_last = min(a.n, s_first + TB) - 1;
    const int c_first = a.sorted_cell[s_first], c_last = a.sorted_cell[s_last];
    BlockRows rows;
    rows.init(c_first, c_last - c_first + 1, a.g.nc[0]);
    bool staged = a.allow != 0 && rows.nseg <= STAGE_MAX_SEGMENTS && rows.nentries <= STAGE_MAX_ENTRIES;
    // Slot 0 is the far-away dummy the padding entries point to.  The cells of one row that follow each other in
    // the box-frame arrays form a run: one bulk copy per run and coordinate plane, so the slots of a run mirror
    // the layout of the frame arrays (even starts, up to two slots of padding between cells).
    int total = 2, nruns = 0;
    bool images = false;
    if (staged) {
        const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
        for (int base = 0; base < rows.nentries; base += 32) {
            const int e = base + lane;
            int src = 0, count = 0, sx = 0, sy = 0, sz = 0;
            int slots = 0;          // distance to the slot of the next entry
            bool run_start = false;
            if (e < rows.nentries) {
                int g = 0, rem = e;
                const int first = 9 * (rows.len0 + 2);
                if (e >= first) {
                    g = 1 + (e - first) / (9 * (nx + 2));
                    rem = (e - first) - (g - 1) * 9 * (nx + 2);
                }
                int xa, len, seg_base;
                rows.segment(g, xa, len, seg_base);
                const int w = len + 2;
                const int plane = rem / w, k = rem - plane * w;
                const int r = rows.r0 + g;
                const int y = r % ny, z = r / ny;
                const int ux = xa - 1 + k, uy = y + (plane % 3) - 1, uz = z + (plane / 3) - 1;
                const int ax = ux < 0 ? ux + nx : (ux >= nx ? ux - nx : ux);
                const int ay = uy < 0 ? uy + ny : (uy >= ny ? uy - ny : uy);
                const int az = uz < 0 ? uz + nz : (uz >= nz ? uz - nz : uz);
                const int cell = (az * ny + ay) * nx + ax;
                const int first_sorted = a.cell_start[cell];
                count = a.cell_start[cell + 1] - first_sorted;
                src = frame_start(cell, first_sorted);  // index into the box-frame arrays
                // periodic image of the cell as seen from the block (the staged kernel works in the frame of the box)
                sx = ux < 0 ? -1 : (ux >= nx ? 1 : 0);
                sy = uy < 0 ? -1 : (uy >= ny ? 1 : 0);
                sz = uz < 0 ? -1 : (uz >= nz ? 1 : 0);
                // the run continues into the next entry when that is the next cell of the same row of the grid
                run_start = k == 0 || ax == 0;
                const bool run_end = k == w - 1 || ax == nx - 1;
                slots = run_end ? ((count + 1) & ~1) : frame_start(cell + 1, a.cell_start[cell + 1]) - src;
            }
            images = images || (sx | sy | sz) != 0;
            // exclusive scan of the slot counts over the warp
            int scan = slots;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, scan, o);
                if (lane >= o) scan += t;
            }
            const int offset = total + scan - slots;
            total += __shfl_sync(0xffffffffu, scan, 31);
            const unsigned starts = __ballot_sync(0xffffffffu, run_start);
            if (e < rows.nentries && offset + slots <= 65535) {
                a.entries[(size_t)block * STAGE_MAX_ENTRIES + e] =
                    make_int4(src, offset | (count << 16), (sx & 0xffff) | (sy << 16), sz);
                if (run_start) {
                    const int r = nruns + __popc(starts & ((1u << lane) - 1u));
                    // bits 16..21: periodic image of the run, two bits per axis (0: none, 1: +L, 3: -L)
                    const int image = (sx & 3) | ((sy & 3) << 2) | ((sz & 3) << 4);
                    a.runs[(size_t)block * STAGE_MAX_ENTRIES + r] = make_int2(src, offset | (image << 16));
                }
            }
            nruns += __popc(starts);
        }
        if (total > a.stage_atoms_max || total > 65535) staged = false;
        __syncwarp();
    }
    if (!staged && lane == 0 && a.allow != 0) atomicAdd(a.flags + FLAG_UNSTAGED, 1);
    // does this rank own an atom of the block?
    bool owned = false;
    for (int s = s_first + lane; s <= s_last; s += 32) {
        const int orig = a.sorted_info[s].w;
        owned = owned || (orig >= a.o_lo && orig < a.o_hi);
    }
    owned = __any_sync(0xffffffffu, owned);
    images = __any_sync(0xffffffffu, images);
    if (lane == 0) {
        a.header[block] =
            make_int4(rows.c0, owned ? rows.K : -rows.K, staged ? (rows.nentries | (nruns << 16)) : -1, images ? -total : total);
    }
}

// ------------------------------------------------------------------------------------------------
// list build
// ------------------------------------------------------------------------------------------------

constexpr unsigned LIST_INDEX_MASK = (1u << 26) - 1u;
constexpr int BUILD_WARPS = REBUILD_WARPS;
constexpr int BUILD_CHUNKS = 8;  // work items per cell of the list build

struct BuildArgs {
    GridView g;
    int ncells;
    int cells_per_warp;  // work items (cell, chunk) per warp
    int o_lo, o_hi;  // original-index range of the atoms this rank owns (lists are built for those)
    int capacity;    // 32-bit entries per atom (a staged column holds as many 16-bit ones in half the space)
    float radius2;   // (cut-off + skin)^2, enlarged by 1e-4 relative
    const int* __restrict__ cell_start;
    const float4* __restrict__ sorted_f32;
    const int4* __restrict__ sorted_info;
    const int4* __restrict__ blk_header;
    const int4* __restrict__ blk_entries;
    unsigned short* __restrict__ self_local;
    unsigned* __restrict__ nlist;
    int* __restrict__ ncount;
    unsigned char* __restrict__ cell_needed;  // cells whose atoms this rank's lists refer to
    int* __restrict__ flags;
};

// offset32: (a, b, c) * edge of the 27 neighbour-cell offsets, in shared memory
__device__ __forceinline__ void list_build_phase(int vb, const BuildArgs& a, const float (*offset32)[3]) {
    const int lane = threadIdx.x & 31;
    const int global_warp = vb * BUILD_WARPS + (threadIdx.x >> 5);
    // work items: (k, cell) for k < BUILD_CHUNKS; item (k, c) takes the 32-atom chunks k, k + BUILD_CHUNKS, ... of
    // cell c, so that dense cells (100 atoms at water density, few cells) still spread over every SM
    const int item_lo = global_warp * a.cells_per_warp;
    const int item_hi = min(a.ncells * BUILD_CHUNKS, item_lo + a.cells_per_warp);

    for (int item = item_lo; item < item_hi; item++) {
        // chunk-major order: the consecutive items of a warp are different cells (at liquid-argon density only chunk 0
        // of a cell holds atoms)
        const int first_chunk = item / a.ncells, c = item - first_chunk * a.ncells;
        const int hs = a.cell_start[c], he = a.cell_start[c + 1];
        if (hs + 32 * first_chunk >= he) continue;
        const int cx = c % a.g.nc[0];
        const int cy = (c / a.g.nc[0]) % a.g.nc[1];
        const int cz = c / (a.g.nc[0] * a.g.nc[1]);
        for (int base = hs + 32 * first_chunk; base < he; base += 32 * BUILD_CHUNKS) {
            const int s_i = base + lane;
            bool active = s_i < he;
            float xf = 1.0e18f, yf = 0.0f, zf = 0.0f;  // lanes without an owned atom sit far away
            // staging table of the lane's block (the lanes of one warp can sit in two blocks)
            bool staged = false;
            const int4* entries = a.blk_entries;
            int entry_first = 0, entry_width = 0;
            if (active) {
                const int block = s_i / TB;
                const int4 header = a.blk_header[block];
                staged = header.z >= 0;
                if (staged) {
                    BlockRows rows;
                    rows.init(header.x, abs(header.y), a.g.nc[0]);
                    int xa, len, seg_base;
                    rows.segment(c / a.g.nc[0] - rows.r0, xa, len, seg_base);
                    entries += (size_t)block * STAGE_MAX_ENTRIES;
                    entry_width = len + 2;
                    entry_first = seg_base + (cx - xa);  // entry of offset code 0 (dx = -1, plane 0)
                    a.self_local[s_i] =
                        (unsigned short)((entries[entry_first + 4 * entry_width + 1].y & 0xffff) + (s_i - hs));
                }
                const int orig = a.sorted_info[s_i].w;
                active = orig >= a.o_lo && orig < a.o_hi;
                if (active) {
                    const float4 f = a.sorted_f32[s_i];
                    xf = f.x;
                    yf = f.y;
                    zf = f.z;
                }
            }
            if (!__any_sync(0xffffffffu, active)) {
                if (s_i < he) a.ncount[s_i] = 0;
                continue;
            }
            if (lane < 27) {
                int mx = cx + (lane % 3) - 1, my = cy + ((lane / 3) % 3) - 1, mz = cz + (lane / 9) - 1;
                mx += mx < 0 ? a.g.nc[0] : (mx >= a.g.nc[0] ? -a.g.nc[0] : 0);
                my += my < 0 ? a.g.nc[1] : (my >= a.g.nc[1] ? -a.g.nc[1] : 0);
                mz += mz < 0 ? a.g.nc[2] : (mz >= a.g.nc[2] ? -a.g.nc[2] : 0);
                a.cell_needed[(mz * a.g.nc[1] + my) * a.g.nc[0] + mx] = 1;
            }
            // The columns of 32 consecutive atoms form one contiguous slab of (capacity / 4) x 32 16-byte words,
            // interleaved word by word: word w of atom i is word (w * 32 + i % 32) of slab i / 32, so that a
            // warp of the force kernel streams its slab front to back with 512-byte coalesced reads.  A word
            // holds 4 global entries (offset code << 26) | j, or 8 16-bit indices into the block's staged copy.
            unsigned* column = a.nlist + ((size_t)(s_i >> 5) * (a.capacity >> 2) * 32 + (s_i & 31)) * 4;
            unsigned short* column16 = reinterpret_cast<unsigned short*>(column);
            int count = 0;
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
                    // what is stored for neighbour s_j: tag + s_j
                    unsigned tag = (unsigned)code << 26;
                    if (staged) tag = (unsigned)((entries[entry_first + row * entry_width + dx].y & 0xffff) - s0);
                    // eight candidates per round, loaded before the first one is tested (uniform addresses: one
                    // broadcast each); the compiler does not hoist these loads over the divergent stores by itself
                    for (int first = s0; first < s1; first += 8) {
                        float4 f[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) f[u] = __ldg(a.sorted_f32 + min(first + u, s1 - 1));
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int s_j = first + u;
                            const float ddx = xr - f[u].x, ddy = yr - f[u].y, ddz = zr - f[u].z;
                            const float r2 = ddx * ddx + ddy * ddy + ddz * ddz;
                            if (r2 < a.radius2 && s_j != s_i && s_j < s1) {
                                if (count < a.capacity) {
                                    const unsigned value = tag + (unsigned)s_j;
                                    if (staged) {
                                        column16[(count >> 3) * 256 + (count & 7)] = (unsigned short)value;
                                    } else {
                                        column[(count >> 2) * 128 + (count & 3)] = value;
                                    }
                                }
                                count++;
                            }
                        }
                    }
                }
            }
            if (active) {
                if (count > a.capacity) {
                    a.flags[FLAG_OVERFLOW] = 1;
                    count = a.capacity;
                }
                a.ncount[s_i] = count;
                // the force kernel reads whole 16-byte words: pad the last one (global format: a valid index that
                // fails the listed test; staged format: the dummy slot 0)
                if (staged) {
                    for (int k = count; (k & 7) != 0; k++) column16[(k >> 3) * 256 + (k & 7)] = 0;
                } else {
                    for (int k = count; (k & 3) != 0; k++) column[(k >> 2) * 128 + (k & 3)] = (unsigned)s_i;
                }
            } else if (s_i < he) {
                a.ncount[s_i] = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// bank-aware order of the staged columns
// ------------------------------------------------------------------------------------------------
//
// The staged copy is an array of (x, y, z) doubles, 24 bytes per atom: slot s sits in the 8-byte bank pairs
// (3 s + c) mod 16, so two lanes of a half warp collide exactly when their slots are different and congruent
// modulo 16.  In arrival order the 32 columns of a warp hit the banks at random (measured: 5.1 wavefronts per
// LDS.64 instead of 2; the shared-memory pipe saturated long before the FP64 pipe).  The order of a column is
// free, so every lane sorts its entries into 16 buckets by slot mod 16 and emits them round robin, lane l
// starting at bucket l mod 16: while its buckets last, a half warp touches 16 different bank pairs at every
// step.  An exhausted bucket is replaced by the next non-empty one (simulation: 2.9 wavefronts per LDS.64).

constexpr int REORDER_THREADS = 32;

// one warp per slab of 32 columns; shared: the slab's entries grouped by bucket, [position][lane]
__global__ void __launch_bounds__(REORDER_THREADS)
    list_reorder_kernel(int n, int capacity, const int4* __restrict__ blk_header, const int* __restrict__ ncount,
                        unsigned* __restrict__ nlist, int epoch, const int* __restrict__ flags) {
    if (flags[FLAG_REBUILD] != epoch) return;
    extern __shared__ unsigned short reorder_smem[];
    unsigned short* sorted = reorder_smem;  // entry p of thread t at [p * 32 + t], grouped by bucket
    __shared__ unsigned short cursor[16][REORDER_THREADS], last[16][REORDER_THREADS];
    const int t = threadIdx.x;
    for (int slab = blockIdx.x; slab * REORDER_THREADS < n; slab += gridDim.x) {
        const int s_i = slab * REORDER_THREADS + t;
        if (s_i >= n) continue;
        if (blk_header[s_i / TB].z < 0) continue;  // global format: no shared-memory gathers
        const int count = ncount[s_i];
        uint4* words = reinterpret_cast<uint4*>(nlist) + (size_t)(s_i >> 5) * (capacity >> 2) * 32 + (s_i & 31);
        const int nwords = (count + 7) >> 3;
#pragma unroll
        for (int b = 0; b < 16; b++) last[b][t] = 0;
        // whole 16-byte words in, eight entries each (the tail of the last word is padding and is not counted)
        for (int w = 0; w < nwords; w++) {
            const uint4 word = words[w * 32];
            const unsigned pairs[4] = {word.x, word.y, word.z, word.w};
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const unsigned short e = (unsigned short)(q & 1 ? pairs[q >> 1] >> 16 : pairs[q >> 1] & 0xffffu);
                if (8 * w + q < count) last[e & 15][t]++;
            }
        }
        int running = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            cursor[b][t] = (unsigned short)running;
            running += last[b][t];
            last[b][t] = (unsigned short)running;
        }
        // stable scatter into the buckets (second pass over the words: they are in L1 / L2 now)
        for (int w = 0; w < nwords; w++) {
            const uint4 word = words[w * 32];
            const unsigned pairs[4] = {word.x, word.y, word.z, word.w};
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const unsigned short e = (unsigned short)(q & 1 ? pairs[q >> 1] >> 16 : pairs[q >> 1] & 0xffffu);
                if (8 * w + q < count) {
                    const int position = cursor[e & 15][t]++;
                    sorted[position * REORDER_THREADS + t] = e;
                }
            }
        }
        // rewind, then emit round robin, whole words out
        running = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            cursor[b][t] = (unsigned short)running;
            running = last[b][t];
        }
        int bucket = s_i & 15;
        for (int w = 0; w < nwords; w++) {
            unsigned pairs[4] = {0u, 0u, 0u, 0u};  // padding entries point at the dummy slot 0
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (8 * w + q < count) {
                    int b = bucket;
                    while (cursor[b][t] == last[b][t]) b = (b + 1) & 15;  // some bucket is non-empty
                    const int position = cursor[b][t]++;
                    pairs[q >> 1] |= (unsigned)sorted[position * REORDER_THREADS + t] << (16 * (q & 1));
                    bucket = (bucket + 1) & 15;
                }
            }
            words[w * 32] = make_uint4(pairs[0], pairs[1], pairs[2], pairs[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the rebuild as one cooperative kernel
// ------------------------------------------------------------------------------------------------
//
// Whether to rebuild is decided on the device: list_update_kernel writes the current epoch into
// flags[FLAG_REBUILD] when an atom moved too far.  The rebuild used to be eleven kernels launched every step,
// each returning at once while no rebuild was due: about 50 us per step of the 1M-atom box (empty grids of
// thousands of blocks are not free).  It is now ONE cooperative kernel whose phases are separated by grid-wide
// barriers, followed by the reorder kernel on a small persistent grid: two cheap launches when nothing is due.

struct RebuildArgs {
    int n, ncells, scan_blocks, nblocks, epoch;
    GridView g;
    const double* position;
    int *cell_of, *slot_of, *cell_count, *cell_start, *scan_scratch, *grouped;
    ScatterArgs scatter;
    TableArgs table;
    BuildArgs build;
    int* flags;
};

__global__ void __launch_bounds__(REBUILD_THREADS) rebuild_kernel(RebuildArgs r) {
    if (r.flags[FLAG_REBUILD] != r.epoch) return;  // every block takes the same decision: nobody waits at a barrier
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ int scan_shared[33];
    __shared__ float offset32[27][3];
    if (threadIdx.x < 27) {
        const int t = threadIdx.x;
        offset32[t][0] = (float)((double)((t % 3) - 1) * r.g.edge[0]);
        offset32[t][1] = (float)((double)(((t / 3) % 3) - 1) * r.g.edge[1]);
        offset32[t][2] = (float)((double)((t / 9) - 1) * r.g.edge[2]);
    }
    const int atom_blocks = (r.n + REBUILD_THREADS - 1) / REBUILD_THREADS;

    cell_zero_phase(r.ncells + 1, r.cell_count, r.build.cell_needed, r.flags);
    grid.sync();
    for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) {
        cell_assign_phase(vb, r.n, r.g, r.position, r.cell_of, r.slot_of, r.cell_count, r.flags);
    }
    grid.sync();
    if (r.flags[FLAG_NONFINITE] != 0) return;  // every block reads the same value: the force kernels return as well
    for (int vb = blockIdx.x; vb < r.scan_blocks; vb += gridDim.x) {
        scan_blocks_phase(vb, r.ncells, r.cell_count, r.cell_start, r.scan_scratch, scan_shared);
    }
    grid.sync();
    if (blockIdx.x == 0) scan_sums_phase(r.scan_blocks, r.scan_scratch, scan_shared);
    grid.sync();
    for (int vb = blockIdx.x; vb < r.scan_blocks; vb += gridDim.x) {
        scan_add_phase(vb, r.ncells, r.cell_start, r.scan_scratch, r.n);
    }
    grid.sync();
    for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) {
        cell_group_phase(vb, r.n, r.cell_of, r.slot_of, r.cell_start, r.grouped);
    }
    grid.sync();
    for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) cell_scatter_phase(vb, r.scatter);
    grid.sync();
    for (int vb = blockIdx.x; vb * REBUILD_WARPS < r.nblocks; vb += gridDim.x) block_table_phase(vb, r.table);
    grid.sync();
    for (int vb = blockIdx.x; vb * BUILD_WARPS * r.build.cells_per_warp < r.ncells * BUILD_CHUNKS; vb += gridDim.x) {
        list_build_phase(vb, r.build, offset32);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) r.flags[FLAG_COUNT] += 1;
}

// ------------------------------------------------------------------------------------------------
// force kernel
// ------------------------------------------------------------------------------------------------

constexpr int NL_THREADS = 128;
constexpr int NL_NV = 16;  // same layout as the all-pairs kernel
constexpr int NL_MODE_FORCES = 0;
constexpr int NL_MODE_FULL = 1;

struct ForceArgs {
    int n;
    int o_lo, o_hi;
    int capacity;
    double edge[3];
    const unsigned* __restrict__ nlist;
    const int* __restrict__ ncount;
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
    double cutoff2;  // (largest cut-off)^2: early-out of the general path
    // LJ fast path
    double lj_sigma2, lj_epsilon24, lj_epsilon48, lj_epsilon4, lj_cutoff2, lj_shift;
    double lj_inv_sigma;                  // staged blocks: lengths in units of sigma
    long long lj_reduced_cutoff2_bits;    // (cut-off / sigma)^2 as an integer: squared distances compare like integers
    const int4* __restrict__ blk_header;
    const int4* __restrict__ blk_entries;
    const int2* __restrict__ blk_runs;
    const unsigned short* __restrict__ self_local;
    const double* __restrict__ frame;  // box-frame positions, x | y | z planes (sorted order)
    size_t frame_stride;
    int ntiles;
    double length[3];
    int write_forces;            // energy-only queries must not clobber the forces the integrator holds
    double* __restrict__ force;  // original order, n x 3
    double* __restrict__ partials;
    const int* __restrict__ flags;  // FLAG_NONFINITE: the list is not valid, nothing is evaluated
};

// 1 / x to full FP64 precision without the slow-path branch of the compiler's division: the hardware seed
// (MUFU.RCP64H, ~2^-23 relative) followed by two Newton-Raphson steps (2^-46, then rounding-limited).
// Only used where x is a squared distance already known to be normal and positive.
__device__ __forceinline__ double reciprocal(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// Lennard-Jones fast path, branch-free so that the four neighbours of a list word form four independent
// dependency chains the scheduler can interleave (the FP64 pipe has a long latency and few warps are resident).
template <int MODE>
__device__ __forceinline__ void evaluate_lj(const ForceArgs& a, double xi, double yi, double zi, int s_i, int s_j,
                                            bool listed, const double4& pj, double& fx, double& fy, double& fz,
                                            double (&acc)[NL_NV]) {
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const bool inside = listed && r2 < a.lj_cutoff2;
    const double rinv2 = reciprocal(inside ? r2 : 1.0);
    const double s2 = a.lj_sigma2 * rinv2;
    const double s6 = s2 * s2 * s2;
    // force(r) / r = -24 eps (s6 - 2 s6^2) / r^2 (functions.rs:85-88)
    double fr = a.lj_epsilon24 * s6 * (2.0 * s6 - 1.0) * rinv2;
    fr = inside ? fr : 0.0;
    fx += fr * dx;
    fy += fr * dy;
    fz += fr * dz;
    if (MODE == NL_MODE_FULL) {
        // like the staged path of lj_force_kernel: half of the energy and virial from each side of the pair
        const bool count = inside;
        const double w = fr;
        acc[0] += count ? a.lj_epsilon4 * (s6 * s6 - s6) - a.lj_shift : 0.0;
        acc[14] += count ? 1.0 : 0.0;
        acc[2] += w * dx * dx;
        acc[3] += w * dx * dy;
        acc[4] += w * dx * dz;
        acc[5] += w * dy * dy;
        acc[6] += w * dy * dz;
        acc[7] += w * dz * dz;
    }
}

// One listed neighbour, FP64: exact cut-off test and pair evaluation for the thread's atom.
template <bool SIMPLE, int MODE>
__device__ __forceinline__ void evaluate_neighbor(const ForceArgs& a, const PairParams* __restrict__ sp, double xi,
                                                  double yi, double zi, double qi, const int4& info_i, int s_i,
                                                  int s_j, const double4& pj, double& fx, double& fy, double& fz,
                                                  double (&acc)[NL_NV]) {
    const double dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const bool count = s_j > s_i;
    if (!(r2 < a.cutoff2 * 1.0000000001)) return;
    // one reciprocal square root serves the distance, f / r and the coulomb terms (no division in the pair loop)
    const double rinv = rsqrt(r2);
    const double r = r2 * rinv;
    const int4 info_j = a.sorted_info[s_j];
    const bool same_molecule = info_i.y == info_j.y;
    const unsigned bits = same_molecule ? a.bond_dist[info_i.z + (info_j.w - info_j.y)] : 0u;
    if (a.do_pairs) {
        const PairParams& pp = sp[info_i.x * a.nkinds + info_j.x];
        if (pp.potential > LUMOL_CUDA_POTENTIAL_NULL && r < pp.cutoff) {
            double scaling;
            if (!restriction_excluded(pp.restriction, bits, pp.scale14, scaling)) {
                double e, f;
                if (SIMPLE) {
                    pair_eval_simple(pp, r, rinv, e, f);
                } else {
                    pair_eval(pp, a.tables, a.table_energy, a.table_force, r, e, f);
                }
                const double fr = scaling * f * rinv;
                fx += fr * dx;
                fy += fr * dy;
                fz += fr * dz;
                if (MODE == NL_MODE_FULL && count) {
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
                ewald_real_pair_rinv(a.coulomb, excluded, qi * qj, r, rinv, e, fr);
            } else if (!excluded) {
                wolf_pair_rinv(a.coulomb, qi * qj, r, rinv, e, fr);
                e *= scaling;
                fr *= scaling;
            } else {
                active = false;
            }
            if (active) {
                fx += fr * dx;
                fy += fr * dy;
                fz += fr * dz;
                if (MODE == NL_MODE_FULL && count) {
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

// Walk of a column in the global format: four entries (one 16-byte word) at a time.  Software pipeline, per
// thread: the list word two iterations ahead and the four neighbour positions one iteration ahead are in flight
// while the current four neighbours are evaluated.
template <bool LJ_ONLY, int MODE, bool SIMPLE = false>
__device__ __forceinline__ void walk_global_column(const ForceArgs& a, const PairParams* sp, const double (*offset64)[3],
                                                   int s_i, const int4& info_i, double& fx, double& fy, double& fz,
                                                   double (&acc)[NL_NV], int part = 0, int nparts = 1) {
    // thread `part` of the `nparts` threads sharing the atom takes the words part, part + nparts, ...
    const double4 pi = a.sorted_pos[s_i];
    const int count = a.ncount[s_i];
    const uint4* words = reinterpret_cast<const uint4*>(a.nlist) + (size_t)(s_i >> 5) * (a.capacity >> 2) * 32 + (s_i & 31) +
                         (size_t)part * 32;
    const int stride = 32 * nparts;
    const int nwords = (((count + 3) >> 2) - part + nparts - 1) / nparts;
    uint4 wcur = nwords > 0 ? words[0] : make_uint4(0, 0, 0, 0);
    uint4 wnext = nwords > 1 ? words[stride] : make_uint4(0, 0, 0, 0);
    double4 pcur[4], pnext[4];
    {
        const unsigned e[4] = {wcur.x, wcur.y, wcur.z, wcur.w};
#pragma unroll
        for (int t = 0; t < 4; t++) pcur[t] = a.sorted_pos[e[t] & LIST_INDEX_MASK];
    }
    // 32-bit shared-memory address of the offset table (one multiply-add per neighbour instead of a
    // generic-pointer computation)
    const unsigned offset_base = (unsigned)__cvta_generic_to_shared(&offset64[0][0]);
    for (int w = 0; w < nwords; w++) {
        uint4 wafter = make_uint4(0, 0, 0, 0);
        if (w + 2 < nwords) wafter = words[(size_t)(w + 2) * stride];
        {
            const unsigned e[4] = {wnext.x, wnext.y, wnext.z, wnext.w};
#pragma unroll
            for (int t = 0; t < 4; t++) pnext[t] = a.sorted_pos[e[t] & LIST_INDEX_MASK];
        }
        const unsigned entries[4] = {wcur.x, wcur.y, wcur.z, wcur.w};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const bool listed = 4 * (w * nparts + part) + t < count;  // the last word is padded
            const unsigned address = offset_base + (entries[t] >> 26) * 24u;
            double ox, oy, oz;
            asm("ld.shared.f64 %0, [%1];" : "=d"(ox) : "r"(address));
            asm("ld.shared.f64 %0, [%1+8];" : "=d"(oy) : "r"(address));
            asm("ld.shared.f64 %0, [%1+16];" : "=d"(oz) : "r"(address));
            const int s_j = (int)(entries[t] & LIST_INDEX_MASK);
            if (LJ_ONLY) {
                evaluate_lj<MODE>(a, pi.x - ox, pi.y - oy, pi.z - oz, s_i, s_j, listed, pcur[t], fx, fy, fz, acc);
            } else if (listed) {
                evaluate_neighbor<SIMPLE, MODE>(a, sp, pi.x - ox, pi.y - oy, pi.z - oz, pi.w, info_i, s_i, s_j,
                                                 pcur[t], fx, fy, fz, acc);
            }
        }
        wcur = wnext;
        wnext = wafter;
#pragma unroll
        for (int t = 0; t < 4; t++) pcur[t] = pnext[t];
    }
}

// General kernel (any potential, restrictions, Ewald real space / Wolf): every block uses the global format.
// SPLIT threads (neighbouring lanes) share an atom and take every SPLIT-th word of its column: a 100k-atom water box
// has only 3072 warps of one thread per atom, too few to hide the latency of the erfc / exp chains.
template <int MODE, int SPLIT, int MINB, bool SIMPLE>
__global__ void __launch_bounds__(NL_THREADS, MINB) list_force_kernel(ForceArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairParams* sp = reinterpret_cast<PairParams*>(smem_raw);
    __shared__ double offset64[27][3];  // (a, b, c) * edge for the 27 neighbour-cell offsets
    __shared__ double scratch[32 * NL_NV];
    if (a.flags[FLAG_NONFINITE] != 0) return;

    {
        const int words = (int)(sizeof(PairParams) / sizeof(double)) * a.nkinds * a.nkinds;
        const double* src = reinterpret_cast<const double*>(a.pairs);
        double* dst = reinterpret_cast<double*>(sp);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    if (threadIdx.x < 27) {
        const int t = threadIdx.x;
        offset64[t][0] = (double)((t % 3) - 1) * a.edge[0];
        offset64[t][1] = (double)(((t / 3) % 3) - 1) * a.edge[1];
        offset64[t][2] = (double)((t / 9) - 1) * a.edge[2];
    }
    __syncthreads();

    double acc[NL_NV];
#pragma unroll
    for (int k = 0; k < NL_NV; k++) acc[k] = 0.0;

    const int s_i = (blockIdx.x * NL_THREADS + threadIdx.x) / SPLIT;
    const int part = threadIdx.x % SPLIT;
    int4 info_i = make_int4(0, 0, 0, -1);
    bool active = s_i < a.n;
    if (active) {
        info_i = a.sorted_info[s_i];
        active = info_i.w >= a.o_lo && info_i.w < a.o_hi;
    }
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (active) walk_global_column<false, MODE, SIMPLE>(a, sp, offset64, s_i, info_i, fx, fy, fz, acc, part, SPLIT);
#pragma unroll
    for (int o = 1; o < SPLIT; o <<= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (active && part == 0 && a.write_forces) {
        a.force[3 * info_i.w] = fx;
        a.force[3 * info_i.w + 1] = fy;
        a.force[3 * info_i.w + 2] = fz;
    }

    if (MODE == NL_MODE_FULL) {
        block_sum<NL_NV>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NL_NV; k++) a.partials[(size_t)blockIdx.x * NL_NV + k] = acc[k];
        }
    }
}

// 1 / x by the hardware seed and one cubically convergent step, y (1 + e + e^2) with e = 1 - x y: three
// FMAs take the seed's 2^-20 to below the FP64 rounding error.
__device__ __forceinline__ double reciprocal3(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// One staged neighbour of the Lennard-Jones kernel: 17 FP64 instructions, no branch.  A padding entry points
// at the dummy slot, far outside any cut-off.
template <int MODE>
__device__ __forceinline__ void staged_lj(const ForceArgs& a, const double* __restrict__ stage, unsigned index,
                                          double xi, double yi, double zi, double& fx, double& fy, double& fz,
                                          double (&acc)[NL_NV]) {
    const double* pj = stage + index;
    const double dx = xi - pj[0], dy = yi - pj[STAGE_SLOTS], dz = zi - pj[2 * STAGE_SLOTS];
    // reduced lengths (x / sigma): 1 / r^2 is s2 directly; the comparison runs on the integer pipe (r2 >= 0, so the
    // bit patterns order like the values; NaN compares as outside)
    const double r2 = dx * dx + dy * dy + dz * dz;
    const bool inside = __double_as_longlong(r2) < a.lj_reduced_cutoff2_bits;
    const double s2 = reciprocal3(r2);
    const double s6 = s2 * s2 * s2;
    // sigma^2 * force(r) / r = -24 eps (s6 - 2 s6^2) s2 (functions.rs:85-88); the thread rescales its sum at the end
    double fr = (s6 * s2) * fma(a.lj_epsilon48, s6, -a.lj_epsilon24);
    fr = inside ? fr : 0.0;
    fx = fma(fr, dx, fx);
    fy = fma(fr, dy, fy);
    fz = fma(fr, dz, fz);
    if (MODE == NL_MODE_FULL) {
        // every pair is visited from both sides: half of the energy and virial from each
        const double e = a.lj_epsilon4 * fma(s6, s6, -s6) - a.lj_shift;
        acc[0] += inside ? e : 0.0;
        acc[14] += inside ? 1.0 : 0.0;
        acc[2] += fr * dx * dx;
        acc[3] += fr * dx * dy;
        acc[4] += fr * dx * dz;
        acc[5] += fr * dy * dy;
        acc[6] += fr * dy * dz;
        acc[7] += fr * dz * dz;
    }
}

constexpr int LJ_THREADS = TB;

// ---- bulk asynchronous copies (TMA, UBLKCP): global -> shared, completion counted on an mbarrier ---------------

__device__ __forceinline__ unsigned shared_address(const void* pointer) { return (unsigned)__cvta_generic_to_shared(pointer); }

__device__ __forceinline__ void barrier_init(unsigned long long* barrier, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(shared_address(barrier)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void barrier_expect_bytes(unsigned long long* barrier, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(shared_address(barrier)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(double* shared_dst, const double* global_src, unsigned bytes, unsigned long long* barrier) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     shared_address(shared_dst)),
                 "l"(global_src), "r"(bytes), "r"(shared_address(barrier))
                 : "memory");
}
__device__ __forceinline__ void barrier_wait(unsigned long long* barrier, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred done;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
        "@done bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(shared_address(barrier)),
        "r"(parity)
        : "memory");
}

// Lennard-Jones kernel: TB consecutive atoms per block, one thread per atom, two blocks per SM.  The block
// issues every load it will wait for at once: each thread the head of its own column, warp 0 the bulk copies
// (one UBLKCP per cell and coordinate plane, box-frame coordinates, completion counted on an mbarrier) of the
// whole neighbourhood.  The staging costs a few hundred instructions and one memory round trip per block.
template <int MODE>
__global__ void __launch_bounds__(LJ_THREADS, 2) lj_force_kernel(ForceArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);  // x | y | z planes of STAGE_SLOTS doubles
    __shared__ double offset64[27][3];
    __shared__ __align__(8) unsigned long long arrived;
    __shared__ int4 images[STAGE_MAX_ENTRIES];  // runs seen through a periodic boundary: first slot, end, image code
    __shared__ int nimages;

    if (a.flags[FLAG_NONFINITE] != 0) return;
    double acc[NL_NV];
#pragma unroll
    for (int k = 0; k < NL_NV; k++) acc[k] = 0.0;

    const int tid = threadIdx.x;
    const int s_i = blockIdx.x * TB + tid;
    const bool present = s_i < a.n;
    // loads that do not depend on the header: atoms of other ranks have empty columns (list_build_kernel)
    const uint4* words = reinterpret_cast<const uint4*>(a.nlist) + (size_t)(s_i >> 5) * (a.capacity >> 2) * 32 + (s_i & 31);
    const int count = present ? a.ncount[s_i] : 0;
    const unsigned self = present ? a.self_local[s_i] : 0u;
    const int orig = present ? a.sorted_info[s_i].w : -1;
    // header: c0, +-K (negative: no atom of this rank), entries (-1: global format), +-staged slots (negative:
    // periodic images among them)
    const int4 header = a.blk_header[blockIdx.x];
    const int2* runs = a.blk_runs + (size_t)blockIdx.x * STAGE_MAX_ENTRIES;
    const bool mine = orig >= a.o_lo && orig < a.o_hi;
    double fx = 0.0, fy = 0.0, fz = 0.0;

    if (header.y <= 0) {
        // no atom of this rank in the block
    } else if (header.z < 0) {
        // ---- block that could not be staged: global gathers ---------------------------------------------
        if (tid < 27) {
            offset64[tid][0] = (double)((tid % 3) - 1) * a.edge[0];
            offset64[tid][1] = (double)(((tid / 3) % 3) - 1) * a.edge[1];
            offset64[tid][2] = (double)((tid / 9) - 1) * a.edge[2];
        }
        __syncthreads();
        if (mine) walk_global_column<true, MODE>(a, nullptr, offset64, s_i, a.sorted_info[s_i], fx, fy, fz, acc);
    } else {
        const int nruns = header.z >> 16;
        uint4 wcur = make_uint4(0, 0, 0, 0), wnext = make_uint4(0, 0, 0, 0);
        if (present) {
            wcur = words[0];
            wnext = words[32];  // inside the slab whatever the count
        }
        if (tid == 0) {
            barrier_init(&arrived, 1);
            nimages = 0;
        }
        if (tid < 3) stage[tid * STAGE_SLOTS] = 1.0e9 * (double)(tid + 1);  // the dummy, far outside any cut-off
        __syncthreads();
        if (tid < 32) {
            // one bulk copy per run and plane; a run ends where the next one starts
            const int total = abs(header.w);
            int2 run[STAGE_MAX_ENTRIES / 32];
            int end[STAGE_MAX_ENTRIES / 32];
#pragma unroll
            for (int k = 0; k < STAGE_MAX_ENTRIES / 32; k++) {
                const int r = tid + 32 * k;
                run[k] = r < nruns ? runs[r] : make_int2(0, 0);
                end[k] = r + 1 < nruns ? (runs[r + 1].y & 0xffff) : total;
            }
            if (tid == 0) barrier_expect_bytes(&arrived, 24u * (unsigned)(total - 2));
            __syncwarp();
#pragma unroll
            for (int k = 0; k < STAGE_MAX_ENTRIES / 32; k++) {
                if (tid + 32 * k < nruns) {
                    const int first = run[k].y & 0xffff;
                    const unsigned bytes = (unsigned)(end[k] - first) * 8u;
                    const double* src = a.frame + run[k].x;
                    double* dst = stage + first;
                    bulk_copy(dst, src, bytes, &arrived);
                    bulk_copy(dst + STAGE_SLOTS, src + a.frame_stride, bytes, &arrived);
                    bulk_copy(dst + 2 * STAGE_SLOTS, src + 2 * a.frame_stride, bytes, &arrived);
                    if ((run[k].y >> 16) != 0) {
                        // seen through a periodic boundary: remembered for the fix-up below
                        const int position = atomicAdd(&nimages, 1);
                        images[position] = make_int4(first, end[k], run[k].y >> 16, 0);
                    }
                }
            }
        }
        barrier_wait(&arrived, 0);
        if (header.w < 0) {
            // add the image vector to what the copies brought for cells across a periodic boundary
            __syncthreads();
            const int lane = tid & 31;
            for (int m = tid >> 5; m < nimages; m += LJ_THREADS / 32) {
                const int4 image = images[m];
                const int ix = (image.z << 30) >> 30, iy = (image.z << 28) >> 30, iz = (image.z << 26) >> 30;
                const double sx = (double)ix * a.length[0], sy = (double)iy * a.length[1], sz = (double)iz * a.length[2];
                for (int t = image.x + lane; t < image.y; t += 32) {
                    stage[t] += sx;
                    stage[STAGE_SLOTS + t] += sy;
                    stage[2 * STAGE_SLOTS + t] += sz;
                }
            }
            __syncthreads();
        }

        const int nwords = (count + 7) >> 3;
        if (nwords > 0 && mine) {
            const double xi = stage[self], yi = stage[STAGE_SLOTS + self], zi = stage[2 * STAGE_SLOTS + self];
            // eight 16-bit entries per 16-byte word; the word two iterations ahead is in flight
            for (int w = 0; w < nwords; w++) {
                uint4 wafter = make_uint4(0, 0, 0, 0);
                if (w + 2 < nwords) wafter = words[(w + 2) * 32];
                const unsigned pairs[4] = {wcur.x, wcur.y, wcur.z, wcur.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    staged_lj<MODE>(a, stage, pairs[t] & 0xffffu, xi, yi, zi, fx, fy, fz, acc);
                    staged_lj<MODE>(a, stage, pairs[t] >> 16, xi, yi, zi, fx, fy, fz, acc);
                }
                wcur = wnext;
                wnext = wafter;
            }
            fx *= a.lj_inv_sigma;
            fy *= a.lj_inv_sigma;
            fz *= a.lj_inv_sigma;
        }
    }
    if (mine && header.y > 0 && a.write_forces) {
        a.force[3 * orig] = fx;
        a.force[3 * orig + 1] = fy;
        a.force[3 * orig + 2] = fz;
    }

    if (MODE == NL_MODE_FULL) {
#pragma unroll
        for (int k = 0; k < NL_NV; k++) acc[k] *= 0.5;
        __syncthreads();
        block_sum<NL_NV>(acc, stage);  // the staged copy is dead: its memory holds the reduction scratch
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NL_NV; k++) a.partials[(size_t)blockIdx.x * NL_NV + k] = acc[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------

__global__ void set_flag_kernel(int* flags, int index, int value) { flags[index] = value; }

static uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
}

int launch_pairs_cells(Context* ctx, const ComputeRequest& req) {
    const int n = (int)ctx->n;
    // single Lennard-Jones interaction, no charges in play: the pipelined kernel of pairs_lj2.cu (neighbour path 2 keeps
    // the first-generation kernels with every block in the global list format, which the tests compare against)
    if (ctx->any_pair && ctx->single_lj && ctx->coulomb.kind == 0 && ctx->forced_path != 2 && lj2_enabled(ctx)) {
        ctx->lj2_active = true;
        return launch_pairs_lj2(ctx, req);
    }
    // water-like charged systems: staged, branch-free Lennard-Jones + Ewald / Wolf kernel of pairs_lj2.cu
    if (cq_applicable(ctx)) {
        ctx->lj2_active = true;
        return launch_pairs_cq(ctx, req);
    }
    ctx->lj2_active = false;
    if (ctx->n >= (int64_t)LIST_INDEX_MASK) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "the neighbour list handles at most %u atoms per GPU", LIST_INDEX_MASK);
    }
    GridView g;
    for (int d = 0; d < 3; d++) {
        g.nc[d] = ctx->ncell[d];
        g.length[d] = ctx->cell.h[4 * d];
        g.edge[d] = g.length[d] / (double)g.nc[d];
    }
    const int ncells = g.nc[0] * g.nc[1] * g.nc[2];
    int64_t o_lo, o_hi;
    ctx->owned_range(ctx->n, o_lo, o_hi);

    // the list covers every interaction family so that it does not depend on which parts are requested
    double cutoff = ctx->any_pair ? ctx->max_pair_cutoff : 0.0;
    if (ctx->coulomb.kind != 0 && ctx->coulomb.rc > cutoff) cutoff = ctx->coulomb.rc;
    const double skin = ctx->skin_effective;
    const double radius = cutoff + skin;

    // ---- buffers ---------------------------------------------------------------------------------
    const double volume = g.length[0] * g.length[1] * g.length[2];
    const double mean_neighbors = 4.0 / 3.0 * PI * radius * radius * radius * (double)n / volume;
    int capacity = (int)(2.0 * mean_neighbors) + 64;
    if (capacity > n) capacity = n;
    capacity = (capacity + 7) / 8 * 8;  // whole 16-byte words in both list formats
    const size_t stride = ((size_t)n + 31) / 32 * 32;
    LUMOL_CUDA_CHECK(ctx, ctx->nl_flags.reserve(16));
    LUMOL_CUDA_CHECK(ctx, ctx->nlist.reserve(stride * (size_t)capacity));
    LUMOL_CUDA_CHECK(ctx, ctx->ncount.reserve(stride));
    LUMOL_CUDA_CHECK(ctx, ctx->xref.reserve((size_t)3 * n));
    LUMOL_CUDA_CHECK(ctx, ctx->rel0.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_of.reserve((size_t)2 * n));  // cell_of + slot_of
    LUMOL_CUDA_CHECK(ctx, ctx->cell_count.reserve((size_t)ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_start.reserve((size_t)ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_needed.reserve((size_t)ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->order.reserve((size_t)2 * n));  // order + grouped
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_pos.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_f32.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_info.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_cell.reserve((size_t)n));
    const size_t frame_stride = ((size_t)n + (size_t)ncells + 2 + 31) / 32 * 32;
    LUMOL_CUDA_CHECK(ctx, ctx->frame_pos.reserve(3 * frame_stride));
    LUMOL_CUDA_CHECK(ctx, ctx->self_local.reserve(stride));
    const int nblocks = (n + TB - 1) / TB;
    LUMOL_CUDA_CHECK(ctx, ctx->blk_header.reserve((size_t)nblocks));
    LUMOL_CUDA_CHECK(ctx, ctx->blk_entries.reserve((size_t)nblocks * STAGE_MAX_ENTRIES));
    LUMOL_CUDA_CHECK(ctx, ctx->blk_runs.reserve((size_t)nblocks * STAGE_MAX_ENTRIES + 1));
    // Lennard-Jones fast path (one LJ interaction, no restriction, no charges in play): staged blocks
    const bool lj_system = ctx->any_pair && ctx->single_lj && ctx->coulomb.kind == 0;
    const bool allow_staging = lj_system && ctx->forced_path != 2;
    // the staged kernel works in units of sigma (one multiplication less per pair)
    const double frame_scale = lj_system ? 1.0 / ctx->host_pairs[0].p[0] : 1.0;
    const int stage_bytes = STAGE_BYTES;
    const int stage_atoms_max = STAGE_ATOMS_MAX;
    const int scan_blocks = (ncells + SCAN_BLOCK - 1) / SCAN_BLOCK;
    LUMOL_CUDA_CHECK(ctx, ctx->scan_scratch.reserve((size_t)scan_blocks + 1));
    int* flags = ctx->nl_flags.ptr;
    int* cell_of = ctx->cell_of.ptr;
    int* slot_of = ctx->cell_of.ptr + n;
    int* order = ctx->order.ptr;
    int* grouped = ctx->order.ptr + n;

    // ---- is the current list still describing this system? ------------------------------------------
    uint64_t signature = mix(0x1234, (uint64_t)n);
    signature = mix(signature, ctx->cell_generation);
    signature = mix(signature, ctx->structure_generation);
    uint64_t bits;
    std::memcpy(&bits, &radius, sizeof(bits));
    signature = mix(signature, bits);
    signature = mix(signature, (uint64_t)o_lo * 1315423911ull + (uint64_t)o_hi);
    signature = mix(signature, (uint64_t)capacity);
    signature = mix(signature, lj_system ? 1 : 0);
    const bool reuse = ctx->list_valid && signature == ctx->list_signature;
    const int blocks = (n + 255) / 256;
    // a rebuild is requested by writing the epoch of the evaluation into flags[FLAG_REBUILD] (never 0)
    ctx->list_epoch = ctx->list_epoch % 1000000000 + 1;
    const int epoch = ctx->list_epoch;
    {
        ScopedClock clock(ctx, &ctx->clk_neighbor);
        if (!reuse) {
            if (!ctx->flags_initialised) {
                LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(flags, 0, 16 * sizeof(int), ctx->stream));
                ctx->flags_initialised = true;
            }
            LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->frame_pos.ptr, 0, 3 * frame_stride * sizeof(double), ctx->stream));
            set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_REBUILD, epoch);
        } else {
            const double half = 0.5 * skin;
            list_update_kernel<<<blocks, 256, 0, ctx->stream>>>(n, g, order, ctx->position.ptr, ctx->xref.ptr, ctx->rel0.ptr,
                                                                ctx->sorted_cell.ptr, ctx->cell_start.ptr, ctx->sorted_pos.ptr,
                                                                allow_staging ? ctx->frame_pos.ptr : nullptr, frame_stride,
                                                                frame_scale, half * half, epoch,
                                                                ctx->nranks > 1 ? ctx->cell_needed.ptr : nullptr, flags);
        }
        ctx->launches++;
        ctx->clk_neighbor.launches++;

        // ---- rebuild pipeline: does nothing unless flags[FLAG_REBUILD] holds this epoch ---------------------------
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
        s.rel0 = ctx->rel0.ptr;
        s.sorted_f32 = ctx->sorted_f32.ptr;
        s.sorted_info = ctx->sorted_info.ptr;
        s.sorted_cell = ctx->sorted_cell.ptr;
        s.frame = ctx->frame_pos.ptr;
        s.frame_stride = frame_stride;
        s.frame_scale = frame_scale;
        s.xref = ctx->xref.ptr;

        TableArgs t;
        t.g = g;
        t.n = n;
        t.nblocks = nblocks;
        t.allow = allow_staging ? 1 : 0;
        t.stage_atoms_max = stage_atoms_max;
        t.cell_start = ctx->cell_start.ptr;
        t.sorted_cell = ctx->sorted_cell.ptr;
        t.header = ctx->blk_header.ptr;
        t.entries = ctx->blk_entries.ptr;
        t.runs = ctx->blk_runs.ptr;
        t.sorted_info = ctx->sorted_info.ptr;
        t.o_lo = (int)o_lo;
        t.o_hi = (int)o_hi;
        t.flags = flags;

        BuildArgs b;
        b.g = g;
        b.ncells = ncells;
        // one warp per `cells_per_warp` consecutive cells: about sixteen work items per resident warp
        b.cells_per_warp = ncells * BUILD_CHUNKS / (ctx->sm_count * 8 * BUILD_WARPS * 16) + 1;
        b.o_lo = (int)o_lo;
        b.o_hi = (int)o_hi;
        b.capacity = capacity;
        b.blk_header = ctx->blk_header.ptr;
        b.blk_entries = ctx->blk_entries.ptr;
        b.self_local = ctx->self_local.ptr;
        b.radius2 = (float)(radius * radius * 1.0001);
        b.cell_start = ctx->cell_start.ptr;
        b.sorted_f32 = ctx->sorted_f32.ptr;
        b.sorted_info = ctx->sorted_info.ptr;
        b.nlist = ctx->nlist.ptr;
        b.ncount = ctx->ncount.ptr;
        b.cell_needed = ctx->cell_needed.ptr;
        b.flags = flags;
        RebuildArgs r;
        r.n = n;
        r.ncells = ncells;
        r.scan_blocks = scan_blocks;
        r.nblocks = nblocks;
        r.epoch = epoch;
        r.g = g;
        r.position = ctx->position.ptr;
        r.cell_of = cell_of;
        r.slot_of = slot_of;
        r.cell_count = ctx->cell_count.ptr;
        r.cell_start = ctx->cell_start.ptr;
        r.scan_scratch = ctx->scan_scratch.ptr;
        r.grouped = grouped;
        r.scatter = s;
        r.table = t;
        r.build = b;
        r.flags = flags;
        if (ctx->rebuild_grid == 0) {
            int per_sm = 0;
            LUMOL_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rebuild_kernel, REBUILD_THREADS, 0));
            if (per_sm < 1) return ctx->fail(LUMOL_CUDA_ERROR_CUDA, "the rebuild kernel does not fit on the device");
            ctx->rebuild_grid = per_sm * ctx->sm_count;
        }
        {
            void* params[] = {&r};
            LUMOL_CUDA_CHECK(ctx, cudaLaunchCooperativeKernel((const void*)rebuild_kernel, dim3(ctx->rebuild_grid),
                                                              dim3(REBUILD_THREADS), params, 0, ctx->stream));
        }
        ctx->launches++;
        ctx->clk_neighbor.launches++;
        const size_t reorder_smem = (size_t)capacity * REORDER_THREADS * sizeof(unsigned short);
        static const bool reorder_disabled = std::getenv("LUMOL_CUDA_NO_REORDER") != nullptr;  // profiling switch
        if (allow_staging && reorder_smem <= 200 * 1024 && !reorder_disabled) {
            if (reorder_smem > 40 * 1024) {
                LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(list_reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           (int)reorder_smem));
            }
            const int slabs = (n + REORDER_THREADS - 1) / REORDER_THREADS;
            const int reorder_grid = slabs < ctx->sm_count * 10 ? slabs : ctx->sm_count * 10;
            list_reorder_kernel<<<reorder_grid, REORDER_THREADS, reorder_smem, ctx->stream>>>(
                n, capacity, ctx->blk_header.ptr, ctx->ncount.ptr, ctx->nlist.ptr, epoch, flags);
            ctx->launches++;
            ctx->clk_neighbor.launches++;
        }
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }
    ctx->list_valid = true;
    ctx->list_signature = signature;

    // ---- forces ------------------------------------------------------------------------------------------
    const bool do_pairs = req.pairs && ctx->any_pair;
    const bool do_coulomb = req.coulomb && ctx->coulomb.kind != 0;
    double active_cutoff = do_pairs ? ctx->max_pair_cutoff : 0.0;
    if (do_coulomb && ctx->coulomb.rc > active_cutoff) active_cutoff = ctx->coulomb.rc;

    ForceArgs a;
    a.n = n;
    a.o_lo = (int)o_lo;
    a.o_hi = (int)o_hi;
    a.capacity = capacity;
    for (int d = 0; d < 3; d++) a.edge[d] = g.edge[d];
    a.nlist = ctx->nlist.ptr;
    a.ncount = ctx->ncount.ptr;
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
    a.cutoff2 = active_cutoff * active_cutoff;
    a.force = ctx->force.ptr;
    a.write_forces = req.forces;
    a.flags = flags;

    // the staged kernel needs lists built for it (lj_system); a system with charges whose coulomb part is
    // not requested still takes the general kernel
    const bool lj_only = do_pairs && !do_coulomb && lj_system;
    a.lj_sigma2 = a.lj_epsilon24 = a.lj_epsilon48 = a.lj_epsilon4 = a.lj_cutoff2 = a.lj_shift = 0.0;
    a.lj_inv_sigma = 1.0;
    a.lj_reduced_cutoff2_bits = 0;
    if (lj_only) {
        const lumol_cuda_pair& p = ctx->host_pairs[0];
        a.lj_sigma2 = p.p[0] * p.p[0];
        a.lj_epsilon24 = 24.0 * p.p[1];
        a.lj_epsilon48 = 48.0 * p.p[1];
        a.lj_epsilon4 = 4.0 * p.p[1];
        a.lj_cutoff2 = p.cutoff * p.cutoff;
        a.lj_inv_sigma = frame_scale;
        const double reduced_cutoff = p.cutoff * frame_scale;
        const double reduced_cutoff2 = reduced_cutoff * reduced_cutoff;
        std::memcpy(&a.lj_reduced_cutoff2_bits, &reduced_cutoff2, sizeof(double));
        a.lj_shift = p.shift;
    }
    a.blk_header = ctx->blk_header.ptr;
    a.blk_entries = ctx->blk_entries.ptr;
    a.blk_runs = ctx->blk_runs.ptr;
    a.self_local = ctx->self_local.ptr;
    a.frame = ctx->frame_pos.ptr;
    a.frame_stride = frame_stride;
    a.ntiles = nblocks;
    for (int d = 0; d < 3; d++) a.length[d] = g.length[d] * frame_scale;  // image vectors of the staged copies
    const size_t smem = lj_only ? (size_t)stage_bytes : sizeof(PairParams) * (size_t)ctx->nkinds * ctx->nkinds;
    if (!lj_only && smem > 100 * 1024) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "too many particle kinds (%d) for the shared pair table", ctx->nkinds);
    }
    const bool full = req.energy || req.virial;
    const int threads = lj_only ? LJ_THREADS : NL_THREADS;
    // general kernel: up to four threads per atom while that does not exceed about a dozen warps per SM slot
    int split = 1;
    if (!lj_only) {
        while (split < 4 && (int64_t)n * split * 2 <= (int64_t)ctx->sm_count * 2048 * 4) split *= 2;
    }
    const int force_blocks = lj_only ? (n + threads - 1) / threads : (int)(((int64_t)n * split + threads - 1) / threads);
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)force_blocks * NL_NV));
    a.partials = ctx->partials.ptr;

    const void* kernel;
    if (lj_only) {
        kernel = full ? (const void*)lj_force_kernel<NL_MODE_FULL> : (const void*)lj_force_kernel<NL_MODE_FORCES>;
    } else {
        // forces-only evaluations (every MD step) are capped at 128 registers: four resident blocks per SM instead
        // of three hide more of the erfc / exp latency (1.42 -> 1.33 ms on the 98k-atom SPC/E box, 64 bytes of spills)
        // SIMPLE: pair tables of Lennard-Jones / harmonic / null entries only (smaller kernel); experiment knob to
        // fall back to the general instance
        static const bool general_only = std::getenv("LUMOL_CUDA_LIST_GENERAL") != nullptr;
        const bool simple = ctx->simple_pairs && !general_only;
#define LUMOL_LIST_KERNEL(S) \
    (full ? (simple ? (const void*)list_force_kernel<NL_MODE_FULL, S, 1, true> : (const void*)list_force_kernel<NL_MODE_FULL, S, 1, false>) \
          : (simple ? (const void*)list_force_kernel<NL_MODE_FORCES, S, 4, true> : (const void*)list_force_kernel<NL_MODE_FORCES, S, 4, false>))
        if (split == 4) {
            kernel = LUMOL_LIST_KERNEL(4);
        } else if (split == 2) {
            kernel = LUMOL_LIST_KERNEL(2);
        } else {
            kernel = LUMOL_LIST_KERNEL(1);
        }
#undef LUMOL_LIST_KERNEL
    }
    if (smem > 40 * 1024) {
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    {
        ScopedClock clock(ctx, &ctx->clk_pair);
        void* params[] = {&a};
        LUMOL_CUDA_CHECK(ctx, cudaLaunchKernel(kernel, dim3(force_blocks), dim3(threads), params, smem, ctx->stream));
        ctx->launches++;
        ctx->clk_pair.launches++;
    }
    if (full) {
        int status = launch_reduce(ctx, force_blocks, NL_NV, RES_E_PAIRS);
        if (status != 0) return status;
    }
    return 0;
}

// Reads the device-side list counters (synchronises the stream).
int neighbor_list_status(Context* ctx, int* rebuilds, int* overflow) {
    *rebuilds = 0;
    *overflow = 0;
    if (ctx->nl_flags.ptr == nullptr || !ctx->flags_initialised) return 0;
    int host[12] = {0};
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(host, ctx->nl_flags.ptr, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    *overflow = host[FLAG_NONFINITE] != 0 ? 2 : host[FLAG_OVERFLOW];
    // pairs_lj2.cu: frame slots of the ghost images, list of the pairs sitting on the cut-off
    if (ctx->lj2_active && *overflow == 0 && (host[11] != 0 || host[10] > ctx->deferred_capacity)) *overflow = 1;
    if (host[FLAG_NONFINITE] != 0) ctx->list_valid = false;  // the next evaluation rebuilds from scratch
    *rebuilds = host[FLAG_COUNT];
    return 0;
}

}  // namespace lumol
