// Pipelined Lennard-Jones pair path for single-LJ systems on the neighbour-list path (the headline kernel of the
// benchmark: liquid argon, sys/compute.rs:37-55 + energy.rs:47-59 + compute.rs:202-216 of the reference for one
// PairInteraction with restriction None).  Second generation of the staged kernel of pairs_cells.cu; what changed and
// why (profiles/r1m_*, VERDICT round 1: FP64 pipe 65 % busy, 30 % of the warp time outside the pair loop, 1.33 listed
// entries per pair inside the cut-off):
//
//   * GHOST CELLS.  The box-frame position planes carry one layer of periodic ghost cells around the grid
//     ((nx+2)(ny+2)(nz+2) cells, the refresh kernel writes up to seven images of a boundary atom).  Every row of a
//     unit's neighbourhood is then ONE contiguous range: one bulk copy per row and plane, no image fix-up pass.
//   * PERSISTENT, PIPELINED.  One CTA of sixteen warps per SM and two staging buffers: while unit k is evaluated out
//     of one, the bulk copies (cp.async.bulk -> mbarrier) of unit k + 1 land in the other.  A warp that finishes a unit
//     goes straight on to the next one; the LAST warp to finish unit k issues the copies of unit k + 2 into the buffer
//     that has just drained.  No block-wide barrier, no warp set aside as a producer (544 threads would cap the
//     kernel at 96 registers).
//   * TWO LANES PER ATOM.  Lanes 2a, 2a+1 share atom a and split every 8-entry list word 4 + 4: sixteen warps work on one
//     256-atom unit, no imbalance between the halves.
//   * RADIAL LEVELS.  At build time every list entry gets the level of its distance, r_b in [rc + k d, rc + (k+1) d),
//     d = skin / 8, and the columns are ordered by level.  A pair can only be inside the cut-off when
//     r_b < rc + 2 dmax(t), dmax = largest displacement of any atom since the rebuild (the refresh kernel reduces it), so
//     an evaluation walks only the prefix of the levels below 2 dmax / d: 1.06 listed per in-cut-off pair right after a
//     rebuild, 1.33 just before the next, no re-pruning pass and no second list.
//   * EXACT CUT-OFF.  A pair whose r^2 falls in a narrow band around rc^2 (top 32 bits equal +-1) is not evaluated here
//     but appended to a short list; a fix-up kernel evaluates those few pairs from the original coordinates with the
//     reference's own arithmetic (minimum image by division and round, r = sqrt(d.d), `r >= rc -> 0`, pairs.rs:186), so
//     the set of interacting pairs is the reference's bit for bit whatever rounding the box-frame arithmetic did.
//
// Every listed pair is evaluated from both sides (full shell): with 17 FP64 instructions per pair, moving a force
// through shared memory or shuffles (24 bytes each way at 128 B/clk/SM) costs more than recomputing it (17 / 64 clk/SM),
// see DESIGN.md section 4.1.
#include "context.hpp"
#include "neighbor_common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <climits>

#include <cstdlib>
#include <cstring>

namespace lumol {

#ifndef LJ2_THREADS_CONFIG
#define LJ2_THREADS_CONFIG 512
#endif
#ifndef LJ2_BUFFERS_CONFIG
#define LJ2_BUFFERS_CONFIG 2
#endif
constexpr int LJ2_THREADS = LJ2_THREADS_CONFIG;
#ifndef LJ2_ILP8
#define LJ2_ILP8 0
#endif
#ifndef LJ2_LPA_CONFIG
#define LJ2_LPA_CONFIG 2
#endif
constexpr int LJ2_LPA = LJ2_LPA_CONFIG;                         // lanes per atom: they split every 8-entry list word 4 + 4 (and alternate words when 4)
constexpr int U_ATOMS = LJ2_THREADS / LJ2_LPA;     // atoms per unit (consecutive in cell order)
constexpr int LJ2_BUFFERS = LJ2_BUFFERS_CONFIG;
constexpr int LJ2_WORD_STRIDE = LJ2_LPA / 2;       // a lane walks words word_first, word_first + LJ2_WORD_STRIDE, ...
constexpr int LJ2_MAX_SEGMENTS = 6;
constexpr int LJ2_MAX_RUNS = 9 * LJ2_MAX_SEGMENTS;
constexpr int LJ2_SLOTS = LJ2_BUFFERS == 2 ? 4528 : 3072;  // atoms of one staged copy (dummy slots 0, 1 included)
constexpr int LJ2_STAGE_BYTES = 3 * LJ2_SLOTS * 8;  // (x, y, z) per slot: 106 KiB, two buffers per CTA
constexpr int LJ2_LEVELS = 8;
constexpr int LJ2_NV = 16;
constexpr unsigned RAW_VALUE_MASK = (1u << 28) - 1u;
constexpr int FLAG_DISP = 8;       // flags[8], flags[9]: float bits of the largest squared displacement, by epoch parity
constexpr int FLAG_DEFERRED = 10;  // number of deferred (on-the-cut-off) pairs of the current evaluation
constexpr int FLAG_FRAME_OVERFLOW = 11;
constexpr int DEFERRED_PER_ATOM = 4;

__host__ __device__ __forceinline__ int pad2(int count) { return (count + 1) & ~1; }

// What a force kernel expects of the units and lists the rebuild makes for it (two kernels share the rebuild: the
// Lennard-Jones kernel above and the charged-system kernel at the end of this file).
struct UnitShape {
    int atoms;           // atoms per unit
    int lanes_per_atom;  // lanes sharing a column (2, 4, ...): they take the halves of its words in turn
    int slots;           // capacity of one staged copy
    int value_factor;    // staged entries hold slot * value_factor (doubles before the slot's x in an (x, y, z) copy: 3)
    int width;           // doubles per frame slot: 3 (x, y, z) or 4 (x, y, z, charge)
    int flags;           // bit 0: entries carry the neighbour's kind (2 bits) and a same-molecule bit above 13 bits of slot;
                         // bit 1: the force kernel reads the raw 32-bit columns of the build (no level / bank order pass)
                         // bit 2: units never cross a row of cells (variable size, unit_start[]): every unit is one segment
};

constexpr int FLAG_UNITS = 13;  // number of units when they are row-aligned (known on the device only)
constexpr int FLAG_HALO_ITEMS = 14;  // sharded runs: number of (cell image, peers) runs the halo kernel copies

// unit of sorted atom s when units are row-aligned
__device__ __forceinline__ int unit_of_atom(int s, int cell, int nx, int unit_atoms, const int* __restrict__ cell_start,
                                            const int* __restrict__ row_units) {
    const int row = cell / nx;
    return row_units[row] + (s - cell_start[row * nx]) / unit_atoms;
}

// ------------------------------------------------------------------------------------------------
// rebuild phases specific to this path
// ------------------------------------------------------------------------------------------------

struct ExtGrid {
    int nx, ny, nz;  // real cells
    int ex, ey, ez;  // with the ghost layer
    __host__ __device__ int count() const { return ex * ey * ez; }
    // real cell seen at extended coordinates (X, Y, Z)
    __device__ __forceinline__ int source(int X, int Y, int Z) const {
        const int x = X == 0 ? nx - 1 : (X == ex - 1 ? 0 : X - 1);
        const int y = Y == 0 ? ny - 1 : (Y == ey - 1 ? 0 : Y - 1);
        const int z = Z == 0 ? nz - 1 : (Z == ez - 1 ? 0 : Z - 1);
        return (z * ny + y) * nx + x;
    }
    __device__ __forceinline__ int index(int X, int Y, int Z) const { return (Z * ey + Y) * ex + X; }
};

struct Scatter2Args {
    int n;
    GridView g;
    const double* __restrict__ pos;   // state order
    const int* __restrict__ state_of_sorted_old;  // unused here (kept for the sorted-resident mode)
    const int* __restrict__ cell_of;
    const int* __restrict__ cell_start;
    const int* __restrict__ grouped;
    int* __restrict__ order;          // sorted slot -> state index
    float4* __restrict__ sorted_f32;  // cell-relative FP32 position (list build)
    int* __restrict__ sorted_cell;
    double* __restrict__ xref;        // 3 n, sorted order: position at the rebuild
    int* __restrict__ kshift;         // n: wrap counts floor(x / L) packed 3 x 10 bits + sign offset
    // charged systems: (first atom of the molecule << 2) | kind rides in the w component of sorted_f32
    const unsigned* __restrict__ kind;
    const int* __restrict__ mol_first;
};

// wrap counts in [-512, 511] per axis
__device__ __forceinline__ int pack_shift(int kx, int ky, int kz) {
    return ((kx + 512) & 1023) | (((ky + 512) & 1023) << 10) | (((kz + 512) & 1023) << 20);
}
__device__ __forceinline__ void unpack_shift(int packed, int& kx, int& ky, int& kz) {
    kx = (packed & 1023) - 512;
    ky = ((packed >> 10) & 1023) - 512;
    kz = ((packed >> 20) & 1023) - 512;
}

// rank inside the cell = number of cell mates with a smaller state index (deterministic order)
__device__ __forceinline__ void scatter2_phase(int vb, const Scatter2Args& a, int* __restrict__ flags) {
    const int s = vb * REBUILD_THREADS + threadIdx.x;
    if (s >= a.n) return;
    const int i = a.grouped[s];
    const int c = a.cell_of[i];
    const int lo = a.cell_start[c], hi = a.cell_start[c + 1];
    int rank = 0;
    for (int t = lo; t < hi; t++) rank += a.grouped[t] < i ? 1 : 0;
    const int dst = lo + rank;
    a.order[dst] = i;
    const int cx = c % a.g.nc[0], cy = (c / a.g.nc[0]) % a.g.nc[1], cz = c / (a.g.nc[0] * a.g.nc[1]);
    const double px = a.pos[3 * i], py = a.pos[3 * i + 1], pz = a.pos[3 * i + 2];
    const double kx = floor(px / a.g.length[0]), ky = floor(py / a.g.length[1]), kz = floor(pz / a.g.length[2]);
    if (fabs(kx) > 500.0 || fabs(ky) > 500.0 || fabs(kz) > 500.0) flags[FLAG_NONFINITE] = 1;  // hundreds of boxes away: refuse
    const double x = px - kx * a.g.length[0] - ((double)cx + 0.5) * a.g.edge[0];
    const double y = py - ky * a.g.length[1] - ((double)cy + 0.5) * a.g.edge[1];
    const double z = pz - kz * a.g.length[2] - ((double)cz + 0.5) * a.g.edge[2];
    const int tag = a.kind != nullptr ? ((a.mol_first[i] << 2) | (int)(a.kind[i] & 3u)) : 0;
    a.sorted_f32[dst] = make_float4((float)x, (float)y, (float)z, __int_as_float(tag));
    a.sorted_cell[dst] = c;
    a.xref[3 * dst] = px;
    a.xref[3 * dst + 1] = py;
    a.xref[3 * dst + 2] = pz;
    a.kshift[dst] = pack_shift((int)kx, (int)ky, (int)kz);
}

// padded atom count of every extended cell (a ghost cell mirrors its source)
__device__ __forceinline__ void ext_count_phase(const ExtGrid& e, const int* __restrict__ cell_start, int* __restrict__ ext_pad) {
    const int total = e.count();
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int X = k % e.ex, Y = (k / e.ex) % e.ey, Z = k / (e.ex * e.ey);
        const int c = e.source(X, Y, Z);
        ext_pad[k] = pad2(cell_start[c + 1] - cell_start[c]);
    }
}

// last pass of the exclusive scan of the padded counts; also writes the grand total behind the last element
__device__ __forceinline__ void scan_add_total_phase(int vb, int count, const int* __restrict__ in, int* __restrict__ out,
                                                     const int* __restrict__ block_sums) {
    const int base = vb * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    const int add = block_sums[vb];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < count) {
            const int value = out[base + k] + add;
            out[base + k] = value;
            if (base + k == count - 1) out[count] = value + in[count - 1];
        }
    }
}

// Frame index of an atom's real copy and of its ghost images; `visit(frame_index, sx, sy, sz)` with the image shift in
// box lengths.  Extended cells start at ext_start[] + 2 (frame slots 0 and 1 are the far-away dummy).
template <typename Visit>
__device__ __forceinline__ void for_each_image(const ExtGrid& e, const int* __restrict__ ext_start, int c, int rank, Visit visit) {
    const int x = c % e.nx, y = (c / e.nx) % e.ny, z = c / (e.nx * e.ny);
    // ghost direction per axis: an atom of the first cell is also seen behind the last one (+L) and vice versa
    const int gx = x == 0 ? 1 : (x == e.nx - 1 ? -1 : 0);
    const int gy = y == 0 ? 1 : (y == e.ny - 1 ? -1 : 0);
    const int gz = z == 0 ? 1 : (z == e.nz - 1 ? -1 : 0);
    for (int m = 0; m < 8; m++) {
        const int sx = (m & 1) ? gx : 0, sy = (m & 2) ? gy : 0, sz = (m & 4) ? gz : 0;
        if (((m & 1) && gx == 0) || ((m & 2) && gy == 0) || ((m & 4) && gz == 0)) continue;
        const int X = sx > 0 ? e.ex - 1 : (sx < 0 ? 0 : x + 1);
        const int Y = sy > 0 ? e.ey - 1 : (sy < 0 ? 0 : y + 1);
        const int Z = sz > 0 ? e.ez - 1 : (sz < 0 ? 0 : z + 1);
        visit(2 + ext_start[e.index(X, Y, Z)] + rank, sx, sy, sz);
    }
}

struct Frame2Args {
    int n;
    GridView g;
    ExtGrid e;
    const int* __restrict__ cell_start;
    const int* __restrict__ sorted_cell;
    const int* __restrict__ ext_start;
    const int* __restrict__ order;
    const double* __restrict__ xref;
    const int* __restrict__ kshift;
    int* __restrict__ fidx;        // n: frame index of the real copy
    int* __restrict__ frame_atom;  // frame slot -> sorted index (images included)
    double* __restrict__ frame;    // (x, y, z) or (x, y, z, charge) per frame slot
    size_t fstride;                // frame slots allocated
    double scale;
    int width;
    const double* __restrict__ charge;  // caller's order (width 4)
};

// frame indices and the frames themselves at the rebuild positions
__device__ __forceinline__ void frame2_phase(int vb, const Frame2Args& a, int* __restrict__ flags) {
    const int s = vb * REBUILD_THREADS + threadIdx.x;
    if (vb == 0 && threadIdx.x < 2) {
        // the dummy that padding entries of unstaged columns point at
        for (int p = 0; p < a.width; p++) a.frame[a.width * threadIdx.x + p] = p < 3 ? 1.0e9 * (double)(p + 1) : 0.0;
        a.frame_atom[threadIdx.x] = -1;
    }
    if (s >= a.n) return;
    const int c = a.sorted_cell[s];
    const int rank = s - a.cell_start[c];
    int kx, ky, kz;
    unpack_shift(a.kshift[s], kx, ky, kz);
    const double x = a.xref[3 * s] - (double)kx * a.g.length[0];
    const double y = a.xref[3 * s + 1] - (double)ky * a.g.length[1];
    const double z = a.xref[3 * s + 2] - (double)kz * a.g.length[2];
    bool first = true;
    for_each_image(a.e, a.ext_start, c, rank, [&](int f, int sx, int sy, int sz) {
        if ((size_t)f + 2 > a.fstride) {
            flags[FLAG_FRAME_OVERFLOW] = 1;
            return;
        }
        if (first) a.fidx[s] = f;
        first = false;
        a.frame_atom[f] = s;
        a.frame[(size_t)a.width * f] = (x + (double)sx * a.g.length[0]) * a.scale;
        a.frame[(size_t)a.width * f + 1] = (y + (double)sy * a.g.length[1]) * a.scale;
        a.frame[(size_t)a.width * f + 2] = (z + (double)sz * a.g.length[2]) * a.scale;
        if (a.width == 4) a.frame[(size_t)a.width * f + 3] = a.charge[a.order[s]];
    });
}

// ---- unit tables -------------------------------------------------------------------------------------------------
//
// Unit u holds the sorted atoms [256 u, 256 u + 256): home cells [c0, c0 + K) of the x-fastest cell order, i.e. up to a
// few row segments.  Segment g (cells x in [xa, xa + len) of grid row r0 + g) brings nine runs, one per (dy, dz): the
// extended cells X in [xa, xa + len + 1] of the extended row (z + dz + 1, y + dy + 1), contiguous in the frame planes.
// Run index 9 g + 3 (dz + 1) + (dy + 1).

struct UnitRows {
    int c0, K, nx, r0, xa0, len0, nseg;
    __host__ __device__ void init(int first_cell, int cells, int cells_x) {
        c0 = first_cell;
        K = cells;
        nx = cells_x;
        r0 = c0 / nx;
        xa0 = c0 - r0 * nx;
        len0 = min(nx - xa0, K);
        nseg = 1 + (K - len0 + nx - 1) / nx;
    }
    __host__ __device__ void segment(int g, int& xa, int& len) const {
        if (g == 0) {
            xa = xa0;
            len = len0;
        } else {
            xa = 0;
            len = min(nx, K - len0 - (g - 1) * nx);
        }
    }
};

struct Table2Args {
    int n, nunits;
    UnitShape shape;
    const int* __restrict__ unit_start;  // row-aligned units: first atom of every unit (and n behind the last)
    ExtGrid e;
    const int* __restrict__ sorted_cell;
    const int* __restrict__ ext_start;
    int4* __restrict__ header;  // c0, K, number of runs (-1: not staged), staged slots
    int4* __restrict__ runs;    // first frame index, slots, first slot
    int* __restrict__ flags;
    // sharded runs: need_mask[c] collects the ranks whose units stage cell c (or one of its periodic images), so that the
    // owner of an atom knows into which GPUs' frames it has to store the atom's new position
    int* __restrict__ need_mask;
    int units_per_rank;
};

// one warp per unit
__device__ __forceinline__ void unit_table_phase(int vb, const Table2Args& a) {
    const int lane = threadIdx.x & 31;
    const int unit = vb * REBUILD_WARPS + (threadIdx.x >> 5);
    if (unit >= a.nunits) return;
    int s_first = unit * a.shape.atoms, s_last = min(a.n, s_first + a.shape.atoms) - 1;
    if (a.shape.flags & 4) {
        if (unit >= a.flags[FLAG_UNITS]) return;
        s_first = a.unit_start[unit];
        s_last = a.unit_start[unit + 1] - 1;
    }
    const int c_first = a.sorted_cell[s_first], c_last = a.sorted_cell[s_last];
    UnitRows rows;
    rows.init(c_first, c_last - c_first + 1, a.e.nx);
    bool staged = rows.nseg <= LJ2_MAX_SEGMENTS;
    int total = 2;
    const int nruns = 9 * rows.nseg;
    if (staged) {
        for (int base = 0; base < nruns; base += 32) {
            const int r = base + lane;
            int f0 = 0, slots = 0;
            if (r < nruns) {
                const int g = r / 9, plane = r - 9 * g;
                int xa, len;
                rows.segment(g, xa, len);
                const int row = rows.r0 + g;
                const int y = row % a.e.ny, z = row / a.e.ny;
                const int Y = y + (plane % 3), Z = z + (plane / 3);  // dy + 1 + y, dz + 1 + z
                const int e0 = a.e.index(xa, Y, Z), e1 = e0 + len + 1;
                f0 = 2 + a.ext_start[e0];
                slots = 2 + a.ext_start[e1 + 1] - f0;
            }
            int scan = slots;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, scan, o);
                if (lane >= o) scan += t;
            }
            if (r < nruns) a.runs[(size_t)unit * LJ2_MAX_RUNS + r] = make_int4(f0, slots, total + scan - slots, 0);
            total += __shfl_sync(0xffffffffu, scan, 31);
            if (a.need_mask != nullptr && r < nruns) {
                const int g = r / 9, plane = r - 9 * g;
                int xa, len;
                rows.segment(g, xa, len);
                const int row = rows.r0 + g;
                const int Y = row % a.e.ny + (plane % 3), Z = row / a.e.ny + (plane / 3);
                const int bit = 1 << (unit / a.units_per_rank);
                for (int X = xa; X <= xa + len + 1; X++) atomicOr(a.need_mask + a.e.source(X, Y, Z), bit);
            }
        }
        if (total > a.shape.slots) staged = false;
    }
    if (!staged && lane == 0) atomicAdd(a.flags + FLAG_UNSTAGED, 1);
    if (lane == 0) a.header[unit] = make_int4(rows.c0, rows.K, staged ? nruns : -1, total);
}

// ---- list build -----------------------------------------------------------------------------------------------------

struct Build2Args {
    GridView g;
    ExtGrid e;
    UnitShape shape;
    const int* __restrict__ row_units;  // row-aligned units: first unit of every row of cells
    const int* __restrict__ order;  // sorted slot -> state index
    int o_lo, o_hi;                 // state-index range of the atoms this rank owns (first-generation sharding of the charged path)
    int ncells;
    int cells_per_warp;
    int s_lo, s_hi;      // sorted range of the atoms this rank owns (lists are built for those)
    int capacity;        // entries per atom
    float radius2;       // (cut-off + skin)^2, enlarged by 1e-4 relative
    float cutoff, inv_delta;  // level of an entry: floor((r_b - cutoff) / delta), delta = skin / 8
    const int* __restrict__ cell_start;
    const int* __restrict__ ext_start;
    const float4* __restrict__ sorted_f32;
    const int4* __restrict__ header;
    const int4* __restrict__ runs;
    unsigned short* __restrict__ self_slot;
    unsigned* __restrict__ nlist;  // raw columns: 32-bit entries (level << 28) | slot or frame index
    int* __restrict__ ncount;
    unsigned char* __restrict__ cell_needed;
    int* __restrict__ flags;
};

constexpr int BUILD2_CHUNKS = 8;

// MUFU.RSQ without the denormal pre-scaling of rsqrtf (squared distances of distinct atoms are far from denormal)
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One warp per 32-atom chunk of a home cell, one lane per atom; the 27 neighbour cells are streamed, eight candidates
// (uniform addresses) loaded before the first is tested.  Survivors are collected in a per-lane shared-memory buffer and
// written as whole 16-byte words (four entries).
template <int FLAGS>
__device__ __forceinline__ void list_build2_phase(int vb, const Build2Args& a, const float (*offset32)[3], unsigned (*pending)[4]) {
    const int lane = threadIdx.x & 31;
    const int global_warp = vb * REBUILD_WARPS + (threadIdx.x >> 5);
    const int item_lo = global_warp * a.cells_per_warp;
    const int item_hi = min(a.ncells * BUILD2_CHUNKS, item_lo + a.cells_per_warp);
    unsigned* mine = pending[threadIdx.x];
    const float level_origin = -a.cutoff * a.inv_delta;

    for (int item = item_lo; item < item_hi; item++) {
        const int first_chunk = item / a.ncells, c = item - first_chunk * a.ncells;
        const int hs = a.cell_start[c], he = a.cell_start[c + 1];
        if (hs + 32 * first_chunk >= he) continue;
        const int cx = c % a.g.nc[0];
        const int cy = (c / a.g.nc[0]) % a.g.nc[1];
        const int cz = c / (a.g.nc[0] * a.g.nc[1]);
        for (int base = hs + 32 * first_chunk; base < he; base += 32 * BUILD2_CHUNKS) {
            const int s_i = base + lane;
            bool active = s_i < he && s_i >= a.s_lo && s_i < a.s_hi;
            if (active && a.o_hi > a.o_lo) {
                const int origin = a.order[s_i];
                active = origin >= a.o_lo && origin < a.o_hi;
            }
            float xf = 1.0e18f, yf = 0.0f, zf = 0.0f;
            int tag_i = 0;
            bool staged = false;
            const int4* runs = a.runs;
            int run_base = 0;
            if (s_i < he) {
                const int unit = (FLAGS & 4) ? unit_of_atom(s_i, c, a.g.nc[0], a.shape.atoms, a.cell_start, a.row_units) : s_i / a.shape.atoms;
                const int4 header = a.header[unit];
                staged = header.z >= 0;
                runs += (size_t)unit * LJ2_MAX_RUNS;
                run_base = 9 * (c / a.g.nc[0] - header.x / a.g.nc[0]);
                if (staged) {
                    const int4 run = runs[run_base + 4];  // dy = dz = 0
                    const int f_self = 2 + a.ext_start[a.e.index(cx + 1, cy + 1, cz + 1)] + (s_i - hs);
                    a.self_slot[s_i] = (unsigned short)(a.shape.value_factor * (run.z + (f_self - run.x)));
                }
            }
            if (active) {
                const float4 f = a.sorted_f32[s_i];
                xf = f.x;
                yf = f.y;
                zf = f.z;
                tag_i = __float_as_int(f.w);
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
            // raw column of atom s: word w (four 32-bit entries) at ((s >> 5) * (capacity / 4) + w) * 32 + (s & 31)
            uint4* column = reinterpret_cast<uint4*>(a.nlist) + (size_t)(s_i >> 5) * (a.capacity >> 2) * 32 + (s_i & 31);
            int count = 0;
            for (int row = 0; row < 9; row++) {
                const int uy = cy + (row % 3) - 1, uz = cz + (row / 3) - 1;
                int ny = uy, nz = uz;
                ny += ny < 0 ? a.g.nc[1] : 0;
                ny -= ny >= a.g.nc[1] ? a.g.nc[1] : 0;
                nz += nz < 0 ? a.g.nc[2] : 0;
                nz -= nz >= a.g.nc[2] ? a.g.nc[2] : 0;
                const int row_base = (nz * a.g.nc[1] + ny) * a.g.nc[0];
                int4 run = make_int4(0, 0, 0, 0);
                if (staged) run = runs[run_base + row];
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const int code = row * 3 + dx;
                    int nx = cx + dx - 1;
                    nx += nx < 0 ? a.g.nc[0] : 0;
                    nx -= nx >= a.g.nc[0] ? a.g.nc[0] : 0;
                    const int s0 = a.cell_start[row_base + nx], s1 = a.cell_start[row_base + nx + 1];
                    const float xr = xf - offset32[code][0], yr = yf - offset32[code][1], zr = zf - offset32[code][2];
                    // value stored for neighbour s_j: its slot in the staged copy times three (the copy holds (x, y, z)
                    // triples: the offset of its x in doubles), or its frame index
                    const int frame_first = 2 + a.ext_start[a.e.index(cx + dx, uy + 1, uz + 1)];  // unwrapped: the image seen from here
                    const int tag = (staged ? run.z + (frame_first - run.x) : frame_first) - s0;
                    const int factor = staged ? a.shape.value_factor : 1;
                    // eight candidates per round, loaded before the first one is tested (uniform addresses: one broadcast each).  A
                    // variant that tested the eight branch-free into a bit mask and sent only the survivors through the divergent
                    // part (re-loading them) was slower: 2.72 against 2.23 ms for the 1M-atom box
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
                                    // r2 * rsqrt(r2) is within a few ulp of the root: the levels are walked with a margin of
                                    // 2e-3 A (Lj2Args::margin), five orders above that
                                    int level = (int)floorf(fmaf(r2 * rsqrt_approx(r2), a.inv_delta, level_origin));
                                    level = max(0, min(LJ2_LEVELS - 1, level));
                                    unsigned value = (unsigned)(factor * (tag + s_j));
                                    if (FLAGS & 1) {
                                        // neighbour's kind and "same molecule" above the slot (13 bits) or the frame index (25 bits)
                                        const int tag_j = __float_as_int(f[u].w);
                                        const unsigned bits = (unsigned)(tag_j & 3) | ((tag_j >> 2) == (tag_i >> 2) ? 4u : 0u);
                                        value |= bits << (staged ? 13 : 25);
                                    }
                                    mine[count & 3] = ((unsigned)level << 28) | value;
                                    if ((count & 3) == 3) column[(count >> 2) * 32] = make_uint4(mine[0], mine[1], mine[2], mine[3]);
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
                if ((count & 3) != 0) {
                    for (int k = count & 3; k < 4; k++) mine[k] = 0u;
                    column[(count >> 2) * 32] = make_uint4(mine[0], mine[1], mine[2], mine[3]);
                }
                a.ncount[s_i] = count;
            } else if (s_i < he) {
                a.ncount[s_i] = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the rebuild as one cooperative kernel
// ------------------------------------------------------------------------------------------------

struct Rebuild2Args {
    int n, ncells, scan_blocks, ext_scan_blocks, nunits, epoch;
    GridView g;
    ExtGrid e;
    const double* position;
    int *cell_of, *slot_of, *cell_count, *cell_start, *scan_scratch, *grouped, *ext_start, *ext_pad;
    int *row_pad, *row_units, *unit_start, *row_scratch;  // row-aligned units
    int nrows, row_scan_blocks;
    Scatter2Args scatter;
    Frame2Args frame;
    Table2Args table;
    Build2Args build;
    int* flags;
    // Sorted-resident molecular dynamics (the state arrays are kept in cell order between rebuilds): `position` is then the
    // state in the OLD cell order; velocities, masses and the map back to the caller's atom order follow the atoms into the
    // new order.  Sharded: the positions and velocities of the other ranks' atoms arrive by peer stores first.
    struct Sorted {
        int active;
        double *x, *v, *m, *tmp;     // state; tmp holds 4 doubles per atom
        int *origin, *tmp_origin;    // sorted slot -> index in the caller's arrays
        int nranks;
        const int* gathered;         // sharded: gathered[rank] == epoch once that rank's state has arrived here
    } sorted;
    // sharded runs (need_mask set): this rank, and where to put the (first double, doubles, peers) of every frame run of its
    // atoms that another rank stages
    int halo_rank;
    int4* halo_items;
};

// velocities, masses and origins of the atoms into scratch, in the new order ...
__device__ __forceinline__ void sorted_collect_phase(int vb, int n, const int* __restrict__ order, const Rebuild2Args::Sorted& a) {
    const int s = vb * REBUILD_THREADS + threadIdx.x;
    if (s >= n) return;
    const int i = order[s];
    a.tmp[4 * (size_t)s] = a.v[3 * (size_t)i];
    a.tmp[4 * (size_t)s + 1] = a.v[3 * (size_t)i + 1];
    a.tmp[4 * (size_t)s + 2] = a.v[3 * (size_t)i + 2];
    a.tmp[4 * (size_t)s + 3] = a.m[i];
    a.tmp_origin[s] = a.origin[i];
}

// ... and back into the state arrays once every block has read the old order
__device__ __forceinline__ void sorted_commit_phase(int vb, int n, const double* __restrict__ xref, const Rebuild2Args::Sorted& a) {
    const int s = vb * REBUILD_THREADS + threadIdx.x;
    if (s >= n) return;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        a.v[3 * (size_t)s + c] = a.tmp[4 * (size_t)s + c];
        a.x[3 * (size_t)s + c] = xref[3 * (size_t)s + c];
    }
    a.m[s] = a.tmp[4 * (size_t)s + 3];
    a.origin[s] = a.tmp_origin[s];
}

template <int FLAGS>
__global__ void __launch_bounds__(REBUILD_THREADS, 3) rebuild2_kernel(Rebuild2Args r) {
    if (r.flags[FLAG_REBUILD] != r.epoch) return;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    if (r.sorted.active && r.sorted.nranks > 1) {
        // every rank takes the rebuild decision from the same number, so every rank is here: wait for their atoms
        if (blockIdx.x == 0 && threadIdx.x < r.sorted.nranks) {
            const volatile int* flag = r.sorted.gathered + threadIdx.x;
            const long long start = clock64();
            while (*flag != r.epoch) {
                if (clock64() - start > 40000000000ll) {
                    r.flags[FLAG_NONFINITE] = 2;  // reported as a communication error by the host
                    break;
                }
            }
            __threadfence_system();
        }
        grid.sync();
        if (r.flags[FLAG_NONFINITE] != 0) return;
    }
    __shared__ int scan_shared[33];
    __shared__ float offset32[27][3];
    __shared__ unsigned pending[REBUILD_THREADS][4];
    if (threadIdx.x < 27) {
        const int t = threadIdx.x;
        offset32[t][0] = (float)((double)((t % 3) - 1) * r.g.edge[0]);
        offset32[t][1] = (float)((double)(((t / 3) % 3) - 1) * r.g.edge[1]);
        offset32[t][2] = (float)((double)((t / 9) - 1) * r.g.edge[2]);
    }
    const int atom_blocks = (r.n + REBUILD_THREADS - 1) / REBUILD_THREADS;

    cell_zero_phase(r.ncells + 1, r.cell_count, r.build.cell_needed, r.flags);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        r.flags[FLAG_FRAME_OVERFLOW] = 0;
        r.flags[FLAG_HALO_ITEMS] = 0;
    }
    if (r.table.need_mask != nullptr) {
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < r.ncells; k += gridDim.x * blockDim.x) r.table.need_mask[k] = 0;
    }
    grid.sync();
    for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) {
        cell_assign_phase(vb, r.n, r.g, r.position, r.cell_of, r.slot_of, r.cell_count, r.flags);
    }
    grid.sync();
    if (r.flags[FLAG_NONFINITE] != 0) return;
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
    // the extended cells only need cell_start: their padded counts are formed alongside
    ext_count_phase(r.e, r.cell_start, r.ext_pad);
    const bool aligned = (FLAGS & 4) != 0;
    if (aligned) {
        // units per row of cells
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < r.nrows; row += gridDim.x * blockDim.x) {
            const int atoms = r.cell_start[(row + 1) * r.g.nc[0]] - r.cell_start[row * r.g.nc[0]];
            r.row_pad[row] = (atoms + r.table.shape.atoms - 1) / r.table.shape.atoms;
        }
    }
    grid.sync();
    for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) scatter2_phase(vb, r.scatter, r.flags);
    grid.sync();
    if (r.flags[FLAG_NONFINITE] != 0) return;
    const int next = r.e.count();
    for (int vb = blockIdx.x; vb < r.ext_scan_blocks; vb += gridDim.x) {
        scan_blocks_phase(vb, next, r.ext_pad, r.ext_start, r.scan_scratch, scan_shared);
    }
    if (aligned) {
        for (int vb = blockIdx.x; vb < r.row_scan_blocks; vb += gridDim.x) {
            scan_blocks_phase(vb, r.nrows, r.row_pad, r.row_units, r.row_scratch, scan_shared);
        }
    }
    grid.sync();
    if (blockIdx.x == 0) scan_sums_phase(r.ext_scan_blocks, r.scan_scratch, scan_shared);
    if (aligned && blockIdx.x == (gridDim.x > 1 ? 1 : 0)) scan_sums_phase(r.row_scan_blocks, r.row_scratch, scan_shared);
    grid.sync();
    for (int vb = blockIdx.x; vb < r.ext_scan_blocks; vb += gridDim.x) {
        scan_add_total_phase(vb, next, r.ext_pad, r.ext_start, r.scan_scratch);
    }
    if (aligned) {
        for (int vb = blockIdx.x; vb < r.row_scan_blocks; vb += gridDim.x) {
            scan_add_total_phase(vb, r.nrows, r.row_pad, r.row_units, r.row_scratch);
        }
        grid.sync();
        // first atom of every unit
        for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < r.nrows; row += gridDim.x * blockDim.x) {
            const int first = r.cell_start[row * r.g.nc[0]];
            const int u0 = r.row_units[row], u1 = r.row_units[row + 1];
            for (int u = u0; u < u1; u++) r.unit_start[u] = first + (u - u0) * r.table.shape.atoms;
            if (row == r.nrows - 1) {
                r.unit_start[u1] = r.n;
                r.flags[FLAG_UNITS] = u1;
            }
        }
    }
    grid.sync();
    for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) frame2_phase(vb, r.frame, r.flags);
    for (int vb = blockIdx.x; vb * REBUILD_WARPS < r.nunits; vb += gridDim.x) unit_table_phase(vb, r.table);
    if (r.sorted.active) {
        for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) sorted_collect_phase(vb, r.n, r.scatter.order, r.sorted);
    }
    grid.sync();
    if (r.sorted.active) {
        for (int vb = blockIdx.x; vb < atom_blocks; vb += gridDim.x) sorted_commit_phase(vb, r.n, r.scatter.xref, r.sorted);
    }
    if (r.table.need_mask != nullptr && r.halo_items != nullptr && r.build.s_hi > r.build.s_lo) {
        // the frame runs of this rank's atoms (every image of every cell they sit in) that other ranks stage: what the halo
        // kernel copies after every kick-drift
        const int c_lo = r.scatter.sorted_cell[r.build.s_lo];
        const int ncells = r.scatter.sorted_cell[r.build.s_hi - 1] - c_lo + 1;
        for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < 8 * ncells; item += gridDim.x * blockDim.x) {
            const int cell = c_lo + (item >> 3), m = item & 7;
            const unsigned targets = (unsigned)r.table.need_mask[cell] & ~(1u << r.halo_rank);
            if (targets == 0u) continue;
            const int first = r.cell_start[cell], end = r.cell_start[cell + 1];
            const int lo = max(first, r.build.s_lo), hi = min(end, r.build.s_hi);
            if (lo >= hi) continue;
            const int x = cell % r.e.nx, y = (cell / r.e.nx) % r.e.ny, z = cell / (r.e.nx * r.e.ny);
            const int gx = x == 0 ? 1 : (x == r.e.nx - 1 ? -1 : 0);
            const int gy = y == 0 ? 1 : (y == r.e.ny - 1 ? -1 : 0);
            const int gz = z == 0 ? 1 : (z == r.e.nz - 1 ? -1 : 0);
            if (((m & 1) && gx == 0) || ((m & 2) && gy == 0) || ((m & 4) && gz == 0)) continue;
            const int sx = (m & 1) ? gx : 0, sy = (m & 2) ? gy : 0, sz = (m & 4) ? gz : 0;
            const int X = sx > 0 ? r.e.ex - 1 : (sx < 0 ? 0 : x + 1);
            const int Y = sy > 0 ? r.e.ey - 1 : (sy < 0 ? 0 : y + 1);
            const int Z = sz > 0 ? r.e.ez - 1 : (sz < 0 ? 0 : z + 1);
            const int begin = 3 * (2 + r.ext_start[r.e.index(X, Y, Z)] + (lo - first));
            r.halo_items[atomicAdd(r.flags + FLAG_HALO_ITEMS, 1)] = make_int4(begin, 3 * (hi - lo), (int)targets, 0);
        }
    }
    for (int vb = blockIdx.x; vb * REBUILD_WARPS * r.build.cells_per_warp < r.ncells * BUILD2_CHUNKS; vb += gridDim.x) {
        list_build2_phase<FLAGS>(vb, r.build, offset32, pending);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        r.flags[FLAG_COUNT] += 1;
        // the frames hold the rebuild positions: nothing has moved yet
        r.flags[FLAG_DISP] = 0;
        r.flags[FLAG_DISP + 1] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// final order of the columns: by level, bank-aware inside a level, two streams per atom
// ------------------------------------------------------------------------------------------------
//
// A staged copy holds (x, y, z) triples of doubles, so coordinate c of slot s sits in the 8-byte bank pair (3 s + c) mod
// 16: two lanes of a half warp collide when their slots are different and congruent modulo 16 (3 is invertible mod 16;
// list entries are 3 s, whose low four bits are the bucket).  Lanes 2a + h of a half warp (eight atoms, two
// halves) read entry 8 w + 4 h + q of their column at sub-step t = 4 w + q: the entry emitted for that position is
// taken, while the buckets of the current level last, from residue (2 (a mod 8) + h + t) mod 16, so the sixteen lanes
// of a half warp touch sixteen different bank pairs.

constexpr int REORDER2_THREADS = 32;

// COUNTER: unsigned char while a column holds at most 255 entries (more resident warps: the kernel is a chain of dependent
// shared-memory accesses), unsigned short otherwise.  Columns of units that could not be staged keep the order of the
// build (their 32-bit entries only lose the level bits; every level counts as walked).
template <typename COUNTER>
__global__ void __launch_bounds__(REORDER2_THREADS)
    list_reorder2_kernel(UnitShape shape, int s_lo, int n, int capacity, const int4* __restrict__ header, const int* __restrict__ ncount,
                         unsigned* __restrict__ nlist, unsigned short* __restrict__ cum_levels, int epoch,
                         const int* __restrict__ flags) {
    if (flags[FLAG_REBUILD] != epoch || flags[FLAG_NONFINITE] != 0) return;
    extern __shared__ unsigned short reorder2_smem[];
    unsigned short* sorted = reorder2_smem;  // entry p of thread t at [p * 32 + t], grouped by (level, residue)
    // buckets: the sixteen bank residues of level 0 (four fifths of the entries: the only ones walked at every evaluation
    // get the bank-aware order), then one bucket per further level
    constexpr int BUCKETS = 16 + LJ2_LEVELS - 1;
    __shared__ COUNTER cursor[BUCKETS][REORDER2_THREADS], last[BUCKETS][REORDER2_THREADS];
    auto bucket_of = [](unsigned entry) { return (entry >> 28) == 0u ? (int)(entry & 15u) : 15 + (int)(entry >> 28); };
    const int t = threadIdx.x;
    const int word_stride = shape.lanes_per_atom / 2;
    // columns [s_lo, n): the atoms of this rank (s_lo is a multiple of the unit size)
    for (int slab = s_lo / REORDER2_THREADS + blockIdx.x; slab * REORDER2_THREADS < n; slab += gridDim.x) {
        const int s_i = slab * REORDER2_THREADS + t;
        if (s_i >= n) continue;
        const bool staged = header[s_i / shape.atoms].z >= 0;
        const int count = ncount[s_i];
        uint4* words = reinterpret_cast<uint4*>(nlist) + (size_t)(s_i >> 5) * (capacity >> 2) * 32 + (s_i & 31);
        const int nwords = (count + 3) >> 2;
        if (!staged) {
            for (int w = 0; w < nwords; w++) {
                uint4 word = words[w * 32];
                word.x &= RAW_VALUE_MASK;
                word.y &= RAW_VALUE_MASK;
                word.z &= RAW_VALUE_MASK;
                word.w &= RAW_VALUE_MASK;
                words[w * 32] = word;
            }
            for (int level = 0; level < LJ2_LEVELS; level++) cum_levels[(size_t)s_i * LJ2_LEVELS + level] = (unsigned short)count;
            continue;
        }
        for (int b = 0; b < BUCKETS; b++) last[b][t] = 0;
        // (four words in flight: the loop is bound by the latency of these loads otherwise)
        for (int w0 = 0; w0 < nwords; w0 += 4) {
            uint4 ahead[4];
#pragma unroll
            for (int u = 0; u < 4; u++) ahead[u] = words[min(w0 + u, nwords - 1) * 32];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int w = w0 + u;
                const unsigned e[4] = {ahead[u].x, ahead[u].y, ahead[u].z, ahead[u].w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (4 * w + q < count) last[bucket_of(e[q])][t]++;
                }
            }
        }
        int running = 0;
        for (int b = 0; b < BUCKETS; b++) {
            cursor[b][t] = (COUNTER)running;
            running += last[b][t];
            last[b][t] = (COUNTER)running;
            if (b >= 15) cum_levels[(size_t)s_i * LJ2_LEVELS + (b - 15)] = (unsigned short)running;
        }
        for (int w0 = 0; w0 < nwords; w0 += 4) {
            uint4 ahead[4];
#pragma unroll
            for (int u = 0; u < 4; u++) ahead[u] = words[min(w0 + u, nwords - 1) * 32];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int w = w0 + u;
                const unsigned e[4] = {ahead[u].x, ahead[u].y, ahead[u].z, ahead[u].w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (4 * w + q < count) {
                        const int position = cursor[bucket_of(e[q])][t]++;
                        sorted[position * REORDER2_THREADS + t] = (unsigned short)(e[q] & 0xffffu);
                    }
                }
            }
        }
        // rewind the residue buckets of level 0
        running = 0;
        unsigned nonempty = 0;  // residue buckets of level 0 that still hold entries
        for (int b = 0; b < 16; b++) {
            cursor[b][t] = (COUNTER)running;
            nonempty |= running != (int)last[b][t] ? 1u << b : 0u;
            running = last[b][t];
        }
        const int level0 = running;
        const int lane_base = shape.lanes_per_atom * (s_i & (16 / shape.lanes_per_atom - 1));
        unsigned packed[4] = {0u, 0u, 0u, 0u};
        for (int p = 0; p < count; p++) {
            int position = p;  // further levels: the order of the counting sort
            if (p < level0) {
                // entry p sits in word p >> 3, half (p >> 2) & 1; with four lanes per atom, lanes {0, 1} walk the even words
                // and {2, 3} the odd ones
                const int word = p >> 3;
                const int h = ((word % word_stride) << 1) | ((p >> 2) & 1), step = ((word / word_stride) << 2) | (p & 3);
                const int want = (lane_base + h + step) & 15;
                // first non-empty bucket at or after `want`, cyclically
                const unsigned rotated = ((nonempty >> want) | (nonempty << (16 - want))) & 0xffffu;
                const int b = (want + __ffs(rotated) - 1) & 15;
                position = cursor[b][t]++;
                if (position + 1 == (int)last[b][t]) nonempty &= ~(1u << b);
            }
            const unsigned slot = sorted[position * REORDER2_THREADS + t];
            packed[(p & 7) >> 1] |= slot << (16 * (p & 1));
            if ((p & 7) == 7) {
                words[(p >> 3) * 32] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                packed[0] = packed[1] = packed[2] = packed[3] = 0u;
            }
        }
        if ((count & 7) != 0) words[(count >> 3) * 32] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// per-evaluation refresh of the frames
// ------------------------------------------------------------------------------------------------

struct Update2Args {
    int n;
    int s_lo, s_hi;  // sorted range refreshed by this rank
    GridView g;
    ExtGrid e;
    const int* __restrict__ order;  // sorted slot -> state index (nullptr: the state arrays are in sorted order)
    const double* __restrict__ pos;
    const double* __restrict__ xref;
    const int* __restrict__ kshift;
    const int* __restrict__ fidx;
    const int* __restrict__ sorted_cell;
    const int* __restrict__ cell_start;
    const int* __restrict__ ext_start;
    double* __restrict__ frame;  // (x, y, z) or (x, y, z, charge) per frame slot
    size_t fstride;
    int width;
    double scale;
    double threshold2;
    int epoch;
    int* __restrict__ flags;
};

// Positions keep the periodic image they had at the rebuild (frame = x - k L with the wrap counts of the rebuild); the
// largest squared displacement since the rebuild is reduced per block and folded with an integer atomicMax of its float
// bits (non-negative floats order like their bit patterns).
__global__ void __launch_bounds__(256) lj2_update_kernel(Update2Args a) {
    __shared__ float block_max[8];
    const int s = a.s_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.flags[FLAG_DISP + ((a.epoch + 1) & 1)] = 0;  // the slot of the next evaluation
        a.flags[FLAG_DEFERRED] = 0;
    }
    float d2f = 0.0f;
    if (s < a.s_hi) {
        const int i = a.order != nullptr ? a.order[s] : s;
        const double px = a.pos[3 * i], py = a.pos[3 * i + 1], pz = a.pos[3 * i + 2];
        const double dx = px - a.xref[3 * s], dy = py - a.xref[3 * s + 1], dz = pz - a.xref[3 * s + 2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (!(d2 <= a.threshold2)) a.flags[FLAG_REBUILD] = a.epoch;  // also catches NaN
        d2f = __double2float_ru(d2);
        if (!(d2f >= 0.0f)) d2f = 3.0e38f;
        int kx, ky, kz;
        unpack_shift(a.kshift[s], kx, ky, kz);
        const double x = px - (double)kx * a.g.length[0];
        const double y = py - (double)ky * a.g.length[1];
        const double z = pz - (double)kz * a.g.length[2];
        const int c = a.sorted_cell[s];
        const int rank = s - a.cell_start[c];
        for_each_image(a.e, a.ext_start, c, rank, [&](int f, int sx, int sy, int sz) {
            a.frame[(size_t)a.width * f] = (x + (double)sx * a.g.length[0]) * a.scale;
            a.frame[(size_t)a.width * f + 1] = (y + (double)sy * a.g.length[1]) * a.scale;
            a.frame[(size_t)a.width * f + 2] = (z + (double)sz * a.g.length[2]) * a.scale;
        });
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2f = fmaxf(d2f, __shfl_xor_sync(0xffffffffu, d2f, o));
    if ((threadIdx.x & 31) == 0) block_max[threadIdx.x >> 5] = d2f;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = block_max[0];
        for (int w = 1; w < 8; w++) m = fmaxf(m, block_max[w]);
        atomicMax(a.flags + FLAG_DISP + (a.epoch & 1), __float_as_int(m));
    }
}

// ------------------------------------------------------------------------------------------------
// force kernel
// ------------------------------------------------------------------------------------------------

struct Lj2Args {
    int n;
    int u_lo, u_hi;  // units of this rank
    int capacity;
    const unsigned* __restrict__ nlist;
    const unsigned short* __restrict__ cum_levels;
    const unsigned short* __restrict__ self_slot;
    const int* __restrict__ fidx;
    const int* __restrict__ order;  // nullptr: state arrays in sorted order
    const int4* __restrict__ header;
    const int4* __restrict__ runs;
    const double* __restrict__ frame;  // (x, y, z) per frame slot
    size_t fstride;
    double epsilon24, epsilon48, epsilon4, shift, inv_sigma;
    int band_lo;           // top 32 bits of (rc / sigma)^2, minus one: below it a pair is inside the cut-off
    float inv_delta, margin;  // levels needed: floor((2 dmax + margin) * inv_delta) + 1
    int epoch;
    int all_levels;  // experiment knob: walk every level whatever the displacement
    int write_forces;
    double* __restrict__ force;
    double* __restrict__ partials;
    int2* __restrict__ deferred;  // (state index of i, frame index of j)
    int deferred_capacity;
    int* __restrict__ flags;
};

__device__ __forceinline__ unsigned smem_address(const void* pointer) { return (unsigned)__cvta_generic_to_shared(pointer); }

__device__ __forceinline__ void mbar_init(unsigned long long* barrier, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_address(barrier)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_expect_bytes(unsigned long long* barrier, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_address(barrier)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* barrier) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_address(barrier)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* barrier, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred done;\n"
        "LJ2_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
        "@done bra LJ2_WAIT_DONE;\n"
        "bra LJ2_WAIT_LOOP;\n"
        "LJ2_WAIT_DONE:\n"
        "}\n" ::"r"(smem_address(barrier)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void async_copy16(void* shared_dst, const void* global_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_address(shared_dst)), "l"(global_src) : "memory");
}
__device__ __forceinline__ void bulk_load(double* shared_dst, const double* global_src, unsigned bytes, unsigned long long* barrier) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_address(shared_dst)),
                 "l"(global_src), "r"(bytes), "r"(smem_address(barrier))
                 : "memory");
}

// One listed neighbour: 17 FP64 instructions, no branch.  Lengths are in units of sigma, so 1 / r^2 is s2 directly; the
// cut-off test compares the top 32 bits of r^2 on the integer pipe (r^2 >= 0: the bit patterns order like the values),
// a pair outside gets a zero reciprocal seed and contributes exactly zero.  `near` collects the pairs whose r^2 falls
// in the three-value band around the cut-off: they are evaluated by the fix-up kernel.
template <int MODE>
__device__ __forceinline__ void lj2_pair(const Lj2Args& a, double xj, double yj, double zj, double xi, double yi, double zi,
                                         double& fx, double& fy, double& fz, double (&acc)[LJ2_NV], unsigned& near, unsigned bit) {
    const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const int top = __double2hiint(r2);
    const bool inside = top < a.band_lo;
    near |= (unsigned)(top - a.band_lo) < 3u ? bit : 0u;
    double seed;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(r2));
    const double y = __hiloint2double(inside ? __double2hiint(seed) : 0, 0);
    // one cubically convergent step, y (1 + e + e^2) with e = 1 - r2 y
    const double e = fma(-r2, y, 1.0);
    const double s2 = fma(y, fma(e, e, e), y);
    const double s4 = s2 * s2;
    const double s6 = s4 * s2;
    // sigma^2 force(r) / r = -24 eps (s6 - 2 s6^2) s2 (functions.rs:85-88); rescaled by 1 / sigma at the end
    const double fr = (s4 * s4) * fma(a.epsilon48, s6, -a.epsilon24);
    fx = fma(fr, dx, fx);
    fy = fma(fr, dy, fy);
    fz = fma(fr, dz, fz);
    if (MODE == 1) {
        const double energy = a.epsilon4 * fma(s6, s6, -s6) - a.shift;
        acc[0] += inside ? energy : 0.0;
        acc[14] += inside ? 1.0 : 0.0;
        acc[2] += fr * dx * dx;
        acc[3] += fr * dx * dy;
        acc[4] += fr * dx * dz;
        acc[5] += fr * dy * dy;
        acc[6] += fr * dy * dz;
        acc[7] += fr * dz * dz;
    }
}

// the four neighbours of a half word: positions out of the staged copy, then the four pair evaluations
struct Staged4 {
    double x[4], y[4], z[4];
    __device__ __forceinline__ void load(const double* __restrict__ buffer, const uint2& word) {
        const unsigned j[4] = {word.x & 0xffffu, word.x >> 16, word.y & 0xffffu, word.y >> 16};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            x[q] = buffer[j[q]];
            y[q] = buffer[j[q] + 1];
            z[q] = buffer[j[q] + 2];
        }
    }
    template <int MODE>
    __device__ __forceinline__ void evaluate(const Lj2Args& a, double xi, double yi, double zi, double& fx, double& fy, double& fz,
                                             double (&acc)[LJ2_NV], unsigned& near) const {
#pragma unroll
        for (int q = 0; q < 4; q++) lj2_pair<MODE>(a, x[q], y[q], z[q], xi, yi, zi, fx, fy, fz, acc, near, 1u << q);
    }
};

// rare: remember the pairs of this half word that sit on the cut-off
__device__ __noinline__ void lj2_defer(const int4* __restrict__ header, const int4* __restrict__ runs, int2* __restrict__ deferred,
                                       int deferred_capacity, int* __restrict__ flags, int unit, int state_i, unsigned near,
                                       unsigned lo, unsigned hi, bool staged) {
    const unsigned entries[2] = {lo, hi};
    for (int k = 0; k < 4; k++) {
        if (!(near & (1u << k))) continue;
        int frame_index;
        if (staged) {
            const int slot = (int)((entries[k >> 1] >> (16 * (k & 1))) & 0xffffu) / 3;
            frame_index = -1;
            const int nruns = header[unit].z;
            for (int r = 0; r < nruns; r++) {
                const int4 run = runs[(size_t)unit * LJ2_MAX_RUNS + r];
                if (slot >= run.z && slot < run.z + run.y) frame_index = run.x + (slot - run.z);
            }
            if (frame_index < 0) continue;  // the dummy
        } else {
            if (k >= 2) continue;
            frame_index = (int)entries[k];
        }
        const int position = atomicAdd(flags + FLAG_DEFERRED, 1);
        if (position < deferred_capacity) deferred[position] = make_int2(state_i, frame_index);
    }
}

// This lane's view of the column of sorted atom s: 8-byte halves of the 16-byte words word_first, word_first + 2, ...
// (uint2 stride 128 between them); word w of atom s sits at ((s >> 5) * slab_words + w * 32 + (s & 31)) in uint4 units.
__device__ __forceinline__ const uint2* lj2_column(const unsigned* __restrict__ nlist, size_t slab_words, int s, int word_first, int half) {  // uint2 stride between its words: 64 * LJ2_WORD_STRIDE
    const uint4* base = reinterpret_cast<const uint4*>(nlist) + (size_t)(s >> 5) * slab_words + (size_t)word_first * 32 + (s & 31);
    return reinterpret_cast<const uint2*>(base) + half;
}

// Bulk copies of one unit's neighbourhood into a staging buffer, issued by one warp: ONE copy per run (the frames hold
// (x, y, z) triples; with one plane per coordinate the three times more numerous, three times smaller copies kept the
// SM's copy engine busy for longer than the pairs took: profiles/r2d_*), completion counted in bytes on the mbarrier.
__device__ __forceinline__ void lj2_issue_copies(const Lj2Args& a, const int4* __restrict__ table, double* buffer, unsigned long long* barrier, int lane) {
    // table[0]: header of the unit, table[1 ...]: its runs (global memory, or the copy a warp left in shared memory)
    const int4 header = table[0];
    if (header.z < 0) {
        // not staged: the consumers gather from global memory, the barrier only has to complete its phase
        if (lane == 0) mbar_arrive(barrier);
        return;
    }
    if (lane == 0) mbar_expect_bytes(barrier, 24u * (unsigned)(header.w - 2));
    __syncwarp();
    for (int r = lane; r < header.z; r += 32) {
        const int4 run = table[1 + r];
        bulk_load(buffer + 3 * run.z, a.frame + 3 * (size_t)run.x, (unsigned)run.y * 24u, barrier);
    }
}

template <int MODE>
__global__ void __launch_bounds__(LJ2_THREADS, 1) lj2_force_kernel(Lj2Args a) {
    extern __shared__ __align__(16) unsigned char lj2_smem[];
    double* stage = reinterpret_cast<double*>(lj2_smem);  // two buffers of three planes
    __shared__ __align__(8) unsigned long long full[LJ2_BUFFERS];
    __shared__ int done[LJ2_BUFFERS];  // warps that have finished with each buffer
    __shared__ int4 refill_table[LJ2_BUFFERS][1 + LJ2_MAX_RUNS];  // header and runs of the unit that will refill each buffer

    if (a.flags[FLAG_NONFINITE] != 0) return;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    if (tid == 0) {
        for (int b = 0; b < LJ2_BUFFERS; b++) {
            mbar_init(&full[b], 1);
            done[b] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 6 * LJ2_BUFFERS) {
        // slots 0 and 1 of every buffer: the dummy, far outside any cut-off
        const int buffer = tid / 6, slot = (tid % 6) / 3, coordinate = tid % 3;
        stage[(size_t)buffer * 3 * LJ2_SLOTS + 3 * slot + coordinate] = 1.0e9 * (double)(coordinate + 1);
    }
    __syncthreads();

    double acc[LJ2_NV];
#pragma unroll
    for (int k = 0; k < LJ2_NV; k++) acc[k] = 0.0;

    const int first_unit = a.u_lo + blockIdx.x, stride = gridDim.x;
    // the first units: copies issued by warp 0; afterwards the LAST warp to finish unit k issues the copies of unit
    // k + LJ2_BUFFERS into the buffer it has just released, so no warp ever waits for a buffer to drain
    if (tid < 32) {
        for (int b = 0; b < LJ2_BUFFERS; b++) {
            const int unit = first_unit + b * stride;
            if (unit >= a.u_hi) break;
            if (lane == 0) refill_table[b][0] = a.header[unit];
            for (int r = lane; r < LJ2_MAX_RUNS; r += 32) refill_table[b][1 + r] = a.runs[(size_t)unit * LJ2_MAX_RUNS + r];
            __syncwarp();
            lj2_issue_copies(a, refill_table[b], stage + (size_t)b * 3 * LJ2_SLOTS, &full[b], lane);
        }
    }
    {
        // ---- consumers ------------------------------------------------------------------------------------------
        // the LJ2_LPA lanes of an atom: lanes {0, 1} take the two halves of its words (of the even words with four lanes per atom,
        // lanes {2, 3} then take the odd ones)
        const int local = tid / LJ2_LPA, sub = tid % LJ2_LPA;
        const int word_first = sub >> 1, half = sub & 1;
        // levels to walk: pairs listed at r_b >= rc + 2 dmax cannot have come inside the cut-off
        const float dmax = sqrtf(__int_as_float(a.flags[FLAG_DISP + (a.epoch & 1)])) * 1.000001f;
        int nlevels = (int)((2.0f * dmax + a.margin) * a.inv_delta) + 1;
        nlevels = max(1, min(LJ2_LEVELS, nlevels));
        if (a.all_levels == 1) nlevels = LJ2_LEVELS;
        const size_t slab_words = (size_t)(a.capacity >> 2) * 32;  // uint4 words per slab of 32 columns

        // head of the first unit
        int s = first_unit * U_ATOMS + local;
        int count = 0;
        unsigned self = 0;
        // the first four words of the lane's walk, loaded one unit ahead (whatever the count: the slab is allocated; what
        // lies behind the walked prefix is discarded when the unit starts)
        uint2 h0 = make_uint2(0, 0), h1 = make_uint2(0, 0), h2 = make_uint2(0, 0), h3 = make_uint2(0, 0);
        if (first_unit < a.u_hi && s < a.n) {
            count = a.cum_levels[(size_t)s * LJ2_LEVELS + nlevels - 1];
            self = a.self_slot[s];
            const uint2* words = lj2_column(a.nlist, slab_words, s, word_first, half);
            h0 = words[0];
            h1 = words[64 * LJ2_WORD_STRIDE];
            h2 = words[2 * 64 * LJ2_WORD_STRIDE];
            h3 = words[3 * 64 * LJ2_WORD_STRIDE];
        }
        int b = 0, phase = 0;  // buffer k mod 3 and the parity of its use number k / 3
        for (int unit = first_unit; unit < a.u_hi; unit += stride) {
            const int s_now = s, count_now = a.all_levels == 2 ? 0 : count;  // experiment knob 2: staging only
            const unsigned self_now = self;
            uint2 wcur = h0, wnext = h1, w2 = h2, w3 = h3;
            // head of the next unit: in flight while this one is evaluated
            const int next_unit = unit + stride;
            s = next_unit * U_ATOMS + local;
            count = 0;
            const uint2* next_words = nullptr;
            if (next_unit < a.u_hi && s < a.n) {
                count = a.cum_levels[(size_t)s * LJ2_LEVELS + nlevels - 1];
                self = a.self_slot[s];
                next_words = lj2_column(a.nlist, slab_words, s, word_first, half);
                h0 = next_words[0];
                h1 = next_words[64 * LJ2_WORD_STRIDE];
                h2 = next_words[2 * 64 * LJ2_WORD_STRIDE];
                h3 = next_words[3 * 64 * LJ2_WORD_STRIDE];
            }
            const int4 header = a.header[unit];
            const bool staged = header.z >= 0;
            const bool present = s_now < a.n;
            const int state_i = present ? (a.order != nullptr ? a.order[s_now] : s_now) : 0;
            const uint2* words = lj2_column(a.nlist, slab_words, s_now, word_first, half);
            double fx = 0.0, fy = 0.0, fz = 0.0;
            mbar_wait(&full[b], phase);
            // the table of the unit that will refill this buffer: fetched by warp 0 with asynchronous copies once this unit's
            // bulk copies have landed (so the table they were issued from is free); it is complete before warp 0 reports this
            // unit done, and used by whichever warp finishes the unit last
            if (tid < 32 && unit + LJ2_BUFFERS * stride < a.u_hi) {
                const int refill = unit + LJ2_BUFFERS * stride;
                if (lane == 0) async_copy16(&refill_table[b][0], a.header + refill);
                for (int r = lane; r < LJ2_MAX_RUNS; r += 32) async_copy16(&refill_table[b][1 + r], a.runs + (size_t)refill * LJ2_MAX_RUNS + r);
            }
            if (staged) {
                const double* buffer = stage + (size_t)b * 3 * LJ2_SLOTS;
                const double xi = buffer[self_now], yi = buffer[self_now + 1], zi = buffer[self_now + 2];
                // words of this lane: word_first, word_first + LJ2_WORD_STRIDE, ... of the (count + 7) / 8 words of the prefix
                const int nwords = (((count_now + 7) >> 3) - word_first + LJ2_WORD_STRIDE - 1) / LJ2_WORD_STRIDE;
                // Software pipeline, unrolled four words deep: the list words of a column come from DRAM (the list is the
                // only stream of the kernel), so four of them are in flight per thread and the lines further ahead are
                // pulled into L2; the twelve position loads of the next half word are issued before the current four
                // pairs are evaluated.
                // words behind the walked prefix may hold anything: never turned into shared-memory addresses
                if (0 >= nwords) wcur = make_uint2(0, 0);
                if (1 >= nwords) wnext = make_uint2(0, 0);
                if (2 >= nwords) w2 = make_uint2(0, 0);
                if (3 >= nwords) w3 = make_uint2(0, 0);
#if LJ2_ILP8
                // eight independent pair chains per thread: the positions of two half words are loaded, then both are evaluated
                // in one basic block (the FP64 pipe is latency-bound with four chains per thread and four warps per scheduler)
                for (int w = 0; w < nwords; w += 4) {
                    Staged4 pa, pb;
                    pa.load(buffer, wcur);
                    pb.load(buffer, wnext);
                    unsigned near = 0, near_b = 0;
                    pa.evaluate<MODE>(a, xi, yi, zi, fx, fy, fz, acc, near);
                    pb.evaluate<MODE>(a, xi, yi, zi, fx, fy, fz, acc, near_b);
                    if ((near | near_b) != 0) {
                        if (near != 0) lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, state_i, near, wcur.x, wcur.y, true);
                        if (near_b != 0) lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, state_i, near_b, wnext.x, wnext.y, true);
                    }
                    wcur = make_uint2(0, 0);
                    wnext = make_uint2(0, 0);
                    if (w + 4 < nwords) wcur = words[(size_t)(w + 4) * 64 * LJ2_WORD_STRIDE];
                    if (w + 5 < nwords) wnext = words[(size_t)(w + 5) * 64 * LJ2_WORD_STRIDE];
                    if (next_words != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(next_words + (size_t)(w + 4) * 64 * LJ2_WORD_STRIDE));
                    if (w + 2 >= nwords) break;
                    pa.load(buffer, w2);
                    pb.load(buffer, w3);
                    near = 0;
                    near_b = 0;
                    pa.evaluate<MODE>(a, xi, yi, zi, fx, fy, fz, acc, near);
                    pb.evaluate<MODE>(a, xi, yi, zi, fx, fy, fz, acc, near_b);
                    if ((near | near_b) != 0) {
                        if (near != 0) lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, state_i, near, w2.x, w2.y, true);
                        if (near_b != 0) lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, state_i, near_b, w3.x, w3.y, true);
                    }
                    w2 = make_uint2(0, 0);
                    w3 = make_uint2(0, 0);
                    if (w + 6 < nwords) w2 = words[(size_t)(w + 6) * 64 * LJ2_WORD_STRIDE];
                    if (w + 7 < nwords) w3 = words[(size_t)(w + 7) * 64 * LJ2_WORD_STRIDE];
                    if (next_words != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(next_words + (size_t)(w + 6) * 64 * LJ2_WORD_STRIDE));
                }
#else
                Staged4 pa, pb;
                pa.load(buffer, wcur);
#define LJ2_STEP(CUR, NEXT_POS, WORD_NOW, WORD_NEXT, AHEAD)                                                              \
    {                                                                                                                      \
        NEXT_POS.load(buffer, WORD_NEXT);                                                                                  \
        unsigned near = 0;                                                                                                 \
        CUR.evaluate<MODE>(a, xi, yi, zi, fx, fy, fz, acc, near);                                                          \
        if (near != 0) {                                                                                                   \
            lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, state_i, near, WORD_NOW.x, WORD_NOW.y, true); \
        }                                                                                                                  \
        WORD_NOW = make_uint2(0, 0);                                                                                       \
        if (w + (AHEAD) < nwords) WORD_NOW = words[(size_t)(w + (AHEAD)) * 64 * LJ2_WORD_STRIDE];                          \
        if (next_words != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(next_words + (size_t)(w + (AHEAD)) * 64 * LJ2_WORD_STRIDE)); \
    }
                for (int w = 0; w < nwords; w += 4) {
                    LJ2_STEP(pa, pb, wcur, wnext, 4)
                    if (w + 1 >= nwords) break;
                    LJ2_STEP(pb, pa, wnext, w2, 5)
                    if (w + 2 >= nwords) break;
                    LJ2_STEP(pa, pb, w2, w3, 6)
                    if (w + 3 >= nwords) break;
                    LJ2_STEP(pb, pa, w3, wcur, 7)
                }
#undef LJ2_STEP
#endif
            } else if (present) {
                // unit whose neighbourhood does not fit in shared memory: 32-bit frame indices, global gathers
                const int f_self = a.fidx[s_now];
                const double xi = a.frame[3 * (size_t)f_self], yi = a.frame[3 * (size_t)f_self + 1], zi = a.frame[3 * (size_t)f_self + 2];
                const int nwords = (((count_now + 3) >> 2) - word_first + LJ2_WORD_STRIDE - 1) / LJ2_WORD_STRIDE;
                for (int w = 0; w < nwords; w++) {
                    const uint2 word = words[(size_t)w * 64 * LJ2_WORD_STRIDE];
                    unsigned near = 0;
                    const unsigned j0 = word.x, j1 = word.y;
                    lj2_pair<MODE>(a, a.frame[3 * (size_t)j0], a.frame[3 * (size_t)j0 + 1], a.frame[3 * (size_t)j0 + 2], xi, yi, zi, fx, fy, fz, acc, near, 1u);
                    lj2_pair<MODE>(a, a.frame[3 * (size_t)j1], a.frame[3 * (size_t)j1 + 1], a.frame[3 * (size_t)j1 + 2], xi, yi, zi, fx, fy, fz, acc, near, 2u);
                    if (near != 0) lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, state_i, near, word.x, word.y, false);
                }
            }
            // the lanes of an atom
#pragma unroll
            for (int o = 1; o < LJ2_LPA; o <<= 1) {
                fx += __shfl_xor_sync(0xffffffffu, fx, o);
                fy += __shfl_xor_sync(0xffffffffu, fy, o);
                fz += __shfl_xor_sync(0xffffffffu, fz, o);
            }
            if (present && sub == 0 && a.write_forces) {
                a.force[3 * state_i] = fx * a.inv_sigma;
                a.force[3 * state_i + 1] = fy * a.inv_sigma;
                a.force[3 * state_i + 2] = fz * a.inv_sigma;
            }
            // hand the buffer back; the last warp of the block refills it
            if (tid < 32) asm volatile("cp.async.wait_all;" ::: "memory");
            __threadfence_block();
            __syncwarp();
            int last = 0;
            if (lane == 0) last = atomicAdd(&done[b], 1) == LJ2_THREADS / 32 - 1 ? 1 : 0;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                if (lane == 0) done[b] = 0;
                const int refill = unit + LJ2_BUFFERS * stride;
                if (refill < a.u_hi) lj2_issue_copies(a, refill_table[b], stage + (size_t)b * 3 * LJ2_SLOTS, &full[b], lane);
            }
            if (++b == LJ2_BUFFERS) {
                b = 0;
                phase ^= 1;
            }
        }
    }

    if (MODE == 1) {
        // every pair was visited from both sides: half of the energy and virial from each
#pragma unroll
        for (int k = 0; k < LJ2_NV; k++) acc[k] *= 0.5;
        __syncthreads();
        block_sum<LJ2_NV>(acc, stage);  // the staged copies are dead
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < LJ2_NV; k++) a.partials[(size_t)blockIdx.x * LJ2_NV + k] = acc[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pairs on the cut-off: the reference's own arithmetic on the original coordinates
// ------------------------------------------------------------------------------------------------

struct Fix2Args {
    const int2* __restrict__ deferred;
    int capacity;
    const int* __restrict__ frame_atom;
    const int* __restrict__ order;  // nullptr: state arrays in sorted order
    const double* __restrict__ pos;
    // sharded sorted-resident runs: the state arrays hold current positions of the rank's own atoms only; the separation is
    // then taken from the frames (positions of both atoms times 1 / sigma, same periodic image up to whole boxes)
    const double* __restrict__ frame;
    const int* __restrict__ fidx;
    double sigma_for_frames;
    double length[3];
    double sigma, epsilon, cutoff, shift;
    int full;
    int write_forces;
    double* __restrict__ force;
    double* __restrict__ results;
    int* __restrict__ flags;
};

__global__ void __launch_bounds__(128) lj2_fixup_kernel(Fix2Args a) {
    const int count = min(a.flags[FLAG_DEFERRED], a.capacity);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const int2 entry = a.deferred[k];
        const int i = entry.x;
        const int s_j = a.frame_atom[entry.y];
        if (s_j < 0) continue;
        const int j = a.order != nullptr ? a.order[s_j] : s_j;
        // nearest_image(i, j) (configuration.rs:399-403, cells.rs:287-291), no contraction: Rust never fuses
        double d[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double raw;
            if (a.frame != nullptr) {
                raw = (a.frame[3 * (size_t)a.fidx[i] + c] - a.frame[3 * (size_t)entry.y + c]) * a.sigma_for_frames;
            } else {
                raw = __dadd_rn(a.pos[3 * i + c], -a.pos[3 * j + c]);
            }
            d[c] = __dadd_rn(raw, -__dmul_rn(round(__ddiv_rn(raw, a.length[c])), a.length[c]));
        }
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
        const double r = sqrt(r2);
        if (r >= a.cutoff) continue;  // pairs.rs:186
        // functions.rs:80-88
        const double s = a.sigma / r;
        const double s2 = s * s;
        const double s6 = s2 * (s2 * s2);
        const double energy = 4.0 * a.epsilon * (s6 * s6 - s6) - a.shift;
        const double force = -24.0 * a.epsilon * (s6 - 2.0 * (s6 * s6)) / r;
        const double fr = force / r;
        if (a.write_forces) {
            atomicAdd(a.force + 3 * i, fr * d[0]);
            atomicAdd(a.force + 3 * i + 1, fr * d[1]);
            atomicAdd(a.force + 3 * i + 2, fr * d[2]);
        }
        if (a.full) {
            atomicAdd(a.results + RES_E_PAIRS, 0.5 * energy);
            atomicAdd(a.results + RES_PAIR_COUNT, 0.5);
            const double w = 0.5 * fr;
            atomicAdd(a.results + RES_W_PAIRS + 0, w * d[0] * d[0]);
            atomicAdd(a.results + RES_W_PAIRS + 1, w * d[0] * d[1]);
            atomicAdd(a.results + RES_W_PAIRS + 2, w * d[0] * d[2]);
            atomicAdd(a.results + RES_W_PAIRS + 3, w * d[1] * d[1]);
            atomicAdd(a.results + RES_W_PAIRS + 4, w * d[1] * d[2]);
            atomicAdd(a.results + RES_W_PAIRS + 5, w * d[2] * d[2]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------

__global__ void lj2_set_flag_kernel(int* flags, int index, int value) { flags[index] = value; }

static uint64_t mix2(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
}

bool lj2_enabled(const Context* ctx) {
    static const char* knob = std::getenv("LUMOL_CUDA_LJ2");
    if (knob != nullptr && knob[0] == '0') return false;
    // evaluations in the caller's atom order are not sharded on this path (the sorted-resident engine below is): with
    // several ranks they keep the first-generation kernels, whose atom-block ownership the reductions in api.cu expect
    return ctx->nranks == 1;
}

// Sizes, ownership and buffers of one evaluation.
struct Lj2Plan {
    UnitShape shape;
    int n, ncells, next, nunits, capacity, deferred_capacity, scan_blocks, ext_scan_blocks;
    size_t stride, fstride;
    GridView g;
    ExtGrid e;
    double cutoff, skin, radius, sigma, epsilon, scale, shift;
    int units_per_rank, u_lo, u_hi, s_lo, s_hi;  // units / sorted atoms of this rank
    int nrows, row_scan_blocks;
    uint64_t signature;
};

// `sharded`: the sorted-resident engine on several ranks (units split between the ranks); otherwise one rank does all.
static UnitShape lj2_shape() {
    UnitShape shape;
    shape.atoms = U_ATOMS;
    shape.lanes_per_atom = LJ2_LPA;
    shape.slots = LJ2_SLOTS;
    shape.value_factor = 3;
    shape.width = 3;
    shape.flags = 0;
    return shape;
}

// `cutoff`: the list radius is cutoff + skin.  `lj`: the single Lennard-Jones interaction (nullptr on the charged path).
static int lj2_plan(Context* ctx, bool sharded, const UnitShape& shape, double cutoff, const lumol_cuda_pair* lj, Lj2Plan& P) {
    const int n = (int)ctx->n;
    P.n = n;
    P.shape = shape;
    for (int d = 0; d < 3; d++) {
        P.g.nc[d] = ctx->ncell[d];
        P.g.length[d] = ctx->cell.h[4 * d];
        P.g.edge[d] = P.g.length[d] / (double)P.g.nc[d];
    }
    P.e.nx = P.g.nc[0];
    P.e.ny = P.g.nc[1];
    P.e.nz = P.g.nc[2];
    P.e.ex = P.e.nx + 2;
    P.e.ey = P.e.ny + 2;
    P.e.ez = P.e.nz + 2;
    P.ncells = P.g.nc[0] * P.g.nc[1] * P.g.nc[2];
    P.next = P.e.count();
    P.cutoff = cutoff;
    P.skin = ctx->skin_effective;
    P.radius = P.cutoff + P.skin;
    P.sigma = lj != nullptr ? lj->p[0] : 1.0;
    P.epsilon = lj != nullptr ? lj->p[1] : 0.0;
    P.scale = 1.0 / P.sigma;
    P.shift = lj != nullptr ? lj->shift : 0.0;

    const double volume = P.g.length[0] * P.g.length[1] * P.g.length[2];
    const double mean_neighbors = 4.0 / 3.0 * PI * P.radius * P.radius * P.radius * (double)n / volume;
    int capacity = (int)(2.0 * mean_neighbors) + 64;
    if (capacity > n) capacity = n;
    P.capacity = (capacity + 7) / 8 * 8;
    P.stride = ((size_t)n + 31) / 32 * 32;
    P.nunits = (n + shape.atoms - 1) / shape.atoms;
    P.nrows = P.g.nc[1] * P.g.nc[2];
    if (shape.flags & 4) P.nunits += P.nrows + 1;  // upper bound: the last unit of every row may be partial
    const int nranks = sharded ? ctx->nranks : 1;
    P.units_per_rank = (P.nunits + nranks - 1) / nranks;
    P.u_lo = sharded ? std::min(P.nunits, P.units_per_rank * ctx->rank) : 0;
    P.u_hi = sharded ? std::min(P.nunits, P.u_lo + P.units_per_rank) : P.nunits;
    P.s_lo = std::min(n, P.u_lo * shape.atoms);
    P.s_hi = (shape.flags & 4) ? n : std::min(n, P.u_hi * shape.atoms);
    // frame slots: the atoms, their ghost images (a boundary atom has up to seven), one slot of padding per extended cell
    const double ghost_ratio = (double)P.next / (double)P.ncells;
    size_t fstride = (size_t)((double)n * (ghost_ratio * 1.5 + 0.25)) + 2 * (size_t)P.next + 64;
    if (fstride > (size_t)8 * n + 2 * (size_t)P.next + 64) fstride = (size_t)8 * n + 2 * (size_t)P.next + 64;
    P.fstride = (fstride + 31) / 32 * 32;
    if (P.fstride >= (size_t)RAW_VALUE_MASK) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "too many atoms per GPU for the neighbour list (%d)", n);
    }
    LUMOL_CUDA_CHECK(ctx, ctx->nl_flags.reserve(16));
    LUMOL_CUDA_CHECK(ctx, ctx->nlist.reserve_zeroed(P.stride * (size_t)P.capacity, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, ctx->ncount.reserve(P.stride));
    LUMOL_CUDA_CHECK(ctx, ctx->xref.reserve((size_t)3 * n));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_of.reserve((size_t)2 * n));  // cell_of + slot_of
    LUMOL_CUDA_CHECK(ctx, ctx->cell_count.reserve((size_t)P.ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_start.reserve((size_t)P.ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->cell_needed.reserve((size_t)P.ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->order.reserve((size_t)2 * n));
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_f32.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sorted_cell.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->self_local.reserve_zeroed(P.stride, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, ctx->ext_start.reserve(2 * ((size_t)P.next + 2)));  // offsets, then the padded counts
    LUMOL_CUDA_CHECK(ctx, ctx->fidx.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->kshift.reserve((size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->frame_atom.reserve(P.fstride));
    LUMOL_CUDA_CHECK(ctx, ctx->frame_pos.reserve(2 * (size_t)shape.width * P.fstride));  // two copies: a sharded run alternates with the step parity
    LUMOL_CUDA_CHECK(ctx, ctx->blk_header.reserve((size_t)P.nunits));
    LUMOL_CUDA_CHECK(ctx, ctx->blk_entries.reserve_zeroed((size_t)P.nunits * LJ2_MAX_RUNS, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, ctx->cum_levels.reserve(P.stride * LJ2_LEVELS));
    P.deferred_capacity = n * DEFERRED_PER_ATOM + 1024;
    ctx->deferred_capacity = P.deferred_capacity;
    LUMOL_CUDA_CHECK(ctx, ctx->deferred.reserve((size_t)P.deferred_capacity));
    P.row_scan_blocks = (P.nrows + SCAN_BLOCK - 1) / SCAN_BLOCK;
    // row-aligned units: padded counts, first unit of every row (+ total), scan scratch, first atom of every unit (+ n)
    LUMOL_CUDA_CHECK(ctx, ctx->unit_rows.reserve(3 * ((size_t)P.nrows + 2) + (size_t)P.row_scan_blocks + 2 + (size_t)P.nunits + 2));
    P.scan_blocks = (P.ncells + SCAN_BLOCK - 1) / SCAN_BLOCK;
    P.ext_scan_blocks = (P.next + SCAN_BLOCK - 1) / SCAN_BLOCK;
    LUMOL_CUDA_CHECK(ctx, ctx->scan_scratch.reserve((size_t)(P.scan_blocks > P.ext_scan_blocks ? P.scan_blocks : P.ext_scan_blocks) + 1));

    uint64_t signature = mix2(0x4c4a32, (uint64_t)n);
    signature = mix2(signature, ctx->cell_generation);
    signature = mix2(signature, ctx->structure_generation);
    uint64_t bits;
    std::memcpy(&bits, &P.radius, sizeof(bits));
    signature = mix2(signature, bits);
    signature = mix2(signature, (uint64_t)P.capacity);
    signature = mix2(signature, (uint64_t)P.fstride);
    signature = mix2(signature, (uint64_t)P.u_lo * 1315423911ull + (uint64_t)P.u_hi);
    signature = mix2(signature, (uint64_t)shape.atoms * 64 + (uint64_t)shape.width * 8 + (uint64_t)shape.flags);
    if (shape.flags && ctx->nranks > 1) signature = mix2(signature, (uint64_t)ctx->rank * 977 + (uint64_t)ctx->nranks);
    P.signature = signature;
    if (!ctx->flags_initialised) {
        LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->nl_flags.ptr, 0, 16 * sizeof(int), ctx->stream));
        ctx->flags_initialised = true;
    }
    return 0;
}

static double* lj2_frame(Context* ctx, const Lj2Plan& P, int parity) { return ctx->frame_pos.ptr + (size_t)parity * P.shape.width * P.fstride; }

// Refresh of the frames from positions in the caller's order (state arrays not in cell order).
static int lj2_launch_update(Context* ctx, const Lj2Plan& P, int epoch) {
    Update2Args u;
    u.n = P.n;
    u.s_lo = 0;
    u.s_hi = P.n;
    u.g = P.g;
    u.e = P.e;
    u.order = ctx->order.ptr;
    u.pos = ctx->position.ptr;
    u.xref = ctx->xref.ptr;
    u.kshift = ctx->kshift.ptr;
    u.fidx = ctx->fidx.ptr;
    u.sorted_cell = ctx->sorted_cell.ptr;
    u.cell_start = ctx->cell_start.ptr;
    u.ext_start = ctx->ext_start.ptr;
    u.frame = lj2_frame(ctx, P, 0);
    u.fstride = P.fstride;
    u.width = P.shape.width;
    u.scale = P.scale;
    u.threshold2 = 0.25 * P.skin * P.skin;
    u.epoch = epoch;
    u.flags = ctx->nl_flags.ptr;
    lj2_update_kernel<<<(P.n + 255) / 256, 256, 0, ctx->stream>>>(u);
    ctx->launches++;
    ctx->clk_neighbor.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// The guarded rebuild (cooperative kernel + final order of the columns): does nothing unless flags[FLAG_REBUILD] holds
// this epoch.  `state`: positions the atoms are binned from (caller's order, or the old cell order of a sorted-resident run).
static int lj2_launch_rebuild(Context* ctx, const Lj2Plan& P, int epoch, const double* state, int parity,
                              const Rebuild2Args::Sorted& sorted, int* need_mask, int owner_lo = 0, int owner_hi = 0) {
    const int n = P.n;
    int* flags = ctx->nl_flags.ptr;
    int* cell_of = ctx->cell_of.ptr;
    int* slot_of = ctx->cell_of.ptr + n;
    int* order = ctx->order.ptr;
    int* grouped = ctx->order.ptr + n;
    Rebuild2Args r;
    r.n = n;
    r.ncells = P.ncells;
    r.scan_blocks = P.scan_blocks;
    r.ext_scan_blocks = P.ext_scan_blocks;
    r.nunits = P.nunits;
    r.epoch = epoch;
    r.g = P.g;
    r.e = P.e;
    r.position = state;
    r.cell_of = cell_of;
    r.slot_of = slot_of;
    r.cell_count = ctx->cell_count.ptr;
    r.cell_start = ctx->cell_start.ptr;
    r.scan_scratch = ctx->scan_scratch.ptr;
    r.grouped = grouped;
    r.ext_start = ctx->ext_start.ptr;
    r.ext_pad = ctx->ext_start.ptr + P.next + 2;
    r.nrows = P.nrows;
    r.row_scan_blocks = P.row_scan_blocks;
    r.row_pad = ctx->unit_rows.ptr;
    r.row_units = r.row_pad + P.nrows + 2;
    r.row_scratch = r.row_units + P.nrows + 2;
    r.unit_start = r.row_scratch + P.row_scan_blocks + 2;
    r.scatter.n = n;
    r.scatter.g = P.g;
    r.scatter.pos = state;
    r.scatter.state_of_sorted_old = nullptr;
    r.scatter.cell_of = cell_of;
    r.scatter.cell_start = ctx->cell_start.ptr;
    r.scatter.grouped = grouped;
    r.scatter.order = order;
    r.scatter.sorted_f32 = ctx->sorted_f32.ptr;
    r.scatter.sorted_cell = ctx->sorted_cell.ptr;
    r.scatter.xref = ctx->xref.ptr;
    r.scatter.kshift = ctx->kshift.ptr;
    r.scatter.kind = (P.shape.flags & 1) ? ctx->kind.ptr : nullptr;
    r.scatter.mol_first = ctx->mol_first.ptr;
    r.frame.n = n;
    r.frame.g = P.g;
    r.frame.e = P.e;
    r.frame.cell_start = ctx->cell_start.ptr;
    r.frame.sorted_cell = ctx->sorted_cell.ptr;
    r.frame.ext_start = ctx->ext_start.ptr;
    r.frame.order = order;
    r.frame.xref = ctx->xref.ptr;
    r.frame.kshift = ctx->kshift.ptr;
    r.frame.fidx = ctx->fidx.ptr;
    r.frame.frame_atom = ctx->frame_atom.ptr;
    r.frame.frame = lj2_frame(ctx, P, parity);
    r.frame.fstride = P.fstride;
    r.frame.scale = P.scale;
    r.frame.width = P.shape.width;
    r.frame.charge = ctx->charge.ptr;
    r.table.n = n;
    r.table.nunits = P.nunits;
    r.table.shape = P.shape;
    r.table.unit_start = r.unit_start;
    r.table.e = P.e;
    r.table.sorted_cell = ctx->sorted_cell.ptr;
    r.table.ext_start = ctx->ext_start.ptr;
    r.table.header = ctx->blk_header.ptr;
    r.table.runs = ctx->blk_entries.ptr;
    r.table.flags = flags;
    r.table.need_mask = need_mask;
    r.table.units_per_rank = P.units_per_rank;
    Build2Args& b = r.build;
    b.g = P.g;
    b.e = P.e;
    b.shape = P.shape;
    b.row_units = r.row_units;
    b.order = order;
    b.o_lo = owner_lo;
    b.o_hi = owner_hi;
    b.ncells = P.ncells;
    b.cells_per_warp = P.ncells * BUILD2_CHUNKS / (ctx->sm_count * 8 * REBUILD_WARPS * 16) + 1;
    b.s_lo = P.s_lo;
    b.s_hi = P.s_hi;
    b.capacity = P.capacity;
    b.radius2 = (float)(P.radius * P.radius * 1.0001);
    b.cutoff = (float)P.cutoff;
    b.inv_delta = P.skin > 0.0 ? (float)((double)LJ2_LEVELS / P.skin) : 0.0f;
    b.cell_start = ctx->cell_start.ptr;
    b.ext_start = ctx->ext_start.ptr;
    b.sorted_f32 = ctx->sorted_f32.ptr;
    b.header = ctx->blk_header.ptr;
    b.runs = ctx->blk_entries.ptr;
    b.self_slot = ctx->self_local.ptr;
    b.nlist = ctx->nlist.ptr;
    b.ncount = ctx->ncount.ptr;
    b.cell_needed = ctx->cell_needed.ptr;
    b.flags = flags;
    r.flags = flags;
    r.sorted = sorted;
    r.halo_rank = ctx->rank;
    r.halo_items = need_mask != nullptr ? ctx->sre_halo_items.ptr : nullptr;
    const void* rebuild = P.shape.flags == 0 ? (const void*)rebuild2_kernel<0> : (const void*)rebuild2_kernel<7>;
    int& grid = P.shape.flags == 0 ? ctx->rebuild2_grid : ctx->rebuild2_grid_flags;
    if (grid == 0) {
        int per_sm = 0;
        LUMOL_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rebuild, REBUILD_THREADS, 0));
        if (per_sm < 1) return ctx->fail(LUMOL_CUDA_ERROR_CUDA, "the rebuild kernel does not fit on the device");
        grid = per_sm * ctx->sm_count;
    }
    {
        void* params[] = {&r};
        LUMOL_CUDA_CHECK(ctx, cudaLaunchCooperativeKernel(rebuild, dim3(grid), dim3(REBUILD_THREADS), params, 0, ctx->stream));
    }
    ctx->launches++;
    ctx->clk_neighbor.launches++;
    if (P.shape.flags & 2) return 0;  // raw columns
    const size_t reorder_smem = (size_t)P.capacity * REORDER2_THREADS * sizeof(unsigned short);
    if (reorder_smem > 190 * 1024 || P.capacity > 65535) {
        return ctx->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "neighbour list columns of %d entries do not fit the reorder kernel", P.capacity);
    }
    const bool narrow = P.capacity <= 255;
    const void* reorder = narrow ? (const void*)list_reorder2_kernel<unsigned char> : (const void*)list_reorder2_kernel<unsigned short>;
    if (reorder_smem > 16 * 1024) {
        LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(reorder, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reorder_smem));
    }
    const int slabs = (P.s_hi - P.s_lo + REORDER2_THREADS - 1) / REORDER2_THREADS;
    const int reorder_grid = std::max(1, slabs < ctx->sm_count * 16 ? slabs : ctx->sm_count * 16);
    {
        UnitShape shape = P.shape;
        int s_lo = P.s_lo, s_hi = P.s_hi, capacity = P.capacity;
        const int4* header = ctx->blk_header.ptr;
        const int* ncount = ctx->ncount.ptr;
        unsigned* nlist = ctx->nlist.ptr;
        unsigned short* cum = ctx->cum_levels.ptr;
        int epoch_value = epoch;
        const int* flag_pointer = flags;
        void* params[] = {&shape, &s_lo, &s_hi, &capacity, &header, &ncount, &nlist, &cum, &epoch_value, &flag_pointer};
        LUMOL_CUDA_CHECK(ctx, cudaLaunchKernel(reorder, dim3(reorder_grid), dim3(REORDER2_THREADS), params, reorder_smem, ctx->stream));
    }
    ctx->launches++;
    ctx->clk_neighbor.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// Force kernel, reduction of the scalar sums, fix-up of the pairs on the cut-off.  `order`: sorted slot -> index in the
// state arrays (nullptr: the state arrays are in cell order).
static int lj2_launch_force(Context* ctx, const Lj2Plan& P, const ComputeRequest& req, int epoch, int parity, const int* order,
                            const double* state_position, double* state_force, bool fix_from_frames = false) {
    const bool full = req.energy || req.virial;
    int* flags = ctx->nl_flags.ptr;
    Lj2Args a;
    a.n = P.n;
    a.u_lo = P.u_lo;
    a.u_hi = P.u_hi;
    a.capacity = P.capacity;
    a.nlist = ctx->nlist.ptr;
    a.cum_levels = ctx->cum_levels.ptr;
    a.self_slot = ctx->self_local.ptr;
    a.fidx = ctx->fidx.ptr;
    a.order = order;
    a.header = ctx->blk_header.ptr;
    a.runs = ctx->blk_entries.ptr;
    a.frame = lj2_frame(ctx, P, parity);
    a.fstride = P.fstride;
    a.epsilon24 = 24.0 * P.epsilon;
    a.epsilon48 = 48.0 * P.epsilon;
    a.epsilon4 = 4.0 * P.epsilon;
    a.shift = P.shift;
    a.inv_sigma = P.scale;
    {
        const double reduced = P.cutoff * P.scale;
        const double reduced2 = reduced * reduced;
        uint64_t pattern;
        std::memcpy(&pattern, &reduced2, sizeof(pattern));
        a.band_lo = (int)(pattern >> 32) - 1;
    }
    a.inv_delta = P.skin > 0.0 ? (float)((double)LJ2_LEVELS / P.skin) : 0.0f;
    a.margin = 2.0e-3f;
    a.epoch = epoch;
    static const char* all_levels = std::getenv("LUMOL_CUDA_LJ2_ALL_LEVELS");  // experiments: 1 all levels, 2 no pairs at all
    a.all_levels = all_levels != nullptr ? std::atoi(all_levels) : 0;
    a.write_forces = req.forces;
    a.force = state_force;
    a.deferred = ctx->deferred.ptr;
    a.deferred_capacity = P.deferred_capacity;
    a.flags = flags;
    const int units = P.u_hi - P.u_lo;
    const int grid = std::max(1, units < ctx->sm_count ? units : ctx->sm_count);
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)grid * LJ2_NV));
    a.partials = ctx->partials.ptr;
    const void* kernel = full ? (const void*)lj2_force_kernel<1> : (const void*)lj2_force_kernel<0>;
    const size_t smem = (size_t)LJ2_BUFFERS * LJ2_STAGE_BYTES;
    LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ScopedClock clock(ctx, &ctx->clk_pair);
        void* params[] = {&a};
        LUMOL_CUDA_CHECK(ctx, cudaLaunchKernel(kernel, dim3(grid), dim3(LJ2_THREADS), params, smem, ctx->stream));
        ctx->launches++;
        ctx->clk_pair.launches++;
    }
    if (full) {
        int status = launch_reduce(ctx, grid, LJ2_NV, RES_E_PAIRS);
        if (status != 0) return status;
    }
    Fix2Args f;
    f.deferred = ctx->deferred.ptr;
    f.capacity = P.deferred_capacity;
    f.frame_atom = ctx->frame_atom.ptr;
    f.order = order;
    f.pos = state_position;
    f.frame = fix_from_frames ? lj2_frame(ctx, P, parity) : nullptr;
    f.fidx = ctx->fidx.ptr;
    f.sigma_for_frames = P.sigma;
    for (int d = 0; d < 3; d++) f.length[d] = P.g.length[d];
    f.sigma = P.sigma;
    f.epsilon = P.epsilon;
    f.cutoff = P.cutoff;
    f.shift = P.shift;
    f.full = full ? 1 : 0;
    f.write_forces = req.forces ? 1 : 0;
    f.force = state_force;
    f.results = ctx->results.ptr;
    f.flags = flags;
    lj2_fixup_kernel<<<8, 128, 0, ctx->stream>>>(f);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// One evaluation with the state arrays in the caller's order (lumol_cuda_compute, and MD outside the sorted-resident engine).
int launch_pairs_lj2(Context* ctx, const ComputeRequest& req) {
    Lj2Plan P;
    int status = lj2_plan(ctx, false, lj2_shape(), ctx->host_pairs[0].cutoff, &ctx->host_pairs[0], P);
    if (status != 0) return status;
    int* flags = ctx->nl_flags.ptr;
    const bool reuse = ctx->list_valid && P.signature == ctx->list_signature;
    ctx->list_epoch = ctx->list_epoch % 1000000000 + 1;
    const int epoch = ctx->list_epoch;
    {
        ScopedClock clock(ctx, &ctx->clk_neighbor);
        if (!reuse) {
            lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_REBUILD, epoch);
            lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_DEFERRED, 0);
            ctx->launches += 2;
        } else if ((status = lj2_launch_update(ctx, P, epoch)) != 0) {
            return status;
        }
        Rebuild2Args::Sorted none{};
        if ((status = lj2_launch_rebuild(ctx, P, epoch, ctx->position.ptr, 0, none, nullptr)) != 0) return status;
    }
    ctx->list_valid = true;
    ctx->list_signature = P.signature;
    return lj2_launch_force(ctx, P, req, epoch, 0, ctx->order.ptr, ctx->position.ptr, ctx->force.ptr);
}


// ------------------------------------------------------------------------------------------------
// sorted-resident molecular dynamics
// ------------------------------------------------------------------------------------------------
//
// lumol_cuda_md_run of a velocity-Verlet NVE run of a single-LJ system on the neighbour-list path
// (lumol-sim/src/md/molecular_dynamics.rs:66-76 around integrators.rs:39-69).  Between the entry and the exit of the call
// the state (x, v, f, m) lives in CELL ORDER: the force kernel stores forces with unit stride, the kick + drift kernel
// refreshes the frames (and the largest displacement) in the same pass over the atoms, and a rebuild moves the state into
// the new order inside the cooperative kernel.  Per step: kick-drift-refresh, rebuild guard (two launches that return at
// once), force kernel, fix-up.
//
// Several ranks (one process per GPU): a rank owns a contiguous range of units, i.e. a slab of cells, and integrates only
// its own atoms.  Every rank has the same frame layout (all of them sort all atoms at a rebuild), so the owner of an
// atom stores its new frame position, ghost images included, straight into the frames of the ranks whose units stage the
// atom's cell (NVLink peer stores, CUDA IPC mappings; need_mask from the unit tables).  Frames alternate between two
// copies with the step parity, so a rank one step ahead never overwrites what a neighbour is still reading.  The last
// block of the kick-drift kernel publishes, on every peer, the rank's largest squared displacement and an arrival stamp;
// a one-block kernel waits for the stamps of all ranks and takes the rebuild decision from the global maximum, so all
// ranks rebuild in the same step.  At a rebuild every rank stores the positions and velocities of its atoms into the
// state arrays of all peers (a device-side all-gather), then all ranks sort identically.  No NCCL call inside the loop.

constexpr int SRE_THREADS = 256;
constexpr int SYNC_ARRIVED = 0;    // [parity][rank]: epoch of the last kick-drift of that rank that has landed here
constexpr int SYNC_DISP = 32;      // [parity][rank]: float bits of its largest squared displacement
constexpr int SYNC_GATHERED = 64;  // [rank]: epoch of the last rebuild gather of that rank that has landed here
constexpr int SYNC_LOCAL = 96;     // [parity]: this rank's running maximum; +2: ticket counter; +3: gather ticket counter
constexpr int SYNC_INTS = 128;

struct SreArgs {
    int n, s_lo, s_hi;
    GridView g;
    ExtGrid e;
    double half_dt, dt;
    double* __restrict__ x;
    double* __restrict__ v;
    const double* __restrict__ f;
    const double* __restrict__ m;
    const double* __restrict__ xref;
    const int* __restrict__ kshift;
    const int* __restrict__ sorted_cell;
    const int* __restrict__ cell_start;
    const int* __restrict__ ext_start;
    size_t fstride;
    double scale, threshold2;
    int epoch, parity;
    int* __restrict__ flags;
    // sharding
    int nranks, rank;
    const int* __restrict__ need_mask;
    double* frame[PEER_MAX_RANKS];  // this parity's frame copy on every rank
    int* sync[PEER_MAX_RANKS];      // the SYNC_* block of every rank
};

// integrators.rs:47-53 (and :55-68 of the previous step when MERGED): a = f / m; v += (0.5 dt) a; x += v dt, with the
// reference's roundings; then the frame images of the new position and the displacement since the rebuild.  One thread per
// atom (a thread-per-component variant with fully coalesced accesses was slower: 0.087 against 0.048 ms for 1M atoms, the
// cell arithmetic of the ghost images was then done three times).
constexpr int SRE_ATOMS = SRE_THREADS;

template <bool MERGED>
__global__ void __launch_bounds__(SRE_THREADS) sre_kick_drift_kernel(SreArgs a) {
    __shared__ float block_max[SRE_THREADS / 32];
    const int s = a.s_lo + blockIdx.x * SRE_ATOMS + threadIdx.x;
    int* local = a.sync[a.rank];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.flags[FLAG_DEFERRED] = 0;
        if (a.nranks == 1) a.flags[FLAG_DISP + (a.parity ^ 1)] = 0;  // the slot of the next evaluation
    }
    float d2f = 0.0f;
    if (s < a.s_hi) {
        const double mass = a.m[s];
        double p[3];
        double d2 = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t k = 3 * (size_t)s + c;
            const double kick = __dmul_rn(a.half_dt, __ddiv_rn(a.f[k], mass));
            double v = __dadd_rn(a.v[k], kick);
            if (MERGED) v = __dadd_rn(v, kick);
            p[c] = __dadd_rn(a.x[k], __dmul_rn(v, a.dt));
            a.v[k] = v;
            a.x[k] = p[c];
            const double d = p[c] - a.xref[k];
            d2 += d * d;
        }
        if (a.nranks == 1 && !(d2 <= a.threshold2)) a.flags[FLAG_REBUILD] = a.epoch;  // also catches NaN
        d2f = __double2float_ru(d2);
        if (!(d2f >= 0.0f)) d2f = 3.0e38f;
        int kx, ky, kz;
        unpack_shift(a.kshift[s], kx, ky, kz);
        const double x = p[0] - (double)kx * a.g.length[0];
        const double y = p[1] - (double)ky * a.g.length[1];
        const double z = p[2] - (double)kz * a.g.length[2];
        const int cell = a.sorted_cell[s];
        const int rank_in_cell = s - a.cell_start[cell];
        // the frames of this rank; the halo kernel below copies what the other ranks stage
        double* frame = a.frame[a.rank];
        for_each_image(a.e, a.ext_start, cell, rank_in_cell, [&](int f, int sx, int sy, int sz) {
            frame[3 * (size_t)f] = (x + (double)sx * a.g.length[0]) * a.scale;
            frame[3 * (size_t)f + 1] = (y + (double)sy * a.g.length[1]) * a.scale;
            frame[3 * (size_t)f + 2] = (z + (double)sz * a.g.length[2]) * a.scale;
        });
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2f = fmaxf(d2f, __shfl_xor_sync(0xffffffffu, d2f, o));
    if ((threadIdx.x & 31) == 0) block_max[threadIdx.x >> 5] = d2f;
    __syncthreads();
    if (a.nranks == 1) {
        if (threadIdx.x == 0) {
            float m = block_max[0];
            for (int w = 1; w < SRE_THREADS / 32; w++) m = fmaxf(m, block_max[w]);
            atomicMax(a.flags + FLAG_DISP + a.parity, __float_as_int(m));
        }
        return;
    }
    // sharded: this rank's running maximum; the halo kernel publishes it with the frames
    if (threadIdx.x == 0) {
        float m = block_max[0];
        for (int w = 1; w < SRE_THREADS / 32; w++) m = fmaxf(m, block_max[w]);
        atomicMax(local + SYNC_LOCAL + a.parity, __float_as_int(m));
    }
}

// Sharded: the new frame positions of this rank's atoms (ghost images included) into the frames of the ranks that stage
// them.  One warp per owned cell and image: a cell's atoms are contiguous in the frames, so the peer stores are runs of
// consecutive doubles (per-atom stores from the kick-drift kernel took 28 us for 0.5 M atoms on two ranks).  The last
// block publishes, on every rank, this rank's largest squared displacement and its arrival stamp.
struct SreHaloArgs {
    const int4* __restrict__ items;  // (first double, doubles, peers): built at the rebuild
    const int* __restrict__ flags;
    int epoch, parity, nranks, rank;
    double* frame[PEER_MAX_RANKS];
    int* sync[PEER_MAX_RANKS];
};

__global__ void __launch_bounds__(SRE_THREADS) sre_halo_kernel(SreHaloArgs a) {
    __shared__ bool last_block;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (SRE_THREADS / 32) + (threadIdx.x >> 5), nwarps = gridDim.x * (SRE_THREADS / 32);
    const double* mine = a.frame[a.rank];
    const int nitems = a.flags[FLAG_HALO_ITEMS];
    for (int item = warp; item < nitems; item += nwarps) {
        const int4 run = a.items[item];
        for (int r = 0; r < a.nranks; r++) {
            if (!(run.z & (1 << r))) continue;
            double* peer = a.frame[r];
            for (int k = lane; k < run.y; k += 32) peer[(size_t)run.x + k] = mine[(size_t)run.x + k];
        }
    }
    __syncthreads();
    int* local = a.sync[a.rank];
    if (threadIdx.x == 0) {
        __threadfence_system();  // cumulative: covers the peer stores of the block, ordered before this thread by the barrier
        last_block = atomicAdd(local + SYNC_LOCAL + 2, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (last_block && threadIdx.x < a.nranks) {
        const int value = *(volatile int*)(local + SYNC_LOCAL + a.parity);
        int* peer = a.sync[threadIdx.x];
        *(volatile int*)(peer + SYNC_DISP + a.parity * PEER_MAX_RANKS + a.rank) = value;
        __threadfence_system();
        *(volatile int*)(peer + SYNC_ARRIVED + a.parity * PEER_MAX_RANKS + a.rank) = a.epoch;
    }
    if (last_block) {
        __syncthreads();
        if (threadIdx.x == 0) {
            local[SYNC_LOCAL + 2] = 0;
            local[SYNC_LOCAL + a.parity] = 0;  // read above; this parity is used again two steps from now
        }
    }
}

// Sharded: waits for the kick-drift of every rank (their halo frames are then in place), takes the rebuild decision from
// the largest displacement of all atoms of all ranks.  One warp.
__global__ void sre_sync_kernel(int nranks, int epoch, int parity, float threshold2, const int* __restrict__ sync,
                                int* __restrict__ flags, double* __restrict__ results) {
    const int lane = threadIdx.x;
    int value = 0, ok = 1;
    if (lane < nranks) {
        const volatile int* stamp = sync + SYNC_ARRIVED + parity * PEER_MAX_RANKS + lane;
        const long long start = clock64();
        while (*stamp != epoch) {
            if (clock64() - start > 40000000000ll) {
                ok = 0;
                break;
            }
        }
        __threadfence_system();
        value = *(const volatile int*)(sync + SYNC_DISP + parity * PEER_MAX_RANKS + lane);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        value = max(value, __shfl_xor_sync(0xffffffffu, value, o));
        ok = min(ok, __shfl_xor_sync(0xffffffffu, ok, o));
    }
    if (lane == 0) {
        flags[FLAG_DISP + parity] = value;
        flags[FLAG_DISP + (parity ^ 1)] = 0;
        const float d2 = __int_as_float(value);
        if (!(d2 <= threshold2)) flags[FLAG_REBUILD] = epoch;
        if (!ok) results[RES_FLAGS + 1] = 1.0;
    }
}

// Sharded, at a rebuild: the positions and velocities of this rank's atoms into the state arrays of every other rank.
struct SreGatherArgs {
    int epoch, nranks, rank;
    size_t lo3, hi3;
    const double* __restrict__ x;
    const double* __restrict__ v;
    double* peer_x[PEER_MAX_RANKS];
    double* peer_v[PEER_MAX_RANKS];
    int* sync[PEER_MAX_RANKS];
    const int* __restrict__ flags;
};

__global__ void __launch_bounds__(SRE_THREADS) sre_gather_kernel(SreGatherArgs a) {
    if (a.flags[FLAG_REBUILD] != a.epoch) return;
    __shared__ bool last_block;
    for (size_t k = a.lo3 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.hi3; k += (size_t)gridDim.x * blockDim.x) {
        const double x = a.x[k], v = a.v[k];
        for (int r = 0; r < a.nranks; r++) {
            if (r == a.rank) continue;
            a.peer_x[r][k] = x;
            a.peer_v[r][k] = v;
        }
    }
    __syncthreads();
    int* local = a.sync[a.rank];
    if (threadIdx.x == 0) {
        __threadfence_system();  // cumulative: covers the peer stores of the block, ordered before this thread by the barrier
        last_block = atomicAdd(local + SYNC_LOCAL + 3, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (last_block) {
        if (threadIdx.x < a.nranks) {
            __threadfence_system();
            *(volatile int*)(a.sync[threadIdx.x] + SYNC_GATHERED + a.rank) = a.epoch;
        }
        if (threadIdx.x == 0) local[SYNC_LOCAL + 3] = 0;
    }
}

// state <-> caller's order
__global__ void __launch_bounds__(SRE_THREADS)
    sre_enter_kernel(int n, const int* __restrict__ order, const double* __restrict__ position, const double* __restrict__ velocity,
                     const double* __restrict__ force, const double* __restrict__ mass, double* __restrict__ x, double* __restrict__ v,
                     double* __restrict__ f, double* __restrict__ m, int* __restrict__ origin) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = order[s];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        x[3 * (size_t)s + c] = position[3 * (size_t)i + c];
        v[3 * (size_t)s + c] = velocity[3 * (size_t)i + c];
        f[3 * (size_t)s + c] = force[3 * (size_t)i + c];
    }
    m[s] = mass[i];
    origin[s] = i;
}

__global__ void __launch_bounds__(SRE_THREADS)
    sre_leave_kernel(int n, const int* __restrict__ origin, const double* __restrict__ x, const double* __restrict__ v,
                     const double* __restrict__ f, double* __restrict__ position, double* __restrict__ velocity,
                     double* __restrict__ force, int* __restrict__ order) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = origin[s];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        position[3 * (size_t)i + c] = x[3 * (size_t)s + c];
        velocity[3 * (size_t)i + c] = v[3 * (size_t)s + c];
        force[3 * (size_t)i + c] = f[3 * (size_t)s + c];
    }
    order[s] = i;  // the list stays valid for evaluations in the caller's order
}

// integrators.rs:55-68: the second half kick of the last step
__global__ void __launch_bounds__(SRE_THREADS)
    sre_kick_kernel(size_t lo3, size_t hi3, double half_dt, const double* __restrict__ f, const double* __restrict__ m, double* __restrict__ v) {
    for (size_t k = lo3 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi3; k += (size_t)gridDim.x * blockDim.x) {
        v[k] = __dadd_rn(v[k], __dmul_rn(half_dt, __ddiv_rn(f[k], m[k / 3])));
    }
}

bool sorted_md_applicable(Context* ctx) {
    const char* knob = std::getenv("LUMOL_CUDA_SORTED_MD");  // read at every call: the tests switch it
    if (knob != nullptr && knob[0] == '0') return false;
    static const char* lj2 = std::getenv("LUMOL_CUDA_LJ2");
    if (lj2 != nullptr && lj2[0] == '0') return false;
    if (ctx->integrator != LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET || ctx->thermostat != LUMOL_CUDA_THERMOSTAT_NONE || ctx->controls != 0) return false;
    if (!(ctx->any_pair && ctx->single_lj && ctx->coulomb.kind == 0) || ctx->nbonds + ctx->nangles + ctx->ndihedrals != 0) return false;
    if (ctx->forced_path == 0 || ctx->forced_path == 2 || ctx->nranks > PEER_MAX_RANKS) return false;
    if (ctx->n >= (int64_t)RAW_VALUE_MASK) return false;
    return choose_neighbor_path(ctx, ctx->max_pair_cutoff) == 1;
}

static int sorted_md_run_checked(Context* ctx, int64_t nsteps);

int sorted_md_run(Context* ctx, int64_t nsteps) {
    const int status = sorted_md_run_checked(ctx, nsteps);
    if (status != 0) ctx->list_valid = false;  // the order map may describe the state arrays, not the caller's
    return status;
}

static int sorted_md_run_checked(Context* ctx, int64_t nsteps) {
    const bool sharded = ctx->nranks > 1;
    ctx->path = 1;
    ctx->lj2_active = true;
    Lj2Plan P;
    int status = lj2_plan(ctx, sharded, lj2_shape(), ctx->host_pairs[0].cutoff, &ctx->host_pairs[0], P);
    if (status != 0) return status;
    const int n = P.n;
    int* flags = ctx->nl_flags.ptr;
    // state arrays padded to whole rank blocks (the exit all-gather moves equal blocks)
    const size_t padded = (size_t)P.units_per_rank * U_ATOMS * (size_t)(sharded ? ctx->nranks : 1);
    LUMOL_CUDA_CHECK(ctx, ctx->sre_x.reserve(3 * padded));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_v.reserve(3 * padded));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_f.reserve(3 * padded));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_m.reserve(padded));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_tmp.reserve(4 * (size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_origin.reserve(2 * (size_t)n));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_need_mask.reserve((size_t)P.ncells + 1));
    LUMOL_CUDA_CHECK(ctx, ctx->sre_halo_items.reserve(8 * (size_t)P.ncells + 8));
    const bool fresh_sync = ctx->sre_sync.ptr == nullptr;
    LUMOL_CUDA_CHECK(ctx, ctx->sre_sync.reserve(SYNC_INTS));
    if (fresh_sync) LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->sre_sync.ptr, 0, SYNC_INTS * sizeof(int), ctx->stream));

    // ---- peers ------------------------------------------------------------------------------------------
    void* peers[4 * PEER_MAX_RANKS] = {};
    if (sharded) {
        void* local[4] = {ctx->sre_x.ptr, ctx->sre_v.ptr, ctx->frame_pos.ptr, ctx->sre_sync.ptr};
        bool ok = false;
        LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        if ((status = comm_map_peer_buffers(ctx, 4, local, peers, &ok)) != 0) return status;
        if (!ok) return ctx->fail(LUMOL_CUDA_ERROR_COMM, "peer memory (CUDA IPC over NVLink) is not available between the ranks");
    } else {
        peers[0 * PEER_MAX_RANKS] = ctx->sre_x.ptr;
        peers[1 * PEER_MAX_RANKS] = ctx->sre_v.ptr;
        peers[2 * PEER_MAX_RANKS] = ctx->frame_pos.ptr;
        peers[3 * PEER_MAX_RANKS] = ctx->sre_sync.ptr;
    }

    // ---- entry: a list valid for the current positions (caller's order), then the state into cell order -----------
    const bool reuse = ctx->list_valid && P.signature == ctx->list_signature;
    ctx->list_epoch = ctx->list_epoch % 1000000000 + 1;
    int epoch = ctx->list_epoch;
    {
        ScopedClock clock(ctx, &ctx->clk_neighbor);
        if (!reuse) {
            lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_REBUILD, epoch);
            lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_DEFERRED, 0);
            ctx->launches += 2;
        } else if ((status = lj2_launch_update(ctx, P, epoch)) != 0) {
            return status;
        }
        Rebuild2Args::Sorted none{};
        if ((status = lj2_launch_rebuild(ctx, P, epoch, ctx->position.ptr, 0, none, sharded ? ctx->sre_need_mask.ptr : nullptr)) != 0) return status;
    }
    ctx->list_valid = true;
    ctx->list_signature = P.signature;
    const int blocks_all = (n + SRE_THREADS - 1) / SRE_THREADS;
    sre_enter_kernel<<<blocks_all, SRE_THREADS, 0, ctx->stream>>>(n, ctx->order.ptr, ctx->position.ptr, ctx->velocity.ptr, ctx->force.ptr,
                                                                  ctx->mass.ptr, ctx->sre_x.ptr, ctx->sre_v.ptr, ctx->sre_f.ptr, ctx->sre_m.ptr,
                                                                  ctx->sre_origin.ptr);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    if (sharded) {
        // nobody pushes into a peer whose buffers are not in place yet
        LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->results.ptr + RES_FLAGS + 1, 0, sizeof(double), ctx->stream));
        if ((status = comm_barrier(ctx)) != 0) return status;
    }

    // ---- steps -------------------------------------------------------------------------------------------------
    const int owned = P.s_hi - P.s_lo;
    const int blocks_owned = std::max(1, (owned + SRE_ATOMS - 1) / SRE_ATOMS);
    const int blocks_kick = std::max(1, std::min(ctx->sm_count * 8, (3 * owned + SRE_THREADS - 1) / SRE_THREADS));
    ComputeRequest req;
    req.forces = true;
    req.pairs = true;
    for (int64_t step = 0; step < nsteps; step++) {
        ctx->list_epoch = ctx->list_epoch % 1000000000 + 1;
        epoch = ctx->list_epoch;
        const int parity = epoch & 1;
        {
            ScopedClock clock(ctx, &ctx->clk_integrate);
            SreArgs a;
            a.n = n;
            a.s_lo = P.s_lo;
            a.s_hi = P.s_hi;
            a.g = P.g;
            a.e = P.e;
            a.half_dt = 0.5 * ctx->dt;
            a.dt = ctx->dt;
            a.x = ctx->sre_x.ptr;
            a.v = ctx->sre_v.ptr;
            a.f = ctx->sre_f.ptr;
            a.m = ctx->sre_m.ptr;
            a.xref = ctx->xref.ptr;
            a.kshift = ctx->kshift.ptr;
            a.sorted_cell = ctx->sorted_cell.ptr;
            a.cell_start = ctx->cell_start.ptr;
            a.ext_start = ctx->ext_start.ptr;
            a.fstride = P.fstride;
            a.scale = P.scale;
            a.threshold2 = 0.25 * P.skin * P.skin;
            a.epoch = epoch;
            a.parity = parity;
            a.flags = flags;
            a.nranks = sharded ? ctx->nranks : 1;
            a.rank = sharded ? ctx->rank : 0;
            a.need_mask = ctx->sre_need_mask.ptr;
            for (int r = 0; r < a.nranks; r++) {
                a.frame[r] = (double*)peers[2 * PEER_MAX_RANKS + r] + (size_t)parity * 3 * P.fstride;
                a.sync[r] = (int*)peers[3 * PEER_MAX_RANKS + r];
            }
            if (step == 0) {
                sre_kick_drift_kernel<false><<<blocks_owned, SRE_THREADS, 0, ctx->stream>>>(a);
            } else {
                sre_kick_drift_kernel<true><<<blocks_owned, SRE_THREADS, 0, ctx->stream>>>(a);
            }
            ctx->launches++;
            ctx->clk_integrate.launches++;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        if (sharded) {
            ScopedClock clock(ctx, &ctx->clk_comm);
            SreHaloArgs h;
            h.items = ctx->sre_halo_items.ptr;
            h.flags = flags;
            h.epoch = epoch;
            h.parity = parity;
            h.nranks = ctx->nranks;
            h.rank = ctx->rank;
            for (int r = 0; r < ctx->nranks; r++) {
                h.frame[r] = (double*)peers[2 * PEER_MAX_RANKS + r] + (size_t)parity * 3 * P.fstride;
                h.sync[r] = (int*)peers[3 * PEER_MAX_RANKS + r];
            }
            sre_halo_kernel<<<ctx->sm_count, SRE_THREADS, 0, ctx->stream>>>(h);
            ctx->launches++;
            ctx->clk_comm.launches++;
            sre_sync_kernel<<<1, 32, 0, ctx->stream>>>(ctx->nranks, epoch, parity, (float)(0.25 * P.skin * P.skin), ctx->sre_sync.ptr, flags,
                                                       ctx->results.ptr);
            SreGatherArgs g;
            g.epoch = epoch;
            g.nranks = ctx->nranks;
            g.rank = ctx->rank;
            g.lo3 = 3 * (size_t)P.s_lo;
            g.hi3 = 3 * (size_t)P.s_hi;
            g.x = ctx->sre_x.ptr;
            g.v = ctx->sre_v.ptr;
            for (int r = 0; r < ctx->nranks; r++) {
                g.peer_x[r] = (double*)peers[0 * PEER_MAX_RANKS + r];
                g.peer_v[r] = (double*)peers[1 * PEER_MAX_RANKS + r];
                g.sync[r] = (int*)peers[3 * PEER_MAX_RANKS + r];
            }
            g.flags = flags;
            sre_gather_kernel<<<ctx->sm_count, SRE_THREADS, 0, ctx->stream>>>(g);
            ctx->launches += 2;
            ctx->clk_comm.launches += 2;
            LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
        }
        {
            ScopedClock clock(ctx, &ctx->clk_neighbor);
            Rebuild2Args::Sorted sorted{};
            sorted.active = 1;
            sorted.x = ctx->sre_x.ptr;
            sorted.v = ctx->sre_v.ptr;
            sorted.m = ctx->sre_m.ptr;
            sorted.tmp = ctx->sre_tmp.ptr;
            sorted.origin = ctx->sre_origin.ptr;
            sorted.tmp_origin = ctx->sre_origin.ptr + n;
            sorted.nranks = sharded ? ctx->nranks : 1;
            sorted.gathered = ctx->sre_sync.ptr + SYNC_GATHERED;

            if ((status = lj2_launch_rebuild(ctx, P, epoch, ctx->sre_x.ptr, parity, sorted, sharded ? ctx->sre_need_mask.ptr : nullptr)) != 0) {
                return status;
            }
        }
        if ((status = lj2_launch_force(ctx, P, req, epoch, parity, nullptr, ctx->sre_x.ptr, ctx->sre_f.ptr, sharded)) != 0) return status;
        ctx->md_step++;
    }
    if (nsteps > 0) {
        ScopedClock clock(ctx, &ctx->clk_integrate);
        sre_kick_kernel<<<blocks_kick, SRE_THREADS, 0, ctx->stream>>>(3 * (size_t)P.s_lo, 3 * (size_t)P.s_hi, 0.5 * ctx->dt, ctx->sre_f.ptr,
                                                                      ctx->sre_m.ptr, ctx->sre_v.ptr);
        ctx->launches++;
        ctx->clk_integrate.launches++;
        LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    }

    // ---- exit: every rank gets every atom back, in the caller's order ---------------------------------------------------
    if (sharded) {
        const size_t chunk = 3 * (size_t)P.units_per_rank * U_ATOMS;
        if ((status = comm_allgather_chunks(ctx, ctx->sre_x.ptr, chunk)) != 0) return status;
        if ((status = comm_allgather_chunks(ctx, ctx->sre_v.ptr, chunk)) != 0) return status;
        if ((status = comm_allgather_chunks(ctx, ctx->sre_f.ptr, chunk)) != 0) return status;
    }
    sre_leave_kernel<<<blocks_all, SRE_THREADS, 0, ctx->stream>>>(n, ctx->sre_origin.ptr, ctx->sre_x.ptr, ctx->sre_v.ptr, ctx->sre_f.ptr,
                                                                  ctx->position.ptr, ctx->velocity.ptr, ctx->force.ptr, ctx->order.ptr);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------
// charged systems: Lennard-Jones (or null) pairs + Ewald real space / Wolf in one branch-free pass
// ------------------------------------------------------------------------------------------------
//
// Replaces, for "water-like" systems on the neighbour-list path, the loops of sys/compute.rs:37-55 + ewald.rs:430-530 /
// wolf.rs:177-283: every pair entry Lennard-Jones, null or absent with restriction None, at most four particle kinds,
// coulomb restriction None or InterMolecular, alpha * rc <= 3.4.  Everything else keeps list_force_kernel.
//
// The first-generation kernel gathered 48 bytes per neighbour from L1 / L2 and branched on kind, restriction and charge:
// 18 of 32 lanes active, FP64 pipe 27 % busy (profiles/r1y_list_force_kernel_spce98k_summary.csv).  Here a unit of 256
// consecutive atoms stages its neighbourhood once, as (x, y, z, charge) quadruples (one bulk copy per row of cells), the
// 16-bit list entries carry the neighbour's kind and a same-molecule bit next to the slot, and one straight-line body
// evaluates both terms for every listed pair:
//   * one reciprocal square root (hardware seed + one cubic step) serves r, 1 / r and 1 / r^2;
//   * epsilon = 0 (pairs without Lennard-Jones term, or outside their cut-off) and q_i q_j = 0 (zero charge, outside the
//     coulomb cut-off, excluded Wolf pair) switch the terms off arithmetically;
//   * ONE exponential exp(-(alpha r)^2) per pair serves the Gaussian term and erfc(alpha r) = exp(-(alpha r)^2) erfcx(alpha r),
//     erfcx by a degree-14 polynomial in (x - 4) / (x + 4) (tools/fit_erfcx.py: 3.4e-15 relative); an excluded Ewald pair
//     takes erfc - 1 = -erf (ewald.rs:395-399, 419-425);
//   * pairs whose r^2 falls in the narrow band around a cut-off are left to the fix-up kernel, which evaluates them with
//     the reference's own arithmetic (`r >= rc` for the pair potential, pairs.rs:186; `r > rc` for Ewald and Wolf,
//     ewald.rs:390, wolf.rs:91).

constexpr int CQ_THREADS = 512;
constexpr int CQ_LPA = 2;
constexpr int CQ_ATOMS = CQ_THREADS / CQ_LPA;
constexpr int CQ_SLOTS = 6784;  // quadruples of one staged copy: 212 KiB
constexpr int CQ_NV = 16;
constexpr int CQ_MAX_KINDS = 4;
constexpr double CQ_XMAX = 3.45;

// tools/fit_erfcx.py
constexpr double CQ_ERFCX_A = 2.1594202898550723, CQ_ERFCX_B = 1.1594202898550725;
__constant__ double CQ_ERFCX[15] = {
    0.3773882991430218,     -0.3429947340586216,   0.17661230872576122,    -0.07210396072897172,  0.02349570533116956,
    -0.00604706965782215,   0.0011867532468606704, -0.00016177461783462791, 1.0614456744267083e-05, 9.443963073478152e-07,
    -2.7237999019525457e-07, 7.795810757036897e-09, 4.342510292920047e-09,  -3.307242319007307e-10, -6.793336512007265e-11,
};

struct CqClass {
    double sigma2, eps24, eps4, shift;
    int band_lo;  // top 32 bits of rc^2, minus one; INT_MIN when the pair has no Lennard-Jones term
    int pad;
};

struct CqArgs {
    int n, capacity;
    const int* __restrict__ unit_start;  // first atom of every unit (row-aligned units)
    const unsigned* __restrict__ nlist;
    const int* __restrict__ ncount;
    const unsigned short* __restrict__ self_slot;
    const int* __restrict__ fidx;
    const int* __restrict__ order;
    const float4* __restrict__ tags;  // sorted_f32: w holds (molecule << 2) | kind
    const int4* __restrict__ header;
    const int4* __restrict__ runs;
    const double* __restrict__ frame;  // (x, y, z, charge) per frame slot
    CqClass classes[CQ_MAX_KINDS * CQ_MAX_KINDS];
    int coulomb_kind;        // 0 none, 1 Ewald, 2 Wolf
    int exclude_same;        // the coulomb restriction is InterMolecular
    int coulomb_band_lo;
    int top_cap;             // top 32 bits of the largest argument of the erfc / exp approximations
    double alpha, gauss, wolf_energy, wolf_force;  // gauss = alpha 2 / sqrt(pi)
    int o_lo, o_hi;          // state-index range owned by this rank
    int write_forces;
    double* __restrict__ force;
    double* __restrict__ partials;
    int2* __restrict__ deferred;
    int deferred_capacity;
    int* __restrict__ flags;
    int* __restrict__ unit_counter;
};

// exp(a) for a in [-12, 0]: a = k ln2 + r, |r| <= ln2 / 2, Taylor polynomial of degree 12 (1.7e-16), 2^k by exponent arithmetic
__device__ __forceinline__ double cq_exp(double a) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: the integer lands in the low word
    const double shifted = fma(a, 1.4426950408889634, magic);
    const int k = __double2loint(shifted);
    const double kd = shifted - magic;
    double r = fma(kd, -6.93147180369123816490e-01, a);
    r = fma(kd, -1.90821492927058770002e-10, r);
    // Horner: Estrin's scheme (shorter dependent chains, three more multiplications) was slower, 0.375 against 0.355 ms on the
    // 98k-atom SPC/E box: the FP64 pipe, not the chain latency, is the limit
    double p = 1.0 / 479001600.0;
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// erfcx(x) = exp(x^2) erfc(x), x in [0, CQ_XMAX]
__device__ __forceinline__ double cq_erfcx(double x) {
    const double d = x + 4.0;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    y = fma(y, fma(e, e, e), y);
    const double zz = fma((x - 4.0) * y, CQ_ERFCX_A, CQ_ERFCX_B);
    double p = CQ_ERFCX[14];
#pragma unroll
    for (int k = 13; k >= 0; k--) p = fma(p, zz, CQ_ERFCX[k]);
    return p;
}

struct CqAtom {
    double x, y, z, q;        // q already divided by 4 pi epsilon_0
    const CqClass* row;       // classes of this atom's kind, in shared memory
};

// One listed neighbour.  `entry`: slot (13 bits) | kind of j (2 bits) | same molecule (1 bit).
template <int MODE>
__device__ __forceinline__ void cq_pair(const CqArgs& a, const CqAtom& me, double xj, double yj, double zj, double qj, unsigned kind_j,
                                        bool same, double& fx, double& fy, double& fz, double (&acc)[CQ_NV], unsigned& near, unsigned bit) {
    const double dx = me.x - xj, dy = me.y - yj, dz = me.z - zj;
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const int top = __double2hiint(r2);
    const CqClass& c = me.row[kind_j];
    const bool lj = c.band_lo != INT_MIN;
    const bool on_cutoff = (lj && (unsigned)(top - c.band_lo) < 3u) || (a.coulomb_kind != 0 && (unsigned)(top - a.coulomb_band_lo) < 3u);
    near |= on_cutoff ? bit : 0u;
    const bool inside_pair = lj && top < c.band_lo && !on_cutoff;
    const bool inside_coulomb = top < a.coulomb_band_lo && !on_cutoff;
    // 1 / r: hardware seed and one cubic step, y0 (1 + e / 2 + 3 e^2 / 8) with e = 1 - r2 y0^2
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
    const double e = fma(-(r2 * y0), y0, 1.0);
    const double rinv = fma(y0 * e, fma(e, 0.375, 0.5), y0);
    const double rinv2 = rinv * rinv;
    const double r = r2 * rinv;
    // Lennard-Jones (functions.rs:80-88): force / r = 24 eps (2 s6^2 - s6) / r^2
    const double s2 = c.sigma2 * rinv2;
    const double s6 = s2 * s2 * s2;
    const double eps24 = inside_pair ? c.eps24 : 0.0;
    double fr = eps24 * (s6 * fma(2.0, s6, -1.0)) * rinv2;
    // coulomb (ewald.rs:387-428, wolf.rs:90-117)
    double qq = me.q * qj;
    const bool wolf_skip = a.coulomb_kind == 2 && same && a.exclude_same;
    qq = (inside_coulomb && !wolf_skip) ? qq : 0.0;
    const double excluded = (a.coulomb_kind == 1 && same && a.exclude_same) ? 1.0 : 0.0;
    // alpha r capped (integer minimum on the top word) to the range of the erfc / exp approximations: only pairs outside the
    // coulomb cut-off (qq = 0), e.g. the far-away slot that padding entries point at, are affected
    const double x_true = a.alpha * r;
    const double x = __hiloint2double(min(__double2hiint(x_true), a.top_cap), __double2loint(x_true));
    const double t = cq_exp(-(x * x));
    const double eor = fma(t, cq_erfcx(x), -excluded) * rinv;  // (erfc(alpha r) - excluded) / r
    fr += qq * fma(rinv2, fma(a.gauss, t, eor), -a.wolf_force * rinv);
    fx = fma(fr, dx, fx);
    fy = fma(fr, dy, fy);
    fz = fma(fr, dz, fz);
    if (MODE == 1) {
        const double fl = eps24 * (s6 * fma(2.0, s6, -1.0)) * rinv2;
        const double fc = fr - fl;
        acc[0] += inside_pair ? c.eps4 * fma(s6, s6, -s6) - c.shift : 0.0;
        acc[14] += inside_pair ? 1.0 : 0.0;
        acc[2] += fl * dx * dx;
        acc[3] += fl * dx * dy;
        acc[4] += fl * dx * dz;
        acc[5] += fl * dy * dy;
        acc[6] += fl * dy * dz;
        acc[7] += fl * dz * dz;
        acc[1] += qq * (eor - a.wolf_energy);
        acc[15] += qq != 0.0 ? 1.0 : 0.0;
        acc[8] += fc * dx * dx;
        acc[9] += fc * dx * dy;
        acc[10] += fc * dx * dz;
        acc[11] += fc * dy * dy;
        acc[12] += fc * dy * dz;
        acc[13] += fc * dz * dz;
    }
}

template <int MODE>
__global__ void __launch_bounds__(CQ_THREADS, 1) cq_force_kernel(CqArgs a) {
    extern __shared__ __align__(16) unsigned char cq_smem[];
    double* stage = reinterpret_cast<double*>(cq_smem);  // (x, y, z, q) per slot
    __shared__ __align__(8) unsigned long long full;
    __shared__ CqClass classes[CQ_MAX_KINDS * CQ_MAX_KINDS];
    __shared__ int next_unit;

    if (a.flags[FLAG_NONFINITE] != 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < CQ_MAX_KINDS * CQ_MAX_KINDS) classes[tid] = a.classes[tid];
    if (tid < 8) stage[tid] = (tid & 3) == 3 ? 0.0 : 1.0e9 * (double)((tid & 3) + 1);  // dummy slots 0, 1: far away, no charge
    double acc[CQ_NV];
#pragma unroll
    for (int k = 0; k < CQ_NV; k++) acc[k] = 0.0;
    // the columns are the raw ones of the build: 32-bit entries (level << 28 | flags | slot), four per 16-byte word; the
    // two lanes of an atom take the even and the odd words
    const size_t slab_words = (size_t)(a.capacity >> 2) * 32;
    const int local = tid / CQ_LPA, half = tid % CQ_LPA;
    unsigned phase = 0;

    const int nunits = a.flags[FLAG_UNITS];
    for (int unit = blockIdx.x;; ) {
        __syncthreads();  // everybody is done with the previous copy (and, the first time, the barrier is initialised)
        if (unit >= nunits) break;
        const int4 header = a.header[unit];
        const bool staged = header.z >= 0;
        if (tid < 32) {
            if (staged) {
                if (lane == 0) mbar_expect_bytes(&full, 32u * (unsigned)(header.w - 2));
                __syncwarp();
                for (int r = lane; r < header.z; r += 32) {
                    const int4 run = a.runs[(size_t)unit * LJ2_MAX_RUNS + r];
                    bulk_load(stage + 4 * run.z, a.frame + 4 * (size_t)run.x, (unsigned)run.y * 32u, &full);
                }
            } else if (lane == 0) {
                mbar_arrive(&full);
            }
            if (lane == 0) next_unit = atomicAdd(a.unit_counter, 1) + (int)gridDim.x;
        }
        // the thread's atom, while the copies are in flight
        const int s = a.unit_start[unit] + local;
        const bool present = s < a.unit_start[unit + 1];
        int count = 0, origin = -1, tag = 0;
        unsigned self = 0;
        if (present) {
            count = a.ncount[s];
            self = a.self_slot[s];
            origin = a.order[s];
            tag = __float_as_int(a.tags[s].w);
        }
        const bool owned = present && origin >= a.o_lo && origin < a.o_hi;
        const int column = present ? s : 0;
        const uint4* words = reinterpret_cast<const uint4*>(a.nlist) + (size_t)(column >> 5) * slab_words + (size_t)half * 32 + (column & 31);
        double fx = 0.0, fy = 0.0, fz = 0.0;
        mbar_wait(&full, phase);
        phase ^= 1u;
        CqAtom me;
        me.row = classes + (tag & 3) * CQ_MAX_KINDS;
        if (staged) {
            me.x = stage[4 * self];
            me.y = stage[4 * self + 1];
            me.z = stage[4 * self + 2];
            me.q = stage[4 * self + 3] * INV_FOUR_PI_EPSILON_0;
            // this lane's words: half, half + 2, ...
            const int nwords = owned ? (((count + 3) >> 2) - half + 1) >> 1 : 0;
            uint4 wcur = nwords > 0 ? words[0] : make_uint4(0, 0, 0, 0);
            uint4 wnext = nwords > 1 ? words[64] : make_uint4(0, 0, 0, 0);
            for (int w = 0; w < nwords; w++) {
                uint4 wafter = make_uint4(0, 0, 0, 0);
                if (w + 2 < nwords) wafter = words[(size_t)(w + 2) * 64];
                if (w + 10 < nwords) asm volatile("prefetch.global.L2 [%0];" ::"l"(words + (size_t)(w + 10) * 64));
                const unsigned entries[4] = {wcur.x & 0xffffu, wcur.y & 0xffffu, wcur.z & 0xffffu, wcur.w & 0xffffu};
                unsigned near = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2* pj = reinterpret_cast<const double2*>(stage + 4 * (entries[q] & 0x1fffu));
                    const double2 xy = pj[0], zq = pj[1];
                    cq_pair<MODE>(a, me, xy.x, xy.y, zq.x, zq.y, (entries[q] >> 13) & 3u, (entries[q] >> 15) != 0u, fx, fy, fz, acc, near, 1u << q);
                }
                if (near != 0) {
                    // the defer routine recovers the frame index of a neighbour from the unit's runs (entries as slot * 3)
                    const unsigned lo = (3u * (entries[0] & 0x1fffu)) | ((3u * (entries[1] & 0x1fffu)) << 16);
                    const unsigned hi = (3u * (entries[2] & 0x1fffu)) | ((3u * (entries[3] & 0x1fffu)) << 16);
                    lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, origin, near, lo, hi, true);
                }
                wcur = wnext;
                wnext = wafter;
            }
        } else if (owned) {
            const int f_self = a.fidx[s];
            me.x = a.frame[4 * (size_t)f_self];
            me.y = a.frame[4 * (size_t)f_self + 1];
            me.z = a.frame[4 * (size_t)f_self + 2];
            me.q = a.frame[4 * (size_t)f_self + 3] * INV_FOUR_PI_EPSILON_0;
            const int nwords = (((count + 3) >> 2) - half + 1) >> 1;
            for (int w = 0; w < nwords; w++) {
                const uint4 word = words[(size_t)w * 64];
                const unsigned entries[4] = {word.x & RAW_VALUE_MASK, word.y & RAW_VALUE_MASK, word.z & RAW_VALUE_MASK, word.w & RAW_VALUE_MASK};
#pragma unroll
                for (int q = 0; q < 4; q += 2) {
                    unsigned near = 0;
                    for (int t = 0; t < 2; t++) {
                        const unsigned entry = entries[q + t];
                        const double* pj = a.frame + 4 * (size_t)(entry & 0x1ffffffu);
                        cq_pair<MODE>(a, me, pj[0], pj[1], pj[2], pj[3], (entry >> 25) & 3u, ((entry >> 27) & 1u) != 0u, fx, fy, fz, acc, near, 1u << t);
                    }
                    if (near != 0) {
                        lj2_defer(a.header, a.runs, a.deferred, a.deferred_capacity, a.flags, unit, origin, near, entries[q] & 0x1ffffffu,
                                  entries[q + 1] & 0x1ffffffu, false);
                    }
                }
            }
        }
        fx += __shfl_xor_sync(0xffffffffu, fx, 1);
        fy += __shfl_xor_sync(0xffffffffu, fy, 1);
        fz += __shfl_xor_sync(0xffffffffu, fz, 1);
        if (owned && half == 0 && a.write_forces) {
            a.force[3 * (size_t)origin] = fx;
            a.force[3 * (size_t)origin + 1] = fy;
            a.force[3 * (size_t)origin + 2] = fz;
        }
        __syncthreads();
        unit = next_unit;
    }
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < CQ_NV; k++) acc[k] *= 0.5;
        block_sum<CQ_NV>(acc, stage);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < CQ_NV; k++) a.partials[(size_t)blockIdx.x * CQ_NV + k] = acc[k];
        }
    }
}

// Pairs left out by a list kernel because they sit on a cut-off: evaluated like the all-pairs kernel evaluates them (the
// reference's arithmetic and cut-off semantics, any pair potential and restriction), one side of the pair per entry.
struct FixArgs {
    const int2* __restrict__ deferred;  // (state index of i, frame index of j)
    int capacity;
    const int* __restrict__ frame_atom;
    const int* __restrict__ order;
    const double* __restrict__ pos;
    const double* __restrict__ charge;
    const unsigned* __restrict__ kind;
    const int* __restrict__ mol_first;
    const int* __restrict__ bd_row;
    const unsigned char* __restrict__ bond_dist;
    int nkinds;
    const PairParams* __restrict__ pairs;
    const TableDesc* __restrict__ tables;
    const double* __restrict__ table_energy;
    const double* __restrict__ table_force;
    CellView cell;
    CoulombView coulomb;
    int do_pairs, do_coulomb, full, write_forces;
    double* __restrict__ force;
    double* __restrict__ results;
    const int* __restrict__ flags;
};

__global__ void __launch_bounds__(128) list_fixup_kernel(FixArgs a) {
    const int count = min(a.flags[FLAG_DEFERRED], a.capacity);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const int2 entry = a.deferred[k];
        const int i = entry.x;
        const int s_j = a.frame_atom[entry.y];
        if (s_j < 0) continue;
        const int j = a.order[s_j];
        double dx = __dadd_rn(a.pos[3 * (size_t)i], -a.pos[3 * (size_t)j]);
        double dy = __dadd_rn(a.pos[3 * (size_t)i + 1], -a.pos[3 * (size_t)j + 1]);
        double dz = __dadd_rn(a.pos[3 * (size_t)i + 2], -a.pos[3 * (size_t)j + 2]);
        vector_image_exact(a.cell, dx, dy, dz);
        const double r = sqrt(dot3_exact(dx, dy, dz, dx, dy, dz));
        const bool same_molecule = a.mol_first[i] == a.mol_first[j];
        const unsigned bits = same_molecule ? a.bond_dist[a.bd_row[i] + (j - a.mol_first[j])] : 0u;
        double fx = 0.0, fy = 0.0, fz = 0.0;
        if (a.do_pairs) {
            const PairParams& pp = a.pairs[a.kind[i] * a.nkinds + a.kind[j]];
            if (pp.potential > LUMOL_CUDA_POTENTIAL_NULL && r < pp.cutoff) {
                double scaling;
                if (!restriction_excluded(pp.restriction, bits, pp.scale14, scaling)) {
                    double e, f;
                    pair_eval(pp, a.tables, a.table_energy, a.table_force, r, e, f);
                    const double fr = scaling * f / r;
                    fx += fr * dx;
                    fy += fr * dy;
                    fz += fr * dz;
                    if (a.full) {
                        const double w = 0.5 * fr;
                        atomicAdd(a.results + RES_E_PAIRS, 0.5 * scaling * e);
                        atomicAdd(a.results + RES_PAIR_COUNT, 0.5);
                        atomicAdd(a.results + RES_W_PAIRS + 0, w * dx * dx);
                        atomicAdd(a.results + RES_W_PAIRS + 1, w * dx * dy);
                        atomicAdd(a.results + RES_W_PAIRS + 2, w * dx * dz);
                        atomicAdd(a.results + RES_W_PAIRS + 3, w * dy * dy);
                        atomicAdd(a.results + RES_W_PAIRS + 4, w * dy * dz);
                        atomicAdd(a.results + RES_W_PAIRS + 5, w * dz * dz);
                    }
                }
            }
        }
        if (a.do_coulomb && a.coulomb.kind != 0 && r <= a.coulomb.rc) {
            const double qi = a.charge[i], qj = a.charge[j];
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
                    if (a.full) {
                        const double w = 0.5 * fr;
                        atomicAdd(a.results + RES_E_COULOMB_REAL, 0.5 * e);
                        atomicAdd(a.results + RES_COULOMB_PAIR_COUNT, 0.5);
                        atomicAdd(a.results + RES_W_COULOMB_REAL + 0, w * dx * dx);
                        atomicAdd(a.results + RES_W_COULOMB_REAL + 1, w * dx * dy);
                        atomicAdd(a.results + RES_W_COULOMB_REAL + 2, w * dx * dz);
                        atomicAdd(a.results + RES_W_COULOMB_REAL + 3, w * dy * dy);
                        atomicAdd(a.results + RES_W_COULOMB_REAL + 4, w * dy * dz);
                        atomicAdd(a.results + RES_W_COULOMB_REAL + 5, w * dz * dz);
                    }
                }
            }
        }
        if (a.write_forces) {
            atomicAdd(a.force + 3 * (size_t)i, fx);
            atomicAdd(a.force + 3 * (size_t)i + 1, fy);
            atomicAdd(a.force + 3 * (size_t)i + 2, fz);
        }
    }
}

static int band_below(double cutoff) {
    const double squared = cutoff * cutoff;
    uint64_t pattern;
    std::memcpy(&pattern, &squared, sizeof(pattern));
    return (int)(pattern >> 32) - 1;
}

bool cq_applicable(const Context* ctx) {
    static const char* knob = std::getenv("LUMOL_CUDA_CQ");
    if (knob != nullptr && knob[0] == '0') return false;
    if (ctx->coulomb.kind == 0 || ctx->forced_path == 2 || ctx->nkinds < 1 || ctx->nkinds > CQ_MAX_KINDS) return false;
    if (ctx->coulomb.restriction != LUMOL_CUDA_RESTRICTION_NONE && ctx->coulomb.restriction != LUMOL_CUDA_RESTRICTION_INTER_MOLECULAR) return false;
    if (!(ctx->coulomb.alpha * ctx->coulomb.rc <= 3.4)) return false;
    for (const lumol_cuda_pair& p : ctx->host_pairs) {
        if (p.potential == LUMOL_CUDA_POTENTIAL_ABSENT || p.potential == LUMOL_CUDA_POTENTIAL_NULL) continue;
        if (p.potential != LUMOL_CUDA_POTENTIAL_LJ || p.restriction != LUMOL_CUDA_RESTRICTION_NONE) return false;
    }
    return ctx->n < (1 << 25);
}

int launch_pairs_cq(Context* ctx, const ComputeRequest& req) {
    UnitShape shape;
    shape.atoms = CQ_ATOMS;
    shape.lanes_per_atom = CQ_LPA;
    shape.slots = CQ_SLOTS;
    shape.value_factor = 1;
    shape.width = 4;
    shape.flags = 7;
    double cutoff = ctx->any_pair ? ctx->max_pair_cutoff : 0.0;
    if (ctx->coulomb.rc > cutoff) cutoff = ctx->coulomb.rc;
    Lj2Plan P;
    int status = lj2_plan(ctx, false, shape, cutoff, nullptr, P);
    if (status != 0) return status;
    int64_t o_lo, o_hi;
    ctx->owned_range(ctx->n, o_lo, o_hi);
    int* flags = ctx->nl_flags.ptr;
    const bool reuse = ctx->list_valid && P.signature == ctx->list_signature;
    ctx->list_epoch = ctx->list_epoch % 1000000000 + 1;
    const int epoch = ctx->list_epoch;
    {
        ScopedClock clock(ctx, &ctx->clk_neighbor);
        if (!reuse) {
            lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_REBUILD, epoch);
            lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, FLAG_DEFERRED, 0);
            ctx->launches += 2;
        } else if ((status = lj2_launch_update(ctx, P, epoch)) != 0) {
            return status;
        }
        Rebuild2Args::Sorted none{};
        const bool sharded = ctx->nranks > 1;
        if ((status = lj2_launch_rebuild(ctx, P, epoch, ctx->position.ptr, 0, none, nullptr, sharded ? (int)o_lo : 0, sharded ? (int)o_hi : 0)) != 0) {
            return status;
        }
    }
    ctx->list_valid = true;
    ctx->list_signature = P.signature;

    const bool do_pairs = req.pairs && ctx->any_pair;
    const bool do_coulomb = req.coulomb;
    const bool full = req.energy || req.virial;
    CqArgs a;
    a.n = P.n;
    a.capacity = P.capacity;
    a.unit_start = ctx->unit_rows.ptr + 2 * ((size_t)P.nrows + 2) + (size_t)P.row_scan_blocks + 2;
    a.nlist = ctx->nlist.ptr;
    a.ncount = ctx->ncount.ptr;
    a.self_slot = ctx->self_local.ptr;
    a.fidx = ctx->fidx.ptr;
    a.order = ctx->order.ptr;
    a.tags = ctx->sorted_f32.ptr;
    a.header = ctx->blk_header.ptr;
    a.runs = ctx->blk_entries.ptr;
    a.frame = lj2_frame(ctx, P, 0);
    for (int ki = 0; ki < CQ_MAX_KINDS; ki++) {
        for (int kj = 0; kj < CQ_MAX_KINDS; kj++) {
            CqClass& c = a.classes[ki * CQ_MAX_KINDS + kj];
            c.sigma2 = c.eps24 = c.eps4 = c.shift = 0.0;
            c.band_lo = INT_MIN;
            c.pad = 0;
            if (ki >= ctx->nkinds || kj >= ctx->nkinds || !do_pairs) continue;
            const lumol_cuda_pair& p = ctx->host_pairs[(size_t)ki * ctx->nkinds + kj];
            if (p.potential != LUMOL_CUDA_POTENTIAL_LJ) continue;
            c.sigma2 = p.p[0] * p.p[0];
            c.eps24 = 24.0 * p.p[1];
            c.eps4 = 4.0 * p.p[1];
            c.shift = p.shift;
            c.band_lo = band_below(p.cutoff);
        }
    }
    a.coulomb_kind = do_coulomb ? ctx->coulomb.kind : 0;
    a.exclude_same = ctx->coulomb.restriction == LUMOL_CUDA_RESTRICTION_INTER_MOLECULAR ? 1 : 0;
    a.coulomb_band_lo = do_coulomb ? band_below(ctx->coulomb.rc) : INT_MIN;
    a.alpha = ctx->coulomb.alpha;
    {
        const double cap = 0.9999 * CQ_XMAX;
        uint64_t pattern;
        std::memcpy(&pattern, &cap, sizeof(pattern));
        a.top_cap = (int)(pattern >> 32);
    }
    a.gauss = ctx->coulomb.alpha * FRAC_2_SQRT_PI;
    a.wolf_energy = ctx->coulomb.kind == 2 ? ctx->coulomb.wolf_energy_constant : 0.0;
    a.wolf_force = ctx->coulomb.kind == 2 ? ctx->coulomb.wolf_force_constant : 0.0;
    a.o_lo = (int)o_lo;
    a.o_hi = (int)o_hi;
    a.write_forces = req.forces;
    a.force = ctx->force.ptr;
    a.deferred = ctx->deferred.ptr;
    a.deferred_capacity = P.deferred_capacity;
    a.flags = flags;
    a.unit_counter = flags + 12;
    const int grid = std::max(1, (P.n + CQ_ATOMS - 1) / CQ_ATOMS < ctx->sm_count ? (P.n + CQ_ATOMS - 1) / CQ_ATOMS : ctx->sm_count);
    LUMOL_CUDA_CHECK(ctx, ctx->partials.reserve((size_t)grid * CQ_NV));
    a.partials = ctx->partials.ptr;
    const void* kernel = full ? (const void*)cq_force_kernel<1> : (const void*)cq_force_kernel<0>;
    const size_t smem = (size_t)CQ_SLOTS * 32;
    LUMOL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        ScopedClock clock(ctx, &ctx->clk_pair);
        lj2_set_flag_kernel<<<1, 1, 0, ctx->stream>>>(flags, 12, 0);
        void* params[] = {&a};
        LUMOL_CUDA_CHECK(ctx, cudaLaunchKernel(kernel, dim3(grid), dim3(CQ_THREADS), params, smem, ctx->stream));
        ctx->launches += 2;
        ctx->clk_pair.launches++;
    }
    if (full) {
        if ((status = launch_reduce(ctx, grid, CQ_NV, RES_E_PAIRS)) != 0) return status;
    }
    FixArgs f;
    f.deferred = ctx->deferred.ptr;
    f.capacity = P.deferred_capacity;
    f.frame_atom = ctx->frame_atom.ptr;
    f.order = ctx->order.ptr;
    f.pos = ctx->position.ptr;
    f.charge = ctx->charge.ptr;
    f.kind = ctx->kind.ptr;
    f.mol_first = ctx->mol_first.ptr;
    f.bd_row = ctx->bd_row.ptr;
    f.bond_dist = ctx->bond_dist.ptr;
    f.nkinds = ctx->nkinds;
    f.pairs = ctx->pairs.ptr;
    f.tables = ctx->tables.ptr;
    f.table_energy = ctx->table_energy.ptr;
    f.table_force = ctx->table_force.ptr;
    f.cell = ctx->cell;
    f.coulomb = ctx->coulomb;
    f.do_pairs = do_pairs ? 1 : 0;
    f.do_coulomb = do_coulomb ? 1 : 0;
    f.full = full ? 1 : 0;
    f.write_forces = req.forces ? 1 : 0;
    f.force = ctx->force.ptr;
    f.results = ctx->results.ptr;
    f.flags = flags;
    list_fixup_kernel<<<8, 128, 0, ctx->stream>>>(f);
    ctx->launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

}  // namespace lumol
