// Pieces of the cell-list rebuild shared by the neighbour-list paths (pairs_cells.cu: general + first-generation
// Lennard-Jones kernels; pairs_lj2.cu: pipelined Lennard-Jones kernel): flags, grid arithmetic, the phases of the
// counting sort.  The reference has no neighbour search; binning wraps like UnitCell::wrap_vector
// (cells.rs:263-279).
#pragma once

#include "context.hpp"

namespace lumol {

// flags[0]: rebuild requested, flags[1]: a list column overflowed, flags[2]: number of rebuilds so far
constexpr int FLAG_REBUILD = 0, FLAG_OVERFLOW = 1, FLAG_COUNT = 2;
// flags[3]: blocks that could not be staged (FLAG_UNSTAGED); flags[4]: a position was not finite at the last rebuild
constexpr int FLAG_UNSTAGED = 3;
constexpr int FLAG_NONFINITE = 4;

// ------------------------------------------------------------------------------------------------
// counting sort (every kernel returns immediately when no rebuild is requested)
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

// The rebuild runs as the phases of ONE cooperative kernel (rebuild_kernel below) separated by grid-wide
// barriers; `vb` is the virtual block a resident block is working on.
constexpr int REBUILD_THREADS = 256;
constexpr int REBUILD_WARPS = REBUILD_THREADS / 32;

__device__ __forceinline__ void cell_zero_phase(int count, int* __restrict__ cell_count, unsigned char* __restrict__ cell_needed,
                                                int* __restrict__ flags) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        flags[3] = 0;  // FLAG_UNSTAGED, counted by block_table_phase
        flags[FLAG_NONFINITE] = 0;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        cell_count[k] = 0;
        cell_needed[k] = 0;
    }
}

__device__ __forceinline__ void cell_assign_phase(int vb, int n, const GridView& g, const double* __restrict__ pos,
                                                  int* __restrict__ cell_of, int* __restrict__ slot_of,
                                                  int* __restrict__ cell_count, int* __restrict__ flags) {
    const int i = vb * REBUILD_THREADS + threadIdx.x;
    if (i >= n) return;
    // an exploded simulation must not turn the sort into a quadratic loop over one cell: the rebuild stops here
    if (!(isfinite(pos[3 * i]) && isfinite(pos[3 * i + 1]) && isfinite(pos[3 * i + 2]))) flags[FLAG_NONFINITE] = 1;
    const int cx = cell_coordinate(wrap_coordinate(pos[3 * i], g.length[0]), g.length[0], g.nc[0]);
    const int cy = cell_coordinate(wrap_coordinate(pos[3 * i + 1], g.length[1]), g.length[1], g.nc[1]);
    const int cz = cell_coordinate(wrap_coordinate(pos[3 * i + 2], g.length[2]), g.length[2], g.nc[2]);
    const int c = (cz * g.nc[1] + cy) * g.nc[0] + cx;
    cell_of[i] = c;
    slot_of[i] = atomicAdd(cell_count + c, 1);
}

// exclusive scan of `count` ints in three passes (block scan, scan of block sums, add back)
constexpr int SCAN_THREADS = REBUILD_THREADS;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int value, int* shared, int& total) {
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

__device__ __forceinline__ void scan_blocks_phase(int vb, int count, const int* __restrict__ in, int* __restrict__ out,
                                                  int* __restrict__ block_sums, int* shared) {
    const int base = vb * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
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
    if (threadIdx.x == 0) block_sums[vb] = total;
}

// one block
__device__ __forceinline__ void scan_sums_phase(int nblocks, int* __restrict__ block_sums, int* shared) {
    int& carry = shared[32];
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

__device__ __forceinline__ void scan_add_phase(int vb, int count, int* __restrict__ out, const int* __restrict__ block_sums,
                                               int total_count) {
    const int base = vb * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    const int add = block_sums[vb];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < count) out[base + k] += add;
    }
    if (vb == 0 && threadIdx.x == 0) out[count] = total_count;
}

// first pass of the scatter: original indices grouped by cell, arrival order
__device__ __forceinline__ void cell_group_phase(int vb, int n, const int* __restrict__ cell_of, const int* __restrict__ slot_of,
                                                 const int* __restrict__ cell_start, int* __restrict__ grouped) {
    const int i = vb * REBUILD_THREADS + threadIdx.x;
    if (i >= n) return;
    grouped[cell_start[cell_of[i]] + slot_of[i]] = i;
}

}  // namespace lumol
