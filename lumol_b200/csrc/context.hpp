// Host-side state of one lumol_cuda_context: device buffers, interaction tables, launch bookkeeping.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "device_math.cuh"

namespace lumol {

// Growable device buffer.
template <typename T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t capacity = 0;
    bool zero_filled = false;  // the current allocation was cleared by reserve_zeroed

    cudaError_t reserve(size_t count) {
        if (count <= capacity) {
            return cudaSuccess;
        }
        if (ptr != nullptr) {
            cudaFree(ptr);
            ptr = nullptr;
            capacity = 0;
        }
        size_t want = count + count / 8 + 32;
        zero_filled = false;
        cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&ptr), want * sizeof(T));
        if (err == cudaSuccess) {
            capacity = want;
        }
        return err;
    }

    // like reserve, and a buffer that had to be (re)allocated starts as zeros: for tables whose tails are read ahead
    // and discarded (list words behind the walked prefix, run slots behind the run count)
    cudaError_t reserve_zeroed(size_t count, cudaStream_t stream) {
        if (count <= capacity && zero_filled) {
            return cudaSuccess;
        }
        cudaError_t err = reserve(count);
        if (err == cudaSuccess) {
            // (also a buffer that another path allocated earlier with plain reserve())
            err = cudaMemsetAsync(ptr, 0, capacity * sizeof(T), stream);
            zero_filled = err == cudaSuccess;
        }
        return err;
    }

    void release() {
        if (ptr != nullptr) {
            cudaFree(ptr);
        }
        ptr = nullptr;
        capacity = 0;
    }
};

// Result slots written by the reduction kernels (device `results` array).
enum ResultSlot {
    // ---- one 16-value block written by the pair kernels -------------------------------------------
    RES_E_PAIRS = 0,
    RES_E_COULOMB_REAL = 1,
    RES_W_PAIRS = 2,          // 6 values: xx xy xz yy yz zz
    RES_W_COULOMB_REAL = 8,   // 6
    RES_PAIR_COUNT = 14,          // pair interactions evaluated inside their cut-off (each pair once)
    RES_COULOMB_PAIR_COUNT = 15,  // same for the coulomb real-space term
    // ---- bonded -----------------------------------------------------------------------------------
    RES_E_BONDS = 16,
    RES_W_BONDS = 17,         // 6, directly after RES_E_BONDS (one 7-value reduction)
    RES_E_ANGLES = 23,
    RES_E_DIHEDRALS = 24,
    // ---- everything below is NOT a per-rank partial sum ---------------------------------------------
    RES_E_KSPACE = 25,
    RES_W_KSPACE = 26,        // 6
    RES_CHARGE2 = 32,         // sum q^2 over the atoms of this rank
    RES_KINETIC = 33,
    RES_KINETIC_TENSOR = 34,  // 6
    RES_MOMENTUM = 40,        // sum m v (3) and sum m (1)
    RES_MOLECULAR_BLOCK = 44,      // same 16-value layout as slots 0..15, from the molecular-virial kernel
    RES_W_MOLECULAR_PAIRS = 46,    // 6
    RES_W_MOLECULAR_COULOMB = 52,  // 6
    RES_W_KSPACE_CORRECTION = 60,  // 9: sum_mol sum_i f_i (x) (x_i - com), ewald.rs:736-753
    RES_SCALE_FACTOR = 69,         // thermostat factor computed on the device
    RES_FLAGS = 70,
    RES_CENTER = 72,      // sum m x (3) and sum m (1): RemoveRotation
    RES_ROTATION = 76,    // angular momentum (3), -sum m d (x) d (6: xx xy xz yy yz zz)
    RES_COUNT = 88
};

struct Timer {
    cudaEvent_t start = nullptr, stop = nullptr;
};

struct KernelClock {
    int64_t launches = 0;
    double ms = 0.0;
};

struct Comm;  // comm.cu

// Peer-memory position exchange (comm.cu): where the drift kernel of a sharded run stores the new positions of its
// atoms on the other GPUs, and how their arrival is signalled.
constexpr int PEER_MAX_RANKS = 8;
struct PeerPush {
    int nranks = 0, rank = 0;
    int epoch = 0;                        // push number; its parity selects the inbox copy
    double* inbox[PEER_MAX_RANKS] = {};   // rank p's inbox copy of this parity (3 n doubles), mapped here
    int* flags[PEER_MAX_RANKS] = {};      // rank p's arrival flags of this parity: flags[source rank] = epoch
    int* counter = nullptr;               // blocks of the pushing kernel that are done
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    int sm_count = 148;

    // ---- particles -------------------------------------------------------------------------
    int64_t n = 0;
    DeviceBuffer<double> position, velocity, force, mass, charge;
    DeviceBuffer<double> aux;  // Verlet: previous positions; LeapFrog: accelerations
    DeviceBuffer<unsigned> kind;
    DeviceBuffer<int> mol_first;  // first atom of the molecule holding atom i
    DeviceBuffer<int> bd_row;     // byte offset of atom i's row in bond_dist
    DeviceBuffer<int> mol_of;     // molecule index of atom i
    DeviceBuffer<unsigned char> bond_dist;
    DeviceBuffer<int> mol_start;  // nmol + 1
    DeviceBuffer<double> mol_com; // nmol x 3
    int64_t nmol = 0;
    int max_mol_size = 1;
    bool has_molecules = false;

    // ---- cell ------------------------------------------------------------------------------
    CellView cell{};
    bool cell_set = false;
    uint64_t cell_generation = 0;

    // ---- interactions ----------------------------------------------------------------------
    int nkinds = 0;
    std::vector<lumol_cuda_pair> host_pairs;
    DeviceBuffer<PairParams> pairs;
    double max_pair_cutoff = 0.0;
    bool any_pair = false;
    bool single_lj = false;  // every present pair is plain LJ with restriction None: fast path
    bool simple_pairs = false;  // every entry is absent, null, LJ or harmonic: the list kernel drops the other closed forms
    std::vector<TableDesc> host_tables;
    std::vector<double> host_table_energy, host_table_force;
    DeviceBuffer<TableDesc> tables;
    DeviceBuffer<double> table_energy, table_force;
    bool tables_dirty = false;

    std::vector<lumol_cuda_potential> host_bonded;
    DeviceBuffer<lumol_cuda_potential> bonded;
    int64_t nbonds = 0, nangles = 0, ndihedrals = 0;
    DeviceBuffer<int> bonds, angles, dihedrals;  // index tuples followed by the potential id
    CoulombView coulomb{};
    int kmax = 0;

    // ---- Ewald k-space -----------------------------------------------------------------------
    uint64_t ewald_generation = ~0ull;  // cell generation the factor table was built for
    int64_t nk = 0;
    double kmax2 = 0.0;
    std::vector<int> host_kindex;
    std::vector<double> host_kenergy;
    DeviceBuffer<short4> kindex;
    DeviceBuffer<double> kenergy;  // EwaldFactor::energy
    DeviceBuffer<double> kvirial;  // 6 per k
    DeviceBuffer<double2> rho, rho_partial;
    DeviceBuffer<int4> krows;        // (h, k) rows of the k list (KRow, ewald.cu)
    int64_t nkrows = 0;
    bool krows_regular = false;
    int kspace_algorithm = -1;       // -1 automatic, 0 direct kernels, 1 tiled kernels
    DeviceBuffer<double> kgmat, kforce_partial;
    DeviceBuffer<double2> kxy_scratch;  // e_x | e_y phase tables of the resident blocks of the tiled force kernel (large kmax)
    double kbasis[9] = {0};  // k_vector of the three unit indices, one per row
    uint64_t ewald_table_version = 0;  // bumped whenever ewald_prepare rebuilds the factor table
    // rho(k) on the device describes the resident positions iff these match positions_epoch / ewald_table_version
    uint64_t rho_positions_epoch = ~0ull, rho_table_version = ~0ull;

    // ---- Monte Carlo trial moves (mc.cu) ---------------------------------------------------------
    uint64_t positions_epoch = 0;          // bumped by every call that changes the resident positions or charges
    std::vector<int> host_mol_start;       // nmol + 1 (empty while every atom is its own molecule)
    DeviceBuffer<int2> mc_trials;          // (molecule, first row of its new positions) per trial
    DeviceBuffer<double> mc_new_pos;       // rows of 3: new positions of the trials, concatenated
    DeviceBuffer<double2> mc_delta_rho;    // ntrials x nk
    DeviceBuffer<double> mc_pair_partials, mc_k_partials, mc_results;
    std::vector<int2> mc_host_trials;      // of the last cost call
    uint64_t mc_positions_epoch = ~0ull;   // resident state the last cost call was evaluated against
    double* mc_host_results = nullptr;     // pinned
    size_t mc_host_capacity = 0;

    // ---- neighbour search ----------------------------------------------------------------------
    int forced_path = -1;
    int path = 0;
    int ncell[3] = {0, 0, 0};
    double skin = 1.0;             // Verlet skin in Angstrom; rebuild when an atom moved more than skin / 2
    double skin_effective = 0.0;   // what fits in the box
    bool list_valid = false;
    bool flags_initialised = false;
    int list_epoch = 0;     // evaluation counter; flags[FLAG_REBUILD] == epoch requests a rebuild
    int rebuild_grid = 0;   // co-resident blocks of the cooperative rebuild kernel
    uint64_t list_signature = 0;
    uint64_t structure_generation = 0;  // bumped by every change the list depends on besides positions
    DeviceBuffer<int> nl_flags;         // rebuild flag, overflow flag, rebuild counter
    DeviceBuffer<unsigned> nlist;       // neighbour list: one slab per 32 atoms, columns interleaved by 16-byte words
    DeviceBuffer<int> ncount;
    DeviceBuffer<double> xref;          // positions at the last rebuild, original order
    DeviceBuffer<double4> rel0;         // cell-relative positions at the last rebuild, sorted order
    DeviceBuffer<int> cell_of, cell_count, cell_start, order;
    DeviceBuffer<double4> sorted_pos;  // x, y, z relative to the centre of the own cell, charge
    DeviceBuffer<float4> sorted_f32;   // the same position in FP32
    DeviceBuffer<int4> sorted_info;    // kind, mol_first, bd_row, original index
    DeviceBuffer<int> scan_scratch;
    DeviceBuffer<unsigned char> cell_needed;   // sharded runs: cells this rank's lists refer to
    DeviceBuffer<int> sorted_cell;             // linear cell index, sorted order
    DeviceBuffer<int2> blk_runs;               // runs of consecutive cells: one bulk copy each
    DeviceBuffer<int4> blk_header, blk_entries;  // staging tables of the Lennard-Jones kernel (pairs_cells.cu)
    DeviceBuffer<double> frame_pos;            // x | y | z planes, sorted order, positions in the frame of the box
    DeviceBuffer<unsigned short> self_local;   // staged slot of each atom inside its own block
    // pipelined Lennard-Jones path (pairs_lj2.cu): frames with ghost cells, radial levels, deferred cut-off pairs
    DeviceBuffer<int> ext_start, fidx, kshift, frame_atom, unit_rows;
    DeviceBuffer<unsigned short> cum_levels;
    DeviceBuffer<int2> deferred;
    int deferred_capacity = 0;
    int rebuild2_grid = 0, rebuild2_grid_flags = 0;
    bool lj2_active = false;  // the current neighbour list was built by pairs_lj2.cu
    // sorted-resident molecular dynamics (pairs_lj2.cu): state in cell order, ownership by units, halo frames pushed to peers
    DeviceBuffer<double> sre_x, sre_v, sre_f, sre_m, sre_tmp;
    DeviceBuffer<int> sre_origin, sre_need_mask, sre_sync;
    DeviceBuffer<int4> sre_halo_items;

    // ---- reductions ------------------------------------------------------------------------------
    DeviceBuffer<double> partials, reduce_scratch;
    DeviceBuffer<double> results;
    double* host_results = nullptr;  // pinned, RES_COUNT doubles

    // ---- molecular dynamics --------------------------------------------------------------------------
    int integrator = -1;
    double dt = 0.0;
    int thermostat = 0;
    double thermostat_temperature = 0.0, thermostat_parameter = 0.0;
    std::vector<double> csvr_noise;
    DeviceBuffer<double> csvr_noise_dev;
    int64_t csvr_cursor = 0, csvr_count = 0;
    unsigned controls = 0;
    // Berendsen barostats (integrators.rs:176-342): target pressure / stress, time scale, current scaling matrix
    double barostat_target[9] = {0}, barostat_tau = 0.0, barostat_eta[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int dof_mode = 0;
    int64_t dof_frozen = 0;
    int64_t md_step = 0;

    // ---- multi-GPU -------------------------------------------------------------------------------------
    Comm* comm = nullptr;
    int rank = 0, nranks = 1;

    // ---- measurement -----------------------------------------------------------------------------------
    bool profiling = false;
    bool discard_downloads = false;  // child rank > 0 of a multi-device context: get_* calls run their collectives, copy nothing
    Timer timer;
    int64_t launches = 0;
    KernelClock clk_pair, clk_kspace, clk_integrate, clk_neighbor, clk_comm;
    double last_pair_count = 0.0, last_coulomb_pair_count = 0.0;

    int fail(int code, const char* fmt, ...) {
        char buffer[512];
        va_list args;
        va_start(args, fmt);
        vsnprintf(buffer, sizeof(buffer), fmt, args);
        va_end(args);
        error = buffer;
        return code;
    }

    // atoms [lo, hi) owned by this rank (contiguous block; in cell-sorted order on the cell path)
    void owned_range(int64_t total, int64_t& lo, int64_t& hi) const {
        int64_t chunk = (total + nranks - 1) / nranks;
        lo = chunk * rank;
        hi = lo + chunk;
        if (lo > total) lo = total;
        if (hi > total) hi = total;
    }
};

#define LUMOL_CUDA_CHECK(ctx, expr)                                                                        \
    do {                                                                                                   \
        cudaError_t err__ = (expr);                                                                        \
        if (err__ != cudaSuccess) {                                                                        \
            return (ctx)->fail(LUMOL_CUDA_ERROR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                               __FILE__, __LINE__);                                                        \
        }                                                                                                  \
    } while (0)

// Scoped event timing of one kernel class on the context stream (only when profiling is enabled).
struct ScopedClock {
    Context* ctx;
    KernelClock* clock;
    ScopedClock(Context* c, KernelClock* k) : ctx(c), clock(k) {
        if (ctx->profiling) {
            cudaEventRecord(ctx->timer.start, ctx->stream);
        }
    }
    ~ScopedClock() {
        if (ctx->profiling) {
            cudaEventRecord(ctx->timer.stop, ctx->stream);
            cudaEventSynchronize(ctx->timer.stop);
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, ctx->timer.start, ctx->timer.stop);
            clock->ms += ms;
        }
    }
};

// ---- launchers implemented in the .cu files ---------------------------------------------------------
struct ComputeRequest {
    bool forces = false;
    bool energy = false;
    bool virial = false;            // atomic
    bool molecular_virial = false;
    bool pairs = false;
    bool bonded = false;
    bool coulomb = false;
};

int launch_reduce(Context* ctx, int nblocks, int nvalues, int first_slot);                       // reduce.cu
int launch_pairs_allpairs(Context* ctx, const ComputeRequest& req);                              // pairs_allpairs.cu
int launch_pairs_cells(Context* ctx, const ComputeRequest& req);                                 // pairs_cells.cu
int launch_pairs_lj2(Context* ctx, const ComputeRequest& req);                                   // pairs_lj2.cu
bool lj2_enabled(const Context* ctx);                                                            // pairs_lj2.cu
int launch_pairs_cq(Context* ctx, const ComputeRequest& req);                                    // pairs_lj2.cu
bool cq_applicable(const Context* ctx);                                                          // pairs_lj2.cu
int choose_neighbor_path(Context* ctx, double cutoff);                                           // pairs_cells.cu
int neighbor_list_status(Context* ctx, int* rebuilds, int* overflow);                            // pairs_cells.cu
int launch_coulomb_self(Context* ctx);                                                           // pairs_allpairs.cu
int launch_molecule_com(Context* ctx);                                                           // pairs_allpairs.cu
int launch_bonded(Context* ctx, const ComputeRequest& req);                                      // bonded.cu
int ewald_prepare(Context* ctx);                                                                 // ewald.cu
int launch_ewald_kspace(Context* ctx, const ComputeRequest& req);                                // ewald.cu
int launch_kinetic(Context* ctx, bool tensor);                                                   // integrate.cu
int md_setup(Context* ctx);                                                                      // integrate.cu
int barostat_step(Context* ctx);                                                                // api.cu
int md_step(Context* ctx, bool first, bool last);                                                                     // integrate.cu
int launch_scale_velocities(Context* ctx, double factor, bool from_device);                      // integrate.cu
int launch_remove_rotation(Context* ctx);                                                        // integrate.cu
int launch_rewrap(Context* ctx);                                                                 // integrate.cu
int launch_barostat_drift(Context* ctx, const double eta[9], bool isotropic);                    // integrate.cu
int launch_second_kick(Context* ctx);                                                            // integrate.cu
int launch_remove_translation(Context* ctx);                                                     // integrate.cu
int evaluate_forces_device(Context* ctx, const ComputeRequest& req);                             // api.cu
int comm_peer_push_begin(Context* ctx, PeerPush* push);                                          // comm.cu
int comm_peer_gather(Context* ctx, const PeerPush& push);                                        // comm.cu
int comm_allgather_positions(Context* ctx);                                                      // comm.cu
int comm_allreduce(Context* ctx, double* data, int64_t count);                                   // comm.cu
int comm_allgather_blocks(Context* ctx, double* data, int64_t total);                            // comm.cu
int comm_allgather_chunks(Context* ctx, double* data, size_t chunk);                             // comm.cu
int comm_map_peer_buffers(Context* ctx, int count, void* const* local, void** peers, bool* ok);  // comm.cu
int comm_barrier(Context* ctx);                                                                  // comm.cu
bool sorted_md_applicable(Context* ctx);                                                         // pairs_lj2.cu
int sorted_md_run(Context* ctx, int64_t nsteps);                                                 // pairs_lj2.cu
void comm_destroy(Context* ctx);                                                                 // comm.cu
int launch_move_cost(Context* ctx, int ntrials, int max_size);                                   // mc.cu
int launch_move_accept(Context* ctx, int trial, int first, int size, int64_t row);               // mc.cu
int measure_fp64_peak(Context* ctx, double* tflops);                                             // peaks.cu
int measure_copy_bandwidth(Context* ctx, double* gbs);                                           // peaks.cu

}  // namespace lumol

// The opaque handle of include/lumol_cuda.h.
namespace lumol {
struct Multi;
}

struct lumol_cuda_context {
    lumol::Context impl;
    lumol::Multi* multi = nullptr;  // lumol_cuda_create_multi: this context only fans calls out to one child per device
};

namespace lumol {
// multi.cu
typedef std::function<int32_t(lumol_cuda_context*, int)> MultiTask;
int32_t multi_run(lumol_cuda_context* parent, const MultiTask& task);  // on every child at once; first failure, else child 0's status
int multi_size(const lumol_cuda_context* parent);
lumol_cuda_context* multi_child(const lumol_cuda_context* parent, int rank);
void multi_destroy(lumol_cuda_context* parent);
void set_create_error(const char* message);  // api.cu: the text lumol_cuda_last_error(NULL) returns
}  // namespace lumol
