// extern "C" entry points of include/lumol_cuda.h: argument checking, host <-> device staging and the
// orchestration of one fused force / energy / virial evaluation.
#include "context.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace lumol;

static std::string g_create_error;

void lumol::set_create_error(const char* message) { g_create_error = message; }

// A context made by lumol_cuda_create_multi hands the call to one host thread per device (multi.cu); `child` and `rank`
// name the per-device context inside `call`.
#define MULTI_FAN(ctx, call)                                                                                      \
    if ((ctx) != nullptr && (ctx)->multi != nullptr) {                                                            \
        return lumol::multi_run((ctx), [=](lumol_cuda_context* child, int rank) -> int32_t { (void)rank; return (call); }); \
    }

#define CTX_OR_FAIL(ctx)                                  \
    if ((ctx) == nullptr) {                               \
        return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;         \
    }                                                     \
    Context* c = &(ctx)->impl;                            \
    if (cudaSetDevice(c->device) != cudaSuccess) {        \
        return c->fail(LUMOL_CUDA_ERROR_CUDA, "cudaSetDevice(%d) failed", c->device); \
    }

// Brackets every call that changes the resident positions (or charges): whatever was derived from them and is
// kept between calls -- the structure factor the Monte Carlo cost calls reuse, a pending trial move -- is stale.
struct PositionsChange {
    Context* c;
    explicit PositionsChange(Context* context) : c(context) { c->positions_epoch++; }
    ~PositionsChange() { c->positions_epoch++; }
};

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------

extern "C" int32_t lumol_cuda_abi_version(void) { return LUMOL_CUDA_ABI_VERSION; }

extern "C" int32_t lumol_cuda_create(int32_t device, lumol_cuda_context** out) {
    if (out == nullptr) {
        g_create_error = "lumol_cuda_create: null output pointer";
        return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (") +
                         (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                         "); lumol_cuda has no CPU fallback";
        cudaGetLastError();
        return LUMOL_CUDA_ERROR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        g_create_error = "lumol_cuda_create: device ordinal out of range";
        return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
    }
    err = cudaSetDevice(device);
    if (err != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice failed: ") + cudaGetErrorString(err);
        return LUMOL_CUDA_ERROR_CUDA;
    }
    lumol_cuda_context* ctx = new (std::nothrow) lumol_cuda_context();
    if (ctx == nullptr) {
        g_create_error = "out of host memory";
        return LUMOL_CUDA_ERROR_CUDA;
    }
    Context* c = &ctx->impl;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->sm_count = prop.multiProcessorCount;
    }
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->timer.start) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->timer.stop) == cudaSuccess;
    ok = ok && cudaMallocHost(reinterpret_cast<void**>(&c->host_results), RES_COUNT * sizeof(double)) == cudaSuccess;
    ok = ok && c->results.reserve(RES_COUNT) == cudaSuccess;
    ok = ok && cudaMemset(c->results.ptr, 0, RES_COUNT * sizeof(double)) == cudaSuccess;
    if (!ok) {
        g_create_error = std::string("CUDA resource creation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return LUMOL_CUDA_ERROR_CUDA;
    }
    c->cell.shape = LUMOL_CUDA_CELL_INFINITE;
    *out = ctx;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_destroy(lumol_cuda_context* ctx) {
    if (ctx == nullptr) return LUMOL_CUDA_SUCCESS;
    if (ctx->multi != nullptr) {
        multi_destroy(ctx);
        delete ctx;
        return LUMOL_CUDA_SUCCESS;
    }
    Context* c = &ctx->impl;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    comm_destroy(c);
    c->position.release(); c->velocity.release(); c->force.release(); c->mass.release(); c->charge.release();
    c->aux.release(); c->kind.release(); c->mol_first.release(); c->bd_row.release(); c->mol_of.release();
    c->bond_dist.release(); c->mol_start.release(); c->mol_com.release(); c->pairs.release(); c->tables.release();
    c->table_energy.release(); c->table_force.release(); c->bonded.release(); c->bonds.release();
    c->angles.release(); c->dihedrals.release(); c->kindex.release(); c->kenergy.release(); c->kvirial.release();
    c->rho.release(); c->rho_partial.release(); c->cell_of.release(); c->cell_count.release();
    c->cell_start.release(); c->order.release(); c->sorted_pos.release(); c->sorted_f32.release(); c->sorted_info.release();
    c->krows.release(); c->kgmat.release(); c->kforce_partial.release(); c->kxy_scratch.release();
    c->sorted_cell.release(); c->cell_needed.release(); c->frame_pos.release(); c->blk_runs.release(); c->blk_header.release(); c->blk_entries.release(); c->self_local.release();
    c->nl_flags.release(); c->nlist.release(); c->ncount.release(); c->xref.release(); c->rel0.release();
    c->scan_scratch.release(); c->partials.release(); c->reduce_scratch.release(); c->results.release();
    c->csvr_noise_dev.release();
    c->mc_trials.release(); c->mc_new_pos.release(); c->mc_delta_rho.release(); c->mc_pair_partials.release();
    c->mc_k_partials.release(); c->mc_results.release();
    if (c->mc_host_results) cudaFreeHost(c->mc_host_results);
    if (c->host_results) cudaFreeHost(c->host_results);
    if (c->timer.start) cudaEventDestroy(c->timer.start);
    if (c->timer.stop) cudaEventDestroy(c->timer.stop);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete ctx;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" const char* lumol_cuda_last_error(const lumol_cuda_context* ctx) {
    if (ctx == nullptr) return g_create_error.c_str();
    return ctx->impl.error.c_str();
}

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------

static void invert3(const double m[9], double r[9]) {
    // Matrix3::inverse (matrix.rs:212-227)
    const double det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                       m[2] * (m[3] * m[7] - m[4] * m[6]);
    const double id = 1.0 / det;
    r[0] = (m[4] * m[8] - m[7] * m[5]) * id;
    r[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    r[3] = (m[5] * m[6] - m[3] * m[8]) * id;
    r[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    r[5] = (m[3] * m[2] - m[0] * m[5]) * id;
    r[6] = (m[3] * m[7] - m[6] * m[4]) * id;
    r[7] = (m[6] * m[1] - m[0] * m[7]) * id;
    r[8] = (m[0] * m[4] - m[3] * m[1]) * id;
}

extern "C" int32_t lumol_cuda_set_cell(lumol_cuda_context* ctx, const double cell[9], int32_t shape) {
    MULTI_FAN(ctx, lumol_cuda_set_cell(child, cell, shape));
    CTX_OR_FAIL(ctx);
    if (shape < 0 || shape > 2 || (shape != LUMOL_CUDA_CELL_INFINITE && cell == nullptr)) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_set_cell: bad shape or null matrix");
    }
    CellView v{};
    v.shape = shape;
    if (shape != LUMOL_CUDA_CELL_INFINITE) {
        for (int k = 0; k < 9; k++) v.h[k] = cell[k];
        const double det = v.h[0] * (v.h[4] * v.h[8] - v.h[7] * v.h[5]) - v.h[1] * (v.h[3] * v.h[8] - v.h[5] * v.h[6]) +
                           v.h[2] * (v.h[3] * v.h[7] - v.h[4] * v.h[6]);
        if (!(std::fabs(det) > 1e-30)) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "The matrix is not invertible!");
        }
        invert3(v.h, v.inv);
    }
    const bool changed = !c->cell_set || std::memcmp(&v, &c->cell, sizeof(CellView)) != 0;
    c->cell = v;
    c->cell_set = true;
    if (changed) c->cell_generation++;  // Ewald::prepare recomputes the factors only then (ewald.rs:354-359)
    return LUMOL_CUDA_SUCCESS;
}

static int upload(Context* c, void* dst, const void* src, size_t bytes) {
    LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

static int reset_molecules(Context* c) {
    // one molecule per atom: bond_dist holds a single FAR byte that is never read (different molecules)
    const int64_t n = c->n;
    std::vector<int> first((size_t)n), row((size_t)n, 0), start((size_t)n + 1);
    for (int64_t i = 0; i < n; i++) {
        first[(size_t)i] = (int)i;
        start[(size_t)i] = (int)i;
    }
    start[(size_t)n] = (int)n;
    const unsigned char far = (unsigned char)BOND_FAR;
    LUMOL_CUDA_CHECK(c, c->mol_first.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->mol_of.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->bd_row.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->mol_start.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->bond_dist.reserve(16));
    int status = 0;
    if (n > 0) {
        if ((status = upload(c, c->mol_first.ptr, first.data(), (size_t)n * sizeof(int)))) return status;
        if ((status = upload(c, c->mol_of.ptr, first.data(), (size_t)n * sizeof(int)))) return status;
        if ((status = upload(c, c->bd_row.ptr, row.data(), (size_t)n * sizeof(int)))) return status;
    }
    if ((status = upload(c, c->mol_start.ptr, start.data(), ((size_t)n + 1) * sizeof(int)))) return status;
    if ((status = upload(c, c->bond_dist.ptr, &far, 1))) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    c->nmol = n;
    c->max_mol_size = 1;
    c->has_molecules = false;
    c->host_mol_start.clear();
    c->structure_generation++;
    c->positions_epoch++;  // a pending trial move refers to molecules that no longer exist
    return 0;
}

extern "C" int32_t lumol_cuda_set_particles(lumol_cuda_context* ctx, int64_t n, const double* position,
                                            const double* velocity, const double* mass, const double* charge,
                                            const uint32_t* kind) {
    MULTI_FAN(ctx, lumol_cuda_set_particles(child, n, position, velocity, mass, charge, kind));
    CTX_OR_FAIL(ctx);
    if (n < 0 || n > 400000000 || (n > 0 && (position == nullptr || mass == nullptr || charge == nullptr || kind == nullptr))) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_set_particles: bad size or null array");
    }
    PositionsChange change(c);
    const bool resized = n != c->n;
    c->n = n;
    c->structure_generation++;
    // blocks of a multi-GPU all-gather are padded to equal size: keep room for nranks * chunk atoms
    const size_t padded = (size_t)n + 64 * 3 + 8;
    LUMOL_CUDA_CHECK(c, c->position.reserve(3 * padded));
    LUMOL_CUDA_CHECK(c, c->velocity.reserve(3 * padded));
    LUMOL_CUDA_CHECK(c, c->force.reserve(3 * padded));
    LUMOL_CUDA_CHECK(c, c->mass.reserve(padded));
    LUMOL_CUDA_CHECK(c, c->charge.reserve(padded));
    LUMOL_CUDA_CHECK(c, c->kind.reserve(padded));
    int status = 0;
    if (n > 0) {
        if ((status = upload(c, c->position.ptr, position, (size_t)n * 3 * sizeof(double)))) return status;
        if (velocity != nullptr) {
            if ((status = upload(c, c->velocity.ptr, velocity, (size_t)n * 3 * sizeof(double)))) return status;
        } else {
            LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->velocity.ptr, 0, (size_t)n * 3 * sizeof(double), c->stream));
        }
        if ((status = upload(c, c->mass.ptr, mass, (size_t)n * sizeof(double)))) return status;
        if ((status = upload(c, c->charge.ptr, charge, (size_t)n * sizeof(double)))) return status;
        if ((status = upload(c, c->kind.ptr, kind, (size_t)n * sizeof(uint32_t)))) return status;
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->force.ptr, 0, (size_t)n * 3 * sizeof(double), c->stream));
    }
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (resized || !c->has_molecules) {
        // topology and bonded lists refer to atom indices: a new size invalidates them
        if ((status = reset_molecules(c))) return status;
        if (resized) {
            c->nbonds = c->nangles = c->ndihedrals = 0;
            c->integrator = -1;
        }
    }
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_owned_positions(lumol_cuda_context* ctx, const double* owned_position) {
    if (ctx != nullptr && ctx->multi != nullptr) return ctx->impl.fail(LUMOL_CUDA_ERROR_STATE, "a multi-device context takes whole arrays: lumol_cuda_set_positions");
    CTX_OR_FAIL(ctx);
    int64_t lo, hi;
    c->owned_range(c->n, lo, hi);
    if (owned_position == nullptr && hi > lo) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null positions");
    PositionsChange change(c);
    if (hi > lo) {
        int status = upload(c, c->position.ptr + 3 * lo, owned_position, (size_t)(hi - lo) * 3 * sizeof(double));
        if (status) return status;
    }
    // the blocks of the other ranks arrive over NVLink instead of over every rank's PCIe link
    if (c->nranks > 1) {
        int status = comm_allgather_positions(c);
        if (status) return status;
    }
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_positions(lumol_cuda_context* ctx, const double* position) {
    if (ctx != nullptr && ctx->multi != nullptr) {
        // every device uploads the block of atoms it owns; the devices exchange the blocks among themselves
        if (position == nullptr) return ctx->impl.fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null positions");
        return multi_run(ctx, [=](lumol_cuda_context* child, int rank) -> int32_t {
            (void)rank;
            int64_t lo, hi;
            child->impl.owned_range(child->impl.n, lo, hi);
            return lumol_cuda_set_owned_positions(child, position + 3 * lo);
        });
    }
    CTX_OR_FAIL(ctx);
    if (position == nullptr && c->n > 0) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null positions");
    PositionsChange change(c);
    if (c->n > 0) {
        int status = upload(c, c->position.ptr, position, (size_t)c->n * 3 * sizeof(double));
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_velocities(lumol_cuda_context* ctx, const double* velocity) {
    MULTI_FAN(ctx, lumol_cuda_set_velocities(child, velocity));
    CTX_OR_FAIL(ctx);
    if (velocity == nullptr && c->n > 0) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null velocities");
    if (c->n > 0) {
        int status = upload(c, c->velocity.ptr, velocity, (size_t)c->n * 3 * sizeof(double));
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    return LUMOL_CUDA_SUCCESS;
}

static int download3(Context* c, double* dst, DeviceBuffer<double>& src, bool gather_blocks) {
    if (dst == nullptr && c->n > 0) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output array");
    if (c->n == 0) return 0;
    if (gather_blocks && c->nranks > 1) {
        int status = comm_allgather_blocks(c, src.ptr, 3 * c->n);
        if (status) return status;
    }
    if (c->discard_downloads) {  // the first child of a multi-device context fills the caller's array
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(dst, src.ptr, (size_t)c->n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int32_t lumol_cuda_get_positions(lumol_cuda_context* ctx, double* position) {
    MULTI_FAN(ctx, lumol_cuda_get_positions(child, position));
    CTX_OR_FAIL(ctx);
    return download3(c, position, c->position, false);  // positions are all-gathered after every drift
}

extern "C" int32_t lumol_cuda_get_velocities(lumol_cuda_context* ctx, double* velocity) {
    MULTI_FAN(ctx, lumol_cuda_get_velocities(child, velocity));
    CTX_OR_FAIL(ctx);
    return download3(c, velocity, c->velocity, true);
}

extern "C" int32_t lumol_cuda_get_forces(lumol_cuda_context* ctx, double* forces) {
    MULTI_FAN(ctx, lumol_cuda_get_forces(child, forces));
    CTX_OR_FAIL(ctx);
    return download3(c, forces, c->force, true);
}

extern "C" int32_t lumol_cuda_set_molecules(lumol_cuda_context* ctx, int64_t nmol, const uint64_t* start,
                                            const uint64_t* bond_distances_offset, const uint8_t* bond_distances,
                                            uint64_t bond_distances_size) {
    MULTI_FAN(ctx, lumol_cuda_set_molecules(child, nmol, start, bond_distances_offset, bond_distances, bond_distances_size));
    CTX_OR_FAIL(ctx);
    if (nmol == 0) {
        return reset_molecules(c);
    }
    if (nmol < 0 || start == nullptr || bond_distances_offset == nullptr || bond_distances == nullptr) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_set_molecules: null array");
    }
    const int64_t n = c->n;
    if (start[0] != 0 || (int64_t)start[nmol] != n) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "molecule ranges must cover [0, %lld)", (long long)n);
    }
    if (bond_distances_size >= 2000000000ull) {
        return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "bond distance matrices are too large");
    }
    std::vector<int> first((size_t)n), row((size_t)n), of((size_t)n), starts((size_t)nmol + 1);
    int max_size = 1;
    for (int64_t m = 0; m < nmol; m++) {
        const int64_t lo = (int64_t)start[m], hi = (int64_t)start[m + 1];
        if (hi <= lo || hi > n) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "molecule %lld has an empty or out-of-range atom range", (long long)m);
        }
        const int64_t size = hi - lo;
        if (bond_distances_offset[m] + (uint64_t)(size * size) > bond_distances_size) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "bond distance matrix of molecule %lld is out of range", (long long)m);
        }
        if (size > max_size) max_size = (int)size;
        starts[(size_t)m] = (int)lo;
        for (int64_t i = lo; i < hi; i++) {
            first[(size_t)i] = (int)lo;
            of[(size_t)i] = (int)m;
            row[(size_t)i] = (int)(bond_distances_offset[m] + (uint64_t)((i - lo) * size));
        }
    }
    starts[(size_t)nmol] = (int)n;
    LUMOL_CUDA_CHECK(c, c->mol_first.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->mol_of.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->bd_row.reserve((size_t)n + 1));
    LUMOL_CUDA_CHECK(c, c->mol_start.reserve((size_t)nmol + 1));
    LUMOL_CUDA_CHECK(c, c->bond_dist.reserve((size_t)bond_distances_size + 16));
    int status = 0;
    if ((status = upload(c, c->mol_first.ptr, first.data(), (size_t)n * sizeof(int)))) return status;
    if ((status = upload(c, c->mol_of.ptr, of.data(), (size_t)n * sizeof(int)))) return status;
    if ((status = upload(c, c->bd_row.ptr, row.data(), (size_t)n * sizeof(int)))) return status;
    if ((status = upload(c, c->mol_start.ptr, starts.data(), ((size_t)nmol + 1) * sizeof(int)))) return status;
    if ((status = upload(c, c->bond_dist.ptr, bond_distances, (size_t)bond_distances_size))) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    c->nmol = nmol;
    c->max_mol_size = max_size;
    c->has_molecules = true;
    c->host_mol_start = starts;
    c->structure_generation++;
    c->positions_epoch++;
    return LUMOL_CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// interactions
// ------------------------------------------------------------------------------------------------

static bool valid_restriction(int r) { return r >= LUMOL_CUDA_RESTRICTION_NONE && r <= LUMOL_CUDA_RESTRICTION_SCALE14; }

extern "C" int32_t lumol_cuda_set_pairs(lumol_cuda_context* ctx, int32_t nkinds, const lumol_cuda_pair* pairs) {
    MULTI_FAN(ctx, lumol_cuda_set_pairs(child, nkinds, pairs));
    CTX_OR_FAIL(ctx);
    if (nkinds < 0 || (nkinds > 0 && pairs == nullptr)) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_set_pairs: bad table");
    }
    const size_t count = (size_t)nkinds * nkinds;
    std::vector<PairParams> params(count > 0 ? count : 1);
    c->max_pair_cutoff = 0.0;
    c->any_pair = false;
    bool single_lj = count > 0;
    bool simple_pairs = true;
    for (size_t k = 0; k < count; k++) {
        const lumol_cuda_pair& p = pairs[k];
        if (p.potential > LUMOL_CUDA_POTENTIAL_HARMONIC) simple_pairs = false;  // anything with exp, pow or a table
        if (p.potential < LUMOL_CUDA_POTENTIAL_ABSENT || p.potential > LUMOL_CUDA_POTENTIAL_TABLE ||
            p.potential == LUMOL_CUDA_POTENTIAL_COSINE_HARMONIC || p.potential == LUMOL_CUDA_POTENTIAL_TORSION) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "pair entry %zu: potential %d is not a pair potential", k, p.potential);
        }
        if (!valid_restriction(p.restriction)) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "pair entry %zu: unknown restriction %d", k, p.restriction);
        }
        if (p.potential == LUMOL_CUDA_POTENTIAL_TABLE && (p.table < 0 || (size_t)p.table >= c->host_tables.size())) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "pair entry %zu: unknown table %d", k, p.table);
        }
        const lumol_cuda_pair& t = pairs[(k % nkinds) * nkinds + k / nkinds];
        if (std::memcmp(&p, &t, sizeof(lumol_cuda_pair)) != 0) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "the pair table must be symmetric (entry %zu)", k);
        }
        PairParams& q = params[k];
        q.potential = p.potential;
        q.restriction = p.restriction;
        q.table = p.table;
        q.pad = 0;
        for (int a = 0; a < 5; a++) q.p[a] = p.p[a];
        q.cutoff = p.cutoff;
        q.shift = p.shift;
        q.scale14 = p.scale14;
        if (p.potential > LUMOL_CUDA_POTENTIAL_NULL) {
            c->any_pair = true;
            if (p.cutoff > c->max_pair_cutoff) c->max_pair_cutoff = p.cutoff;
        }
        if (p.potential != LUMOL_CUDA_POTENTIAL_LJ || p.restriction != LUMOL_CUDA_RESTRICTION_NONE ||
            std::memcmp(&p, &pairs[0], sizeof(lumol_cuda_pair)) != 0) {
            single_lj = false;
        }
    }
    c->single_lj = single_lj;
    c->simple_pairs = simple_pairs;
    c->nkinds = nkinds;
    c->structure_generation++;
    c->host_pairs.assign(pairs, pairs + count);
    LUMOL_CUDA_CHECK(c, c->pairs.reserve(count + 1));
    if (count > 0) {
        int status = upload(c, c->pairs.ptr, params.data(), count * sizeof(PairParams));
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_add_table(lumol_cuda_context* ctx, int32_t size, double max, const double* energy,
                                        const double* force) {
    MULTI_FAN(ctx, lumol_cuda_add_table(child, size, max, energy, force));
    CTX_OR_FAIL(ctx);
    if (size < 2 || !(max > 0.0) || energy == nullptr || force == nullptr) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_add_table: need size >= 2, max > 0 and both tables");
    }
    TableDesc d;
    d.size = size;
    d.offset = (int)c->host_table_energy.size();
    d.delta = max / (double)size;  // computations.rs:103
    c->host_tables.push_back(d);
    c->host_table_energy.insert(c->host_table_energy.end(), energy, energy + size);
    c->host_table_force.insert(c->host_table_force.end(), force, force + size);
    c->tables_dirty = true;
    return (int32_t)c->host_tables.size() - 1;
}

extern "C" int32_t lumol_cuda_clear_tables(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_clear_tables(child));
    CTX_OR_FAIL(ctx);
    c->host_tables.clear();
    c->host_table_energy.clear();
    c->host_table_force.clear();
    c->tables_dirty = true;
    return LUMOL_CUDA_SUCCESS;
}

static int sync_tables(Context* c) {
    if (!c->tables_dirty) return 0;
    const size_t nt = c->host_tables.size(), nv = c->host_table_energy.size();
    LUMOL_CUDA_CHECK(c, c->tables.reserve(nt + 1));
    LUMOL_CUDA_CHECK(c, c->table_energy.reserve(nv + 1));
    LUMOL_CUDA_CHECK(c, c->table_force.reserve(nv + 1));
    int status = 0;
    if (nt > 0) {
        if ((status = upload(c, c->tables.ptr, c->host_tables.data(), nt * sizeof(TableDesc)))) return status;
        if ((status = upload(c, c->table_energy.ptr, c->host_table_energy.data(), nv * sizeof(double)))) return status;
        if ((status = upload(c, c->table_force.ptr, c->host_table_force.data(), nv * sizeof(double)))) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    c->tables_dirty = false;
    return 0;
}

extern "C" int32_t lumol_cuda_set_bonded_potentials(lumol_cuda_context* ctx, int32_t npotentials,
                                                    const lumol_cuda_potential* potentials) {
    MULTI_FAN(ctx, lumol_cuda_set_bonded_potentials(child, npotentials, potentials));
    CTX_OR_FAIL(ctx);
    if (npotentials < 0 || (npotentials > 0 && potentials == nullptr)) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_set_bonded_potentials: bad table");
    }
    for (int k = 0; k < npotentials; k++) {
        if (potentials[k].potential < LUMOL_CUDA_POTENTIAL_NULL || potentials[k].potential > LUMOL_CUDA_POTENTIAL_TORSION) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "bonded potential %d: unsupported kind %d", k, potentials[k].potential);
        }
    }
    c->host_bonded.assign(potentials, potentials + npotentials);
    LUMOL_CUDA_CHECK(c, c->bonded.reserve((size_t)npotentials + 1));
    if (npotentials > 0) {
        int status = upload(c, c->bonded.ptr, potentials, (size_t)npotentials * sizeof(lumol_cuda_potential));
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    return LUMOL_CUDA_SUCCESS;
}

static int set_terms(Context* c, int arity, int64_t count, const int64_t* atoms, const int32_t* potential,
                     DeviceBuffer<int>& dst, int64_t& stored) {
    if (count < 0 || (count > 0 && (atoms == nullptr || potential == nullptr))) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "bonded list: bad size or null array");
    }
    // every rank keeps the whole list (all positions are resident) and evaluates the terms touching its atoms
    std::vector<int> packed;
    packed.reserve((size_t)count * (arity + 1) + 8);
    int64_t kept = 0;
    for (int64_t t = 0; t < count; t++) {
        for (int a = 0; a < arity; a++) {
            const int64_t idx = atoms[t * arity + a];
            if (idx < 0 || idx >= c->n) {
                return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "bonded list: atom index %lld out of range", (long long)idx);
            }
        }
        if (potential[t] >= (int)c->host_bonded.size()) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "bonded list: unknown potential %d", potential[t]);
        }
        for (int a = 0; a < arity; a++) packed.push_back((int)atoms[t * arity + a]);
        packed.push_back(potential[t]);
        kept++;
    }
    LUMOL_CUDA_CHECK(c, dst.reserve(packed.size() + 1));
    if (!packed.empty()) {
        int status = upload(c, dst.ptr, packed.data(), packed.size() * sizeof(int));
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    stored = kept;
    return 0;
}

extern "C" int32_t lumol_cuda_set_bonds(lumol_cuda_context* ctx, int64_t n, const int64_t* atoms, const int32_t* potential) {
    MULTI_FAN(ctx, lumol_cuda_set_bonds(child, n, atoms, potential));
    CTX_OR_FAIL(ctx);
    return set_terms(c, 2, n, atoms, potential, c->bonds, c->nbonds);
}

extern "C" int32_t lumol_cuda_set_angles(lumol_cuda_context* ctx, int64_t n, const int64_t* atoms, const int32_t* potential) {
    MULTI_FAN(ctx, lumol_cuda_set_angles(child, n, atoms, potential));
    CTX_OR_FAIL(ctx);
    return set_terms(c, 3, n, atoms, potential, c->angles, c->nangles);
}

extern "C" int32_t lumol_cuda_set_dihedrals(lumol_cuda_context* ctx, int64_t n, const int64_t* atoms, const int32_t* potential) {
    MULTI_FAN(ctx, lumol_cuda_set_dihedrals(child, n, atoms, potential));
    CTX_OR_FAIL(ctx);
    return set_terms(c, 4, n, atoms, potential, c->dihedrals, c->ndihedrals);
}

extern "C" int32_t lumol_cuda_set_coulomb_none(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_set_coulomb_none(child));
    CTX_OR_FAIL(ctx);
    if (c->coulomb.kind != 0) c->structure_generation++;
    c->coulomb = CoulombView{};
    c->kmax = 0;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_coulomb_ewald(lumol_cuda_context* ctx, double cutoff, double alpha, int32_t kmax,
                                                int32_t restriction) {
    MULTI_FAN(ctx, lumol_cuda_set_coulomb_ewald(child, cutoff, alpha, kmax, restriction));
    CTX_OR_FAIL(ctx);
    // Ewald::new panics (ewald.rs:280-286)
    if (cutoff < 0.0) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "the cutoff can not be negative in Ewald");
    if (alpha < 0.0) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "alpha can not be negative in Ewald");
    if (kmax <= 0) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "kmax can not be 0 in Ewald");
    if (!valid_restriction(restriction)) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "unknown restriction %d", restriction);
    if (restriction == LUMOL_CUDA_RESTRICTION_SCALE14) {
        return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "Scaling restriction scheme using Ewald are not implemented");  // ewald.rs:388
    }
    CoulombView v{};
    v.kind = 1;
    v.restriction = restriction;
    v.scale14 = 1.0;
    v.rc = cutoff;
    v.alpha = alpha;
    if (c->coulomb.kind != 1 || c->coulomb.alpha != alpha || c->kmax != kmax) {
        c->ewald_generation = ~0ull;  // new factors
    }
    if (c->coulomb.kind != 1 || c->coulomb.rc != cutoff) c->structure_generation++;
    c->coulomb = v;
    c->kmax = kmax;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_coulomb_wolf(lumol_cuda_context* ctx, double cutoff, int32_t restriction, double scale14) {
    MULTI_FAN(ctx, lumol_cuda_set_coulomb_wolf(child, cutoff, restriction, scale14));
    CTX_OR_FAIL(ctx);
    if (!(cutoff > 0.0)) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "Got a negative cutoff in Wolf summation");  // wolf.rs:69
    if (!valid_restriction(restriction)) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "unknown restriction %d", restriction);
    CoulombView v{};
    v.kind = 2;
    v.restriction = restriction;
    v.scale14 = scale14;
    v.rc = cutoff;
    // wolf.rs:70-77
    v.alpha = PI / cutoff;
    const double alpha_cutoff = v.alpha * cutoff;
    const double alpha_cutoff_2 = alpha_cutoff * alpha_cutoff;
    v.wolf_energy_constant = std::erfc(alpha_cutoff) / cutoff;
    v.wolf_force_constant = std::erfc(alpha_cutoff) / (cutoff * cutoff) + FRAC_2_SQRT_PI * v.alpha * std::exp(-alpha_cutoff_2) / cutoff;
    if (c->coulomb.kind != 2 || c->coulomb.rc != cutoff) c->structure_generation++;
    c->coulomb = v;
    c->kmax = 0;
    return LUMOL_CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------------

static double volume_of(const CellView& cell) {
    if (cell.shape == LUMOL_CUDA_CELL_INFINITE) return 0.0;
    if (cell.shape == LUMOL_CUDA_CELL_ORTHORHOMBIC) return cell.h[0] * cell.h[4] * cell.h[8];
    const double* h = cell.h;
    const double a[3] = {h[0], h[3], h[6]}, b[3] = {h[1], h[4], h[7]}, cc[3] = {h[2], h[5], h[8]};
    return a[0] * (b[1] * cc[2] - b[2] * cc[1]) + a[1] * (b[2] * cc[0] - b[0] * cc[2]) + a[2] * (b[0] * cc[1] - b[1] * cc[0]);
}

namespace lumol {

// Everything that runs on the device for one evaluation; leaves forces in ctx->force and the scalar
// sums in ctx->results.  Used by lumol_cuda_compute and by the MD loop.
int evaluate_forces_device(Context* c, const ComputeRequest& req) {
    int status = sync_tables(c);
    if (status) return status;
    const int64_t n = c->n;
    if (n == 0) return 0;

    const bool do_pairs = req.pairs && c->any_pair;
    const bool do_coulomb = req.coulomb && c->coulomb.kind != 0;
    double cutoff = 0.0;
    if (do_pairs) cutoff = c->max_pair_cutoff;
    if (do_coulomb && c->coulomb.rc > cutoff) cutoff = c->coulomb.rc;

    if (do_pairs || do_coulomb) {
        int path = choose_neighbor_path(c, cutoff);
        if (path < 0) {
            return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED,
                           "the cell list needs an orthorhombic cell with at least 3 cut-off lengths per edge");
        }
        if (req.molecular_virial && !req.forces && !req.energy && !req.virial) {
            path = 0;  // the molecular virial lives in the all-pairs kernel
        }
        c->path = path;
        if (path == 1) {
            ComputeRequest r = req;
            r.molecular_virial = false;
            status = launch_pairs_cells(c, r);
            if (status == 0 && req.molecular_virial) {
                ComputeRequest m{};
                m.molecular_virial = true;
                m.pairs = req.pairs;
                m.coulomb = req.coulomb;
                status = launch_pairs_allpairs(c, m);
            }
        } else {
            status = launch_pairs_allpairs(c, req);
        }
        if (status) return status;
    } else {
        if (req.forces) {
            LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->force.ptr, 0, (size_t)n * 3 * sizeof(double), c->stream));
        }
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_E_PAIRS, 0, 16 * sizeof(double), c->stream));
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_MOLECULAR_BLOCK, 0, 16 * sizeof(double), c->stream));
    }

    if (req.bonded) {
        status = launch_bonded(c, req);
        if (status) return status;
    } else {
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_E_BONDS, 0, 9 * sizeof(double), c->stream));
    }

    if (req.energy || req.virial || req.molecular_virial) {  // the host reads no sums after a forces-only evaluation
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_E_KSPACE, 0, 7 * sizeof(double), c->stream));
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_W_KSPACE_CORRECTION, 0, 9 * sizeof(double), c->stream));
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_CHARGE2, 0, sizeof(double), c->stream));
    }
    if (do_coulomb) {
        if (req.energy) {
            status = launch_coulomb_self(c);
            if (status) return status;
        }
        if (c->coulomb.kind == 1) {
            status = launch_ewald_kspace(c, req);
            if (status) return status;
        }
    }
    return 0;
}

}  // namespace lumol

extern "C" int32_t lumol_cuda_compute(lumol_cuda_context* ctx, uint32_t what, uint32_t parts, double* forces,
                                      lumol_cuda_energy* energy, double virial[9]) {
    if (ctx != nullptr && ctx->multi != nullptr) {
        // every device downloads the block of forces it owns straight into the caller's array; sums come from the first
        return multi_run(ctx, [=](lumol_cuda_context* child, int rank) -> int32_t {
            lumol_cuda_energy unused_energy;
            double unused_virial[9];
            int64_t lo = 0, hi = 0;
            child->impl.owned_range(child->impl.n, lo, hi);
            const bool download = forces != nullptr && (what & LUMOL_CUDA_FORCES) != 0 && (what & ~31u) == 0;
            return lumol_cuda_compute(child, download ? (what | LUMOL_CUDA_OWNED_FORCES) : what, parts, download ? forces + 3 * lo : forces,
                                      rank == 0 || energy == nullptr ? energy : &unused_energy,
                                      rank == 0 || virial == nullptr ? virial : unused_virial);
        });
    }
    CTX_OR_FAIL(ctx);
    if ((what & ~31u) != 0 || (parts & ~7u) != 0) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_compute: unknown bits in what/parts");
    }
    ComputeRequest req;
    req.forces = (what & LUMOL_CUDA_FORCES) != 0;
    req.energy = (what & LUMOL_CUDA_ENERGY) != 0;
    req.virial = (what & LUMOL_CUDA_ATOMIC_VIRIAL) != 0;
    req.molecular_virial = !req.virial && (what & LUMOL_CUDA_MOLECULAR_VIRIAL) != 0;
    req.pairs = (parts & LUMOL_CUDA_PART_PAIRS) != 0;
    req.bonded = (parts & LUMOL_CUDA_PART_BONDED) != 0;
    req.coulomb = (parts & LUMOL_CUDA_PART_COULOMB) != 0;
    if ((req.virial || req.molecular_virial) && c->cell.shape == LUMOL_CUDA_CELL_INFINITE) {
        return c->fail(LUMOL_CUDA_ERROR_INFINITE_CELL, "Can not compute virial for infinite cell");  // compute.rs:199
    }
    if (req.coulomb && c->coulomb.kind == 1 && c->cell.shape == LUMOL_CUDA_CELL_INFINITE) {
        return c->fail(LUMOL_CUDA_ERROR_INFINITE_CELL, "Ewald is not defined with infinite unit cell");  // ewald.rs:124
    }
    if (req.molecular_virial && c->nranks > 1) {
        return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "the molecular virial is evaluated on a single GPU only");
    }
    // forces == NULL with LUMOL_CUDA_FORCES is allowed: the forces stay on the device (lumol_cuda_get_forces)

    int status = evaluate_forces_device(c, req);
    if (status) return status;

    // per-rank partial sums -> global sums.  Slots 0..22 (pairs, coulomb real, bonded) and the charge sum
    // are partial per rank; the k-space sums come from the all-reduced rho and are already global.
    if (c->nranks > 1 && (req.energy || req.virial)) {
        status = comm_allreduce(c, c->results.ptr, RES_E_KSPACE);
        if (status) return status;
        status = comm_allreduce(c, c->results.ptr + RES_CHARGE2, 1);
        if (status) return status;
    }

    if (c->n > 0) {
        LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(c->host_results, c->results.ptr, RES_COUNT * sizeof(double),
                                            cudaMemcpyDeviceToHost, c->stream));
    } else {
        std::memset(c->host_results, 0, RES_COUNT * sizeof(double));
    }
    if (req.forces && forces != nullptr && c->n > 0) {
        if ((what & LUMOL_CUDA_OWNED_FORCES) != 0 && c->nranks > 1) {
            // only this rank's block: no all-gather, 1 / nranks of the bytes to the host
            int64_t lo, hi;
            c->owned_range(c->n, lo, hi);
            if (hi > lo) {
                LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(forces, c->force.ptr + 3 * lo, (size_t)(hi - lo) * 3 * sizeof(double),
                                                    cudaMemcpyDeviceToHost, c->stream));
            }
        } else {
            if (c->nranks > 1) {
                status = comm_allgather_blocks(c, c->force.ptr, 3 * c->n);
                if (status) return status;
            }
            LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(forces, c->force.ptr, (size_t)c->n * 3 * sizeof(double),
                                                cudaMemcpyDeviceToHost, c->stream));
        }
    }
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (c->path == 1) {
        int rebuilds = 0, overflow = 0;
        status = neighbor_list_status(c, &rebuilds, &overflow);
        if (status) return status;
        if (overflow == 2) {
            return c->fail(LUMOL_CUDA_ERROR_NOT_FINITE, "a particle position is not finite: the neighbour list cannot be built");
        }
        if (overflow) {
            return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "neighbour list capacity exceeded (strongly inhomogeneous density)");
        }
    }
    const double* r = c->host_results;
    if (req.energy || req.virial) {
        c->last_pair_count = r[RES_PAIR_COUNT];
        c->last_coulomb_pair_count = r[RES_COULOMB_PAIR_COUNT];
    }

    // tail corrections are host scalars: sum over kinds present in the system, both orders
    // (energy.rs:63-79, compute.rs:219-228)
    double tail_energy = 0.0, tail_virial = 0.0;
    if (req.pairs && c->cell.shape != LUMOL_CUDA_CELL_INFINITE && c->nkinds > 0 && (req.energy || req.virial || req.molecular_virial)) {
        std::vector<uint32_t> kinds((size_t)c->n);
        if (c->n > 0) {
            LUMOL_CUDA_CHECK(c, cudaMemcpy(kinds.data(), c->kind.ptr, (size_t)c->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        }
        std::vector<int64_t> counts((size_t)c->nkinds, 0);
        for (uint32_t k : kinds) {
            if ((int)k >= c->nkinds) {
                return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "particle kind %u has no row in the pair table", k);
            }
            counts[k]++;
        }
        const double volume = volume_of(c->cell);
        for (int i = 0; i < c->nkinds; i++) {
            if (counts[(size_t)i] == 0) continue;
            for (int j = 0; j < c->nkinds; j++) {
                if (counts[(size_t)j] == 0) continue;
                const lumol_cuda_pair& p = c->host_pairs[(size_t)i * c->nkinds + j];
                if (p.potential == LUMOL_CUDA_POTENTIAL_ABSENT) continue;
                const double two_pi_density = 2.0 * PI * (double)counts[(size_t)i] * (double)counts[(size_t)j] / volume;
                tail_energy += two_pi_density * p.tail_energy;
                tail_virial += two_pi_density * (p.tail_virial * (1.0 / 3.0));
            }
        }
    }

    if (energy != nullptr) {
        std::memset(energy, 0, sizeof(*energy));
        if (req.energy) {
            energy->pairs = r[RES_E_PAIRS];
            energy->pairs_tail = tail_energy;
            energy->bonds = r[RES_E_BONDS];
            energy->angles = r[RES_E_ANGLES];
            energy->dihedrals = r[RES_E_DIHEDRALS];
            energy->coulomb_real = r[RES_E_COULOMB_REAL];
            if (req.coulomb && c->coulomb.kind == 1) {
                // ewald.rs:619-626
                energy->coulomb_self = -c->coulomb.alpha / std::sqrt(PI) * r[RES_CHARGE2] / FOUR_PI_EPSILON_0;
                energy->coulomb_kspace = r[RES_E_KSPACE];
            } else if (req.coulomb && c->coulomb.kind == 2) {
                // wolf.rs:99-101: sum_i q_i^2 * 0.5 * (energy_constant + alpha * 2/sqrt(pi)) / 4 pi eps0, subtracted
                energy->coulomb_self = -r[RES_CHARGE2] * 0.5 * (c->coulomb.wolf_energy_constant + c->coulomb.alpha * FRAC_2_SQRT_PI) / FOUR_PI_EPSILON_0;
            }
            const double total = energy->pairs + energy->pairs_tail + energy->bonds + energy->angles + energy->dihedrals +
                                 energy->coulomb_real + energy->coulomb_self + energy->coulomb_kspace;
            if (!std::isfinite(total)) {
                return c->fail(LUMOL_CUDA_ERROR_NOT_FINITE, "Potential energy is infinite!");  // compute.rs:125
            }
        }
    }

    if (virial != nullptr) {
        for (int k = 0; k < 9; k++) virial[k] = 0.0;
        static const int map[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
        if (req.virial) {
            for (int k = 0; k < 9; k++) {
                virial[k] = r[RES_W_PAIRS + map[k]] + r[RES_W_BONDS + map[k]] + r[RES_W_COULOMB_REAL + map[k]] +
                            r[RES_W_KSPACE + map[k]];
            }
            virial[0] += tail_virial;
            virial[4] += tail_virial;
            virial[8] += tail_virial;
        } else if (req.molecular_virial) {
            // compute.rs:281-363 (bond virials are ignored there) and ewald.rs:916-923, 736-753
            for (int k = 0; k < 9; k++) {
                virial[k] = r[RES_W_MOLECULAR_PAIRS + map[k]] + r[RES_W_MOLECULAR_COULOMB + map[k]] +
                            r[RES_W_KSPACE + map[k]] - r[RES_W_KSPACE_CORRECTION + k];
            }
            virial[0] += tail_virial;
            virial[4] += tail_virial;
            virial[8] += tail_virial;
        }
    }
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_owned_range(lumol_cuda_context* ctx, int64_t* first, int64_t* count) {
    if (ctx != nullptr && ctx->multi != nullptr) {  // the devices of a multi-device context own everything between them
        if (first == nullptr || count == nullptr) return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
        *first = 0;
        *count = multi_child(ctx, 0)->impl.n;
        return LUMOL_CUDA_SUCCESS;
    }
    CTX_OR_FAIL(ctx);
    if (first == nullptr || count == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output");
    int64_t lo, hi;
    c->owned_range(c->n, lo, hi);
    *first = lo;
    *count = hi - lo;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_kinetic_energy(lumol_cuda_context* ctx, double* kinetic) {
    if (ctx != nullptr && ctx->multi != nullptr) {
        if (kinetic == nullptr) return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
        return multi_run(ctx, [=](lumol_cuda_context* child, int rank) -> int32_t {
            double unused = 0.0;
            return lumol_cuda_kinetic_energy(child, rank == 0 ? kinetic : &unused);
        });
    }
    CTX_OR_FAIL(ctx);
    if (kinetic == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output");
    *kinetic = 0.0;
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    int status = launch_kinetic(c, false);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(c->host_results, c->results.ptr + RES_KINETIC, 11 * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->stream));
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    *kinetic = c->host_results[0];
    if (!std::isfinite(*kinetic)) return c->fail(LUMOL_CUDA_ERROR_NOT_FINITE, "Kinetic energy is infinite!");  // compute.rs:141
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_kinetic_tensor(lumol_cuda_context* ctx, double tensor[9]) {
    if (ctx != nullptr && ctx->multi != nullptr) {
        if (tensor == nullptr) return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
        return multi_run(ctx, [=](lumol_cuda_context* child, int rank) -> int32_t {
            double unused[9];
            return lumol_cuda_kinetic_tensor(child, rank == 0 ? tensor : unused);
        });
    }
    CTX_OR_FAIL(ctx);
    if (tensor == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output");
    for (int k = 0; k < 9; k++) tensor[k] = 0.0;
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    int status = launch_kinetic(c, true);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(c->host_results, c->results.ptr + RES_KINETIC, 11 * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->stream));
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    static const int map[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    for (int k = 0; k < 9; k++) tensor[k] = c->host_results[1 + map[k]];
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_ewald_kvectors(lumol_cuda_context* ctx, int64_t capacity, int64_t* count, int32_t* index,
                                             double* energy_factor, double* rho) {
    MULTI_FAN(ctx, rank != 0 ? 0 : lumol_cuda_ewald_kvectors(child, capacity, count, index, energy_factor, rho));
    CTX_OR_FAIL(ctx);
    if (c->coulomb.kind != 1) return c->fail(LUMOL_CUDA_ERROR_STATE, "Ewald is not the active coulomb potential");
    int status = ewald_prepare(c);
    if (status) return status;
    if (count != nullptr) *count = c->nk;
    if (capacity < c->nk) return LUMOL_CUDA_SUCCESS;
    if (index != nullptr) std::memcpy(index, c->host_kindex.data(), (size_t)c->nk * 3 * sizeof(int32_t));
    if (energy_factor != nullptr) std::memcpy(energy_factor, c->host_kenergy.data(), (size_t)c->nk * sizeof(double));
    if (rho != nullptr && c->nk > 0) {
        LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(rho, c->rho.ptr, (size_t)c->nk * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    return LUMOL_CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// Monte Carlo energy cache (sys/cache.rs)
// ------------------------------------------------------------------------------------------------

extern "C" int32_t lumol_cuda_move_molecules_cost(lumol_cuda_context* ctx, int64_t ntrials, const int64_t* molecules,
                                                  const double* new_positions, lumol_cuda_energy* costs) {
    if (ctx != nullptr && ctx->multi != nullptr) {
        return multi_run(ctx, [=](lumol_cuda_context* child, int rank) -> int32_t {
            std::vector<lumol_cuda_energy> unused((size_t)(rank == 0 || ntrials < 0 || ntrials > 65535 ? 0 : ntrials));
            return lumol_cuda_move_molecules_cost(child, ntrials, molecules, new_positions, rank == 0 || costs == nullptr ? costs : unused.data());
        });
    }
    CTX_OR_FAIL(ctx);
    if (ntrials < 0 || ntrials > 65535 || (ntrials > 0 && (molecules == nullptr || new_positions == nullptr || costs == nullptr))) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_move_molecules_cost: bad trial count or null array");
    }
    if (c->nranks > 1) {
        return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "Monte Carlo trial moves are evaluated on a single GPU only");
    }
    if (c->coulomb.kind == 1 && c->cell.shape == LUMOL_CUDA_CELL_INFINITE) {
        return c->fail(LUMOL_CUDA_ERROR_INFINITE_CELL, "Ewald is not defined with infinite unit cell");  // ewald.rs:124
    }
    c->mc_positions_epoch = ~0ull;
    if (ntrials == 0) return LUMOL_CUDA_SUCCESS;

    // rows of new_positions per trial, from the molecule sizes
    c->mc_host_trials.resize((size_t)ntrials);
    int64_t rows = 0;
    int max_size = 1;
    for (int64_t t = 0; t < ntrials; t++) {
        const int64_t m = molecules[t];
        if (m < 0 || m >= c->nmol) {
            return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "trial %lld: molecule %lld is out of range", (long long)t, (long long)m);
        }
        const int size = c->has_molecules ? c->host_mol_start[(size_t)m + 1] - c->host_mol_start[(size_t)m] : 1;
        if (size > max_size) max_size = size;
        c->mc_host_trials[(size_t)t] = make_int2((int)m, (int)rows);
        rows += size;
        if (rows > 2000000000ll / 3) return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "too many trial positions");
    }
    int status = sync_tables(c);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, c->mc_trials.reserve((size_t)ntrials));
    LUMOL_CUDA_CHECK(c, c->mc_new_pos.reserve((size_t)rows * 3));
    if ((status = upload(c, c->mc_trials.ptr, c->mc_host_trials.data(), (size_t)ntrials * sizeof(int2)))) return status;
    if ((status = upload(c, c->mc_new_pos.ptr, new_positions, (size_t)rows * 3 * sizeof(double)))) return status;

    if (c->coulomb.kind == 1) {
        // the reference reads rho(k) from the cache its last energy / forces call left (ewald.rs:810-819); here it is
        // recomputed only when the resident positions, the cell or the Ewald parameters changed since it was formed
        if ((status = ewald_prepare(c))) return status;
        if (c->rho_positions_epoch != c->positions_epoch || c->rho_table_version != c->ewald_table_version) {
            ComputeRequest rho_only{};
            rho_only.coulomb = true;
            if ((status = launch_ewald_kspace(c, rho_only))) return status;
        }
    }
    if ((status = launch_move_cost(c, (int)ntrials, max_size))) return status;

    if (c->mc_host_capacity < (size_t)ntrials * 6) {
        if (c->mc_host_results) cudaFreeHost(c->mc_host_results);
        c->mc_host_results = nullptr;
        c->mc_host_capacity = 0;
        const size_t want = (size_t)ntrials * 6 + 60;
        LUMOL_CUDA_CHECK(c, cudaMallocHost(reinterpret_cast<void**>(&c->mc_host_results), want * sizeof(double)));
        c->mc_host_capacity = want;
    }
    LUMOL_CUDA_CHECK(c, cudaMemcpyAsync(c->mc_host_results, c->mc_results.ptr, (size_t)ntrials * 6 * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->stream));
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int64_t t = 0; t < ntrials; t++) {
        const double* r = c->mc_host_results + 6 * t;
        std::memset(&costs[t], 0, sizeof(lumol_cuda_energy));
        costs[t].pairs = r[0] - r[1];           // pair tail, bonds, angles, dihedrals do not change (cache.rs:165-167)
        costs[t].coulomb_real = r[2] - r[3];    // ewald.rs:612 / wolf.rs:160; no self cost (ewald.rs:942)
        costs[t].coulomb_kspace = r[4] - r[5];  // ewald.rs:838
    }
    c->mc_positions_epoch = c->positions_epoch;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_move_molecule_cost(lumol_cuda_context* ctx, int64_t molecule, const double* new_positions,
                                                 lumol_cuda_energy* cost) {
    return lumol_cuda_move_molecules_cost(ctx, 1, &molecule, new_positions, cost);
}

extern "C" int32_t lumol_cuda_move_molecule_accept(lumol_cuda_context* ctx, int64_t trial) {
    MULTI_FAN(ctx, lumol_cuda_move_molecule_accept(child, trial));
    CTX_OR_FAIL(ctx);
    if (c->mc_positions_epoch != c->positions_epoch) {
        // cache.rs:123-126
        return c->fail(LUMOL_CUDA_ERROR_STATE, "called EnergyCache::update without call a `*_cost` function first");
    }
    if (trial < 0 || trial >= (int64_t)c->mc_host_trials.size()) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "trial %lld is not one of the last cost call", (long long)trial);
    }
    const int2 entry = c->mc_host_trials[(size_t)trial];
    const int first = c->has_molecules ? c->host_mol_start[(size_t)entry.x] : entry.x;
    const int size = c->has_molecules ? c->host_mol_start[(size_t)entry.x + 1] - first : 1;
    const bool rho_current = c->coulomb.kind == 1 && c->rho_positions_epoch == c->positions_epoch;
    int status = launch_move_accept(c, (int)trial, first, size, entry.y);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    // the resident positions changed; rho(k) was updated with them; the other trials of the batch are stale
    c->positions_epoch++;
    if (rho_current) c->rho_positions_epoch = c->positions_epoch;
    c->mc_positions_epoch = ~0ull;
    return LUMOL_CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// molecular dynamics
// ------------------------------------------------------------------------------------------------

extern "C" int32_t lumol_cuda_md_setup(lumol_cuda_context* ctx, int32_t integrator, double timestep) {
    MULTI_FAN(ctx, lumol_cuda_md_setup(child, integrator, timestep));
    CTX_OR_FAIL(ctx);
    if (integrator < LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET || integrator > LUMOL_CUDA_INTEGRATOR_ANISO_BERENDSEN_BAROSTAT) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "unknown integrator %d", integrator);
    }
    c->integrator = integrator;
    c->dt = timestep;
    for (int k = 0; k < 9; k++) c->barostat_eta[k] = (k % 4 == 0) ? 1.0 : 0.0;  // eta starts at one (integrators.rs:196, 276)
    int status = md_setup(c);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_md_set_degrees_of_freedom(lumol_cuda_context* ctx, int32_t mode, int64_t frozen) {
    MULTI_FAN(ctx, lumol_cuda_md_set_degrees_of_freedom(child, mode, frozen));
    CTX_OR_FAIL(ctx);
    if (mode != LUMOL_CUDA_DOF_PARTICLES && mode != LUMOL_CUDA_DOF_MOLECULES) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "unknown degrees-of-freedom mode %d", mode);
    }
    c->dof_mode = mode;
    c->dof_frozen = frozen;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_md_set_thermostat(lumol_cuda_context* ctx, int32_t thermostat, double temperature, double parameter) {
    MULTI_FAN(ctx, lumol_cuda_md_set_thermostat(child, thermostat, temperature, parameter));
    CTX_OR_FAIL(ctx);
    if (thermostat < LUMOL_CUDA_THERMOSTAT_NONE || thermostat > LUMOL_CUDA_THERMOSTAT_CSVR) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "unknown thermostat %d", thermostat);
    }
    if (thermostat != LUMOL_CUDA_THERMOSTAT_NONE && temperature < 0.0) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "The temperature must be positive in thermostats.");  // thermostats.rs:42
    }
    if (thermostat == LUMOL_CUDA_THERMOSTAT_BERENDSEN && !(parameter >= 1.0)) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "The timestep must be larger than 1 in berendsen thermostat.");  // thermostats.rs:100
    }
    if (thermostat == LUMOL_CUDA_THERMOSTAT_CSVR && !(parameter >= 1.0)) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "The timestep must be larger than 1 in CSVR thermostat.");  // thermostats.rs:160
    }
    c->thermostat = thermostat;
    c->thermostat_temperature = temperature;
    c->thermostat_parameter = parameter;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_md_set_csvr_noise(lumol_cuda_context* ctx, int64_t nsteps, const double* noise) {
    MULTI_FAN(ctx, lumol_cuda_md_set_csvr_noise(child, nsteps, noise));
    CTX_OR_FAIL(ctx);
    if (nsteps < 0 || (nsteps > 0 && noise == nullptr)) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "bad noise array");
    LUMOL_CUDA_CHECK(c, c->csvr_noise_dev.reserve((size_t)2 * nsteps + 2));
    if (nsteps > 0) {
        int status = upload(c, c->csvr_noise_dev.ptr, noise, (size_t)2 * nsteps * sizeof(double));
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    c->csvr_count = nsteps;
    c->csvr_cursor = 0;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_md_set_controls(lumol_cuda_context* ctx, uint32_t controls) {
    MULTI_FAN(ctx, lumol_cuda_md_set_controls(child, controls));
    CTX_OR_FAIL(ctx);
    if (controls & ~7u) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "unknown control bits");
    c->controls = controls;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_md_set_barostat(lumol_cuda_context* ctx, const double target[9], double tau) {
    MULTI_FAN(ctx, lumol_cuda_md_set_barostat(child, target, tau));
    CTX_OR_FAIL(ctx);
    if (target == nullptr || !(tau > 0.0)) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_md_set_barostat: bad arguments");
    for (int k = 0; k < 9; k++) c->barostat_target[k] = target[k];
    c->barostat_tau = tau;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_get_cell(lumol_cuda_context* ctx, double cell[9]) {
    MULTI_FAN(ctx, rank != 0 ? 0 : lumol_cuda_get_cell(child, cell));
    CTX_OR_FAIL(ctx);
    if (cell == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output");
    for (int k = 0; k < 9; k++) cell[k] = c->cell.h[k];
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_remove_rotation(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_remove_rotation(child));
    CTX_OR_FAIL(ctx);
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    int status = launch_remove_rotation(c);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_rewrap(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_rewrap(child));
    CTX_OR_FAIL(ctx);
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    PositionsChange change(c);
    int status = launch_rewrap(c);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

namespace lumol {

// UnitCell::lengths (cells.rs:134-146): distances between opposite faces
static void cell_lengths(const CellView& cell, double out[3]) {
    const double a[3] = {cell.h[0], cell.h[3], cell.h[6]}, b[3] = {cell.h[1], cell.h[4], cell.h[7]}, cc[3] = {cell.h[2], cell.h[5], cell.h[8]};
    auto cross = [](const double* u, const double* v, double* w) {
        w[0] = u[1] * v[2] - u[2] * v[1];
        w[1] = u[2] * v[0] - u[0] * v[2];
        w[2] = u[0] * v[1] - u[1] * v[0];
    };
    auto project = [](const double* normal, const double* v) {
        const double norm = std::sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
        return std::fabs((normal[0] / norm) * v[0] + (normal[1] / norm) * v[1] + (normal[2] / norm) * v[2]);
    };
    double na[3], nb[3], nc[3];
    cross(b, cc, na);
    cross(cc, a, nb);
    cross(a, b, nc);
    out[0] = project(na, a);
    out[1] = project(nb, b);
    out[2] = project(nc, cc);
}

// One step of BerendsenBarostat::integrate / AnisoBerendsenBarostat::integrate (integrators.rs:211-255, 295-341).
// Positions, velocities and forces stay on the device; the host sees the virial and the kinetic sums.
int barostat_step(Context* c) {
    lumol_cuda_context* handle = reinterpret_cast<lumol_cuda_context*>(c);
    const bool isotropic = c->integrator == LUMOL_CUDA_INTEGRATOR_BERENDSEN_BAROSTAT;
    if (c->nranks > 1) return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "the barostats are not available in sharded runs");
    if (c->cell.shape == LUMOL_CUDA_CELL_INFINITE) return c->fail(LUMOL_CUDA_ERROR_INFINITE_CELL, "can not scale infinite cells");  // cells.rs:204
    if (!(c->barostat_tau > 0.0)) return c->fail(LUMOL_CUDA_ERROR_STATE, "lumol_cuda_md_set_barostat was not called");
    double* eta = c->barostat_eta;
    int status = launch_barostat_drift(c, eta, isotropic);
    if (status) return status;
    // system.cell.scale_mut(factor): self.cell *= factor (cells.rs:203-207)
    double factor[9] = {0};
    if (isotropic) {
        factor[0] = factor[4] = factor[8] = eta[0] * eta[0] * eta[0] * 1.0;
    } else {
        for (int k = 0; k < 9; k++) factor[k] = eta[k];
    }
    double scaled[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            scaled[3 * i + j] = c->cell.h[3 * i] * factor[j] + c->cell.h[3 * i + 1] * factor[3 + j] + c->cell.h[3 * i + 2] * factor[6 + j];
    // UnitCell::scale_mut multiplies the matrix and keeps the shape (cells.rs:203-207): an orthorhombic cell scaled by a
    // matrix with off-diagonal terms still takes the orthorhombic branches (diagonal lengths) everywhere, as in lumol
    const int shape = c->cell.shape;
    status = lumol_cuda_set_cell(handle, scaled, shape);
    if (status) return status;
    double maximum_cutoff = -1.0;
    if (c->any_pair) maximum_cutoff = c->max_pair_cutoff;
    if (c->coulomb.kind != 0 && c->coulomb.rc > maximum_cutoff) maximum_cutoff = c->coulomb.rc;
    if (maximum_cutoff > 0.0) {
        double lengths[3];
        cell_lengths(c->cell, lengths);
        for (int d = 0; d < 3; d++) {
            if (0.5 * lengths[d] <= maximum_cutoff) {
                return c->fail(LUMOL_CUDA_ERROR_STATE,
                               "Tried to decrease the cell size in %sBerendesen barostat but the new size is smaller than the "
                               "interactions cut off radius. You can try to increase the cell size or the number of particles.",
                               isotropic ? "" : "anisotropic ");
            }
        }
    }
    // system.pressure() / system.stress() and system.forces() at the new positions: one device pass
    double virial[9];
    // system.pressure() / stress() use `Virial`, which is the molecular virial when the simulated degrees of freedom are
    // molecules (compute.rs:372-391)
    const uint32_t which = c->dof_mode == LUMOL_CUDA_DOF_MOLECULES ? LUMOL_CUDA_MOLECULAR_VIRIAL : LUMOL_CUDA_ATOMIC_VIRIAL;
    status = lumol_cuda_compute(handle, LUMOL_CUDA_FORCES | which, LUMOL_CUDA_PART_PAIRS | LUMOL_CUDA_PART_BONDED | LUMOL_CUDA_PART_COULOMB,
                                nullptr, nullptr, virial);
    if (status) return status;
    const double volume = volume_of(c->cell);
    if (isotropic) {
        double kinetic = 0.0;
        status = lumol_cuda_kinetic_energy(handle, &kinetic);
        if (status) return status;
        // compute.rs:134-171, 393-413: T = 2 K / (dof kB); P = (dof kB T + tr W) / (3 V)
        const double dof = c->dof_mode == LUMOL_CUDA_DOF_MOLECULES ? 3.0 * (double)c->nmol : (double)(3 * c->n - c->dof_frozen);
        const double temperature = 1.0 / (dof * K_BOLTZMANN) * 2.0 * kinetic;
        const double pressure = (dof * K_BOLTZMANN * temperature + (virial[0] + virial[4] + virial[8])) / (3.0 * volume);
        const double eta3 = 1.0 - 7372.0 / c->barostat_tau * (c->barostat_target[0] - pressure);
        eta[0] = eta[4] = eta[8] = std::cbrt(eta3);
    } else {
        double kinetic[9];
        status = lumol_cuda_kinetic_tensor(handle, kinetic);
        if (status) return status;
        const double scale = c->dt * 7372.0 / c->barostat_tau;
        for (int k = 0; k < 9; k++) {
            const double stress = (kinetic[k] + virial[k]) / volume;  // compute.rs:452-480
            eta[k] = ((k % 4 == 0) ? 1.0 : 0.0) - scale * (c->barostat_target[k] - stress);
        }
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < i; j++) {
                eta[3 * i + j] = 0.5 * (eta[3 * i + j] + eta[3 * j + i]);
                eta[3 * j + i] = eta[3 * i + j];
            }
        }
    }
    return launch_second_kick(c);
}

}  // namespace lumol

extern "C" int32_t lumol_cuda_md_run(lumol_cuda_context* ctx, int64_t nsteps) {
    MULTI_FAN(ctx, lumol_cuda_md_run(child, nsteps));
    CTX_OR_FAIL(ctx);
    if (c->integrator < 0) return c->fail(LUMOL_CUDA_ERROR_STATE, "lumol_cuda_md_setup was not called");
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    PositionsChange change(c);
    if (c->nranks > 1) {
        // a peer-exchange time-out of an earlier run must not be reported again
        LUMOL_CUDA_CHECK(c, cudaMemsetAsync(c->results.ptr + RES_FLAGS + 1, 0, sizeof(double), c->stream));
    }
    // Single-LJ systems on the neighbour-list path, NVE velocity-Verlet: the state lives in cell order for the duration of the
    // call, units of atoms are owned by ranks and only halo frames cross NVLink (pairs_lj2.cu)
    if (nsteps > 0 && sorted_md_applicable(c)) {
        int status = sync_tables(c);
        if (status) return status;
        status = sorted_md_run(c, nsteps);
        if (status) return status;
        LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        if (c->nranks > 1) {
            double timed_out = 0.0;
            LUMOL_CUDA_CHECK(c, cudaMemcpy(&timed_out, c->results.ptr + RES_FLAGS + 1, sizeof(double), cudaMemcpyDeviceToHost));
            if (timed_out != 0.0) {
                c->list_valid = false;
                return c->fail(LUMOL_CUDA_ERROR_COMM, "the positions of another rank did not arrive (peer exchange timed out)");
            }
        }
        int rebuilds = 0, overflow = 0;
        status = neighbor_list_status(c, &rebuilds, &overflow);
        if (status) return status;
        if (overflow == 2) {
            return c->fail(LUMOL_CUDA_ERROR_NOT_FINITE, "a particle position is not finite: the neighbour list cannot be built");
        }
        if (overflow) {
            return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "neighbour list capacity exceeded (strongly inhomogeneous density)");
        }
        return LUMOL_CUDA_SUCCESS;
    }
    // Small systems (all-pairs path: every system of the reference's tests and benches) are bound by launch latency:
    // one step in the middle of the run is captured into a CUDA graph and replayed.  The first
    // step runs eagerly (it allocates and settles the path), the last one as well (it ends with the half kick).
    int64_t s = 0;
    static const bool graphs_disabled = std::getenv("LUMOL_CUDA_NO_GRAPH") != nullptr;
    const bool plain_integrator = c->integrator >= LUMOL_CUDA_INTEGRATOR_VELOCITY_VERLET && c->integrator <= LUMOL_CUDA_INTEGRATOR_LEAP_FROG;
    if (!graphs_disabled && nsteps >= 8 && c->nranks == 1 && !c->profiling && plain_integrator && c->thermostat != LUMOL_CUDA_THERMOSTAT_CSVR) {
        int status = md_step(c, true, false);
        if (status) return status;
        s = 1;
        if (c->path == 0) {
            const int64_t launches_before = c->launches, step_before = c->md_step;
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            bool captured = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (captured) {
                status = md_step(c, false, false);
                captured = cudaStreamEndCapture(c->stream, &graph) == cudaSuccess && status == 0 && graph != nullptr;
                // what the capture "ran" was only recorded
                c->launches = launches_before;
                c->md_step = step_before;
                if (captured) captured = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            }
            if (captured) {
                const int64_t replays = nsteps - 2;
                int64_t per_step = 0;  // kernel and memset nodes of one step
                {
                    size_t nodes = 0;
                    if (cudaGraphGetNodes(graph, nullptr, &nodes) == cudaSuccess) per_step = (int64_t)nodes;
                }
                for (int64_t r = 0; r < replays; r++) {
                    LUMOL_CUDA_CHECK(c, cudaGraphLaunch(exec, c->stream));
                }
                c->launches += per_step * replays;
                c->md_step += replays;
                s += replays;
            } else {
                cudaGetLastError();
                if (status) return status;
            }
            if (exec != nullptr) cudaGraphExecDestroy(exec);
            if (graph != nullptr) cudaGraphDestroy(graph);
        }
    }
    for (; s < nsteps; s++) {
        int status = md_step(c, s == 0, s + 1 == nsteps);
        if (status) return status;
    }
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (c->nranks > 1) {
        double timed_out = 0.0;
        LUMOL_CUDA_CHECK(c, cudaMemcpy(&timed_out, c->results.ptr + RES_FLAGS + 1, sizeof(double), cudaMemcpyDeviceToHost));
        if (timed_out != 0.0) {
            return c->fail(LUMOL_CUDA_ERROR_COMM, "the positions of another rank did not arrive (peer exchange timed out)");
        }
    }
    if (c->path == 1) {
        int rebuilds = 0, overflow = 0;
        int status = neighbor_list_status(c, &rebuilds, &overflow);
        if (status) return status;
        if (overflow == 2) {
            return c->fail(LUMOL_CUDA_ERROR_NOT_FINITE, "a particle position is not finite: the neighbour list cannot be built");
        }
        if (overflow) {
            return c->fail(LUMOL_CUDA_ERROR_UNSUPPORTED, "neighbour list capacity exceeded (strongly inhomogeneous density)");
        }
    }
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_scale_velocities(lumol_cuda_context* ctx, double factor) {
    MULTI_FAN(ctx, lumol_cuda_scale_velocities(child, factor));
    CTX_OR_FAIL(ctx);
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    int status = launch_scale_velocities(c, factor, false);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_remove_translation(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_remove_translation(child));
    CTX_OR_FAIL(ctx);
    if (c->n == 0) return LUMOL_CUDA_SUCCESS;
    int status = launch_remove_translation(c);
    if (status) return status;
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------
// measurement
// ------------------------------------------------------------------------------------------------

extern "C" int32_t lumol_cuda_set_profiling(lumol_cuda_context* ctx, int32_t enabled) {
    MULTI_FAN(ctx, lumol_cuda_set_profiling(child, enabled));
    CTX_OR_FAIL(ctx);
    c->profiling = enabled != 0;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_get_stats(lumol_cuda_context* ctx, lumol_cuda_stats* stats) {
    MULTI_FAN(ctx, rank != 0 ? 0 : lumol_cuda_get_stats(child, stats));
    CTX_OR_FAIL(ctx);
    if (stats == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null stats");
    std::memset(stats, 0, sizeof(*stats));
    stats->natoms = c->n;
    stats->kernel_launches = c->launches;
    stats->neighbor_path = c->path;
    for (int d = 0; d < 3; d++) stats->ncells[d] = c->ncell[d];
    stats->nkvectors = c->nk;
    stats->pair_launches = c->clk_pair.launches;
    stats->pair_ms = c->clk_pair.ms;
    stats->kspace_launches = c->clk_kspace.launches;
    stats->kspace_ms = c->clk_kspace.ms;
    stats->integrate_launches = c->clk_integrate.launches;
    stats->integrate_ms = c->clk_integrate.ms;
    stats->neighbor_launches = c->clk_neighbor.launches;
    stats->neighbor_ms = c->clk_neighbor.ms;
    stats->comm_launches = c->clk_comm.launches;
    stats->comm_ms = c->clk_comm.ms;
    stats->pair_count = c->last_pair_count;
    stats->coulomb_pair_count = c->last_coulomb_pair_count;
    int rebuilds = 0, overflow = 0;
    int status = neighbor_list_status(c, &rebuilds, &overflow);
    if (status) return status;
    stats->neighbor_rebuilds = rebuilds;
    stats->neighbor_skin = c->skin_effective;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_reset_stats(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_reset_stats(child));
    CTX_OR_FAIL(ctx);
    c->launches = 0;
    c->clk_pair = c->clk_kspace = c->clk_integrate = c->clk_neighbor = c->clk_comm = KernelClock{};
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_neighbor_path(lumol_cuda_context* ctx, int32_t path) {
    MULTI_FAN(ctx, lumol_cuda_set_neighbor_path(child, path));
    CTX_OR_FAIL(ctx);
    if (path < -1 || path > 2) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "neighbour path must be -1, 0, 1 or 2");
    c->forced_path = path;
    c->structure_generation++;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_kspace_algorithm(lumol_cuda_context* ctx, int32_t algorithm) {
    MULTI_FAN(ctx, lumol_cuda_set_kspace_algorithm(child, algorithm));
    CTX_OR_FAIL(ctx);
    if (algorithm < -1 || algorithm > 1) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "the k-space algorithm must be -1, 0 or 1");
    }
    c->kspace_algorithm = algorithm;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_set_neighbor_skin(lumol_cuda_context* ctx, double skin) {
    MULTI_FAN(ctx, lumol_cuda_set_neighbor_skin(child, skin));
    CTX_OR_FAIL(ctx);
    if (!(skin >= 0.0)) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "the neighbour-list skin must be >= 0");
    c->skin = skin;
    c->structure_generation++;
    return LUMOL_CUDA_SUCCESS;
}

extern "C" void* lumol_cuda_stream(lumol_cuda_context* ctx) {
    if (ctx != nullptr && ctx->multi != nullptr) return (void*)multi_child(ctx, 0)->impl.stream;  // the first device's
    return ctx ? (void*)ctx->impl.stream : nullptr;
}

extern "C" int32_t lumol_cuda_synchronize(lumol_cuda_context* ctx) {
    MULTI_FAN(ctx, lumol_cuda_synchronize(child));
    CTX_OR_FAIL(ctx);
    LUMOL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_measure_fp64_peak(lumol_cuda_context* ctx, double* tflops) {
    MULTI_FAN(ctx, rank != 0 ? 0 : lumol_cuda_measure_fp64_peak(child, tflops));
    CTX_OR_FAIL(ctx);
    if (tflops == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output");
    return measure_fp64_peak(c, tflops);
}

extern "C" int32_t lumol_cuda_measure_copy_bandwidth(lumol_cuda_context* ctx, double* gbs) {
    MULTI_FAN(ctx, rank != 0 ? 0 : lumol_cuda_measure_copy_bandwidth(child, gbs));
    CTX_OR_FAIL(ctx);
    if (gbs == nullptr) return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "null output");
    return measure_copy_bandwidth(c, gbs);
}
