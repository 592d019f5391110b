// Multi-GPU plumbing: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The library does not link NCCL: it dlopen()s libnccl.so.2 when lumol_cuda_comm_init is called, so a
// single-GPU user (and the CPU-only symbol tests) never need it, and a process that already loaded
// PyTorch's bundled NCCL shares that copy.
//
// Sharding: every rank keeps all positions resident and owns one contiguous block of
// atoms, for which it evaluates forces (full neighbour shell, so no force reduction) and integrates.
// Collectives per MD step: one all-gather of the positions after the drift; per Ewald evaluation one
// all-reduce of rho(k); per energy/virial query one all-reduce of the scalar sums.
#include "context.hpp"

#include <dlfcn.h>

#include <cstring>

namespace lumol {

// minimal NCCL ABI (nccl.h): opaque communicator, 128-byte unique id, enums as ints
struct NcclUniqueId {
    char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

constexpr int NCCL_FLOAT64 = 8;  // ncclDouble
constexpr int NCCL_SUM = 0;      // ncclSum

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_get_error_string get_error_string = nullptr;
};

static NcclApi g_nccl;

static bool load_nccl(std::string& error) {
    if (g_nccl.handle != nullptr) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* handle = nullptr;
    for (const char* name : names) {
        handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (handle != nullptr) break;
    }
    if (handle == nullptr) {
        error = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(handle, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(handle, "ncclCommInitRank");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(handle, "ncclCommDestroy");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(handle, "ncclAllReduce");
    g_nccl.all_gather = (fn_all_gather)dlsym(handle, "ncclAllGather");
    g_nccl.get_error_string = (fn_get_error_string)dlsym(handle, "ncclGetErrorString");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce ||
        !g_nccl.all_gather || !g_nccl.get_error_string) {
        error = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    g_nccl.handle = handle;
    return true;
}

struct Comm {
    NcclComm comm = nullptr;
    DeviceBuffer<double> staging;
};

#define NCCL_CHECK(ctx, expr)                                                                               \
    do {                                                                                                    \
        int rc__ = (expr);                                                                                  \
        if (rc__ != 0) {                                                                                    \
            return (ctx)->fail(LUMOL_CUDA_ERROR_COMM, "%s failed: %s", #expr, g_nccl.get_error_string(rc__)); \
        }                                                                                                   \
    } while (0)

// all-gather of equal blocks, in place: rank r owns elements [r * chunk * width, (r + 1) * chunk * width)
static int allgather_in_place(Context* ctx, double* data, int64_t n_items, int width) {
    if (ctx->nranks <= 1) return 0;
    const int64_t chunk = (n_items + ctx->nranks - 1) / ctx->nranks;
    const size_t count = (size_t)chunk * width;
    ScopedClock clock(ctx, &ctx->clk_comm);
    NCCL_CHECK(ctx, g_nccl.all_gather(data + (size_t)ctx->rank * count, data, count, NCCL_FLOAT64, ctx->comm->comm,
                                      ctx->stream));
    ctx->clk_comm.launches++;
    return 0;
}

int comm_allgather_positions(Context* ctx) { return allgather_in_place(ctx, ctx->position.ptr, ctx->n, 3); }

int comm_allgather_blocks(Context* ctx, double* data, int64_t total) { return allgather_in_place(ctx, data, total / 3, 3); }

int comm_allreduce(Context* ctx, double* data, int64_t count) {
    if (ctx->nranks <= 1) return 0;
    ScopedClock clock(ctx, &ctx->clk_comm);
    NCCL_CHECK(ctx, g_nccl.all_reduce(data, data, (size_t)count, NCCL_FLOAT64, NCCL_SUM, ctx->comm->comm, ctx->stream));
    ctx->clk_comm.launches++;
    return 0;
}

void comm_destroy(Context* ctx) {
    if (ctx->comm != nullptr) {
        if (ctx->comm->comm != nullptr && g_nccl.comm_destroy != nullptr) {
            g_nccl.comm_destroy(ctx->comm->comm);
        }
        ctx->comm->staging.release();
        delete ctx->comm;
        ctx->comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
}

}  // namespace lumol

using namespace lumol;

extern "C" int32_t lumol_cuda_comm_unique_id(uint8_t id[128]) {
    std::string error;
    if (id == nullptr) return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
    if (!load_nccl(error)) return LUMOL_CUDA_ERROR_COMM;
    NcclUniqueId uid;
    if (g_nccl.get_unique_id(&uid) != 0) return LUMOL_CUDA_ERROR_COMM;
    std::memcpy(id, uid.internal, 128);
    return LUMOL_CUDA_SUCCESS;
}

extern "C" int32_t lumol_cuda_comm_init(lumol_cuda_context* ctx, int32_t nranks, int32_t rank, const uint8_t id[128]) {
    if (ctx == nullptr) return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
    Context* c = &ctx->impl;
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks || (nranks > 1 && id == nullptr)) {
        return c->fail(LUMOL_CUDA_ERROR_INVALID_ARGUMENT, "lumol_cuda_comm_init: bad rank %d of %d", rank, nranks);
    }
    if (cudaSetDevice(c->device) != cudaSuccess) return c->fail(LUMOL_CUDA_ERROR_CUDA, "cudaSetDevice failed");
    comm_destroy(c);
    if (nranks == 1) return LUMOL_CUDA_SUCCESS;
    std::string error;
    if (!load_nccl(error)) return c->fail(LUMOL_CUDA_ERROR_COMM, "%s", error.c_str());
    c->comm = new Comm();
    NcclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    NCCL_CHECK(c, g_nccl.comm_init_rank(&c->comm->comm, nranks, uid, rank));
    c->nranks = nranks;
    c->rank = rank;
    c->structure_generation++;
    return LUMOL_CUDA_SUCCESS;
}
