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
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <vector>

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
constexpr int NCCL_CHAR = 0;     // ncclChar
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
    // ---- peer-memory position exchange --------------------------------------------------------------
    // Instead of an NCCL all-gather after the drift, the drift kernel itself stores the new positions of the rank's
    // atoms into an inbox on every other GPU (NVLink peer stores through CUDA IPC mappings) and raises a flag
    // there; a gather kernel waits for the flags of all peers and copies their blocks out of the local inbox.
    // Two inbox copies alternate with the parity of the push number: a rank that is one step ahead writes into
    // the copy its peers are not reading.
    int peer_state = 0;  // 0: not set up, 1: ready, -1: unavailable (NCCL all-gather is used)
    int64_t peer_atoms = 0;
    int push_epoch = 0;
    DeviceBuffer<double> inbox;      // 2 x 3n
    DeviceBuffer<int> inbox_flags;   // 2 x PEER_MAX_RANKS flags, then the block counter
    DeviceBuffer<unsigned char> handle_staging;
    double* peer_inbox[PEER_MAX_RANKS] = {};
    int* peer_flags[PEER_MAX_RANKS] = {};
    std::vector<void*> opened;
    // buffers of the sorted-resident engine mapped from the other ranks (comm_map_peer_buffers)
    std::vector<void*> mapped;
    std::vector<void*> mapped_local;  // the local pointers the current mapping was made for
    std::vector<void*> mapped_peers;  // [buffer][rank]
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

// ------------------------------------------------------------------------------------------------
// peer-memory position exchange
// ------------------------------------------------------------------------------------------------

// What one rank tells the others about a device buffer.  Ranks of other processes map it through CUDA IPC; ranks of the
// same process (lumol_cuda_create_multi: one host process driving several devices) cannot open their own process's
// IPC handles and use the pointer itself after enabling peer access between the two devices.
struct PeerHandle {
    int32_t pid;
    int32_t device;
    void* raw;
    int64_t exported;  // the IPC handle below is valid
    cudaIpcMemHandle_t ipc;
};

static bool export_handle(Context* ctx, void* local, PeerHandle* out) {
    std::memset(out, 0, sizeof(PeerHandle));
    out->pid = (int32_t)getpid();
    out->device = ctx->device;
    out->raw = local;
    out->exported = cudaIpcGetMemHandle(&out->ipc, local) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    return true;  // a rank of another process finds out when it opens the handle
}

// `opened` collects what must be closed with cudaIpcCloseMemHandle later
static bool open_handle(Context* ctx, const PeerHandle& theirs, void** pointer, std::vector<void*>& opened) {
    if (theirs.pid == (int32_t)getpid()) {
        if (theirs.device != ctx->device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctx->device, theirs.device) != cudaSuccess || can == 0) return false;
            const cudaError_t status = cudaDeviceEnablePeerAccess(theirs.device, 0);
            if (status != cudaSuccess && status != cudaErrorPeerAccessAlreadyEnabled) return false;
            cudaGetLastError();
        }
        *pointer = theirs.raw;
        return true;
    }
    if (theirs.exported == 0 || cudaIpcOpenMemHandle(pointer, theirs.ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return false;
    opened.push_back(*pointer);
    return true;
}

static void peer_close(Comm* comm) {
    for (void* pointer : comm->opened) cudaIpcCloseMemHandle(pointer);
    comm->opened.clear();
    comm->peer_state = 0;
}

// Collective over the ranks: allocates the inboxes, exchanges their IPC handles through NCCL and maps the peers'.
// Any failure on any rank (no peer access, IPC not permitted) makes every rank fall back to the NCCL all-gather.
static int peer_setup(Context* ctx) {
    Comm* comm = ctx->comm;
    peer_close(comm);
    comm->peer_atoms = ctx->n;
    comm->push_epoch = 0;
    const int nranks = ctx->nranks;
    // Measured on 8 x B200 (1M-atom box): pushing from the drift kernel beats the NCCL all-gather with two ranks
    // (22 + 35 us against 58 + 19 us) but not with four or eight, where one kernel storing to three or seven peers
    // reaches only ~400 GB/s; LUMOL_CUDA_PEER_PUSH=1 / =0 forces the choice.
    const char* forced = std::getenv("LUMOL_CUDA_PEER_PUSH");
    bool ok = nranks <= PEER_MAX_RANKS && (forced != nullptr ? forced[0] == '1' : nranks == 2);
    // the second inbox copy starts at an even number of doubles: the drift kernel stores double2 (16 bytes) into it
    const size_t n3 = ((size_t)3 * ctx->n + 1) & ~(size_t)1;
    PeerHandle mine[2];
    std::memset(mine, 0, sizeof(mine));
    if (ok) {
        ok = comm->inbox.reserve(2 * n3) == cudaSuccess && comm->inbox_flags.reserve(2 * PEER_MAX_RANKS + 8) == cudaSuccess;
        if (ok) {
            ok = cudaMemsetAsync(comm->inbox_flags.ptr, 0, (2 * PEER_MAX_RANKS + 8) * sizeof(int), ctx->stream) == cudaSuccess &&
                 export_handle(ctx, comm->inbox.ptr, &mine[0]) && export_handle(ctx, comm->inbox_flags.ptr, &mine[1]);
        }
        cudaGetLastError();
    }
    // exchange: [ok byte + padding | two handles] per rank
    const size_t record = 16 + sizeof(mine);
    std::vector<unsigned char> host(record * (size_t)nranks, 0);
    LUMOL_CUDA_CHECK(ctx, comm->handle_staging.reserve(record * (size_t)nranks));
    unsigned char* own = host.data() + record * (size_t)ctx->rank;
    own[0] = ok ? 1 : 0;
    std::memcpy(own + 16, mine, sizeof(mine));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(comm->handle_staging.ptr + record * (size_t)ctx->rank, own, record, cudaMemcpyHostToDevice,
                                          ctx->stream));
    NCCL_CHECK(ctx, g_nccl.all_gather(comm->handle_staging.ptr + record * (size_t)ctx->rank, comm->handle_staging.ptr, record, NCCL_CHAR,
                                      comm->comm, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(host.data(), comm->handle_staging.ptr, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < nranks; p++) ok = ok && host[record * (size_t)p] == 1;
    if (ok) {
        for (int p = 0; p < nranks && ok; p++) {
            if (p == ctx->rank) {
                comm->peer_inbox[p] = comm->inbox.ptr;
                comm->peer_flags[p] = comm->inbox_flags.ptr;
                continue;
            }
            PeerHandle theirs[2];
            std::memcpy(theirs, host.data() + record * (size_t)p + 16, sizeof(theirs));
            void* inbox = nullptr;
            void* flags = nullptr;
            ok = open_handle(ctx, theirs[0], &inbox, comm->opened) && open_handle(ctx, theirs[1], &flags, comm->opened);
            comm->peer_inbox[p] = (double*)inbox;
            comm->peer_flags[p] = (int*)flags;
        }
        cudaGetLastError();
    }
    // every rank must take the same path: agree on the outcome
    double verdict = ok ? 0.0 : 1.0;
    LUMOL_CUDA_CHECK(ctx, comm->staging.reserve(8));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(comm->staging.ptr, &verdict, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, g_nccl.all_reduce(comm->staging.ptr, comm->staging.ptr, 1, NCCL_FLOAT64, NCCL_SUM, comm->comm, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(&verdict, comm->staging.ptr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (verdict != 0.0) {
        peer_close(comm);
        comm->peer_state = -1;
    } else {
        comm->peer_state = 1;
    }
    return 0;
}

// Fills `push` for the next drift kernel (collective: every rank calls it at the same step).  push->nranks stays 0
// when the peer path is unavailable and the caller must use comm_allgather_positions instead.
int comm_peer_push_begin(Context* ctx, PeerPush* push) {
    *push = PeerPush();
    if (ctx->nranks <= 1 || ctx->comm == nullptr) return 0;
    Comm* comm = ctx->comm;
    if (comm->peer_state == 0 || comm->peer_atoms != ctx->n) {
        int status = peer_setup(ctx);
        if (status != 0) return status;
    }
    if (comm->peer_state != 1) return 0;
    comm->push_epoch++;
    const int parity = comm->push_epoch & 1;
    push->nranks = ctx->nranks;
    push->rank = ctx->rank;
    push->epoch = comm->push_epoch;
    for (int p = 0; p < ctx->nranks; p++) {
        push->inbox[p] = comm->peer_inbox[p] + (size_t)parity * (((size_t)3 * ctx->n + 1) & ~(size_t)1);
        push->flags[p] = comm->peer_flags[p] + parity * PEER_MAX_RANKS;
    }
    push->counter = comm->inbox_flags.ptr + 2 * PEER_MAX_RANKS;
    return 0;
}

// Waits until every peer's block of this push has arrived in the local inbox, then copies the blocks into the
// position array.  RES_FLAGS + 1 is raised (and the wait abandoned) after about twenty seconds.
__global__ void __launch_bounds__(256)
    peer_gather_kernel(int nranks, int rank, int epoch, const int* __restrict__ flags, const double* __restrict__ inbox,
                       double* __restrict__ position, int64_t chunk3, int64_t n3, double* __restrict__ results) {
    __shared__ int proceed;
    if (threadIdx.x == 0) {
        int ok = 1;
        const long long start = clock64();
        for (int p = 0; p < nranks; p++) {
            if (p == rank) continue;
            const volatile int* flag = flags + p;
            while (*flag < epoch) {
                if (clock64() - start > 40000000000ll) {
                    ok = 0;
                    break;
                }
            }
        }
        __threadfence_system();
        proceed = ok;
    }
    __syncthreads();
    if (!proceed) {
        if (blockIdx.x == 0 && threadIdx.x == 0) results[RES_FLAGS + 1] = 1.0;
        return;
    }
    const int64_t own_lo = chunk3 * rank, own_hi = min(n3, own_lo + chunk3);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += (int64_t)gridDim.x * blockDim.x) {
        if (k < own_lo || k >= own_hi) position[k] = __ldcv(inbox + k);
    }
}

int comm_peer_gather(Context* ctx, const PeerPush& push) {
    if (push.nranks <= 1) return 0;
    ScopedClock clock(ctx, &ctx->clk_comm);
    const int64_t n3 = 3 * ctx->n;
    const int64_t chunk3 = 3 * ((ctx->n + ctx->nranks - 1) / ctx->nranks);
    int blocks = (int)((n3 + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;  // every block spins on the flags: all must be resident
    peer_gather_kernel<<<blocks, 256, 0, ctx->stream>>>(push.nranks, push.rank, push.epoch, push.flags[push.rank], push.inbox[push.rank],
                                                        ctx->position.ptr, chunk3, n3, ctx->results.ptr);
    ctx->launches++;
    ctx->clk_comm.launches++;
    LUMOL_CUDA_CHECK(ctx, cudaGetLastError());
    return 0;
}

// all-gather of equal blocks of `chunk` doubles, in place: rank r owns [r * chunk, (r + 1) * chunk)
int comm_allgather_chunks(Context* ctx, double* data, size_t chunk) {
    if (ctx->nranks <= 1 || chunk == 0) return 0;
    ScopedClock clock(ctx, &ctx->clk_comm);
    NCCL_CHECK(ctx, g_nccl.all_gather(data + (size_t)ctx->rank * chunk, data, chunk, NCCL_FLOAT64, ctx->comm->comm, ctx->stream));
    ctx->clk_comm.launches++;
    return 0;
}

static void unmap_peer_buffers(Comm* comm) {
    for (void* pointer : comm->mapped) cudaIpcCloseMemHandle(pointer);
    comm->mapped.clear();
    comm->mapped_local.clear();
    comm->mapped_peers.clear();
}

// Collective over the ranks: maps `count` device buffers of every rank into this process through CUDA IPC (NVLink peer
// access).  peers[b * PEER_MAX_RANKS + p] is rank p's buffer b (the local pointer for p == rank).  The mapping is kept
// until one of the local pointers changes (every rank calls this at the same points, so they re-map together).
// *ok = false when some rank could not export or open a handle: every rank then takes its fallback.
int comm_map_peer_buffers(Context* ctx, int count, void* const* local, void** peers, bool* ok) {
    *ok = false;
    if (ctx->nranks <= 1 || ctx->comm == nullptr || ctx->nranks > PEER_MAX_RANKS) return 0;
    Comm* comm = ctx->comm;
    const int nranks = ctx->nranks;
    bool same = (int)comm->mapped_local.size() == count;
    for (int b = 0; same && b < count; b++) same = comm->mapped_local[(size_t)b] == local[b];
    // every rank must agree on whether to re-map: a rank whose pointers moved forces all of them
    double moved = same ? 0.0 : 1.0;
    LUMOL_CUDA_CHECK(ctx, comm->staging.reserve(8));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(comm->staging.ptr, &moved, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, g_nccl.all_reduce(comm->staging.ptr, comm->staging.ptr, 1, NCCL_FLOAT64, NCCL_SUM, comm->comm, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(&moved, comm->staging.ptr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (moved == 0.0) {
        for (int k = 0; k < count * PEER_MAX_RANKS; k++) peers[k] = comm->mapped_peers[(size_t)k];
        *ok = true;
        return 0;
    }
    unmap_peer_buffers(comm);
    bool fine = true;
    std::vector<PeerHandle> mine((size_t)count);
    for (int b = 0; b < count; b++) fine = fine && export_handle(ctx, local[b], &mine[(size_t)b]);
    cudaGetLastError();
    const size_t record = 16 + (size_t)count * sizeof(PeerHandle);
    std::vector<unsigned char> host(record * (size_t)nranks, 0);
    LUMOL_CUDA_CHECK(ctx, comm->handle_staging.reserve(record * (size_t)nranks));
    unsigned char* own = host.data() + record * (size_t)ctx->rank;
    own[0] = fine ? 1 : 0;
    std::memcpy(own + 16, mine.data(), (size_t)count * sizeof(PeerHandle));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(comm->handle_staging.ptr + record * (size_t)ctx->rank, own, record, cudaMemcpyHostToDevice,
                                          ctx->stream));
    NCCL_CHECK(ctx, g_nccl.all_gather(comm->handle_staging.ptr + record * (size_t)ctx->rank, comm->handle_staging.ptr, record, NCCL_CHAR,
                                      comm->comm, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(host.data(), comm->handle_staging.ptr, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < nranks; p++) fine = fine && host[record * (size_t)p] == 1;
    std::vector<void*> table((size_t)count * PEER_MAX_RANKS, nullptr);
    for (int p = 0; p < nranks && fine; p++) {
        for (int b = 0; b < count && fine; b++) {
            if (p == ctx->rank) {
                table[(size_t)b * PEER_MAX_RANKS + p] = local[b];
                continue;
            }
            PeerHandle theirs;
            std::memcpy(&theirs, host.data() + record * (size_t)p + 16 + (size_t)b * sizeof(PeerHandle), sizeof(theirs));
            void* pointer = nullptr;
            fine = open_handle(ctx, theirs, &pointer, comm->mapped);
            if (fine) table[(size_t)b * PEER_MAX_RANKS + p] = pointer;
        }
    }
    cudaGetLastError();
    double verdict = fine ? 0.0 : 1.0;
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(comm->staging.ptr, &verdict, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, g_nccl.all_reduce(comm->staging.ptr, comm->staging.ptr, 1, NCCL_FLOAT64, NCCL_SUM, comm->comm, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaMemcpyAsync(&verdict, comm->staging.ptr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (verdict != 0.0) {
        unmap_peer_buffers(comm);
        return 0;
    }
    comm->mapped_local.assign(local, local + count);
    comm->mapped_peers = table;
    for (int k = 0; k < count * PEER_MAX_RANKS; k++) peers[k] = table[(size_t)k];
    *ok = true;
    return 0;
}

// barrier over the ranks on the context stream (host returns when every rank has reached it)
int comm_barrier(Context* ctx) {
    if (ctx->nranks <= 1 || ctx->comm == nullptr) return 0;
    LUMOL_CUDA_CHECK(ctx, ctx->comm->staging.reserve(8));
    LUMOL_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->comm->staging.ptr, 0, sizeof(double), ctx->stream));
    NCCL_CHECK(ctx, g_nccl.all_reduce(ctx->comm->staging.ptr, ctx->comm->staging.ptr, 1, NCCL_FLOAT64, NCCL_SUM, ctx->comm->comm, ctx->stream));
    LUMOL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

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
        peer_close(ctx->comm);
        unmap_peer_buffers(ctx->comm);
        ctx->comm->staging.release();
        ctx->comm->inbox.release();
        ctx->comm->inbox_flags.release();
        ctx->comm->handle_staging.release();
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
    if (ctx->multi != nullptr) {
        return c->fail(LUMOL_CUDA_ERROR_STATE, "lumol_cuda_comm_init: a multi-device context already shards over its own devices");
    }
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
