// lumol_cuda_create_multi: one context, several devices of one process.  The parent context owns one child context per
// device (the ordinary sharded contexts of comm.cu, rank r on devices[r]) and one host thread per child; every entry
// point of the ABI called on the parent is handed to all the threads at once and returns when the last one has, so
// the collectives inside a call (NCCL, peer-memory flags) meet the way they do with one process per GPU.  A lumol
// process keeps its single `System` and its single thread of control and still uses every GPU.
#include "context.hpp"

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>

namespace lumol {

struct Multi {
    std::vector<lumol_cuda_context*> children;
    std::vector<std::thread> threads;
    std::vector<int32_t> status;
    std::mutex mutex;
    std::condition_variable wake, done;
    const MultiTask* task = nullptr;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
};

static void worker(Multi* multi, int rank) {
    uint64_t seen = 0;
    for (;;) {
        const MultiTask* task = nullptr;
        {
            std::unique_lock<std::mutex> lock(multi->mutex);
            multi->wake.wait(lock, [&] { return multi->stop || multi->generation != seen; });
            if (multi->stop) return;
            seen = multi->generation;
            task = multi->task;
        }
        const int32_t status = (*task)(multi->children[(size_t)rank], rank);
        {
            std::lock_guard<std::mutex> lock(multi->mutex);
            multi->status[(size_t)rank] = status;
            multi->pending--;
        }
        multi->done.notify_one();
    }
}

int32_t multi_run(lumol_cuda_context* parent, const MultiTask& task) {
    Multi* multi = parent->multi;
    {
        std::unique_lock<std::mutex> lock(multi->mutex);
        multi->task = &task;
        multi->pending = (int)multi->children.size();
        multi->generation++;
        multi->wake.notify_all();
        multi->done.wait(lock, [&] { return multi->pending == 0; });
        multi->task = nullptr;
    }
    for (size_t r = 0; r < multi->children.size(); r++) {
        if (multi->status[r] < 0) {
            parent->impl.error = multi->children[r]->impl.error;
            return multi->status[r];
        }
    }
    return multi->status[0];
}

int multi_size(const lumol_cuda_context* parent) { return parent->multi == nullptr ? 0 : (int)parent->multi->children.size(); }

lumol_cuda_context* multi_child(const lumol_cuda_context* parent, int rank) { return parent->multi->children[(size_t)rank]; }

void multi_destroy(lumol_cuda_context* parent) {
    Multi* multi = parent->multi;
    if (multi == nullptr) return;
    {
        std::lock_guard<std::mutex> lock(multi->mutex);
        multi->stop = true;
    }
    multi->wake.notify_all();
    for (std::thread& thread : multi->threads) thread.join();
    // the communicators are torn down together (ncclCommDestroy of one rank may wait for the others)
    std::vector<std::thread> closers;
    for (lumol_cuda_context* child : multi->children) closers.emplace_back([child] { lumol_cuda_destroy(child); });
    for (std::thread& thread : closers) thread.join();
    delete multi;
    parent->multi = nullptr;
}

}  // namespace lumol

using namespace lumol;

extern "C" int32_t lumol_cuda_create_multi(const int32_t* devices, int32_t ndevices, lumol_cuda_context** out) {
    if (out == nullptr) return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
    *out = nullptr;
    if (devices == nullptr || ndevices < 1 || ndevices > PEER_MAX_RANKS) {
        set_create_error("lumol_cuda_create_multi: between 1 and 8 devices");
        return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
    }
    for (int a = 0; a < ndevices; a++) {
        for (int b = 0; b < a; b++) {
            if (devices[a] == devices[b]) {
                set_create_error("lumol_cuda_create_multi: a device is listed twice");
                return LUMOL_CUDA_ERROR_INVALID_ARGUMENT;
            }
        }
    }
    if (ndevices == 1) return lumol_cuda_create(devices[0], out);
    lumol_cuda_context* parent = new (std::nothrow) lumol_cuda_context();
    Multi* multi = new (std::nothrow) Multi();
    if (parent == nullptr || multi == nullptr) {
        delete parent;
        delete multi;
        set_create_error("out of host memory");
        return LUMOL_CUDA_ERROR_CUDA;
    }
    parent->impl.device = devices[0];
    for (int r = 0; r < ndevices; r++) {
        lumol_cuda_context* child = nullptr;
        const int32_t status = lumol_cuda_create(devices[r], &child);  // leaves its message for lumol_cuda_last_error(NULL)
        if (status < 0) {
            for (lumol_cuda_context* made : multi->children) lumol_cuda_destroy(made);
            delete multi;
            delete parent;
            return status;
        }
        child->impl.discard_downloads = r > 0;
        multi->children.push_back(child);
    }
    multi->status.assign((size_t)ndevices, 0);
    parent->multi = multi;
    for (int r = 0; r < ndevices; r++) multi->threads.emplace_back(worker, multi, r);
    uint8_t id[128];
    int32_t status = lumol_cuda_comm_unique_id(id);
    if (status == 0) {
        const uint8_t* shared = id;
        status = multi_run(parent, [=](lumol_cuda_context* child, int rank) -> int32_t { return lumol_cuda_comm_init(child, ndevices, rank, shared); });
    } else {
        set_create_error("lumol_cuda_create_multi: cannot load NCCL (libnccl.so.2)");
    }
    if (status < 0) {
        if (!parent->impl.error.empty()) set_create_error(parent->impl.error.c_str());
        multi_destroy(parent);
        delete parent;
        return status;
    }
    *out = parent;
    return LUMOL_CUDA_SUCCESS;
}
