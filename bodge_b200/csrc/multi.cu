// Several GPUs driven by ONE process through the C ABI: column shards, a replica of the matrix per
// GPU, no communication during the recursion, ONE NCCL collective over NVLink at the end (all-reduce of the summed
// moments, or all-gather of the per-column ones).  The reference has nothing of the kind ("no support for e.g. MPI",
// README.md:36-39: one box is the whole machine); with torch.distributed the same partition runs one process per GPU
// (bodge_b200/distributed.py) -- this is the path for callers without torch.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2, or of $BDG_NCCL_LIB): libbdg.so keeps linking against nothing
// but the CUDA runtime, and a single-GPU caller never needs the library to be present.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "bdg_internal.h"

namespace {

typedef struct ncclComm *ncclComm_t;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;  // nccl.h: ncclDataType_t / ncclRedOp_t

struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

std::mutex g_mutex;
NcclApi g_nccl;
std::map<std::vector<int>, std::vector<ncclComm_t>> g_comms;  // one clique per device list, created once

int load_nccl() {
    if (g_nccl.lib) return BDG_OK;
    const char *env = getenv("BDG_NCCL_LIB");
    const char *names[] = {env && *env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *name : names) {
        lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        bdg_set_error("NCCL not found (%s): set BDG_NCCL_LIB to the path of libnccl.so.2", dlerror());
        return BDG_E_INVALID;
    }
    NcclApi api;
    api.lib = lib;
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(lib, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!api.CommInitAll || !api.CommDestroy || !api.AllReduce || !api.AllGather || !api.GroupStart || !api.GroupEnd) {
        bdg_set_error("the NCCL library lacks a required symbol");
        return BDG_E_INVALID;
    }
    g_nccl = api;
    return BDG_OK;
}

#define BDG_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        int rc__ = (expr);                                                                               \
        if (rc__ != 0) {                                                                                 \
            bdg_set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc__) : "?"); \
            return BDG_E_CUDA;                                                                           \
        }                                                                                                \
    } while (0)

int clique(const std::vector<int> &devices, std::vector<ncclComm_t> **out) {
    auto it = g_comms.find(devices);
    if (it == g_comms.end()) {
        std::vector<ncclComm_t> comms(devices.size(), nullptr);
        BDG_NCCL(g_nccl.CommInitAll(comms.data(), (int)devices.size(), devices.data()));
        it = g_comms.emplace(devices, std::move(comms)).first;
    }
    *out = &it->second;
    return BDG_OK;
}

// contiguous, balanced split: the first n % parts shards get one extra item (bodge_b200/distributed.py: shard_range)
void shard(int64_t n, int part, int parts, int64_t *lo, int64_t *hi) {
    const int64_t base = n / parts, extra = n % parts;
    *lo = part * base + std::min<int64_t>(part, extra);
    *hi = *lo + base + (part < extra ? 1 : 0);
}

}  // namespace

extern "C" int bdg_cheb_moments_multi(bdg_t **sys, int n_gpu, int kind, int64_t n_cols, const int64_t *probe_rows,
                                      uint64_t seed, double scale, int32_t n_moments, int reduce, double *mu) {
    BDG_REQUIRE(sys && n_gpu >= 1 && n_gpu <= 64, "need 1..64 handles");
    BDG_REQUIRE(n_cols >= n_gpu, "need at least one column per GPU (%lld columns, %d GPUs)", (long long)n_cols, n_gpu);
    BDG_REQUIRE(n_moments >= 1 && mu, "bad output arguments");
    BDG_REQUIRE(reduce == BDG_MU_PER_COLUMN || reduce == BDG_MU_SUM, "unknown reduce mode");
    std::vector<int> devices;
    for (int g = 0; g < n_gpu; ++g) {
        BDG_REQUIRE(sys[g] != nullptr, "null handle %d", g);
        for (int d : devices) BDG_REQUIRE(d != sys[g]->device, "two handles on device %d: one replica per GPU", d);
        devices.push_back(sys[g]->device);
    }
    if (n_gpu == 1)
        return bdg_cheb_moments(sys[0], kind, (int32_t)n_cols, probe_rows, seed, 0, scale, n_moments, reduce, mu, 0);

    std::lock_guard<std::mutex> lock(g_mutex);
    BDG_TRY(load_nccl());
    std::vector<ncclComm_t> *comms = nullptr;
    BDG_TRY(clique(devices, &comms));

    // 1. every GPU starts the recursion on its shard of the columns (launches only: the GPUs run concurrently)
    std::vector<int64_t> lo(n_gpu), hi(n_gpu);
    int64_t widest = 0;
    for (int g = 0; g < n_gpu; ++g) {
        shard(n_cols, g, n_gpu, &lo[g], &hi[g]);
        widest = std::max(widest, hi[g] - lo[g]);
        BDG_TRY(bdg_cheb_begin(sys[g], kind, (int32_t)(hi[g] - lo[g]), probe_rows ? probe_rows + lo[g] : nullptr, seed, lo[g],
                               scale, BDG_KERNEL_AUTO_MOMENTS));
    }
    for (int g = 0; g < n_gpu; ++g)
        BDG_TRY(bdg_cheb_steps(sys[g], std::max(0, (n_moments + 1) / 2 - 1 - sys[g]->cheb.steps_done), nullptr));

    // 2. local moments into device buffers, 3. ONE collective on the handles' own streams
    const size_t send_count = reduce == BDG_MU_SUM ? (size_t)n_moments : (size_t)n_moments * (size_t)widest;
    for (int g = 0; g < n_gpu; ++g) {
        BDG_CUDA(cudaSetDevice(sys[g]->device));
        BDG_TRY(dev_alloc(sys[g], sys[g]->multi_send, send_count * sizeof(double)));
        if (reduce != BDG_MU_SUM) {
            BDG_TRY(dev_alloc(sys[g], sys[g]->multi_recv, send_count * n_gpu * sizeof(double)));
            BDG_CUDA(cudaMemsetAsync(sys[g]->multi_send.ptr, 0, send_count * sizeof(double), sys[g]->stream));
        }
        BDG_TRY(bdg_cheb_moments_read(sys[g], n_moments, reduce, sys[g]->multi_send.as<double>(), 1));
    }
    BDG_NCCL(g_nccl.GroupStart());
    for (int g = 0; g < n_gpu; ++g) {
        if (reduce == BDG_MU_SUM)
            BDG_NCCL(g_nccl.AllReduce(sys[g]->multi_send.ptr, sys[g]->multi_send.ptr, send_count, kNcclFloat64, kNcclSum, (*comms)[g],
                                      sys[g]->stream));
        else
            BDG_NCCL(g_nccl.AllGather(sys[g]->multi_send.ptr, sys[g]->multi_recv.ptr, send_count, kNcclFloat64, (*comms)[g],
                                      sys[g]->stream));
    }
    BDG_NCCL(g_nccl.GroupEnd());

    // 4. the result is on every GPU; the host reads GPU 0's copy
    BDG_CUDA(cudaSetDevice(sys[0]->device));
    if (reduce == BDG_MU_SUM) {
        BDG_CUDA(cudaMemcpyAsync(mu, sys[0]->multi_send.ptr, send_count * sizeof(double), cudaMemcpyDeviceToHost, sys[0]->stream));
        BDG_CUDA(cudaStreamSynchronize(sys[0]->stream));
    } else {
        std::vector<double> all(send_count * n_gpu);
        BDG_CUDA(cudaMemcpyAsync(all.data(), sys[0]->multi_recv.ptr, all.size() * sizeof(double), cudaMemcpyDeviceToHost, sys[0]->stream));
        BDG_CUDA(cudaStreamSynchronize(sys[0]->stream));
        for (int g = 0; g < n_gpu; ++g) {  // shard g wrote [n_moments][k_g] at the start of its piece
            const int64_t k = hi[g] - lo[g];
            const double *piece = all.data() + (size_t)g * send_count;
            for (int32_t n = 0; n < n_moments; ++n)
                for (int64_t c = 0; c < k; ++c) mu[(size_t)n * n_cols + lo[g] + c] = piece[(size_t)n * k + c];
        }
    }
    for (int g = 1; g < n_gpu; ++g) {  // the other GPUs' streams are done with the buffers before the call returns
        BDG_CUDA(cudaSetDevice(sys[g]->device));
        BDG_CUDA(cudaStreamSynchronize(sys[g]->stream));
    }
    return BDG_OK;
}

extern "C" int bdg_multi_release(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_nccl.lib)
        for (auto &kv : g_comms)
            for (ncclComm_t c : kv.second)
                if (c) g_nccl.CommDestroy(c);
    g_comms.clear();
    return BDG_OK;
}
