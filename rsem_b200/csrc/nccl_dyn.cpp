// NCCL is loaded with dlopen at first use so that librsem_b200.so itself has no link-time
// dependency on it: single-GPU runs (and loading the library on a machine without NCCL) never
// touch it, and inside a PyTorch process the already-loaded libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"

namespace rsem_b200 {

// minimal, ABI-stable subset of nccl.h
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*);
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
    int (*CommDestroy)(NcclComm);
    const char* (*GetErrorString)(int);
};

static NcclApi g_api;
static bool g_loaded = false, g_tried = false;
static std::mutex g_mu;

const NcclApi* nccl_api() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_loaded) return &g_api;
    if (g_tried) { set_error("libnccl.so.2 could not be loaded"); return nullptr; }
    g_tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        return nullptr;
    }
    g_api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    g_api.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    g_api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllReduce || !g_api.CommDestroy) {
        set_error("libnccl.so.2 lacks a required symbol");
        return nullptr;
    }
    g_loaded = true;
    return &g_api;
}

static int nccl_fail(const NcclApi* api, int rc, const char* what) {
    std::string msg = std::string("NCCL error in ") + what + ": ";
    msg += api->GetErrorString ? api->GetErrorString(rc) : "unknown";
    set_error(msg);
    return RSEM_B200_ERR_NCCL;
}

int nccl_unique_id(void* id128) {
    const NcclApi* api = nccl_api();
    if (!api) return RSEM_B200_ERR_NCCL;
    NcclUniqueId id;
    int rc = api->GetUniqueId(&id);
    if (rc != kNcclSuccess) return nccl_fail(api, rc, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof id);
    return 0;
}

int nccl_comm_init(void** comm, const void* id128, int n_ranks, int rank) {
    const NcclApi* api = nccl_api();
    if (!api) return RSEM_B200_ERR_NCCL;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof id);
    int rc = api->CommInitRank(comm, n_ranks, id, rank);
    if (rc != kNcclSuccess) return nccl_fail(api, rc, "ncclCommInitRank");
    return 0;
}

int nccl_allreduce_sum_f64(void* comm, double* buf, size_t n, cudaStream_t stream) {
    const NcclApi* api = nccl_api();
    if (!api) return RSEM_B200_ERR_NCCL;
    int rc = api->AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, comm, stream);
    if (rc != kNcclSuccess) return nccl_fail(api, rc, "ncclAllReduce");
    return 0;
}

int nccl_allgather_bytes(void* comm, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t stream) {
    const NcclApi* api = nccl_api();
    if (!api) return RSEM_B200_ERR_NCCL;
    if (!api->AllGather) { set_error("libnccl.so.2 lacks ncclAllGather"); return RSEM_B200_ERR_NCCL; }
    int rc = api->AllGather(send, recv, bytes_per_rank, /* ncclInt8 */ 0, comm, stream);
    if (rc != kNcclSuccess) return nccl_fail(api, rc, "ncclAllGather");
    return 0;
}

int nccl_comm_destroy(void* comm) {
    const NcclApi* api = nccl_api();
    if (!api) return RSEM_B200_ERR_NCCL;
    api->CommDestroy(comm);
    return 0;
}

}  // namespace rsem_b200
