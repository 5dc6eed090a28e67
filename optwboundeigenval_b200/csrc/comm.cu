// comm.cu -- thin run-time binding to NCCL (see comm.h).
#include <dlfcn.h>
#include <string.h>

#include "comm.h"
#include "common.cuh"

namespace b2s {

// minimal slice of nccl.h (ABI stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        set_error("NCCL not found: %s", dlerror());
        return nullptr;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
        set_error("NCCL symbols missing in the loaded libnccl");
        api.handle = nullptr;
        return nullptr;
    }
    return &api;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

#define B2S_NCCL(api, call)                                                                     \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != ncclSuccess) {                                                                 \
            set_error("%s failed: %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r_) : "?"); \
            return -8;                                                                           \
        }                                                                                        \
    } while (0)

int comm_unique_id(void* h_id128) {
    NcclApi* a = nccl();
    if (!a) return -8;
    ncclUniqueId id;
    B2S_NCCL(a, a->GetUniqueId(&id));
    memcpy(h_id128, &id, sizeof(id));
    return 0;
}

int comm_init(Comm** out, const void* h_id128, int rank, int world) {
    NcclApi* a = nccl();
    if (!a) return -8;
    ncclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    Comm* c = new Comm();
    c->rank = rank; c->world = world;
    ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", a->GetErrorString ? a->GetErrorString(r) : "?");
        delete c;
        return -8;
    }
    *out = c;
    return 0;
}

int comm_destroy(Comm* c) {
    if (!c) return 0;
    NcclApi* a = nccl();
    if (a && c->comm) a->CommDestroy(c->comm);
    delete c;
    return 0;
}

int comm_allreduce_f32(Comm* c, float* buf, long long n, cudaStream_t st) {
    NcclApi* a = nccl();
    if (!a) return -8;
    B2S_NCCL(a, a->AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, c->comm, st));
    count_launch();
    return 0;
}
int comm_allreduce_f64(Comm* c, double* buf, long long n, cudaStream_t st) {
    NcclApi* a = nccl();
    if (!a) return -8;
    B2S_NCCL(a, a->AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, c->comm, st));
    count_launch();
    return 0;
}

}  // namespace b2s
