// comm.cu -- thin run-time binding to NCCL (see comm.h).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "comm.h"
#include "common.cuh"
#include "peer.cuh"

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
    void* xbuf = nullptr;              // [flags][2 parities x kPeerMax sources x kPeerCap doubles] (peer.cuh)
    void* peer_base[kPeerMax] = {};
    unsigned long long* d_seq = nullptr;
    unsigned int* d_ticket = nullptr;
    unsigned int* d_chan_ticket = nullptr;   // [kPeerCap + 1]: per-channel tickets, then the done ticket (peer_exchange_channel)
    int* h_error = nullptr;            // mapped pinned host memory (peer.cuh: a peer did not arrive in time)
    bool peers_ready = false;
    PeerCtx ctx = {};
    PeerCtx* d_ctx = nullptr;          // device copy for kernels that do the exchange themselves (bn.cu)
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
    for (int r = 0; r < kPeerMax; ++r)
        if (c->peer_base[r] && r != c->rank) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->xbuf) cudaFree(c->xbuf);
    if (c->d_seq) cudaFree(c->d_seq);
    if (c->d_ticket) cudaFree(c->d_ticket);
    if (c->d_chan_ticket) cudaFree(c->d_chan_ticket);
    if (c->h_error) cudaFreeHost(c->h_error);
    if (c->d_ctx) cudaFree(c->d_ctx);
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
// ---- one-shot peer all-reduce ------------------------------------------------------------------------
__global__ void __launch_bounds__(512) peer_allreduce_kernel(double* __restrict__ buf, const int n, const PeerCtx ctx) {
    peer_exchange_block(buf, n, ctx);
}

int comm_peer_local(Comm* c, void* h_handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!c->xbuf) {
        const size_t bytes = kPeerBufBytes;
        B2S_CUDA(cudaMalloc(&c->xbuf, bytes));
        B2S_CUDA(cudaMemset(c->xbuf, 0, bytes));
        B2S_CUDA(cudaMalloc(&c->d_seq, sizeof(unsigned long long)));
        B2S_CUDA(cudaMemset(c->d_seq, 0, sizeof(unsigned long long)));
        B2S_CUDA(cudaMalloc(&c->d_ticket, sizeof(unsigned int)));
        B2S_CUDA(cudaMemset(c->d_ticket, 0, sizeof(unsigned int)));
        B2S_CUDA(cudaMalloc(&c->d_chan_ticket, (kPeerCap + 1) * sizeof(unsigned int)));
        B2S_CUDA(cudaMemset(c->d_chan_ticket, 0, (kPeerCap + 1) * sizeof(unsigned int)));
        B2S_CUDA(cudaHostAlloc(&c->h_error, sizeof(int), cudaHostAllocMapped));
        *c->h_error = 0;
        B2S_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    B2S_CUDA(cudaIpcGetMemHandle(&h, c->xbuf));
    memcpy(h_handle64, &h, sizeof(h));
    return 0;
}

int comm_peer_attach(Comm* c, const void* h_handles) {
    if (!c->xbuf) { set_error("comm_peer_attach: call comm_peer_local first"); return -1; }
    if (c->world > kPeerMax) return 0;                          // larger worlds stay on NCCL
    const char* hs = static_cast<const char*>(h_handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->peer_base[r] = c->xbuf; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)r * sizeof(h), sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            // no peer access between these devices: keep NCCL for everything
            fprintf(stderr, "b2s: cudaIpcOpenMemHandle(rank %d) failed (%s); small all-reduces stay on NCCL\n", r, cudaGetErrorString(e));
            cudaGetLastError();
            c->peer_base[r] = nullptr;
            return 0;
        }
    }
    PeerCtx& x = c->ctx;
    for (int r = 0; r < c->world; ++r) {
        x.flag[r] = reinterpret_cast<unsigned long long*>(c->peer_base[r]);
        x.data[r] = reinterpret_cast<double*>(static_cast<char*>(c->peer_base[r]) + kPeerFlagBytes);
        x.ll[r] = reinterpret_cast<uint4*>(static_cast<char*>(c->peer_base[r]) + kPeerLLOffset);
    }
    x.own_ll = reinterpret_cast<uint4*>(static_cast<char*>(c->xbuf) + kPeerLLOffset);
    x.chan_ticket = c->d_chan_ticket;
    x.done_ticket = c->d_chan_ticket + kPeerCap;
    x.own_flag = reinterpret_cast<unsigned long long*>(c->xbuf);
    x.own_data = reinterpret_cast<double*>(static_cast<char*>(c->xbuf) + kPeerFlagBytes);
    x.seq = c->d_seq;
    x.ticket = c->d_ticket;
    {
        int* dev_err = nullptr;
        B2S_CUDA(cudaHostGetDevicePointer(&dev_err, c->h_error, 0));
        x.error = dev_err;
        const char* e = getenv("B2S_PEER_TIMEOUT_S");
        const double secs = e ? atof(e) : 600.0;
        x.timeout_ns = (unsigned long long)((secs > 0 ? secs : 600.0) * 1e9);
    }
    x.rank = c->rank; x.world = c->world;
    if (!c->d_ctx) B2S_CUDA(cudaMalloc(&c->d_ctx, sizeof(PeerCtx)));
    B2S_CUDA(cudaMemcpy(c->d_ctx, &x, sizeof(PeerCtx), cudaMemcpyHostToDevice));
    static const bool off = getenv("B2S_PEER_AR") && atoi(getenv("B2S_PEER_AR")) == 0;
    c->peers_ready = !off;
    return 0;
}

// ---- single-GPU form of the per-channel exchange ----------------------------------------------------------
// world = 1: the "exchange" is the channel's last block writing two packets into local memory and the channel's other
// blocks polling them -- a per-channel barrier in place of the grid-wide one of the cooperative BatchNorm kernels.
struct LocalPeer {
    void* ll = nullptr;
    unsigned long long* seq = nullptr;
    unsigned int* tickets = nullptr;
    int* h_error = nullptr;
    PeerCtx* d_ctx = nullptr;
};
int local_peer_create(LocalPeer** out) {
    LocalPeer* l = new LocalPeer();
    *out = l;
    B2S_CUDA(cudaMalloc(&l->ll, kPeerLLBytes));
    B2S_CUDA(cudaMemset(l->ll, 0, kPeerLLBytes));
    B2S_CUDA(cudaMalloc(&l->seq, sizeof(unsigned long long)));
    B2S_CUDA(cudaMemset(l->seq, 0, sizeof(unsigned long long)));
    B2S_CUDA(cudaMalloc(&l->tickets, (kPeerCap + 1) * sizeof(unsigned int)));
    B2S_CUDA(cudaMemset(l->tickets, 0, (kPeerCap + 1) * sizeof(unsigned int)));
    B2S_CUDA(cudaHostAlloc(&l->h_error, sizeof(int), cudaHostAllocMapped));
    *l->h_error = 0;
    PeerCtx x = {};
    x.ll[0] = x.own_ll = static_cast<uint4*>(l->ll);
    x.seq = l->seq;
    x.chan_ticket = l->tickets;
    x.done_ticket = l->tickets + kPeerCap;
    int* dev_err = nullptr;
    B2S_CUDA(cudaHostGetDevicePointer(&dev_err, l->h_error, 0));
    x.error = dev_err;
    x.timeout_ns = 600ull * 1000000000ull;
    x.rank = 0; x.world = 1;
    B2S_CUDA(cudaMalloc(&l->d_ctx, sizeof(PeerCtx)));
    B2S_CUDA(cudaMemcpy(l->d_ctx, &x, sizeof(PeerCtx), cudaMemcpyHostToDevice));
    B2S_CUDA(cudaDeviceSynchronize());
    return 0;
}
void local_peer_destroy(LocalPeer* l) {
    if (!l) return;
    cudaFree(l->ll); cudaFree(l->seq); cudaFree(l->tickets); cudaFree(l->d_ctx);
    if (l->h_error) cudaFreeHost(l->h_error);
    delete l;
}
const PeerCtx* local_peer_ctx(const LocalPeer* l) { return l ? l->d_ctx : nullptr; }

int comm_peer_ready(const Comm* c) { return (c && c->peers_ready) ? 1 : 0; }
int comm_peer_error(Comm* c) {
    if (!c || !c->h_error) return 0;
    const int e = *((volatile int*)c->h_error);
    if (e) *c->h_error = 0;
    return e;
}
void comm_peer_disable(Comm* c) { if (c) c->peers_ready = false; }

// Default: the per-channel packet exchange inside the fused BatchNorm kernels (peer_exchange_channel: no grid barrier,
// one NVLink traversal per layer; B2S_PEER_LL=0 switches it off).  The older in-kernel form -- block 0 exchanges
// between two grid barriers, B2S_PEER_FUSED=1 -- idles the whole grid while one block talks to the peers (measured at
// 2 GPUs, DenseNet3: 3.60 ms per step against 3.33 ms for the separate one-CTA exchange kernel and 3.35 ms for NCCL).
int comm_peer_ll(const Comm* c) {
    static const bool on = !(getenv("B2S_PEER_LL") && atoi(getenv("B2S_PEER_LL")) == 0);
    return (c && c->peers_ready && on) ? 1 : 0;
}
const PeerCtx* comm_peer_ctx(Comm* c) {
    static const bool fused = getenv("B2S_PEER_FUSED") && atoi(getenv("B2S_PEER_FUSED")) != 0;
    return (c && c->peers_ready && (fused || comm_peer_ll(c))) ? c->d_ctx : nullptr;
}

const PeerCtx* comm_peer_tail_ctx(Comm* c) {
    // measured at 2 GPUs: 3.39 ms per step with the exchange in the statistics kernel's tail, 3.36 ms with the
    // separate one-block kernel -- no gain, so the simpler form is the default.  Opt-in: B2S_PEER_TAIL=1.
    static const bool on = getenv("B2S_PEER_TAIL") && atoi(getenv("B2S_PEER_TAIL")) != 0;
    return (c && c->peers_ready && on) ? c->d_ctx : nullptr;
}

int comm_allreduce_f64(Comm* c, double* buf, long long n, cudaStream_t st) {
    if (c->peers_ready && n <= kPeerCap) {
        peer_allreduce_kernel<<<1, n > 256 ? 512 : 256, 0, st>>>(buf, (int)n, c->ctx);
        B2S_LAUNCH_CHECK();
        return 0;
    }
    NcclApi* a = nccl();
    if (!a) return -8;
    B2S_NCCL(a, a->AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, c->comm, st));
    count_launch();
    return 0;
}

}  // namespace b2s
