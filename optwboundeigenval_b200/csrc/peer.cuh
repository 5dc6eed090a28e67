// peer.cuh -- one-shot all-reduce of small fp64 vectors over NVLink peer memory, callable from inside a kernel.
//
// Every rank owns an exchange buffer [2 sequence flags, 128 B apart][2 slots x kPeerCap doubles] that all peers
// have mapped with cudaIpc (comm.cu).  peer_exchange_block() is executed by ONE thread block per rank:
// publish the vector in slot (call number & 1), raise the flag (st.release.sys), poll the peers' flags
// (ld.acquire.sys through NVSwitch), add the peers' vectors in rank order (bitwise identical on all ranks).
// A peer that does not arrive within B2S_PEER_TIMEOUT_S seconds (default 600) sets a host-visible error flag.
// A rank can be at most one call ahead of a peer -- it needs the peer's flag of the previous call to get
// there -- so two slots are enough and no slot is overwritten while a peer still reads it.
// The BatchNorm kernels call it between their statistics and apply phases (bn.cu), which makes the per-layer
// statistics all-reduce part of the kernel that produces and consumes the sums instead of a separate
// NCCL launch between two kernels.
#pragma once
#include <cuda_runtime.h>

namespace b2s {

constexpr int kPeerMax = 8;            // GPUs of one box
constexpr int kPeerCap = 4096;         // doubles per slot
constexpr int kPeerFlagBytes = 256;    // two 8-byte sequence flags, 128 bytes apart

struct PeerCtx {
    const double* data[kPeerMax];      // every rank's exchange data (own = local pointer), [2 slots][kPeerCap]
    const unsigned long long* flag[kPeerMax];
    double* own_data;
    unsigned long long* own_flag;
    unsigned long long* seq;           // device-resident call counter (graph replays advance it)
    unsigned int* ticket;              // arrival counter of the kernel whose last block performs the exchange
    int* error;                        // mapped host memory: set to 1 + peer rank when a peer did not arrive in time
    unsigned long long timeout_ns;     // B2S_PEER_TIMEOUT_S (default 600 s), measured with %globaltimer
    int rank, world;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// buf[0..n) <- sum over ranks of buf[0..n); all threads of the calling block take part (n <= kPeerCap)
__device__ __forceinline__ void peer_exchange_block(double* buf, const int n, const PeerCtx& ctx) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x, nth = blockDim.x * blockDim.y;
    const unsigned long long seq = *((volatile unsigned long long*)ctx.seq) + 1;
    const int slot = (int)(seq & 1);
    double* mine = ctx.own_data + (size_t)slot * kPeerCap;
    for (int i = tid; i < n; i += nth) mine[i] = __ldcg(buf + i);
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_release_sys(ctx.own_flag + slot * 16, seq);
    if (tid < ctx.world && tid != ctx.rank) {
        const unsigned long long* f = ctx.flag[tid] + slot * 16;
        // a peer that is merely late (graph instantiation, lazy module load, a stalled data loader) is waited for;
        // one that never arrives within the (wall-clock, configurable) limit raises the host-visible error flag and
        // the exchange returns with an incomplete sum -- the host reports it at the end of the call (plan.cu leave()),
        // the context stays usable (no trap)
        unsigned long long t0 = 0;
        unsigned int spins = 0;
        while (ld_acquire_sys(f) < seq) {
            if ((++spins & 0xfffu) == 0) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > ctx.timeout_ns) { *((volatile int*)ctx.error) = 1 + tid; break; }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nth) {
        double s = 0.0;
        for (int r = 0; r < ctx.world; ++r) s += __ldcv(ctx.data[r] + (size_t)slot * kPeerCap + i);
        buf[i] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *ctx.seq = seq;
}

// Tail of a reduction kernel whose blocks have just added their partial sums into `sums` with atomics: the block
// that arrives last (ticket) all-reduces the sums over the GPUs before the kernel ends, so the statistics
// all-reduce costs no launch of its own.  Every thread of every block must call it.
__device__ __forceinline__ void peer_exchange_tail(double* sums, const int n, const PeerCtx& ctx) {
    __shared__ int is_last;
    __threadfence();                    // this block's atomics are visible device-wide before its ticket
    __syncthreads();
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    if (tid == 0) {
        const unsigned int t = atomicAdd(ctx.ticket, 1u);
        is_last = t == gridDim.x * gridDim.y * gridDim.z - 1;
    }
    __syncthreads();
    if (is_last) {
        if (tid == 0) *ctx.ticket = 0;  // kernels that use the ticket are serialised on one stream
        __threadfence();
        peer_exchange_block(sums, n, ctx);
    }
}

}  // namespace b2s
