// peer.cuh -- one-shot all-reduce of small fp64 vectors over NVLink peer memory, callable from inside a kernel.
//
// Every rank owns an exchange buffer that all peers have mapped with cudaIpc (comm.cu):
//     flags [2 parities][kPeerMax sources]            (16 B apart)
//     data  [2 parities][kPeerMax sources][kPeerCap doubles]
// peer_exchange_block() is executed by ONE thread block per rank and PUSHES: it writes its vector into slot
// (parity, own rank) of EVERY rank's buffer (posted stores through NVSwitch, nothing waits for them), fences, raises
// flag (parity, own rank) in every buffer (st.release.sys), then polls the flags of its OWN buffer -- local memory --
// and sums the slots in rank order (bitwise identical on all ranks).  The round-1 form published locally and pulled:
// every poll of a remote flag and every load of remote data was an NVLink round trip on the critical path of the
// BatchNorm layer that waits for the sums.
// A rank can be at most one call ahead of a peer -- it needs the peer's flag of the previous call to get there -- so
// two parities are enough and no slot is overwritten while a peer still reads it.
// A peer that does not arrive within B2S_PEER_TIMEOUT_S seconds (default 600) sets a host-visible error flag.
// The BatchNorm kernels call it between their statistics and apply phases (bn.cu), which makes the per-layer
// statistics all-reduce part of the kernel that produces and consumes the sums instead of a separate
// NCCL launch between two kernels.
#pragma once
#include <cuda_runtime.h>

namespace b2s {

constexpr int kPeerMax = 8;            // GPUs of one box
constexpr int kPeerCap = 4096;         // doubles per slot
constexpr int kPeerFlagBytes = 2 * kPeerMax * 16;     // [parity][source] 8-byte sequence flags, 16 bytes apart
constexpr size_t kPeerDataBytes = 2 * (size_t)kPeerMax * kPeerCap * sizeof(double);
// per-channel low-latency packets (peer_exchange_channel): [2 parities][kPeerMax sources][kPeerCap packets] of 16 bytes
constexpr size_t kPeerLLOffset = kPeerFlagBytes + kPeerDataBytes;
constexpr size_t kPeerLLBytes = 2 * (size_t)kPeerMax * kPeerCap * 16;
constexpr size_t kPeerBufBytes = kPeerLLOffset + kPeerLLBytes;

struct PeerCtx {
    double* data[kPeerMax];            // every rank's exchange data (own = local pointer), [2][kPeerMax][kPeerCap]
    unsigned long long* flag[kPeerMax];    // every rank's flags, [2][kPeerMax] (stride 2 words)
    double* own_data;
    unsigned long long* own_flag;
    unsigned long long* seq;           // device-resident call counter (graph replays advance it)
    unsigned int* ticket;              // arrival counter of the kernel whose last block performs the exchange
    int* error;                        // mapped host memory: set to 1 + peer rank when a peer did not arrive in time
    unsigned long long timeout_ns;     // B2S_PEER_TIMEOUT_S (default 600 s), measured with %globaltimer
    int rank, world;
    // per-channel low-latency exchange (peer_exchange_channel)
    uint4* ll[kPeerMax];               // every rank's packet area (own = local pointer)
    uint4* own_ll;
    unsigned int* chan_ticket;         // [kPeerCap] arrival counters: the last block of a channel publishes its sums
    unsigned int* done_ticket;         // blocks that have finished the exchange; the last one advances seq
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
    const int par = (int)(seq & 1);
    const size_t slot = ((size_t)par * kPeerMax + ctx.rank) * kPeerCap;          // where MY vector goes in every buffer
    for (int i = tid; i < n; i += nth) {
        const double v = __ldcg(buf + i);
        for (int r = 0; r < ctx.world; ++r) ctx.data[r][slot + i] = v;             // own buffer and every peer's
    }
    __threadfence_system();
    __syncthreads();
    if (tid < ctx.world) st_release_sys(ctx.flag[tid] + (par * kPeerMax + ctx.rank) * 2, seq);
    if (tid < ctx.world && tid != ctx.rank) {
        // a peer that is merely late (graph instantiation, lazy module load, a stalled data loader) is waited for;
        // one that never arrives within the (wall-clock, configurable) limit raises the host-visible error flag and
        // the exchange returns with an incomplete sum -- the host reports it at the end of the call (plan.cu leave()),
        // the context stays usable (no trap)
        const unsigned long long* f = ctx.own_flag + (par * kPeerMax + tid) * 2;      // local memory
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
    const double* mine = ctx.own_data + (size_t)par * kPeerMax * kPeerCap;
    for (int i = tid; i < n; i += nth) {
        double s = 0.0;
        for (int r = 0; r < ctx.world; ++r) s += __ldcv(mine + (size_t)r * kPeerCap + i);
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

// ---- per-channel exchange without any grid-wide barrier ------------------------------------------------------
// A (C x splits) reduction grid has added its partial sums of channel c = blockIdx.x into sums[q * C + c] (q < NQ) with
// atomics.  Every block of the grid calls this with all its threads:
//   * the block that arrives last FOR ITS CHANNEL (per-channel ticket) pushes the channel's NQ sums to every rank --
//     its own included -- as 16-byte packets {low word, tag, high word, tag} (the layout of NCCL's LL protocol: the
//     sequence tag travels inside each 8-byte half of the store, so data and "flag" are ONE NVLink traversal; the
//     data / fence / flag form above costs a round trip for the system fence before the flag can leave);
//   * every block of the channel polls the packets of all ranks in its OWN buffer (local memory) and sums them in
//     rank order: the local packet doubles as the barrier between the channel's blocks, so the cooperative BatchNorm
//     kernels need no grid.sync() at all under data parallelism, and channels proceed independently of each other.
// tot[q] = the global sums (bitwise identical on all ranks and blocks); block (c, 0) also stores them into sums[]
// for the kernels that read them later.  Spinning blocks only wait for blocks of the same channel: the kernel must be
// launched cooperatively (all blocks co-resident) unless the grid has one block per channel.
// Two parities: a rank needs every peer's packets of call k to leave call k, so it can be at most one call ahead.
__device__ __forceinline__ void st_ll(uint4* p, const double v, const unsigned int tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(tag),
                 "r"((unsigned int)(b >> 32)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ld_ll(const uint4* p, const unsigned int tag, double& v) {
    uint4 u;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
    if (u.y != tag || u.w != tag) return false;
    v = __longlong_as_double((long long)(((unsigned long long)u.z << 32) | u.x));
    return true;
}

template <int NQ>
__device__ __forceinline__ void peer_exchange_channel(double* sums, const int C, const int c, const PeerCtx& ctx, double (&tot)[NQ]) {
    __shared__ double s_val[kPeerMax * NQ];
    __shared__ int s_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const unsigned long long seq = *((volatile unsigned long long*)ctx.seq) + 1;
    const unsigned int tag = (unsigned int)seq;
    const int par = (int)(seq & 1);
    const int world = ctx.world;
    __threadfence();                    // this block's atomics are visible device-wide before its ticket
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(ctx.chan_ticket + c, 1u);
        s_last = t == gridDim.y - 1;
        if (s_last) {
            ctx.chan_ticket[c] = 0;     // kernels that use the tickets are serialised on one stream
            __threadfence();
        }
    }
    __syncthreads();
    if (s_last && tid < world * NQ) {
        const int r = tid / NQ, q = tid - r * NQ;
        const double v = __ldcg(sums + (size_t)q * C + c);
        st_ll(ctx.ll[r] + ((size_t)par * kPeerMax + ctx.rank) * kPeerCap + (size_t)q * C + c, v, tag);
    }
    if (tid < world * NQ) {
        const int r = tid / NQ, q = tid - r * NQ;
        const uint4* pk = ctx.own_ll + ((size_t)par * kPeerMax + r) * kPeerCap + (size_t)q * C + c;
        double v = 0.0;
        unsigned long long t0 = 0;
        unsigned int spins = 0;
        while (!ld_ll(pk, tag, v)) {
            if ((++spins & 0xfffu) == 0) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > ctx.timeout_ns) { *((volatile int*)ctx.error) = 1 + r; v = 0.0; break; }
            }
        }
        s_val[tid] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += s_val[r * NQ + q];
        tot[q] = s;
    }
    if (tid == 0) {
        if (blockIdx.y == 0) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) sums[(size_t)q * C + c] = tot[q];
        }
        const unsigned int d = atomicAdd(ctx.done_ticket, 1u);
        if (d == gridDim.x * gridDim.y - 1) {       // every block has read seq and finished polling
            *ctx.done_ticket = 0;
            *ctx.seq = seq;
        }
    }
}

}  // namespace b2s
