// comm.h -- NCCL data parallelism for the sharded minibatch (internal).
// The reference is single-process / single-device; sharding the minibatch over
// the GPUs of one box needs one all-reduce(sum) of the P-vector per pass and, for train-mode
// BatchNorm, one small all-reduce of the per-channel sums per BN layer and direction
// (SURVEY.md 8e).  NCCL is resolved at run time from the already loaded libnccl.so.2
// (the one PyTorch ships), so the library has no link-time dependency on it.
#pragma once
#include <cuda_runtime.h>

namespace b2s {
struct Comm;
int comm_unique_id(void* h_id128);
int comm_init(Comm** out, const void* h_id128, int rank, int world);
int comm_destroy(Comm* c);
int comm_allreduce_f32(Comm* c, float* buf, long long n, cudaStream_t st);
int comm_allreduce_f64(Comm* c, double* buf, long long n, cudaStream_t st);
// One-shot all-reduce of small fp64 vectors over NVLink peer memory (the per-layer BatchNorm sums: 2C doubles,
// latency-bound): every rank publishes its vector in an exchange buffer that the peers have mapped
// (cudaIpc), raises a sequence flag, polls the peers' flags and sums the peers' vectors in rank order with
// direct loads through NVSwitch -- one small kernel, no ring, no proxy thread.  comm_allreduce_f64 uses it
// for n <= the exchange capacity once the peers are attached, NCCL otherwise.
int comm_peer_local(Comm* c, void* h_handle64);                 // allocate the exchange buffer, return its IPC handle
int comm_peer_attach(Comm* c, const void* h_handles);           // world x 64 bytes, rank order
// whether this rank mapped every peer; all ranks must agree before the peer path is used (hvp_operator.init_comm
// all-reduces the flags and disables the path everywhere when one rank could not attach)
int comm_peer_ready(const Comm* c);
void comm_peer_disable(Comm* c);
// 0, or 1 + the rank of a peer that did not reach an exchange within B2S_PEER_TIMEOUT_S (reading clears it)
int comm_peer_error(Comm* c);
struct PeerCtx;
// device-resident exchange context for kernels that do the exchange themselves (peer.cuh), NULL when unavailable
const PeerCtx* comm_peer_ctx(Comm* c);
// 1 when the in-kernel exchange is the per-channel packet form (peer.cuh: peer_exchange_channel), the default
int comm_peer_ll(const Comm* c);
// ... for reduction kernels whose last block performs the exchange (peer_exchange_tail), NULL when unavailable
const PeerCtx* comm_peer_tail_ctx(Comm* c);
// single-GPU form of the per-channel exchange: a per-channel barrier inside the cooperative BatchNorm kernels
struct LocalPeer;
int local_peer_create(LocalPeer** out);
void local_peer_destroy(LocalPeer* l);
const PeerCtx* local_peer_ctx(const LocalPeer* l);
}  // namespace b2s
