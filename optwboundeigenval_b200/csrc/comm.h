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
}  // namespace b2s
