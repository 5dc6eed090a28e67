/*
 * b200_spectral.h -- C ABI of libb200spectral.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for the spectral-radius hot path of ars2240/optWBoundEigenval.
 * The reference has no FFI: the seam is the Python class `HVPOperator`
 * (opt.py:48-192) and the methods `comp_rho` / `comp_gradrho` of
 * `OptWBoundEignVal` (opt.py:418-542).  Every entry point below names the
 * reference lines whose work it replaces.  The Python side
 * (optwboundeigenval_b200/hvp_operator.py) binds these with ctypes and passes
 * raw device pointers obtained from `tensor.data_ptr()`; see INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, a negative code on failure; the
 *    message is available from b2s_last_error() (thread local);
 *  - all pointers named d_* are DEVICE pointers borrowed for the call, h_* are
 *    HOST pointers; the library owns only its plan and workspaces;
 *  - work is enqueued on the plan's stream (set with b2s_plan_set_stream, the
 *    default is the legacy default stream); only functions documented as
 *    synchronising wait for the device;
 *  - a plan is not re-entrant: one thread per plan at a time;
 *  - flat parameter vectors follow `model.parameters()` order, shared
 *    parameters once (opt.py:102,135,146,191).
 */
#ifndef B200_SPECTRAL_H
#define B200_SPECTRAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_ABI_VERSION 1

/* ---- tape description ------------------------------------------------------------ */

/* One activation tensor, NCHW inside a (possibly larger) per-sample buffer; a
 * channel-concatenation is expressed by views with different `offset` into the
 * same buffer (densenet.py:21,42 `torch.cat`). Linear activations use H=W=1. */
typedef struct {
    int32_t buf;            /* buffer id, 0 is the network input */
    int32_t C, H, W;
    int64_t offset;         /* element offset of this view inside one sample of the buffer */
    int64_t sample_stride;  /* elements between consecutive samples of the buffer */
} b2s_tensor;

enum {
    B2S_OP_CONV = 1,     /* nn.Conv2d and nn.Linear (H=W=1, 1x1 kernel) */
    B2S_OP_BN = 2,       /* train-mode nn.BatchNorm2d / BatchNorm1d */
    B2S_OP_RELU = 3,     /* standalone ReLU (normally fused into its producer) */
    B2S_OP_MAXPOOL = 4,
    B2S_OP_AVGPOOL = 5,  /* kernel == stride, no padding; adaptive (1,1) maps to kernel = H */
    B2S_OP_COPY = 6,     /* physical copy of a view (fallback for concatenations that cannot alias) */
    B2S_OP_ADD = 7       /* residual connection out = in + tensor[slot] (torchvision Bottleneck, dcnn.py:219-236) */
};

enum {
    B2S_F_RELU = 1,      /* a ReLU is fused behind this op */
    B2S_F_FIRST = 2,     /* input is network data: no input tangent, no input adjoint */
    B2S_F_BWD_ACC = 4,   /* backward accumulates into the input adjoint instead of overwriting it */
    B2S_F_BWD_ACC2 = 8   /* B2S_OP_ADD: the same for the second operand's adjoint */
};

typedef struct {
    int32_t kind;
    int32_t flags;
    int32_t in;             /* tensor id */
    int32_t out;            /* tensor id */
    int64_t w_off;          /* weight (Conv/Linear) or gamma (BN) offset in the flat parameter vector, -1 if none */
    int64_t b_off;          /* bias / beta offset, -1 if none */
    int32_t kh, kw, sh, sw, ph, pw;
    int32_t slot;           /* BN: index into the running-statistics pointer tables; MAXPOOL: argmax slot */
    float eps;              /* BN epsilon */
    float momentum;         /* BN momentum (<0: cumulative average, nn.BatchNorm2d(momentum=None)) */
} b2s_op;

/* loss heads */
enum {
    B2S_HEAD_CE = 1,            /* logits -> CrossEntropyLoss(mean)                       densenet.py:121 */
    B2S_HEAD_SOFTMAX_CE = 2,    /* logits -> softmax -> CrossEntropyLoss(mean)            forest_data.py:88, usps_data.py:335 */
    B2S_HEAD_WBCE = 3,          /* logits -> W_BCEWithLogitsLoss                          dcnn.py:375-400 */
    B2S_HEAD_SIGMOID_WBCE = 4   /* logits -> sigmoid -> W_BCEWithLogitsLoss               dcnn.py:275,375-400 */
};

typedef struct b2s_plan b2s_plan;

/* ---- library ----------------------------------------------------------------------- */
int b2s_abi_version(void);
const char* b2s_last_error(void);
/* number of kernels this library has launched (or captured graph kernel nodes replayed) since load */
int64_t b2s_launch_count(void);

/* Which contractions run on the tcgen05 / TMEM tensor-core path (3xTF32, fp32-accurate): 0 = none
 * (CUDA-core kernels only), 1 = automatic (default: layers with >= 64 destination channels), 2 = every
 * legal shape (tests).  Process-wide; set it before plans capture their graphs. */
int b2s_set_tensor_core_mode(int32_t mode);

/* ---- plan -------------------------------------------------------------------------- */
/* Compiles a tape into an execution plan and allocates its workspaces (value, tangent
 * and adjoint caches for `max_batch` samples).  buf_elems[b] = elements per sample of
 * buffer b.  logits = tensor id feeding the loss head. Replaces the autograd graph that
 * HVPOperator.prepare_grad builds with create_graph=True (opt.py:175-192). */
int b2s_plan_create(const b2s_tensor* tensors, int32_t n_tensors,
                    const int64_t* buf_elems, int32_t n_bufs,
                    const b2s_op* ops, int32_t n_ops,
                    int32_t logits, int32_t head, int64_t n_params,
                    int32_t max_batch, int32_t device, b2s_plan** out);
int b2s_plan_destroy(b2s_plan* p);
int b2s_plan_set_stream(b2s_plan* p, void* cuda_stream);
/* use_graphs != 0: each pass is captured once per batch size into a CUDA graph and replayed */
int b2s_plan_set_graphs(b2s_plan* p, int32_t use_graphs);
int64_t b2s_plan_workspace_bytes(const b2s_plan* p);
/* Third-order derivative through train-mode BatchNorm used by b2s_vghv.
 * exact = 0 (default): reproduce what nested torch.autograd returns for HVPOperator.vGHv
 *   (opt.py:132-143).  torch's batchnorm_double_backward reads the batch mean / inverse std from
 *   saved non-differentiable tensors, so its third sweep drops every path through them; the library
 *   computes the exact result and subtracts exactly those dropped terms (DESIGN.md "BatchNorm").
 * exact = 1: the true gradient of v^T H v (agrees with finite differences of the gradient). */
int b2s_plan_set_bn_third_order(b2s_plan* p, int32_t exact);
/* device pointers (float*) of the running_mean / running_var tensors of BN slot `slot`
 * (updated in place by the base pass exactly like a train-mode forward, opt.py:181,421) */
int b2s_plan_set_bn_buffers(b2s_plan* p, int32_t slot, void* d_running_mean, void* d_running_var);

/* ---- the three passes ---------------------------------------------------------------- */
/* Base pass = HVPOperator.prepare_grad (opt.py:175-192): forward, loss, gradient.
 * d_params: fp32 [n_params]; d_x: fp32 [batch, ...]; d_y: int64 [batch] class labels
 * (CE heads) or fp32 [batch, C] targets with d_coef fp32 [batch, C] per-entry weights
 * already divided by the valid counts (WBCE heads; d_coef NULL for CE).
 * loss_scale multiplies the loss (1/global_batch for mean-reduced CE).
 * d_grad_out: fp64 [n_params] or NULL.  Asynchronous. */
int b2s_base_pass(b2s_plan* p, const float* d_params, const float* d_x, const void* d_y,
                  const float* d_coef, int32_t batch, double loss_scale,
                  double* d_grad_out, double* d_loss_out);
/* H*v = HVPOperator.Hv(vec, storedGrad=True) (opt.py:77-108). d_v fp64 [n_params]
 * (rounded to fp32 as the reference's cast does), d_out fp64 [n_params]. Asynchronous. */
int b2s_hv(b2s_plan* p, const double* d_v, double* d_out);
/* comp_f (opt.py:544-572) and, through it, every forward pass of test_model (opt.py:954): forward only, BatchNorm in
 * evaluation mode (running statistics read, not updated), loss of the head into d_loss_out (fp64 [1]) and the raw
 * logits [batch, classes] into d_logits_out (optional; a softmax / sigmoid tail of the model is the caller's).
 * Arguments as b2s_base_pass.  Local to the calling rank (no collective); forgets the cached base pass. */
int b2s_eval_pass(b2s_plan* p, const float* d_params, const float* d_x, const void* d_y, const float* d_coef,
                  int32_t batch, double loss_scale, float* d_logits_out, double* d_loss_out);
/* grad_w (v^T H v) = HVPOperator.vGHv (opt.py:110-152). Asynchronous. */
int b2s_vghv(b2s_plan* p, const double* d_v, double* d_out);
/* fp32 results of the last passes, device pointers owned by the plan (float [n_params]) */
const float* b2s_grad_f32(const b2s_plan* p);
const float* b2s_hv_f32(const b2s_plan* p);

/* test hook: copies one cached tensor (adjoint = 0: value jets, 1: adjoint jets; order 0..2) of the
 * last passes to host memory, densely packed [batch, C, H, W]. Synchronises. */
int b2s_debug_read(b2s_plan* p, int32_t adjoint, int32_t order, int32_t tensor, float* h_out);

/* Built-in per-launch timer: runs pass `order` (0 base, 1 Hv, 2 second order, 3 BatchNorm
 * compatibility sweep) eagerly `reps` times with a CUDA-event pair around every kernel on the
 * launching stream and returns, per kernel family, launches / device milliseconds / algorithmic
 * FLOPs and bytes PER PASS.  bench.py derives the roofline line from this. Synchronises. */
typedef struct {
    char name[48];
    int32_t launches;
    double ms;
    double flops;
    double bytes;
} b2s_prof_entry;
int b2s_profile_pass(b2s_plan* p, int32_t order, int32_t reps, b2s_prof_entry* out, int32_t cap, int32_t* n_out);

/* ---- spectral-radius iteration ------------------------------------------------------- */
typedef struct {
    int32_t max_iter;        /* min(ndim, max_pow_iter)                                     opt.py:447 */
    double eps;              /* pow_iter_eps                                                opt.py:480 */
    const double* h_alpha;   /* host array [max_iter] of relaxation factors, NULL = all 1   opt.py:489 */
    int32_t precond;         /* 1: v_new = v + alpha * T(r) with the K-FAC map installed    opt.py:491-493 */
} b2s_power_cfg;

typedef struct {
    int32_t iters;           /* index i of the last iteration, as comp_rho returns it       opt.py:533 */
    int32_t converged;       /* 0 when all three stopping values stayed above eps           opt.py:513 */
    double lam, norm, rn, vnn;
    double stop[3];
} b2s_power_result;

/* comp_rho's loop (opt.py:447-498) on the device: HVP + fused vector kernels per
 * iteration, 4 scalars read back per iteration for the stopping test.  d_v fp64
 * [n_params] in/out: start vector on entry, on exit the vector self.v would hold
 * (opt.py:508).  h_traj: NULL or host array [max_iter*5] receiving the rows
 * (i, lam, n, rn, vnn) of the verbose log (opt.py:466).  Synchronises. */
int b2s_power_iterate(b2s_plan* p, double* d_v, const b2s_power_cfg* cfg,
                      b2s_power_result* h_result, double* h_traj);

/* ---- the vector part of the iteration on its own -----------------------------------------
 * A b2s_pistate holds the fp64 eigenvector / residual double buffers and every scalar of the
 * loop on the device.  b2s_power_iterate drives one internally; it is exposed so the two fused
 * vector kernels can be measured against the HBM roofline at any length n and parity-tested
 * against opt.py:455-498 without a network. */
typedef struct b2s_pistate b2s_pistate;
int b2s_pi_create(int64_t n, int32_t max_iter_capacity, int32_t device, b2s_pistate** out);
int b2s_pi_destroy(b2s_pistate* s);
/* start a run: copies d_v0 (fp64 [n]) in, rounds it to fp32, clears lam_old / r_old (opt.py:436) */
int b2s_pi_reset(b2s_pistate* s, const double* d_v0, const b2s_power_cfg* cfg, void* stream);
/* fp32 rounding of the current vector: the input of the next HVP (device pointer, fixed for the
 * lifetime of the state) */
const float* b2s_pi_v32(const b2s_pistate* s);
/* vector work of one iteration given d_hv = H*v (fp32 [n]): pass A + pass B. Asynchronous. */
int b2s_pi_step(b2s_pistate* s, const float* d_hv, void* stream);
/* preconditioned variant only (cfg.precond = 1): after b2s_pi_step, d_r = b2s_pi_residual() holds
 * r; the caller maps it through T and finishes the update v <- normalise(v + alpha*Tr). */
const double* b2s_pi_residual(b2s_pistate* s, void* stream);   /* synchronises */
int b2s_pi_precond_update(b2s_pistate* s, const double* d_Tr, void* stream);
/* 1 once a stopping test fired or the iterations ran out (synchronises `stream`) */
int b2s_pi_done(b2s_pistate* s, void* stream);
/* reads the scalars back (synchronises `stream`); d_v_out (fp64 [n]) receives the vector self.v
 * would hold, may be NULL; h_traj [iters+1][5] may be NULL */
int b2s_pi_result(b2s_pistate* s, b2s_power_result* h_result, double* h_traj, double* d_v_out,
                  void* stream);

/* ---- K-FAC preconditioner of the `lobpcg=True` variant (opt.py:362-416) ---------------------
 * Layers are named by the index of their Conv/Linear op in the tape.  b2s_kfac_build forms, from the
 * cached base pass of the current batch, A = 0.95 I + 0.05 E[a a^T] of op_a's input patches (bias
 * column of ones; kfac.py:292-311,52-58) and G = 0.95 I + 0.05 B*S*sum(g g^T) of op_g's output
 * adjoint (kfac.py:337-367,60-65) into caller-provided device buffers [da,da] / [dg,dg] (op_a / op_g
 * differ only for a module applied twice: last forward use for A, first for G, as the reference's
 * hooks leave them).  The host runs eigh on them (kfac.py:87-93) and installs the two inverses
 * Q diag(1/d) Q^T (borrowed device pointers, row-major fp32) with b2s_kfac_set.  b2s_kfac_apply maps
 * r -> T r = G^-1 R A^-1 per layer, identity elsewhere (opt.py:384-416, kfac.py:118-120);
 * b2s_power_iterate with cfg.precond = 1 uses it for v <- normalise(v + alpha T r) (opt.py:491-498). */
int b2s_kfac_dims(const b2s_plan* p, int32_t op_index, int32_t* dim_a, int32_t* dim_g);
int b2s_kfac_build(b2s_plan* p, int32_t op_a, int32_t op_g, float* d_A, float* d_G);
int b2s_kfac_clear(b2s_plan* p);
int b2s_kfac_set(b2s_plan* p, int32_t op_index, const float* d_Ainv, const float* d_Ginv);
int b2s_kfac_apply(b2s_plan* p, const double* d_r, double* d_out);

/* ---- step assembly of the regularised minibatch step (opt.py:616-639 assembly, 654-659 scatter) ----------
 * p = grad f + coef * grad rho with coef = mu * sign (opt.py:631-639; d_gradrho NULL when g == 0, opt.py:636),
 * written as fp64 (d_p, optional) and as the fp32 flat vector d_p32 whose slices in model.parameters() order become
 * param.grad (the reference's per-parameter `p[i:i+n].view(s).float()` loop, opt.py:654-659).  All vectors device
 * resident, 16-byte aligned, enqueued on `stream`. */
int b2s_step_assemble(const double* d_gradf, const double* d_gradrho, double coef, int64_t n, double* d_p, float* d_p32,
                      void* stream);

/* The remainder of iter()'s minibatch body fused on flat vectors (opt.py:535-542 clip, 616-659 assembly, 696-699
 * optimizer step).  b2s_clip_norm: d_out2 = {|x|, clip > 0 && |x| > clip ? clip / |x| : 1} without a host sync
 * (d_scratch: b2s_clip_scratch_doubles() doubles, zeroed once by the caller).  b2s_step_fused over n elements:
 *   p = grad f + coef * s * grad rho          (s = d_scale2[1] when given; grad rho itself is rescaled in place when
 *                                              write_gradrho, as the reference's `self.gradrho *= clip / grn`)
 *   d_p (fp64, optional), d_p32 = (float)p    (param.grad)
 *   kind 1: torch.optim.SGD update, kind 2: torch.optim.Adam update of d_params with state vectors d_state1
 *   (momentum buffer / exp_avg) and d_state2 (exp_avg_sq), fp32 arithmetic in torch's operation order. */
typedef struct {
    int32_t kind;            /* 0 assemble only, 1 SGD, 2 Adam */
    int32_t first_step;      /* SGD: this step initialises the momentum buffer (buf = grad) */
    int32_t nesterov, maximize, write_gradrho;
    double lr, momentum, dampening, weight_decay;
    double beta1, beta2, eps;
    double step_size;        /* Adam: lr / (1 - beta1^t) */
    double bias2_sqrt;       /* Adam: sqrt(1 - beta2^t) */
} b2s_step_opt;
int b2s_clip_norm(const double* d_x, int64_t n, double clip, double* d_scratch, double* d_out2, void* stream);
int b2s_clip_scratch_doubles(void);
int b2s_step_fused(const double* d_gradf, double* d_gradrho, double coef, const double* d_scale2, int64_t n, double* d_p,
                   float* d_p32, float* d_params, float* d_state1, float* d_state2, const b2s_step_opt* opt, void* stream);

/* ---- data parallelism (one process per GPU) ------------------------------------------- */
int b2s_comm_unique_id(void* h_id128);                       /* 128 bytes, rank 0 */
int b2s_comm_init(b2s_plan* p, const void* h_id128, int32_t rank, int32_t world);
int b2s_comm_destroy(b2s_plan* p);
/* Ragged shards: the number of samples over ALL ranks of the next base pass (0 = batch x world).  The loss scale
 * handed to b2s_base_pass must be 1 / global_batch; BatchNorm statistics and the K-FAC factors divide by it. */
int b2s_plan_set_global_batch(b2s_plan* p, int64_t global_batch);
/* One-shot all-reduce of the small per-layer BatchNorm sums over NVLink peer memory instead of NCCL (the reference
 * has no counterpart: it is single-device).  After b2s_comm_init every rank calls
 * b2s_comm_peer_local (allocates its exchange buffer, returns the 64-byte cudaIpcMemHandle), the handles are
 * all-gathered by the host side, and b2s_comm_peer_attach maps the peers' buffers.  Optional: without it
 * (or when peer access is unavailable) every all-reduce stays on NCCL. */
int b2s_comm_peer_local(b2s_plan* p, void* h_handle64);
int b2s_comm_peer_attach(b2s_plan* p, const void* h_handles /* world x 64 bytes, rank order */);
/* 1 when this rank mapped every peer.  The ranks must agree: the host side all-reduces (min) the flags and calls
 * b2s_comm_peer_disable everywhere when one rank could not attach (mixed NCCL / peer ranks would dead-lock). */
int b2s_comm_peer_ready(const b2s_plan* p);
int b2s_comm_peer_disable(b2s_plan* p);

#ifdef __cplusplus
}
#endif
#endif /* B200_SPECTRAL_H */
