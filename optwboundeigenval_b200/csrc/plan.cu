// plan.cu -- tape interpreter + C ABI of libb200spectral.
//
// A plan is the static replacement of the autograd graph that the reference rebuilds for every
// minibatch (HVPOperator.prepare_grad with create_graph=True, opt.py:175-192): the tape of layer
// ops, the flat-parameter offsets and preallocated caches for every tensor's value jets
// (x, xdot, xddot) and adjoint jets (xbar, R xbar, R^2 xbar).  One "pass of order k" walks the
// tape forward then backward propagating the order-k components (rop.py:69-164 generalised):
//     order 0 -> loss and gradient         (prepare_grad,        opt.py:175-192)
//     order 1 -> Hessian-vector product    (Hv,                  opt.py:77-108)
//     order 2 -> grad_w (v^T H v)          (second half of vGHv, opt.py:110-152)
// Each pass is a fixed kernel sequence, captured once per batch size into a CUDA graph and
// replayed (one graph launch per HVP instead of the reference's per-op autograd dispatch).
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200_spectral.h"
#include "comm.h"
#include "conv_args.h"
#include "kernels.h"
#include "vec.h"

namespace b2s {

// ---- error / counters ---------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static long long g_launch_total = 0;
static std::mutex g_launch_mu;
thread_local long long g_launches = 0;
static thread_local bool g_counting_paused = false;
static thread_local long long g_capture_count = 0;

bool skip_family(const char* name) {
    static const char* list = getenv("B2S_SKIP");
    if (!list) return false;
    const char* hit = strstr(list, name);
    if (!hit) return false;
    const char end = hit[strlen(name)];
    return (hit == list || hit[-1] == ',') && (end == 0 || end == ',');
}
bool pdl_enabled() {
    static const bool on = getenv("B2S_PDL") && atoi(getenv("B2S_PDL")) != 0;   // measured: no gain inside captured graphs, off by default
    return on;
}
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
struct ProfRec {
    const char* name;
    double flops, bytes;
    cudaEvent_t e0, e1;
};
struct Profiler {
    std::vector<ProfRec> recs;
};
static thread_local Profiler* g_prof = nullptr;

ProfScope::ProfScope(const char* name, double flops, double bytes, cudaStream_t stream) : idx(-1), st(stream) {
    if (!g_prof) return;
    ProfRec r{name, flops, bytes, nullptr, nullptr};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    g_prof->recs.push_back(r);
    idx = (int)g_prof->recs.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0 && g_prof) cudaEventRecord(g_prof->recs[idx].e1, st);
}

void count_launch(int n) {
    g_capture_count += n;
    if (g_counting_paused) return;
    std::lock_guard<std::mutex> lk(g_launch_mu);
    g_launch_total += n;
}

}  // namespace b2s

using namespace b2s;

struct b2s_pistate {
    int device = 0;
    long long n = 0;
    int cap = 0;
    PiDev h{};               // host mirror used to initialise the device copy
    PiDev* d = nullptr;
    double* vbuf[2] = {nullptr, nullptr};
    double* rbuf[2] = {nullptr, nullptr};
    float* v32 = nullptr;
    double* alpha = nullptr;
    double* traj = nullptr;
    double* scratch = nullptr;
    bool have_alpha = false;
    bool own_v32 = true;     // false: v32 aliases the owning plan's tangent-direction buffer
};

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    long long kernels = 0;
    double loss_scale = 0.0;          // baked into the captured head kernels
    long long global_batch = 0;       // baked into the captured BatchNorm kernels
};

struct b2s_plan {
    int device = 0;
    cudaStream_t caller = nullptr;    // stream the caller's tensors are ordered on
    cudaStream_t stream = nullptr;    // plan-owned work stream (graphs cannot be captured on stream 0)
    static constexpr int kMaxSide = 8;
    cudaStream_t sides[kMaxSide] = {};   // weight-gradient contractions run here (round robin), off the adjoint critical path
    cudaEvent_t ev_joins[kMaxSide] = {};
    int n_side = 4;                   // B2S_SIDE_STREAMS: independent weight-gradient kernels of different layers overlap
    int side_next = 0;
    unsigned side_mask = 0;           // side streams with work in flight in this pass
    cudaStream_t side = nullptr;      // the side stream chosen by the last fork_side()
    cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_fork = nullptr;
    bool side_used = false;
    bool use_graphs = true;
    std::vector<b2s_tensor> tensors;
    std::vector<b2s_op> ops;
    std::vector<long long> buf_elems;
    std::vector<long long> buf_off;   // element offset of each buffer inside an arena (per max_batch)
    long long arena_elems = 0;
    int logits = 0, head = 0;
    long long P = 0;
    int max_batch = 0;
    int batch = 0;                    // batch of the cached base pass
    double loss_scale = 1.0;
    long long workspace = 0;

    float* fw[3] = {nullptr, nullptr, nullptr};   // value jets
    float* bw[3] = {nullptr, nullptr, nullptr};   // adjoint jets
    float* params = nullptr;
    float* v32 = nullptr;             // tangent direction (fp32 rounding of v)
    float* out32[3] = {nullptr, nullptr, nullptr};
    double* loss = nullptr;

    // batch norm
    int n_bn = 0;
    std::vector<long long> bn_off;    // per op index: offset (in doubles) of its [2][C] block, -1 if not BN
    long long bn_total = 0;
    double* fsum[3] = {nullptr, nullptr, nullptr};
    double* bsum[3] = {nullptr, nullptr, nullptr};
    double* csum = nullptr;           // sums of the third-order compatibility sweep
    float* out_corr = nullptr;        // its parameter-space result
    bool bn_exact_third_order = false;
    bool bn_fused = true;             // statistics+apply in one cooperative launch (single GPU)
    bool wgrad_side = true;           // weight-gradient contractions on the side stream
    std::vector<float*> bn_rm, bn_rv;

    // max pool
    std::vector<int32_t*> argmax;     // per op index

    // labels
    long long* labels = nullptr;
    float* target = nullptr;
    float* coef = nullptr;
    int n_classes = 0;

    std::map<long long, GraphEntry> graphs;   // key = order * 2^32 + batch
    bool v_is_current = false;        // order-1 caches correspond to the contents of v32

    struct KfacLayer {
        int op = -1;                  // op whose parameter slots / geometry the layer uses
        int dg = 0, da = 0, dw = 0;
        const float* Ainv = nullptr;  // [da,da] borrowed from the caller
        const float* Ginv = nullptr;  // [dg,dg]
    };
    std::vector<KfacLayer> kfac;
    float* kfac_m = nullptr;          // scratch [dg,da] x2
    long long kfac_m_elems = 0;
    double* kfac_tr = nullptr;        // T r

    // tcgen05 path: packed 3xTF32 weight images (conv_tc.cu)
    std::vector<long long> tc_off[2];   // per op index and mode: float offset of the image, -1 = not packed
    TcPackJob* tc_jobs = nullptr;       // device copy of the job list
    int tc_njobs = 0;
    long long tc_total = 0;             // elements over all jobs
    float* tc_packW = nullptr;          // images of the parameters (rebuilt by every base pass)
    float* tc_packV = nullptr;          // images of the tangent direction (rebuilt by every order-1 pass)

    b2s_pistate* pi = nullptr;
    Comm* comm = nullptr;
    LocalPeer* local_peer = nullptr;  // single GPU: per-channel barrier of the fused BatchNorm kernels (B2S_BN_CHANSYNC)
    int world = 1;
    long long global_batch = 0;       // samples over all ranks of the cached base pass (0: batch * world)

    // eigen-iteration loop control (b2s_power_iterate)
    int* h_done = nullptr;            // pinned ring of `done` flags read back with a lag
    static constexpr int kMaxLag = 8;
    cudaEvent_t ev_poll[kMaxLag + 1] = {};
    int poll_lag = -1;                // -1: not measured yet
    struct LoopGraph {                // whole loop as ONE graph launch: WHILE node around (HVP pass + vector kernels)
        cudaGraphExec_t exec = nullptr;
        long long kernels_per_iter = 0;
        double loss_scale = 0.0;
        long long global_batch = 0;
    };
    std::map<long long, LoopGraph> loops;   // key = batch
    int device_loop = -1;             // -1 undecided, 0 unavailable / disabled, 1 in use
};

namespace b2s {

static inline float* tptr(const b2s_plan* p, float* const* arena, int order, int t) {
    const b2s_tensor& T = p->tensors[t];
    return arena[order] ? arena[order] + p->buf_off[T.buf] + T.offset : nullptr;
}
static inline View tview(const b2s_plan* p, float* const* arena, int order, int t) {
    const b2s_tensor& T = p->tensors[t];
    View v;
    v.p = tptr(p, arena, order, t);
    v.C = T.C; v.H = T.H; v.W = T.W; v.sstride = T.sample_stride;
    return v;
}

static int alloc_order(b2s_plan* p, int k) {
    if (p->fw[k]) return 0;
    const size_t bytes = (size_t)p->arena_elems * sizeof(float);
    B2S_CUDA(cudaMalloc(&p->fw[k], bytes));
    B2S_CUDA(cudaMalloc(&p->bw[k], bytes));
    B2S_CUDA(cudaMemsetAsync(p->fw[k], 0, bytes, p->stream));
    B2S_CUDA(cudaMemsetAsync(p->bw[k], 0, bytes, p->stream));
    B2S_CUDA(cudaMalloc(&p->out32[k], (size_t)(p->P + 4) * sizeof(float)));
    p->workspace += 2 * (long long)bytes + (p->P + 4) * 4;
    if (p->bn_total > 0) {
        B2S_CUDA(cudaMalloc(&p->fsum[k], (size_t)p->bn_total * sizeof(double)));
        B2S_CUDA(cudaMalloc(&p->bsum[k], (size_t)p->bn_total * sizeof(double)));
        p->workspace += 2 * p->bn_total * 8;
    }
    return 0;
}

static ConvGeom conv_geom(const b2s_plan* p, const b2s_op& op) {
    const b2s_tensor& I = p->tensors[op.in];
    const b2s_tensor& O = p->tensors[op.out];
    ConvGeom g;
    g.batch = p->batch;
    g.Cin = I.C; g.H = I.H; g.W = I.W; g.in_sstride = I.sample_stride;
    g.Cout = O.C; g.OH = O.H; g.OW = O.W; g.out_sstride = O.sample_stride;
    g.KH = op.kh; g.KW = op.kw; g.sh = op.sh; g.sw = op.sw; g.ph = op.ph; g.pw = op.pw;
    return g;
}

static BnArgs bn_args(b2s_plan* p, int oi, int K) {
    const b2s_op& op = p->ops[oi];
    const b2s_tensor& I = p->tensors[op.in];
    const b2s_tensor& O = p->tensors[op.out];
    const bool first = op.flags & B2S_F_FIRST;
    BnArgs a{};
    a.batch = p->batch; a.C = I.C; a.HW = I.H * I.W;
    a.in_sstride = I.sample_stride; a.out_sstride = O.sample_stride;
    for (int k = 0; k < 3; ++k) {
        a.x[k] = (k <= K && (k == 0 || !first)) ? tptr(p, p->fw, k, op.in) : nullptr;
        a.g[k] = k <= K ? tptr(p, p->bw, k, op.out) : nullptr;
        a.fsum[k] = p->fsum[k] ? p->fsum[k] + p->bn_off[oi] : nullptr;
        a.bsum[k] = p->bsum[k] ? p->bsum[k] + p->bn_off[oi] : nullptr;
    }
    a.y0 = tptr(p, p->fw, 0, op.out);
    a.yk = tptr(p, p->fw, K, op.out);
    a.xbar = first ? nullptr : tptr(p, p->bw, K, op.in);
    a.gamma = p->params + op.w_off; a.beta = p->params + op.b_off;
    a.vgamma = p->v32 + op.w_off; a.vbeta = p->v32 + op.b_off;
    a.out_gamma = p->out32[K] + op.w_off; a.out_beta = p->out32[K] + op.b_off;
    a.running_mean = op.slot >= 0 && op.slot < (int)p->bn_rm.size() ? p->bn_rm[op.slot] : nullptr;
    a.running_var = op.slot >= 0 && op.slot < (int)p->bn_rv.size() ? p->bn_rv[op.slot] : nullptr;
    a.eps = op.eps; a.momentum = op.momentum;
    a.relu = (op.flags & B2S_F_RELU) ? 1 : 0;
    a.first = first ? 1 : 0;
    a.count = (p->global_batch > 0 ? p->global_batch : (long long)p->batch * p->world) * a.HW;
    a.accumulate = (op.flags & B2S_F_BWD_ACC) ? 1 : 0;
    a.pgrad_scale = 1.0f / (float)p->world;
    a.peer = (p->comm && 2 * a.C <= 4096) ? comm_peer_ctx(p->comm) : nullptr;
    a.peer_tail = (p->comm && 2 * a.C <= 4096) ? comm_peer_tail_ctx(p->comm) : nullptr;
    a.peer_ll = (a.peer && comm_peer_ll(p->comm)) ? 1 : 0;
    if (!p->comm && p->local_peer && 2 * a.C <= 4096) { a.peer = local_peer_ctx(p->local_peer); a.peer_ll = 2; }
    return a;
}

// packed image of a weight operand (W lives in p->params, V in p->v32), NULL when the layer is not packed
static inline const float* tc_image(const b2s_plan* p, int oi, int mode, const float* wt) {
    if (!p->tc_packW || p->tc_off[mode][oi] < 0) return nullptr;
    const bool is_v = wt >= p->v32 && wt < p->v32 + p->P;
    return (is_v ? p->tc_packV : p->tc_packW) + p->tc_off[mode][oi];
}

// ---- forward sweep of order K -----------------------------------------------------------------
static int forward(b2s_plan* p, int K, bool eval_mode = false) {
    cudaStream_t st = p->stream;
    if (p->tc_njobs > 0 && get_tc_mode() != 0) {
        if (K == 0) B2S_TRY(launch_tc_pack(st, p->tc_jobs, p->tc_njobs, p->tc_total, p->params, p->tc_packW));
        if (K == 1) B2S_TRY(launch_tc_pack(st, p->tc_jobs, p->tc_njobs, p->tc_total, p->v32, p->tc_packV));
    }
    if (p->bn_total > 0) B2S_CUDA(cudaMemsetAsync(p->fsum[K], 0, (size_t)p->bn_total * sizeof(double), st));
    for (size_t oi = 0; oi < p->ops.size(); ++oi) {
        const b2s_op& op = p->ops[oi];
        const bool first = op.flags & B2S_F_FIRST;
        const bool relu = op.flags & B2S_F_RELU;
        switch (op.kind) {
        case B2S_OP_CONV: {
            const ConvGeom g = conv_geom(p, op);
            const float* W = p->params + op.w_off;
            const float* V = p->v32 + op.w_off;
            const float* act[kMaxPairs];
            const float* wt[kMaxPairs];
            float sc[kMaxPairs];
            int np = 0;
            const float* bias = nullptr;
            float* out = tptr(p, p->fw, K, op.out);
            const float* y0 = tptr(p, p->fw, 0, op.out);
            if (K == 0) {
                act[np] = tptr(p, p->fw, 0, op.in); wt[np] = W; sc[np] = 1.f; ++np;
                if (op.b_off >= 0) bias = p->params + op.b_off;
            } else if (K == 1) {
                if (!first) { act[np] = tptr(p, p->fw, 1, op.in); wt[np] = W; sc[np] = 1.f; ++np; }
                act[np] = tptr(p, p->fw, 0, op.in); wt[np] = V; sc[np] = 1.f; ++np;
                if (op.b_off >= 0) bias = p->v32 + op.b_off;
            } else {
                if (first) {   // x has no tangent: yddot = 0
                    B2S_TRY(launch_zero_view(st, tview(p, p->fw, 2, op.out), p->batch));
                    break;
                }
                act[np] = tptr(p, p->fw, 2, op.in); wt[np] = W; sc[np] = 1.f; ++np;
                act[np] = tptr(p, p->fw, 1, op.in); wt[np] = V; sc[np] = 2.f; ++np;
            }
            const float* pk[kMaxPairs];
            for (int q = 0; q < np; ++q) pk[q] = tc_image(p, (int)oi, MODE_FWD, wt[q]);
            B2S_TRY(launch_conv_fwd(st, g, np, act, wt, sc, bias, relu ? (K == 0 ? 1 : 2) : 0, y0, out, 0, pk));
            break;
        }
        case B2S_OP_BN: {
            const BnArgs a = bn_args(p, (int)oi, K);
            if (eval_mode) {
                B2S_TRY(launch_bn_eval(st, a));
                break;
            }
            const int do_stats = !(first && K > 0);
            int rc = 1;
            if ((!p->comm || a.peer) && p->bn_fused) rc = launch_bn_fwd_fused(st, K, a, do_stats);
            if (rc < 0) return rc;
            if (rc == 1) {
                if (do_stats) {
                    B2S_TRY(launch_bn_fwd_stats(st, K, a));
                    if (p->comm && !a.peer_tail) B2S_TRY(comm_allreduce_f64(p->comm, a.fsum[K], 2 * a.C, st));
                }
                B2S_TRY(launch_bn_fwd_apply(st, K, a));
            }
            break;
        }
        case B2S_OP_RELU:
            B2S_TRY(launch_relu_fwd(st, K, tview(p, p->fw, 0, op.in), tview(p, p->fw, K, op.in),
                                    tview(p, p->fw, K, op.out), p->batch));
            break;
        case B2S_OP_MAXPOOL:
            B2S_TRY(launch_maxpool_fwd(st, K, tview(p, p->fw, K, op.in), tview(p, p->fw, K, op.out),
                                       p->argmax[oi], p->batch, op.kh, op.kw, op.sh, op.sw, op.ph, op.pw));
            break;
        case B2S_OP_AVGPOOL:
            B2S_TRY(launch_avgpool_fwd(st, tview(p, p->fw, K, op.in), tview(p, p->fw, K, op.out), p->batch,
                                       op.kh));
            break;
        case B2S_OP_COPY:
            B2S_TRY(launch_copy_view(st, tview(p, p->fw, K, op.in), tview(p, p->fw, K, op.out), p->batch, 0));
            break;
        case B2S_OP_ADD:
            B2S_TRY(launch_add_fwd(st, K, relu ? 1 : 0, tview(p, p->fw, K, op.in), tview(p, p->fw, K, op.slot),
                                   tview(p, p->fw, 0, op.out), tview(p, p->fw, K, op.out), p->batch));
            break;
        default:
            set_error("unknown op kind %d", op.kind);
            return -5;
        }
    }
    // loss head
    const b2s_tensor& L = p->tensors[p->logits];
    HeadArgs h{};
    h.kind = p->head; h.batch = p->batch; h.C = L.C * L.H * L.W;
    for (int k = 0; k < 3; ++k) h.z[k] = k <= K ? tptr(p, p->fw, k, p->logits) : nullptr;
    h.zs = L.sample_stride;
    h.labels = p->labels; h.target = p->target; h.coef = p->coef;
    h.loss_scale = p->loss_scale;
    h.zbar = tptr(p, p->bw, K, p->logits);
    h.loss = K == 0 ? p->loss : nullptr;
    if (K == 0) B2S_CUDA(cudaMemsetAsync(p->loss, 0, sizeof(double), st));
    B2S_TRY(launch_head(st, K, h));
    if (K == 0 && p->comm) B2S_TRY(comm_allreduce_f64(p->comm, p->loss, 1, st));
    return 0;
}

// The weight/bias gradient of a layer is a leaf of the backward sweep: nothing downstream reads it
// before the pass ends.  Fork it onto the side stream (ordered after everything enqueued so far on
// the main stream) so the adjoint chain  dgrad -> BN backward -> dgrad ...  does not wait for it.
static int fork_side(b2s_plan* p) {
    if (!p->wgrad_side) return 0;
    const int i = p->side_next;
    p->side_next = (i + 1) % p->n_side;
    p->side = p->sides[i];
    B2S_CUDA(cudaEventRecord(p->ev_fork, p->stream));
    B2S_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
    p->side_mask |= 1u << i;
    p->side_used = true;
    return 0;
}
static int join_side(b2s_plan* p) {
    if (!p->side_used) return 0;
    for (int i = 0; i < p->n_side; ++i) {
        if (!(p->side_mask & (1u << i))) continue;
        B2S_CUDA(cudaEventRecord(p->ev_joins[i], p->sides[i]));
        B2S_CUDA(cudaStreamWaitEvent(p->stream, p->ev_joins[i], 0));
    }
    p->side_mask = 0;
    p->side_next = 0;
    p->side_used = false;
    return 0;
}

// The weight gradient of a layer is a leaf of the sweep: it goes to a side stream right away.  (Measured and dropped in
// round 2: queueing the leaves of a dense block and contracting them in one persistent launch per block -- 2.65..2.78 ms
// per HVP against 2.50 ms: the per-layer launches fill the SMs the latency-bound adjoint chain leaves idle, a grouped
// launch can only start when its block is finished and the largest group ends up exposed at the end of the sweep.)
static int conv_wgrad_leaf(b2s_plan* p, const ConvGeom& g, int np, const float* const* a_, const float* const* b_,
                           const float* sc, float* wbar) {
    B2S_TRY(fork_side(p));
    return launch_conv_wgrad(p->wgrad_side ? p->side : p->stream, g, np, a_, b_, sc, wbar);
}

// adjoint of a residual connection: both operands receive (mask .) the output adjoint of arena order `ord`
static int add_backward(b2s_plan* p, const b2s_op& op, int ord) {
    const int in_buf = p->tensors[0].buf;
    const bool has_a = p->tensors[op.in].buf != in_buf, has_b = p->tensors[op.slot].buf != in_buf;
    const View y0 = tview(p, p->fw, 0, op.out), g = tview(p, p->bw, ord, op.out);
    const View ga = tview(p, p->bw, ord, op.in), gb = tview(p, p->bw, ord, op.slot);
    return launch_add_bwd(p->stream, (op.flags & B2S_F_RELU) ? &y0 : nullptr, g, has_a ? &ga : nullptr,
                          (op.flags & B2S_F_BWD_ACC) ? 1 : 0, has_b ? &gb : nullptr, (op.flags & B2S_F_BWD_ACC2) ? 1 : 0,
                          p->batch);
}

// ---- backward sweep of order K ----------------------------------------------------------------
static int backward(b2s_plan* p, int K) {
    cudaStream_t st = p->stream;
    B2S_CUDA(cudaMemsetAsync(p->out32[K], 0, (size_t)p->P * sizeof(float), st));
    if (p->bn_total > 0) B2S_CUDA(cudaMemsetAsync(p->bsum[K], 0, (size_t)p->bn_total * sizeof(double), st));
    for (int oi = (int)p->ops.size() - 1; oi >= 0; --oi) {
        const b2s_op& op = p->ops[oi];
        const bool first = op.flags & B2S_F_FIRST;
        const bool relu = op.flags & B2S_F_RELU;
        const int acc = (op.flags & B2S_F_BWD_ACC) ? 1 : 0;
        switch (op.kind) {
        case B2S_OP_CONV: {
            const ConvGeom g = conv_geom(p, op);
            if (relu) B2S_TRY(launch_mask_inplace(st, tview(p, p->fw, 0, op.out), tview(p, p->bw, K, op.out), p->batch));
            const float* W = p->params + op.w_off;
            const float* V = p->v32 + op.w_off;
            const float* x[3];
            const float* gk[3];
            for (int k = 0; k < 3; ++k) {
                x[k] = k <= K ? tptr(p, p->fw, k, op.in) : nullptr;
                gk[k] = k <= K ? tptr(p, p->bw, k, op.out) : nullptr;
            }
            const float* a_[kMaxPairs];
            const float* b_[kMaxPairs];
            float sc[kMaxPairs];
            int np = 0;
            // weight gradient of order K:  sum_j binom(K,j) corr(x_j, g_{K-j})
            a_[np] = x[0]; b_[np] = gk[K]; sc[np] = 1.f; ++np;
            if (!first && K >= 1) { a_[np] = x[1]; b_[np] = gk[K - 1]; sc[np] = K == 2 ? 2.f : 1.f; ++np; }
            if (!first && K == 2) { a_[np] = x[2]; b_[np] = gk[0]; sc[np] = 1.f; ++np; }
            B2S_TRY(conv_wgrad_leaf(p, g, np, a_, b_, sc, p->out32[K] + op.w_off));
            if (op.b_off >= 0) {
                B2S_TRY(fork_side(p));
                B2S_TRY(launch_bias_grad(p->wgrad_side ? p->side : st, gk[K], p->batch, g.Cout, g.OH * g.OW, g.out_sstride,
                                         p->out32[K] + op.b_off));
            }
            if (!first) {
                np = 0;
                a_[np] = gk[K]; b_[np] = W; sc[np] = 1.f; ++np;
                if (K >= 1) { a_[np] = gk[K - 1]; b_[np] = V; sc[np] = K == 2 ? 2.f : 1.f; ++np; }
                const float* pk[kMaxPairs];
                for (int q = 0; q < np; ++q) pk[q] = tc_image(p, oi, MODE_DGRAD, b_[q]);
                B2S_TRY(launch_conv_dgrad(st, g, np, a_, b_, sc, tptr(p, p->bw, K, op.in), acc, pk));
            }
            break;
        }
        case B2S_OP_BN: {
            const BnArgs a = bn_args(p, oi, K);
            int rc = 1;
            if ((!p->comm || a.peer) && p->bn_fused) rc = launch_bn_bwd_fused(st, K, a);
            if (rc < 0) return rc;
            if (rc == 1) {
                B2S_TRY(launch_bn_bwd_stats(st, K, a));
                if (p->comm && !a.peer_tail) B2S_TRY(comm_allreduce_f64(p->comm, a.bsum[K], 2 * a.C, st));
                B2S_TRY(launch_bn_bwd_apply(st, K, a));
            }
            break;
        }
        case B2S_OP_RELU:
            if (!first)
                B2S_TRY(launch_relu_bwd(st, tview(p, p->fw, 0, op.in), tview(p, p->bw, K, op.out),
                                        tview(p, p->bw, K, op.in), p->batch, acc));
            break;
        case B2S_OP_MAXPOOL:
            if (!first) {
                if (!acc) B2S_TRY(launch_zero_view(st, tview(p, p->bw, K, op.in), p->batch));
                B2S_TRY(launch_maxpool_bwd(st, tview(p, p->bw, K, op.out), tview(p, p->bw, K, op.in),
                                           p->argmax[oi], p->batch));
            }
            break;
        case B2S_OP_AVGPOOL:
            if (!first)
                B2S_TRY(launch_avgpool_bwd(st, tview(p, p->bw, K, op.out), tview(p, p->bw, K, op.in), p->batch,
                                           op.kh, acc));
            break;
        case B2S_OP_COPY:
            if (!first)
                B2S_TRY(launch_copy_view(st, tview(p, p->bw, K, op.out), tview(p, p->bw, K, op.in), p->batch, acc));
            break;
        case B2S_OP_ADD:
            B2S_TRY(add_backward(p, op, K));
            break;
        default:
            set_error("unknown op kind %d", op.kind);
            return -5;
        }
    }
    B2S_TRY(join_side(p));
    if (p->comm) B2S_TRY(comm_allreduce_f32(p->comm, p->out32[K], p->P, st));
    return 0;
}

// First-order backward sweep with zero loss seed that carries the adjoint torch's third-order
// BatchNorm derivative drops (bn.cu, "reference-compatible third order").  Runs after pass 2 and
// reuses its adjoint arena bw[2]; result in p->out_corr.
static int backward_correction(b2s_plan* p) {
    cudaStream_t st = p->stream;
    const int ns = bn_corr_sums();
    B2S_CUDA(cudaMemsetAsync(p->out_corr, 0, (size_t)p->P * sizeof(float), st));
    B2S_CUDA(cudaMemsetAsync(p->csum, 0, (size_t)p->bn_total / 2 * ns * sizeof(double), st));
    B2S_TRY(launch_zero_view(st, tview(p, p->bw, 2, p->logits), p->batch));
    for (int oi = (int)p->ops.size() - 1; oi >= 0; --oi) {
        const b2s_op& op = p->ops[oi];
        const bool first = op.flags & B2S_F_FIRST;
        const bool relu = op.flags & B2S_F_RELU;
        const int acc = (op.flags & B2S_F_BWD_ACC) ? 1 : 0;
        switch (op.kind) {
        case B2S_OP_CONV: {
            const ConvGeom g = conv_geom(p, op);
            if (relu) B2S_TRY(launch_mask_inplace(st, tview(p, p->fw, 0, op.out), tview(p, p->bw, 2, op.out), p->batch));
            const float* x0 = tptr(p, p->fw, 0, op.in);
            const float* gc = tptr(p, p->bw, 2, op.out);
            const float* W = p->params + op.w_off;
            const float one = 1.f;
            B2S_TRY(conv_wgrad_leaf(p, g, 1, &x0, &gc, &one, p->out_corr + op.w_off));
            if (op.b_off >= 0) {
                B2S_TRY(fork_side(p));
                B2S_TRY(launch_bias_grad(p->wgrad_side ? p->side : st, gc, p->batch, g.Cout, g.OH * g.OW, g.out_sstride,
                                         p->out_corr + op.b_off));
            }
            if (!first) {
                const float* pk = tc_image(p, oi, MODE_DGRAD, W);
                B2S_TRY(launch_conv_dgrad(st, g, 1, &gc, &W, &one, tptr(p, p->bw, 2, op.in), acc, &pk));
            }
            break;
        }
        case B2S_OP_BN: {
            BnArgs a = bn_args(p, oi, 1);
            a.csum = p->csum + p->bn_off[oi] / 2 * ns;
            a.out_gamma = p->out_corr + op.w_off;
            a.out_beta = p->out_corr + op.b_off;
            const float* gc = tptr(p, p->bw, 2, op.out);
            B2S_TRY(launch_bn_corr_stats(st, a, gc));
            if (p->comm) B2S_TRY(comm_allreduce_f64(p->comm, a.csum, ns * a.C, st));
            B2S_TRY(launch_bn_corr_apply(st, a, gc, first ? nullptr : tptr(p, p->bw, 2, op.in)));
            break;
        }
        case B2S_OP_RELU:
            if (!first)
                B2S_TRY(launch_relu_bwd(st, tview(p, p->fw, 0, op.in), tview(p, p->bw, 2, op.out),
                                        tview(p, p->bw, 2, op.in), p->batch, acc));
            break;
        case B2S_OP_MAXPOOL:
            if (!first) {
                if (!acc) B2S_TRY(launch_zero_view(st, tview(p, p->bw, 2, op.in), p->batch));
                B2S_TRY(launch_maxpool_bwd(st, tview(p, p->bw, 2, op.out), tview(p, p->bw, 2, op.in),
                                           p->argmax[oi], p->batch));
            }
            break;
        case B2S_OP_AVGPOOL:
            if (!first)
                B2S_TRY(launch_avgpool_bwd(st, tview(p, p->bw, 2, op.out), tview(p, p->bw, 2, op.in), p->batch,
                                           op.kh, acc));
            break;
        case B2S_OP_COPY:
            if (!first)
                B2S_TRY(launch_copy_view(st, tview(p, p->bw, 2, op.out), tview(p, p->bw, 2, op.in), p->batch, acc));
            break;
        case B2S_OP_ADD:
            B2S_TRY(add_backward(p, op, 2));
            break;
        default:
            set_error("unknown op kind %d", op.kind);
            return -5;
        }
    }
    B2S_TRY(join_side(p));
    if (p->comm) B2S_TRY(comm_allreduce_f32(p->comm, p->out_corr, p->P, st));
    return 0;
}

static int run_pass_eager_impl(b2s_plan* p, int K);
static int run_pass_eager(b2s_plan* p, int K) {
    set_tc_splitk_allowed(K != 0 && K != 4);      // base and evaluation passes: deterministic summation (conv_tc.cu)
    const int rc = run_pass_eager_impl(p, K);
    set_tc_splitk_allowed(true);
    return rc;
}
static int run_pass_eager_impl(b2s_plan* p, int K) {
    if (K == 3) return backward_correction(p);
    if (K == 4) {                      // evaluation pass: local to the rank, no collective
        Comm* comm = p->comm;
        p->comm = nullptr;
        const int rc = forward(p, 0, true);
        p->comm = comm;
        return rc;
    }
    B2S_TRY(forward(p, K));
    B2S_TRY(backward(p, K));
    return 0;
}

// one pass of order K on p->stream, through a CUDA graph when enabled
static int run_pass(b2s_plan* p, int K) {
    if (K < 3) B2S_TRY(alloc_order(p, K));
    if (!p->use_graphs) return run_pass_eager(p, K);
    // (K = 3: compatibility sweep, K = 4: evaluation pass -- both live in the order-2 / order-0 arenas)
    const long long key = ((long long)K << 32) | (unsigned)p->batch;
    auto it = p->graphs.find(key);
    if (it != p->graphs.end() && (it->second.loss_scale != p->loss_scale || it->second.global_batch != p->global_batch)) {
        if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
        p->graphs.erase(it);
        it = p->graphs.end();
    }
    if (it == p->graphs.end()) {
        cudaGraph_t graph = nullptr;
        g_counting_paused = true;
        g_capture_count = 0;
        cudaError_t e = cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            g_counting_paused = false;
            set_error("cudaStreamBeginCapture: %s", cudaGetErrorString(e));
            return -2;
        }
        const int rc = run_pass_eager(p, K);
        e = cudaStreamEndCapture(p->stream, &graph);
        g_counting_paused = false;
        if (rc != 0) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (e != cudaSuccess) {
            set_error("cudaStreamEndCapture: %s", cudaGetErrorString(e));
            return -2;
        }
        GraphEntry ge;
        ge.kernels = g_capture_count;
        ge.loss_scale = p->loss_scale;
        ge.global_batch = p->global_batch;
        e = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
            return -2;
        }
        it = p->graphs.emplace(key, ge).first;
    }
    B2S_CUDA(cudaGraphLaunch(it->second.exec, p->stream));
    count_launch((int)it->second.kernels);
    return 0;
}

// order the plan stream after the caller's stream / the caller's stream after the plan stream
static int peer_status(b2s_plan* p) {
    if (!p->comm) return 0;
    const int e = comm_peer_error(p->comm);
    if (e) {
        set_error("data-parallel exchange: rank %d did not reach a BatchNorm statistics exchange within B2S_PEER_TIMEOUT_S; "
                  "the results of the last call are incomplete", e - 1);
        return -9;
    }
    return 0;
}
static void drop_graphs(b2s_plan* p) {
    for (auto& kv : p->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    p->graphs.clear();
    for (auto& kv : p->loops) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    p->loops.clear();
}
static int enter(b2s_plan* p) {
    B2S_CUDA(cudaSetDevice(p->device));
    B2S_TRY(peer_status(p));
    B2S_CUDA(cudaEventRecord(p->ev_in, p->caller));
    B2S_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in, 0));
    return 0;
}
static int leave(b2s_plan* p) {
    B2S_CUDA(cudaEventRecord(p->ev_out, p->stream));
    B2S_CUDA(cudaStreamWaitEvent(p->caller, p->ev_out, 0));
    return peer_status(p);
}

static int pi_upload(b2s_pistate* s, cudaStream_t st) {
    B2S_CUDA(cudaMemcpyAsync(s->d, &s->h, sizeof(PiDev), cudaMemcpyHostToDevice, st));
    // the host mirror may be rewritten right after this call returns
    B2S_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int pi_reset_impl(b2s_pistate* s, const double* d_v0, const b2s_power_cfg* cfg, cudaStream_t st) {
    if (cfg->max_iter > s->cap) {
        set_error("power iteration: max_iter %d exceeds the state's capacity %d", cfg->max_iter, s->cap);
        return -6;
    }
    B2S_CUDA(cudaMemcpyAsync(s->vbuf[0], d_v0, (size_t)s->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    B2S_TRY(launch_cast_f64_f32(st, s->vbuf[0], s->v32, s->n));
    s->have_alpha = cfg->h_alpha != nullptr;
    if (s->have_alpha)
        B2S_CUDA(cudaMemcpyAsync(s->alpha, cfg->h_alpha, (size_t)cfg->max_iter * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
    PiDev& h = s->h;
    memset(&h, 0, sizeof(h));
    h.n = s->n;
    h.vbuf[0] = s->vbuf[0]; h.vbuf[1] = s->vbuf[1];
    h.rbuf[0] = s->rbuf[0]; h.rbuf[1] = s->rbuf[1];
    h.v32 = s->v32;
    h.alpha = s->have_alpha ? s->alpha : nullptr;
    h.traj = s->traj;
    h.scratch = s->scratch;
    h.precond = cfg->precond;
    h.max_iter = cfg->max_iter;
    h.last_iter = -1;
    h.eps = cfg->eps;
    h.stop[0] = h.stop[1] = h.stop[2] = INFINITY;
    if (cfg->max_iter <= 0) h.done = 1;
    return pi_upload(s, st);
}

static int pi_result_impl(b2s_pistate* s, b2s_power_result* r, double* h_traj, double* d_v_out, cudaStream_t st) {
    PiDev h;
    B2S_CUDA(cudaMemcpyAsync(&h, s->d, sizeof(PiDev), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    if (r) {
        r->iters = h.last_iter;
        r->converged = h.converged;
        r->lam = h.lam; r->norm = h.norm; r->rn = h.rn; r->vnn = h.vnn;
        for (int k = 0; k < 3; ++k) r->stop[k] = h.stop[k];
    }
    if (h_traj && h.last_iter >= 0) {
        std::vector<double> tmp((size_t)(h.last_iter + 1) * 4);
        B2S_CUDA(cudaMemcpy(tmp.data(), s->traj, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i <= h.last_iter; ++i) {
            h_traj[i * 5 + 0] = i;
            for (int q = 0; q < 4; ++q) h_traj[i * 5 + 1 + q] = tmp[(size_t)i * 4 + q];
        }
    }
    if (d_v_out)
        B2S_CUDA(cudaMemcpyAsync(d_v_out, s->vbuf[h.cur], (size_t)s->n * sizeof(double),
                                 cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace b2s

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int b2s_abi_version(void) { return B2S_ABI_VERSION; }
const char* b2s_last_error(void) { return g_err; }
int64_t b2s_launch_count(void) {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    return g_launch_total;
}

int b2s_plan_create(const b2s_tensor* tensors, int32_t n_tensors, const int64_t* buf_elems, int32_t n_bufs,
                    const b2s_op* ops, int32_t n_ops, int32_t logits, int32_t head, int64_t n_params,
                    int32_t max_batch, int32_t device, b2s_plan** out) {
    if (!tensors || !buf_elems || !ops || !out || n_tensors <= 0 || n_bufs <= 0 || n_ops <= 0 || max_batch <= 0) {
        set_error("b2s_plan_create: invalid arguments");
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("b2s_plan_create: no CUDA device is available; this library has no CPU fallback");
        return -7;
    }
    B2S_CUDA(cudaSetDevice(device));
    b2s_plan* p = new b2s_plan();
    p->device = device;
    if (const char* e = getenv("B2S_BN_FUSED")) p->bn_fused = atoi(e) != 0;      // experiment switch
    if (const char* e = getenv("B2S_WGRAD_SIDE")) p->wgrad_side = atoi(e) != 0;
    {
        const char* e = getenv("B2S_BN_CHANSYNC");
        if (e && atoi(e) != 0 && local_peer_create(&p->local_peer) != 0) { delete p; return -2; }
    }
    p->tensors.assign(tensors, tensors + n_tensors);
    p->ops.assign(ops, ops + n_ops);
    p->buf_elems.assign(buf_elems, buf_elems + n_bufs);
    p->logits = logits; p->head = head; p->P = n_params; p->max_batch = max_batch;
    // validate
    for (int t = 0; t < n_tensors; ++t) {
        const b2s_tensor& T = tensors[t];
        if (T.buf < 0 || T.buf >= n_bufs || T.offset < 0 ||
            T.offset + (long long)T.C * T.H * T.W > buf_elems[T.buf] || T.sample_stride != buf_elems[T.buf]) {
            set_error("b2s_plan_create: tensor %d does not fit its buffer", t);
            delete p;
            return -1;
        }
    }
    p->buf_off.resize(n_bufs);
    long long off = 0;
    for (int b = 0; b < n_bufs; ++b) {
        p->buf_off[b] = off;
        long long e = buf_elems[b] * (long long)max_batch;
        off += (e + 63) / 64 * 64;     // 256-byte alignment of every buffer
    }
    p->arena_elems = off;
    p->bn_off.assign(n_ops, -1);
    p->argmax.assign(n_ops, nullptr);
    int max_slot = -1;
    for (int i = 0; i < n_ops; ++i) {
        const b2s_op& op = ops[i];
        if (op.in < 0 || op.in >= n_tensors || op.out < 0 || op.out >= n_tensors) {
            set_error("b2s_plan_create: op %d references an unknown tensor", i);
            delete p;
            return -1;
        }
        if (op.kind == B2S_OP_ADD && (op.slot < 0 || op.slot >= n_tensors)) {
            set_error("b2s_plan_create: op %d (add) references an unknown second operand", i);
            delete p;
            return -1;
        }
        if (op.kind == B2S_OP_BN) {
            p->bn_off[i] = p->bn_total;
            p->bn_total += 2LL * tensors[op.in].C;
            if (op.slot > max_slot) max_slot = op.slot;
            ++p->n_bn;
        }
    }
    p->bn_rm.assign(max_slot + 1, nullptr);
    p->bn_rv.assign(max_slot + 1, nullptr);
    int rc = 0;
    auto fail = [&](int code) { b2s_plan_destroy(p); return code; };
    // the adjoint chain runs at the highest stream priority, the weight-gradient leaves at the lowest: thread blocks of a
    // chain kernel are dispatched ahead of pending leaf blocks whenever an SM frees up (B2S_STREAM_PRIO=0: equal)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (const char* e = getenv("B2S_STREAM_PRIO")) { if (atoi(e) == 0) prio_lo = prio_hi = 0; }
    if (cudaStreamCreateWithPriority(&p->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming) != cudaSuccess) {
        set_error("b2s_plan_create: stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(-2);
    }
    if (const char* e = getenv("B2S_SIDE_STREAMS")) p->n_side = std::max(1, std::min((int)b2s_plan::kMaxSide, atoi(e)));
    for (int i = 0; i < p->n_side; ++i) {
        if (cudaStreamCreateWithPriority(&p->sides[i], cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_joins[i], cudaEventDisableTiming) != cudaSuccess) {
            set_error("b2s_plan_create: side stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(-2);
        }
    }
    p->side = p->sides[0];
    if ((rc = alloc_order(p, 0)) != 0) return fail(rc);
    if ((rc = alloc_order(p, 1)) != 0) return fail(rc);
    const b2s_tensor& L = tensors[logits];
    p->n_classes = L.C * L.H * L.W;
    const size_t lab = (size_t)max_batch * p->n_classes;
    if (cudaMalloc(&p->params, (size_t)(n_params + 4) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&p->v32, (size_t)(n_params + 4) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&p->loss, sizeof(double)) != cudaSuccess ||
        cudaMalloc(&p->labels, (size_t)max_batch * sizeof(long long)) != cudaSuccess ||
        cudaMalloc(&p->target, lab * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&p->coef, lab * sizeof(float)) != cudaSuccess) {
        set_error("b2s_plan_create: out of device memory: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(-2);
    }
    cudaMemsetAsync(p->v32, 0, (size_t)(n_params + 4) * sizeof(float), p->stream);
    cudaMemsetAsync(p->params, 0, (size_t)(n_params + 4) * sizeof(float), p->stream);
    p->workspace += 2 * (n_params + 4) * 4 + (long long)lab * 8;
    {   // tcgen05 path: one pack job per eligible (Conv/Linear op, contraction mode); ops that share a
        // weight tensor (forest's fc2) get their own image -- the geometry could differ
        std::vector<TcPackJob> jobs;
        long long floats = 0, elems = 0;
        p->tc_off[0].assign(n_ops, -1);
        p->tc_off[1].assign(n_ops, -1);
        for (int i = 0; i < n_ops; ++i) {
            const b2s_op& op = ops[i];
            if (op.kind != B2S_OP_CONV) continue;
            const b2s_tensor& I = tensors[op.in];
            const b2s_tensor& O = tensors[op.out];
            for (int mode = 0; mode < 2; ++mode) {
                if (mode == MODE_DGRAD && (op.flags & B2S_F_FIRST)) continue;
                const int Cs = mode == MODE_FWD ? I.C : O.C, Cd = mode == MODE_FWD ? O.C : I.C;
                const int Hs = mode == MODE_FWD ? I.H : O.H, Ws = mode == MODE_FWD ? I.W : O.W;
                if (!tc_shape_ok(Cs, Cd, Hs, Ws)) continue;
                TcPackJob jb{};
                jb.src_off = op.w_off; jb.dst_off = floats; jb.begin = elems;
                jb.Cs = Cs; jb.Cd = Cd; jb.KHW = op.kh * op.kw; jb.mode = mode;
                jb.BN = tc_choose_bn(Cd); jb.nchunks = (Cs + 31) / 32;
                const long long f = tc_pack_floats(Cs, Cd, jb.KHW);
                p->tc_off[mode][i] = floats;
                floats += f;
                elems += f / 2;
                jobs.push_back(jb);
            }
        }
        if (!jobs.empty()) {
            if (cudaMalloc(&p->tc_jobs, jobs.size() * sizeof(TcPackJob)) != cudaSuccess ||
                cudaMalloc(&p->tc_packW, (size_t)floats * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&p->tc_packV, (size_t)floats * sizeof(float)) != cudaSuccess) {
                set_error("b2s_plan_create: out of device memory (packed weight images)");
                return fail(-2);
            }
            cudaMemcpy(p->tc_jobs, jobs.data(), jobs.size() * sizeof(TcPackJob), cudaMemcpyHostToDevice);
            p->tc_njobs = (int)jobs.size();
            p->tc_total = elems;
            p->workspace += 2 * floats * 4;
        }
    }
    for (int i = 0; i < n_ops; ++i) {
        if (ops[i].kind == B2S_OP_MAXPOOL) {
            const b2s_tensor& O = tensors[ops[i].out];
            const size_t n = (size_t)max_batch * O.C * O.H * O.W;
            if (cudaMalloc(&p->argmax[i], n * sizeof(int32_t)) != cudaSuccess) {
                set_error("b2s_plan_create: out of device memory (argmax)");
                return fail(-2);
            }
            p->workspace += (long long)n * 4;
        }
    }
    *out = p;
    return 0;
}

int b2s_plan_destroy(b2s_plan* p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    drop_graphs(p);
    if (p->h_done) cudaFreeHost(p->h_done);
    for (auto& e : p->ev_poll) if (e) cudaEventDestroy(e);
    for (int k = 0; k < 3; ++k) {
        cudaFree(p->fw[k]); cudaFree(p->bw[k]); cudaFree(p->out32[k]);
        cudaFree(p->fsum[k]); cudaFree(p->bsum[k]);
    }
    for (auto a : p->argmax) cudaFree(a);
    cudaFree(p->csum); cudaFree(p->out_corr);
    cudaFree(p->tc_jobs); cudaFree(p->tc_packW); cudaFree(p->tc_packV);
    cudaFree(p->kfac_m); cudaFree(p->kfac_tr);
    cudaFree(p->params); cudaFree(p->v32); cudaFree(p->loss);
    cudaFree(p->labels); cudaFree(p->target); cudaFree(p->coef);
    if (p->pi) b2s_pi_destroy(p->pi);
    if (p->comm) comm_destroy(p->comm);
    local_peer_destroy(p->local_peer);
    for (int i = 0; i < b2s_plan::kMaxSide; ++i) {
        if (p->sides[i]) { cudaStreamSynchronize(p->sides[i]); cudaStreamDestroy(p->sides[i]); }
        if (p->ev_joins[i]) cudaEventDestroy(p->ev_joins[i]);
    }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_in) cudaEventDestroy(p->ev_in);
    if (p->ev_out) cudaEventDestroy(p->ev_out);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return 0;
}

int b2s_plan_set_stream(b2s_plan* p, void* cuda_stream) {
    if (!p) return -1;
    p->caller = (cudaStream_t)cuda_stream;
    return 0;
}
int b2s_plan_set_graphs(b2s_plan* p, int32_t use_graphs) {
    if (!p) return -1;
    p->use_graphs = (use_graphs & 1) != 0;
    p->bn_fused = (use_graphs & 2) == 0;       // bit 1: disable the cooperative BatchNorm kernels (debug)
    return 0;
}
int64_t b2s_plan_workspace_bytes(const b2s_plan* p) { return p ? p->workspace : 0; }
int b2s_set_tensor_core_mode(int32_t mode) {
    if (mode < 0 || mode > 2) { set_error("b2s_set_tensor_core_mode: mode must be 0, 1 or 2"); return -1; }
    set_tc_mode(mode);
    return 0;
}
int b2s_plan_set_bn_third_order(b2s_plan* p, int32_t exact) {
    if (!p) return -1;
    p->bn_exact_third_order = exact != 0;
    return 0;
}

int b2s_plan_set_bn_buffers(b2s_plan* p, int32_t slot, void* rm, void* rv) {
    if (!p || slot < 0 || slot >= (int)p->bn_rm.size()) {
        set_error("b2s_plan_set_bn_buffers: bad slot %d", slot);
        return -1;
    }
    if (p->bn_rm[slot] != rm || p->bn_rv[slot] != rv) {
        // captured order-0 graphs hold the old pointers
        for (auto it = p->graphs.begin(); it != p->graphs.end();) {
            if ((it->first >> 32) == 0 || (it->first >> 32) == 4) {
                if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
                it = p->graphs.erase(it);
            } else {
                ++it;
            }
        }
    }
    p->bn_rm[slot] = (float*)rm;
    p->bn_rv[slot] = (float*)rv;
    return 0;
}

int b2s_base_pass(b2s_plan* p, const float* d_params, const float* d_x, const void* d_y, const float* d_coef,
                  int32_t batch, double loss_scale, double* d_grad_out, double* d_loss_out) {
    if (!p || !d_params || !d_x || !d_y) { set_error("b2s_base_pass: null argument"); return -1; }
    if (batch <= 0 || batch > p->max_batch) {
        set_error("b2s_base_pass: batch %d outside (0, %d]", batch, p->max_batch);
        return -1;
    }
    const bool wbce = p->head == B2S_HEAD_WBCE || p->head == B2S_HEAD_SIGMOID_WBCE;
    if (wbce && !d_coef) { set_error("b2s_base_pass: weighted-BCE head needs d_coef"); return -1; }
    B2S_TRY(enter(p));
    cudaStream_t st = p->stream;
    p->batch = batch;
    p->loss_scale = loss_scale;
    p->v_is_current = false;
    B2S_CUDA(cudaMemcpyAsync(p->params, d_params, (size_t)p->P * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const b2s_tensor& X = p->tensors[0];     // tensor 0 is the network input by convention
    B2S_CUDA(cudaMemcpyAsync(p->fw[0] + p->buf_off[X.buf], d_x,
                             (size_t)batch * p->buf_elems[X.buf] * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (wbce) {
        const size_t nb = (size_t)batch * p->n_classes * sizeof(float);
        B2S_CUDA(cudaMemcpyAsync(p->target, d_y, nb, cudaMemcpyDeviceToDevice, st));
        B2S_CUDA(cudaMemcpyAsync(p->coef, d_coef, nb, cudaMemcpyDeviceToDevice, st));
    } else {
        B2S_CUDA(cudaMemcpyAsync(p->labels, d_y, (size_t)batch * sizeof(long long), cudaMemcpyDeviceToDevice, st));
    }
    B2S_TRY(run_pass(p, 0));
    if (d_grad_out) B2S_TRY(launch_cast_f32_f64(st, p->out32[0], d_grad_out, p->P, 1.0));
    if (d_loss_out) B2S_CUDA(cudaMemcpyAsync(d_loss_out, p->loss, sizeof(double), cudaMemcpyDeviceToDevice, st));
    return leave(p);
}

static int hv_impl(b2s_plan* p, const double* d_v) {
    if (p->batch <= 0) { set_error("b2s_hv: call b2s_base_pass first"); return -1; }
    B2S_TRY(launch_cast_f64_f32(p->stream, d_v, p->v32, p->P));
    B2S_TRY(run_pass(p, 1));
    p->v_is_current = true;
    return 0;
}

int b2s_hv(b2s_plan* p, const double* d_v, double* d_out) {
    if (!p || !d_v || !d_out) { set_error("b2s_hv: null argument"); return -1; }
    B2S_TRY(enter(p));
    B2S_TRY(hv_impl(p, d_v));
    B2S_TRY(launch_cast_f32_f64(p->stream, p->out32[1], d_out, p->P, 1.0));
    return leave(p);
}

// Forward-only evaluation pass (comp_f, opt.py:544-572; every forward of test_model goes through it, opt.py:954):
// values of order 0 with BatchNorm in evaluation mode and the loss of the head.  It overwrites the value caches of
// the base pass, so the plan forgets its cached minibatch (a later b2s_hv needs a new b2s_base_pass).
int b2s_eval_pass(b2s_plan* p, const float* d_params, const float* d_x, const void* d_y, const float* d_coef, int32_t batch,
                  double loss_scale, float* d_logits_out, double* d_loss_out) {
    if (!p || !d_params || !d_x || !d_y) { set_error("b2s_eval_pass: null argument"); return -1; }
    if (batch <= 0 || batch > p->max_batch) {
        set_error("b2s_eval_pass: batch %d outside (0, %d]", batch, p->max_batch);
        return -1;
    }
    const bool wbce = p->head == B2S_HEAD_WBCE || p->head == B2S_HEAD_SIGMOID_WBCE;
    if (wbce && !d_coef) { set_error("b2s_eval_pass: weighted-BCE head needs d_coef"); return -1; }
    B2S_TRY(enter(p));
    cudaStream_t st = p->stream;
    p->batch = batch;
    p->loss_scale = loss_scale;
    p->v_is_current = false;
    B2S_CUDA(cudaMemcpyAsync(p->params, d_params, (size_t)p->P * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const b2s_tensor& X = p->tensors[0];
    B2S_CUDA(cudaMemcpyAsync(p->fw[0] + p->buf_off[X.buf], d_x, (size_t)batch * p->buf_elems[X.buf] * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    if (wbce) {
        const size_t nb = (size_t)batch * p->n_classes * sizeof(float);
        B2S_CUDA(cudaMemcpyAsync(p->target, d_y, nb, cudaMemcpyDeviceToDevice, st));
        B2S_CUDA(cudaMemcpyAsync(p->coef, d_coef, nb, cudaMemcpyDeviceToDevice, st));
    } else {
        B2S_CUDA(cudaMemcpyAsync(p->labels, d_y, (size_t)batch * sizeof(long long), cudaMemcpyDeviceToDevice, st));
    }
    const int rc = run_pass(p, 4);
    if (rc != 0) { p->batch = 0; return rc; }
    if (d_logits_out) {
        const b2s_tensor& L = p->tensors[p->logits];
        const size_t row = (size_t)L.C * L.H * L.W * sizeof(float);
        B2S_CUDA(cudaMemcpy2DAsync(d_logits_out, row, tptr(p, p->fw, 0, p->logits), (size_t)L.sample_stride * sizeof(float),
                                   row, (size_t)batch, cudaMemcpyDeviceToDevice, st));
    }
    if (d_loss_out) B2S_CUDA(cudaMemcpyAsync(d_loss_out, p->loss, sizeof(double), cudaMemcpyDeviceToDevice, st));
    p->batch = 0;
    return leave(p);
}

int b2s_vghv(b2s_plan* p, const double* d_v, double* d_out) {
    if (!p || !d_v || !d_out) { set_error("b2s_vghv: null argument"); return -1; }
    B2S_TRY(enter(p));
    B2S_TRY(alloc_order(p, 2));
    B2S_TRY(hv_impl(p, d_v));          // order-1 caches for this v
    B2S_TRY(run_pass(p, 2));
    if (p->n_bn > 0 && !p->bn_exact_third_order) {
        if (!p->out_corr) {
            B2S_CUDA(cudaMalloc(&p->out_corr, (size_t)(p->P + 4) * sizeof(float)));
            B2S_CUDA(cudaMalloc(&p->csum, (size_t)p->bn_total / 2 * bn_corr_sums() * sizeof(double)));
        }
        B2S_TRY(run_pass(p, 3));
        B2S_TRY(launch_sub_cast_f32_f64(p->stream, p->out32[2], p->out_corr, d_out, p->P));
    } else {
        B2S_TRY(launch_cast_f32_f64(p->stream, p->out32[2], d_out, p->P, 1.0));
    }
    return leave(p);
}

int b2s_debug_read(b2s_plan* p, int32_t adjoint, int32_t order, int32_t tensor, float* h_out) {
    if (!p || !h_out || order < 0 || order > 2 || tensor < 0 || tensor >= (int)p->tensors.size()) {
        set_error("b2s_debug_read: invalid arguments");
        return -1;
    }
    float* const* arena = adjoint ? p->bw : p->fw;
    if (!arena[order] || p->batch <= 0) { set_error("b2s_debug_read: nothing cached for that order"); return -1; }
    B2S_CUDA(cudaSetDevice(p->device));
    B2S_CUDA(cudaStreamSynchronize(p->stream));
    const b2s_tensor& T = p->tensors[tensor];
    const size_t n = (size_t)T.C * T.H * T.W;
    B2S_CUDA(cudaMemcpy2D(h_out, n * sizeof(float), tptr(p, arena, order, tensor), (size_t)T.sample_stride * sizeof(float),
                          n * sizeof(float), (size_t)p->batch, cudaMemcpyDeviceToHost));
    return 0;
}

int b2s_profile_pass(b2s_plan* p, int32_t order, int32_t reps, b2s_prof_entry* out, int32_t cap, int32_t* n_out) {
    const bool raw = (order & 0x100) != 0;       // one entry per launch, in launch order
    order &= 0xff;
    if (!p || !out || !n_out || order < 0 || order > 3 || reps <= 0) { set_error("b2s_profile_pass: invalid arguments"); return -1; }
    if (p->batch <= 0) { set_error("b2s_profile_pass: call b2s_base_pass first"); return -1; }
    if (order == 3 && !p->out_corr) { set_error("b2s_profile_pass: run b2s_vghv once before profiling the compatibility sweep"); return -1; }
    B2S_TRY(enter(p));
    if (order < 3) B2S_TRY(alloc_order(p, order));
    Profiler prof;
    int rc = 0;
    // every kernel is timed ALONE: the weight-gradient leaves run on the main stream while profiling (on their side
    // streams the event pair of a leaf would also measure its wait for SMs behind the main chain's kernels -- round 2's
    // bench line carried 39 us per weight-gradient launch for a kernel that takes 17-26 us)
    const bool side = p->wgrad_side;
    p->wgrad_side = false;
    rc = run_pass_eager(p, order);                   // warm-up, untimed
    if (rc == 0) {
        g_prof = &prof;
        for (int r = 0; r < reps && rc == 0; ++r) rc = run_pass_eager(p, order);
        g_prof = nullptr;
    }
    cudaStreamSynchronize(p->stream);
    p->wgrad_side = side;
    std::map<std::string, b2s_prof_entry> agg;
    std::vector<std::string> order_seen;
    const size_t per_pass = prof.recs.size() / (size_t)reps;
    size_t ri = 0;
    for (auto& r : prof.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        char key[64];
        if (raw) snprintf(key, sizeof(key), "%s#%03d", r.name, (int)(per_pass ? ri % per_pass : ri));
        else snprintf(key, sizeof(key), "%s", r.name);
        ++ri;
        auto it = agg.find(key);
        if (it == agg.end()) {
            b2s_prof_entry e{};
            snprintf(e.name, sizeof(e.name), "%s", key);
            it = agg.emplace(key, e).first;
            order_seen.push_back(key);
        }
        it->second.launches += 1;
        it->second.ms += ms;
        it->second.flops += r.flops;
        it->second.bytes += r.bytes;
    }
    int n = 0;
    for (auto& name : order_seen) {
        if (n >= cap) break;
        b2s_prof_entry e = agg[name];
        e.launches /= reps; e.ms /= reps; e.flops /= reps; e.bytes /= reps;
        out[n++] = e;
    }
    *n_out = n;
    if (rc != 0) return rc;
    return leave(p);
}

const float* b2s_grad_f32(const b2s_plan* p) { return p ? p->out32[0] : nullptr; }
const float* b2s_hv_f32(const b2s_plan* p) { return p ? p->out32[1] : nullptr; }

// ---- iteration state ---------------------------------------------------------------------------
static int pi_create_impl(int64_t n, int32_t cap, int32_t device, float* external_v32, b2s_pistate** out) {
    if (n <= 0 || cap <= 0 || !out) { set_error("b2s_pi_create: invalid arguments"); return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("b2s_pi_create: no CUDA device is available; this library has no CPU fallback");
        return -7;
    }
    B2S_CUDA(cudaSetDevice(device));
    b2s_pistate* s = new b2s_pistate();
    s->device = device; s->n = n; s->cap = cap;
    const size_t vb = (size_t)(n + 4) * sizeof(double);
    bool ok = true;
    for (int i = 0; i < 2; ++i) {
        ok = ok && cudaMalloc(&s->vbuf[i], vb) == cudaSuccess;
        ok = ok && cudaMalloc(&s->rbuf[i], vb) == cudaSuccess;
    }
    if (external_v32) { s->v32 = external_v32; s->own_v32 = false; }
    else ok = ok && cudaMalloc(&s->v32, (size_t)(n + 4) * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->alpha, (size_t)cap * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->traj, (size_t)cap * 4 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->scratch, (size_t)pi_scratch_doubles() * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d, sizeof(PiDev)) == cudaSuccess;
    if (!ok) {
        set_error("b2s_pi_create: out of device memory: %s", cudaGetErrorString(cudaGetLastError()));
        b2s_pi_destroy(s);
        return -2;
    }
    cudaMemset(s->d, 0, sizeof(PiDev));
    *out = s;
    return 0;
}

int b2s_pi_create(int64_t n, int32_t cap, int32_t device, b2s_pistate** out) {
    return pi_create_impl(n, cap, device, nullptr, out);
}

int b2s_pi_destroy(b2s_pistate* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) { cudaFree(s->vbuf[i]); cudaFree(s->rbuf[i]); }
    if (s->own_v32) cudaFree(s->v32);
    cudaFree(s->alpha); cudaFree(s->traj); cudaFree(s->scratch); cudaFree(s->d);
    delete s;
    return 0;
}

int b2s_pi_reset(b2s_pistate* s, const double* d_v0, const b2s_power_cfg* cfg, void* stream) {
    if (!s || !d_v0 || !cfg) { set_error("b2s_pi_reset: null argument"); return -1; }
    B2S_CUDA(cudaSetDevice(s->device));
    return pi_reset_impl(s, d_v0, cfg, (cudaStream_t)stream);
}
const float* b2s_pi_v32(const b2s_pistate* s) { return s ? s->v32 : nullptr; }

int b2s_pi_step(b2s_pistate* s, const float* d_hv, void* stream) {
    if (!s || !d_hv) { set_error("b2s_pi_step: null argument"); return -1; }
    return launch_pi_step((cudaStream_t)stream, s->d, s->n, d_hv);
}

const double* b2s_pi_residual(b2s_pistate* s, void* stream) {
    if (!s) return nullptr;
    PiDev h;
    if (cudaMemcpyAsync(&h, s->d, sizeof(PiDev), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
        set_error("b2s_pi_residual: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return s->rbuf[h.r_last];
}

int b2s_pi_precond_update(b2s_pistate* s, const double* d_Tr, void* stream) {
    if (!s || !d_Tr) { set_error("b2s_pi_precond_update: null argument"); return -1; }
    return launch_pi_precond_update((cudaStream_t)stream, s->d, s->n, d_Tr);
}

int b2s_pi_done(b2s_pistate* s, void* stream) {
    if (!s) return -1;
    int done = 0;
    B2S_CUDA(cudaMemcpyAsync(&done, &s->d->done, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    B2S_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return done;
}

int b2s_pi_result(b2s_pistate* s, b2s_power_result* r, double* h_traj, double* d_v_out, void* stream) {
    if (!s) { set_error("b2s_pi_result: null argument"); return -1; }
    return pi_result_impl(s, r, h_traj, d_v_out, (cudaStream_t)stream);
}

static int kfac_apply_impl(b2s_plan* p, const double* d_r, double* d_out);

namespace b2s {
__global__ void loop_condition_kernel(cudaGraphConditionalHandle handle, const PiDev* S) {
    cudaGraphSetConditional(handle, S->done ? 0u : 1u);
}
}  // namespace b2s

// The whole eigen-iteration as ONE graph launch: a WHILE conditional node whose body is the order-1 pass plus
// the two vector kernels; the last kernel of the body feeds the device-side `done` flag to the condition.  No host
// round trip per iteration (the reference does >= 5 `.item()` syncs per iteration, opt.py:457-464) and no work
// after convergence.  Returns 0 = ran, 1 = not available here (caller uses the polled loop), < 0 = error.
static int power_loop_on_device(b2s_plan* p, b2s_pistate* s) {
    if (p->device_loop < 0) {
        const char* e = getenv("B2S_DEVICE_LOOP");
        p->device_loop = (e && atoi(e) == 0) || !p->use_graphs || p->comm ? 0 : 1;
        if (!p->device_loop) return 1;
    }
    cudaStream_t st = p->stream;
    auto it = p->loops.find(p->batch);
    if (it != p->loops.end() && (it->second.loss_scale != p->loss_scale || it->second.global_batch != p->global_batch)) {
        if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
        p->loops.erase(it);
        it = p->loops.end();
    }
    if (it == p->loops.end()) {
        cudaGraph_t graph = nullptr;
        b2s_plan::LoopGraph lg;
        bool ok = cudaGraphCreate(&graph, 0) == cudaSuccess;
        cudaGraphConditionalHandle handle{};
        ok = ok && cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
        cudaGraphNodeParams np{};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node = nullptr;
        ok = ok && cudaGraphAddNode(&node, graph, nullptr, 0, &np) == cudaSuccess;
        int rc = 0;
        if (ok) {
            cudaGraph_t body = np.conditional.phGraph_out[0];
            g_counting_paused = true;
            g_capture_count = 0;
            ok = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                rc = run_pass_eager(p, 1);
                if (rc == 0) rc = launch_pi_step(st, s->d, s->n, p->out32[1]);
                if (rc == 0) {
                    loop_condition_kernel<<<1, 1, 0, st>>>(handle, s->d);
                    count_launch();
                }
                cudaGraph_t out = nullptr;
                ok = cudaStreamEndCapture(st, &out) == cudaSuccess && rc == 0;
            }
            g_counting_paused = false;
            lg.kernels_per_iter = g_capture_count;
        }
        ok = ok && cudaGraphInstantiate(&lg.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            // conditional nodes unavailable (driver) or a node kind the body does not admit: polled loop from now on
            cudaGetLastError();
            p->device_loop = 0;
            if (rc < 0 && rc != -2) return rc;
            // the failed capture may have left the pass half enqueued nowhere; the state was not advanced
            return 1;
        }
        lg.loss_scale = p->loss_scale;
        lg.global_batch = p->global_batch;
        it = p->loops.emplace(p->batch, lg).first;
    }
    B2S_CUDA(cudaGraphLaunch(it->second.exec, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    int last = -1;
    B2S_CUDA(cudaMemcpy(&last, &s->d->last_iter, sizeof(int), cudaMemcpyDeviceToHost));
    count_launch((int)(it->second.kernels_per_iter * (last + 1)));
    return 0;
}

// Host-driven loop (data parallel runs, eager mode): the `done` flag is read back with a lag of a few iterations so
// that the device queue never drains on short iterations; every kernel of the loop is gated on the flag, so the
// iterations enqueued past convergence change nothing.
static int power_loop_polled(b2s_plan* p, b2s_pistate* s, int max_iter) {
    cudaStream_t st = p->stream;
    if (!p->h_done) {
        B2S_CUDA(cudaHostAlloc(&p->h_done, (b2s_plan::kMaxLag + 1) * sizeof(int), cudaHostAllocDefault));
        for (auto& e : p->ev_poll) B2S_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    int lag = p->poll_lag;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (p->comm) {
        // data parallel: every rank must enqueue the same number of passes (they contain collectives), so the loop is
        // synchronous by default.  The all-reduced Hv and the rank-ordered BatchNorm sums are bitwise identical on all
        // ranks and the vector kernels are deterministic, so a lag would be legal (B2S_POLL_LAG_DP=1) -- measured at
        // 2 GPUs: 2.679 against 2.683 ms per step, no gain, so the conservative form stays.
        static const int lag_dp = getenv("B2S_POLL_LAG_DP") ? std::max(0, std::min((int)b2s_plan::kMaxLag, atoi(getenv("B2S_POLL_LAG_DP")))) : 0;
        lag = lag_dp;
    } else if (lag < 0) {
        if (const char* e = getenv("B2S_POLL_LAG")) p->poll_lag = std::max(0, std::min((int)b2s_plan::kMaxLag, atoi(e)));
        lag = 0;                                   // first call: synchronous, and the first iteration is timed
        if (p->poll_lag < 0) {
            cudaEventCreate(&t0); cudaEventCreate(&t1);
        }
    }
    const int ring = b2s_plan::kMaxLag + 1;
    for (int it = 0; it < max_iter; ++it) {
        if (t0 && it == 1) cudaEventRecord(t0, st);
        B2S_TRY(run_pass(p, 1));
        B2S_TRY(launch_pi_step(st, s->d, s->n, p->out32[1]));
        if (t0 && it == 1) cudaEventRecord(t1, st);
        B2S_CUDA(cudaMemcpyAsync(&p->h_done[it % ring], &s->d->done, sizeof(int), cudaMemcpyDeviceToHost, st));
        B2S_CUDA(cudaEventRecord(p->ev_poll[it % ring], st));
        if (it >= lag) {
            const int w = (it - lag) % ring;
            B2S_CUDA(cudaEventSynchronize(p->ev_poll[w]));
            if (p->h_done[w]) break;
        }
    }
    if (t0) {
        float ms = 0.f;
        if (cudaEventSynchronize(t1) == cudaSuccess && cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess && ms > 0.f)
            p->poll_lag = ms < 0.03f ? 4 : ms < 0.12f ? 2 : ms < 0.5f ? 1 : 0;     // keep ~100 us of work queued
        cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
    return 0;
}


int b2s_power_iterate(b2s_plan* p, double* d_v, const b2s_power_cfg* cfg, b2s_power_result* h_result,
                      double* h_traj) {
    if (!p || !d_v || !cfg) { set_error("b2s_power_iterate: null argument"); return -1; }
    if (p->batch <= 0) { set_error("b2s_power_iterate: call b2s_base_pass first"); return -1; }
    if (cfg->precond && p->kfac.empty()) {
        set_error("b2s_power_iterate: precond = 1 but no K-FAC factors are installed (b2s_kfac_set)");
        return -1;
    }
    B2S_TRY(enter(p));
    if (cfg->precond && !p->kfac_tr) B2S_CUDA(cudaMalloc(&p->kfac_tr, (size_t)(p->P + 4) * sizeof(double)));
    if (!p->pi || p->pi->cap < cfg->max_iter) {
        if (p->pi) { b2s_pi_destroy(p->pi); p->pi = nullptr; }
        const int cap = cfg->max_iter > 1024 ? cfg->max_iter : 1024;
        B2S_TRY(pi_create_impl(p->P, cap, p->device, p->v32, &p->pi));
    }
    b2s_pistate* s = p->pi;
    cudaStream_t st = p->stream;
    B2S_TRY(pi_reset_impl(s, d_v, cfg, st));
    B2S_TRY(alloc_order(p, 1));
    if (cfg->precond) {
        int done = cfg->max_iter <= 0;
        while (!done) {
            // the HVP reads p->v32, which the previous iteration's update wrote (the state aliases it)
            B2S_TRY(run_pass(p, 1));
            B2S_TRY(launch_pi_step(st, s->d, s->n, p->out32[1]));
            PiDev h;                                              // v <- normalise(v + alpha T r)   opt.py:491-498
            B2S_CUDA(cudaMemcpyAsync(&h, s->d, sizeof(PiDev), cudaMemcpyDeviceToHost, st));
            B2S_CUDA(cudaStreamSynchronize(st));
            done = h.done;
            if (!done) {
                B2S_TRY(kfac_apply_impl(p, s->rbuf[h.r_last], p->kfac_tr));
                B2S_TRY(launch_pi_precond_update(st, s->d, s->n, p->kfac_tr));
                done = h.iter >= h.max_iter;                      // what pi_precond_commit_kernel decides
            }
        }
    } else if (cfg->max_iter > 0) {
        int rc = 1;
        if (p->device_loop != 0) rc = power_loop_on_device(p, s);
        if (rc < 0) return rc;
        if (rc == 1) B2S_TRY(power_loop_polled(p, s, cfg->max_iter));
    }
    p->v_is_current = false;
    B2S_TRY(pi_result_impl(s, h_result, h_traj, d_v, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    return leave(p);
}

// ---- K-FAC preconditioner ---------------------------------------------------------------------------
static int kfac_dims(const b2s_plan* p, int op_index, int* da, int* dg, int* dw) {
    if (!p || op_index < 0 || op_index >= (int)p->ops.size() || p->ops[op_index].kind != B2S_OP_CONV) {
        set_error("b2s_kfac: op %d is not a Conv2d/Linear op", op_index);
        return -1;
    }
    const b2s_op& op = p->ops[op_index];
    const int w = p->tensors[op.in].C * op.kh * op.kw;
    if (dw) *dw = w;
    if (da) *da = w + (op.b_off >= 0 ? 1 : 0);
    if (dg) *dg = p->tensors[op.out].C;
    return 0;
}

int b2s_kfac_dims(const b2s_plan* p, int32_t op_index, int32_t* dim_a, int32_t* dim_g) {
    return kfac_dims(p, op_index, dim_a, dim_g, nullptr);
}

int b2s_kfac_build(b2s_plan* p, int32_t op_a, int32_t op_g, float* d_A, float* d_G) {
    int da = 0, dg = 0;
    B2S_TRY(kfac_dims(p, op_a, &da, nullptr, nullptr));
    B2S_TRY(kfac_dims(p, op_g, nullptr, &dg, nullptr));
    if (!d_A || !d_G) { set_error("b2s_kfac_build: null output"); return -1; }
    if (p->batch <= 0) { set_error("b2s_kfac_build: call b2s_base_pass first"); return -1; }
    B2S_TRY(enter(p));
    cudaStream_t st = p->stream;
    // data parallel: the factors are batch means, so every rank adds its samples' share and 1/world of
    // the identity term, then one all-reduce per factor; eigh and the apply stay replicated
    const long long gb = p->global_batch > 0 ? p->global_batch : (long long)p->batch * p->world;
    {   // A = 0.95 I + 0.05 a^T a / B, a = patches / spatial with a ones column (kfac.py:292-311, 52-58)
        const b2s_op& op = p->ops[op_a];
        const b2s_tensor& I = p->tensors[op.in];
        const b2s_tensor& O = p->tensors[op.out];
        const float S = (float)(O.H * O.W);
        B2S_TRY(launch_gram(st, tptr(p, p->fw, 0, op.in), I.sample_stride, p->batch, I.C, I.H, I.W, O.H, O.W, op.kh, op.kw,
                            op.sh, op.sw, op.ph, op.pw, op.b_off >= 0 ? 1 : 0, 1.f / S, 1.f / S,
                            0.05f / (float)gb, 0.95f / (float)p->world, d_A));
        if (p->comm) B2S_TRY(comm_allreduce_f32(p->comm, d_A, (long long)da * da, st));
    }
    {   // G = 0.95 I + 0.05 * B*S * g^T g  (kfac.py:337-367 with batch_averaged=True, 60-65)
        const b2s_op& op = p->ops[op_g];
        const b2s_tensor& O = p->tensors[op.out];
        const float S = (float)(O.H * O.W);
        B2S_TRY(launch_gram(st, tptr(p, p->bw, 0, op.out), O.sample_stride, p->batch, O.C, O.H, O.W, O.H, O.W, 1, 1, 1, 1,
                            0, 0, 0, 1.f, 0.f, 0.05f * (float)gb * S, 0.95f / (float)p->world, d_G));
        if (p->comm) B2S_TRY(comm_allreduce_f32(p->comm, d_G, (long long)dg * dg, st));
    }
    return leave(p);
}

int b2s_kfac_clear(b2s_plan* p) {
    if (!p) return -1;
    p->kfac.clear();
    return 0;
}

int b2s_kfac_set(b2s_plan* p, int32_t op_index, const float* d_Ainv, const float* d_Ginv) {
    b2s_plan::KfacLayer L;
    B2S_TRY(kfac_dims(p, op_index, &L.da, &L.dg, &L.dw));
    if (!d_Ainv || !d_Ginv) { set_error("b2s_kfac_set: null matrix"); return -1; }
    L.op = op_index; L.Ainv = d_Ainv; L.Ginv = d_Ginv;
    for (auto& e : p->kfac)
        if (p->ops[e.op].w_off == p->ops[op_index].w_off) { e = L; return 0; }
    p->kfac.push_back(L);
    return 0;
}

static int kfac_apply_impl(b2s_plan* p, const double* d_r, double* d_out) {
    cudaStream_t st = p->stream;
    long long need = 0;
    for (auto& L : p->kfac) need = std::max(need, (long long)L.dg * L.da);
    if (need > p->kfac_m_elems) {
        B2S_CUDA(cudaStreamSynchronize(st));
        cudaFree(p->kfac_m);
        B2S_CUDA(cudaMalloc(&p->kfac_m, (size_t)need * 2 * sizeof(float)));
        p->kfac_m_elems = need;
    }
    if (d_out != d_r)
        B2S_CUDA(cudaMemcpyAsync(d_out, d_r, (size_t)p->P * sizeof(double), cudaMemcpyDeviceToDevice, st));
    for (auto& L : p->kfac) {
        const b2s_op& op = p->ops[L.op];
        float* M = p->kfac_m;
        float* T = p->kfac_m + p->kfac_m_elems;
        B2S_TRY(launch_kfac_gather(st, d_r, op.w_off, op.b_off, L.dg, L.dw, M));                 // opt.py:400-407
        B2S_TRY(launch_sgemm_small(st, L.dg, L.da, L.dg, L.Ginv, L.dg, 1, M, L.da, 1, T, L.da));  // G^-1 R
        B2S_TRY(launch_sgemm_small(st, L.dg, L.da, L.da, T, L.da, 1, L.Ainv, L.da, 1, M, L.da));  // (.) A^-1
        B2S_TRY(launch_kfac_scatter(st, M, op.w_off, op.b_off, L.dg, L.dw, d_out));               // opt.py:409-411
    }
    return 0;
}

int b2s_kfac_apply(b2s_plan* p, const double* d_r, double* d_out) {
    if (!p || !d_r || !d_out) { set_error("b2s_kfac_apply: null argument"); return -1; }
    B2S_TRY(enter(p));
    B2S_TRY(kfac_apply_impl(p, d_r, d_out));
    return leave(p);
}

// ---- data parallelism ---------------------------------------------------------------------------
int b2s_step_assemble(const double* d_gradf, const double* d_gradrho, double coef, int64_t n, double* d_p, float* d_p32,
                      void* stream) {
    if (!d_gradf || !d_p32 || n <= 0) { set_error("b2s_step_assemble: null argument"); return -1; }
    return launch_step_assemble((cudaStream_t)stream, d_gradf, d_gradrho, coef, (long long)n, d_p, d_p32);
}

int b2s_clip_norm(const double* d_x, int64_t n, double clip, double* d_scratch, double* d_out2, void* stream) {
    if (!d_x || !d_scratch || !d_out2 || n <= 0) { set_error("b2s_clip_norm: null argument"); return -1; }
    return launch_clip_norm((cudaStream_t)stream, d_x, (long long)n, clip, d_scratch, d_out2);
}

int b2s_clip_scratch_doubles(void) { return pi_scratch_doubles() + 2; }

int b2s_step_fused(const double* d_gradf, double* d_gradrho, double coef, const double* d_scale2, int64_t n, double* d_p,
                   float* d_p32, float* d_params, float* d_state1, float* d_state2, const b2s_step_opt* opt, void* stream) {
    if (!d_gradf || !d_p32 || !opt || n <= 0) { set_error("b2s_step_fused: null argument"); return -1; }
    if (opt->kind < 0 || opt->kind > 2) { set_error("b2s_step_fused: optimizer kind %d", opt->kind); return -1; }
    if (opt->kind != 0 && !d_params) { set_error("b2s_step_fused: optimizer update without a parameter vector"); return -1; }
    if ((opt->kind == 1 && opt->momentum != 0.0 && !d_state1) || (opt->kind == 2 && (!d_state1 || !d_state2))) {
        set_error("b2s_step_fused: optimizer state vector missing");
        return -1;
    }
    StepOpt o{};
    o.kind = opt->kind; o.first = opt->first_step; o.nesterov = opt->nesterov; o.maximize = opt->maximize;
    o.write_gradrho = opt->write_gradrho;
    o.lr = (float)opt->lr; o.momentum = (float)opt->momentum; o.dampening = (float)opt->dampening;
    o.weight_decay = (float)opt->weight_decay;
    o.beta1 = (float)opt->beta1; o.beta2 = (float)opt->beta2; o.eps = (float)opt->eps;
    o.step_size = (float)opt->step_size; o.bias2_sqrt = (float)opt->bias2_sqrt;
    return launch_step_fused((cudaStream_t)stream, d_gradf, d_gradrho, coef, d_scale2, (long long)n, d_p, d_p32, d_params,
                             d_state1, d_state2, o);
}

int b2s_plan_set_global_batch(b2s_plan* p, int64_t global_batch) {
    if (!p || global_batch < 0) { set_error("b2s_plan_set_global_batch: invalid arguments"); return -1; }
    p->global_batch = global_batch;              // graphs captured with another count are re-captured on use
    return 0;
}

int b2s_comm_unique_id(void* h_id128) { return comm_unique_id(h_id128); }
int b2s_comm_init(b2s_plan* p, const void* h_id128, int32_t rank, int32_t world) {
    if (!p || !h_id128) { set_error("b2s_comm_init: null argument"); return -1; }
    B2S_CUDA(cudaSetDevice(p->device));
    if (p->comm) { comm_destroy(p->comm); p->comm = nullptr; }
    drop_graphs(p);
    p->world = world;
    if (world <= 1) return 0;
    return comm_init(&p->comm, h_id128, rank, world);
}
int b2s_comm_peer_local(b2s_plan* p, void* h_handle64) {
    if (!p || !h_handle64) { set_error("b2s_comm_peer_local: null argument"); return -1; }
    if (!p->comm) { set_error("b2s_comm_peer_local: call b2s_comm_init first"); return -1; }
    B2S_CUDA(cudaSetDevice(p->device));
    return comm_peer_local(p->comm, h_handle64);
}
int b2s_comm_peer_attach(b2s_plan* p, const void* h_handles) {
    if (!p || !h_handles) { set_error("b2s_comm_peer_attach: null argument"); return -1; }
    if (!p->comm) { set_error("b2s_comm_peer_attach: call b2s_comm_init first"); return -1; }
    B2S_CUDA(cudaSetDevice(p->device));
    drop_graphs(p);
    return comm_peer_attach(p->comm, h_handles);
}
int b2s_comm_peer_ready(const b2s_plan* p) { return (p && p->comm) ? comm_peer_ready(p->comm) : 0; }
int b2s_comm_peer_disable(b2s_plan* p) {
    if (!p || !p->comm) return 0;
    drop_graphs(p);
    comm_peer_disable(p->comm);
    return 0;
}
int b2s_comm_destroy(b2s_plan* p) {
    if (!p) return 0;
    if (p->comm) { comm_destroy(p->comm); p->comm = nullptr; }
    p->world = 1;
    drop_graphs(p);
    return 0;
}

}  // extern "C"
