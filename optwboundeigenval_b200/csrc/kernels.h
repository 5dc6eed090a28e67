// kernels.h -- host-side launchers of the hand-written sm_100a kernels (internal).
#pragma once
#include "common.cuh"

namespace b2s {

constexpr int kMaxPairs = 3;

// Geometry of one Conv2d / Linear op. The same struct drives the forward gather
// (y = conv(x, W)), the input-adjoint gather (xbar = conv^T(ybar, W)) and the
// weight-gradient contraction (Wbar = corr(x, ybar)).
struct ConvGeom {
    int batch;
    int Cin, H, W;          // conv input
    long long in_sstride;
    int Cout, OH, OW;       // conv output
    long long out_sstride;
    int KH, KW, sh, sw, ph, pw;
};

// out (+)= sum_p scale[p] * conv(act[p], wt[p]) (+ bias) (relu / relu-mask)
//   relu_mode 0: none, 1: out = max(out, 0), 2: out = relu_ref > 0 ? out : 0
int launch_conv_fwd(cudaStream_t st, const ConvGeom& g, int npairs, const float* const* act,
                    const float* const* wt, const float* scale, const float* bias, int relu_mode,
                    const float* relu_ref, float* out, int accumulate, const float* const* pack = nullptr);
// xbar (+)= sum_p scale[p] * conv^T(adj[p], wt[p])
int launch_conv_dgrad(cudaStream_t st, const ConvGeom& g, int npairs, const float* const* adj,
                      const float* const* wt, const float* scale, float* out, int accumulate,
                      const float* const* pack = nullptr);
// wbar += sum_p scale[p] * corr(act[p], adj[p])   (atomic accumulation into the flat vector)
int launch_conv_wgrad(cudaStream_t st, const ConvGeom& g, int npairs, const float* const* act,
                      const float* const* adj, const float* scale, float* wbar);
// bbar[c] += sum_{n,pix} adj[n,c,pix]
int launch_bias_grad(cudaStream_t st, const float* adj, int batch, int C, int HW, long long sstride,
                     float* bbar);

// ---- elementwise / pooling (elementwise.cu) ---------------------------------------------
struct View {               // NCHW view: element (n,c,h,w) at p[n*sstride + (c*H + h)*W + w]
    float* p;
    int C, H, W;
    long long sstride;
};
int launch_relu_fwd(cudaStream_t st, int order, const View& x0, const View& xk, const View& yk, int batch);
// adj_in (+)= adj_out * (x0 > 0)
int launch_relu_bwd(cudaStream_t st, const View& ref0, const View& adj_out, const View& adj_in, int batch,
                    int accumulate);
// in place: adj *= (ref0 > 0)
int launch_mask_inplace(cudaStream_t st, const View& ref0, const View& adj, int batch);
int launch_maxpool_fwd(cudaStream_t st, int order, const View& xk, const View& yk, int32_t* argmax, int batch,
                       int kh, int kw, int sh, int sw, int ph, int pw);
int launch_maxpool_bwd(cudaStream_t st, const View& adj_out, const View& adj_in, const int32_t* argmax,
                       int batch);   // atomic scatter-add, adj_in must be initialised
int launch_avgpool_fwd(cudaStream_t st, const View& x, const View& y, int batch, int k);
int launch_avgpool_bwd(cudaStream_t st, const View& adj_out, const View& adj_in, int batch, int k,
                       int accumulate);
int launch_copy_view(cudaStream_t st, const View& src, const View& dst, int batch, int accumulate);
int launch_zero_view(cudaStream_t st, const View& v, int batch);
// residual connection (B2S_OP_ADD): y_k = a_k + b_k with an optional fused ReLU decided on order 0
int launch_add_fwd(cudaStream_t st, int order, int relu, const View& a, const View& b, const View& y0, const View& yk, int batch);
int launch_add_bwd(cudaStream_t st, const View* y0_or_null, const View& g, const View* ga, int acc_a, const View* gb, int acc_b,
                   int batch);
int launch_cast_f64_f32(cudaStream_t st, const double* in, float* out, long long n);
int launch_cast_f32_f64(cudaStream_t st, const float* in, double* out, long long n, double scale);

// ---- batch norm (bn.cu) ---------------------------------------------------------------
struct BnArgs {
    int batch, C, HW;
    long long in_sstride, out_sstride;
    const float* x[3];        // input jets (orders 0..2; higher ones NULL when unused)
    const float* y0;          // order-0 output (post-ReLU when fused) for the mask
    float* yk;                // forward output of the current order
    const float* g[3];        // adjoints of the output (orders 0..2)
    float* xbar;              // input adjoint of the current order
    const float* gamma; const float* beta;     // parameters
    const float* vgamma; const float* vbeta;   // tangent direction slices (order >= 1)
    double* fsum[3];          // forward sums per order, each [2][C]: {sum x_k, sum (c*c)_k}
    double* bsum[3];          // backward sums per order, each [2][C]: {sum g_k, sum (g*xh)_k}
    float* out_gamma; float* out_beta;         // where gamma-bar / beta-bar of this order go (flat vector)
    float* running_mean; float* running_var;   // updated by the order-0 forward (may be NULL)
    float eps, momentum;
    int relu;
    int first;                // input has no tangent (network data)
    long long count;          // elements per channel over the GLOBAL batch (batch*HW on one GPU)
    int accumulate;
    float pgrad_scale;        // 1/world when the sums were all-reduced (the flat vector is summed again later)
    double* csum;             // third-order compatibility sweep: [bn_corr_sums()][C]
    const struct PeerCtx* peer;   // multi-GPU: exchange context for the in-kernel all-reduce of the sums (peer.cuh), else NULL
    const struct PeerCtx* peer_tail;   // multi-GPU: the statistics kernels' last block all-reduces the sums, else NULL
    int peer_ll;              // with peer: 1 = per-channel packet exchange, no grid barrier (peer_exchange_channel); 2 = its
                              // single-GPU form (per-channel barrier); 0 = block 0 exchanges between two grid barriers
};
int launch_bn_fwd_stats(cudaStream_t st, int order, const BnArgs& a);
int launch_bn_fwd_apply(cudaStream_t st, int order, const BnArgs& a);
int launch_bn_bwd_stats(cudaStream_t st, int order, const BnArgs& a);
int launch_bn_bwd_apply(cudaStream_t st, int order, const BnArgs& a);
int launch_bn_eval(cudaStream_t st, const BnArgs& a);     // evaluation mode: running statistics, no update
// fused statistics+apply (cooperative launch); return 1 = not applicable, use the two-kernel form
int launch_bn_fwd_fused(cudaStream_t st, int order, const BnArgs& a, int do_stats);
int launch_bn_bwd_fused(cudaStream_t st, int order, const BnArgs& a);
// reference-compatible third order (see bn.cu): first-order sweep with the dropped adjoint injected
int bn_corr_sums();
int launch_bn_corr_stats(cudaStream_t st, const BnArgs& a, const float* gc);
int launch_bn_corr_apply(cudaStream_t st, const BnArgs& a, const float* gc, float* xbar);
int launch_sub_cast_f32_f64(cudaStream_t st, const float* a, const float* b, double* out, long long n);

// ---- loss heads (head.cu) -------------------------------------------------------------
struct HeadArgs {
    int kind, batch, C;
    const float* z[3];        // logits jets, [batch, C] with sample stride zs
    long long zs;
    const long long* labels;  // CE heads
    const float* target;      // WBCE heads [batch, C]
    const float* coef;        // WBCE heads [batch, C]
    double loss_scale;        // CE: 1/global_batch
    float* zbar;              // adjoint of the logits of the current order, stride zs
    double* loss;             // order 0 only
};
int launch_head(cudaStream_t st, int order, const HeadArgs& a);

// ---- K-FAC preconditioner (kfac.cu) ------------------------------------------------------------
int launch_gram(cudaStream_t st, const float* src, long long sstride, int batch, int C, int H, int W, int OH, int OW,
                int KH, int KW, int sh, int sw, int ph, int pw, int has_ones, float elem_scale, float ones_value,
                float scale, float diag_add, float* out);
int launch_sgemm_small(cudaStream_t st, int M, int N, int K, const float* A, int a_rs, int a_cs, const float* B,
                       int b_rs, int b_cs, float* C, int ldc);
int launch_kfac_gather(cudaStream_t st, const double* r, long long w_off, long long b_off, int dg, int dw, float* M);
int launch_kfac_scatter(cudaStream_t st, const float* M, long long w_off, long long b_off, int dg, int dw, double* out);

}  // namespace b2s
