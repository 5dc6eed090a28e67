// elementwise.cu -- ReLU / pooling jets, casts and view utilities (memory-bound kernels).
//
// ReLU and MaxPool are piecewise linear: phi'' = phi''' = 0, so every jet order is routed by the
// mask / argmax fixed on the order-0 pass (SURVEY.md Appendix B, section 7 "tie-breaking").
// AvgPool is linear.  grid = (chunks over C*H*W, batch): no integer division per element.
#include "kernels.h"

namespace b2s {

static inline dim3 ew_grid(long long per_sample, int batch, int threads = 256, int per_thread = 4) {
    long long blocks = (per_sample + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
    if (blocks < 1) blocks = 1;
    if (blocks > 65535) blocks = 65535;
    return dim3((unsigned)blocks, (unsigned)batch);
}

// y_k = order==0 ? max(x0,0) : (x0>0 ? x_k : 0)
__global__ void relu_fwd_kernel(int order, const float* __restrict__ x0, long long x0s,
                                const float* __restrict__ xk, long long xks, float* __restrict__ yk,
                                long long ys, long long per) {
    const int n = blockIdx.y;
    x0 += n * x0s; xk += n * xks; yk += n * ys;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const float r = x0[i];
        yk[i] = r > 0.f ? (order == 0 ? r : xk[i]) : 0.f;
    }
}

__global__ void relu_bwd_kernel(const float* __restrict__ ref, long long rs, const float* __restrict__ go,
                                long long gos, float* __restrict__ gi, long long gis, long long per, int acc) {
    const int n = blockIdx.y;
    ref += n * rs; go += n * gos; gi += n * gis;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        float v = ref[i] > 0.f ? go[i] : 0.f;
        if (acc) v += gi[i];
        gi[i] = v;
    }
}

__global__ void mask_inplace_kernel(const float* __restrict__ ref, long long rs, float* __restrict__ g,
                                    long long gs, long long per) {
    const int n = blockIdx.y;
    ref += n * rs; g += n * gs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        if (!(ref[i] > 0.f)) g[i] = 0.f;
    }
}

// residual connection y = a + b with an optional fused ReLU: order 0 -> max(a0 + b0, 0), order k -> (y0 > 0) (a_k + b_k)
__global__ void add_fwd_kernel(int order, int relu, const float* __restrict__ a, long long as, const float* __restrict__ b,
                               long long bs, const float* __restrict__ y0, long long y0s, float* __restrict__ yk, long long ys,
                               long long per) {
    const int n = blockIdx.y;
    a += n * as; b += n * bs; y0 += n * y0s; yk += n * ys;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        float v = a[i] + b[i];
        if (relu) v = order == 0 ? fmaxf(v, 0.f) : (y0[i] > 0.f ? v : 0.f);
        yk[i] = v;
    }
}

// both operand adjoints (+)= (y0 > 0) g ; a NULL target is skipped (network data has no adjoint)
__global__ void add_bwd_kernel(const float* __restrict__ y0, long long y0s, const float* __restrict__ g, long long gs,
                               float* __restrict__ ga, long long gas, int acc_a, float* __restrict__ gb, long long gbs, int acc_b,
                               long long per) {
    const int n = blockIdx.y;
    if (y0) y0 += n * y0s;
    g += n * gs;
    if (ga) ga += n * gas;
    if (gb) gb += n * gbs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const float v = (!y0 || y0[i] > 0.f) ? g[i] : 0.f;
        if (ga) ga[i] = acc_a ? ga[i] + v : v;
        if (gb) gb[i] = acc_b ? gb[i] + v : v;
    }
}

// order 0: scan the window in (ky,kx) order, first maximum wins (ATen max_pool2d), padding = -inf;
// order>0: y_k = x_k[argmax]
__global__ void maxpool_fwd_kernel(int order, const float* __restrict__ x, long long xs, float* __restrict__ y,
                                   long long ys, int32_t* __restrict__ arg, int C, int H, int W, int OH, int OW,
                                   int kh, int kw, int sh, int sw, int ph, int pw) {
    const int n = blockIdx.y;
    const long long per = (long long)C * OH * OW;
    x += n * xs; y += n * ys; arg += n * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW);
        const long long t = i / OW;
        const int oy = (int)(t % OH);
        const int c = (int)(t / OH);
        const float* xc = x + (long long)c * H * W;
        if (order == 0) {
            float best = -INFINITY;
            int bi = -1;
            for (int ky = 0; ky < kh; ++ky) {
                const int iy = oy * sh - ph + ky;
                if (iy < 0 || iy >= H) continue;
                for (int kx = 0; kx < kw; ++kx) {
                    const int ix = ox * sw - pw + kx;
                    if (ix < 0 || ix >= W) continue;
                    const float v = xc[iy * W + ix];
                    if (bi < 0 || v > best || v != v) { best = v; bi = iy * W + ix; }
                }
            }
            y[i] = best;
            arg[i] = bi;
        } else {
            const int bi = arg[i];
            y[i] = bi >= 0 ? xc[bi] : 0.f;
        }
    }
}

__global__ void maxpool_bwd_kernel(const float* __restrict__ go, long long gos, float* __restrict__ gi,
                                   long long gis, const int32_t* __restrict__ arg, int C, int HW, int OHW) {
    const int n = blockIdx.y;
    const long long per = (long long)C * OHW;
    go += n * gos; gi += n * gis; arg += n * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / OHW);
        const int bi = arg[i];
        if (bi >= 0) atomicAdd(gi + (long long)c * HW + bi, go[i]);
    }
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, long long xs, float* __restrict__ y,
                                   long long ys, int C, int H, int W, int OH, int OW, int k) {
    const int n = blockIdx.y;
    const long long per = (long long)C * OH * OW;
    x += n * xs; y += n * ys;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW);
        const long long t = i / OW;
        const int oy = (int)(t % OH);
        const int c = (int)(t / OH);
        const float* xc = x + ((long long)c * H + oy * k) * W + ox * k;
        float s = 0.f;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) s += xc[ky * W + kx];
        y[i] = s * inv;
    }
}

// gi[c,iy,ix] (+)= go[c, iy/k, ix/k] / k^2 (zero outside the pooled region)
__global__ void avgpool_bwd_kernel(const float* __restrict__ go, long long gos, float* __restrict__ gi,
                                   long long gis, int C, int H, int W, int OH, int OW, int k, int acc) {
    const int n = blockIdx.y;
    const long long per = (long long)C * H * W;
    go += n * gos; gi += n * gis;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % W);
        const long long t = i / W;
        const int iy = (int)(t % H);
        const int c = (int)(t / H);
        const int oy = iy / k, ox = ix / k;
        float v = (oy < OH && ox < OW) ? go[((long long)c * OH + oy) * OW + ox] * inv : 0.f;
        if (acc) v += gi[i];
        gi[i] = v;
    }
}

__global__ void copy_view_kernel(const float* __restrict__ src, long long ss, float* __restrict__ dst,
                                 long long ds, long long per, int acc) {
    const int n = blockIdx.y;
    src += n * ss; dst += n * ds;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = acc ? dst[i] + src[i] : src[i];
}

__global__ void zero_view_kernel(float* __restrict__ dst, long long ds, long long per) {
    const int n = blockIdx.y;
    dst += n * ds;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = 0.f;
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}
__global__ void cast_f32_f64_kernel(const float* __restrict__ in, double* __restrict__ out, long long n,
                                    double scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = (double)in[i] * scale;
}

__global__ void sub_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ out,
                                long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = (double)(a[i] - b[i]);
}

static inline long long per_sample(const View& v) { return (long long)v.C * v.H * v.W; }

int launch_relu_fwd(cudaStream_t st, int order, const View& x0, const View& xk, const View& yk, int batch) {
    ProfScope prof("relu_fwd", 0.0, 8.0 * batch * (double)yk.C * yk.H * yk.W, st);
    const long long per = per_sample(yk);
    relu_fwd_kernel<<<ew_grid(per, batch), 256, 0, st>>>(order, x0.p, x0.sstride, xk.p, xk.sstride, yk.p,
                                                          yk.sstride, per);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_relu_bwd(cudaStream_t st, const View& ref0, const View& adj_out, const View& adj_in, int batch,
                    int accumulate) {
    ProfScope prof("relu_bwd", 0.0, 12.0 * batch * (double)adj_in.C * adj_in.H * adj_in.W, st);
    const long long per = per_sample(adj_in);
    relu_bwd_kernel<<<ew_grid(per, batch), 256, 0, st>>>(ref0.p, ref0.sstride, adj_out.p, adj_out.sstride,
                                                          adj_in.p, adj_in.sstride, per, accumulate);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_mask_inplace(cudaStream_t st, const View& ref0, const View& adj, int batch) {
    ProfScope prof("relu_mask", 0.0, 12.0 * batch * (double)adj.C * adj.H * adj.W, st);
    const long long per = per_sample(adj);
    mask_inplace_kernel<<<ew_grid(per, batch), 256, 0, st>>>(ref0.p, ref0.sstride, adj.p, adj.sstride, per);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_maxpool_fwd(cudaStream_t st, int order, const View& xk, const View& yk, int32_t* argmax, int batch,
                       int kh, int kw, int sh, int sw, int ph, int pw) {
    ProfScope prof("maxpool_fwd", 0.0, 4.0 * batch * ((double)xk.C * xk.H * xk.W + 2.0 * yk.C * yk.H * yk.W), st);
    const long long per = per_sample(yk);
    maxpool_fwd_kernel<<<ew_grid(per, batch, 256, 1), 256, 0, st>>>(order, xk.p, xk.sstride, yk.p, yk.sstride,
                                                                     argmax, xk.C, xk.H, xk.W, yk.H, yk.W, kh,
                                                                     kw, sh, sw, ph, pw);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_maxpool_bwd(cudaStream_t st, const View& adj_out, const View& adj_in, const int32_t* argmax,
                       int batch) {
    ProfScope prof("maxpool_bwd", 0.0, 4.0 * batch * ((double)adj_in.C * adj_in.H * adj_in.W + 2.0 * adj_out.C * adj_out.H * adj_out.W), st);
    const long long per = per_sample(adj_out);
    maxpool_bwd_kernel<<<ew_grid(per, batch, 256, 1), 256, 0, st>>>(adj_out.p, adj_out.sstride, adj_in.p,
                                                                     adj_in.sstride, argmax, adj_in.C,
                                                                     adj_in.H * adj_in.W,
                                                                     adj_out.H * adj_out.W);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_avgpool_fwd(cudaStream_t st, const View& x, const View& y, int batch, int k) {
    ProfScope prof("avgpool_fwd", 0.0, 4.0 * batch * ((double)x.C * x.H * x.W + (double)y.C * y.H * y.W), st);
    const long long per = per_sample(y);
    avgpool_fwd_kernel<<<ew_grid(per, batch, 256, 1), 256, 0, st>>>(x.p, x.sstride, y.p, y.sstride, x.C, x.H,
                                                                     x.W, y.H, y.W, k);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_avgpool_bwd(cudaStream_t st, const View& adj_out, const View& adj_in, int batch, int k,
                       int accumulate) {
    ProfScope prof("avgpool_bwd", 0.0, 4.0 * batch * ((double)adj_in.C * adj_in.H * adj_in.W + (double)adj_out.C * adj_out.H * adj_out.W), st);
    const long long per = per_sample(adj_in);
    avgpool_bwd_kernel<<<ew_grid(per, batch), 256, 0, st>>>(adj_out.p, adj_out.sstride, adj_in.p,
                                                             adj_in.sstride, adj_in.C, adj_in.H, adj_in.W,
                                                             adj_out.H, adj_out.W, k, accumulate);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_copy_view(cudaStream_t st, const View& src, const View& dst, int batch, int accumulate) {
    ProfScope prof("copy_view", 0.0, 8.0 * batch * (double)dst.C * dst.H * dst.W, st);
    const long long per = per_sample(dst);
    copy_view_kernel<<<ew_grid(per, batch), 256, 0, st>>>(src.p, src.sstride, dst.p, dst.sstride, per,
                                                           accumulate);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_zero_view(cudaStream_t st, const View& v, int batch) {
    ProfScope prof("zero_view", 0.0, 4.0 * batch * (double)v.C * v.H * v.W, st);
    const long long per = per_sample(v);
    zero_view_kernel<<<ew_grid(per, batch), 256, 0, st>>>(v.p, v.sstride, per);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_cast_f64_f32(cudaStream_t st, const double* in, float* out, long long n) {
    int blocks = cdiv(n, 256 * 4);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    if (blocks < 1) blocks = 1;
    cast_f64_f32_kernel<<<blocks, 256, 0, st>>>(in, out, n);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_cast_f32_f64(cudaStream_t st, const float* in, double* out, long long n, double scale) {
    int blocks = cdiv(n, 256 * 4);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    if (blocks < 1) blocks = 1;
    cast_f32_f64_kernel<<<blocks, 256, 0, st>>>(in, out, n, scale);
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_sub_cast_f32_f64(cudaStream_t st, const float* a, const float* b, double* out, long long n) {
    int blocks = cdiv(n, 256 * 4);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    if (blocks < 1) blocks = 1;
    sub_cast_kernel<<<blocks, 256, 0, st>>>(a, b, out, n);
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_add_fwd(cudaStream_t st, int order, int relu, const View& a, const View& b, const View& y0, const View& yk, int batch) {
    const long long per = (long long)a.C * a.H * a.W;
    ProfScope prof("add_fwd", (double)per * batch, 12.0 * per * batch, st);
    add_fwd_kernel<<<ew_grid(per, batch), 256, 0, st>>>(order, relu, a.p, a.sstride, b.p, b.sstride, y0.p, y0.sstride, yk.p,
                                                      yk.sstride, per);
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_add_bwd(cudaStream_t st, const View* y0, const View& g, const View* ga, int acc_a, const View* gb, int acc_b, int batch) {
    const long long per = (long long)g.C * g.H * g.W;
    ProfScope prof("add_bwd", (double)per * batch, 16.0 * per * batch, st);
    add_bwd_kernel<<<ew_grid(per, batch), 256, 0, st>>>(y0 ? y0->p : nullptr, y0 ? y0->sstride : 0, g.p, g.sstride,
                                                      ga ? ga->p : nullptr, ga ? ga->sstride : 0, acc_a,
                                                      gb ? gb->p : nullptr, gb ? gb->sstride : 0, acc_b, per);
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s
