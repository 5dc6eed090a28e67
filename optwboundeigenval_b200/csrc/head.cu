// head.cu -- loss heads: adjoint jets of the logits.
//
// For a head with closed-form gradient grad_z l(z), evaluating that closed form on the jet
// (z, zdot, zddot) yields (zbar, R zbar, R^2 zbar) = (grad l, hess l . zdot,
// d3 l[zdot,zdot] + hess l . zddot) in one go (rop.py:129,147 do this by hand for MSE).
// Heads of the BASELINE configs:
//   CE            densenet.py:121 + nn.CrossEntropyLoss
//   SOFTMAX_CE    forest_data.py:88 / usps_data.py:335: the model already ends in a softmax and
//                 CrossEntropyLoss applies log-softmax AGAIN
//   WBCE          dcnn.py:375-400 on raw logits (MyVggNet16_bn)
//   SIGMOID_WBCE  dcnn.py:275: Linear -> Sigmoid, then BCE-with-logits applies a sigmoid again
// One thread per sample, arithmetic in fp64 (inputs and outputs fp32).  Up to 64 classes the class
// dimension lives in thread-local arrays; wider heads (CIFAR-100: cifar100_ResNet_mu0.py) run one
// single-thread block per sample with the jets in shared memory (B x C is tiny either way).  A label outside [0, C) poisons the sample's loss and
// adjoint with NaN instead of reading out of bounds.
#include "kernels.h"
#include "../../include/b200_spectral.h"

namespace b2s {

constexpr int kLocalClasses = 64;
constexpr int kMaxClasses = 640;       // 3 jets x 3 doubles x 640 classes = 46 KB of shared memory

template <int K>
using JD = Jet<K, double>;

template <int K>
__device__ inline void softmax_jet(const JD<K>* u, JD<K>* p, int C, double* logsumexp0) {
    double m = u[0].c[0];
    for (int i = 1; i < C; ++i) m = fmax(m, u[i].c[0]);
    JD<K> S;
    for (int i = 0; i < C; ++i) {
        JD<K> t = u[i];
        t.c[0] -= m;
        p[i] = jet_exp(t);
        S = S + p[i];
    }
    const JD<K> inv = jet_recip(S);
    for (int i = 0; i < C; ++i) p[i] = p[i] * inv;
    if (logsumexp0) *logsumexp0 = m + log(S.c[0]);
}

template <int K>
__device__ inline JD<K> sigmoid_jet(const JD<K>& u) {
    JD<K> t = scale(u, -1.0);
    JD<K> e = jet_exp(t);
    e.c[0] += 1.0;
    return jet_recip(e);
}

__device__ inline double bce_with_logits(double u, double t) {
    return fmax(u, 0.0) - u * t + log1p(exp(-fabs(u)));
}

template <int K>
__device__ inline void head_sample(const HeadArgs& a, const int n, JD<K>* z, JD<K>* p, JD<K>* q) {
    const int C = a.C;
    if (a.kind == B2S_HEAD_CE || a.kind == B2S_HEAD_SOFTMAX_CE) {
        const long long y = a.labels[n];
        if (y < 0 || y >= C) {                       // invalid label: visible, not out of bounds
            float* bad = a.zbar + (long long)n * a.zs;
            for (int i = 0; i < C; ++i) bad[i] = __int_as_float(0x7fc00000);
            if (K == 0 && a.loss) atomicAdd(a.loss, (double)__int_as_float(0x7fc00000));
            return;
        }
    }
    for (int i = 0; i < C; ++i) {
        z[i] = JD<K>();
        const long long idx = (long long)n * a.zs + i;
        z[i].c[0] = a.z[0][idx];
        if (K >= 1) z[i].c[1] = a.z[1][idx];
        if (K >= 2) z[i].c[2] = a.z[2][idx];
    }
    float* out = a.zbar + (long long)n * a.zs;
    double loss = 0.0;
    if (a.kind == B2S_HEAD_CE) {
        double lse;
        softmax_jet<K>(z, p, C, &lse);
        const int y = (int)a.labels[n];
        for (int i = 0; i < C; ++i) {
            double v = p[i].c[K];
            if (K == 0 && i == y) v -= 1.0;
            out[i] = (float)(v * a.loss_scale);
        }
        loss = -(z[y].c[0] - lse) * a.loss_scale;
    } else if (a.kind == B2S_HEAD_SOFTMAX_CE) {
        softmax_jet<K>(z, p, C, nullptr);
        double lse;
        softmax_jet<K>(p, q, C, &lse);
        const int y = (int)a.labels[n];
        JD<K> dotp;
        for (int i = 0; i < C; ++i) {
            if (i == y) q[i].c[0] -= 1.0;          // q now holds pbar / loss_scale
            dotp = dotp + q[i] * p[i];
        }
        for (int i = 0; i < C; ++i) {
            const JD<K> zb = p[i] * (q[i] - dotp);
            out[i] = (float)(zb.c[K] * a.loss_scale);
        }
        loss = -(p[y].c[0] - lse) * a.loss_scale;
    } else if (a.kind == B2S_HEAD_WBCE) {
        for (int i = 0; i < C; ++i) {
            const long long ti = (long long)n * C + i;
            const double cf = a.coef[ti], t = a.target[ti];
            JD<K> s = sigmoid_jet<K>(z[i]);
            if (K == 0) s.c[0] -= t;
            out[i] = (float)(cf * s.c[K]);
            if (cf != 0.0) loss += cf * bce_with_logits(z[i].c[0], t);
        }
    } else {   // B2S_HEAD_SIGMOID_WBCE
        for (int i = 0; i < C; ++i) {
            const long long ti = (long long)n * C + i;
            const double cf = a.coef[ti], t = a.target[ti];
            const JD<K> s = sigmoid_jet<K>(z[i]);
            JD<K> q = sigmoid_jet<K>(s);
            q.c[0] -= t;                             // sbar / coef
            JD<K> one_minus = scale(s, -1.0);
            one_minus.c[0] += 1.0;
            const JD<K> zb = q * s * one_minus;
            out[i] = (float)(cf * zb.c[K]);
            if (cf != 0.0) loss += cf * bce_with_logits(s.c[0], t);
        }
    }
    if (K == 0 && a.loss) atomicAdd(a.loss, loss);
}

template <int K>
__global__ void __launch_bounds__(128) head_kernel(const HeadArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.batch) return;
    JD<K> z[kLocalClasses], p[kLocalClasses], q[kLocalClasses];
    head_sample<K>(a, n, z, p, q);
}

template <int K>
__global__ void __launch_bounds__(32) head_wide_kernel(const HeadArgs a) {
    extern __shared__ double head_smem[];
    JD<K>* z = reinterpret_cast<JD<K>*>(head_smem);
    if (threadIdx.x == 0) head_sample<K>(a, blockIdx.x, z, z + a.C, z + 2 * a.C);
}

int launch_head(cudaStream_t st, int order, const HeadArgs& a) {
    if (a.C > kMaxClasses) {
        set_error("loss head supports at most %d classes, got %d", kMaxClasses, a.C);
        return -4;
    }
    ProfScope prof("loss_head", 0.0, 4.0 * a.batch * a.C * (order + 2), st);
    const int blocks = cdiv(a.batch, 128);
    if (a.C <= kLocalClasses) {
        if (order == 0) head_kernel<0><<<blocks, 128, 0, st>>>(a);
        else if (order == 1) head_kernel<1><<<blocks, 128, 0, st>>>(a);
        else head_kernel<2><<<blocks, 128, 0, st>>>(a);
    } else {
        const size_t sm = (size_t)3 * a.C * sizeof(JD<0>);
        if (order == 0) head_wide_kernel<0><<<a.batch, 32, sm, st>>>(a);
        else if (order == 1) head_wide_kernel<1><<<a.batch, 32, sm, st>>>(a);
        else head_wide_kernel<2><<<a.batch, 32, sm, st>>>(a);
    }
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s
