// vec.cu -- fused vector kernels of the spectral-radius iteration (comp_rho, opt.py:447-498).
//
// The reference spends ~15 ATen kernels and >= 5 host syncs per iteration on length-P fp64
// vectors (opt.py:455-498).  Here one iteration is TWO memory-bound passes over device-resident
// state, with every scalar (lambda, residual norms, stopping test, relaxation, double-buffer
// indices) kept on the device:
//   pass A  pi_dot     reads Hv (fp32) and v (fp64):            d = <Hv,v>, |Hv|^2, |v|^2
//                      last block: sign flip, lambda, |v + alpha (Hv - v)| from the three sums
//   pass B  pi_update  reads Hv, v, r_old; writes r = s Hv - lambda v, v_next (fp64) and its fp32
//                      rounding (the next HVP's input): |r|^2, |r - r_old|^2, |r + r_old|^2
//                      last block: stopping test of opt.py:479-481, bookkeeping of opt.py:483-498
// Algorithmic traffic per iteration: A 12 B/elem + B (4+8+8 read, 8+8+4 written) 40 B/elem = 52 P bytes
// (the reference's own sequence moves ~224 P).  Both kernels are gated on the
// device-side `done` flag, so a host that runs ahead never corrupts a converged state.
// 128-bit loads/stores, warp-shuffle + shared-memory block reduction, deterministic last-block
// final sum (no floating-point atomics).
#include "kernels.h"
#include "vec.h"

namespace b2s {

constexpr int kVecThreads = 256;
constexpr int kVecBlocksMax = kNumSMs * 8;

__device__ __forceinline__ bool last_block_arrives(unsigned* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicInc(counter, gridDim.x - 1);   // wraps to 0: self resetting
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    return is_last;
}

// deterministic final reduction of `nv` columns of scratch[gridDim.x][nv] by the last block
template <int NV>
__device__ __forceinline__ void final_sum(const double* scratch, double (&tot)[NV], double* red) {
    double acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] += __ldcg(scratch + (size_t)b * NV + q);
    }
    block_sum<NV, double>(acc, red);
#pragma unroll
    for (int q = 0; q < NV; ++q) tot[q] = acc[q];
}

__global__ void __launch_bounds__(kVecThreads) pi_dot_kernel(PiDev* __restrict__ S, const float* __restrict__ hv) {
    if (S->done) return;
    __shared__ double red[3 * 32];
    const long long n = S->n;
    const double* __restrict__ v = S->vbuf[S->cur];
    double d = 0, hh = 0, vv = 0;
    const long long n4 = n >> 2;
    const float4* hv4 = reinterpret_cast<const float4*>(hv);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 h = __ldg(hv4 + i);
        const double2 a = v2[2 * i], b = v2[2 * i + 1];
        d += (double)h.x * a.x + (double)h.y * a.y + (double)h.z * b.x + (double)h.w * b.y;
        hh += (double)h.x * h.x + (double)h.y * h.y + (double)h.z * h.z + (double)h.w * h.w;
        vv += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
    }
    if (blockIdx.x == 0) {
        const long long i = (n4 << 2) + threadIdx.x;
        if (i < n) {
            const double h = hv[i], a = v[i];
            d += h * a; hh += h * h; vv += a * a;
        }
    }
    double part[3] = {d, hh, vv};
    block_sum<3, double>(part, red);
    if (threadIdx.x == 0) {
        S->scratch[blockIdx.x * 3 + 0] = part[0];
        S->scratch[blockIdx.x * 3 + 1] = part[1];
        S->scratch[blockIdx.x * 3 + 2] = part[2];
    }
    if (!last_block_arrives(&S->counter)) return;
    double tot[3];
    final_sum<3>(S->scratch, tot, red);
    if (threadIdx.x == 0) {
        const double dd = tot[0], h2 = tot[1], v2n = tot[2];
        const double sign = dd < 0 ? -1.0 : 1.0;           // opt.py:458-460
        const double lam = dd * sign;
        const double alpha = S->alpha ? S->alpha[S->iter] : 1.0;
        // |v + alpha (s Hv - v)|^2 = (1-a)^2 |v|^2 + 2 a (1-a) lam + a^2 |Hv|^2     (opt.py:495-498)
        const double nn = (1 - alpha) * (1 - alpha) * v2n + 2 * alpha * (1 - alpha) * lam + alpha * alpha * h2;
        S->sign = sign;
        S->lam = lam;
        S->vnn = sqrt(h2);
        S->cur_alpha = alpha;
        S->inv_norm = nn > 0 ? 1.0 / sqrt(nn) : 0.0;
    }
}

__global__ void __launch_bounds__(kVecThreads) pi_update_kernel(PiDev* __restrict__ S, const float* __restrict__ hv) {
    if (S->done) return;
    __shared__ double red[3 * 32];
    const long long n = S->n;
    const int cur = S->cur, rcur = S->rcur;
    const double* __restrict__ v = S->vbuf[cur];
    double* __restrict__ vn = S->vbuf[1 - cur];
    const double* __restrict__ ro = S->rbuf[rcur];
    double* __restrict__ rn = S->rbuf[1 - rcur];
    float* __restrict__ v32 = S->v32;
    const bool has_old = S->has_old != 0;
    const bool plain = S->precond == 0;     // preconditioned variant updates v elsewhere
    const double s = S->sign, lam = S->lam, al = S->cur_alpha, inv = S->inv_norm;
    const double cv = (1.0 - al) * inv, ch = al * s * inv;
    double rr = 0, dm = 0, dp = 0;
    const long long n4 = n >> 2;
    const float4* hv4 = reinterpret_cast<const float4*>(hv);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    const double2* ro2 = reinterpret_cast<const double2*>(ro);
    double2* rn2 = reinterpret_cast<double2*>(rn);
    double2* vn2 = reinterpret_cast<double2*>(vn);
    float4* v324 = reinterpret_cast<float4*>(v32);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 h = __ldg(hv4 + i);
        const double2 a = v2[2 * i], b = v2[2 * i + 1];
        double2 o0 = make_double2(0, 0), o1 = make_double2(0, 0);
        if (has_old) { o0 = ro2[2 * i]; o1 = ro2[2 * i + 1]; }
        double2 r0, r1;
        r0.x = s * h.x - lam * a.x; r0.y = s * h.y - lam * a.y;
        r1.x = s * h.z - lam * b.x; r1.y = s * h.w - lam * b.y;
        rr += r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y;
        double t;
        t = r0.x - o0.x; dm += t * t; t = r0.y - o0.y; dm += t * t;
        t = r1.x - o1.x; dm += t * t; t = r1.y - o1.y; dm += t * t;
        t = r0.x + o0.x; dp += t * t; t = r0.y + o0.y; dp += t * t;
        t = r1.x + o1.x; dp += t * t; t = r1.y + o1.y; dp += t * t;
        rn2[2 * i] = r0; rn2[2 * i + 1] = r1;
        if (plain) {
            double2 w0, w1;
            w0.x = cv * a.x + ch * h.x; w0.y = cv * a.y + ch * h.y;
            w1.x = cv * b.x + ch * h.z; w1.y = cv * b.y + ch * h.w;
            vn2[2 * i] = w0; vn2[2 * i + 1] = w1;
            v324[i] = make_float4((float)w0.x, (float)w0.y, (float)w1.x, (float)w1.y);
        }
    }
    if (blockIdx.x == 0) {
        const long long i = (n4 << 2) + threadIdx.x;
        if (i < n) {
            const double h = hv[i], a = v[i], o = has_old ? ro[i] : 0.0;
            const double r = s * h - lam * a;
            rr += r * r; dm += (r - o) * (r - o); dp += (r + o) * (r + o);
            rn[i] = r;
            if (plain) {
                const double w = cv * a + ch * h;
                vn[i] = w;
                v32[i] = (float)w;
            }
        }
    }
    double part[3] = {rr, dm, dp};
    block_sum<3, double>(part, red);
    if (threadIdx.x == 0) {
        S->scratch[blockIdx.x * 3 + 0] = part[0];
        S->scratch[blockIdx.x * 3 + 1] = part[1];
        S->scratch[blockIdx.x * 3 + 2] = part[2];
    }
    if (!last_block_arrives(&S->counter)) return;
    double tot[3];
    final_sum<3>(S->scratch, tot, red);
    if (threadIdx.x == 0) {
        const int i = S->iter;
        const double nres = sqrt(tot[0]);
        const double rnv = sqrt(fmin(tot[1], tot[2]));                           // opt.py:463
        const double inf = INFINITY;
        const double s0 = nres;
        const double s1 = S->n_old != 0 ? rnv / S->n_old : inf;                    // opt.py:479
        const double s2 = S->lam_old != 0 ? fabs(lam - S->lam_old) / S->lam_old : inf;
        S->norm = nres; S->rn = rnv;
        S->stop[0] = s0; S->stop[1] = s1; S->stop[2] = s2;
        if (S->traj) {
            double* row = S->traj + (size_t)i * 4;
            row[0] = lam; row[1] = nres; row[2] = rnv; row[3] = S->vnn;
        }
        S->last_iter = i;
        const double eps = S->eps;
        if (s0 < eps || s1 < eps || s2 < eps) {                                     // opt.py:480-481
            S->converged = 1;
            S->done = 1;
        } else {
            if (i < S->max_iter - 1) {                                              // opt.py:483-485
                S->lam_old = lam; S->n_old = nres;
                S->rcur = 1 - rcur; S->has_old = 1;
            }
            S->r_last = 1 - rcur;
            if (plain) S->cur = 1 - cur;                                           // opt.py:498
            S->iter = i + 1;
            if (plain && i + 1 >= S->max_iter) S->done = 1;
        }
        S->pending = plain ? 0 : 1;   // preconditioned variant: host must finish the update
    }
}

// v_next = (v + alpha * Tr) / |.|  for the K-FAC preconditioned variant (opt.py:491-498): two tiny passes
__global__ void __launch_bounds__(kVecThreads) pi_precond_norm_kernel(PiDev* __restrict__ S, const double* __restrict__ Tr) {
    __shared__ double red[32];
    const long long n = S->n;
    const double* __restrict__ v = S->vbuf[S->cur];
    const double al = S->cur_alpha;
    double acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double w = v[i] + al * Tr[i];
        acc += w * w;
    }
    double part[1] = {acc};
    block_sum<1, double>(part, red);
    if (threadIdx.x == 0) S->scratch[blockIdx.x] = part[0];
    if (!last_block_arrives(&S->counter)) return;
    double tot[1];
    final_sum<1>(S->scratch, tot, red);
    if (threadIdx.x == 0) S->inv_norm = tot[0] > 0 ? 1.0 / sqrt(tot[0]) : 0.0;
}
__global__ void __launch_bounds__(kVecThreads) pi_precond_apply_kernel(PiDev* __restrict__ S, const double* __restrict__ Tr) {
    const long long n = S->n;
    const int cur = S->cur;
    const double* __restrict__ v = S->vbuf[cur];
    double* __restrict__ vn = S->vbuf[1 - cur];
    const double al = S->cur_alpha, inv = S->inv_norm;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double w = (v[i] + al * Tr[i]) * inv;
        vn[i] = w;
        S->v32[i] = (float)w;
    }
}
__global__ void pi_precond_commit_kernel(PiDev* S) {
    S->cur = 1 - S->cur;
    S->pending = 0;
    if (S->iter >= S->max_iter) S->done = 1;
}

static int vec_blocks(long long n) {
    long long b = (n / 4 + kVecThreads - 1) / kVecThreads;
    if (b > kVecBlocksMax) b = kVecBlocksMax;
    if (b < 1) b = 1;
    return (int)b;
}

int pi_scratch_doubles() { return kVecBlocksMax * 3; }

int launch_pi_step(cudaStream_t st, PiDev* dS, long long n, const float* hv) {
    const int blocks = vec_blocks(n);
    pi_dot_kernel<<<blocks, kVecThreads, 0, st>>>(dS, hv);
    B2S_LAUNCH_CHECK();
    pi_update_kernel<<<blocks, kVecThreads, 0, st>>>(dS, hv);
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_pi_precond_update(cudaStream_t st, PiDev* dS, long long n, const double* Tr) {
    const int blocks = vec_blocks(n);
    pi_precond_norm_kernel<<<blocks, kVecThreads, 0, st>>>(dS, Tr);
    B2S_LAUNCH_CHECK();
    pi_precond_apply_kernel<<<blocks, kVecThreads, 0, st>>>(dS, Tr);
    B2S_LAUNCH_CHECK();
    pi_precond_commit_kernel<<<1, 1, 0, st>>>(dS);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---- step assembly of the regularised minibatch step (opt.py:616-659) ---------------------------------
// p = grad f + coef * grad rho  (coef = mu * sign, or no grad rho at all when g == 0), written as fp64 (the
// reference's `p`) and as the fp32 flat vector whose slices become param.grad -- one pass over 8P (+8P) bytes in,
// 8P + 4P bytes out, instead of one slice + cast kernel per parameter tensor in a Python loop.
__global__ void __launch_bounds__(256) step_assemble_kernel(const double* __restrict__ gf, const double* __restrict__ gr, const double coef,
                                                            const long long n, double* __restrict__ p64, float* __restrict__ p32) {
    const long long n2 = n >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 a = reinterpret_cast<const double2*>(gf)[i];
        if (gr) {
            const double2 b = reinterpret_cast<const double2*>(gr)[i];
            a.x += coef * b.x;
            a.y += coef * b.y;
        }
        if (p64) reinterpret_cast<double2*>(p64)[i] = a;
        reinterpret_cast<float2*>(p32)[i] = make_float2((float)a.x, (float)a.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double a = gf[n - 1];
        if (gr) a += coef * gr[n - 1];
        if (p64) p64[n - 1] = a;
        p32[n - 1] = (float)a;
    }
}

int launch_step_assemble(cudaStream_t st, const double* gf, const double* gr, double coef, long long n, double* p64, float* p32) {
    // 16-byte vector accesses need 16-byte aligned bases (torch allocations are 256-byte aligned; slices may not be)
    if (((uintptr_t)gf | (uintptr_t)gr | (uintptr_t)p64) & 15 || ((uintptr_t)p32 & 7)) {
        set_error("b2s_step_assemble: vectors must be 16-byte aligned");
        return -1;
    }
    int blocks = cdiv(n, 256 * 4);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    if (blocks < 1) blocks = 1;
    step_assemble_kernel<<<blocks, 256, 0, st>>>(gf, gr, coef, n, p64, p32);
    B2S_LAUNCH_CHECK();
    return 0;
}


// ---- fused regularised step (opt.py:535-542, 616-659, 696-699) ---------------------------------------------
// The rest of iter()'s minibatch body on flat vectors: the clip of grad rho (norm from a first pass, scale kept on
// the device: no host sync), p = grad f + mu * sign * grad rho, its fp32 rounding (= param.grad), and the
// optimizer update (torch.optim.SGD with momentum / dampening / nesterov / weight decay, or torch.optim.Adam) applied
// in place to the flat fp32 parameter vector that the model's parameters are views of -- one kernel instead of the
// reference's per-parameter Python loop plus the optimizer's own per-tensor kernels.
__global__ void __launch_bounds__(kVecThreads) clip_norm_kernel(const double* __restrict__ x, const long long n, const double clip,
                                                                double* __restrict__ scratch, unsigned* __restrict__ counter,
                                                                double* __restrict__ out2) {
    __shared__ double red[32];
    double acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += x[i] * x[i];
    double part[1] = {acc};
    block_sum<1, double>(part, red);
    if (threadIdx.x == 0) scratch[blockIdx.x] = part[0];
    if (!last_block_arrives(counter)) return;
    double tot[1];
    final_sum<1>(scratch, tot, red);
    if (threadIdx.x == 0) {
        const double nrm = sqrt(tot[0]);
        out2[0] = nrm;
        out2[1] = (clip > 0 && nrm > clip) ? clip / nrm : 1.0;      // opt.py:539-542
    }
}

__global__ void __launch_bounds__(256) step_fused_kernel(const double* __restrict__ gf, double* __restrict__ gr, const double coef,
                                                         const double* __restrict__ scale2, const long long n, double* __restrict__ p64,
                                                         float* __restrict__ p32, float* __restrict__ w, float* __restrict__ s1,
                                                         float* __restrict__ s2, const StepOpt o) {
    const double sc = (gr && scale2) ? scale2[1] : 1.0;
    const double cs = coef * sc;
    const bool rescale = gr && sc != 1.0 && o.write_gradrho;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double a = gf[i];
        if (gr) {
            const double b = gr[i];
            a += cs * b;
            if (rescale) gr[i] = sc * b;
        }
        if (p64) p64[i] = a;
        float g = (float)a;                      // param.grad = p[i:i+n].view(s).float()   opt.py:658
        p32[i] = g;
        if (o.kind == 0) continue;
        float x = w[i];
        if (o.maximize) g = -g;
        if (o.weight_decay != 0.f) g = g + o.weight_decay * x;
        if (o.kind == 1) {                       // torch.optim.SGD (_single_tensor_sgd)
            if (o.momentum != 0.f) {
                float b = o.first ? g : o.momentum * s1[i] + (1.f - o.dampening) * g;
                s1[i] = b;
                g = o.nesterov ? g + o.momentum * b : b;
            }
            w[i] = x - o.lr * g;
        } else {                                 // torch.optim.Adam (_single_tensor_adam, amsgrad = False)
            float m = s1[i], v = s2[i];
            m = m + (g - m) * (1.f - o.beta1);
            v = o.beta2 * v + (1.f - o.beta2) * g * g;
            s1[i] = m; s2[i] = v;
            const float denom = sqrtf(v) / o.bias2_sqrt + o.eps;
            w[i] = x - o.step_size * (m / denom);
        }
    }
}

int launch_clip_norm(cudaStream_t st, const double* x, long long n, double clip, double* scratch, double* out2) {
    const int blocks = vec_blocks(n);
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + kVecBlocksMax);
    clip_norm_kernel<<<blocks, kVecThreads, 0, st>>>(x, n, clip, scratch, counter, out2);
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_step_fused(cudaStream_t st, const double* gf, double* gr, double coef, const double* scale2, long long n, double* p64,
                      float* p32, float* w, float* s1, float* s2, const StepOpt& o) {
    int blocks = cdiv(n, 256 * 2);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    if (blocks < 1) blocks = 1;
    step_fused_kernel<<<blocks, 256, 0, st>>>(gf, gr, coef, scale2, n, p64, p32, w, s1, s2, o);
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s
