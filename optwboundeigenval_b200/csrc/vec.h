// vec.h -- device-resident state of the spectral-radius iteration (internal).
#pragma once
#include "common.cuh"

namespace b2s {

struct PiDev {
    long long n;
    double* vbuf[2];        // current / next eigenvector estimate (fp64, as self.v in opt.py:508)
    double* rbuf[2];        // residual double buffer (r_old of opt.py:463,485)
    float* v32;             // fp32 rounding of the current vector = input of the next HVP
    const double* alpha;    // relaxation per iteration (device array) or NULL = 1
    double* traj;           // [max_iter][4] rows (lam, n, rn, vnn) or NULL
    double* scratch;        // per-block partial sums
    unsigned counter;       // last-block detection (self resetting)
    int cur, rcur, has_old, r_last;
    int precond;            // 1: K-FAC preconditioned update, finished by launch_pi_precond_update
    int pending;
    int iter, max_iter, last_iter;
    int done, converged;
    double eps;
    double sign, lam, vnn, cur_alpha, inv_norm;
    double norm, rn, n_old, lam_old;
    double stop[3];
};

int pi_scratch_doubles();
// one iteration's vector work: pass A (dots) + pass B (residual, stopping test, update)
int launch_pi_step(cudaStream_t st, PiDev* dS, long long n, const float* hv);
int launch_pi_precond_update(cudaStream_t st, PiDev* dS, long long n, const double* Tr);
// p64 (optional) = gf + coef * gr (gr optional), p32 = (float)p64 : the step assembly of opt.py:616-659
int launch_step_assemble(cudaStream_t st, const double* gf, const double* gr, double coef, long long n, double* p64, float* p32);


struct StepOpt {
    int kind;               // 0: assemble only, 1: SGD, 2: Adam
    int first;              // SGD: momentum buffers are initialised by this step (torch: buf = clone(grad))
    int nesterov, maximize, write_gradrho;
    float lr, momentum, dampening, weight_decay;
    float beta1, beta2, eps, step_size, bias2_sqrt;     // Adam: step_size = lr / (1 - beta1^t), bias2_sqrt = sqrt(1 - beta2^t)
};
// out2 = {|x|, clip > 0 && |x| > clip ? clip / |x| : 1}; scratch: pi_scratch_doubles() + 1 doubles, zero-initialised once
int launch_clip_norm(cudaStream_t st, const double* x, long long n, double clip, double* scratch, double* out2);
int launch_step_fused(cudaStream_t st, const double* gf, double* gr, double coef, const double* scale2, long long n, double* p64,
                      float* p32, float* w, float* s1, float* s2, const StepOpt& o);

}  // namespace b2s
