// common.cuh -- shared device helpers of libb200spectral (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b2s {

// ---- error plumbing -----------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local long long g_launches;    // per-thread tally, folded into the global counter
void count_launch(int n = 1);

#define B2S_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            b2s::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                           cudaGetErrorString(e_));                                        \
            return -2;                                                                     \
        }                                                                                  \
    } while (0)

#define B2S_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess) {                                                           \
            b2s::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,        \
                           cudaGetErrorString(e_));                                        \
            return -3;                                                                     \
        }                                                                                  \
        b2s::count_launch();                                                               \
    } while (0)

#define B2S_TRY(expr)                                                                      \
    do {                                                                                   \
        int rc_ = (expr);                                                                  \
        if (rc_ != 0) return rc_;                                                          \
    } while (0)

// ---- built-in per-launch timer (used by b2s_profile_pass for the roofline numbers) -----------
// When a profiler is active on this thread, every launcher brackets its kernel with CUDA events on
// the launching stream and tags it with its algorithmic FLOPs and bytes.
struct ProfScope {
    int idx;
    cudaStream_t st;
    ProfScope(const char* name, double flops, double bytes, cudaStream_t stream);
    ~ProfScope();
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The passes are chains of ~200 short dependent kernels; with PDL the next kernel's CTAs are scheduled
// as soon as every CTA of the previous kernel has called pdl_trigger() (first statement of the kernels
// that take part), run their prologue (barrier init, TMEM allocation) and block in pdl_wait() until the
// previous grid has completed and its writes are visible.  Both instructions are no-ops for a kernel
// launched the ordinary way.  Measured on DenseNet3 (graph replay): 2.78 vs 2.76 ms, i.e. nothing -- the
// attribute is only set with B2S_PDL=1.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// timing experiments only (B2S_SKIP=conv_fwd,bn_bwd,...): the named kernel families are not launched, so
// that the difference of two graph-replayed pass times is the in-situ cost of a family.  Results are garbage.
bool skip_family(const char* name);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---- jets: value + first + second directional derivative ------------------------------
// Derivative convention (not Taylor-normalised): c[1] = d/dt, c[2] = d^2/dt^2 at t = 0 of a
// quantity evaluated at w + t v.  Product rule: (ab)'' = a''b + 2a'b' + ab''.
// Math spec: rop.py:69-164 (R-op / R^2-op recurrences), generalised.
template <int K, typename T = float>
struct Jet {
    T c[3];   // components above K are never read (dead code after inlining)
    __host__ __device__ Jet() { c[0] = c[1] = c[2] = T(0); }
    __host__ __device__ explicit Jet(T a0) { c[0] = a0; c[1] = c[2] = T(0); }
    __host__ __device__ Jet(T a0, T a1, T a2) { c[0] = a0; c[1] = K >= 1 ? a1 : T(0); c[2] = K >= 2 ? a2 : T(0); }
    __host__ __device__ T top() const { return c[K]; }
};

template <int K, typename T>
__host__ __device__ inline Jet<K, T> operator+(const Jet<K, T>& a, const Jet<K, T>& b) {
    Jet<K, T> r;
    r.c[0] = a.c[0] + b.c[0];
    if (K >= 1) r.c[1] = a.c[1] + b.c[1];
    if (K >= 2) r.c[2] = a.c[2] + b.c[2];
    return r;
}
template <int K, typename T>
__host__ __device__ inline Jet<K, T> operator-(const Jet<K, T>& a, const Jet<K, T>& b) {
    Jet<K, T> r;
    r.c[0] = a.c[0] - b.c[0];
    if (K >= 1) r.c[1] = a.c[1] - b.c[1];
    if (K >= 2) r.c[2] = a.c[2] - b.c[2];
    return r;
}
template <int K, typename T>
__host__ __device__ inline Jet<K, T> operator*(const Jet<K, T>& a, const Jet<K, T>& b) {
    Jet<K, T> r;
    r.c[0] = a.c[0] * b.c[0];
    if (K >= 1) r.c[1] = a.c[1] * b.c[0] + a.c[0] * b.c[1];
    if (K >= 2) r.c[2] = a.c[2] * b.c[0] + T(2) * a.c[1] * b.c[1] + a.c[0] * b.c[2];
    return r;
}
template <int K, typename T>
__host__ __device__ inline Jet<K, T> scale(const Jet<K, T>& a, T s) {
    Jet<K, T> r;
    r.c[0] = a.c[0] * s;
    if (K >= 1) r.c[1] = a.c[1] * s;
    if (K >= 2) r.c[2] = a.c[2] * s;
    return r;
}
// s^(-1/2)
template <int K, typename T>
__host__ __device__ inline Jet<K, T> jet_rsqrt(const Jet<K, T>& s) {
    Jet<K, T> r;
    T r0 = T(1) / sqrt(s.c[0]);
    T r3 = r0 * r0 * r0;
    r.c[0] = r0;
    if (K >= 1) r.c[1] = T(-0.5) * r3 * s.c[1];
    if (K >= 2) r.c[2] = T(0.75) * r3 * r0 * r0 * s.c[1] * s.c[1] - T(0.5) * r3 * s.c[2];
    return r;
}
template <int K, typename T>
__host__ __device__ inline Jet<K, T> jet_exp(const Jet<K, T>& a) {
    Jet<K, T> r;
    T e = exp(a.c[0]);
    r.c[0] = e;
    if (K >= 1) r.c[1] = e * a.c[1];
    if (K >= 2) r.c[2] = e * (a.c[2] + a.c[1] * a.c[1]);
    return r;
}
template <int K, typename T>
__host__ __device__ inline Jet<K, T> jet_recip(const Jet<K, T>& a) {
    Jet<K, T> r;
    T q = T(1) / a.c[0];
    r.c[0] = q;
    if (K >= 1) r.c[1] = -a.c[1] * q * q;
    if (K >= 2) r.c[2] = (T(2) * a.c[1] * a.c[1] * q - a.c[2]) * q * q;
    return r;
}

// ---- reductions -----------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0. blockDim.x multiple of 32, <= 1024.
template <int NV, typename T>
__device__ __forceinline__ void block_sum(T (&v)[NV], T* smem /* [NV*32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            T x = lane < nw ? smem[i * 32 + lane] : T(0);
            v[i] = warp_sum(x);
        }
    }
}

}  // namespace b2s
