// kfac.cu -- K-FAC preconditioner of the `lobpcg=True` variant (opt.py:362-416, kfac.py:50-130,277-367).
//
// The reference's "LOBPCG" step is v <- normalise(v + alpha * T r) with T the layer-wise Kronecker
// map  R -> G^-1 R A^-1  (opt.py:491-493, kfac.py:118-120), identity on parameters that are not
// Conv2d/Linear.  The factors come from one forward/backward over the batch:
//     A = 0.95 I + 0.05 * a^T a / B      a = im2col patches / spatial, bias column of ones   kfac.py:292-311
//     G = 0.95 I + 0.05 * B*S * g^T g    g = adjoint of the layer output                      kfac.py:337-367
// Both are Gram matrices of a gathered [pixels x K] matrix; they are built straight from the plan's
// cached base pass (activations and output adjoints are already in HBM).  eigh stays on the host
// side (kfac.py:87-93); the two inverses it yields are applied here with a small tiled SGEMM.
#include "kernels.h"

namespace b2s {

// ---- Gram matrix of gathered patches ---------------------------------------------------------------
// out[k1][k2] = diag_add*(k1==k2) + scale * sum_{n,oy,ox} P[j][k1] * P[j][k2]
// P[j][k] = src[n, c, oy*sh+ky-ph, ox*sw+kx-pw] * elem_scale for k = (c,ky,kx) < C*KH*KW, and
// P[j][K-1] = ones_value when has_ones (the bias column).
struct GramArgs {
    const float* src;
    long long sstride;
    int batch, C, H, W, OH, OW, KH, KW, sh, sw, ph, pw;
    int K;            // matrix size including the ones column
    int has_ones;
    float elem_scale, ones_value, scale, diag_add;
    float* out;
    int j_chunk;
};

__device__ __forceinline__ float gram_elem(const GramArgs& a, int n, int oy, int ox, int k) {
    const int KHW = a.KH * a.KW;
    const int Kp = a.C * KHW;
    if (k >= Kp) return a.ones_value;
    const int c = k / KHW, t = k - c * KHW;
    const int ky = t / a.KW, kx = t - ky * a.KW;
    const int sy = oy * a.sh + ky - a.ph, sx = ox * a.sw + kx - a.pw;
    if (sy < 0 || sy >= a.H || sx < 0 || sx >= a.W) return 0.f;
    return a.src[(long long)n * a.sstride + ((long long)c * a.H + sy) * a.W + sx] * a.elem_scale;
}

constexpr int GT = 32;   // output tile
constexpr int GJ = 32;   // pixels per stage

__global__ void __launch_bounds__(256) gram_kernel(const GramArgs a) {
    __shared__ float P1[GJ][GT + 1];
    __shared__ float P2[GJ][GT + 1];
    const int k1_0 = blockIdx.y * GT, k2_0 = blockIdx.x * GT;
    if (k2_0 > k1_0 + GT - 1) return;                // symmetric: lower block triangle only
    const int OHW = a.OH * a.OW;
    const long long J = (long long)a.batch * OHW;
    const long long jbeg = (long long)blockIdx.z * a.j_chunk;
    const long long jend = jbeg + a.j_chunk < J ? jbeg + a.j_chunk : J;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;          // each thread: 2x2 outputs
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (long long j0 = jbeg; j0 < jend; j0 += GJ) {
        for (int e = tid; e < GJ * GT; e += 256) {
            const int jj = e / GT, kk = e - jj * GT;
            const long long j = j0 + jj;
            float v1 = 0.f, v2 = 0.f;
            if (j < jend) {
                const int n = (int)(j / OHW);
                const int pix = (int)(j - (long long)n * OHW);
                const int oy = pix / a.OW, ox = pix - oy * a.OW;
                if (k1_0 + kk < a.K) v1 = gram_elem(a, n, oy, ox, k1_0 + kk);
                if (k2_0 + kk < a.K) v2 = gram_elem(a, n, oy, ox, k2_0 + kk);
            }
            P1[jj][kk] = v1;
            P2[jj][kk] = v2;
        }
        __syncthreads();
#pragma unroll 8
        for (int jj = 0; jj < GJ; ++jj) {
            const float a0 = P1[jj][ty * 2], a1 = P1[jj][ty * 2 + 1];
            const float b0 = P2[jj][tx * 2], b1 = P2[jj][tx * 2 + 1];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k1 = k1_0 + ty * 2 + i, k2 = k2_0 + tx * 2 + j;
            if (k1 < a.K && k2 < a.K) {
                const float v = acc[i][j] * a.scale;
                atomicAdd(a.out + (long long)k1 * a.K + k2, v);
                if (k2_0 != k1_0) atomicAdd(a.out + (long long)k2 * a.K + k1, v);   // mirror off-diagonal blocks
            }
        }
}

__global__ void gram_init_kernel(float* out, int K, float diag) {
    const long long n = (long long)K * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (i / K == i % K) ? diag : 0.f;
}

int launch_gram(cudaStream_t st, const float* src, long long sstride, int batch, int C, int H, int W, int OH, int OW,
                int KH, int KW, int sh, int sw, int ph, int pw, int has_ones, float elem_scale, float ones_value,
                float scale, float diag_add, float* out) {
    GramArgs a{};
    a.src = src; a.sstride = sstride; a.batch = batch; a.C = C; a.H = H; a.W = W; a.OH = OH; a.OW = OW;
    a.KH = KH; a.KW = KW; a.sh = sh; a.sw = sw; a.ph = ph; a.pw = pw;
    a.K = C * KH * KW + (has_ones ? 1 : 0);
    a.has_ones = has_ones; a.elem_scale = elem_scale; a.ones_value = ones_value; a.scale = scale; a.diag_add = diag_add;
    a.out = out;
    const int tiles = cdiv(a.K, GT);
    const long long J = (long long)batch * OH * OW;
    long long splits = (4LL * kNumSMs) / ((long long)tiles * (tiles + 1) / 2) + 1;
    const long long maxs = (J + GJ - 1) / GJ;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    long long per = (maxs + splits - 1) / splits;
    a.j_chunk = (int)(per * GJ);
    splits = (maxs + per - 1) / per;
    {
        ProfScope prof("kfac_gram_init", 0.0, 4.0 * a.K * a.K, st);
        gram_init_kernel<<<cdiv((long long)a.K * a.K, 256), 256, 0, st>>>(out, a.K, diag_add);
        B2S_LAUNCH_CHECK();
    }
    ProfScope prof("kfac_gram", 2.0 * (double)J * a.K * a.K, 4.0 * (double)J * a.K, st);
    gram_kernel<<<dim3(tiles, tiles, (unsigned)splits), 256, 0, st>>>(a);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---- small SGEMM: C[M,N] = A[M,K] * B[K,N], row-major, optional transposes by strides ----------------
__global__ void __launch_bounds__(256) sgemm_small_kernel(int M, int N, int K, const float* __restrict__ A, int a_rs,
                                                          int a_cs, const float* __restrict__ B, int b_rs, int b_cs,
                                                          float* __restrict__ C, int ldc) {
    __shared__ float As[16][16 + 1];
    __shared__ float Bs[16][16 + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        As[ty][tx] = (m < M && k0 + tx < K) ? A[(long long)m * a_rs + (long long)(k0 + tx) * a_cs] : 0.f;
        Bs[ty][tx] = (k0 + ty < K && n < N) ? B[(long long)(k0 + ty) * b_rs + (long long)n * b_cs] : 0.f;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc = fmaf(As[ty][kk], Bs[kk][tx], acc);
        __syncthreads();
    }
    if (m < M && n < N) C[(long long)m * ldc + n] = acc;
}

int launch_sgemm_small(cudaStream_t st, int M, int N, int K, const float* A, int a_rs, int a_cs, const float* B,
                       int b_rs, int b_cs, float* C, int ldc) {
    ProfScope prof("kfac_sgemm", 2.0 * M * N * (double)K, 4.0 * ((double)M * K + (double)K * N + (double)M * N), st);
    sgemm_small_kernel<<<dim3(cdiv(N, 16), cdiv(M, 16)), 256, 0, st>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---- gather r -> [dg x da] fp32 matrix (weights, bias as last column) and scatter back ---------------
__global__ void kfac_gather_kernel(const double* __restrict__ r, long long w_off, long long b_off, int dg, int dw,
                                   float* __restrict__ M) {
    const int da = dw + (b_off >= 0 ? 1 : 0);
    const long long n = (long long)dg * da;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / da), col = (int)(i - (long long)row * da);
        M[i] = col < dw ? (float)r[w_off + (long long)row * dw + col] : (float)r[b_off + row];
    }
}
__global__ void kfac_scatter_kernel(const float* __restrict__ M, long long w_off, long long b_off, int dg, int dw,
                                    double* __restrict__ out) {
    const int da = dw + (b_off >= 0 ? 1 : 0);
    const long long n = (long long)dg * da;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / da), col = (int)(i - (long long)row * da);
        if (col < dw) out[w_off + (long long)row * dw + col] = (double)M[i];
        else out[b_off + row] = (double)M[i];
    }
}

int launch_kfac_gather(cudaStream_t st, const double* r, long long w_off, long long b_off, int dg, int dw, float* M) {
    const long long n = (long long)dg * (dw + (b_off >= 0 ? 1 : 0));
    kfac_gather_kernel<<<cdiv(n, 256) > 1184 ? 1184 : cdiv(n, 256), 256, 0, st>>>(r, w_off, b_off, dg, dw, M);
    B2S_LAUNCH_CHECK();
    return 0;
}
int launch_kfac_scatter(cudaStream_t st, const float* M, long long w_off, long long b_off, int dg, int dw, double* out) {
    const long long n = (long long)dg * (dw + (b_off >= 0 ? 1 : 0));
    kfac_scatter_kernel<<<cdiv(n, 256) > 1184 ? 1184 : cdiv(n, 256), 256, 0, st>>>(M, w_off, b_off, dg, dw, out);
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s
