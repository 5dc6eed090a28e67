// conv.cu -- implicit-GEMM Conv2d / Linear jet contractions (fp32 SIMT path).
//
// One templated tile kernel serves the three contractions a Conv2d/Linear layer needs at
// every jet order (math spec rop.py:83-98,115-164):
//   forward   ydot  = conv(xdot, W) + conv(x, V)                (K-concatenated "pairs")
//   dgrad     Rxbar = conv^T(Rybar, W) + conv^T(ybar, V)
//   wgrad     RWbar = corr(x, Rybar) + corr(xdot, ybar)         -> Hv slice of this layer
// Instead of three separate library calls per term (what autograd issues, opt.py:99,132,143),
// up to three (activation, weight) pairs are contracted in ONE launch into one accumulator,
// so every output element is written once.  Layout is NCHW with an explicit per-sample
// stride so channel-concatenated DenseNet features are addressed in place.
//
// This is the fp32-exact CUDA-core path (parity tolerance rtol 1e-4 leaves no room for
// single-pass TF32).  GEMM view:
//   forward / dgrad:  M = destination channels, N = batch*pixels, K = source channels*KH*KW
//   wgrad:            M = Cout, N = Cin*KH*KW, K = batch*pixels (split across blockIdx.z)
#include <algorithm>
#include "conv_args.h"
#include "kernels.h"

namespace b2s {

// ConvKArgs / MODE_* live in conv_args.h (shared with the tcgen05 path, conv_tc.cu)

// ---------------------------------------------------------------------------------------------
// forward / dgrad
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN, int MODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_gather_gemm_kernel(const ConvKArgs a) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TXN = BN / TN;
    static_assert(NT % BN == 0, "thread count must be a multiple of BN");
    static_assert(NT % BK == 0, "thread count must be a multiple of BK");
    constexpr int B_ROWS = NT / BN;          // k rows of the B tile loaded per pass
    constexpr int A_ROWS = NT / BK;          // m rows of the A tile loaded per pass
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];

    const ConvGeom& g = a.g;
    // destination / source roles
    const int Cd = MODE == MODE_FWD ? g.Cout : g.Cin;
    const int Hd = MODE == MODE_FWD ? g.OH : g.H;
    const int Wd = MODE == MODE_FWD ? g.OW : g.W;
    const long long d_ss = MODE == MODE_FWD ? g.out_sstride : g.in_sstride;
    const int Cs = MODE == MODE_FWD ? g.Cin : g.Cout;
    const int Hs = MODE == MODE_FWD ? g.H : g.OH;
    const int Ws = MODE == MODE_FWD ? g.W : g.OW;
    const long long s_ss = MODE == MODE_FWD ? g.in_sstride : g.out_sstride;
    const int KHW = g.KH * g.KW;
    const int Ktot = Cs * KHW;
    const int HWd = Hd * Wd;
    const long long J = (long long)g.batch * HWd;

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.y * BM;
    const long long j0 = (long long)blockIdx.x * BN;

    // this thread's pixel for the B-tile loads
    const int jb = tid % BN;
    const int kb0 = tid / BN;
    const long long jl = j0 + jb;
    const bool j_ok = jl < J;
    int n_l = 0, y_l = 0, x_l = 0;
    if (j_ok) {
        n_l = (int)(jl / HWd);
        int pix = (int)(jl - (long long)n_l * HWd);
        y_l = pix / Wd;
        x_l = pix - y_l * Wd;
    }
    const int ka = tid % BK;
    const int ma0 = tid / BK;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int p = 0; p < a.npairs; ++p) {
        const float* __restrict__ src = a.act[p] + (long long)n_l * s_ss;
        const float* __restrict__ wt = a.wt[p];
        const float sc = a.scale[p];
        for (int k0 = 0; k0 < Ktot; k0 += BK) {
            // ---- A tile: weights
#pragma unroll
            for (int i = 0; i < BM / A_ROWS; ++i) {
                const int mb = ma0 + i * A_ROWS;
                const int m = m0 + mb, k = k0 + ka;
                float v = 0.f;
                if (m < Cd && k < Ktot) {
                    if (MODE == MODE_FWD) {
                        v = wt[(long long)m * Ktot + k];
                    } else {
                        const int c = k / KHW, t = k - c * KHW;       // c = output channel of the conv
                        v = wt[((long long)c * g.Cin + m) * KHW + t];
                    }
                }
                As[ka][mb] = v * sc;
            }
            // ---- B tile: gathered activations
#pragma unroll
            for (int i = 0; i < BK / B_ROWS; ++i) {
                const int kb = kb0 + i * B_ROWS;
                const int k = k0 + kb;
                float v = 0.f;
                if (j_ok && k < Ktot) {
                    int c, ky, kx;
                    if (KHW == 1) { c = k; ky = 0; kx = 0; }
                    else { c = k / KHW; const int t = k - c * KHW; ky = t / g.KW; kx = t - ky * g.KW; }
                    int sy, sx;
                    bool ok;
                    if (MODE == MODE_FWD) {
                        sy = y_l * g.sh + ky - g.ph;
                        sx = x_l * g.sw + kx - g.pw;
                        ok = sy >= 0 && sy < Hs && sx >= 0 && sx < Ws;
                    } else {
                        const int ty_ = y_l + g.ph - ky, tx_ = x_l + g.pw - kx;
                        ok = ty_ >= 0 && tx_ >= 0;
                        sy = ty_ / g.sh; sx = tx_ / g.sw;
                        ok = ok && (sy * g.sh == ty_) && (sx * g.sw == tx_) && sy < Hs && sx < Ws;
                    }
                    if (ok) v = src[((long long)c * Hs + sy) * Ws + sx];
                }
                Bs[kb][jb] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float av[TM], bv[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx + j * TXN];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const long long jj = j0 + tx + j * TXN;
        if (jj >= J) continue;
        const int n = (int)(jj / HWd);
        const int pix = (int)(jj - (long long)n * HWd);
        const long long base = (long long)n * d_ss + pix;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + ty * TM + i;
            if (m >= Cd) continue;
            const long long o = base + (long long)m * HWd;
            float v = acc[i][j];
            if (a.bias) v += a.bias[m];
            if (a.accumulate) v += a.out[o];
            if (a.relu_mode == 1) v = v > 0.f ? v : 0.f;
            else if (a.relu_mode == 2) v = a.relu_ref[o] > 0.f ? v : 0.f;
            a.out[o] = v;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// forward / dgrad for FEW destination channels (DenseNet growth convs: 48->12 3x3, Cin->48 1x1 and
// their transposes).  One thread per destination pixel keeps MT destination channels in registers;
// warps of a CTA split the destination channels (groups) and the source channels (K slices), the
// slices are summed through shared memory.  Weights of one (activation, weight) pair are staged in
// shared memory as [k][m] so that a thread reads its MT weights with 128-bit broadcast loads.
// Replaces the tile kernel where its grid collapses (8..128 CTAs for the 8x8 / 16x16 maps).
// ---------------------------------------------------------------------------------------------
template <int MT, int KHW_T, int MODE>
__global__ void __launch_bounds__(512)
conv_px_kernel(const ConvKArgs a, const int groups, const int slices, const int pw_count) {
    extern __shared__ __align__(16) float px_smem[];
    const ConvGeom& g = a.g;
    const int Cd = MODE == MODE_FWD ? g.Cout : g.Cin;
    const int Hd = MODE == MODE_FWD ? g.OH : g.H;
    const int Wd = MODE == MODE_FWD ? g.OW : g.W;
    const long long d_ss = MODE == MODE_FWD ? g.out_sstride : g.in_sstride;
    const int Cs = MODE == MODE_FWD ? g.Cin : g.Cout;
    const int Hs = MODE == MODE_FWD ? g.H : g.OH;
    const int Ws = MODE == MODE_FWD ? g.W : g.OW;
    const long long s_ss = MODE == MODE_FWD ? g.in_sstride : g.out_sstride;
    const int KHW = KHW_T > 0 ? KHW_T : g.KH * g.KW;
    const int Ktot = Cs * KHW;
    const int HWd = Hd * Wd, HWs = Hs * Ws;
    const long long J = (long long)g.batch * HWd;
    const int Mpad = groups * MT;
    float* Wsm = px_smem;                       // [Ktot][Mpad]
    float* red = px_smem + (size_t)Ktot * Mpad; // [slices][pw_count][Mpad][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pw = warp % pw_count;
    const int gi = (warp / pw_count) % groups;
    const int sl = warp / (pw_count * groups);
    const long long j = ((long long)blockIdx.x * pw_count + pw) * 32 + lane;
    const bool ok = j < J;
    int n = 0, y = 0, x = 0, pix = 0;
    if (ok) {
        n = (int)(j / HWd);
        pix = (int)(j - (long long)n * HWd);
        y = pix / Wd;
        x = pix - y * Wd;
    }
    // per-thread tap table (offset inside a source channel plane, -1 when the tap falls outside)
    int toff[KHW_T > 0 ? KHW_T : 1];
    if (KHW_T > 0) {
#pragma unroll
        for (int t = 0; t < (KHW_T > 0 ? KHW_T : 1); ++t) {
            const int ky = t / g.KW, kx = t - ky * g.KW;
            int sy, sx;
            bool v;
            if (MODE == MODE_FWD) {
                sy = y * g.sh + ky - g.ph; sx = x * g.sw + kx - g.pw;
                v = sy >= 0 && sy < Hs && sx >= 0 && sx < Ws;
            } else {
                const int ty_ = y + g.ph - ky, tx_ = x + g.pw - kx;
                sy = ty_ / g.sh; sx = tx_ / g.sw;
                v = ty_ >= 0 && tx_ >= 0 && sy * g.sh == ty_ && sx * g.sw == tx_ && sy < Hs && sx < Ws;
            }
            toff[t] = (ok && v) ? sy * Ws + sx : -1;
        }
    }

    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;

    for (int p = 0; p < a.npairs; ++p) {
        __syncthreads();
        const float* __restrict__ wt = a.wt[p];
        const float sc = a.scale[p];
        if (MODE == MODE_FWD) {
            // global layout [m][k]: consecutive threads read consecutive k (coalesced), scatter into [k][m]
            for (int idx = tid; idx < Ktot * Mpad; idx += blockDim.x) {
                const int m = idx / Ktot, k = idx - m * Ktot;
                Wsm[(size_t)k * Mpad + m] = m < Cd ? wt[(long long)m * Ktot + k] * sc : 0.f;
            }
        } else {
            // global layout [c][m][t] (c = conv output channel): read contiguous runs of (m,t)
            const int MK = Mpad * KHW;
            for (int idx = tid; idx < Cs * MK; idx += blockDim.x) {
                const int c = idx / MK, r = idx - c * MK;
                const int m = r / KHW, t = r - m * KHW;
                Wsm[((size_t)c * KHW + t) * Mpad + m] = m < Cd ? wt[((long long)c * g.Cin + m) * KHW + t] * sc : 0.f;
            }
        }
        __syncthreads();
        const float* __restrict__ src = a.act[p] + (long long)n * s_ss;
        for (int c = sl; c < Cs; c += slices) {
            const float* __restrict__ sc_ = src + (long long)c * HWs;
            const float* wbase = Wsm + (size_t)c * KHW * Mpad + gi * MT;
            if (KHW_T > 0) {
                float xv[KHW_T > 0 ? KHW_T : 1];
#pragma unroll
                for (int t = 0; t < (KHW_T > 0 ? KHW_T : 1); ++t) xv[t] = toff[t] >= 0 ? sc_[toff[t]] : 0.f;
#pragma unroll
                for (int t = 0; t < (KHW_T > 0 ? KHW_T : 1); ++t) {
                    const float4* w4 = reinterpret_cast<const float4*>(wbase + t * Mpad);
#pragma unroll
                    for (int q = 0; q < MT / 4; ++q) {
                        const float4 w = w4[q];
                        acc[4 * q + 0] = fmaf(w.x, xv[t], acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(w.y, xv[t], acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(w.z, xv[t], acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(w.w, xv[t], acc[4 * q + 3]);
                    }
                }
            } else {
                for (int ky = 0; ky < g.KH; ++ky)
                    for (int kx = 0; kx < g.KW; ++kx) {
                        int sy, sx;
                        bool v;
                        if (MODE == MODE_FWD) {
                            sy = y * g.sh + ky - g.ph; sx = x * g.sw + kx - g.pw;
                            v = sy >= 0 && sy < Hs && sx >= 0 && sx < Ws;
                        } else {
                            const int ty_ = y + g.ph - ky, tx_ = x + g.pw - kx;
                            sy = ty_ / g.sh; sx = tx_ / g.sw;
                            v = ty_ >= 0 && tx_ >= 0 && sy * g.sh == ty_ && sx * g.sw == tx_ && sy < Hs && sx < Ws;
                        }
                        const float xv = (ok && v) ? sc_[sy * Ws + sx] : 0.f;
                        const float4* w4 = reinterpret_cast<const float4*>(wbase + (ky * g.KW + kx) * Mpad);
#pragma unroll
                        for (int q = 0; q < MT / 4; ++q) {
                            const float4 w = w4[q];
                            acc[4 * q + 0] = fmaf(w.x, xv, acc[4 * q + 0]);
                            acc[4 * q + 1] = fmaf(w.y, xv, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(w.z, xv, acc[4 * q + 2]);
                            acc[4 * q + 3] = fmaf(w.w, xv, acc[4 * q + 3]);
                        }
                    }
            }
        }
    }
    // ---- sum the K slices: every slice parks its partials, then slice `sl` of each (pixel warp, group)
    // adds up channels mm = sl, sl+slices, ... over all slices in fixed order (deterministic)
    const size_t slice_stride = (size_t)pw_count * Mpad * 32;
    float* myred = red + ((size_t)pw * Mpad + gi * MT) * 32 + lane;
    {
        float* dst = myred + (size_t)sl * slice_stride;
#pragma unroll
        for (int m = 0; m < MT; ++m) dst[m * 32] = acc[m];
    }
    __syncthreads();
    if (!ok) return;
    const long long base = (long long)n * d_ss + pix;
    for (int mm = sl; mm < MT; mm += slices) {
        const int m = gi * MT + mm;
        if (m >= Cd) break;
        const long long o = base + (long long)m * HWd;
        float v = 0.f;
        for (int r = 0; r < slices; ++r) v += myred[(size_t)r * slice_stride + mm * 32];
        if (a.bias) v += a.bias[m];
        if (a.accumulate) v += a.out[o];
        if (a.relu_mode == 1) v = v > 0.f ? v : 0.f;
        else if (a.relu_mode == 2) v = a.relu_ref[o] > 0.f ? v : 0.f;
        a.out[o] = v;
    }
}

template <int MT, int MODE>
static int launch_px_t(cudaStream_t st, const ConvKArgs& a, int groups, int slices, int pw, int khw, size_t smem,
                       int blocks) {
    const int threads = 32 * pw * groups * slices;
#define B2S_PX_LAUNCH(KT)                                                                                  \
    do {                                                                                                   \
        if (smem > 48 * 1024)                                                                              \
            cudaFuncSetAttribute(conv_px_kernel<MT, KT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem);                                                              \
        conv_px_kernel<MT, KT, MODE><<<blocks, threads, smem, st>>>(a, groups, slices, pw);                \
    } while (0)
    if (khw == 1) B2S_PX_LAUNCH(1);
    else if (khw == 9) B2S_PX_LAUNCH(9);
    else B2S_PX_LAUNCH(0);
#undef B2S_PX_LAUNCH
    return 0;
}

// returns true when the pixel-thread kernel was used
template <int MODE>
static bool try_launch_px(cudaStream_t st, const ConvKArgs& a) {
    const ConvGeom& g = a.g;
    const int Cd = MODE == MODE_FWD ? g.Cout : g.Cin;
    const int Cs = MODE == MODE_FWD ? g.Cin : g.Cout;
    const int khw = g.KH * g.KW;
    const long long J = (long long)g.batch * (MODE == MODE_FWD ? g.OH * g.OW : g.H * g.W);
    if (Cd > 160) return false;
    // register block: least padding, then the widest
    const int cand[5] = {32, 24, 16, 12, 8};
    int MT = 8, best_pad = 1 << 30;
    for (int c : cand) {
        const int pad = (Cd + c - 1) / c * c;
        if (pad < best_pad) { best_pad = pad; MT = c; }
    }
    const int groups = best_pad / MT;
    if (groups > 16) return false;
    const size_t wbytes = (size_t)Cs * khw * best_pad * sizeof(float);
    if (wbytes > 160 * 1024) return false;
    int pw = 4;
    while (pw > 1 && (pw * groups > 16 || J / (32 * pw) < 2 * kNumSMs)) pw >>= 1;
    if (pw * groups > 16) return false;
    int slices = 16 / (pw * groups);
    if (slices > Cs) slices = Cs;
    if (slices < 1) slices = 1;
    size_t smem = wbytes + (size_t)slices * pw * best_pad * 32 * sizeof(float);
    while (smem > 200 * 1024 && slices > 1) {
        slices >>= 1;
        smem = wbytes + (size_t)slices * pw * best_pad * 32 * sizeof(float);
    }
    if (smem > 200 * 1024) return false;
    const int blocks = (int)((J + 32 * pw - 1) / (32 * pw));
    switch (MT) {
    case 32: launch_px_t<32, MODE>(st, a, groups, slices, pw, khw, smem, blocks); break;
    case 24: launch_px_t<24, MODE>(st, a, groups, slices, pw, khw, smem, blocks); break;
    case 16: launch_px_t<16, MODE>(st, a, groups, slices, pw, khw, smem, blocks); break;
    case 12: launch_px_t<12, MODE>(st, a, groups, slices, pw, khw, smem, blocks); break;
    default: launch_px_t<8, MODE>(st, a, groups, slices, pw, khw, smem, blocks); break;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// wgrad: Wbar[m, (c,ky,kx)] += sum_j adj[n, m, oy, ox] * act[n, c, oy*sh+ky-ph, ox*sw+kx-pw]
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_wgrad_kernel(const ConvKArgs a) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TXN = BN / TN;
    static_assert(NT % BK == 0, "thread count must be a multiple of BK");
    constexpr int ROWS = NT / BK;            // rows (m or n) loaded per pass
    constexpr int A_IT = (BM + ROWS - 1) / ROWS, B_IT = (BN + ROWS - 1) / ROWS;
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];

    const ConvGeom& g = a.g;
    const int KHW = g.KH * g.KW;
    const int Ncol = g.Cin * KHW;
    const int OHW = g.OH * g.OW;
    const long long J = (long long)g.batch * OHW;
    const long long jbeg = (long long)blockIdx.z * a.k_chunk;
    const long long jend = jbeg + a.k_chunk < J ? jbeg + a.k_chunk : J;

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int jb = tid % BK, r0 = tid / BK;

    // decode the tile's B columns once per CTA (reads below are warp-uniform broadcasts)
    __shared__ int col_off[BN];
    __shared__ int col_tap[BN];
    for (int nb = tid; nb < BN; nb += NT) {
        const int col = n0 + nb;
        if (col < Ncol) {
            const int c = col / KHW, t = col - c * KHW;
            const int ky = t / g.KW, kx = t - ky * g.KW;
            col_off[nb] = c * g.H * g.W;
            col_tap[nb] = (ky << 16) | kx;
        } else {
            col_off[nb] = -1;
            col_tap[nb] = 0;
        }
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int p = 0; p < a.npairs; ++p) {
        const float* __restrict__ act = a.act[p];
        const float* __restrict__ adj = a.wt[p];
        const float sc = a.scale[p];
        for (long long k0 = jbeg; k0 < jend; k0 += BK) {
            const long long j = k0 + jb;
            const bool ok = j < jend;
            int n = 0, oy = 0, ox = 0, pix = 0;
            if (ok) {
                n = (int)(j / OHW);
                pix = (int)(j - (long long)n * OHW);
                oy = pix / g.OW;
                ox = pix - oy * g.OW;
            }
            const float* __restrict__ adj_n = adj + (long long)n * g.out_sstride + pix;
            const float* __restrict__ act_n = act + (long long)n * g.in_sstride;
#pragma unroll
            for (int i = 0; i < A_IT; ++i) {
                const int mb = r0 + i * ROWS;
                if (mb < BM) {
                    const int m = m0 + mb;
                    float v = 0.f;
                    if (ok && m < g.Cout) v = adj_n[(long long)m * OHW] * sc;
                    As[jb][mb] = v;
                }
            }
            const int iy0 = oy * g.sh - g.ph, ix0 = ox * g.sw - g.pw;
#pragma unroll
            for (int i = 0; i < B_IT; ++i) {
                const int nb = r0 + i * ROWS;
                if (nb < BN) {
                    float v = 0.f;
                    const int coff = col_off[nb];
                    if (ok && coff >= 0) {
                        const int tap = col_tap[nb];
                        const int sy = iy0 + (tap >> 16), sx = ix0 + (tap & 0xffff);
                        if (sy >= 0 && sy < g.H && sx >= 0 && sx < g.W) v = act_n[coff + sy * g.W + sx];
                    }
                    Bs[jb][nb] = v;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float av[TM], bv[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
                for (int jn = 0; jn < TN; ++jn) bv[jn] = Bs[kk][tx + jn * TXN];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int jn = 0; jn < TN; ++jn) acc[i][jn] = fmaf(av[i], bv[jn], acc[i][jn]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.Cout) continue;
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) {
            const int col = n0 + tx + jn * TXN;
            if (col < Ncol) atomicAdd(a.out + (long long)m * Ncol + col, acc[i][jn]);
        }
    }
}

// bbar[c] += sum over (n, pix) of adj.  grid = (C, samples, chunks of a plane): contiguous 128-bit loads of one
// (sample, channel) plane slice per block, no integer division per element (round 1: a 64-bit division per element
// made the 15 launches of VGG16 cost 5.5 ms per HVP).
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ adj, int batch, int C, int HW,
                                                        long long sstride, float* __restrict__ bbar) {
    __shared__ float red[32];
    const int c = blockIdx.x;
    float s = 0.f;
    for (int n = blockIdx.y; n < batch; n += gridDim.y) {
        const float* __restrict__ p = adj + (long long)n * sstride + (long long)c * HW;
        if (((HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4* p4 = reinterpret_cast<const float4*>(p);
            for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < (HW >> 2); i += gridDim.z * blockDim.x) {
                const float4 v = __ldg(p4 + i);
                s += (v.x + v.y) + (v.z + v.w);
            }
        } else {
            for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < HW; i += gridDim.z * blockDim.x) s += __ldg(p + i);
        }
    }
    float v[1] = {s};
    block_sum<1, float>(v, red);
    if (threadIdx.x == 0) atomicAdd(bbar + c, v[0]);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int MODE>
static int launch_gather(cudaStream_t st, const ConvKArgs& a) {
    const ConvGeom& g = a.g;
    const int Cd = MODE == MODE_FWD ? g.Cout : g.Cin;
    const long long J = (long long)g.batch * (MODE == MODE_FWD ? g.OH * g.OW : g.H * g.W);
    const double macs = (double)g.batch * g.OH * g.OW * g.Cout * g.Cin * g.KH * g.KW;      // per pair
    const double io = 4.0 * ((double)g.batch * g.Cin * g.H * g.W * a.npairs + (double)g.batch * g.Cout * g.OH * g.OW +
                             (double)g.Cout * g.Cin * g.KH * g.KW * a.npairs);
    ProfScope prof(MODE == MODE_FWD ? "conv_fwd" : "conv_dgrad", 2.0 * macs * a.npairs, io, st);
    if (skip_family(MODE == MODE_FWD ? "conv_fwd" : "conv_dgrad")) return 0;
    {
        const int rc = try_launch_conv_tma(MODE, st, a);     // TMA-fed tcgen05 path (stride 1, 16-byte row pitch)
        if (rc < 0) return rc;
        if (rc == 1) {
            B2S_LAUNCH_CHECK();
            return 0;
        }
    }
    {
        const int rc = try_launch_conv_tc(MODE, st, a);      // tcgen05 path for wide layers
        if (rc < 0) return rc;
        if (rc == 1) {
            B2S_LAUNCH_CHECK();
            return 0;
        }
    }
    if (try_launch_px<MODE>(st, a)) {
        B2S_LAUNCH_CHECK();
        return 0;
    }
    if (Cd <= 16) {
        dim3 grid(cdiv(J, 256), cdiv(Cd, 16));
        conv_gather_gemm_kernel<16, 256, 16, 4, 4, MODE><<<grid, 256, 0, st>>>(a);
    } else if (Cd <= 32) {
        dim3 grid(cdiv(J, 128), cdiv(Cd, 32));
        conv_gather_gemm_kernel<32, 128, 16, 4, 4, MODE><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid(cdiv(J, 64), cdiv(Cd, 64));
        conv_gather_gemm_kernel<64, 64, 16, 4, 4, MODE><<<grid, 256, 0, st>>>(a);
    }
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_conv_fwd(cudaStream_t st, const ConvGeom& g, int npairs, const float* const* act,
                    const float* const* wt, const float* scale, const float* bias, int relu_mode,
                    const float* relu_ref, float* out, int accumulate, const float* const* pack) {
    ConvKArgs a{};
    a.g = g;
    a.npairs = npairs;
    for (int p = 0; p < npairs; ++p) { a.act[p] = act[p]; a.wt[p] = wt[p]; a.scale[p] = scale[p]; a.pack[p] = pack ? pack[p] : nullptr; }
    a.bias = bias; a.relu_mode = relu_mode; a.relu_ref = relu_ref; a.out = out; a.accumulate = accumulate;
    return launch_gather<MODE_FWD>(st, a);
}

int launch_conv_dgrad(cudaStream_t st, const ConvGeom& g, int npairs, const float* const* adj,
                      const float* const* wt, const float* scale, float* out, int accumulate,
                      const float* const* pack) {
    ConvKArgs a{};
    a.g = g;
    a.npairs = npairs;
    for (int p = 0; p < npairs; ++p) { a.act[p] = adj[p]; a.wt[p] = wt[p]; a.scale[p] = scale[p]; a.pack[p] = pack ? pack[p] : nullptr; }
    a.bias = nullptr; a.relu_mode = 0; a.relu_ref = nullptr; a.out = out; a.accumulate = accumulate;
    return launch_gather<MODE_DGRAD>(st, a);
}

int launch_conv_wgrad(cudaStream_t st, const ConvGeom& g, int npairs, const float* const* act,
                      const float* const* adj, const float* scale, float* wbar) {
    ConvKArgs a{};
    a.g = g;
    a.npairs = npairs;
    for (int p = 0; p < npairs; ++p) { a.act[p] = act[p]; a.wt[p] = adj[p]; a.scale[p] = scale[p]; }
    a.out = wbar;
    const double macs = (double)g.batch * g.OH * g.OW * g.Cout * g.Cin * g.KH * g.KW;
    ProfScope prof("conv_wgrad", 2.0 * macs * npairs,
                   4.0 * npairs * ((double)g.batch * g.Cin * g.H * g.W + (double)g.batch * g.Cout * g.OH * g.OW), st);
    if (skip_family("conv_wgrad")) return 0;
    {
        int rc = try_launch_wgrad_tma(st, a);                // tcgen05 path, TMA-staged (small maps: DenseNet3, USPS)
        if (rc == 0) rc = try_launch_wgrad_tc(st, a);        // tcgen05 path, operands gathered by the transform threads
        if (rc < 0) return rc;
        if (rc == 1) {
            B2S_LAUNCH_CHECK();
            return 0;
        }
    }
    const int Ncol = g.Cin * g.KH * g.KW;
    const long long J = (long long)g.batch * g.OH * g.OW;
    constexpr int BK = 32;
    const long long ktiles = (J + BK - 1) / BK;
    auto pick_split = [&](int tiles_mn) {
        long long want = (4LL * kNumSMs + tiles_mn - 1) / tiles_mn;
        if (want < 1) want = 1;
        if (want > ktiles) want = ktiles;
        long long per = (ktiles + want - 1) / want;     // k tiles per split
        a.k_chunk = (int)(per * BK);
        return (int)((ktiles + per - 1) / per);
    };
    if (g.Cout <= 16) {
        const int gm = cdiv(g.Cout, 16), gn = cdiv(Ncol, 256);
        dim3 grid(gn, gm, pick_split(gm * gn));
        conv_wgrad_kernel<16, 256, BK, 4, 4><<<grid, 256, 0, st>>>(a);
    } else if (g.Cout <= 32) {
        const int gm = cdiv(g.Cout, 32), gn = cdiv(Ncol, 128);
        dim3 grid(gn, gm, pick_split(gm * gn));
        conv_wgrad_kernel<32, 128, BK, 4, 4><<<grid, 256, 0, st>>>(a);
    } else {
        const int gm = cdiv(g.Cout, 64), gn = cdiv(Ncol, 64);
        dim3 grid(gn, gm, pick_split(gm * gn));
        conv_wgrad_kernel<64, 64, BK, 4, 4><<<grid, 256, 0, st>>>(a);
    }
    B2S_LAUNCH_CHECK();
    return 0;
}

int launch_bias_grad(cudaStream_t st, const float* adj, int batch, int C, int HW, long long sstride,
                     float* bbar) {
    // about 4 blocks per SM: samples first, then chunks of a plane (each thread should still see >= 8 float4)
    const int cap = std::max(1, (4 * kNumSMs + C - 1) / C);
    const int ny = std::min(batch, cap);
    int nz = std::min(std::max(1, cap / ny), std::max(1, HW / (256 * 4 * 8)));
    if (nz > 64) nz = 64;
    ProfScope prof("bias_grad", (double)batch * C * HW, 4.0 * batch * C * HW, st);
    dim3 grid(C, ny, nz);
    bias_grad_kernel<<<grid, 256, 0, st>>>(adj, batch, C, HW, sstride, bbar);
    B2S_LAUNCH_CHECK();
    return 0;
}

}  // namespace b2s
