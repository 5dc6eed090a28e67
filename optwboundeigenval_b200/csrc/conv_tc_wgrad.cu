// conv_tc_wgrad.cu -- tcgen05 / TMEM weight-gradient contraction of the Conv2d / Linear jets (sm_100a).
//
//   Wbar[co][ci][ky][kx] += sum_p scale_p * sum_{n,oy,ox} g_p[n,co,oy,ox] * x_p[n,ci,oy+ky-ph,ox+kx-pw]
//
// (order K of a layer: up to three (x_j, g_{K-j}) pairs with binomial scales; this is
// the Hv slice of the layer for K = 1.)  GEMM view: the contraction index K is the OUTPUT PIXEL, which is
// the contiguous dimension of both NCHW operands, so both are K-major as they lie in memory:
//   A [128 rows = (tap, ci)] x [32 pixels]   the input, shifted per tap, zero outside the image
//   B [BN  rows = co]        x [32 pixels]   the output adjoint
//   D [128 x BN] in TMEM, one fresh accumulator per 32-pixel k-block, drained into fp32 registers
//   (same fp32-accurate 3xTF32 scheme and the same reason as conv_tc.cu).
// A warp instruction loads an [8 rows x 4 pixel-groups] patch with 128-bit accesses (64 contiguous bytes per
// row), splits it hi/lo in registers and stores it with conflict-free 128-bit shared-memory stores straight
// into the canonical K-major core-matrix layout; a tap with a horizontal shift adds one scalar load per
// patch for the element that crosses the 16-byte boundary.  fence.proxy.async hands the stage to the
// tensor core (SS MMA).  The grid splits the pixel range (split-K) so that all 148 SMs work; each CTA
// atomically adds its partial tile into the flat gradient vector.
//
// 128-bit variant: stride 1, horizontal shift in {-1, 0, +1}, OW and W multiples of 4.  Every other geometry takes the
// scalar-gather variant (VEC = false: one predicated load per element, pixels decoded one by one).
// Contractions too small to be worth a tensor-core launch stay on the CUDA-core kernel (conv.cu).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "conv_args.h"
#include "tc_common.cuh"

namespace b2s {

constexpr int WG_NST = 3;                      // shared-memory stages
constexpr int WG_G = 4;                        // k-blocks accumulated in TMEM between drains (as conv_tma.cu: 16 truncating
                                               // accumulations of K = 8 instead of 4; the drain was ~3 NACC BN TMEM columns
                                               // per thread and k-block)
constexpr int WG_A_FLOATS = TC_M * TC_KB;      // one of hi / lo

template <int BN>
struct WgSmem {
    static constexpr int B_FLOATS = BN * TC_KB;
    static constexpr int STAGE_FLOATS = 2 * WG_A_FLOATS + 2 * B_FLOATS;
    static constexpr size_t BYTES = (size_t)WG_NST * STAGE_FLOATS * sizeof(float) + 256;
};

// hi = TF32 round-to-nearest (integer add of half an ulp, low 13 bits cleared: 2 instructions instead of the 4 of
// cvt.rna's NaN-safe expansion), lo = v - hi exact in fp32; the tensor core reads the upper 19 bits of lo
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi.x = tf32_hi(v.x); lo.x = v.x - hi.x;
    hi.y = tf32_hi(v.y); lo.y = v.y - hi.y;
    hi.z = tf32_hi(v.z); lo.z = v.z - hi.z;
    hi.w = tf32_hi(v.w); lo.w = v.w - hi.w;
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad_kernel(const ConvKArgs a, const int mtiles, const int ntiles, const int ksplit, const int swapped) {
    extern __shared__ __align__(1024) uint8_t wg_smem[];
    using S = WgSmem<BN>;
    float* stages = reinterpret_cast<float*>(wg_smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wg_smem + (size_t)WG_NST * S::STAGE_FLOATS * sizeof(float));
    uint64_t* ab_full = bars;                  // [NST] the 128 transform threads of the owning group
    uint64_t* ab_free = bars + WG_NST;         // [NST] tcgen05.commit
    uint64_t* d_full = bars + 2 * WG_NST;      // [2]
    uint64_t* d_empty = d_full + 2;            // [2]   128 drain threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const ConvGeom& g = a.g;
    const int KHW = g.KH * g.KW;
    const int OHW = g.OH * g.OW, HW = g.H * g.W;
    const long long J = (long long)g.batch * OHW;
    const int nkb = (int)((J + TC_KB - 1) / TC_KB);
    const int Mrows = KHW * g.Cin;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mt = blockIdx.x % mtiles;
    const int nt = (blockIdx.x / mtiles) % ntiles;
    const int ksl = blockIdx.x / (mtiles * ntiles);
    const int per = (nkb + ksplit - 1) / ksplit;
    const int kb0 = ksl * per;
    const int nloc = max(0, min(nkb, kb0 + per) - kb0);         // k-blocks of this CTA per pair
    const int total = a.npairs * nloc;

    if (tid == 0) {
        for (int s = 0; s < WG_NST; ++s) {
            mbar_init(&ab_full[s], 128);
            mbar_init(&ab_free[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // row groups past the last row of the A tile are never written again: zero the stages once
        float4* z = reinterpret_cast<float4*>(wg_smem);
        const int n4 = WG_NST * S::STAGE_FLOATS / 4;
        for (int i = tid; i < n4; i += TC_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ===================== operand transform: global -> registers (split) -> shared memory =========
        const int grp = warp >> 2, wq = warp & 3;
        const int l8 = lane & 7, l4 = lane >> 3;
        // per-lane constants of this warp's four A row groups: channel-plane offset (+ vertical tap shift),
        // vertical / horizontal shift, row validity.  A row group that lies entirely past the last row is
        // never touched (its shared memory was zeroed once at kernel start).
        int rowoff[4], dyv[4], dxv[4];
        bool rok[4], gok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int m0 = mt * TC_M + (wq * 4 + q) * 8;
            const int m = m0 + l8;
            gok[q] = m0 < Mrows;
            rok[q] = m < Mrows;
            const int t = rok[q] ? m / g.Cin : 0;
            const int ci = rok[q] ? m - t * g.Cin : 0;
            const int ky = t / g.KW, kx = t - ky * g.KW;
            dyv[q] = ky - g.ph;
            dxv[q] = kx - g.pw;
            rowoff[q] = ci * HW + dyv[q] * g.W + (VEC ? 0 : dxv[q]);
        }
        constexpr int NB = BN / 16;                          // B patches of this warp
        int boff[NB];
        bool bok[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int idx = wq + 4 * u;
            const int co = nt * BN + (idx >> 1) * 8 + l8;
            bok[u] = co < g.Cout;
            boff[u] = bok[u] ? co * OHW : 0;
        }
        for (int i = grp; i < total; i += 2) {
            const int p = i / nloc;
            const int kb = kb0 + (i - p * nloc);
            const int s = i % WG_NST;
            const uint32_t round = i / WG_NST;
            float* Ahi = stages + (size_t)s * S::STAGE_FLOATS;
            float* Alo = Ahi + WG_A_FLOATS;
            float* Bhi = Alo + WG_A_FLOATS;
            float* Blo = Bhi + S::B_FLOATS;
            const float* __restrict__ xp = a.act[p];
            const float* __restrict__ gp = a.wt[p];
            float4 ca[4][2];
            float ea[4][2];
            float4 cb[NB];
            if (VEC) {
            // the two 4-pixel groups of this lane inside the k-block: k-groups l4 and l4 + 4
            int oy_[2], ox_[2];
            long long xo[2], go[2];
            bool pv[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long j = (long long)kb * TC_KB + (h * 4 + l4) * 4;
                pv[h] = j < J;
                const unsigned jj = pv[h] ? (unsigned)j : 0u;
                const unsigned n = jj / (unsigned)OHW;
                const unsigned rem = jj - n * OHW;
                oy_[h] = rem / g.OW;
                ox_[h] = rem - oy_[h] * g.OW;
                xo[h] = (long long)n * g.in_sstride + oy_[h] * g.W + ox_[h];
                go[h] = (long long)n * g.out_sstride + rem;
            }
            // ---- issue every load of this warp's share of the stage
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!gok[q]) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int sy = oy_[h] + dyv[q];
                    const bool ok = rok[q] && pv[h] && sy >= 0 && sy < g.H;
                    const float* ptr = xp + xo[h] + rowoff[q];
                    ca[q][h] = ok ? __ldg(reinterpret_cast<const float4*>(ptr)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float e = 0.f;
                    if (dxv[q] < 0) { if (ok && ox_[h] > 0) e = __ldg(ptr - 1); }
                    else if (dxv[q] > 0) { if (ok && ox_[h] + 4 < g.W) e = __ldg(ptr + 4); }
                    ea[q][h] = e;
                }
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int h = (wq + 4 * u) & 1;
                const bool ok = pv[h] && bok[u];
                cb[u] = ok ? __ldg(reinterpret_cast<const float4*>(gp + go[h] + boff[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            } else {
                // general geometry (any width, stride, kernel size): the four pixels of a k-group are decoded one by one
                // (a group may straddle image rows or samples) and every element is one predicated scalar load; the tap
                // shift is part of rowoff.  Serves the 14 x 14 / 7 x 7 maps of the chest models (VGG16 conv5_x, DenseNet121
                // blocks 3 / 4), whose rows are not 16-byte aligned.
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long j = (long long)kb * TC_KB + (h * 4 + l4) * 4;
                    const unsigned jj = j < J ? (unsigned)j : 0u;
                    int n = (int)(jj / (unsigned)OHW);
                    const int rem = (int)(jj - (unsigned)n * OHW);
                    int oy = rem / g.OW;
                    int ox = rem - oy * g.OW;
                    float av[4][4], bv[NB][4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool pv = j + e < J;
                        const int iy0 = oy * g.sh, ix0 = ox * g.sw;
                        const float* __restrict__ xb = xp + (long long)n * g.in_sstride + iy0 * g.W + ix0;
                        const float* __restrict__ gb = gp + (long long)n * g.out_sstride + oy * g.OW + ox;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bool ok = gok[q] && rok[q] && pv && (unsigned)(iy0 + dyv[q]) < (unsigned)g.H &&
                                            (unsigned)(ix0 + dxv[q]) < (unsigned)g.W;
                            av[q][e] = ok ? __ldg(xb + rowoff[q]) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            if (((wq + 4 * u) & 1) == h) bv[u][e] = (pv && bok[u]) ? __ldg(gb + boff[u]) : 0.f;
                        }
                        if (++ox == g.OW) { ox = 0; if (++oy == g.OH) { oy = 0; ++n; } }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        ca[q][h] = make_float4(av[q][0], av[q][1], av[q][2], av[q][3]);
                        ea[q][h] = 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < NB; ++u)
                        if (((wq + 4 * u) & 1) == h) cb[u] = make_float4(bv[u][0], bv[u][1], bv[u][2], bv[u][3]);
                }
            }
            // ---- the stage must have been read by the MMAs of its previous use
            if (round > 0) mbar_wait(&ab_free[s], (round - 1) & 1);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!gok[q]) continue;
                const int rg = wq * 4 + q;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 v = ca[q][h];
                    if (VEC) {
                        if (dxv[q] < 0) v = make_float4(ea[q][h], v.x, v.y, v.z);
                        else if (dxv[q] > 0) v = make_float4(v.y, v.z, v.w, ea[q][h]);
                    }
                    float4 hi, lo;
                    split4(v, hi, lo);
                    const int o = rg * 256 + h * 128 + lane * 4;       // ((rg*8 + kgroup) * 32) + (row%8)*4, kgroup = h*4 + lane/8
                    *reinterpret_cast<float4*>(Ahi + o) = hi;
                    *reinterpret_cast<float4*>(Alo + o) = lo;
                }
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int idx = wq + 4 * u;
                const int rgB = idx >> 1, h = idx & 1;
                float4 hi, lo;
                split4(cb[u], hi, lo);
                const int o = rgB * 256 + h * 128 + lane * 4;
                *reinterpret_cast<float4*>(Bhi + o) = hi;
                *reinterpret_cast<float4*>(Blo + o) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
            mbar_arrive(&ab_full[s]);
        }
    } else if (warp < 12) {
        // ===================== drain + epilogue ======================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_DRAIN));
        const int q = warp - 8;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        constexpr int NACC = tc_nacc(BN) > 3 ? 3 : tc_nacc(BN);
        float acc[BN];
#pragma unroll
        for (int i = 0; i < BN; ++i) acc[i] = 0.f;
        uint32_t grp = 0;
        for (int p = 0; p < a.npairs; ++p) {
            const float sc = a.scale[p];
            for (int k0 = 0; k0 < nloc; k0 += WG_G, ++grp) {
                const int b = grp & 1;
                mbar_wait(&d_full[b], (grp >> 1) & 1);
                __syncwarp();
                tc_fence_after();
                const uint32_t d0 = lane_addr + TC_DCOL0 + b * TC_DCOLS;
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[NACC][16];
#pragma unroll
                    for (int q2 = 0; q2 < NACC; ++q2) tmem_ld16(d0 + q2 * BN + c0, v[q2]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float d;
                        if (NACC >= 3) d = (__uint_as_float(v[0][e]) + __uint_as_float(v[2][e])) + __uint_as_float(v[1][e]);
                        else if (NACC == 2) d = __uint_as_float(v[1][e]) + __uint_as_float(v[0][e]);
                        else d = __uint_as_float(v[0][e]);
                        acc[c0 + e] = fmaf(d, sc, acc[c0 + e]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&d_empty[b]);
            }
        }
        const int m = mt * TC_M + r;
        if (total > 0 && m < Mrows) {
            const int t = m / g.Cin, ci = m - t * g.Cin;
            if (!swapped) {
                // row = (tap, input channel), column = output channel:  Wbar[co][ci][t]
                float* __restrict__ wrow = a.out + (long long)ci * KHW + t;
#pragma unroll
                for (int i = 0; i < BN; ++i) {
                    const int co = nt * BN + i;
                    if (co < g.Cout) atomicAdd(wrow + (long long)co * g.Cin * KHW, acc[i]);
                }
            } else {
                // operands were exchanged (launcher): row = (mirrored tap, OUTPUT channel), column = INPUT channel;
                // g.Cin is the layer's Cout and g.Cout its Cin here
                float* __restrict__ wrow = a.out + (long long)ci * g.Cout * KHW + (KHW - 1 - t);
#pragma unroll
                for (int i = 0; i < BN; ++i) {
                    const int c2 = nt * BN + i;
                    if (c2 < g.Cout) atomicAdd(wrow + (long long)c2 * KHW, acc[i]);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_MISC));
        if (warp == TC_WARP_MMA) {
            // ===================== MMA issuer ======================================================
            const uint32_t idesc = umma_idesc_tf32(TC_M, BN);
            const uint32_t idesc2 = umma_idesc_tf32(TC_M, 2 * BN <= 256 ? 2 * BN : BN);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            const uint64_t desc0 = umma_desc(smem_u32(wg_smem), 128, 1024);      // stage 0, A hi, k-step 0
            constexpr int NACC = tc_nacc(BN) > 3 ? 3 : tc_nacc(BN);
            constexpr uint32_t STAGE16 = (uint32_t)(S::STAGE_FLOATS * 4) >> 4;
            constexpr uint32_t A16 = (uint32_t)(WG_A_FLOATS * 4) >> 4, B16 = (uint32_t)(S::B_FLOATS * 4) >> 4;
            uint32_t grp = 0;
            for (int i = 0; i < total; ++i) {
                const int s = i % WG_NST;
                const uint32_t round = i / WG_NST;
                const int kk = i % nloc;                              // k-block inside its pair: drain groups never span pairs
                const bool first = (kk % WG_G) == 0;
                const bool last = (kk % WG_G) == WG_G - 1 || kk == nloc - 1;
                const int b = grp & 1;
                const uint32_t use = grp >> 1;
                mbar_wait(&ab_full[s], round & 1);
                if (first && use > 0) mbar_wait(&d_empty[b], (use - 1) & 1);
                __syncwarp();
                tc_fence_after();
                const uint32_t d_addr = tmem_u + TC_DCOL0 + b * TC_DCOLS;
                const uint64_t dAh = desc0 + (uint64_t)(s * STAGE16), dAl = dAh + A16;
                const uint64_t dBh = dAl + A16, dBl = dBh + B16;
                const uint32_t cont = first ? 0u : 1u;                // accumulate onto the group's earlier k-blocks
                if (elect_one()) {
                    // set layout [lo*hi | hi*hi | hi*lo] (3 accumulators) or [hi*hi | hi*lo + lo*hi] (2): the hi and lo tiles of
                    // B are adjacent in shared memory, so A_hi x [B_hi | B_lo] is one MMA with N = 2 BN (see conv_tma.cu)
#pragma unroll
                    for (int ks = 0; ks < TC_KB / 8; ++ks) {
                        const uint64_t ko = (uint64_t)(ks * 16);
                        const uint32_t accf = ks >= 1 ? 1u : cont;
                        if (NACC == 3) {
                            umma_tf32_ss(d_addr + BN, dAh + ko, dBh + ko, idesc2, accf);
                            umma_tf32_ss(d_addr, dAl + ko, dBh + ko, idesc, accf);
                        } else if (NACC == 2) {
                            umma_tf32_ss(d_addr, dAh + ko, dBh + ko, idesc2, accf);
                            umma_tf32_ss(d_addr + BN, dAl + ko, dBh + ko, idesc, 1u);
                        } else {
                            umma_tf32_ss(d_addr, dAh + ko, dBl + ko, idesc, accf);
                            umma_tf32_ss(d_addr, dAl + ko, dBh + ko, idesc, 1u);
                            umma_tf32_ss(d_addr, dAh + ko, dBh + ko, idesc, 1u);
                        }
                    }
                    umma_commit(&ab_free[s]);
                    if (last) umma_commit(&d_full[b]);
                }
                __syncwarp();
                if (last) ++grp;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int BN, bool VEC>
static int launch_wg_v(cudaStream_t st, const ConvKArgs& a, int mtiles, int ntiles, int ksplit, int swapped) {
    constexpr size_t smem = WgSmem<BN>::BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<BN, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_tc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        attr_set = true;
    }
    conv_tc_wgrad_kernel<BN, VEC><<<mtiles * ntiles * ksplit, TC_THREADS, smem, st>>>(a, mtiles, ntiles, ksplit, swapped);
    return 1;
}
template <int BN>
static int launch_wg_t(cudaStream_t st, const ConvKArgs& a, int mtiles, int ntiles, int ksplit, int swapped, bool vec) {
    return vec ? launch_wg_v<BN, true>(st, a, mtiles, ntiles, ksplit, swapped) : launch_wg_v<BN, false>(st, a, mtiles, ntiles, ksplit, swapped);
}

// Returns 1 when the tensor-core kernel was launched, 0 when the layer is not eligible, <0 on error.
int try_launch_wgrad_tc(cudaStream_t st, const ConvKArgs& a0) {
    const int mode = get_tc_mode();
    if (mode == 0) return 0;
    ConvKArgs a = a0;
    ConvGeom& g = a.g;
    const long long J = (long long)g.batch * g.OH * g.OW;
    static const bool dbg = getenv("B2S_WG_DEBUG") != nullptr;
#define WG_REJECT(why)                                                                                                   \
    do {                                                                                                                  \
        if (dbg) fprintf(stderr, "wgrad_tc: Cin %d Cout %d %dx%d k%d J %lld not eligible: %s\n", g.Cin, g.Cout, g.H, g.W, g.KH, J, why); \
        return 0;                                                                                                         \
    } while (0)
    // 128-bit loads need stride 1, rows that start 16-byte aligned and a horizontal shift of at most one pixel; every
    // other geometry (14 x 14 and 7 x 7 maps, strided or wide kernels) takes the scalar-gather variant of the transform
    static const int scalar_ok = getenv("B2S_WG_SCALAR") ? atoi(getenv("B2S_WG_SCALAR")) : 1;
    bool vec = !(g.sh != 1 || g.sw != 1 || (g.OW & 3) || (g.W & 3) || g.KW - 1 - g.pw > 1 || g.pw > 1) &&
               !((g.in_sstride & 3) || (g.out_sstride & 3));
    for (int p = 0; p < a.npairs; ++p)
        if (((uintptr_t)a.act[p] & 15) || ((uintptr_t)a.wt[p] & 15)) vec = false;
    if (!vec && !scalar_ok) WG_REJECT("stride / width / padding / alignment (scalar variant disabled)");
    if ((uintptr_t)a.out & 3) WG_REJECT("output alignment");
    if (J >= (1LL << 31) || (long long)g.Cin * g.H * g.W >= (1LL << 31) || (long long)g.Cout * g.OH * g.OW >= (1LL << 31)) WG_REJECT("index range");
    if (mode == 1 && !tc_worth_it(J, g.Cin, g.Cout, g.KH * g.KW)) WG_REJECT("too little work");
#undef WG_REJECT
    // The shifted (per-tap) operand is replicated KH*KW times along the M dimension: shift the one with fewer
    // channels.  For a "same" convolution the sum over output pixels of g[co,px] x[ci,px+s] equals the sum
    // over input pixels of x[ci,px'] g[co,px'-s]: exchange the operands, mirror the taps.
    int swapped = 0;
    if (g.Cout < g.Cin && g.H == g.OH && g.W == g.OW && 2 * g.ph == g.KH - 1 && 2 * g.pw == g.KW - 1) {
        swapped = 1;
        for (int p = 0; p < a.npairs; ++p) std::swap(a.act[p], a.wt[p]);
        std::swap(g.Cin, g.Cout);
        std::swap(g.in_sstride, g.out_sstride);
    }
    const int BN = tc_choose_bn(g.Cout);
    const int mtiles = (g.KH * g.KW * g.Cin + TC_M - 1) / TC_M;
    const int ntiles = (g.Cout + BN - 1) / BN;
    const int nkb = (int)((J + TC_KB - 1) / TC_KB);
    // CTAs of one launch: the weight gradient is a leaf on a side stream; the fewer SMs it holds at a time, the less the
    // latency-bound adjoint chain on the main stream waits for an SM (B2S_WG_MAXCTAS, default: all SMs)
    static const int max_ctas = getenv("B2S_WG_MAXCTAS") ? std::max(1, atoi(getenv("B2S_WG_MAXCTAS"))) : kNumSMs;
    int ksplit = std::max(1, max_ctas / (mtiles * ntiles));
    static const int min_kb = getenv("B2S_WG_MINKB") ? atoi(getenv("B2S_WG_MINKB")) : 8;
    ksplit = std::min(ksplit, std::max(1, nkb / min_kb));      // at least min_kb k-blocks per CTA and pair
    switch (BN) {
    case 16: return launch_wg_t<16>(st, a, mtiles, ntiles, ksplit, swapped, vec);
    case 32: return launch_wg_t<32>(st, a, mtiles, ntiles, ksplit, swapped, vec);
    case 48: return launch_wg_t<48>(st, a, mtiles, ntiles, ksplit, swapped, vec);
    case 64: return launch_wg_t<64>(st, a, mtiles, ntiles, ksplit, swapped, vec);
    case 96: return launch_wg_t<96>(st, a, mtiles, ntiles, ksplit, swapped, vec);
    default: return launch_wg_t<128>(st, a, mtiles, ntiles, ksplit, swapped, vec);
    }
}

}  // namespace b2s
