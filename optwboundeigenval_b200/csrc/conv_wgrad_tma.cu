// conv_wgrad_tma.cu -- TMA-staged tcgen05 weight-gradient contraction for the small-map layers (sm_100a):
// every convolution of DenseNet3 (32x32 / 16x16 / 8x8 maps) and of the USPS CNN.
//
//   Wbar[co][ci][ky][kx] += sum_p scale_p * sum_{n,oy,ox} g_p[n,co,oy,ox] * x_p[n,ci,oy+ky-ph,ox+kx-pw]
//
// Same GEMM view, 3xTF32 scheme, drain and epilogue as conv_tc_wgrad.cu (K = output pixel, A rows = (tap, channel) of
// the shifted tensor, B rows = channels of the other one).  What changes is how the operands reach the tensor core.
// In conv_tc_wgrad.cu every transform thread loads its share of the [128 + BN rows] x [32 pixels] stage from global
// memory -- the shifted operand once PER TAP -- and can only ask for the next k-block once it has stored the current
// one: a k-block costs the L2 latency plus the 9-fold re-read (measured ~1 us per k-block and CTA, 25-47 us per
// DenseNet3 launch whatever its size).  Here:
//
//   * a k-block is 32 consecutive output pixels = BH full image rows (W in {8, 16, 32});  ONE cp.async.bulk.tensor box
//     [W][BH rows][channels][1 image] per vertical tap lands the rows y0+dy .. of ALL channels of the shifted tensor in
//     shared memory (the vertical shift is a coordinate, top / bottom padding is TMA's zero fill), one more box the
//     other tensor: 3 x 1.5 KB + 6 KB per k-block for a DenseNet3 bottleneck 3x3 instead of 20 KB of per-thread loads,
//     issued up to NR k-blocks ahead by one producer thread;
//   * the three horizontal taps read the same box at a shared-memory offset of -1 / 0 / +1 pixel (aligned 128-bit
//     load + one scalar for the element that crosses the 16-byte boundary, left / right padding is a mask);
//   * two teams of 128 transform threads take alternate k-blocks: raw box -> registers -> hi / lo split -> the
//     canonical K-major SWIZZLE_128B operand layout (a 32-pixel row is exactly one 128-byte swizzle row, so the raw
//     reads -- 8 lanes = 128 contiguous bytes -- and the swizzled 128-bit stores are both bank-conflict free);
//   * SS-form tcgen05.mma from the swizzled stages, accumulators drained every WT_G k-blocks, fp32 atomics into the
//     flat gradient vector: unchanged.
//
// Eligible: stride 1, "same" size, W in {8, 16, 32} with H*W a multiple of 32, horizontal shift in {-1, 0, +1}, and either
// a 1x1 kernel or all taps of the shifted tensor in one 128-row tile (taps x channels <= 128; for a "same" convolution
// the tensor with fewer channels is the shifted one, as in conv_tc_wgrad.cu).  Everything else: conv_tc_wgrad.cu.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv_args.h"
#include "tc_common.cuh"
#include "tma_common.h"

namespace b2s {

constexpr int WT_THREADS = 512;
constexpr int WT_G = 4;                        // k-blocks accumulated in TMEM between drains (see conv_tc_wgrad.cu)
constexpr int WT_WARP_MMA = 12;
constexpr int WT_WARP_TMA = 13;
constexpr uint32_t WT_A_BYTES = TC_M * 128;    // 128 rows x 32 pixels of fp32: one of hi / lo, also the raw A area

// FLAT = false: k-block = BH full image rows, 4D boxes [W][BH][C][1] (W in {8, 16, 32}).
// FLAT = true : k-block = 32 consecutive pixels of the FLATTENED image plane (any W % 4 == 0: the 224 .. 28 wide maps of the
//               chest models), 3D boxes [40][C][1] starting 4 pixels early (16-byte aligned; the vertical tap is an offset
//               of dy * W pixels, planes are contiguous in NCHW, top / bottom padding is the zero fill outside the plane,
//               a ragged last k-block of an image is zero-filled as well); the horizontal taps read the box at +-1 pixel
//               and mask the pixels whose neighbour lies in the next image row.
constexpr int WT_FLAT_ROW = 40;                // floats per raw row of the shifted tensor in FLAT mode (4 + 32 + 4)
// TS = true (small maps, W in {16, 32}): the 128-row operand goes through TMEM (tcgen05.mma with A in TMEM, as conv_tma.cu)
//   instead of shared-memory hi / lo images: one thread per row reads the row's 32 pixels from the raw box -- which TMA
//   lands with a 128-byte (W = 32) or 64-byte (W = 16) swizzle, so that the 8 lanes of a quarter warp hit 8 (4) different
//   bank groups although their rows are 128 bytes apart --, shifts them in registers (the neighbours outside the k-block's
//   row are padding), splits and stores hi / lo with tcgen05.st.  Only the narrow operand keeps swizzled hi / lo images
//   in shared memory.  Per k-block the shared-memory traffic drops from 126 KB to ~68 KB (BN = 48).
template <int BN, bool FLAT, bool TS>
struct WtSmem {
    static constexpr int NST = BN >= 96 ? 2 : 3;                              // operand stages
    // raw stages: the TS form consumes a k-block in ~600 clocks, so a box (~1600 clocks from request to landing) has to be
    // requested more than three k-blocks ahead -- with four stages the timeline showed the teams waiting for TMA
    static constexpr int NR = FLAT ? 2 : TS ? (BN <= 48 ? 8 : 6) : (BN <= 48 ? 4 : 3);
    static constexpr uint32_t B_BYTES = BN * 128;
    static constexpr uint32_t B_OFF = TS ? 0u : 2 * WT_A_BYTES;               // B hi image inside a stage
    static constexpr uint32_t STAGE_BYTES = B_OFF + 2 * B_BYTES;              // [A hi | A lo |] B hi | B lo
    static constexpr uint32_t RAW_A_BYTES = FLAT ? TC_M * WT_FLAT_ROW * 4 : WT_A_BYTES;
    static constexpr uint32_t RAW_BYTES = RAW_A_BYTES + B_BYTES;              // raw A boxes | raw B box
    static constexpr size_t BYTES = (size_t)NST * STAGE_BYTES + (size_t)NR * RAW_BYTES + 256 + 1024;
    static_assert(BYTES <= 227 * 1024, "shared memory budget");
    static_assert(!(TS && FLAT) && !(TS && BN > 64), "TS form: small maps only");
};

struct alignas(64) WtMaps {
    CUtensorMap a[kMaxPairs];                  // shifted tensor of every pair
    CUtensorMap b[kMaxPairs];                  // the other tensor
};

struct WtGeom {
    int mtiles, ntiles, ksplit, swapped;
    int BH;                                    // image rows per k-block: W * BH = 32
    int ndy;                                   // vertical taps = boxes of the shifted tensor per k-block (1 for a 1x1 kernel)
    int boxCa;                                 // channels per box of the shifted tensor
    int a_box_al;                              // bytes between the boxes of consecutive vertical taps (1024-byte aligned)
    // FLAT mode: an m-tile = (vertical tap ky, channel group of G channels), rows = kx * G + channel
    int G, ncg;                                // channels per group, groups
    int kpi;                                   // k-blocks per image = ceil(H * W / 32)
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                    // leading byte offset: unused for swizzled K-major layouts
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// debug timeline (compiled in only with -DB2S_TC_TRACE_ENABLED): trace[(role * WT_TRACE_IT + i) * 4 + slot] = clock64() of CTA 0
constexpr int WT_TRACE_IT = 72;
__device__ __forceinline__ void wt_stamp(long long* trace, int role, int i, int slot) {
#ifdef B2S_TC_TRACE_ENABLED
    if (trace && blockIdx.x == 0 && i < WT_TRACE_IT) trace[((size_t)role * WT_TRACE_IT + i) * 4 + slot] = clock64();
#else
    (void)trace; (void)role; (void)i; (void)slot;
#endif
}

__device__ __forceinline__ float wt_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void wt_split_store(uint8_t* hi_img, uint8_t* lo_img, const uint32_t off, const float4 v) {
    float4 hi, lo;
    hi.x = wt_hi(v.x); lo.x = v.x - hi.x;
    hi.y = wt_hi(v.y); lo.y = v.y - hi.y;
    hi.z = wt_hi(v.z); lo.z = v.z - hi.z;
    hi.w = wt_hi(v.w); lo.w = v.w - hi.w;
    *reinterpret_cast<float4*>(hi_img + off) = hi;
    *reinterpret_cast<float4*>(lo_img + off) = lo;
}

template <int BN, bool FLAT, bool TS>
__global__ void __launch_bounds__(WT_THREADS, 1)
conv_wgrad_tma_kernel(const ConvKArgs a, const __grid_constant__ WtMaps maps, const WtGeom wg) {
    extern __shared__ uint8_t wt_smem_raw[];
    using S = WtSmem<BN, FLAT, TS>;
    constexpr int WT_NST = S::NST;
    uint8_t* smem = wt_smem_raw + ((1024u - (smem_u32(wt_smem_raw) & 1023u)) & 1023u);      // swizzle atoms: 1024-byte aligned
    uint8_t* stages = smem;
    uint8_t* raws = smem + (size_t)WT_NST * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raws + (size_t)S::NR * S::RAW_BYTES);
    uint64_t* raw_full = bars;                     // [NR]  TMA boxes landed (expect_tx)
    uint64_t* raw_free = raw_full + S::NR;         // [NR]  the 4 warps of the team that owns the k-block have read it (one arrival per warp)
    uint64_t* ab_full = raw_free + S::NR;          // [NST] the 4 warps of the owning team
    uint64_t* ab_free = ab_full + WT_NST;          // [NST] tcgen05.commit
    uint64_t* d_full = ab_free + WT_NST;           // [2]
    uint64_t* d_empty = d_full + 2;                // [2]   the 4 drain warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const ConvGeom& g = a.g;
    const int KHW = g.KH * g.KW;
    const int HW = g.H * g.W;
    const int W = g.W;
    const int nkb = FLAT ? g.batch * wg.kpi : (int)(((long long)g.batch * HW) / TC_KB);
    const int Mrows = KHW * g.Cin;
    const bool one = KHW == 1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mt = blockIdx.x % wg.mtiles;
    const int nt = (blockIdx.x / wg.mtiles) % wg.ntiles;
    const int ksl = blockIdx.x / (wg.mtiles * wg.ntiles);
    const int per = (nkb + wg.ksplit - 1) / wg.ksplit;
    const int kb0 = ksl * per;
    const int nloc = max(0, min(nkb, kb0 + per) - kb0);         // k-blocks of this CTA per pair
    const int total = a.npairs * nloc;
    // FLAT: tile mt = (ky, channel group cg); its rows are kx * G + (channel - cg * G)
    const int f_ky = FLAT ? mt / wg.ncg : 0, f_cg = FLAT ? mt - f_ky * wg.ncg : 0;
    const int f_ch = FLAT ? min(wg.G, g.Cin - f_cg * wg.G) : 0;                    // channels of this group
    const int rows_here = FLAT ? (g.KW - 1) * wg.G + f_ch : min(TC_M, Mrows - mt * TC_M);   // A rows of this tile (upper bound)

    if (tid == 0) wt_stamp(a.trace, 3, WT_TRACE_IT - 1, 2);
    // the producer thread initialises its own barriers and requests the first NR k-blocks at once: their flight overlaps
    // the rest of the prologue (TMEM allocation, row table, zero fill) -- the timeline showed the first box landing
    // ~1700 clocks after a request that was only issued behind the block-wide barrier
    const uint32_t a_box = (uint32_t)wg.boxCa * 128u, b_box = (uint32_t)BN * 128u;
    auto request = [&](const int i) {
        const int p = i / nloc;
        const int kb = kb0 + (i - p * nloc);
        const int rs = i % S::NR;
        uint8_t* dst = raws + (size_t)rs * S::RAW_BYTES;
        if (FLAT) {
            const int n = kb / wg.kpi;
            const int j0 = (kb - n * wg.kpi) * TC_KB;
            mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)wg.G * (WT_FLAT_ROW * 4u) + b_box);
            tma_load_3d(dst, &maps.a[p], j0 - 4 + (f_ky - g.ph) * W, f_cg * wg.G, n, &raw_full[rs]);
            tma_load_3d(dst + S::RAW_A_BYTES, &maps.b[p], j0, nt * BN, n, &raw_full[rs]);
        } else {
            const int j0 = kb * TC_KB;
            const int n = j0 / HW;
            const int y0 = (j0 - n * HW) / W;
            mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)wg.ndy * a_box + b_box);
            for (int d = 0; d < wg.ndy; ++d)
                tma_load_4d(dst + (size_t)d * wg.a_box_al, &maps.a[p], 0, one ? y0 : y0 + d - g.ph, one ? mt * TC_M : 0, n, &raw_full[rs]);
            tma_load_4d(dst + S::RAW_A_BYTES, &maps.b[p], 0, y0, nt * BN, n, &raw_full[rs]);
        }
    };
    const int prefill = min(S::NR < 4 ? S::NR : 4, total);        // the rest of a deeper ring is requested behind the barrier (32 TMA issues before it delayed every role by ~3000 clocks)
    if (warp == WT_WARP_TMA && lane == 0) {
        for (int s = 0; s < S::NR; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_free[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < prefill; ++i) request(i);
    }
    if (tid == 0) {
        for (int s = 0; s < WT_NST; ++s) {
            mbar_init(&ab_full[s], 4);
            mbar_init(&ab_free[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WT_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // per-row constants of the A tile, decoded once per row (the divisions cost ~2400 clocks when every transform thread
    // did them for its eight rows): (raw float offset of the row's channel inside the boxes) * 4 + (dx + 1), -1 = no row
    __shared__ int row_tab[TC_M];
    if (tid < TC_M) {
        int e = -1;
        if (FLAT) {
            const int kx = tid / wg.G, cl = tid - kx * wg.G;
            if (kx < g.KW && cl < f_ch) e = (cl * WT_FLAT_ROW + 4) * 4 + (kx - g.pw + 1);
        } else {
            const int m = mt * TC_M + tid;
            if (m < Mrows) {
                if (one) e = (tid * 32) * 4 + 1;
                else {
                    const int t = m / g.Cin, ca = m - t * g.Cin;
                    const int ky = t / g.KW, kx = t - ky * g.KW;
                    e = (ky * (wg.a_box_al >> 2) + ca * 32) * 4 + (kx - g.pw + 1);
                }
            }
        }
        row_tab[tid] = e;
    }
    if (!TS && rows_here < TC_M) {   // A rows past the last valid one are never written: zero the stages once
        float4* z = reinterpret_cast<float4*>(stages);
        const int n4 = WT_NST * (int)S::STAGE_BYTES / 16;
        for (int i = tid; i < n4; i += WT_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (TS && warp < 8) {
        // ===================== transform, TS form: thread = row of the 128-row operand -> TMEM; narrow operand -> smem =====
        const int team = warp >> 2;
        const int r = tid & 127;
        const int kg = r & 7, rowsub = r >> 3;                  // narrow operand: 4-pixel group kg of rows rowsub + 16 u
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int e = row_tab[r];
        const bool row_ok = e >= 0;
        const int adx = row_ok ? (e & 3) - 1 : 0;
        const uint32_t row_off = row_ok ? (uint32_t)(e >> 2) * 4u : 0u;           // byte offset of this row inside the raw A area
        // physical 16-byte chunk of logical chunk j: TMA's swizzle XORs the chunk index with the 128-byte row index
        const uint32_t swz = W == 32 ? ((row_off >> 7) & 7u) : ((row_off >> 7) & 3u);
        const bool warp_ok = __any_sync(0xffffffffu, row_ok);
        uint32_t doff[BN / 16];
#pragma unroll
        for (int u = 0; u < BN / 16; ++u) {
            const int rl = rowsub + 16 * u;
            doff[u] = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128 + ((kg ^ (rl & 7)) << 4));
        }
        constexpr int NB = BN / 16;
        for (int i = team; i < total; i += 2) {
            const int rs = i % S::NR;
            const int s = i % WT_NST;
            const uint32_t round = i / WT_NST;
            const uint8_t* rawA = raws + (size_t)rs * S::RAW_BYTES;
            const float* __restrict__ rawB = reinterpret_cast<const float*>(rawA + S::RAW_A_BYTES);
            if (r == 0) wt_stamp(a.trace, team, i, 0);
            mbar_wait(&raw_full[rs], (i / S::NR) & 1);
            if (r == 0) wt_stamp(a.trace, team, i, 1);
            float v[34];                                        // v[1 + k] = pixel k of this row; v[0], v[33]: outside the k-block = padding
            v[0] = 0.f; v[33] = 0.f;
            if (warp_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t ph = W == 32 ? (((uint32_t)j ^ swz) << 4) : ((uint32_t)(j >> 2) * 64u + ((((uint32_t)j & 3u) ^ swz) << 4));
                    const float4 c4 = row_ok ? *reinterpret_cast<const float4*>(rawA + row_off + ph) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[1 + 4 * j] = c4.x; v[2 + 4 * j] = c4.y; v[3 + 4 * j] = c4.z; v[4 + 4 * j] = c4.w;
                }
            }
            float4 cb[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) cb[u] = *reinterpret_cast<const float4*>(rawB + (rowsub + 16 * u) * 32 + kg * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_free[rs]);          // the boxes are in registers: the producer may refill the stage
            if (round > 0) mbar_wait(&ab_free[s], (round - 1) & 1);
            __syncwarp();
            tc_fence_after();
            if (r == 0) wt_stamp(a.trace, team, i, 2);
            if (warp_ok) {
                const uint32_t col = (uint32_t)(s * TC_ACOLS);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int t8 = 0; t8 < 8; ++t8) {
                        const int k = j8 * 8 + t8;
                        // horizontal tap: the neighbour inside the image row, zero across a row end (W = 16: two rows per k-block)
                        float f = v[1 + k];
                        if (adx < 0) f = (W == 16 && k == 16) ? 0.f : v[k];
                        else if (adx > 0) f = (W == 16 && k == 15) ? 0.f : v[2 + k];
                        hi[t8] = (__float_as_uint(f) + 0x1000u) & 0xffffe000u;
                        lo[t8] = __float_as_uint(f - __uint_as_float(hi[t8]));
                    }
                    tmem_st8(lane_addr + col + j8 * 8, hi);
                    tmem_st8(lane_addr + col + 32 + j8 * 8, lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            uint8_t* Bhi = stages + (size_t)s * S::STAGE_BYTES + S::B_OFF;
            uint8_t* Blo = Bhi + S::B_BYTES;
#pragma unroll
            for (int u = 0; u < NB; ++u) wt_split_store(Bhi, Blo, doff[u], cb[u]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ab_full[s]);
            if (r == 0) wt_stamp(a.trace, team, i, 3);
        }
    } else if (warp < 8) {
        // ===================== transform: raw boxes -> registers (shift, split) -> swizzled operand stage =========
        const int team = warp >> 2;
        const int r = tid & 127;
        const int kg = r & 7;                                   // 4-pixel group inside the k-block
        const int rowsub = r >> 3;                              // rows rowsub + 16 q
        const int x0 = (kg * 4) % W;
        // per-row constants: raw offset (floats), horizontal shift, validity; swizzled destination offset (bytes)
        int aoff[8], adx[8];
        uint32_t doff[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int rl = rowsub + 16 * q;
            doff[q] = (uint32_t)((rl >> 3) * 1024 + (rl & 7) * 128 + ((kg ^ (rl & 7)) << 4));
            const int e = row_tab[rl];
            aoff[q] = e < 0 ? -1 : (e >> 2) + kg * 4;
            adx[q] = e < 0 ? 0 : (e & 3) - 1;
        }
        bool has_l = x0 > 0, has_r = x0 + 4 < W;
        constexpr int NB = BN / 16;
        for (int i = team; i < total; i += 2) {
            const int rs = i % S::NR;
            const int s = i % WT_NST;
            const uint32_t round = i / WT_NST;
            const float* __restrict__ rawA = reinterpret_cast<const float*>(raws + (size_t)rs * S::RAW_BYTES);
            const float* __restrict__ rawB = rawA + S::RAW_A_BYTES / 4;
            if (FLAT) {                                         // column of this lane's 4-pixel group inside its image row
                const int kb = kb0 + (i % nloc);
                const int xx = ((kb % wg.kpi) * TC_KB + kg * 4) % W;
                has_l = xx > 0;
                has_r = xx + 4 < W;
            }
            if (r == 0) wt_stamp(a.trace, team, i, 0);
            mbar_wait(&raw_full[rs], (i / S::NR) & 1);
            if (r == 0) wt_stamp(a.trace, team, i, 1);
            float4 ca[8];
            float4 cb[NB];
            float el[8], er[8];
            // every load first (branch-free: a warp holds rows of different taps), then the shift as selects
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (16 * q >= rows_here) continue;              // block-uniform
                const bool ok = aoff[q] >= 0;
                const float* __restrict__ src = rawA + (ok ? aoff[q] : 0);
                ca[q] = ok ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                el[q] = (ok && adx[q] < 0 && has_l) ? src[-1] : 0.f;
                er[q] = (ok && adx[q] > 0 && has_r) ? src[4] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) cb[u] = *reinterpret_cast<const float4*>(rawB + (rowsub + 16 * u) * 32 + kg * 4);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (16 * q >= rows_here) continue;
                const float4 v = ca[q];
                if (adx[q] < 0) ca[q] = make_float4(el[q], v.x, v.y, v.z);
                else if (adx[q] > 0) ca[q] = make_float4(v.y, v.z, v.w, er[q]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_free[rs]);          // the boxes are in registers: the producer may refill the stage
            if (round > 0) mbar_wait(&ab_free[s], (round - 1) & 1);
            __syncwarp();
            if (r == 0) wt_stamp(a.trace, team, i, 2);
            uint8_t* Ahi = stages + (size_t)s * S::STAGE_BYTES;
            uint8_t* Alo = Ahi + WT_A_BYTES;
            uint8_t* Bhi = Ahi + S::B_OFF;
            uint8_t* Blo = Bhi + S::B_BYTES;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (16 * q >= rows_here) continue;
                wt_split_store(Ahi, Alo, doff[q], ca[q]);
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) wt_split_store(Bhi, Blo, doff[u], cb[u]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&ab_full[s]);
            if (r == 0) wt_stamp(a.trace, team, i, 3);
        }
    } else if (warp < 12) {
        // ===================== drain + epilogue (as conv_tc_wgrad.cu) =================================
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
            for (int k0 = 0; k0 < nloc; k0 += WT_G, ++grp) {
                const int b = grp & 1;
                if (r == 0) wt_stamp(a.trace, 3, (int)grp, 0);
                mbar_wait(&d_full[b], (grp >> 1) & 1);
                __syncwarp();
                tc_fence_after();
                if (r == 0) wt_stamp(a.trace, 3, (int)grp, 1);
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
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[b]);
                if (r == 0) wt_stamp(a.trace, 3, (int)grp, 2);
            }
        }
        if (r == 0) wt_stamp(a.trace, 3, WT_TRACE_IT - 1, 0);
        int t, ci;
        bool row_ok;
        if (FLAT) {
            const int kx = r / wg.G, cl = r - kx * wg.G;
            row_ok = kx < g.KW && cl < f_ch;
            t = f_ky * g.KW + kx;
            ci = f_cg * wg.G + cl;
        } else {
            const int m = mt * TC_M + r;
            row_ok = m < Mrows;
            t = m / g.Cin;
            ci = m - t * g.Cin;
        }
        if (total > 0 && row_ok) {
            if (!wg.swapped) {
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
        if (warp == WT_WARP_MMA) {
            // ===================== MMA issuer ======================================================
            const uint32_t idesc = umma_idesc_tf32(TC_M, BN);
            const uint32_t idesc2 = umma_idesc_tf32(TC_M, 2 * BN <= 256 ? 2 * BN : BN);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(stages));            // stage 0, A hi, k-step 0
            constexpr int NACC = tc_nacc(BN) > 3 ? 3 : tc_nacc(BN);
            constexpr uint32_t STAGE16 = S::STAGE_BYTES >> 4, A16 = WT_A_BYTES >> 4, B16 = S::B_BYTES >> 4;
            uint32_t grp = 0;
            for (int i = 0; i < total; ++i) {
                const int s = i % WT_NST;
                const uint32_t round = i / WT_NST;
                const int kk = i % nloc;                              // k-block inside its pair: drain groups never span pairs
                const bool first = (kk % WT_G) == 0;
                const bool last = (kk % WT_G) == WT_G - 1 || kk == nloc - 1;
                const int b = grp & 1;
                const uint32_t use = grp >> 1;
                if (lane == 0) wt_stamp(a.trace, 2, i, 0);
                mbar_wait(&ab_full[s], round & 1);
                if (lane == 0) wt_stamp(a.trace, 2, i, 1);
                if (first && use > 0) mbar_wait(&d_empty[b], (use - 1) & 1);
                __syncwarp();
                tc_fence_after();
                if (lane == 0) wt_stamp(a.trace, 2, i, 2);
                const uint32_t d_addr = tmem_u + TC_DCOL0 + b * TC_DCOLS;
                const uint64_t dAh = desc0 + (uint64_t)(s * STAGE16), dAl = dAh + A16;
                const uint64_t dBh = desc0 + (uint64_t)(s * STAGE16) + (uint64_t)(S::B_OFF >> 4), dBl = dBh + B16;
                const uint32_t a_hi = tmem_u + (uint32_t)(s * TC_ACOLS), a_lo = a_hi + 32;      // TS form: A stage in TMEM
                const uint32_t cont = first ? 0u : 1u;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < TC_KB / 8; ++ks) {
                        const uint64_t ko = (uint64_t)(ks * 2);           // 32 bytes per k-step inside the 128-byte swizzle row
                        const uint32_t accf = ks >= 1 ? 1u : cont;
                        if (TS) {
                            if (NACC == 3) {
                                umma_tf32_ts(d_addr + BN, a_hi + ks * 8, dBh + ko, idesc2, accf);     // [hi*hi | hi*lo]
                                umma_tf32_ts(d_addr, a_lo + ks * 8, dBh + ko, idesc, accf);           // lo*hi
                            } else if (NACC == 2) {
                                umma_tf32_ts(d_addr, a_hi + ks * 8, dBh + ko, idesc2, accf);
                                umma_tf32_ts(d_addr + BN, a_lo + ks * 8, dBh + ko, idesc, 1u);
                            } else {
                                umma_tf32_ts(d_addr, a_hi + ks * 8, dBl + ko, idesc, accf);
                                umma_tf32_ts(d_addr, a_lo + ks * 8, dBh + ko, idesc, 1u);
                                umma_tf32_ts(d_addr, a_hi + ks * 8, dBh + ko, idesc, 1u);
                            }
                        } else if (NACC == 3) {
                            umma_tf32_ss(d_addr + BN, dAh + ko, dBh + ko, idesc2, accf);      // [hi*hi | hi*lo]
                            umma_tf32_ss(d_addr, dAl + ko, dBh + ko, idesc, accf);            // lo*hi
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
                if (lane == 0) wt_stamp(a.trace, 2, i, 3);
                if (last) ++grp;
            }
        } else if (warp == WT_WARP_TMA) {
            // ===================== producer: the boxes of a k-block, NR k-blocks ahead ===============
            for (int i = prefill; i < total; ++i) {
                if (lane == 0) {
                    if (i >= S::NR) mbar_wait(&raw_free[i % S::NR], ((uint32_t)(i / S::NR) - 1) & 1);
                    request(i);
                }
                __syncwarp();
            }
        }
    }

    if (tid == 256) wt_stamp(a.trace, 3, WT_TRACE_IT - 1, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == WT_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int BN, bool FLAT, bool TS>
static int launch_wt_t(cudaStream_t st, const ConvKArgs& a, const WtMaps& maps, const WtGeom& wg) {
    constexpr size_t smem = WtSmem<BN, FLAT, TS>::BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tma_kernel<BN, FLAT, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_wgrad_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        attr_set = true;
    }
    static const bool want_trace = getenv("B2S_WG_TRACE") != nullptr;
    static int traced = 0;
    if (want_trace && traced < 6 && a.g.KH == (getenv("B2S_WG_TRACE_K") ? atoi(getenv("B2S_WG_TRACE_K")) : 3) &&
        (!getenv("B2S_WG_TRACE_W") || a.g.W == atoi(getenv("B2S_WG_TRACE_W")))) {   // debug: synchronous launch + timeline of CTA 0
        ++traced;
        long long* d_tr = nullptr;
        const size_t n = 4 * WT_TRACE_IT * 4;
        cudaMalloc(&d_tr, n * sizeof(long long));
        cudaMemset(d_tr, 0, n * sizeof(long long));
        ConvKArgs b = a;
        b.trace = d_tr;
        conv_wgrad_tma_kernel<BN, FLAT, TS><<<wg.mtiles * wg.ntiles * wg.ksplit, WT_THREADS, smem, st>>>(b, maps, wg);
        cudaStreamSynchronize(st);
        std::vector<long long> h(n);
        cudaMemcpy(h.data(), d_tr, n * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(d_tr);
        auto raw = [&](int role, int i, int slot) { return h[((size_t)role * WT_TRACE_IT + i) * 4 + slot]; };
        const long long t0 = raw(3, WT_TRACE_IT - 1, 2);
        auto at = [&](int role, int i, int slot) { const long long v = raw(role, i, slot); return v ? v - t0 : -1; };
        fprintf(stderr, "WGT trace BN=%d flat=%d ts=%d Cin=%d Cout=%d k=%d W=%d grid=%d (clocks since kernel start); drain loop end %lld, epilogue end %lld\n", BN,
                (int)FLAT, (int)TS, a.g.Cin, a.g.Cout, a.g.KH, a.g.W, wg.mtiles * wg.ntiles * wg.ksplit, at(3, WT_TRACE_IT - 1, 0), at(3, WT_TRACE_IT - 1, 1));
        for (int i = 0; i < WT_TRACE_IT - 1; ++i) {
            const int team = i & 1;
            if (at(2, i, 0) < 0) break;
            fprintf(stderr, "  it %2d  X%d: poll %6lld raw_ok %6lld ab_free %6lld stored %6lld | MMA: poll %6lld ab_ok %6lld d_ok %6lld issued %6lld | D[%d]: poll %6lld full %6lld done %6lld\n",
                    i, team, at(team, i, 0), at(team, i, 1), at(team, i, 2), at(team, i, 3), at(2, i, 0), at(2, i, 1), at(2, i, 2), at(2, i, 3),
                    i / WT_G, at(3, i / WT_G, 0), at(3, i / WT_G, 1), at(3, i / WT_G, 2));
        }
        return 1;
    }
    conv_wgrad_tma_kernel<BN, FLAT, TS><<<wg.mtiles * wg.ntiles * wg.ksplit, WT_THREADS, smem, st>>>(a, maps, wg);
    return 1;
}

// Returns 1 when the kernel was launched, 0 when the layer is not eligible, <0 on error.
int try_launch_wgrad_tma(cudaStream_t st, const ConvKArgs& a0) {
    static const int enabled = getenv("B2S_WGRAD_TMA") ? atoi(getenv("B2S_WGRAD_TMA")) : 1;     // 0 = off, 1 = both forms, 2 = small maps only
    const int mode = get_tc_mode();
    if (!enabled || mode == 0) return 0;
    ConvKArgs a = a0;
    ConvGeom& g = a.g;
    const long long J = (long long)g.batch * g.OH * g.OW;
    if (g.sh != 1 || g.sw != 1 || g.H != g.OH || g.W != g.OW) return 0;
    if (g.pw > 1 || g.KW - 1 - g.pw > 1 || J >= (1LL << 31)) return 0;
    if ((g.in_sstride & 3) || (g.out_sstride & 3) || ((uintptr_t)a.out & 3)) return 0;
    for (int p = 0; p < a.npairs; ++p)
        if (((uintptr_t)a.act[p] & 15) || ((uintptr_t)a.wt[p] & 15)) return 0;
    if (mode == 1 && !tc_worth_it(J, g.Cin, g.Cout, g.KH * g.KW)) return 0;
    const int KHW = g.KH * g.KW;
    const int BH = (g.W == 8 || g.W == 16 || g.W == 32) ? TC_KB / g.W : 0;
    bool flat = !(BH > 0 && g.H % BH == 0);
    // which tensor is shifted (A rows = taps x its channels): for a "same" convolution the sum over output pixels of
    // g[co,px] x[ci,px+s] equals the sum over input pixels of x[ci,px'] g[co,px'-s] -- exchange the operands, mirror the
    // taps.  k x k: the tensor with fewer channels; 1x1: the one with MORE channels supplies the rows, the other one the
    // (narrower) columns.
    const bool same = 2 * g.ph == g.KH - 1 && 2 * g.pw == g.KW - 1;
    int swapped = 0;
    if (same && (KHW > 1 ? g.Cout < g.Cin : g.Cout > g.Cin)) {
        swapped = 1;
        for (int p = 0; p < a.npairs; ++p) std::swap(a.act[p], a.wt[p]);
        std::swap(g.Cin, g.Cout);
        std::swap(g.in_sstride, g.out_sstride);
    }
    if (!flat && KHW > 1 && KHW * g.Cin > TC_M) flat = true;           // the taps of the shifted tensor do not fit one tile
    const int BN = tc_choose_bn(g.Cout);
    if (!flat && BN > 64) flat = true;
    if (flat && (enabled == 2 || (g.W & 3) || (KHW > 1 && !same))) return 0;
    // a handful of shifted channels (VGG16 conv1_1: 3): a (vertical tap, channel group) tile would hold 9 of 128 rows --
    // the gather kernel packs all 27 (tap, channel) rows into one tile (61 against 104 us)
    if (flat && KHW > 1 && KHW * g.Cin <= 64) return 0;
    TmEncodeFn enc = tm_encoder();
    if (!enc) return 0;
    WtGeom wg{};
    wg.swapped = swapped;
    static const int ts_on = getenv("B2S_WGRAD_TS") ? atoi(getenv("B2S_WGRAD_TS")) : 1;
    const bool ts = ts_on && !flat && (g.W == 32 || g.W == 16) && BN <= 64;
    WtMaps maps;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    int nkb;
    if (!flat) {
        wg.BH = BH;
        wg.ndy = KHW == 1 ? 1 : g.KH;
        wg.boxCa = KHW == 1 ? std::min(TC_M, (g.Cin + 7) & ~7) : g.Cin;
        wg.a_box_al = (wg.boxCa * 128 + 1023) & ~1023;
        if ((long long)wg.ndy * wg.a_box_al > (long long)WT_A_BYTES) return 0;
        wg.mtiles = (KHW * g.Cin + TC_M - 1) / TC_M;
        nkb = (int)(J / TC_KB);
        for (int p = 0; p < a.npairs; ++p) {
            for (int which = 0; which < 2; ++which) {
                const float* base = which == 0 ? a.act[p] : a.wt[p];
                const int C = which == 0 ? g.Cin : g.Cout;
                const long long ss = which == 0 ? g.in_sstride : g.out_sstride;
                const cuuint64_t dims[4] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)C, (cuuint64_t)g.batch};
                const cuuint64_t strides[3] = {(cuuint64_t)g.W * 4, (cuuint64_t)g.H * g.W * 4, (cuuint64_t)ss * 4};
                const cuuint32_t box[4] = {(cuuint32_t)g.W, (cuuint32_t)BH, (cuuint32_t)(which == 0 ? wg.boxCa : BN), 1};
                // TS form: the 128-row operand's box is landed swizzled so that one thread per row reads it without bank conflicts
                const CUtensorMapSwizzle swz = (ts && which == 0) ? (g.W == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)
                                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
                const CUresult r = enc(which == 0 ? &maps.a[p] : &maps.b[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims,
                                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return 0;
            }
        }
    } else {
        // m-tile = (vertical tap, channel group): its KW horizontal taps share ONE box of the group's channels
        const int gmax = TC_M / g.KW;                                  // 42 channels for a 3 x 3 kernel, 128 for 1 x 1
        wg.ncg = (g.Cin + gmax - 1) / gmax;
        wg.G = (g.Cin + wg.ncg - 1) / wg.ncg;
        wg.mtiles = g.KH * wg.ncg;
        wg.kpi = (g.H * g.W + TC_KB - 1) / TC_KB;
        nkb = g.batch * wg.kpi;
        const long long HW = (long long)g.H * g.W;
        for (int p = 0; p < a.npairs; ++p) {
            for (int which = 0; which < 2; ++which) {
                const float* base = which == 0 ? a.act[p] : a.wt[p];
                const int C = which == 0 ? g.Cin : g.Cout;
                const long long ss = which == 0 ? g.in_sstride : g.out_sstride;
                const cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)g.batch};
                const cuuint64_t strides[2] = {(cuuint64_t)HW * 4, (cuuint64_t)ss * 4};
                const cuuint32_t box[3] = {(cuuint32_t)(which == 0 ? WT_FLAT_ROW : TC_KB), (cuuint32_t)(which == 0 ? wg.G : BN), 1};
                const CUresult r = enc(which == 0 ? &maps.a[p] : &maps.b[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims,
                                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return 0;
            }
        }
    }
    wg.ntiles = (g.Cout + BN - 1) / BN;
    static const int min_kb = getenv("B2S_WGT_MINKB") ? std::max(1, atoi(getenv("B2S_WGT_MINKB"))) : 16;
    static const int max_ctas = getenv("B2S_WG_MAXCTAS") ? std::max(1, atoi(getenv("B2S_WG_MAXCTAS"))) : kNumSMs;
    // split-K: the number of slices that minimises (waves of CTAs) x (k-blocks per CTA + ~4 k-blocks' worth of prologue and
    // epilogue), at least min_kb k-blocks per slice -- 156 tiles (VGG16 conv4_x) would otherwise run as one full wave plus
    // a second wave of 8 CTAs
    const int tiles = wg.mtiles * wg.ntiles;
    const int ks_max = std::max(1, nkb / min_kb);
    int ksplit = 1;
    double best = 1e30;
    for (int ks = 1; ks <= ks_max && ks <= 4 * max_ctas; ++ks) {
        const int waves = (tiles * ks + max_ctas - 1) / max_ctas;
        const double cost = waves * ((double)((nkb + ks - 1) / ks) + 4.0);
        if (cost < best - 1e-9) { best = cost; ksplit = ks; }
    }
    wg.ksplit = ksplit;
    for (int p = a.npairs; p < kMaxPairs; ++p) { maps.a[p] = maps.a[0]; maps.b[p] = maps.b[0]; }
    if (flat) {
        switch (BN) {
        case 16: return launch_wt_t<16, true, false>(st, a, maps, wg);
        case 32: return launch_wt_t<32, true, false>(st, a, maps, wg);
        case 48: return launch_wt_t<48, true, false>(st, a, maps, wg);
        case 64: return launch_wt_t<64, true, false>(st, a, maps, wg);
        case 96: return launch_wt_t<96, true, false>(st, a, maps, wg);
        default: return launch_wt_t<128, true, false>(st, a, maps, wg);
        }
    }
    if (ts) {
        switch (BN) {
        case 16: return launch_wt_t<16, false, true>(st, a, maps, wg);
        case 32: return launch_wt_t<32, false, true>(st, a, maps, wg);
        case 48: return launch_wt_t<48, false, true>(st, a, maps, wg);
        default: return launch_wt_t<64, false, true>(st, a, maps, wg);
        }
    }
    switch (BN) {
    case 16: return launch_wt_t<16, false, false>(st, a, maps, wg);
    case 32: return launch_wt_t<32, false, false>(st, a, maps, wg);
    case 48: return launch_wt_t<48, false, false>(st, a, maps, wg);
    default: return launch_wt_t<64, false, false>(st, a, maps, wg);
    }
}

}  // namespace b2s
