// conv_tma.cu -- TMA-fed tcgen05 / TMEM implicit GEMM for the stride-1 "same"-width Conv2d jet contractions
// (sm_100a): every DenseNet3 layer.
//
// Same contraction, same fp32-accurate 3xTF32 scheme, same MMA / drain / epilogue roles as conv_tc.cu
// (A operand in TMEM, weights as packed K-major shared-memory images, a fresh TMEM accumulator per k-block
// drained into fp32 registers).  What changes is how the activation operand gets there.  In conv_tc.cu
// every transform thread computes im2col addresses and issues 32 global loads per k-block -- the profile
// showed the kernel bound by exactly that instruction stream (~1150 clocks per k-block, two thirds of
// it address arithmetic, load issue and L2 latency).  Here:
//
//   * the tile is 128 consecutive pixels = BH full image rows (of BI images when an image has < 128 pixels);
//     ONE cp.async.bulk.tensor box  [W][BH rows][KC channels][BI images]  per (pair, vertical tap, channel
//     chunk) lands the rows y0+dy .. y0+dy+BH-1 of 32 channels in shared memory: the vertical shift is a
//     coordinate, the top / bottom zero padding is TMA's out-of-bounds fill, channels past the end of the
//     tensor are zero-filled too;
//   * the box is reused by the KW horizontal taps: the shift dx is an offset of the shared-memory read
//     (bank-conflict free: consecutive lanes = consecutive pixels), left / right padding is a per-thread mask.
//     (TMA cannot do this shift itself: a box must start 16-byte aligned in the contiguous dimension --
//     measured: a start coordinate of -1 float raises an illegal-instruction fault.)
//   * two teams of 128 transform threads (thread = pixel) take alternate k-blocks: 32 conflict-free LDS at
//     immediate offsets, split into hi = rna_tf32(x) (integer add + mask), lo = x - hi (exact; the tensor core
//     ignores the low 13 bits of a TF32 operand), tcgen05.st into the A stage in TMEM;
//   * the timeline of the first version showed the SM issue-bound (transform ~130 and drain ~250 instructions
//     per thread and k-block), so the accumulator is drained every TM_G k-blocks instead of every k-block:
//     TM_G * 4 truncating accumulations of K = 8 (<= 2^-24 relative bias each) stay far below the 1e-4 budget.
//
// Warp roles (512 threads): warps 0-7 transform, 8-11 drain + epilogue (208 registers), 12 MMA issuer,
// 13 weight producer (bulk copies), 14 activation producer (TMA).  k-block order: pair, ky, chunk, kx.
#include <cuda.h>            // CUtensorMap and its enums; the encoder is resolved at run time (no -lcuda)

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "conv_args.h"
#include "tc_common.cuh"
#include "tma_common.h"

namespace b2s {

// Transform teams of 128 threads.  The per-role timeline of the TS-form weight gradient (conv_wgrad_tma.cu) suggested that the
// split -> tcgen05.st -> wait::st -> fence -> arrive phase (~1000 clocks per team and k-block, one warp per scheduler working
// through ~260 dependent instructions) paces this kernel: two teams give one k-block per ~455 clocks.  Experiment
// (B2S_BUILD_TM_TEAMS=3): narrow tiles (BN <= 48) with THREE teams -- 640 threads at 96 registers, each team owning one of
// the three A stages.  Measured: DenseNet3 HVP 2.273 -> 2.262 ms, i.e. nothing -- the k-loop is not paced by the transform
// alone -- so the default stays at two teams.  What the experiment did teach: setmaxnreg.inc can only take what
// setmaxnreg.dec of the SAME CTA has released (the pool is the CTA's launch allocation, not the SM's register file): with
// 640 threads at 96 registers the four misc warps release 4 x 32 x (96 - 40), so the drain warps can go to 152 at most;
// asking for 168 blocked forever.
#ifndef B2S_TM_TEAMS
#define B2S_TM_TEAMS 2
#endif
__host__ __device__ constexpr int tm_teams(int BN) { return (B2S_TM_TEAMS == 3 && BN <= 48) ? 3 : 2; }
__host__ __device__ constexpr int tm_threads(int BN) { return (tm_teams(BN) * 4 + 8) * 32; }
__host__ __device__ constexpr int tm_regs_drain(int BN) { return tm_teams(BN) == 3 ? 152 : TC_REGS_DRAIN; }
constexpr int TM_NR = 4;                               // raw activation stages (TMA boxes of <= 16 KB)
constexpr int TM_RAW_FLOATS = TC_M * TC_KB;            // floats reserved per raw stage
#ifndef B2S_TM_G
#define B2S_TM_G 4
#endif
constexpr int TM_G = B2S_TM_G;                         // k-blocks accumulated in TMEM between drains

struct alignas(64) TmaMaps { CUtensorMap m[kMaxPairs]; };

// pixel tile of the destination: BH full rows of BI consecutive images, W * BH * BI = 128
struct TmaTiling {
    int BH, BI;
    int KC;                  // channels per box (8..32, multiple of 8)
    int dbg;                 // B2S_TMA_DBG: 4 = soft barrier time-outs (flag instead of trap) + synchronous launch report
};

// debug (B2S_TMA_DBG & 4): the first barrier wait that times out records its id here and every later wait
// returns at once, so that a protocol error ends the kernel with a diagnosis instead of a trap
__device__ unsigned int tm_timeout_flag = 0;
__device__ __forceinline__ void tm_wait(uint64_t* bar, uint32_t parity, int dbg, unsigned id) {
    if (dbg & 8) { mbar_wait_spin(bar, parity); return; }          // experiment: spin instead of NANOSLEEP.SYNCS
    if (!(dbg & 4)) { mbar_wait(bar, parity); return; }
    const uint32_t addr = smem_u32(bar);
    if (*((volatile unsigned int*)&tm_timeout_flag)) return;
    const long long t0 = clock64();
    while (!mbar_try(addr, parity)) {
        if (*((volatile unsigned int*)&tm_timeout_flag)) return;
        if (clock64() - t0 > 20000000LL) { atomicCAS(&tm_timeout_flag, 0u, id); return; }
    }
}

// debug timeline (compiled in only with -DB2S_TC_TRACE_ENABLED): trace[(role * TM_TRACE_KB + kbg) * 4 + slot] = clock64()
constexpr int TM_TRACE_KB = 80;
__device__ __forceinline__ void tm_stamp(long long* trace, int role, uint32_t kbg, int slot) {
#ifdef B2S_TC_TRACE_ENABLED
    if (trace && blockIdx.x == 0 && kbg < TM_TRACE_KB) trace[((size_t)role * TM_TRACE_KB + kbg) * 4 + slot] = clock64();
#else
    (void)trace; (void)role; (void)kbg; (void)slot;
#endif
}

template <int BN, int MODE, int PT>
__global__ void __launch_bounds__(tm_threads(BN), 1)
conv_tma_kernel(const ConvKArgs a, const __grid_constant__ TmaMaps maps, const TmaTiling tg) {
    extern __shared__ __align__(1024) uint8_t tm_smem[];
    pdl_trigger();                                            // the next kernel may be scheduled behind this one
    constexpr int B_TILE_FLOATS = BN * TC_KB;                 // one of hi / lo
    constexpr uint32_t B_STAGE_BYTES = 2u * B_TILE_FLOATS * sizeof(float);
    constexpr uint32_t RAW_BYTES = TM_RAW_FLOATS * sizeof(float);
    float* rawbuf = reinterpret_cast<float*>(tm_smem);
    uint8_t* bbuf = tm_smem + TM_NR * RAW_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + TC_NST * B_STAGE_BYTES);
    uint64_t* raw_full = bars;                 // [NR]  TMA box landed (expect_tx)
    uint64_t* raw_free = bars + TM_NR;         // [NR]  256 transform threads have read the box for the last time
    uint64_t* a_full = raw_free + TM_NR;       // [NST] the 128 transform threads of the owning team + the weight producer's expect_tx
    uint64_t* ab_free = a_full + TC_NST;       // [NST] tcgen05.commit: the MMAs have read the stage
    uint64_t* d_full = ab_free + TC_NST;       // [2]   tcgen05.commit: block accumulator complete
    uint64_t* d_empty = d_full + 2;            // [2]   128 drain threads arrive
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const ConvGeom& g = a.g;
    const int Cd = MODE == MODE_FWD ? g.Cout : g.Cin;
    const int Hd = MODE == MODE_FWD ? g.OH : g.H;
    const int W = g.W;                                         // = g.OW
    const long long d_ss = MODE == MODE_FWD ? g.out_sstride : g.in_sstride;
    const int Cs = MODE == MODE_FWD ? g.Cin : g.Cout;
    const int HWd = Hd * W;
    const long long J = (long long)g.batch * HWd;
    const int nchunks = (Cs + TC_KB - 1) / TC_KB;
    const int KBp = g.KH * g.KW * nchunks;                     // k-blocks per pair
    const int n_jt = (int)((J + TC_M - 1) / TC_M);
    const int n_nt = (Cd + BN - 1) / BN;
    const int total_tiles = n_jt * n_nt;
    const int last_ksteps = (Cs - (nchunks - 1) * TC_KB + 7) >> 3;
    const int dbg = tg.dbg;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = tm_teams(BN);                           // transform teams
    constexpr int TM_XFORM_THREADS = NT * 128;
    constexpr int TM_WARP_DRAIN0 = NT * 4;                     // first drain warp
    constexpr int TM_WARP_MMA = NT * 4 + 4, TM_WARP_B = NT * 4 + 5, TM_WARP_RAW = NT * 4 + 6;

    if (tid == 0) {
        for (int s = 0; s < TM_NR; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_free[s], TM_XFORM_THREADS);
        }
        for (int s = 0; s < TC_NST; ++s) {
            mbar_init(&a_full[s], 128 + 1);
            mbar_init(&ab_free[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TM_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                               // everything above overlapped the previous kernel's tail

    if (warp < TM_WARP_DRAIN0) {
        // ===================== transform: landed box -> registers (shift, split) -> TMEM =====================
        // PT = pixels of one image inside the tile (128, or the image size when the tile holds several images)
        const int r = tid & 127, team = tid >> 7, q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int nl = r / PT, pl = r - nl * PT;
        const int x = pl % W;
        const int roff = nl * tg.KC * PT + pl;                  // this thread's pixel, channel 0
        uint32_t kbg = 0, gidx = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            for (int p = 0; p < a.npairs; ++p) {
                for (int ky = 0; ky < g.KH; ++ky) {
                    for (int cc = 0; cc < nchunks; ++cc, ++gidx) {
                        const int rs = gidx % TM_NR;
                        const int ksteps = cc == nchunks - 1 ? last_ksteps : TC_KB / 8;
                        if (tid == 0) tm_stamp(a.trace, 3, kbg, 2);
                        tm_wait(&raw_full[rs], (gidx / TM_NR) & 1, dbg, 100 + rs);
                        if (tid == 0) tm_stamp(a.trace, 3, kbg, 3);
                        const float* __restrict__ raw = rawbuf + rs * TM_RAW_FLOATS + roff;
                        for (int kx = 0; kx < g.KW; ++kx, ++kbg) {
                            if ((int)(kbg % NT) != team) continue;
                            if ((tid & 127) == 0) tm_stamp(a.trace, 0, kbg, 0);
                            const int dx = MODE == MODE_FWD ? kx - g.pw : g.pw - kx;
                            const bool ok = (unsigned)(x + dx) < (unsigned)W;
                            const float* __restrict__ src = raw + dx;
                            float v[TC_KB];
#pragma unroll
                            for (int j = 0; j < TC_KB / 8; ++j) {
                                if (j < ksteps) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[j * 8 + i] = ok ? src[(j * 8 + i) * PT] : 0.f;
                                }
                            }
                            const int s = kbg % TC_NST;
                            const uint32_t round = kbg / TC_NST;
                            if (round > 0) tm_wait(&ab_free[s], (round - 1) & 1, dbg, 200 + s);
                            __syncwarp();
                            tc_fence_after();
                            if ((tid & 127) == 0) tm_stamp(a.trace, 0, kbg, 1);
                            const uint32_t col = (uint32_t)(s * TC_ACOLS);
#pragma unroll
                            for (int j = 0; j < TC_KB / 8; ++j) {
                                if (j < ksteps) {
                                    uint32_t hi[8], lo[8];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        // hi = TF32 round-to-nearest (magnitude + half ulp, low 13 bits cleared);
                                        // lo = f - hi is exact in fp32 (<= 12 significant bits); the tensor core
                                        // reads its upper 19 bits
                                        const float f = v[j * 8 + i];
#ifdef B2S_TMA_TRUNC                     // experiment (B2S_BUILD_TRUNC=1): truncating split, one ALU instruction less per element
                                        hi[i] = __float_as_uint(f) & 0xffffe000u;
#else
                                        hi[i] = (__float_as_uint(f) + 0x1000u) & 0xffffe000u;
#endif
                                        lo[i] = __float_as_uint(f - __uint_as_float(hi[i]));
                                    }
                                    tmem_st8(lane_addr + col + j * 8, hi);
                                    tmem_st8(lane_addr + col + 32 + j * 8, lo);
                                }
                            }
                            if ((tid & 127) == 0) tm_stamp(a.trace, 0, kbg, 2);
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                            tc_fence_before();
                            mbar_arrive(&a_full[s]);
                            if ((tid & 127) == 0) tm_stamp(a.trace, 0, kbg, 3);
                        }
                        mbar_arrive(&raw_free[rs]);            // every read of the box has been consumed by the stores above
                    }
                }
            }
        }
    } else if (warp < TM_WARP_DRAIN0 + 4) {
        // ===================== drain + epilogue: TMEM -> fp32 registers -> NCHW global ===============
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(tm_regs_drain(BN)));
        const int q = warp - TM_WARP_DRAIN0;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t grp = 0;                                       // drain groups of TM_G k-blocks (never across pairs)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int jt = tile % n_jt, nt = tile / n_jt;
            float acc[BN];
#pragma unroll
            for (int i = 0; i < BN; ++i) acc[i] = 0.f;
            constexpr int NACC = tc_nacc(BN);
            for (int p = 0; p < a.npairs; ++p) {
                const float sc = a.scale[p];
                for (int k0 = 0; k0 < KBp; k0 += TM_G, ++grp) {
                    const int b = grp & 1;
                    if (tid == TM_XFORM_THREADS) tm_stamp(a.trace, 2, grp, 0);
                    tm_wait(&d_full[b], (grp >> 1) & 1, dbg, 300 + b);
                    __syncwarp();
                    tc_fence_after();
                    if (tid == TM_XFORM_THREADS) tm_stamp(a.trace, 2, grp, 1);
                    // with 6 accumulators the odd-k-step set is only written by k-blocks of >= 2 k-steps
                    // (k-block order (ky, chunk, kx): the chunk index is (kbl / KW) % nchunks)
                    bool odd_set = false;
                    if (NACC == 6) {
                        if (last_ksteps >= 2) odd_set = true;
                        else
                            for (int kbl = k0; kbl < min(k0 + TM_G, KBp); ++kbl) odd_set |= ((kbl / g.KW) % nchunks) != nchunks - 1;
                    }
                    const uint32_t d0 = lane_addr + TC_DCOL0 + b * TC_DCOLS;
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 16) {
                        uint32_t v[NACC >= 3 ? 3 : NACC][16];
#pragma unroll
                        for (int q2 = 0; q2 < (NACC >= 3 ? 3 : NACC); ++q2) tmem_ld16(d0 + q2 * BN + c0, v[q2]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            // set layout [lo*hi | hi*hi | hi*lo] (3+ accumulators) or [hi*hi | hi*lo + lo*hi] (2): small terms first
                            float d;
                            if (NACC >= 3) d = (__uint_as_float(v[0][i]) + __uint_as_float(v[2][i])) + __uint_as_float(v[1][i]);
                            else if (NACC == 2) d = __uint_as_float(v[1][i]) + __uint_as_float(v[0][i]);
                            else d = __uint_as_float(v[0][i]);
                            acc[c0 + i] = fmaf(d, sc, acc[c0 + i]);
                        }
                        if (NACC == 6 && odd_set) {
                            uint32_t w[3][16];
#pragma unroll
                            for (int q2 = 0; q2 < 3; ++q2) tmem_ld16(d0 + (3 + q2) * BN + c0, w[q2]);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                acc[c0 + i] = fmaf((__uint_as_float(w[0][i]) + __uint_as_float(w[2][i])) + __uint_as_float(w[1][i]), sc, acc[c0 + i]);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&d_empty[b]);
                    if (tid == TM_XFORM_THREADS) tm_stamp(a.trace, 2, grp, 2);
                }
            }
            // epilogue: lane = pixel, so every per-channel store is one coalesced 128-byte row.  The variant
            // (accumulate, ReLU mode) is chosen ONCE per tile: the generic form (flags tested per element) compiled
            // to ~20 instructions and a branch per channel, executed by one warp per scheduler -- 6600 clocks per
            // 48-channel tile in the timeline, more than the whole k-loop of a 1x1 convolution.
            const long long j = (long long)jt * TC_M + r;
            if (j < J) {
                const int n = (int)(j / HWd);
                const int pix = (int)(j - (long long)n * HWd);
                const int m0 = nt * BN;
                const long long off = (long long)n * d_ss + pix + (long long)m0 * HWd;
                const int mrem = Cd - m0;                 // valid channels of this tile
                tc_store_tile<BN>(acc, a, off, HWd, m0, mrem);
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_MISC));
        if (warp == TM_WARP_MMA) {
            // ===================== MMA issuer ======================================================
            const uint32_t idesc = umma_idesc_tf32(TC_M, BN);
            const uint32_t idesc2 = umma_idesc_tf32(TC_M, 2 * BN <= 256 ? 2 * BN : BN);   // hi and lo weight tiles in one MMA
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);           // provably warp-uniform
            const uint64_t desc0 = umma_desc(smem_u32(bbuf), 128, 1024);         // stage 0, hi tile, k-step 0
            constexpr int NACC = tc_nacc(BN);
            uint32_t kbg = 0, grp = 0;
            const int mma_spin = (tg.dbg & 16) ? 8 : 0;           // B2S_TMA_DBG bit 4: the MMA issuer spins on its barriers
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int p = 0; p < a.npairs; ++p) {
                    int kbl = 0;
                    bool even_w = false, odd_w = false;            // accumulator sets already written in this drain group
                    for (int ky = 0; ky < g.KH; ++ky) {
                        for (int cc = 0; cc < nchunks; ++cc) {
                            const int ksteps = cc == nchunks - 1 ? last_ksteps : TC_KB / 8;
                            for (int kx = 0; kx < g.KW; ++kx, ++kbl, ++kbg) {
                                const int s = kbg % TC_NST;
                                const uint32_t round = kbg / TC_NST;
                                const int b = grp & 1;
                                const uint32_t use = grp >> 1;
                                const bool first = (kbl % TM_G) == 0;
                                const bool last = (kbl % TM_G) == TM_G - 1 || kbl == KBp - 1;
                                if (lane == 0) tm_stamp(a.trace, 1, kbg, 0);
                                tm_wait(&a_full[s], round & 1, dbg | mma_spin, 400 + s);
                                if (lane == 0) tm_stamp(a.trace, 1, kbg, 1);
                                if (first && use > 0) tm_wait(&d_empty[b], (use - 1) & 1, dbg | mma_spin, 500 + b);
                                __syncwarp();
                                tc_fence_after();
                                if (lane == 0) tm_stamp(a.trace, 1, kbg, 2);
                                const uint32_t d_addr = tmem_u + TC_DCOL0 + b * TC_DCOLS;
                                const uint32_t a_hi = tmem_u + s * TC_ACOLS, a_lo = a_hi + 32;
                                const uint64_t dBh = desc0 + (uint64_t)((s * B_STAGE_BYTES) >> 4);
                                const uint64_t dBl = dBh + (uint64_t)((B_TILE_FLOATS * 4) >> 4);
                                const uint32_t ew = even_w ? 1u : 0u, ow = odd_w ? 1u : 0u;
                                if (elect_one()) {
                                    // Accumulator set of a k-step: with 6 accumulators odd and even k-steps use different sets
                                    // (consecutive MMAs into one accumulator serialise on its read-modify-write latency).
                                    // Layout of a set: [lo*hi | hi*hi | hi*lo], BN columns each.  The hi and lo weight tiles
                                    // are adjacent in shared memory, so  A_hi x [B_hi | B_lo]  is ONE MMA with N = 2 BN:
                                    // two instead of three MMAs per k-step (the issuing thread is the limiter of this kernel).
#pragma unroll
                                    for (int ks = 0; ks < TC_KB / 8; ++ks) {
                                        if (ks < ksteps) {
                                            const uint64_t ko = (uint64_t)(ks * 16);      // 2 core matrices of 128 B per k-step
                                            // accumulate onto what this drain group has already put into the accumulator
                                            const uint32_t accf = NACC == 6 ? ((ks & 1) ? (ks >= 2 ? 1u : ow) : (ks >= 2 ? 1u : ew)) : (ks >= 1 ? 1u : ew);
                                            if (NACC >= 3) {
                                                const uint32_t set = d_addr + (NACC == 6 ? 3 * (ks & 1) * BN : 0);
                                                umma_tf32_ts(set + BN, a_hi + ks * 8, dBh + ko, idesc2, accf);     // [hi*hi | hi*lo]
                                                umma_tf32_ts(set, a_lo + ks * 8, dBh + ko, idesc, accf);           // lo*hi
                                            } else if (NACC == 2) {
                                                umma_tf32_ts(d_addr, a_hi + ks * 8, dBh + ko, idesc2, accf);       // [hi*hi | hi*lo]
                                                umma_tf32_ts(d_addr + BN, a_lo + ks * 8, dBh + ko, idesc, 1u);     // hi*lo += lo*hi
                                            } else {
                                                umma_tf32_ts(d_addr, a_hi + ks * 8, dBl + ko, idesc, accf);
                                                umma_tf32_ts(d_addr, a_lo + ks * 8, dBh + ko, idesc, 1u);
                                                umma_tf32_ts(d_addr, a_hi + ks * 8, dBh + ko, idesc, 1u);
                                            }
                                        }
                                    }
                                    umma_commit(&ab_free[s]);                 // arrives when these MMAs have read the stage
                                    if (last) umma_commit(&d_full[b]);        // ... and when the group's accumulator is complete
                                }
                                __syncwarp();
                                if (lane == 0) tm_stamp(a.trace, 1, kbg, 3);
                                even_w = true;
                                odd_w = odd_w || ksteps >= 2;
                                if (last) { ++grp; even_w = false; odd_w = false; }
                            }
                        }
                    }
                }
            }
        } else if (warp == TM_WARP_B) {
            // ===================== weight producer: bulk copies of the packed images ====================
            uint32_t kbg = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile / n_jt;
                for (int p = 0; p < a.npairs; ++p) {
                    const float* __restrict__ img = a.pack[p] + (long long)nt * KBp * 2 * B_TILE_FLOATS;
                    for (int ky = 0; ky < g.KH; ++ky) {
                        for (int cc = 0; cc < nchunks; ++cc) {
                            for (int kx = 0; kx < g.KW; ++kx, ++kbg) {
                                const int kbl = (ky * g.KW + kx) * nchunks + cc;          // index in the packed image (tap major)
                                const int s = kbg % TC_NST;
                                const uint32_t round = kbg / TC_NST;
                                if (lane == 0) {
                                    if (round > 0) tm_wait(&ab_free[s], (round - 1) & 1, dbg, 600 + s);
                                    mbar_arrive_expect_tx(&a_full[s], B_STAGE_BYTES);
                                    bulk_g2s(bbuf + s * B_STAGE_BYTES, img + (long long)kbl * 2 * B_TILE_FLOATS, B_STAGE_BYTES, &a_full[s]);
                                }
                                __syncwarp();
                            }
                        }
                    }
                }
            }
        } else if (warp == TM_WARP_RAW) {
            // ===================== activation producer: one TMA box per (pair, ky, chunk) =================
            uint32_t gidx = 0;
            const uint32_t box_bytes = (uint32_t)(TC_M * tg.KC * 4);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int jt = tile % n_jt;
                const long long j0 = (long long)jt * TC_M;
                const int n0 = (int)(j0 / HWd);
                const int y0 = (int)(j0 - (long long)n0 * HWd) / W;
                for (int p = 0; p < a.npairs; ++p) {
                    for (int ky = 0; ky < g.KH; ++ky) {
                        const int dy = MODE == MODE_FWD ? ky - g.ph : g.ph - ky;
                        for (int cc = 0; cc < nchunks; ++cc, ++gidx) {
                            const int rs = gidx % TM_NR;
                            const uint32_t round = gidx / TM_NR;
                            if (lane == 0) {
                                if (round > 0) tm_wait(&raw_free[rs], (round - 1) & 1, dbg, 700 + rs);
                                mbar_arrive_expect_tx(&raw_full[rs], box_bytes);
                                tma_load_4d(rawbuf + rs * TM_RAW_FLOATS, &maps.m[p], 0, y0 + dy, cc * TC_KB, n0, &raw_full[rs]);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TM_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
TmEncodeFn tm_encoder() {
    static TmEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TmEncodeFn>(p);
    }();
    return fn;
}

template <int BN, int MODE, int PT>
static int launch_tma_t(cudaStream_t st, const ConvKArgs& a, const TmaMaps& maps, const TmaTiling& tg, long long J, int Cd) {
    constexpr size_t smem = (size_t)TM_NR * TM_RAW_FLOATS * 4 + (size_t)TC_NST * 2 * BN * TC_KB * sizeof(float) + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tma_kernel<BN, MODE, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        attr_set = true;
    }
    const long long tiles = ((J + TC_M - 1) / TC_M) * ((Cd + BN - 1) / BN);
    const unsigned grid = (unsigned)std::min<long long>(tiles, kNumSMs);
    static const bool want_trace = getenv("B2S_TC_TRACE") != nullptr;
    static int traced = 0;
    if (want_trace && traced < 8) {            // debug: synchronous launch + timeline dump of CTA 0 (eager mode only)
        ++traced;
        long long* d_tr = nullptr;
        const size_t n = 4 * TM_TRACE_KB * 4;
        cudaMalloc(&d_tr, n * sizeof(long long));
        cudaMemset(d_tr, 0, n * sizeof(long long));
        ConvKArgs b = a;
        b.trace = d_tr;
        conv_tma_kernel<BN, MODE, PT><<<grid, tm_threads(BN), smem, st>>>(b, maps, tg);
        cudaStreamSynchronize(st);
        std::vector<long long> h(n);
        cudaMemcpy(h.data(), d_tr, n * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(d_tr);
        long long t0 = 0;
        for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        fprintf(stderr, "TMA trace BN=%d mode=%d J=%lld Cs=%d Cd=%d KHW=%d pairs=%d grid=%u (clocks since first stamp)\n", BN, MODE, J,
                MODE == MODE_FWD ? a.g.Cin : a.g.Cout, Cd, a.g.KH * a.g.KW, a.npairs, grid);
        for (int k = 0; k < TM_TRACE_KB; ++k) {
            auto at = [&](int role, int slot) { long long v = h[((size_t)role * TM_TRACE_KB + k) * 4 + slot]; return v ? v - t0 : -1; };
            if (at(1, 0) < 0) break;
            fprintf(stderr, "  kb %2d  raw: wait %6lld ok %6lld | A: start %6lld free %6lld st %6lld done %6lld | MMA: poll %6lld a_ok %6lld all_ok %6lld issued %6lld | D: poll %6lld full %6lld done %6lld\n",
                    k, at(3, 2), at(3, 3), at(0, 0), at(0, 1), at(0, 2), at(0, 3), at(1, 0), at(1, 1), at(1, 2), at(1, 3), at(2, 0), at(2, 1), at(2, 2));
        }
        return 1;
    }
    {
        const cudaError_t e = launch_pdl(conv_tma_kernel<BN, MODE, PT>, dim3(grid), dim3(tm_threads(BN)), smem, st, a, maps, tg);
        if (e != cudaSuccess) { set_error("conv_tma launch: %s", cudaGetErrorString(e)); return -3; }
    }
    if (tg.dbg & 7) {
        const cudaError_t e = cudaStreamSynchronize(st);
        unsigned flag = 0;
        cudaMemcpyFromSymbol(&flag, tm_timeout_flag, sizeof(flag));
        fprintf(stderr, "conv_tma dbg=%d BN=%d mode=%d grid=%u W=%d BH=%d BI=%d KC=%d tiles=%lld: %s, timeout id %u\n", tg.dbg, BN, MODE, grid,
                a.g.W, tg.BH, tg.BI, tg.KC, tiles, cudaGetErrorString(e), flag);
    }
    return 1;
}

template <int MODE, int PT>
static int launch_tma_pt(cudaStream_t st, const ConvKArgs& a, const TmaMaps& maps, const TmaTiling& tg, long long J, int Cd) {
    switch (tc_choose_bn(Cd)) {
    case 16: return launch_tma_t<16, MODE, PT>(st, a, maps, tg, J, Cd);
    case 32: return launch_tma_t<32, MODE, PT>(st, a, maps, tg, J, Cd);
    case 48: return launch_tma_t<48, MODE, PT>(st, a, maps, tg, J, Cd);
    case 64: return launch_tma_t<64, MODE, PT>(st, a, maps, tg, J, Cd);
    case 96: return launch_tma_t<96, MODE, PT>(st, a, maps, tg, J, Cd);
    default: return launch_tma_t<128, MODE, PT>(st, a, maps, tg, J, Cd);
    }
}
template <int MODE>
static int launch_tma_mode(cudaStream_t st, const ConvKArgs& a, const TmaMaps& maps, const TmaTiling& tg, long long J, int Cd) {
    const int PT = tg.BH * a.g.W;               // pixels of one image inside the tile
    switch (PT) {
    case 128: return launch_tma_pt<MODE, 128>(st, a, maps, tg, J, Cd);
    case 64: return launch_tma_pt<MODE, 64>(st, a, maps, tg, J, Cd);
    default: return 0;                          // smaller images stay on the gather kernel
    }
}

// Returns 1 when the kernel was launched, 0 when the layer is not eligible, <0 on error.
int try_launch_conv_tma(int mode, cudaStream_t st, const ConvKArgs& a) {
    static const int enabled = getenv("B2S_TMA") ? atoi(getenv("B2S_TMA")) : 1;
    if (!enabled || get_tc_mode() == 0) return 0;
    for (int p = 0; p < a.npairs; ++p)
        if (!a.pack[p]) return 0;                     // the plan did not pack this layer (shape not eligible)
    const ConvGeom& g = a.g;
    if (g.sh != 1 || g.sw != 1 || g.W != g.OW) return 0;
    const bool fwd = mode == MODE_FWD;
    const int Cd = fwd ? g.Cout : g.Cin, Cs = fwd ? g.Cin : g.Cout;
    const int Hd = fwd ? g.OH : g.H, Hs = fwd ? g.H : g.OH;
    const int W = g.W;
    const long long s_ss = fwd ? g.in_sstride : g.out_sstride;
    const long long J = (long long)g.batch * Hd * W;
    if (get_tc_mode() == 1 && !tc_worth_it(J, Cs, Cd, g.KH * g.KW)) return 0;
    // the tile is 128 consecutive pixels made of full rows: W in {4, 8, 16, 32}; it either divides an image or
    // holds whole images
    if (W < 4 || W > 32 || (TC_M % W) != 0 || (s_ss & 3)) return 0;
    if (Hd * W < 64) return 0;
    const int HWd = Hd * W;
    TmaTiling tg;
    if (HWd % TC_M == 0) { tg.BH = TC_M / W; tg.BI = 1; }
    else if (TC_M % HWd == 0) { tg.BH = Hd; tg.BI = TC_M / HWd; }
    else return 0;
    for (int p = 0; p < a.npairs; ++p)
        if ((uintptr_t)a.act[p] & 15) return 0;
    tg.KC = std::min(TC_KB, (Cs + 7) & ~7);
    static const int dbg = getenv("B2S_TMA_DBG") ? atoi(getenv("B2S_TMA_DBG")) : 0;
    tg.dbg = dbg;
    TmEncodeFn enc = tm_encoder();
    if (!enc) return 0;
    TmaMaps maps;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)Hs, (cuuint64_t)Cs, (cuuint64_t)g.batch};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)Hs * W * 4, (cuuint64_t)s_ss * 4};
    const cuuint32_t box[4] = {(cuuint32_t)W, (cuuint32_t)tg.BH, (cuuint32_t)tg.KC, (cuuint32_t)tg.BI};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int p = 0; p < a.npairs; ++p) {
        const CUresult r = enc(&maps.m[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.act[p]), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            static bool warned = false;
            if (!warned) {
                warned = true;
                fprintf(stderr, "b2s: cuTensorMapEncodeTiled failed (%d) for W=%d H=%d C=%d N=%d sstride=%lld; using the gather kernel\n",
                        (int)r, W, Hs, Cs, g.batch, s_ss);
            }
            return 0;
        }
    }
    for (int p = a.npairs; p < kMaxPairs; ++p) maps.m[p] = maps.m[0];
    return fwd ? launch_tma_mode<MODE_FWD>(st, a, maps, tg, J, Cd) : launch_tma_mode<MODE_DGRAD>(st, a, maps, tg, J, Cd);
}

}  // namespace b2s
