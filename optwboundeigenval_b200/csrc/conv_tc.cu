// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM for the Conv2d / Linear jet contractions (sm_100a).
//
// Same contraction as conv.cu (forward and input-adjoint, up to three K-concatenated
// (activation, weight) pairs per launch) on the 5th-generation tensor cores, fp32-accurate:
//
//   D[128 pixels x BN channels] (fp32, TMEM)  =  A[128 x 32] * B[BN x 32]^T      per k-block
//
// fp32 accuracy (rtol 1e-4 parity) needs two things on this hardware:
//  (1) 3xTF32 operand split: x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi); each k-step issues
//      hi*hi + hi*lo + lo*hi (lo*lo ~ 2^-22 is dropped);
//  (2) the tensor core adds into its fp32 accumulator with truncation (measured with
//      tools/tc_accuracy.py: a K = 4608 all-positive contraction accumulated in TMEM comes out 5e-5
//      low), so the accumulator is NOT kept in TMEM across the whole K loop: every k-block (K = 32)
//      starts a fresh TMEM accumulator, and dedicated drain warps add the finished block into fp32
//      registers with round-to-nearest FMAs while the tensor core works on the next block in the
//      other TMEM buffer.
//
// Data movement, per CTA (persistent over output tiles), warp-specialised:
//   warps 0-3  A transform: lane = pixel (coalesced NCHW reads, im2col on the fly with zero padding),
//              split hi/lo in registers, tcgen05.st straight into TMEM -- the A operand never touches
//              shared memory (tcgen05.mma with A in TMEM), no bank conflicts, no proxy fence;
//   warp  9    B producer: weights are pre-split and pre-arranged ONCE per pass by tc_pack_kernel into
//              the exact shared-memory image (K-major, SWIZZLE_NONE core matrices) of every k-block,
//              so a stage is one cp.async.bulk (bulk-copy engine, mbarrier complete_tx);
//   warp  8    MMA issuer: one thread, 3 x tcgen05.mma.kind::tf32 per k-step, tcgen05.commit to the
//              stage-free and accumulator-full mbarriers;
//   warps 4-7  drain + epilogue: tcgen05.ld the block accumulator, acc += scale * d, and after the last
//              k-block bias / ReLU mask / accumulate and coalesced NCHW stores (lane = pixel).
// Four stages of A (TMEM) and B (smem); two accumulator buffers.  Descriptor encodings follow
// cute/arch/mma_sm100_desc.hpp and cute/arch/mma_sm100_umma.hpp (SM100_MMA_TF32_TS).
#include <algorithm>
#include <cstdlib>

#include "conv_args.h"

namespace b2s {

// process-wide switch; the environment variable B2S_TC_MODE (0/1/2) overrides the default for debugging
static int initial_tc_mode() {
    const char* e = getenv("B2S_TC_MODE");
    return (e && e[0] >= '0' && e[0] <= '2' && e[1] == 0) ? e[0] - '0' : 1;
}
static int g_tc_mode = initial_tc_mode();
void set_tc_mode(int mode) { g_tc_mode = mode; }
int get_tc_mode() { return g_tc_mode; }

constexpr int TC_M = 128;        // pixels per tile (UMMA M, cta_group::1)
constexpr int TC_KB = 32;        // k per k-block = 4 UMMA k-steps of 8 (tf32)
constexpr int TC_NST = 4;        // pipeline stages (A in TMEM, B in shared memory)
constexpr int TC_ACOLS = 64;     // TMEM columns per A stage: 32 hi + 32 lo
constexpr int TC_DCOL0 = TC_NST * TC_ACOLS;    // first accumulator column (two buffers of BN columns)
constexpr int TC_THREADS = 320;  // 10 warps, roles above

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (between the two 16 B K-chunks of a
//   k-step = adjacent core matrices along K), [32,46) stride byte offset >> 4 (between 8-row groups),
//   [46,48) version = 1 (Blackwell), [61,64) layout type = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6), a/b_format TF32 = 2 @
// [7,10)/[10,13), a/b major K = 0 @ 15/16, n_dim = N >> 3 @ [17,23), m_dim = M >> 4 @ [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T   (SM100_MMA_TF32_TS)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol error must become a trap (launch failure), never a hang
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t to_tf32_bits(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// ---------------------------------------------------------------------------------------------
// weight packing: flat parameter vector -> per (op, mode) images of every k-block's B tile
// ---------------------------------------------------------------------------------------------
// k ordering of a pair: k-block kbl = tap * nchunks + chunk, element k inside = source channel
// chunk * 32 + k (zero padded), so that one k-block of the A gather is 32 channels of ONE tap
// (one shifted pixel per thread, channel stride = plane size).
// image[(ntile * KBp + kbl) * 2 + {hi, lo}][BN x 32 tile], tile element (m, k) at float offset
// ((m / 8) * 8 + k / 4) * 32 + (m % 8) * 4 + k % 4   (8 x 16 B core matrices, LBO 128 B, SBO 1024 B)
__global__ void __launch_bounds__(256) tc_pack_kernel(const TcPackJob* __restrict__ jobs, int njobs, long long total,
                                                      const float* __restrict__ src_base, float* __restrict__ dst_base) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].begin <= e) lo = mid; else hi = mid - 1;
        }
        const TcPackJob jb = jobs[lo];
        unsigned l = (unsigned)(e - jb.begin);                    // one image has < 2^32 elements
        const int k = (int)(l % TC_KB); l /= TC_KB;
        const int m = (int)(l % (unsigned)jb.BN); l /= (unsigned)jb.BN;
        const int KBp = jb.KHW * jb.nchunks;
        const int kbl = (int)(l % (unsigned)KBp);
        const int nt = (int)(l / (unsigned)KBp);
        const int t = kbl / jb.nchunks, cc = kbl - t * jb.nchunks;
        const int c = cc * TC_KB + k;
        const int mm = nt * jb.BN + m;
        float v = 0.f;
        if (c < jb.Cs && mm < jb.Cd) {
            const float* w = src_base + jb.src_off;
            if (jb.mode == MODE_FWD) v = w[((long long)mm * jb.Cs + c) * jb.KHW + t];     // [Cout = Cd][Cin = Cs][tap]
            else v = w[((long long)c * jb.Cd + mm) * jb.KHW + t];                         // [Cout = Cs][Cin = Cd][tap]
        }
        const uint32_t h = to_tf32_bits(v);
        const uint32_t lw = to_tf32_bits(v - __uint_as_float(h));
        float* tile = dst_base + jb.dst_off + ((long long)nt * KBp + kbl) * 2 * jb.BN * TC_KB;
        const int o = ((m >> 3) * 8 + (k >> 2)) * 32 + (m & 7) * 4 + (k & 3);
        tile[o] = __uint_as_float(h);
        tile[jb.BN * TC_KB + o] = __uint_as_float(lw);
    }
}

int tc_choose_bn(int Cd) {
    const int cand[6] = {16, 32, 48, 64, 96, 128};
    for (int c : cand)
        if (Cd <= c) return c;
    return 128;
}

long long tc_pack_floats(int Cs, int Cd, int KHW) {
    const int BN = tc_choose_bn(Cd);
    const int ntiles = (Cd + BN - 1) / BN;
    const int nchunks = (Cs + TC_KB - 1) / TC_KB;
    return (long long)ntiles * KHW * nchunks * 2 * BN * TC_KB;
}

bool tc_shape_ok(int Cs, int Cd, int Hs, int Ws) {
    return Cs >= 8 && Cd >= 8 && Hs < 32768 && Ws < 32768;
}

int launch_tc_pack(cudaStream_t st, const TcPackJob* d_jobs, int njobs, long long total, const float* src_base,
                   float* dst_base) {
    if (njobs <= 0 || total <= 0) return 0;
    ProfScope prof("tc_pack", 0.0, 12.0 * (double)total, st);
    long long blocks = (total + 255) / 256;
    if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
    tc_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_jobs, njobs, total, src_base, dst_base);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the contraction
// ---------------------------------------------------------------------------------------------
template <int BN, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvKArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int B_TILE_FLOATS = BN * TC_KB;                 // one of hi / lo
    constexpr uint32_t B_STAGE_BYTES = 2u * B_TILE_FLOATS * sizeof(float);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + TC_NST * B_STAGE_BYTES);
    uint64_t* a_full = bars;                   // [NST] 128 transform threads arrive
    uint64_t* b_full = bars + TC_NST;          // [NST] bulk copy complete_tx
    uint64_t* ab_free = bars + 2 * TC_NST;     // [NST] tcgen05.commit: the MMAs have read the stage
    uint64_t* d_full = bars + 3 * TC_NST;      // [2]   tcgen05.commit: block accumulator complete
    uint64_t* d_empty = d_full + 2;            // [2]   128 drain threads arrive
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const ConvGeom& g = a.g;
    const int Cd = MODE == MODE_FWD ? g.Cout : g.Cin;
    const int Hd = MODE == MODE_FWD ? g.OH : g.H;
    const int Wd = MODE == MODE_FWD ? g.OW : g.W;
    const long long d_ss = MODE == MODE_FWD ? g.out_sstride : g.in_sstride;
    const int Cs = MODE == MODE_FWD ? g.Cin : g.Cout;
    const int Hs = MODE == MODE_FWD ? g.H : g.OH;
    const int Ws = MODE == MODE_FWD ? g.W : g.OW;
    const long long s_ss = MODE == MODE_FWD ? g.in_sstride : g.out_sstride;
    const int KHW = g.KH * g.KW;
    const int HWd = Hd * Wd, HWs = Hs * Ws;
    const long long J = (long long)g.batch * HWd;
    const int nchunks = (Cs + TC_KB - 1) / TC_KB;
    const int KBp = KHW * nchunks;                            // k-blocks per pair
    const int n_jt = (int)((J + TC_M - 1) / TC_M);
    const int n_nt = (Cd + BN - 1) / BN;
    const int total_tiles = n_jt * n_nt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < TC_NST; ++s) {
            mbar_init(&a_full[s], 128);
            mbar_init(&b_full[s], 1);
            mbar_init(&ab_free[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        // ===================== A transform: global -> registers (split) -> TMEM =====================
        const int r = warp * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t kbg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int jt = tile % n_jt;
            const long long j = (long long)jt * TC_M + r;
            const bool in_range = j < J;
            int n = 0, y = 0, x = 0;
            if (in_range) {
                n = (int)(j / HWd);
                const int pix = (int)(j - (long long)n * HWd);
                y = pix / Wd;
                x = pix - y * Wd;
            }
            for (int p = 0; p < a.npairs; ++p) {
                const float* __restrict__ plane0 = a.act[p] + (long long)n * s_ss;
                for (int t = 0; t < KHW; ++t) {
                    const int ky = t / g.KW, kx = t - ky * g.KW;
                    int sy, sx;
                    bool ok;
                    if (MODE == MODE_FWD) {
                        sy = y * g.sh + ky - g.ph; sx = x * g.sw + kx - g.pw;
                        ok = sy >= 0 && sy < Hs && sx >= 0 && sx < Ws;
                    } else {
                        const int ty_ = y + g.ph - ky, tx_ = x + g.pw - kx;
                        sy = ty_ / g.sh; sx = tx_ / g.sw;
                        ok = ty_ >= 0 && tx_ >= 0 && sy * g.sh == ty_ && sx * g.sw == tx_ && sy < Hs && sx < Ws;
                    }
                    ok = ok && in_range;
                    const float* __restrict__ src = plane0 + (ok ? sy * Ws + sx : 0);
                    for (int cc = 0; cc < nchunks; ++cc) {
                        const int nvalid = min(TC_KB, Cs - cc * TC_KB);
                        const int ksteps = (nvalid + 7) >> 3;
                        const int s = kbg % TC_NST;
                        const uint32_t round = kbg / TC_NST;
                        // issue the loads before waiting for the stage: they only need registers
                        float v[TC_KB];
                        const float* __restrict__ sc_ = src + (long long)cc * TC_KB * HWs;
#pragma unroll
                        for (int c = 0; c < TC_KB; ++c) v[c] = (ok && c < nvalid) ? __ldg(sc_ + (long long)c * HWs) : 0.f;
                        if (round > 0) mbar_wait(&ab_free[s], (round - 1) & 1);
                        __syncwarp();
                        tc_fence_after();
                        const uint32_t col = (uint32_t)(s * TC_ACOLS);
#pragma unroll
                        for (int ks = 0; ks < TC_KB / 8; ++ks) {
                            if (ks < ksteps) {
                                uint32_t hi[8], lo[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float f = v[ks * 8 + i];
                                    hi[i] = to_tf32_bits(f);
                                    lo[i] = to_tf32_bits(f - __uint_as_float(hi[i]));
                                }
                                tmem_st8(lane_addr + col + ks * 8, hi);
                                tmem_st8(lane_addr + col + 32 + ks * 8, lo);
                            }
                        }
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        mbar_arrive(&a_full[s]);
                        ++kbg;
                    }
                }
            }
        }
    } else if (warp < 8) {
        // ===================== drain + epilogue: TMEM -> fp32 registers -> NCHW global ===============
        const int q = warp - 4;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t kbg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int jt = tile % n_jt, nt = tile / n_jt;
            float acc[BN];
#pragma unroll
            for (int i = 0; i < BN; ++i) acc[i] = 0.f;
            for (int p = 0; p < a.npairs; ++p) {
                const float sc = a.scale[p];
                for (int kbl = 0; kbl < KBp; ++kbl) {
                    const int b = kbg & 1;
                    mbar_wait(&d_full[b], (kbg >> 1) & 1);
                    __syncwarp();
                    tc_fence_after();
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(lane_addr + TC_DCOL0 + b * BN + c0, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[c0 + i] = fmaf(__uint_as_float(v[i]), sc, acc[c0 + i]);
                    }
                    tc_fence_before();
                    mbar_arrive(&d_empty[b]);
                    ++kbg;
                }
            }
            const long long j = (long long)jt * TC_M + r;
            if (j < J) {
                const int n = (int)(j / HWd);
                const int pix = (int)(j - (long long)n * HWd);
                const long long base = (long long)n * d_ss + pix;
                const int m0 = nt * BN;
#pragma unroll
                for (int i = 0; i < BN; ++i) {
                    const int m = m0 + i;
                    if (m < Cd) {
                        const long long o = base + (long long)m * HWd;
                        float val = acc[i];
                        if (a.bias) val += a.bias[m];
                        if (a.accumulate) val += a.out[o];
                        if (a.relu_mode == 1) val = val > 0.f ? val : 0.f;
                        else if (a.relu_mode == 2) val = a.relu_ref[o] > 0.f ? val : 0.f;
                        a.out[o] = val;
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer ==========================================================
        const uint32_t idesc = umma_idesc_tf32(TC_M, BN);
        const uint32_t b_ring = smem_u32(tc_smem);
        uint32_t kbg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            for (int p = 0; p < a.npairs; ++p) {
                for (int kbl = 0; kbl < KBp; ++kbl) {
                    const int cc = kbl % nchunks;
                    const int nvalid = min(TC_KB, Cs - cc * TC_KB);
                    const int ksteps = (nvalid + 7) >> 3;
                    const int s = kbg % TC_NST;
                    const uint32_t round = kbg / TC_NST;
                    const int b = kbg & 1;
                    const uint32_t use = kbg >> 1;
                    if (lane == 0) {
                        mbar_wait(&a_full[s], round & 1);
                        mbar_wait(&b_full[s], round & 1);
                        if (use > 0) mbar_wait(&d_empty[b], (use - 1) & 1);
                        tc_fence_after();
                        const uint32_t d_addr = tmem + TC_DCOL0 + b * BN;
                        const uint32_t a_hi = tmem + s * TC_ACOLS, a_lo = a_hi + 32;
                        const uint32_t b_hi = b_ring + s * B_STAGE_BYTES, b_lo = b_hi + B_TILE_FLOATS * 4;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t koff = ks * 256;             // 2 core matrices of 128 B per k-step
                            const uint64_t dBh = umma_desc(b_hi + koff, 128, 1024), dBl = umma_desc(b_lo + koff, 128, 1024);
                            umma_tf32_ts(d_addr, a_hi + ks * 8, dBl, idesc, ks > 0 ? 1u : 0u);    // small terms first
                            umma_tf32_ts(d_addr, a_lo + ks * 8, dBh, idesc, 1u);
                            umma_tf32_ts(d_addr, a_hi + ks * 8, dBh, idesc, 1u);
                        }
                        umma_commit(&ab_free[s]);          // arrives when these MMAs have read the stage
                        umma_commit(&d_full[b]);           // ... and when the block accumulator is complete
                    }
                    __syncwarp();
                    ++kbg;
                }
            }
        }
    } else {
        // ===================== B producer: bulk copies of the packed weight images ===================
        uint32_t kbg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile / n_jt;
            for (int p = 0; p < a.npairs; ++p) {
                const float* __restrict__ img = a.pack[p] + (long long)nt * KBp * 2 * B_TILE_FLOATS;
                for (int kbl = 0; kbl < KBp; ++kbl) {
                    const int s = kbg % TC_NST;
                    const uint32_t round = kbg / TC_NST;
                    if (lane == 0) {
                        if (round > 0) mbar_wait(&ab_free[s], (round - 1) & 1);
                        mbar_arrive_expect_tx(&b_full[s], B_STAGE_BYTES);
                        bulk_g2s(tc_smem + s * B_STAGE_BYTES, img + (long long)kbl * 2 * B_TILE_FLOATS, B_STAGE_BYTES,
                                 &b_full[s]);
                    }
                    __syncwarp();
                    ++kbg;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int BN, int MODE>
static int launch_tc_t(cudaStream_t st, const ConvKArgs& a, long long J, int Cd) {
    constexpr size_t smem = (size_t)TC_NST * 2 * BN * TC_KB * sizeof(float) + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        attr_set = true;
    }
    const long long tiles = ((J + TC_M - 1) / TC_M) * ((Cd + BN - 1) / BN);
    const unsigned grid = (unsigned)std::min<long long>(tiles, kNumSMs);
    conv_tc_kernel<BN, MODE><<<grid, TC_THREADS, smem, st>>>(a);
    return 1;
}

template <int MODE>
static int launch_tc_mode(cudaStream_t st, const ConvKArgs& a, long long J, int Cd) {
    switch (tc_choose_bn(Cd)) {
    case 16: return launch_tc_t<16, MODE>(st, a, J, Cd);
    case 32: return launch_tc_t<32, MODE>(st, a, J, Cd);
    case 48: return launch_tc_t<48, MODE>(st, a, J, Cd);
    case 64: return launch_tc_t<64, MODE>(st, a, J, Cd);
    case 96: return launch_tc_t<96, MODE>(st, a, J, Cd);
    default: return launch_tc_t<128, MODE>(st, a, J, Cd);
    }
}

int try_launch_conv_tc(int mode, cudaStream_t st, const ConvKArgs& a) {
    if (g_tc_mode == 0) return 0;
    for (int p = 0; p < a.npairs; ++p)
        if (!a.pack[p]) return 0;                     // the plan did not pack this layer (shape not eligible)
    const ConvGeom& g = a.g;
    const int Cd = mode == MODE_FWD ? g.Cout : g.Cin;
    const int Cs = mode == MODE_FWD ? g.Cin : g.Cout;
    const long long J = (long long)g.batch * (mode == MODE_FWD ? g.OH * g.OW : g.H * g.W);
    // automatic: wide layers with enough pixel tiles to be worth a tensor-core launch; narrow and
    // classifier-sized problems stay on the CUDA-core kernels
    if (g_tc_mode == 1 && (J < 1024 || Cd < 64 || Cs < 64)) return 0;
    return mode == MODE_FWD ? launch_tc_mode<MODE_FWD>(st, a, J, Cd) : launch_tc_mode<MODE_DGRAD>(st, a, J, Cd);
}

}  // namespace b2s
