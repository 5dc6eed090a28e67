// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM for the wide Conv2d / Linear contractions (sm_100a).
//
// Same contraction as conv.cu (forward and input-adjoint, up to three K-concatenated
// (activation, weight) pairs per launch) but on the 5th-generation tensor cores:
//
//   D[128 pixels x BN channels] (fp32, in TMEM)  +=  A[128 x 32] * B[BN x 32]^T   per k-block
//
// fp32 accuracy (rtol 1e-4 parity) is kept with the 3xTF32 split: every operand element
// is split while it is staged, x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi), and each
// k-step issues hi*hi + hi*lo + lo*hi (lo*lo ~ 2^-22 is dropped) into the same TMEM accumulator.
// Operands are gathered by the CTA's threads (im2col on the fly from NCHW activations, weights
// straight from the flat parameter vector), written to shared memory in the canonical K-major,
// non-swizzled UMMA layout (8 x 16 B core matrices: one warp store = one core matrix, conflict
// free), made visible to the async proxy with fence.proxy.async, and consumed by tcgen05.mma issued
// by a single thread.  Two shared-memory stages: tcgen05.commit -> mbarrier tells the stagers when
// a stage may be overwritten, so staging of k-block i+1 overlaps the MMAs of k-block i.  The
// epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 bit x 16 columns per instruction),
// adds bias / applies the ReLU mask / accumulates, and stores NCHW (lane = pixel: coalesced).
//
// TMA is not used: the operands need a gather AND an arithmetic split on the way to shared memory,
// which the copy engine cannot do; descriptor encodings follow cute/arch/mma_sm100_desc.hpp.
#include "conv_args.h"

namespace b2s {

static int g_tc_mode = 1;
void set_tc_mode(int mode) { g_tc_mode = mode; }
int get_tc_mode() { return g_tc_mode; }

constexpr int TC_BM = 128;      // pixels per CTA (UMMA M, cta_group::1)
constexpr int TC_BK = 32;       // k per stage = 4 UMMA k-steps of 8 (tf32)
constexpr int TC_THREADS = 256;

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

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <int BN, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvKArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int A_FLOATS = TC_BM * TC_BK;            // 4096 floats = 16 KB
    constexpr int B_FLOATS = BN * TC_BK;
    constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    float* stage_base = reinterpret_cast<float*>(tc_smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + 2 * STAGE_FLOATS * sizeof(float));   // [0,1] stage free, [2] done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    int* pix_yx = reinterpret_cast<int*>(tmem_slot + 4);                        // [128] (y << 16) | x, -1 past the end
    long long* pix_base = reinterpret_cast<long long*>(pix_yx + TC_BM);         // [128] sample base offset in the source

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
    const int Ktot = Cs * KHW;
    const int HWd = Hd * Wd, HWs = Hs * Ws;
    const long long J = (long long)g.batch * HWd;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long j0 = (long long)blockIdx.x * TC_BM;
    const int m0 = blockIdx.y * BN;

    // ---- one-time setup: barriers, TMEM, pixel table ---------------------------------------------
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < TC_BM) {
        const long long j = j0 + tid;
        if (j < J) {
            const int n = (int)(j / HWd);
            const int pix = (int)(j - (long long)n * HWd);
            const int y = pix / Wd, x = pix - y * Wd;
            pix_yx[tid] = (y << 16) | x;
            pix_base[tid] = (long long)n * s_ss;
        } else {
            pix_yx[tid] = -1;
            pix_base[tid] = 0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    // staging roles: warp w owns the k-group kg = w (4 consecutive k); lane = 4 * (row % 8) + (k % 4)
    const int kk = lane & 3, rr = lane >> 2;
    const int kblocks_per_pair = (Ktot + TC_BK - 1) / TC_BK;
    const int total_kb = a.npairs * kblocks_per_pair;
    const uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
    uint32_t free_phase[2] = {0, 0};

    for (int kb = 0; kb < total_kb; ++kb) {
        const int s = kb & 1;
        const int p = kb / kblocks_per_pair;
        const int k0 = (kb - p * kblocks_per_pair) * TC_BK;
        float* Ahi = stage_base + s * STAGE_FLOATS;
        float* Alo = Ahi + A_FLOATS;
        float* Bhi = Alo + A_FLOATS;
        float* Blo = Bhi + B_FLOATS;
        if (kb >= 2) {                       // the MMAs that read this stage (k-block kb-2) must have retired
            mbar_wait(&bars[s], free_phase[s]);
            free_phase[s] ^= 1;
        }
        // ---- stage A: gathered activations, core matrix (rg, kg=warp) at float offset (rg*8 + warp)*32 + lane
        const int k = k0 + warp * 4 + kk;
        const bool k_ok = k < Ktot;
        int c = 0, ky = 0, kx = 0;
        if (k_ok) {
            if (KHW == 1) c = k;
            else { c = k / KHW; const int t = k - c * KHW; ky = t / g.KW; kx = t - ky * g.KW; }
        }
        const float* __restrict__ src = a.act[p] + (long long)c * HWs;
#pragma unroll 4
        for (int rg = 0; rg < TC_BM / 8; ++rg) {
            const int r = rg * 8 + rr;
            const int yx = pix_yx[r];
            float v = 0.f;
            if (k_ok && yx >= 0) {
                const int y = yx >> 16, x = yx & 0xffff;
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
                if (ok) v = src[pix_base[r] + (long long)sy * Ws + sx];
            }
            const float hi = to_tf32(v);
            const float lo = to_tf32(v - hi);
            const int o = (rg * 8 + warp) * 32 + lane;
            Ahi[o] = hi;
            Alo[o] = lo;
        }
        // ---- stage B: weights (rows = destination channels), same core-matrix mapping
        const float* __restrict__ wt = a.wt[p];
        const float sc = a.scale[p];
#pragma unroll 4
        for (int rg = 0; rg < BN / 8; ++rg) {
            const int m = m0 + rg * 8 + rr;
            float v = 0.f;
            if (k_ok && m < Cd) {
                if (MODE == MODE_FWD) v = wt[(long long)m * Ktot + k];
                else v = wt[((long long)c * g.Cin + m) * KHW + (ky * g.KW + kx)];
                v *= sc;
            }
            const float hi = to_tf32(v);
            const float lo = to_tf32(v - hi);
            const int o = (rg * 8 + warp) * 32 + lane;
            Bhi[o] = hi;
            Blo[o] = lo;
        }
        // generic-proxy writes -> visible to the tensor core (async proxy), then hand over
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
                const uint32_t koff = ks * 256;                 // 2 core matrices of 128 B per k-step
                const uint64_t dAh = umma_desc(a_hi + koff, 128, 1024), dAl = umma_desc(a_lo + koff, 128, 1024);
                const uint64_t dBh = umma_desc(b_hi + koff, 128, 1024), dBl = umma_desc(b_lo + koff, 128, 1024);
                umma_tf32(tmem_d, dAh, dBh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                umma_tf32(tmem_d, dAl, dBh, idesc, 1u);
            }
            umma_commit(&bars[s]);                              // arrives when these MMAs have read the stage
            if (kb == total_kb - 1) umma_commit(&bars[2]);      // ... and when the accumulator is complete
        }
    }

    // ---- epilogue: TMEM -> registers -> NCHW global ------------------------------------------------
    mbar_wait(&bars[2], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int q = warp & 3;                                  // TMEM lane quarter this warp may read
        const int half = warp >> 2;                              // column half
        const int r = q * 32 + lane;
        const long long j = j0 + r;
        const bool ok = j < J;
        int n = 0, pix = 0;
        if (ok) { n = (int)(j / HWd); pix = (int)(j - (long long)n * HWd); }
        const long long base = (long long)n * d_ss + pix;
#pragma unroll 1
        for (int cc = 0; cc < BN / 2; cc += 16) {
            const int col0 = half * (BN / 2) + cc;
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int m = m0 + col0 + i;
                    if (m < Cd) {
                        const long long o = base + (long long)m * HWd;
                        float val = __uint_as_float(v[i]);
                        if (a.bias) val += a.bias[m];
                        if (a.accumulate) val += a.out[o];
                        if (a.relu_mode == 1) val = val > 0.f ? val : 0.f;
                        else if (a.relu_mode == 2) val = a.relu_ref[o] > 0.f ? val : 0.f;
                        a.out[o] = val;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(BN) : "memory");
    }
}

template <int BN, int MODE>
static int launch_tc_t(cudaStream_t st, const ConvKArgs& a, long long J, int Cd) {
    constexpr size_t smem = 2 * (2 * TC_BM * TC_BK + 2 * BN * TC_BK) * sizeof(float) + 4 * sizeof(uint64_t) +
                            4 * sizeof(uint32_t) + TC_BM * sizeof(int) + TC_BM * sizeof(long long) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        attr_set = true;
    }
    dim3 grid((unsigned)((J + TC_BM - 1) / TC_BM), (unsigned)((Cd + BN - 1) / BN));
    conv_tc_kernel<BN, MODE><<<grid, TC_THREADS, smem, st>>>(a);
    return 1;
}

int try_launch_conv_tc(int mode, cudaStream_t st, const ConvKArgs& a) {
    if (g_tc_mode == 0) return 0;
    const ConvGeom& g = a.g;
    const int Cd = mode == MODE_FWD ? g.Cout : g.Cin;
    const int Cs = mode == MODE_FWD ? g.Cin : g.Cout;
    const long long J = (long long)g.batch * (mode == MODE_FWD ? g.OH * g.OW : g.H * g.W);
    const int Ktot = Cs * g.KH * g.KW;
    const int Hs = mode == MODE_FWD ? g.H : g.OH, Ws = mode == MODE_FWD ? g.W : g.OW;
    if (Hs >= 65536 || Ws >= 65536) return 0;
    if (g_tc_mode == 1) {
        // automatic: wide layers with enough pixels to fill the machine; the narrow DenseNet-BC layers stay
        // on the pixel-thread kernel (their cost is the gather, not the FLOPs)
        if (Cd < 64 || Ktot < 64 || J < 128LL * 32) return 0;
    } else {
        if (Cd < 8) return 0;
    }
    if (Cd > 64) {
        return mode == MODE_FWD ? launch_tc_t<128, MODE_FWD>(st, a, J, Cd) : launch_tc_t<128, MODE_DGRAD>(st, a, J, Cd);
    }
    return mode == MODE_FWD ? launch_tc_t<64, MODE_FWD>(st, a, J, Cd) : launch_tc_t<64, MODE_DGRAD>(st, a, J, Cd);
}

}  // namespace b2s
