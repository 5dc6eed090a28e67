// tc_common.cuh -- shared device helpers of the tcgen05 / TMEM kernels (conv_tc.cu, conv_tc_wgrad.cu):
// pipeline constants, UMMA descriptors, mbarrier / tcgen05 wrappers.  Encodings follow
// cute/arch/mma_sm100_desc.hpp and cute/arch/mma_sm100_umma.hpp.
#pragma once
#include "conv_args.h"

namespace b2s {

constexpr int TC_M = 128;        // pixels per tile (UMMA M, cta_group::1)
constexpr int TC_KB = 32;        // k per k-block = 4 UMMA k-steps of 8 (tf32)
constexpr int TC_NST = 3;        // pipeline stages (A in TMEM, B in shared memory)
constexpr int TC_ACOLS = 64;     // TMEM columns per A stage: 32 hi + 32 lo
constexpr int TC_DCOL0 = TC_NST * TC_ACOLS;    // first accumulator column
constexpr int TC_DCOLS = (512 - TC_DCOL0) / 2; // columns of one of the two accumulator buffers (160)
// Consecutive tcgen05.mma into the SAME accumulator serialise on its ~90-clock read-modify-write latency
// (measured: 12 dependent MMAs take ~1100 clocks whether N is 16 or 128).  Narrow tiles therefore spread
// the three 3xTF32 terms (hi*lo, lo*hi, hi*hi) -- and for BN = 16 also odd / even k-steps -- over
// independent accumulators inside the buffer; the drain adds them up.
__host__ __device__ constexpr int tc_nacc(int BN) { return 6 * BN <= TC_DCOLS ? 6 : 3 * BN <= TC_DCOLS ? 3 : 2 * BN <= TC_DCOLS ? 2 : 1; }
constexpr int TC_THREADS = 512;  // four warpgroups: 2 x A transform, drain, {MMA issuer, B producer, 2 idle warps}
constexpr int TC_WARP_MMA = 12;
constexpr int TC_WARP_B = 13;
// registers per thread after the role split (setmaxnreg): 128 (launch) for the transform warpgroups,
// TC_REGS_DRAIN for the warpgroup that holds up to 128 accumulators per thread, TC_REGS_MISC for the rest;
// 128 * (128 + 128 + 208 + 40) <= 64 K registers
constexpr int TC_REGS_DRAIN = 208;
constexpr int TC_REGS_MISC = 40;

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
// D[tmem] (+)= A[smem] * B[smem]^T   (SM100_MMA_TF32_SS)
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// One probe of the barrier phase.  The suspend-time hint lets the hardware park the thread until the
// phase completes (or ~the hint elapses) instead of returning at once: without it the waiting roles
// (drain, MMA issuer, producers) spin on try_wait and -- measured with ncu's per-instruction counts --
// issue more than half of all instructions of the kernel, starving the transform warps of issue slots.
__device__ __forceinline__ uint32_t mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    return done;
}
// bounded wait: a protocol error must become a trap (launch failure), never a hang.  The clock is only
// read once the first probe has failed.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try(addr, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(addr, parity)) {
        if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
    }
}
// Spinning variant without the suspend hint, for the ONE thread whose wake-up latency is on the critical path of every
// k-block (the MMA issuer): NANOSLEEP.SYNCS parks the warp and the wake-up after the arrival is not immediate.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
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
// one lane of a converged warp; operands computed warp-uniformly OUTSIDE the elected branch stay in uniform
// registers, so each tcgen05.mma is a single UTCHMMA (a branch on lane == 0 makes the compiler wrap every
// MMA in an R2UR broadcast loop: ~90 clocks per instruction)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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


// one pixel's BN output channels: out[c] (+)= acc[c] with the ReLU variant; channel stride cs, mrem valid channels.
// Loads (previous value, ReLU reference) are issued eight channels ahead of the stores that need them.
template <int BN, bool ACC, int RELU>
__device__ __forceinline__ void tc_epilogue(const float (&acc)[BN], float* __restrict__ outp, const float* __restrict__ refp,
                                            const int cs, const int mrem) {
    if (mrem >= BN) {                             // full tile: no per-channel predicates at all
#pragma unroll
        for (int i0 = 0; i0 < BN; i0 += 8) {
            float prev[8], ref[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (ACC) prev[i] = outp[(long long)(i0 + i) * cs];
                if (RELU == 2) ref[i] = __ldg(refp + (long long)(i0 + i) * cs);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float val = acc[i0 + i];
                if (ACC) val += prev[i];
                if (RELU == 1) val = fmaxf(val, 0.f);
                if (RELU == 2) val = ref[i] > 0.f ? val : 0.f;
                outp[(long long)(i0 + i) * cs] = val;
            }
        }
    } else {
#pragma unroll
        for (int i0 = 0; i0 < BN; i0 += 8) {
            float prev[8], ref[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool okc = i0 + i < mrem;
                if (ACC) prev[i] = okc ? outp[(long long)(i0 + i) * cs] : 0.f;
                if (RELU == 2) ref[i] = okc ? __ldg(refp + (long long)(i0 + i) * cs) : 1.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i0 + i < mrem) {
                    float val = acc[i0 + i];
                    if (ACC) val += prev[i];
                    if (RELU == 1) val = fmaxf(val, 0.f);
                    if (RELU == 2) val = ref[i] > 0.f ? val : 0.f;
                    outp[(long long)(i0 + i) * cs] = val;
                }
            }
        }
    }
}

// picks the variant ONCE per tile: the generic form (flags tested per element) compiled to ~20 instructions and a branch
// per channel, executed by one warp per scheduler -- 6600 clocks per 48-channel tile in the timeline of conv_tma_kernel
template <int BN>
__device__ __forceinline__ void tc_store_tile(float (&acc)[BN], const ConvKArgs& a, const long long off, const int cs, const int m0,
                                              const int mrem) {
    if (a.bias) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
            if (i < mrem) acc[i] += __ldg(a.bias + m0 + i);
    }
    const int variant = (a.accumulate != 0 ? 1 : 0) + 2 * a.relu_mode;
    switch (variant) {
    case 0: tc_epilogue<BN, false, 0>(acc, a.out + off, nullptr, cs, mrem); break;
    case 1: tc_epilogue<BN, true, 0>(acc, a.out + off, nullptr, cs, mrem); break;
    case 2: tc_epilogue<BN, false, 1>(acc, a.out + off, nullptr, cs, mrem); break;
    case 3: tc_epilogue<BN, true, 1>(acc, a.out + off, nullptr, cs, mrem); break;
    case 4: tc_epilogue<BN, false, 2>(acc, a.out + off, a.relu_ref + off, cs, mrem); break;
    default: tc_epilogue<BN, true, 2>(acc, a.out + off, a.relu_ref + off, cs, mrem); break;
    }
}


// Split-K form: this CTA holds a PARTIAL tile; partial tiles are summed with fp32 atomics into an output that was
// zeroed beforehand (overwrite mode) or already holds the value to accumulate onto.  The ReLU mask of relu_mode 2 is
// multiplicative, so it is applied to every partial; the bias is added by the slice that starts at pair-tap 0.
template <int BN>
__device__ __forceinline__ void tc_store_tile_split(float (&acc)[BN], const ConvKArgs& a, const long long off, const int cs, const int m0,
                                                    const int mrem, const bool first_slice) {
    float* __restrict__ outp = a.out + off;
    const float* __restrict__ refp = a.relu_mode == 2 ? a.relu_ref + off : nullptr;
#pragma unroll
    for (int i0 = 0; i0 < BN; i0 += 8) {
        float ref[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ref[i] = (refp && i0 + i < mrem) ? __ldg(refp + (long long)(i0 + i) * cs) : 1.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i0 + i < mrem) {
                float val = acc[i0 + i];
                if (first_slice && a.bias) val += __ldg(a.bias + m0 + i0 + i);
                if (ref[i] > 0.f) atomicAdd(outp + (long long)(i0 + i) * cs, val);
            }
        }
    }
}

}  // namespace b2s
