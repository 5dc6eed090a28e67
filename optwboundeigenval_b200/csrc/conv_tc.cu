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
//   warps 0-7  A transform (two warpgroups on alternate k-blocks, loads of the next k-block in flight
//              while the current one is converted): lane = pixel (coalesced NCHW reads, im2col on the fly
//              with zero padding), split hi/lo in registers, tcgen05.st straight into TMEM -- the A operand
//              never touches shared memory (tcgen05.mma with A in TMEM), no bank conflicts, no proxy fence;
//   warp  13   B producer: weights are pre-split and pre-arranged ONCE per pass by tc_pack_kernel into
//              the exact shared-memory image (K-major, SWIZZLE_NONE core matrices) of every k-block,
//              so a stage is one cp.async.bulk (bulk-copy engine, mbarrier complete_tx);
//   warp  12   MMA issuer: one thread, 3 x tcgen05.mma.kind::tf32 per k-step, tcgen05.commit to the
//              stage-free and accumulator-full mbarriers;
//   warps 8-11 drain + epilogue (setmaxnreg gives this warpgroup 208 registers): tcgen05.ld the block accumulator, acc += scale * d, and after the last
//              k-block bias / ReLU mask / accumulate and coalesced NCHW stores (lane = pixel).
// Four stages of A (TMEM) and B (smem); two accumulator buffers.  Descriptor encodings follow
// cute/arch/mma_sm100_desc.hpp and cute/arch/mma_sm100_umma.hpp (SM100_MMA_TF32_TS).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "conv_args.h"
#include "tc_common.cuh"

namespace b2s {

// process-wide switch; the environment variable B2S_TC_MODE (0/1/2) overrides the default for debugging
static int initial_tc_mode() {
    const char* e = getenv("B2S_TC_MODE");
    return (e && e[0] >= '0' && e[0] <= '2' && e[1] == 0) ? e[0] - '0' : 1;
}
static int g_tc_mode = initial_tc_mode();
void set_tc_mode(int mode) { g_tc_mode = mode; }
int get_tc_mode() { return g_tc_mode; }

// ---------------------------------------------------------------------------------------------
// weight packing: flat parameter vector -> per (op, mode) images of every k-block's B tile
// ---------------------------------------------------------------------------------------------
// k ordering of a pair: k-block kbl = tap * nchunks + chunk, element k inside = source channel
// chunk * 32 + k (zero padded), so that one k-block of the A gather is 32 channels of ONE tap
// (one shifted pixel per thread, channel stride = plane size).
// image[(ntile * KBp + kbl) * 2 + {hi, lo}][BN x 32 tile], tile element (m, k) at float offset
// ((m / 8) * 8 + k / 4) * 32 + (m % 8) * 4 + k % 4   (8 x 16 B core matrices, LBO 128 B, SBO 1024 B)
// One thread = four consecutive k of one tile row: the job lookup and the index arithmetic are paid once per 16-byte store
// (the first version did a binary search, four divisions and two scalar stores per ELEMENT: 0.35 ms per pass for VGG16's
// 19.4 M weights, 1.3 TB/s).
__global__ void __launch_bounds__(256) tc_pack_kernel(const TcPackJob* __restrict__ jobs, int njobs, long long total,
                                                      const float* __restrict__ src_base, float* __restrict__ dst_base) {
    const long long groups = total >> 2;                             // every image is a multiple of 32 elements
    for (long long e4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; e4 < groups;
         e4 += (long long)gridDim.x * blockDim.x) {
        const long long e = e4 << 2;
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].begin <= e) lo = mid; else hi = mid - 1;
        }
        const TcPackJob jb = jobs[lo];
        unsigned l = (unsigned)(e - jb.begin);                    // one image has < 2^32 elements
        const int k = (int)(l % TC_KB); l /= TC_KB;               // multiple of 4
        const int m = (int)(l % (unsigned)jb.BN); l /= (unsigned)jb.BN;
        const int KBp = jb.KHW * jb.nchunks;
        const int kbl = (int)(l % (unsigned)KBp);
        const int nt = (int)(l / (unsigned)KBp);
        const int t = kbl / jb.nchunks, cc = kbl - t * jb.nchunks;
        const int c = cc * TC_KB + k;
        const int mm = nt * jb.BN + m;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (mm < jb.Cd) {
            const float* w = src_base + jb.src_off;
            // forward: [Cout = Cd][Cin = Cs][tap]; input adjoint: [Cout = Cs][Cin = Cd][tap]
            const long long base = jb.mode == MODE_FWD ? ((long long)mm * jb.Cs + c) * jb.KHW + t : ((long long)c * jb.Cd + mm) * jb.KHW + t;
            const long long step = jb.mode == MODE_FWD ? (long long)jb.KHW : (long long)jb.Cd * jb.KHW;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (c + i < jb.Cs) v[i] = __ldg(w + base + i * step);
        }
        float4 h4, l4;
        uint32_t h;
        h = to_tf32_bits(v[0]); h4.x = __uint_as_float(h); l4.x = __uint_as_float(to_tf32_bits(v[0] - h4.x));
        h = to_tf32_bits(v[1]); h4.y = __uint_as_float(h); l4.y = __uint_as_float(to_tf32_bits(v[1] - h4.y));
        h = to_tf32_bits(v[2]); h4.z = __uint_as_float(h); l4.z = __uint_as_float(to_tf32_bits(v[2] - h4.z));
        h = to_tf32_bits(v[3]); h4.w = __uint_as_float(h); l4.w = __uint_as_float(to_tf32_bits(v[3] - h4.w));
        float* tile = dst_base + jb.dst_off + ((long long)nt * KBp + kbl) * 2 * jb.BN * TC_KB;
        const int o = ((m >> 3) * 8 + (k >> 2)) * 32 + (m & 7) * 4;
        *reinterpret_cast<float4*>(tile + o) = h4;
        *reinterpret_cast<float4*>(tile + jb.BN * TC_KB + o) = l4;
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

bool tc_splitk_allowed();

bool tc_shape_ok(int Cs, int Cd, int Hs, int Ws) {
    return Cs >= 8 && Cd >= 8 && Hs < 32768 && Ws < 32768;
}

int launch_tc_pack(cudaStream_t st, const TcPackJob* d_jobs, int njobs, long long total, const float* src_base,
                   float* dst_base) {
    if (njobs <= 0 || total <= 0) return 0;
    ProfScope prof("tc_pack", 0.0, 12.0 * (double)total, st);
    long long blocks = ((total >> 2) + 255) / 256;
    if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
    tc_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_jobs, njobs, total, src_base, dst_base);
    B2S_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the contraction
// ---------------------------------------------------------------------------------------------
// Walks the k-blocks of this CTA in issue order: tile (persistent stride), pair, tap, channel chunk.
// The A-transform warps keep the per-pixel gather state here (one thread = one pixel row of the tile).
// lanes whose tap falls outside the image read this instead (stride 0): their loads need no predicate
__device__ const float tc_zero_words[4] = {0.f, 0.f, 0.f, 0.f};

// Virtual tile -> (pixel tile, channel tile, k-slice).  Split-K (a.k_chunk = number of slices > 1) cuts the pair-tap
// range [0, npairs * KH * KW) of a tile into slices handled by different CTAs, so that layers with few pixel tiles
// (VGG16 conv5_x at batch 4: 28 tiles for 148 SMs, 288 k-blocks each) fill the machine; the partial tiles are added
// with fp32 atomics (tc_store_tile_split).
__device__ __forceinline__ void tc_decode_tile(const ConvKArgs& a, const int n_jt, const int vt, const int Cd_unused,
                                               int& jt, int& nt, int& pt0, int& pt1) {
    (void)Cd_unused;
    const int PT = a.npairs * a.g.KH * a.g.KW;
    const int ksplit = a.k_chunk > 1 ? a.k_chunk : 1;
    const int real = vt / ksplit, ks = vt - real * ksplit;       // slices of a tile are neighbours: they share the A tile in L2
    jt = real % n_jt;
    nt = real / n_jt;
    const int per = (PT + ksplit - 1) / ksplit;
    pt0 = min(PT, ks * per);
    pt1 = min(PT, pt0 + per);
}

template <int MODE>
struct TcCursor {
    // geometry is re-read from the kernel parameters (constant bank) instead of being cached in
    // registers: this role keeps two 32-value load buffers live
    const ConvKArgs& a;
    int r;                 // pixel row of this thread inside the tile
    int tile, p, ky, kx, cc;   // tile = VIRTUAL tile: (pixel tile, channel tile, k-slice), see tc_decode_tile
    int pt, pt_end;            // pair-tap index (p * KHW + tap) and the end of this k-slice
    uint32_t kbg;
    int y, x;              // destination pixel (y < 0: past the end of the pixel range)
    long long nbase;       // sample offset in the source
    const float* src;      // channel 0 of the current (pair, tap) at this thread's pixel, or the zero word
    long long cstride;     // elements between consecutive source channels for this thread (0 on the zero word)

    __device__ __forceinline__ int Hd() const { return MODE == MODE_FWD ? a.g.OH : a.g.H; }
    __device__ __forceinline__ int Wd() const { return MODE == MODE_FWD ? a.g.OW : a.g.W; }
    __device__ __forceinline__ int Cs() const { return MODE == MODE_FWD ? a.g.Cin : a.g.Cout; }
    __device__ __forceinline__ int Hs() const { return MODE == MODE_FWD ? a.g.H : a.g.OH; }
    __device__ __forceinline__ int Ws() const { return MODE == MODE_FWD ? a.g.W : a.g.OW; }
    __device__ __forceinline__ long long s_ss() const { return MODE == MODE_FWD ? a.g.in_sstride : a.g.out_sstride; }
    __device__ __forceinline__ long long J() const { return (long long)a.g.batch * (Hd() * Wd()); }
    __device__ __forceinline__ int n_jt() const { return (int)((J() + TC_M - 1) / TC_M); }
    __device__ __forceinline__ int nchunks() const { return (Cs() + TC_KB - 1) / TC_KB; }

    __device__ __forceinline__ TcCursor(const ConvKArgs& a_, int r_) : a(a_), r(r_) {}
    __device__ __forceinline__ void start(int total_tiles) {
        tile = blockIdx.x; p = 0; ky = 0; kx = 0; cc = 0; kbg = 0; pt = 0; pt_end = 0;
        if (tile < total_tiles) { set_tile(); set_tap(); }
    }
    __device__ __forceinline__ void set_tile() {
        int jt_, nt_, pt0_, pt1_;
        tc_decode_tile(a, n_jt(), tile, MODE == MODE_FWD ? a.g.Cout : a.g.Cin, jt_, nt_, pt0_, pt1_);
        pt = pt0_; pt_end = pt1_;
        const int jt = jt_;
        const long long j = (long long)jt * TC_M + r;
        const int HWd = Hd() * Wd();
        y = -1; x = 0; nbase = 0;
        if (j < J()) {
            const int n = (int)(j / HWd);
            const int pix = (int)(j - (long long)n * HWd);
            y = pix / Wd();
            x = pix - y * Wd();
            nbase = (long long)n * s_ss();
        }
    }
    __device__ __forceinline__ void set_tap() {
        const ConvGeom& g = a.g;
        {
            const int KHW = g.KH * g.KW;
            p = pt / KHW;
            const int t = pt - p * KHW;
            ky = t / g.KW;
            kx = t - ky * g.KW;
        }
        int sy, sx;
        bool ok;
        if (MODE == MODE_FWD) {
            sy = y * g.sh + ky - g.ph; sx = x * g.sw + kx - g.pw;
            ok = sy >= 0 && sy < Hs() && sx >= 0 && sx < Ws();
        } else {
            const int ty_ = y + g.ph - ky, tx_ = x + g.pw - kx;
            sy = ty_ / g.sh; sx = tx_ / g.sw;
            ok = ty_ >= 0 && tx_ >= 0 && sy * g.sh == ty_ && sx * g.sw == tx_ && sy < Hs() && sx < Ws();
        }
        ok = ok && y >= 0;
        src = ok ? a.act[p] + nbase + (sy * Ws() + sx) : tc_zero_words;
        cstride = ok ? (long long)Hs() * Ws() : 0;
    }
    __device__ __forceinline__ int nvalid() const { return min(TC_KB, Cs() - cc * TC_KB); }
    // returns false when the CTA has no further k-block
    __device__ __forceinline__ bool advance(int total_tiles) {
        ++kbg;
        if (++cc < nchunks()) return true;
        cc = 0;
        if (++pt == pt_end) {
            tile += gridDim.x;
            if (tile >= total_tiles) return false;
            set_tile();
        }
        set_tap();
        return true;
    }
    // the 32 (zero padded) source channels of the current k-block: all loads independent, unpredicated
    // (out-of-image lanes read the zero word), addresses by pointer increments; only a ragged last group
    // of 8 channels pays per-channel selects
    __device__ __forceinline__ void load(float (&v)[TC_KB]) const {
        const int nv = nvalid();
        const float* __restrict__ ptr = src + (long long)cc * TC_KB * cstride;
#pragma unroll
        for (int g8 = 0; g8 < TC_KB / 8; ++g8) {
            if (g8 * 8 + 8 <= nv) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[g8 * 8 + i] = __ldg(ptr); ptr += cstride; }
            } else if (g8 * 8 < nv) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool in = g8 * 8 + i < nv;
                    v[g8 * 8 + i] = in ? __ldg(ptr) : 0.f;
                    if (in) ptr += cstride;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[g8 * 8 + i] = 0.f;
            }
        }
    }
};

// debug timeline: trace[(role * TC_TRACE_KB + kbg) * 4 + slot] = clock64(), CTA 0, first TC_TRACE_KB k-blocks
constexpr int TC_TRACE_KB = 48;
// (compiled in only with -DB2S_TC_TRACE_ENABLED: `B2S_BUILD_TRACE=1 python -m optwboundeigenval_b200.build --force`)
__device__ __forceinline__ void tc_stamp(long long* trace, int role, uint32_t kbg, int slot) {
#ifdef B2S_TC_TRACE_ENABLED
    if (trace && blockIdx.x == 0 && kbg < TC_TRACE_KB) trace[((size_t)role * TC_TRACE_KB + kbg) * 4 + slot] = clock64();
#else
    (void)trace; (void)role; (void)kbg; (void)slot;
#endif
}

// split one k-block of A values and hand it to the tensor core: wait for the TMEM stage, tcgen05.st
// hi / lo, publish on a_full
__device__ __forceinline__ void tc_commit_a(const float (&v)[TC_KB], uint32_t kbg, int ksteps, uint32_t lane_addr,
                                            uint64_t* a_full, uint64_t* ab_free, long long* trace) {
    const int s = kbg % TC_NST;
    const uint32_t round = kbg / TC_NST;
    const bool tr = (threadIdx.x & 127) == 0;
    if (tr) tc_stamp(trace, 0, kbg, 0);
    if (round > 0) mbar_wait(&ab_free[s], (round - 1) & 1);
    __syncwarp();
    tc_fence_after();
    if (tr) tc_stamp(trace, 0, kbg, 1);
    const uint32_t col = (uint32_t)(s * TC_ACOLS);
#pragma unroll
    for (int ks = 0; ks < TC_KB / 8; ++ks) {
        if (ks < ksteps) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // hi = round-to-nearest TF32; lo = f - hi is exact (<= 12 significant bits) and is cut to
                // TF32 by clearing its low 13 bits (2^-22 relative to f; one conversion-pipe op per element)
                const float f = v[ks * 8 + i];
                hi[i] = to_tf32_bits(f);
                lo[i] = __float_as_uint(f - __uint_as_float(hi[i])) & 0xffffe000u;
            }
            tmem_st8(lane_addr + col + ks * 8, hi);
            tmem_st8(lane_addr + col + 32 + ks * 8, lo);
        }
    }
    if (tr) tc_stamp(trace, 0, kbg, 2);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    mbar_arrive(&a_full[s]);
    if (tr) tc_stamp(trace, 0, kbg, 3);
}

template <int BN, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const ConvKArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    constexpr int B_TILE_FLOATS = BN * TC_KB;                 // one of hi / lo
    constexpr uint32_t B_STAGE_BYTES = 2u * B_TILE_FLOATS * sizeof(float);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + TC_NST * B_STAGE_BYTES);
    uint64_t* a_full = bars;                   // [NST] operands landed: the 128 transform threads of the owning group
                                               //       + the B producer's arrive.expect_tx and the bulk copy's bytes
    uint64_t* b_full = a_full;                 // (same barrier)
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
    const int KHW = g.KH * g.KW;
    const int HWd = Hd * Wd;
    const long long J = (long long)g.batch * HWd;
    const int nchunks = (Cs + TC_KB - 1) / TC_KB;
    const int KBp = KHW * nchunks;                            // k-blocks per pair
    const int n_jt = (int)((J + TC_M - 1) / TC_M);
    const int n_nt = (Cd + BN - 1) / BN;
    const int ksplit = a.k_chunk > 1 ? a.k_chunk : 1;
    const int total_tiles = n_jt * n_nt * ksplit;               // virtual tiles (tc_decode_tile)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
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
    if (warp == TC_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ===================== A transform: global -> registers (split) -> TMEM =====================
        // two warpgroups take alternate k-blocks; each keeps the loads of its NEXT k-block in flight while
        // it converts and stores the current one (two register buffers)
        const int grp = warp >> 2, q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        TcCursor<MODE> c(a, q * 32 + lane);
        c.start(total_tiles);
        bool more = c.tile < total_tiles;
        if (grp == 1 && more) more = c.advance(total_tiles);
        float va[TC_KB], vb[TC_KB];
        uint32_t ka = 0, kb = 0;
        int sa = 0, sb = 0;
        // fetch(): issue the loads of this group's next k-block and step the cursor two k-blocks on
#define TC_FETCH(V, KI, SI)                                           \
        do {                                                          \
            if ((tid & 127) == 0) tc_stamp(a.trace, 3, c.kbg, 0);      \
            c.load(V); KI = c.kbg; SI = (c.nvalid() + 7) >> 3;         \
            if ((tid & 127) == 0) tc_stamp(a.trace, 3, KI, 1);         \
            more = c.advance(total_tiles);                            \
            if (more) more = c.advance(total_tiles);                  \
            if ((tid & 127) == 0) tc_stamp(a.trace, 3, KI, 2);         \
        } while (0)
        bool ha = more;
        if (ha) TC_FETCH(va, ka, sa);
        while (ha) {
            const bool hb = more;
            if (hb) TC_FETCH(vb, kb, sb);
            tc_commit_a(va, ka, sa, lane_addr, a_full, ab_free, a.trace);
            if (!hb) break;
            ha = more;
            if (ha) TC_FETCH(va, ka, sa);
            tc_commit_a(vb, kb, sb, lane_addr, a_full, ab_free, a.trace);
        }
#undef TC_FETCH
    } else if (warp < 12) {
        // ===================== drain + epilogue: TMEM -> fp32 registers -> NCHW global ===============
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_DRAIN));
        const int q = warp - 8;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t kbg = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int jt, nt, pt0, pt1;
            tc_decode_tile(a, n_jt, tile, Cd, jt, nt, pt0, pt1);
            float acc[BN];
#pragma unroll
            for (int i = 0; i < BN; ++i) acc[i] = 0.f;
            constexpr int NACC = tc_nacc(BN);
            const int last_ksteps = (Cs - (nchunks - 1) * TC_KB + 7) >> 3;
            for (int pt = pt0; pt < pt1; ++pt) {
                const float sc = a.scale[pt / KHW];
                for (int kbl = 0; kbl < nchunks; ++kbl) {
                    const int b = kbg & 1;
                    if (tid == 256) tc_stamp(a.trace, 2, kbg, 0);
                    mbar_wait(&d_full[b], (kbg >> 1) & 1);
                    __syncwarp();
                    tc_fence_after();
                    if (tid == 256) tc_stamp(a.trace, 2, kbg, 1);
                    // with 6 accumulators the odd-k-step set is only written by k-blocks of >= 2 k-steps
                    const bool odd_set = NACC == 6 && (((kbl % nchunks) == nchunks - 1 ? last_ksteps : TC_KB / 8) >= 2);
                    const uint32_t d0 = lane_addr + TC_DCOL0 + b * TC_DCOLS;
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 16) {
                        uint32_t v[NACC >= 3 ? 3 : NACC][16];
#pragma unroll
                        for (int q2 = 0; q2 < (NACC >= 3 ? 3 : NACC); ++q2) tmem_ld16(d0 + q2 * BN + c0, v[q2]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float d = __uint_as_float(v[0][i]);
                            if (NACC >= 2) d += __uint_as_float(v[1][i]);
                            if (NACC >= 3) d += __uint_as_float(v[2][i]);
                            acc[c0 + i] = fmaf(d, sc, acc[c0 + i]);
                        }
                        if (NACC == 6 && odd_set) {
                            uint32_t w[3][16];
#pragma unroll
                            for (int q2 = 0; q2 < 3; ++q2) tmem_ld16(d0 + (3 + q2) * BN + c0, w[q2]);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                acc[c0 + i] = fmaf(__uint_as_float(w[0][i]) + __uint_as_float(w[1][i]) + __uint_as_float(w[2][i]), sc, acc[c0 + i]);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&d_empty[b]);
                    if (tid == 256) tc_stamp(a.trace, 2, kbg, 2);
                    ++kbg;
                }
            }
            // epilogue: lane = pixel, so every per-channel store is one coalesced 128-byte row.  Reads the
            // epilogue depends on (previous adjoint, ReLU reference) are issued in batches of 8 ahead of
            // the stores so that they overlap instead of forming a load -> store chain.
            const long long j = (long long)jt * TC_M + r;
            if (j < J) {
                const int n = (int)(j / HWd);
                const int pix = (int)(j - (long long)n * HWd);
                const int m0 = nt * BN;
                const long long off = (long long)n * d_ss + pix + (long long)m0 * HWd;
                const int mrem = Cd - m0;                 // valid channels of this tile
                if (ksplit > 1) tc_store_tile_split<BN>(acc, a, off, HWd, m0, mrem, pt0 == 0);
                else tc_store_tile<BN>(acc, a, off, HWd, m0, mrem);
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_MISC));
        if (warp == TC_WARP_MMA) {
            // ===================== MMA issuer ======================================================
            // one thread; everything loop-invariant is hoisted, the per-k-block path is two barrier probes,
            // <= 12 MMAs and two commits
            const uint32_t idesc = umma_idesc_tf32(TC_M, BN);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);           // provably warp-uniform
            const uint64_t desc0 = umma_desc(smem_u32(tc_smem), 128, 1024);      // stage 0, hi tile, k-step 0
            const int last_ksteps = (Cs - (nchunks - 1) * TC_KB + 7) >> 3;
            constexpr int NACC = tc_nacc(BN);
            uint32_t kbg = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int jt_, nt_, pt0, pt1;
                tc_decode_tile(a, n_jt, tile, Cd, jt_, nt_, pt0, pt1);
                for (int pt = pt0; pt < pt1; ++pt) {
                    for (int cc = 0; cc < nchunks; ++cc) {
                        const int ksteps = cc == nchunks - 1 ? last_ksteps : TC_KB / 8;
                        const int s = kbg % TC_NST;
                        const uint32_t round = kbg / TC_NST;
                        const int b = kbg & 1;
                        const uint32_t use = kbg >> 1;
                        if (lane == 0) tc_stamp(a.trace, 1, kbg, 0);
                        mbar_wait(&a_full[s], round & 1);
                        if (lane == 0) tc_stamp(a.trace, 1, kbg, 1);
                        if (use > 0) mbar_wait(&d_empty[b], (use - 1) & 1);
                        __syncwarp();
                        tc_fence_after();
                        if (lane == 0) tc_stamp(a.trace, 1, kbg, 2);
                        const uint32_t d_addr = tmem_u + TC_DCOL0 + b * TC_DCOLS;
                        const uint32_t a_hi = tmem_u + s * TC_ACOLS, a_lo = a_hi + 32;
                        const uint64_t dBh = desc0 + (uint64_t)((s * B_STAGE_BYTES) >> 4);
                        const uint64_t dBl = dBh + (uint64_t)((B_TILE_FLOATS * 4) >> 4);
                        if (elect_one()) {
                            // accumulator of (term, k-step): terms 0 = hi*lo, 1 = lo*hi (small, added first), 2 = hi*hi
                            auto acc_of = [](int term, int ks) {
                                return NACC == 6 ? term + 3 * (ks & 1) : NACC == 3 ? term : NACC == 2 ? (term == 2 ? 1 : 0) : 0;
                            };
#pragma unroll
                            for (int ks = 0; ks < TC_KB / 8; ++ks) {
                                if (ks < ksteps) {
                                    const uint64_t ko = (uint64_t)(ks * 16);      // 2 core matrices of 128 B per k-step
                                    const uint32_t fresh = NACC == 6 ? (ks >= 2) : (ks >= 1);   // accumulate onto this k-block's earlier k-steps
                                    umma_tf32_ts(d_addr + acc_of(0, ks) * BN, a_hi + ks * 8, dBl + ko, idesc, fresh);
                                    umma_tf32_ts(d_addr + acc_of(1, ks) * BN, a_lo + ks * 8, dBh + ko, idesc, NACC <= 2 ? 1u : fresh);
                                    umma_tf32_ts(d_addr + acc_of(2, ks) * BN, a_hi + ks * 8, dBh + ko, idesc, NACC == 1 ? 1u : fresh);
                                }
                            }
                            umma_commit(&ab_free[s]);      // arrives when these MMAs have read the stage
                            umma_commit(&d_full[b]);       // ... and when the block accumulator is complete
                        }
                        __syncwarp();
                        if (lane == 0) tc_stamp(a.trace, 1, kbg, 3);
                        ++kbg;
                    }
                }
            }
        } else if (warp == TC_WARP_B) {
            // ===================== B producer: bulk copies of the packed weight images ===============
            uint32_t kbg = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int jt_, nt, pt0, pt1;
                tc_decode_tile(a, n_jt, tile, Cd, jt_, nt, pt0, pt1);
                for (int pt = pt0; pt < pt1; ++pt) {
                    const int p = pt / KHW;
                    const float* __restrict__ img = a.pack[p] + (long long)nt * KBp * 2 * B_TILE_FLOATS;
                    for (int kbl = (pt - p * KHW) * nchunks; kbl < (pt - p * KHW + 1) * nchunks; ++kbl) {
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int BN, int MODE>
static int launch_tc_t(cudaStream_t st, const ConvKArgs& a_in, long long J, int Cd) {
    constexpr size_t smem = (size_t)TC_NST * 2 * BN * TC_KB * sizeof(float) + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        attr_set = true;
    }
    long long tiles = ((J + TC_M - 1) / TC_M) * ((Cd + BN - 1) / BN);
    ConvKArgs a = a_in;
    a.k_chunk = 1;
    {   // split-K for layers whose tiles fill less than half of the SMs (tc_decode_tile)
        static const int enabled = getenv("B2S_TC_SPLITK") ? atoi(getenv("B2S_TC_SPLITK")) : 1;
        const int Cs = MODE == MODE_FWD ? a.g.Cin : a.g.Cout;
        const int nchunks = (Cs + TC_KB - 1) / TC_KB;
        const int PT = a.npairs * a.g.KH * a.g.KW;
        if (enabled && tc_splitk_allowed() && a.relu_mode != 1 && tiles * 2 <= kNumSMs && PT >= 2) {
            int ksplit = (int)std::min<long long>(PT, kNumSMs / tiles);
            while (ksplit > 1 && ((PT + ksplit - 1) / ksplit) * nchunks < 6) --ksplit;     // >= 6 k-blocks per slice
            if (ksplit > 1) {
                const int per = (PT + ksplit - 1) / ksplit;
                ksplit = (PT + per - 1) / per;                                              // no empty slice
            }
            if (ksplit > 1) {
                a.k_chunk = ksplit;
                tiles *= ksplit;
                if (!a.accumulate) {           // the partial tiles are added with atomics: start from zero
                    const int HWd = MODE == MODE_FWD ? a.g.OH * a.g.OW : a.g.H * a.g.W;
                    const long long d_ss = MODE == MODE_FWD ? a.g.out_sstride : a.g.in_sstride;
                    cudaError_t e = cudaMemset2DAsync(a.out, (size_t)d_ss * sizeof(float), 0, (size_t)Cd * HWd * sizeof(float),
                                                      (size_t)a.g.batch, st);
                    if (e != cudaSuccess) { set_error("conv_tc split-K: cudaMemset2DAsync: %s", cudaGetErrorString(e)); return -2; }
                }
            }
        }
    }
    const unsigned grid = (unsigned)std::min<long long>(tiles, kNumSMs);
    static const bool want_trace = getenv("B2S_TC_TRACE") != nullptr;
    static int traced = 0;
    if (want_trace && traced < 6) {            // debug: synchronous launch + timeline dump of CTA 0
        ++traced;
        long long* d_tr = nullptr;
        const size_t n = 4 * TC_TRACE_KB * 4;
        cudaMalloc(&d_tr, n * sizeof(long long));
        cudaMemset(d_tr, 0, n * sizeof(long long));
        ConvKArgs b = a;
        b.trace = d_tr;
        conv_tc_kernel<BN, MODE><<<grid, TC_THREADS, smem, st>>>(b);
        cudaStreamSynchronize(st);
        std::vector<long long> h(n);
        cudaMemcpy(h.data(), d_tr, n * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(d_tr);
        long long t0 = 0;
        for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        fprintf(stderr, "TC trace BN=%d mode=%d J=%lld Cs=%d Cd=%d KHW=%d pairs=%d grid=%u (clocks since first stamp)\n", BN, MODE, J,
                MODE == MODE_FWD ? a.g.Cin : a.g.Cout, Cd, a.g.KH * a.g.KW, a.npairs, grid);
        for (int k = 0; k < TC_TRACE_KB; ++k) {
            auto at = [&](int role, int slot) { long long v = h[((size_t)role * TC_TRACE_KB + k) * 4 + slot]; return v ? v - t0 : -1; };
            if (at(1, 0) < 0) break;
            fprintf(stderr, "  kb %2d  F: start %6lld loads %6lld adv %6lld | A: start %6lld stage %6lld st %6lld done %6lld | MMA: poll %6lld a_ok %6lld all_ok %6lld issued %6lld | D: poll %6lld full %6lld done %6lld\n",
                    k, at(3, 0), at(3, 1), at(3, 2), at(0, 0), at(0, 1), at(0, 2), at(0, 3), at(1, 0), at(1, 1), at(1, 2), at(1, 3), at(2, 0), at(2, 1), at(2, 2));
        }
        return 1;
    }
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

// The base pass (values, ReLU / max-pool decisions, gradient) keeps the deterministic single-CTA-per-tile summation:
// its decisions are what every later pass is conditioned on.  The jet passes (linear given the decisions) may split.
static thread_local bool g_splitk_allowed = true;
void set_tc_splitk_allowed(bool on) { g_splitk_allowed = on; }
bool tc_splitk_allowed() { return g_splitk_allowed; }

int try_launch_conv_tc(int mode, cudaStream_t st, const ConvKArgs& a) {
    if (g_tc_mode == 0) return 0;
    for (int p = 0; p < a.npairs; ++p)
        if (!a.pack[p]) return 0;                     // the plan did not pack this layer (shape not eligible)
    const ConvGeom& g = a.g;
    const int Cd = mode == MODE_FWD ? g.Cout : g.Cin;
    const int Cs = mode == MODE_FWD ? g.Cin : g.Cout;
    const long long J = (long long)g.batch * (mode == MODE_FWD ? g.OH * g.OW : g.H * g.W);
    // automatic: enough pixel tiles to be worth a tensor-core launch; classifier-sized problems stay on the
    // CUDA-core kernels
    if (g_tc_mode == 1 && !tc_worth_it(J, Cs, Cd, g.KH * g.KW)) return 0;
    return mode == MODE_FWD ? launch_tc_mode<MODE_FWD>(st, a, J, Cd) : launch_tc_mode<MODE_DGRAD>(st, a, J, Cd);
}

}  // namespace b2s
