// Thin-channel convolutions on the tensor cores with the ACTIVATION operand in tensor memory (tcgen05.mma with A in
// TMEM), 3xTF32 operand split: the default configuration's encoder layers behind the head (EVAL mode; reference layers
// HiddenStateExtractor/vq_vae.py:278-289 and the ResidualBlock convs at :203-209).
//
// Why this form.  These layers are implicit GEMMs  D[128 pixels x Cout] += A[128 x K] * B[Cout x K]^T  with K = C*kh*kw
// = 32..256 but only N = Cout = 16..32 output channels.  With both operands in shared memory (conv_tc.cu) every
// activation element crosses the shared-memory port five times (TMA write, split read, split writes of hi and lo, MMA
// reads of hi and lo) for 16 multiply-adds: no faster than the CUDA-core kernel.  Here the im2col row of a pixel is
// built in REGISTERS by the thread that owns the pixel (one 8-byte shared-memory load per input row and channel, the
// horizontal neighbours by warp shuffle), split hi/lo in registers, and written straight into tensor memory
// (tcgen05.st, lane = pixel, column = k).  The tensor core then reads A from TMEM; only the small weight operand
// ([b_hi; b_lo] rows of one K-major 128-byte-swizzled tile set, resident for the whole persistent CTA) is read from
// shared memory.  Shared-memory traffic per activation element: one TMA write + one read.
//
//   warps 0-3  one thread = one output pixel = one TMEM lane: gather + split + tcgen05.st of K-chunk `ky` into one of
//              two A buffers; later the epilogue (tcgen05.ld of [main | small] -> bias, skip, ReLU -> NCHW store)
//   warp 4     issues, per K=8 step, a_hi * [b_hi; b_lo] (N = 2*Cout) and a_lo * b_hi (N = Cout, into the `small`
//              columns); tcgen05.commit releases the A buffer / publishes the accumulator
//   warp 5     TMA producer: the input rows of the next tiles (zero padding above/below = out-of-bounds fill)
#include "common.cuh"

#include <algorithm>
#include <stdlib.h>

#include <cuda.h>
#include <cudaTypedefs.h>

namespace dmb {
namespace {

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {      // keeps the warp converged
    uint32_t done;
    do {
        done = mbar_try(bar, parity) ? 1u : 0u;
    } while (!__all_sync(0xffffffffu, done != 0));
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, TF32 inputs, fp32 accumulation; converged warp, one elected lane issues
__device__ __forceinline__ void tc_mma_tf32_ts_elect(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive columns, registers -> tensor memory: thread t of the warp writes lane (base lane + t)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t tf32_rna_bits(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return r;
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (see conv_tc.cu)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int TM_THREADS = 320;      // 8 gather/epilogue warps (two per TMEM lane quadrant) + MMA warp + TMA warp
constexpr int cpow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

// waits of the two helper warps back off: a bare try_wait loop returns every ~10 cycles and the spinning producer
// thread alone took 20 % of the SM's issue slots (ncu, profiles/r2_ncu_conv_tm_enc4_v1.txt)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(128);
}
__device__ __forceinline__ void mbar_wait_warp_sleep(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        done = mbar_try(bar, parity) ? 1u : 0u;
        if (!done) __nanosleep(64);
    } while (!__all_sync(0xffffffffu, done != 0));
}
// tensor memory -> registers, 32 lanes x N consecutive columns
template <int N> struct TmemLd;
template <> struct TmemLd<8> {
    static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr) : "memory");
    }
};
template <> struct TmemLd<16> {
    static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr) : "memory");
    }
};
template <> struct TmemLd<32> {
    static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v) {
        TmemLd<16>::ld(taddr, v);
        TmemLd<16>::ld(taddr + 16u, v + 16);
    }
};
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// FUSE_ (3x3, 16 -> 32 only): the rest of a ResidualBlock layer rides on the same tile -- ReLU, the 1x1 convolution back
// to COUT2 = 16 channels as a second small GEMM (its A operand is the first GEMM's activated output, split and stored
// to tensor memory by the epilogue warps), bias and the skip connection (vq_vae.py:203-209, :222-225).
// BN_ (train-mode BatchNorm around the layer, DMB_BN_PER_SAMPLE / DMB_BN_BATCH): the producer's pending affine + ReLU is
// applied to every gathered value (relu?(x * scale[c] + shift[c]), tables per channel or per patch and channel; rows
// outside the image stay zero), the layer's raw output is stored, and the epilogue leaves one (sum, sum of squares)
// partial per (patch, tile, channel) for bn_finalize -- the contract of conv_tma.cuh / conv_fwd.cu.
// DG_ (data gradient of a training step, whole-batch statistics): plain input (the gradient at the layer's output, any
// BatchNorm-backward transform already applied), and the epilogue of the CUDA-core data-gradient kernels (conv_fwd.cu):
// ReLU gate of the layer back-propagated into ([mask_src * mask_s[c] + mask_t[c] > 0]), skip gradient, store, and the
// next BatchNorm backward's sums (sum o, sum o * stat_src) -- or (sum o, sum o^2) for a bias gradient -- per CTA.
// DUAL_ (with DG_): the BatchNorm backward in front of the layer is applied on load,  x = A[c]*g + Bc[c]*y + Cc[c]  with g
// and the raw activation y staged side by side (two TMA boxes per stage) -- otherwise the caller materialises it.
// CT_: a ConvTranspose2d(4x4, stride 2, padding 1) CIN -> COUT/4 written as the 3x3 stride-1 convolution it is on the
// INPUT grid: output pixel (2y + py, 2x + px) only sees input pixels (y + dy, x + dx) with dy in {-1, 0} (py = 0) or
// {0, 1} (py = 1), so the four output phases are 4 * Cout "channels" n = (2 py + px) * Cout + co of one 3x3 kernel whose
// unused taps are zero (pack_tm image with ct = 1); the epilogue scatters them (pixel shuffle): warp group = py, a
// thread stores the px pair of its input pixel as one float2.  Statistics rows: one per (CTA, quadrant, py).
template <int KS_, int S_, int CIN_, int COUT_, int WIN_, bool FUSE_ = false, bool INRELU_ = false, bool BN_ = false,
          bool DG_ = false, bool DUAL_ = false, bool CT_ = false>
struct TM {
    static constexpr bool CT = CT_;
    static constexpr int CT_C = COUT_ / 4;                   // CT: output channels of the transposed convolution
    static constexpr bool BN = BN_;
    static constexpr bool DG = DG_;
    static constexpr bool DUAL = DUAL_;
    static constexpr bool STATS = BN_ || DG_;
    static constexpr int KS = KS_, S = S_, CIN = CIN_, COUT = COUT_, W = WIN_, H = WIN_;
    static constexpr bool FUSE = FUSE_;
    static constexpr bool INRELU = INRELU_;                  // ReLU on load, compiled in (a run-time flag left 16
                                                             // predicated FMNMX per chunk in the layers without it)
    static constexpr int COUT2 = 16;                         // FUSE: output channels of the 1x1
    static constexpr int K2 = COUT;                          // FUSE: its reduction length
    static constexpr int NROWS2 = 2 * COUT2;
    static constexpr int PAD = (KS == 1) ? 0 : 1;
    static constexpr int WO = W / S, HO = H / S;
    static constexpr int TH = 128 / WO;                      // output rows per tile (one tile = 128 pixels)
    static constexpr int TILES = HO / TH;                    // tiles per patch
    static constexpr int RIN = (TH - 1) * S + KS;            // input rows per tile
    // K order: k = (ky*CIN + ci)*KS + kx.  A chunk is CPC input channels of one kernel row (16..48 k values): what one
    // warp group gathers, splits and stores to tensor memory at a time
    static constexpr int CPC = (KS * CIN > 48) ? CIN / 2 : CIN;
    static constexpr int CPR = CIN / CPC;                    // chunks per kernel row
    static constexpr int KC = KS * CPC;                      // K of one chunk
    static constexpr int NCH = KS * CPR;                     // chunks per tile
    // Three chunks per tile (the 3x3 layers with 16 input channels) would leave warp group 0 with two chunks and group 1
    // with one: there BOTH groups build every chunk, half of its channels each (24 k-values), into the same A buffer
    static constexpr bool SPLIT = (NCH == 3) && (CPC % 2 == 0) && ((KC / 2) % 8 == 0);
    static constexpr int CPG = SPLIT ? CPC / 2 : CPC;        // channels a group gathers per chunk
    static constexpr int KP = KS * CPG;                      // k-values a group builds per chunk
    static constexpr int K = KS * KS * CIN;
    static constexpr int NT = (K + 31) / 32;                 // 128-byte operand tiles along K
    static constexpr int NROWS = 2 * COUT;                   // [b_hi; b_lo]
    static constexpr int B1_FLOATS = NT * NROWS * 32;
    static constexpr int B_FLOATS = B1_FLOATS + (FUSE ? (K2 / 32) * NROWS2 * 32 : 0);
    static constexpr int IN_FLOATS = CIN * RIN * W;          // one input stage
    static constexpr int IN_BYTES = IN_FLOATS * 4;
    static constexpr int A_COLS = 2 * KC;                    // hi | lo of one chunk
    static constexpr int D_COL = 2 * A_COLS;                 // two A buffers, then the accumulators [main | small]
    // The tensor core TRUNCATES when it aligns and adds into the fp32 accumulator, a bias that grows with the number of
    // MMAs chained on one accumulator (measured: 2.8e-6 on z_before after five layers with one accumulator, against
    // 1.2e-6 for the FFMA kernels).  Chunk ky therefore accumulates into accumulator ky % NACC and the epilogue adds the
    // partial sums in round-to-nearest fp32.  Two accumulators where tensor memory allows it without losing the second
    // resident CTA: reading them back (LDTM, 64 B/clk per SM) is what a third and fourth would cost.
    static constexpr int NACC = (NCH >= 2 && COUT <= 32 && cpow2(D_COL + 2 * 2 * COUT) <= cpow2(D_COL + 2 * COUT)) ? 2 : 1;
    static constexpr int TMEM_COLS = cpow2(D_COL + NACC * 2 * COUT);
    // Two accumulator sets where they fit: the epilogue of tile t then runs AFTER this warp group's first chunk of tile
    // t+1 (ncu: 30 % of the stall samples sat on the wait for the tile's last MMAs in front of the epilogue)
    // (only where each warp group still has a chunk to build AFTER its deferred epilogue: the tile's accumulator cannot
    // then complete -- and the single d_full barrier cannot run two phases ahead of a waiting group -- before that wait)
    static constexpr bool DBUF = !FUSE && NCH >= 4 && (D_COL + 2 * NACC * 2 * COUT <= TMEM_COLS);
    static constexpr int D_SET = NACC * 2 * COUT;
    static constexpr int CTAS = (TMEM_COLS <= 256) ? 2 : 1;  // per SM
    static constexpr int SMEM_BUDGET = (CTAS == 2 ? 110 : 200) * 1024;
    static constexpr int STAGE_BYTES = (((DUAL ? 2 : 1) * IN_BYTES) + 127) & ~127;
    static constexpr int NSTAGE_RAW = (SMEM_BUDGET - B_FLOATS * 4 - 2048) / STAGE_BYTES;
    static constexpr int NSTAGE = NSTAGE_RAW > 4 ? 4 : NSTAGE_RAW;
    static constexpr size_t SMEM = 1024 + (size_t)B_FLOATS * 4 + (size_t)NSTAGE * STAGE_BYTES + 256 + 2048;  // + barriers + stat_red
    static constexpr int HALF = COUT / 2;                    // output channels per epilogue warp group
    static constexpr int NS = CT ? CT_C : HALF;              // statistics channels a thread holds
    static constexpr int WARP_ROWS = CT ? 8 : 4;             // whole-batch statistics: warp rows folded into ONE row per CTA
    static constexpr int CS = CT ? CT_C : COUT;              // channels of a statistics row
    static constexpr int STAT_ROWS = 1;                      // whole-batch statistics rows per CTA
    static_assert(WARP_ROWS * CS * 16 <= 2048, "the CTA's statistics fold fits its shared-memory slot");
    static_assert(128 % WO == 0 && HO % TH == 0, "a tile is 128 consecutive output pixels of one patch");
    static_assert(WO <= 32 && 32 % WO == 0, "a warp covers whole output rows (shuffle neighbours)");
    static_assert(KC % 16 == 0 && KC <= 48 && CIN % CPC == 0, "chunk = 16..48 k values");
    static_assert(COUT == 16 || COUT == 32 || (CT && COUT == 64), "N = Cout and 2*Cout must be legal MMA widths");
    static_assert(!CT || (KS == 3 && S == 1 && !FUSE && !BN && !INRELU), "transposed form: 3x3 on the input grid");
    static_assert((NROWS * 128) % 1024 == 0, "operand tiles stay 1024-byte aligned");
    static_assert(TMEM_COLS <= 512 && NSTAGE >= 2, "resources");
    static_assert((KS == 1 && S == 1) || (KS == 3 && S == 1) || (KS == 4 && S == 2), "unsupported kernel");
    static_assert((W * 4) % 16 == 0 && W <= 256 && RIN <= 256 && CIN <= 256, "TMA box");
    static_assert(!FUSE || (KS == 3 && COUT == 32 && CIN == COUT2 && NACC == 1 && A_COLS >= 2 * K2),
                  "the fused tail is the 3x3 16 -> 32 -> 1x1 -> 16 residual layer");
    static_assert(!BN || (!FUSE && !INRELU), "BatchNorm form: ReLU on load is a run-time flag of the transform");
    static_assert(!DG || (!FUSE && !INRELU && !BN), "data-gradient form: plain input");
    static_assert(!DUAL || (DG && KS != 4 && IN_BYTES % 128 == 0), "dual-tensor load: data-gradient form, stride 1");
};

struct TmKArgs {
    const float* wtm;       // pack_tm_weights image: [NT][NROWS][32] swizzled
    const float* bias;      // [Cout]
    float* y;               // (B, Cout, Ho, Wo)
    const float* skip;      // (B, Cout, Ho, Wo) or nullptr
    const float* bias2;     // FUSE: bias of the 1x1 [16]
    const float* in_scale;  // BN: pending affine of the input, [Cin] or [B][Cin] (nullptr = identity)
    const float* in_shift;
    int in_per_sample;
    double* stats;          // BN: [B][TILES*4][Cout][2] partial (sum, sum of squares) per warp (32 pixels)
    int stats_batch;        // BN: whole-batch statistics -- every warp adds up its tiles' partials and leaves ONE row at
                            // the end: [gridDim.x * 4][Cout][2]
    int64_t ntiles;
    int in_relu, out_relu;
    int dbg;                // DMB_TM_DBG skip experiments (results are wrong): 1 no a_lo*b_hi MMAs, 2 no a_hi MMAs,
                            // 4 no gather (zeros), 8 no split / tensor-memory stores
};
// extras of the data-gradient form: a separate kernel parameter (growing TmKArgs by as little as 16 bytes changed the
// register allocation of EVERY variant: 0 -> 28..116 bytes of spills in the plain forward kernels)
struct TmDgArgs {
    const float* mask_src;  // (B, Cout, Ho, Wo) or nullptr
    const float* mask_s;    // its affine [Cout] (nullptr = identity)
    const float* mask_t;
    const float* stat_src;  // (B, Cout, Ho, Wo) or nullptr
    const float* in_a;      // DUAL: x = in_a[c] * g + in_b[c] * y + in_c[c] per input channel (nullptr: g alone, one box)
    const float* in_b;
    const float* in_c;
};

template <class C>
__global__ void __launch_bounds__(TM_THREADS, C::CTAS)
conv_tm_kernel(const __grid_constant__ CUtensorMap tmap, const TmKArgs a, const TmDgArgs dg,
               const __grid_constant__ CUtensorMap tmap2, const TmFinArgs fin) {
    constexpr int KS = C::KS, S = C::S, CIN = C::CIN, COUT = C::COUT, W = C::W, KC = C::KC, HALF = C::HALF;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    float* bs = reinterpret_cast<float*>(bp);                                  // weight operand tiles
    const uint32_t bs_u = base;
    uint8_t* stage0 = bp + (size_t)C::B_FLOATS * 4;
    const uint32_t stage0_u = base + (uint32_t)C::B_FLOATS * 4u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + (size_t)C::NSTAGE * C::STAGE_BYTES);
    const uint32_t bars_u = stage0_u + (uint32_t)(C::NSTAGE * C::STAGE_BYTES);
    // barrier slots: in_full[NSTAGE] | in_empty[NSTAGE] | a_full[2] | a_empty[2] | d_full | d_empty | a2_full | d2_full | slot
    const uint32_t in_full = bars_u, in_empty = bars_u + 8u * C::NSTAGE, a_full = bars_u + 16u * C::NSTAGE;
    const uint32_t a_empty = a_full + 16u, d_full = a_empty + 16u, d_empty = d_full + 8u;
    const uint32_t a2_full = d_empty + 8u, d2_full = a2_full + 8u;
    uint32_t* slot_mem = reinterpret_cast<uint32_t*>(bars + 2 * C::NSTAGE + 8);
    [[maybe_unused]] double2* stat_red = reinterpret_cast<double2*>(bars + 32);  // 2 KB behind the 256-byte barrier block
    const uint32_t slot = smem_u32(slot_mem);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(in_full + 8u * s, 1u); mbar_init(in_empty + 8u * s, 8u); }
        for (int s = 0; s < 2; ++s) { mbar_init(a_full + 8u * s, C::SPLIT ? 8u : 4u); mbar_init(a_empty + 8u * s, 1u); }
        mbar_init(d_full, 1u);
        mbar_init(d_empty, 8u);
        mbar_init(a2_full, 8u);
        mbar_init(d2_full, 1u);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmap) : "memory");
    }
    if (warp == 8) tmem_alloc(slot, (uint32_t)C::TMEM_COLS);
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernels' output
    for (int i = tid; i < C::B_FLOATS / 4; i += TM_THREADS)
        reinterpret_cast<float4*>(bs)[i] = __ldg(reinterpret_cast<const float4*>(a.wtm) + i);
    fence_proxy_async();                     // weight tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot_mem);

    if (warp == 9) {
        // ---- TMA producer
        if (lane == 0) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
                const int stage = it % C::NSTAGE, n = it / C::NSTAGE;
                if (n > 0) mbar_wait_sleep(in_empty + 8u * stage, (uint32_t)((n - 1) & 1));
                const int b = (int)(tile / C::TILES), t = (int)(tile % C::TILES);
                const uint32_t bar = in_full + 8u * stage;
                if constexpr (C::DUAL) {
                    const bool dual = dg.in_a != nullptr;
                    mbar_expect_tx(bar, (uint32_t)(dual ? 2 * C::IN_BYTES : C::IN_BYTES));
                    tma_load_4d(stage0_u + (uint32_t)(stage * C::STAGE_BYTES), &tmap, bar, 0, t * C::TH * S - C::PAD, b, 0);
                    if (dual)
                        tma_load_4d(stage0_u + (uint32_t)(stage * C::STAGE_BYTES + C::IN_BYTES), &tmap2, bar, 0,
                                    t * C::TH * S - C::PAD, b, 0);
                    continue;
                }
                mbar_expect_tx(bar, (uint32_t)C::IN_BYTES);
                tma_load_4d(stage0_u + (uint32_t)(stage * C::STAGE_BYTES), &tmap, bar, 0, t * C::TH * S - C::PAD, b, 0);
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer (whole warp converged, one elected lane per instruction)
        const uint32_t idesc_main = make_idesc_tf32(128, 2 * COUT), idesc_lo = make_idesc_tf32(128, COUT);
        const uint32_t d_tmem = tmem_base + (uint32_t)C::D_COL;
        uint32_t nuse0 = 0, nuse1 = 0;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
            // the epilogue (both warp groups) of the tile that last used this accumulator set has read it
            constexpr int DIST = C::DBUF ? 2 : 1;
            if (it >= DIST) mbar_wait_warp_sleep(d_empty, (uint32_t)((it - DIST) & 1));
            const uint32_t d_set = d_tmem + (C::DBUF ? (uint32_t)((it & 1) * C::D_SET) : 0u);
#pragma unroll
            for (int ch = 0; ch < C::NCH; ++ch) {
                const uint32_t buf = (uint32_t)(ch & 1);
                const uint32_t n = buf ? nuse1 : nuse0;
                mbar_wait_warp_sleep(a_full + 8u * buf, n & 1u);
                if (buf) ++nuse1; else ++nuse0;
                tc_fence_after();
                const uint32_t a_hi = tmem_base + buf * (uint32_t)C::A_COLS, a_lo = a_hi + (uint32_t)KC;
                const uint32_t dacc = d_set + (uint32_t)((ch % C::NACC) * 2 * COUT);
#pragma unroll
                for (int s = 0; s < KC / 8; ++s) {
                    const int kg = ch * KC + 8 * s;
                    const uint64_t bd = make_desc_sw128(bs_u + (uint32_t)((kg >> 5) * C::NROWS * 128)) +
                                        (uint64_t)(2 * ((kg & 31) >> 3));
                    if (!(a.dbg & 2)) tc_mma_tf32_ts_elect(dacc, a_hi + 8u * s, bd, idesc_main, (ch >= C::NACC || s != 0) ? 1u : 0u);
                    if (!(a.dbg & 1)) tc_mma_tf32_ts_elect(dacc + (uint32_t)COUT, a_lo + 8u * s, bd, idesc_lo, 1u);
                }
                tc_commit_elect(a_empty + 8u * buf);
            }
            tc_commit_elect(d_full);
            if constexpr (C::FUSE) {
                // second GEMM: relu(3x3 output) [128 x 32] (A buffer 0, written by the epilogue warps) x W2 [16 x 32]^T
                mbar_wait_warp_sleep(a2_full, (uint32_t)(it & 1));
                tc_fence_after();
                const uint32_t idesc2_main = make_idesc_tf32(128, 2 * C::COUT2), idesc2_lo = make_idesc_tf32(128, C::COUT2);
                const uint32_t a2_hi = tmem_base, a2_lo = tmem_base + (uint32_t)C::K2;
                const uint64_t bd2 = make_desc_sw128(bs_u + (uint32_t)(C::B1_FLOATS * 4));
#pragma unroll
                for (int s = 0; s < C::K2 / 8; ++s) {
                    tc_mma_tf32_ts_elect(d_tmem, a2_hi + 8u * s, bd2 + (uint64_t)(2 * s), idesc2_main, s != 0 ? 1u : 0u);
                    tc_mma_tf32_ts_elect(d_tmem + (uint32_t)C::COUT2, a2_lo + 8u * s, bd2 + (uint64_t)(2 * s), idesc2_lo, 1u);
                }
                tc_commit_elect(d2_full);
            }
        }
    } else {
        // ---- gather / split / store into tensor memory, then the epilogue.  thread = pixel = TMEM lane; the two warp
        // groups (warps 0-3, 4-7) share the pixels: group g builds the chunks g, g+2, ... (it owns A buffer g) and
        // writes output channels [g*Cout/2, (g+1)*Cout/2)
        const int wg = warp >> 2, q = warp & 3;
        const int p = q * 32 + lane;                         // pixel inside the tile, row-major (row, ox)
        const int prow = p / C::WO, ox = p % C::WO;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool left = (ox == 0), right = (ox == C::WO - 1);
        float bias_r[HALF];
#pragma unroll
        for (int c = 0; c < HALF; ++c) bias_r[c] = __ldg(a.bias + (C::CT ? c % C::CT_C : wg * HALF + c));
        uint32_t my_n = 0;                                   // chunks this group has produced (uses of its A buffer)
        [[maybe_unused]] double acc_s = 0.0, acc_q = 0.0;    // BN, stats_batch: this lane's channel over all tiles of the CTA
        [[maybe_unused]] int acc_chan = 0;

        // gather + split + tensor-memory store of chunk `ch` of the tile staged at `tin` (tile = patch b, rows from row0)
        auto produce = [&](const float* tin, int ch, int b = 0, int row0 = 0, int it = 0) {
            // gather first (shared memory only), then wait for the buffer: the MMAs of this group's previous chunk
            // overlap the loads
            float v[C::KP];
            const int ky = ch / C::CPR, ci0 = (ch % C::CPR) * C::CPC + (C::SPLIT ? wg * C::CPG : 0);
            const float* rp = tin + ((size_t)ci0 * C::RIN + prow * S + ky) * W + S * ox;
            // BN: affine + ReLU of the producer on every value; an input row outside the image contributes zeros (the TMA
            // zero fill would otherwise turn into `shift`), the left / right border columns are zeroed below as always
            [[maybe_unused]] const float* tsc = nullptr;
            [[maybe_unused]] const float* tsh = nullptr;
            [[maybe_unused]] bool row_ok = true;
            [[maybe_unused]] float xf_lo = 0.f;
            if constexpr (C::BN) {
                const size_t tb = (a.in_per_sample ? (size_t)b * CIN : 0) + ci0;
                tsc = a.in_scale ? a.in_scale + tb : nullptr;
                tsh = a.in_scale ? a.in_shift + tb : nullptr;
                row_ok = (unsigned)(row0 + prow * S + ky) < (unsigned)C::H;
                xf_lo = a.in_relu ? 0.f : -INFINITY;
            }
            [[maybe_unused]] bool dual = false;
            if constexpr (C::DUAL) {
                dual = dg.in_a != nullptr;
                row_ok = (unsigned)(row0 + prow * S + ky) < (unsigned)C::H;
            }
            if (a.dbg & 4) {
#pragma unroll
                for (int j = 0; j < C::KP; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int ci = 0; ci < C::CPG; ++ci) {
                if (a.dbg & 4) break;
                [[maybe_unused]] float sc = 1.f, sh = 0.f;
                if constexpr (C::BN) {
                    if (tsc) { sc = __ldg(tsc + ci); sh = __ldg(tsh + ci); }
                    // (selects, not a branch: rows of one warp can differ and the shuffles below are full-mask)
                    if (!row_ok) { sc = 0.f; sh = 0.f; }
                }
                if constexpr (KS == 4) {
                    float2 f = *reinterpret_cast<const float2*>(rp + ci * C::RIN * W);
                    if constexpr (C::INRELU) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
                    if constexpr (C::BN) { f.x = fmaxf(fmaf(f.x, sc, sh), xf_lo); f.y = fmaxf(fmaf(f.y, sc, sh), xf_lo); }
                    const float up = __shfl_up_sync(0xffffffffu, f.y, 1), dn = __shfl_down_sync(0xffffffffu, f.x, 1);
                    v[ci * 4 + 0] = left ? 0.f : up;
                    v[ci * 4 + 1] = f.x;
                    v[ci * 4 + 2] = f.y;
                    v[ci * 4 + 3] = right ? 0.f : dn;
                } else if constexpr (KS == 3) {
                    float f = rp[ci * C::RIN * W];
                    if constexpr (C::INRELU) f = fmaxf(f, 0.f);
                    if constexpr (C::BN) f = fmaxf(fmaf(f, sc, sh), xf_lo);
                    if constexpr (C::DUAL) {
                        if (dual) {      // BatchNorm backward on load; rows outside the image stay zero
                            float ca = __ldg(dg.in_a + ci0 + ci), cb = __ldg(dg.in_b + ci0 + ci), cc = __ldg(dg.in_c + ci0 + ci);
                            if (!row_ok) { ca = 0.f; cb = 0.f; cc = 0.f; }
                            f = fmaf(f, ca, fmaf(rp[C::IN_FLOATS + ci * C::RIN * W], cb, cc));
                        }
                    }
                    const float up = __shfl_up_sync(0xffffffffu, f, 1), dn = __shfl_down_sync(0xffffffffu, f, 1);
                    v[ci * 3 + 0] = left ? 0.f : up;
                    v[ci * 3 + 1] = f;
                    v[ci * 3 + 2] = right ? 0.f : dn;
                } else {
                    float f = rp[ci * C::RIN * W];
                    if constexpr (C::INRELU) f = fmaxf(f, 0.f);
                    if constexpr (C::BN) f = fmaxf(fmaf(f, sc, sh), xf_lo);
                    if constexpr (C::DUAL) {
                        if (dual)
                            f = fmaf(f, __ldg(dg.in_a + ci0 + ci),
                                     fmaf(rp[C::IN_FLOATS + ci * C::RIN * W], __ldg(dg.in_b + ci0 + ci), __ldg(dg.in_c + ci0 + ci)));
                    }
                    v[ci] = f;
                }
            }
            // A buffer and how often it has been used before: a group's own buffer, or (SPLIT) buffer ch & 1 shared by
            // both groups -- buffer 0 takes chunks 0 and 2 of every tile, buffer 1 chunk 1
            const uint32_t bufi = C::SPLIT ? (uint32_t)(ch & 1) : (uint32_t)wg;
            const uint32_t used = C::SPLIT ? (uint32_t)(it * ((ch & 1) ? 1 : 2) + (ch >> 1)) : my_n;
            if (used > 0) mbar_wait(a_empty + 8u * bufi, (used - 1) & 1u);
            tc_fence_after();
            const uint32_t a_hi = lane_base + bufi * (uint32_t)C::A_COLS + (C::SPLIT ? (uint32_t)(wg * C::KP) : 0u);
            const uint32_t a_lo = a_hi + (uint32_t)KC;
#pragma unroll
            for (int j0 = 0; j0 < C::KP; j0 += 16) {
                if (a.dbg & 8) break;
                // hi = x rounded to TF32 on the bit pattern (cvt.rna.tf32 compiles to a five-instruction sequence),
                // lo = x - hi exactly; the tensor core drops the 13 low bits of lo
                constexpr int NV16 = 16;
                uint32_t hi[NV16], lo[NV16];
#pragma unroll
                for (int j = 0; j < NV16; ++j) {
                    if (j0 + j < C::KP) {
                        hi[j] = (__float_as_uint(v[j0 + j]) + 0x1000u) & 0xFFFFE000u;
                        lo[j] = __float_as_uint(v[j0 + j] - __uint_as_float(hi[j]));
                    } else {
                        hi[j] = 0u; lo[j] = 0u;
                    }
                }
                if (j0 + 16 <= C::KP) {
                    tmem_st16(a_hi + (uint32_t)j0, hi);
                    tmem_st16(a_lo + (uint32_t)j0, lo);
                } else {                                 // (KP = 24: the last eight k-values)
                    tmem_st8(a_hi + (uint32_t)j0, hi);
                    tmem_st8(a_lo + (uint32_t)j0, lo);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8u * bufi);
            ++my_n;
        };

        // epilogue of tile `tile` (the it-th of this CTA): this group's half of the output channels
        auto epilogue = [&](int64_t tile, int it, const float* tin, int stage) {
            mbar_wait(d_full, (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t d_tmem = lane_base + (uint32_t)C::D_COL + (uint32_t)(wg * HALF) +
                                    (C::DBUF ? (uint32_t)((it & 1) * C::D_SET) : 0u);
            uint32_t r[C::NACC][2][HALF];
#pragma unroll
            for (int j = 0; j < C::NACC; ++j) {
                TmemLd<HALF>::ld(d_tmem + (uint32_t)(j * 2 * COUT), r[j][0]);
                TmemLd<HALF>::ld(d_tmem + (uint32_t)(j * 2 * COUT + COUT), r[j][1]);
            }
            tmem_ld_wait();
            const int b = (int)(tile / C::TILES), t = (int)(tile % C::TILES);
            const size_t pix = (size_t)(t * C::TH + prow) * C::WO + ox;
            if constexpr (!C::FUSE) {
                tc_fence_before();           // the accumulators may be overwritten once every warp has arrived
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty);
                [[maybe_unused]] float ssum[C::NS], ssq[C::NS];
                auto accv = [&](int c) {
                    float val = __uint_as_float(r[0][0][c]) + __uint_as_float(r[0][1][c]);
                    if constexpr (C::NACC == 2)
                        val += __uint_as_float(r[1][0][c]) + __uint_as_float(r[1][1][c]);
                    return val + bias_r[c];
                };
                if constexpr (C::CT) {
                    // pixel shuffle: this group's columns are n = wg*HALF + px*CT_C + co  ->  output (co, 2y + wg, 2x + px)
                    constexpr int CT_C = C::CT_C;
                    constexpr size_t PLANE = (size_t)4 * C::HO * C::WO;
                    const size_t o0 = (((size_t)b * CT_C) * (2 * C::HO) + 2 * (t * C::TH + prow) + wg) * (2 * C::WO) + 2 * ox;
#pragma unroll
                    for (int co = 0; co < CT_C; ++co) {
                        float v0 = accv(co), v1 = accv(CT_C + co);
                        [[maybe_unused]] float2 mraw = make_float2(0.f, 0.f);
                        if constexpr (C::DG) {
                            if (dg.mask_src) {
                                mraw = __ldg(reinterpret_cast<const float2*>(dg.mask_src + o0 + co * PLANE));
                                float m0 = mraw.x, m1 = mraw.y;
                                if (dg.mask_s) {
                                    const float ms = __ldg(dg.mask_s + co), mt = __ldg(dg.mask_t + co);
                                    m0 = fmaf(m0, ms, mt); m1 = fmaf(m1, ms, mt);
                                }
                                if (!(m0 > 0.f)) v0 = 0.f;
                                if (!(m1 > 0.f)) v1 = 0.f;
                            }
                        }
                        if (a.out_relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                        *reinterpret_cast<float2*>(a.y + o0 + co * PLANE) = make_float2(v0, v1);
                        if constexpr (C::STATS) {
                            ssum[co] = v0 + v1; ssq[co] = fmaf(v0, v0, v1 * v1);
                            if (dg.stat_src) {
                                const float2 sv = (dg.stat_src == dg.mask_src)
                                                      ? mraw : __ldg(reinterpret_cast<const float2*>(dg.stat_src + o0 + co * PLANE));
                                ssq[co] = fmaf(v0, sv.x, v1 * sv.y);
                            }
                        }
                    }
                } else {
                const size_t chan0 = ((size_t)b * COUT + wg * HALF) * (C::HO * C::WO) + pix;
                float* yp = a.y + chan0;
                const float* sp = a.skip ? a.skip + chan0 : nullptr;
#pragma unroll
                for (int c = 0; c < HALF; ++c) {
                    float val = accv(c);
                    [[maybe_unused]] float mraw = 0.f;
                    if constexpr (C::DG) {
                        if (dg.mask_src) {
                            float mv = __ldg(dg.mask_src + chan0 + (size_t)c * (C::HO * C::WO));
                            mraw = mv;
                            if (dg.mask_s) mv = fmaf(mv, __ldg(dg.mask_s + wg * HALF + c), __ldg(dg.mask_t + wg * HALF + c));
                            if (!(mv > 0.f)) val = 0.f;
                        }
                    }
                    if (sp) val += __ldg(sp + (size_t)c * (C::HO * C::WO));
                    if (a.out_relu) val = fmaxf(val, 0.f);
                    yp[(size_t)c * (C::HO * C::WO)] = val;
                    if constexpr (C::STATS) {
                        ssum[c] = val; ssq[c] = val * val;
                        if constexpr (C::DG) {
                            // (the BatchNorm whose backward sums these are is usually the one whose output gates: one load)
                            if (dg.stat_src)
                                ssq[c] = val * (dg.stat_src == dg.mask_src ? mraw
                                                                           : __ldg(dg.stat_src + chan0 + (size_t)c * (C::HO * C::WO)));
                        }
                    }
                }
                }
                if constexpr (C::STATS) {
                    // (sum, sum of squares) of this WARP's 32 pixels per channel, one partial row per warp (the four
                    // quadrants of a tile are four rows of the [B][TILES*4][Cout][2] partials: no shared memory, no
                    // barrier between the warps).  Recursive halving: at each of the first log2(HALF) steps a lane keeps
                    // half of the channels it holds and adds its partner's values for them, then plain butterflies --
                    // 9 (HALF = 8) or 16 (HALF = 16) shuffles per quantity instead of 5 per channel; fixed order.
                    if (a.stats) {
                        int held = C::NS;
                        int chan = 0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            if (held > 1) {
                                const int half = held >> 1;
                                const bool upper = (lane & o) != 0;
#pragma unroll
                                for (int i = 0; i < (C::NS + 1) / 2; ++i) {
                                    if (i < half) {
                                        const float send_s = upper ? ssum[i] : ssum[i + half];
                                        const float send_q = upper ? ssq[i] : ssq[i + half];
                                        const float keep_s = upper ? ssum[i + half] : ssum[i];
                                        const float keep_q = upper ? ssq[i + half] : ssq[i];
                                        ssum[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, o);
                                        ssq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, o);
                                    }
                                }
                                chan += upper ? half : 0;
                                held = half;
                            } else {
                                ssum[0] += __shfl_xor_sync(0xffffffffu, ssum[0], o);
                                ssq[0] += __shfl_xor_sync(0xffffffffu, ssq[0], o);
                            }
                        }
                        // lanes whose low bits (those of the plain butterfly steps) are zero publish their channel
                        constexpr int PLAIN = 32 / C::NS - 1;            // mask of the butterfly-only lane bits
                        if (a.stats_batch) {
                            acc_s += (double)ssum[0]; acc_q += (double)ssq[0]; acc_chan = chan;
                        } else if ((lane & PLAIN) == 0) {
                            double* dst = a.stats + ((((size_t)b * (C::TILES * 4) + t * 4 + q) * COUT) + wg * HALF + chan) * 2;
                            dst[0] = (double)ssum[0]; dst[1] = (double)ssq[0];
                        }
                    }
                }
            } else {
                // relu(conv3x3 + bias) for this group's 16 middle channels -> hi / lo -> A operand of the 1x1 in A buffer 0
                // (every MMA of this tile has completed: d_full), columns [16 wg, 16 wg + 16) of hi [0, 32) and lo [32, 64)
                {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float h = fmaxf(__uint_as_float(r[0][0][c]) + __uint_as_float(r[0][1][c]) + bias_r[c], 0.f);
                        hi[c] = (__float_as_uint(h) + 0x1000u) & 0xFFFFE000u;
                        lo[c] = __float_as_uint(h - __uint_as_float(hi[c]));
                    }
                    tmem_st16(lane_base + (uint32_t)(wg * 16), hi);
                    tmem_st16(lane_base + (uint32_t)(C::K2 + wg * 16), lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a2_full);
                // skip connection: the block input at this pixel is the centre tap of the input tile (raw, no ReLU)
                constexpr int H2 = C::COUT2 / 2;
                float xin[H2];
#pragma unroll
                for (int c = 0; c < H2; ++c) xin[c] = tin[((wg * H2 + c) * C::RIN + prow + 1) * W + ox];
                __syncwarp();
                if (lane == 0) mbar_arrive(in_empty + 8u * stage);
                mbar_wait(d2_full, (uint32_t)(it & 1));
                tc_fence_after();
                uint32_t m2[H2], s2[H2];
                const uint32_t d2 = lane_base + (uint32_t)C::D_COL + (uint32_t)(wg * H2);
                TmemLd<H2>::ld(d2, m2);
                TmemLd<H2>::ld(d2 + (uint32_t)C::COUT2, s2);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty);
                float* yp = a.y + ((size_t)b * C::COUT2 + wg * H2) * (C::HO * C::WO) + pix;
#pragma unroll
                for (int c = 0; c < H2; ++c) {
                    // (no ReLU here: out_relu named the one between the two convolutions, applied above)
                    yp[(size_t)c * (C::HO * C::WO)] =
                        __uint_as_float(m2[c]) + __uint_as_float(s2[c]) + __ldg(a.bias2 + wg * H2 + c) + xin[c];
                }
            }
        };

        // DG: the epilogue reads the gate (and skip) tensor at the tile's output pixels -- cold lines whose DRAM latency
        // nothing hid (the gather warps ARE the epilogue warps: +14 .. +30 us per launch at batch 256).  Each thread
        // therefore asks for one or two of the tile's lines in L2 before it starts on the tile.
        [[maybe_unused]] auto prefetch_out = [&](const float* base, int64_t tile) {
            const int b = (int)(tile / C::TILES), t = (int)(tile % C::TILES);
            constexpr int LPC = C::CT ? 16 : 4;                                  // 128-byte lines per channel and tile
            constexpr int NL = (C::CT ? C::CT_C : COUT) * LPC;
            const int me = warp * 32 + lane;
#pragma unroll
            for (int l0 = 0; l0 < NL; l0 += 256) {
                const int l = l0 + me;
                if (l < NL) {
                    const int ch = l / LPC, seg = l % LPC;
                    const float* p = C::CT ? base + ((size_t)b * C::CT_C + ch) * (4 * C::HO * C::WO) + (size_t)t * 512 + seg * 32
                                           : base + ((size_t)b * COUT + ch) * (C::HO * C::WO) + (size_t)t * 128 + seg * 32;
                    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
                }
            }
        };
        int it = 0;
        int64_t prev_tile = -1;
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
            const int stage = it % C::NSTAGE;
            if constexpr (C::DG) {
                if (dg.mask_src) prefetch_out(dg.mask_src, tile);
                if (a.skip) prefetch_out(a.skip, tile);
            }
            mbar_wait(in_full + 8u * stage, (uint32_t)((it / C::NSTAGE) & 1));
            const float* tin = reinterpret_cast<const float*>(stage0 + (size_t)stage * C::STAGE_BYTES);
            const int tb = (int)(tile / C::TILES), trow0 = (int)(tile % C::TILES) * C::TH * S - C::PAD;
            if constexpr (C::DBUF) {
                // this group's first chunk of the tile lets the MMA warp start on it; the previous tile's epilogue
                // (its MMAs have had a whole chunk to drain) comes next, then the remaining chunks
                if (wg < C::NCH) produce(tin, wg, tb, trow0);
                if (prev_tile >= 0) epilogue(prev_tile, it - 1, nullptr, 0);
#pragma unroll 1
                for (int ch = wg + 2; ch < C::NCH; ch += 2) produce(tin, ch, tb, trow0);
                __syncwarp();
                if (lane == 0) mbar_arrive(in_empty + 8u * stage);  // this warp is done with the input stage
                prev_tile = tile;
            } else {
#pragma unroll 1
                for (int ch = C::SPLIT ? 0 : wg; ch < C::NCH; ch += C::SPLIT ? 1 : 2) produce(tin, ch, tb, trow0, it);
                if constexpr (!C::FUSE) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(in_empty + 8u * stage);
                }
                epilogue(tile, it, tin, stage);
            }
        }
        if constexpr (C::DBUF) {
            if (prev_tile >= 0) epilogue(prev_tile, it - 1, nullptr, 0);
        }
        if constexpr (C::STATS) {
            // whole-batch statistics: the warps' sums meet in shared memory ([quadrant (x py)][channel]) ...
            constexpr int PLAIN = 32 / C::NS - 1;
            if (a.stats && a.stats_batch && (lane & PLAIN) == 0)         // (every CTA of the grid has at least one tile)
                stat_red[C::CT ? (q * 2 + wg) * C::CS + acc_chan : q * C::CS + wg * HALF + acc_chan] = make_double2(acc_s, acc_q);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (C::STATS) {
        // ... and leave ONE row per CTA, added in a fixed order (296 rows per launch for the finalize to fold).  The two
        // helper warps do it: they have no output stores in flight, so their __threadfence below is cheap.
        const bool rows_on = a.stats && a.stats_batch;
        if (rows_on && tid >= 256 && tid - 256 < C::CS) {
            const int c = tid - 256;
            double s = 0.0, q2 = 0.0;
#pragma unroll
            for (int r = 0; r < C::WARP_ROWS; ++r) { const double2 v = stat_red[r * C::CS + c]; s += v.x; q2 += v.y; }
            double* dst = a.stats + ((size_t)blockIdx.x * C::CS + c) * 2;
            dst[0] = s; dst[1] = q2;
            if (fin.mode != 0) __threadfence();
        }
        if (fin.mode != 0 && rows_on) {
            // the last CTA to get here folds every CTA's row and finalises the BatchNorm (no launch of its own, no wait
            // for SM room beside the weight-gradient stream); fixed order: group g adds rows g, g + G, ..., then the groups
            __shared__ bool last;
            __syncthreads();
            if (tid == 256) last = (atomicAdd(fin.ticket, 1u) == gridDim.x - 1);
            __syncthreads();
            if (last) {
                __threadfence();
                constexpr int G = 256 / C::CS;
                double2* part = reinterpret_cast<double2*>(stage0);          // [G][CS]: the input stages are free by now
                if (tid < 256) {
                    const int c = tid % C::CS, g = tid / C::CS;
                    double s = 0.0, q2 = 0.0;
                    const double2* rows = reinterpret_cast<const double2*>(a.stats) + c;
                    const int n = (int)gridDim.x;
                    for (int r0 = g; r0 < n; r0 += 8 * G) {                  // eight independent loads in flight
                        double2 v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int r = r0 + j * G;
                            v[j] = r < n ? __ldcg(rows + (size_t)r * C::CS) : make_double2(0.0, 0.0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) { s += v[j].x; q2 += v[j].y; }
                    }
                    part[g * C::CS + c] = make_double2(s, q2);
                }
                __syncthreads();
                if (tid < C::CS) {
                    const int c = tid;
                    double s = 0.0, q2 = 0.0;
#pragma unroll
                    for (int g = 0; g < G; ++g) { const double2 v = part[g * C::CS + c]; s += v.x; q2 += v.y; }
                    if (fin.mode == 1) {             // bn_finalize_batch_kernel
                        const double mean = s / fin.cnt;
                        double var = q2 / fin.cnt - mean * mean;
                        if (var < 0.0) var = 0.0;
                        const float invstd = (float)(1.0 / sqrt(var + (double)fin.eps));
                        const float sc = fin.gamma[c] * invstd;
                        fin.scale[c] = sc;
                        fin.shift[c] = fin.beta[c] - (float)mean * sc;
                        if (fin.save_mean) { fin.save_mean[c] = (float)mean; fin.save_invstd[c] = invstd; }
                        if (fin.running_mean) {
                            const double unbiased = fin.cnt > 1.0 ? var * fin.cnt / (fin.cnt - 1.0) : var;
                            fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
                            fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
                        }
                    } else {                         // bn_bwd_finalize_batch_kernel: s = sum g, q2 = sum g * y
                        const double N = fin.cnt;
                        const double mu = fin.mean[c], is = fin.invstd[c], gm = fin.gamma[c];
                        const double dbeta = s;
                        const double dgamma = is * (q2 - mu * s);
                        fin.A[c] = (float)(gm * is);
                        fin.Bc[c] = (float)(-gm * is * is * dgamma / N);
                        fin.Cc[c] = (float)(-gm * is * dbeta / N + gm * is * is * mu * dgamma / N);
                        fin.dgamma[c] = (float)dgamma;
                        fin.dbeta[c] = (float)dbeta;
                    }
                }
                if (tid == 0) *fin.ticket = 0u;
            }
        }
    }
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)C::TMEM_COLS);
    }
}

// ---- weight image ---------------------------------------------------------------------------------------------
// w_packed [Cin][ks][ks][Cout] -> [NT][2*Cout][32] K-major rows, 16-byte chunks XOR-swizzled by (row & 7); row n < Cout:
// tf32(w) of channel n, row Cout + n: w - tf32(w); k = ky*(ks*Cin) + ci*ks + kx, zero beyond K.
__global__ void pack_tm_kernel(const float* __restrict__ w, float* __restrict__ out, int cin, int cout, int ks) {
    pdl_wait();
    const int KC = ks * cin, K = ks * KC, NT = (K + 31) / 32, NR = 2 * cout;
    const int total = NT * 32 * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / cout, co = i - k * cout;
        float hi = 0.f, lo = 0.f;
        if (k < K) {
            const int ky = k / KC, r = k - ky * KC, ci = r / ks, kx = r - ci * ks;
            const float wv = __ldg(w + ((size_t)(ci * ks + ky) * ks + kx) * cout + co);
            hi = __uint_as_float(tf32_rna_bits(wv));
            lo = wv - hi;
        }
        const int tile = k >> 5, q = (k & 31) >> 2, e = k & 3;
        const size_t tb = (size_t)tile * NR * 32;
        out[tb + (size_t)co * 32 + (((q ^ (co & 7)) << 2) | e)] = hi;
        const int n2 = cout + co;
        out[tb + (size_t)n2 * 32 + (((q ^ (n2 & 7)) << 2) | e)] = lo;
    }
}

// the same for every layer of a model in ONE launch (blockIdx.y = layer): a training step re-packs every step
struct PackTmJobs { int n; TmPackJob j[TM_PACK_MAX]; };
__global__ void pack_tm_multi_kernel(const PackTmJobs jobs) {
    pdl_wait();
    const TmPackJob& jb = jobs.j[blockIdx.y];
    const float* __restrict__ w = jb.w;
    float* __restrict__ out = jb.out;
    // ct: the source is a transposed convolution's [Cin][4][4][Cout]; the image is that of the 3x3 convolution over the
    // input grid with 4 * Cout phase channels n = (2 py + px) * Cout + co (conv_tm.cu, CT form)
    const int cin = jb.cin, ks = jb.ct ? 3 : jb.ks, cout = jb.ct ? 4 * jb.cout : jb.cout;
    const int KC = ks * cin, K = ks * KC, NT = (K + 31) / 32, NR = 2 * cout;
    const int total = NT * 32 * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / cout, co = i - k * cout;
        float hi = 0.f, lo = 0.f;
        if (k < K) {
            const int ky = k / KC, r = k - ky * KC, ci = r / ks, kx = r - ci * ks;
            float wv;
            if (jb.ct) {
                const int ph = co / jb.cout, c = co - ph * jb.cout, py = ph >> 1, px = ph & 1;
                // output row 2y + py takes input row y + (ky - 1) through tap  py + 1 - 2 (ky - 1)  of the 4x4 kernel
                const int ty = py + 3 - 2 * ky, tx = px + 3 - 2 * kx;
                const bool on = (ky - 1 == -1 + py || ky - 1 == py) && (kx - 1 == -1 + px || kx - 1 == px);
                wv = on ? __ldg(w + ((size_t)(ci * 4 + ty) * 4 + tx) * jb.cout + c) : 0.f;
            } else {
                wv = __ldg(w + ((size_t)(ci * ks + ky) * ks + kx) * cout + co);
            }
            hi = __uint_as_float(tf32_rna_bits(wv));
            lo = wv - hi;
        }
        const int tile = k >> 5, q = (k & 31) >> 2, e = k & 3;
        const size_t tb = (size_t)tile * NR * 32;
        out[tb + (size_t)co * 32 + (((q ^ (co & 7)) << 2) | e)] = hi;
        const int n2 = cout + co;
        out[tb + (size_t)n2 * 32 + (((q ^ (n2 & 7)) << 2) | e)] = lo;
    }
}

PFN_cuTensorMapEncodeTiled tm_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }();
    return fn;
}

template <class C>
int launch_tm(const ConvTmArgs& a, cudaStream_t st) {
    PFN_cuTensorMapEncodeTiled enc = tm_encoder();
    DMB_CHECK(enc != nullptr, "conv_tm: cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    const cuuint64_t gdim[4] = {(cuuint64_t)C::W, (cuuint64_t)C::H, (cuuint64_t)a.B, (cuuint64_t)C::CIN};
    const cuuint64_t gstr[3] = {(cuuint64_t)C::W * 4, (cuuint64_t)C::W * C::H * C::CIN * 4, (cuuint64_t)C::W * C::H * 4};
    const cuuint32_t box[4] = {(cuuint32_t)C::W, (cuuint32_t)C::RIN, 1u, (cuuint32_t)C::CIN};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.x), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DMB_CHECK(r == CUDA_SUCCESS, "conv_tm: cuTensorMapEncodeTiled failed (%d)", (int)r);
    TmKArgs k{};
    k.wtm = a.wtm; k.bias = a.bias; k.y = a.y; k.skip = a.skip; k.bias2 = a.bias2;
    k.in_scale = a.in_scale; k.in_shift = a.in_shift; k.in_per_sample = a.in_per_sample; k.stats = a.stats;
    k.stats_batch = a.stats_batch;
    TmDgArgs d{a.mask_src, a.mask_s, a.mask_t, a.stat_src, a.in_a, a.in_b, a.in_c};
    CUtensorMap map2 = map;
    if constexpr (C::DUAL) {
        if (a.in_a) {
            DMB_CHECK(a.x2 && a.in_b && a.in_c && !(reinterpret_cast<uintptr_t>(a.x2) & 15), "conv_tm: the dual-tensor load needs x2 (16-byte aligned) and all three coefficient tables");
            const CUresult r2 = enc(&map2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.x2), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            DMB_CHECK(r2 == CUDA_SUCCESS, "conv_tm: cuTensorMapEncodeTiled failed (%d)", (int)r2);
        }
    } else {
        DMB_CHECK(!a.in_a, "conv_tm: this layer shape has no dual-tensor load (materialise the BatchNorm-backward gradient)");
    }
    k.ntiles = (int64_t)a.B * C::TILES;
    k.in_relu = a.in_relu; k.out_relu = a.out_relu;
    { const char* e = getenv("DMB_TM_DBG"); k.dbg = e ? atoi(e) : 0; }
    auto kern = conv_tm_kernel<C>;
    int dev = 0;
    DMB_CUDA(cudaGetDevice(&dev));
    DMB_CHECK(dev >= 0 && dev < 64, "conv_tm: device index %d out of range", dev);
    static bool configured[64] = {false};
    if (!configured[dev]) {
        DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[dev] = true;
    }
    int sms = 148;
    DMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (sms > TM_MAX_SMS) sms = TM_MAX_SMS;
    const int64_t grid = std::min<int64_t>(k.ntiles, (int64_t)sms * C::CTAS);
    DMB_CHECK(grid > 0, "conv_tm: empty launch");
    if (a.stat_rows) *a.stat_rows = (int)grid * C::STAT_ROWS;
    TmFinArgs fin{};
    if (a.fin) {
        if constexpr (C::STATS) {
            DMB_CHECK(a.fin->ticket && a.stats && a.stats_batch && (a.fin->mode == 1 || a.fin->mode == 2), "conv_tm: the in-kernel finalize needs a ticket and whole-batch statistics");
            fin = *a.fin;
        } else {
            DMB_CHECK(false, "conv_tm: this form leaves no statistics to finalise");
        }
    }
    DMB_LAUNCH((kern), (unsigned)grid, TM_THREADS, C::SMEM, st, map, k, d, map2, fin);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace

bool conv_tm_supported(int cin, int cout, int ks, int stride, int H, int W) {
    if (H != W) return false;
    return (ks == 4 && stride == 2 && cin == 8 && cout == 16 && W == 64) ||
           (ks == 4 && stride == 2 && cin == 16 && cout == 16 && W == 32) ||
           (ks == 3 && stride == 1 && cin == 16 && cout == 16 && W == 16) ||
           (ks == 3 && stride == 1 && cin == 16 && cout == 32 && W == 16) ||
           (ks == 1 && stride == 1 && cin == 32 && cout == 16 && W == 16);
}

// shapes of the data-gradient form, named by the data-gradient convolution's own (cin, cout): the default model's
// residual 1x1 (16 -> 32) and 3x3 (32 -> 16), enc.10 (16 -> 16) and the stride-2 convolutions that back-propagate through
// a ConvTranspose2d 16 -> 8 (output 64 x 64 or 32 x 32: the default decoder's first layer) and 16 -> 16 (output 32 x 32)
bool conv_tm_dg_supported(int cin, int cout, int ks, int stride, int H, int W) {
    if (H != W) return false;
    return (ks == 1 && stride == 1 && cin == 16 && cout == 32 && W == 16) ||
           (ks == 3 && stride == 1 && cin == 32 && cout == 16 && W == 16) ||
           (ks == 3 && stride == 1 && cin == 16 && cout == 16 && W == 16) ||
           (ks == 4 && stride == 2 && cin == 8 && cout == 16 && W == 64) ||
           (ks == 4 && stride == 2 && cin == 8 && cout == 16 && W == 32) ||
           (ks == 4 && stride == 2 && cin == 16 && cout == 16 && W == 32);
}

// transposed form (ConvTranspose2d 4x4 s2 p1, cin -> cout, input H x W): the default decoder's first layer (plain: bias,
// ReLU on store) and the data gradients of the encoder's stride-2 convolutions 16 -> 16 @32 and 8 -> 16 @64 (dg)
bool conv_tm_ct_supported(int cin, int cout, int H, int W, bool dg) {
    if (H != W) return false;
    if (!dg) return cin == 16 && cout == 8 && W == 16;
    return (cin == 16 && cout == 16 && W == 16) || (cin == 16 && cout == 8 && W == 32);
}

// ... of which these apply a BatchNorm backward on load (x2 / in_a / in_b / in_c)
bool conv_tm_dg_dual(int cin, int cout, int ks, int stride, int H, int W) {
    return conv_tm_dg_supported(cin, cout, ks, stride, H, W) && stride == 1 && !(ks == 3 && cin == 32);
}

int conv_tm_bands(int cin, int cout, int ks, int stride, int H, int W) {
    if (!conv_tm_supported(cin, cout, ks, stride, H, W)) return 0;
    return (H / stride) * (W / stride) / 32;        // one partial row per warp: 32 output pixels
}

int64_t conv_tm_weight_floats(int cin, int cout, int ks) {
    const int K = ks * ks * cin;
    return (int64_t)((K + 31) / 32) * 2 * cout * 32;
}

int pack_tm_weights_ct(const float* w_packed, float* out, int cin, int cout, cudaStream_t st) {
    TmPackJob j{w_packed, out, cin, cout, 4, 1};
    return pack_tm_weights_multi(&j, 1, st);
}

int pack_tm_weights(const float* w_packed, float* out, int cin, int cout, int ks, cudaStream_t st) {
    const int total = ((ks * ks * cin + 31) / 32) * 32 * cout;
    DMB_LAUNCH((pack_tm_kernel), (total + 255) / 256, 256, 0, st, w_packed, out, cin, cout, ks);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int pack_tm_weights_multi(const TmPackJob* jobs, int n, cudaStream_t st) {
    for (int i0 = 0; i0 < n; i0 += TM_PACK_MAX) {
        PackTmJobs pj{};
        pj.n = (n - i0 < TM_PACK_MAX) ? n - i0 : TM_PACK_MAX;
        int most = 0;
        for (int i = 0; i < pj.n; ++i) {
            pj.j[i] = jobs[i0 + i];
            const int total = pj.j[i].ct ? ((9 * pj.j[i].cin + 31) / 32) * 32 * 4 * pj.j[i].cout
                                         : ((pj.j[i].ks * pj.j[i].ks * pj.j[i].cin + 31) / 32) * 32 * pj.j[i].cout;
            if (total > most) most = total;
        }
        DMB_LAUNCH((pack_tm_multi_kernel), dim3((most + 255) / 256, pj.n), 256, 0, st, pj);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
    }
    return 0;
}

int conv_tm(const ConvTmArgs& a, cudaStream_t st) {
    DMB_CHECK(!(reinterpret_cast<uintptr_t>(a.x) & 15) && !(reinterpret_cast<uintptr_t>(a.wtm) & 15),
              "conv_tm: x and the weight image must be 16-byte aligned");
    DMB_CHECK(a.B > 0, "conv_tm: empty batch");
    if (a.ct) {
        // ConvTranspose2d 4x4 s2 p1 as a 3x3 convolution on the input grid + pixel shuffle
        DMB_CHECK(conv_tm_ct_supported(a.Cin, a.Cout, a.H, a.W, a.dg != 0), "conv_tm: unsupported transposed layer %d->%d @%dx%d%s",
                  a.Cin, a.Cout, a.H, a.W, a.dg ? " (data gradient)" : "");
        DMB_CHECK(!a.bn && !a.bias2 && !a.in_relu && !a.in_scale && !a.skip, "conv_tm: the transposed form takes a plain input and no skip");
        DMB_CHECK(!(reinterpret_cast<uintptr_t>(a.y) & 7), "conv_tm: y must be 8-byte aligned");
        if (!a.dg) {
            DMB_CHECK(!a.mask_src && !a.stat_src && !a.stats && !a.in_a, "conv_tm: gate / sums belong to the data-gradient form");
            return launch_tm<TM<3, 1, 16, 32, 16, false, false, false, false, false, true>>(a, st);
        }
        DMB_CHECK(!a.out_relu, "conv_tm: no ReLU on store in the data-gradient form");
        DMB_CHECK((a.mask_s == nullptr) == (a.mask_t == nullptr) && (a.mask_src || !a.mask_s), "conv_tm: mask affine without a mask");
        DMB_CHECK(!a.stats || (a.stats_batch && a.stat_rows), "conv_tm: the data-gradient form leaves whole-batch sums");
        DMB_CHECK(a.stats || !a.stat_src, "conv_tm: stat_src without stats");
        if (a.Cout == 16) return launch_tm<TM<3, 1, 16, 64, 16, false, false, false, true, true, true>>(a, st);
        return launch_tm<TM<3, 1, 16, 32, 32, false, false, false, true, true, true>>(a, st);
    }
    DMB_CHECK(a.dg ? conv_tm_dg_supported(a.Cin, a.Cout, a.ks, a.stride, a.H, a.W)
                   : conv_tm_supported(a.Cin, a.Cout, a.ks, a.stride, a.H, a.W),
              "conv_tm: unsupported layer %dx%d s%d %d->%d @%dx%d", a.ks, a.ks, a.stride, a.Cin, a.Cout, a.H, a.W);
    if (a.dg) {
        // data gradient of a training step: plain input, gate / skip / BatchNorm-backward sums in the epilogue
        DMB_CHECK(!a.bn && !a.bias2 && !a.in_relu && !a.out_relu && !a.in_scale, "conv_tm: the data-gradient form takes a plain input");
        DMB_CHECK((a.mask_s == nullptr) == (a.mask_t == nullptr) && (a.mask_src || !a.mask_s), "conv_tm: mask affine without a mask");
        DMB_CHECK(!a.stats || (a.stats_batch && a.stat_rows), "conv_tm: the data-gradient form leaves whole-batch sums");
        DMB_CHECK(a.stats || !a.stat_src, "conv_tm: stat_src without stats");
        if (a.ks == 1) return launch_tm<TM<1, 1, 16, 32, 16, false, false, false, true, true>>(a, st);
        if (a.ks == 3 && a.Cin == 32) return launch_tm<TM<3, 1, 32, 16, 16, false, false, false, true>>(a, st);
        if (a.ks == 3) return launch_tm<TM<3, 1, 16, 16, 16, false, false, false, true, true>>(a, st);
        if (a.Cin == 8 && a.W == 64) return launch_tm<TM<4, 2, 8, 16, 64, false, false, false, true>>(a, st);
        if (a.Cin == 8) return launch_tm<TM<4, 2, 8, 16, 32, false, false, false, true>>(a, st);
        return launch_tm<TM<4, 2, 16, 16, 32, false, false, false, true>>(a, st);
    }
    DMB_CHECK(!a.mask_src && !a.stat_src, "conv_tm: gate / stat_src belong to the data-gradient form (dg = 1)");
    if (a.bn) {
        // train-mode BatchNorm around the layer: transform on load, raw output, statistics partials
        DMB_CHECK(!a.bias2 && !a.skip && !a.out_relu, "conv_tm: the BatchNorm form stores the raw convolution output");
        DMB_CHECK((a.in_scale == nullptr) == (a.in_shift == nullptr), "conv_tm: scale / shift come together");
        if (a.ks == 4 && a.Cin == 8) return launch_tm<TM<4, 2, 8, 16, 64, false, false, true>>(a, st);
        if (a.ks == 4) return launch_tm<TM<4, 2, 16, 16, 32, false, false, true>>(a, st);
        if (a.ks == 3 && a.Cout == 16) return launch_tm<TM<3, 1, 16, 16, 16, false, false, true>>(a, st);
        if (a.ks == 3) return launch_tm<TM<3, 1, 16, 32, 16, false, false, true>>(a, st);
        return launch_tm<TM<1, 1, 32, 16, 16, false, false, true>>(a, st);
    }
    DMB_CHECK(!a.in_scale && !a.stats, "conv_tm: a pending affine / statistics need the BatchNorm form (bn = 1)");
    if (a.ks == 3 && a.Cout == 32) {
        // the residual block's 3x3: ReLU on load compiled in or out; with bias2 the whole layer is fused
        if (a.bias2) {
            DMB_CHECK(!a.skip && a.out_relu && a.in_relu, "conv_tm: the fused residual layer applies ReLU on load and "
                      "between the two convolutions and takes its skip from the input tile");
            return launch_tm<TM<3, 1, 16, 32, 16, true, true>>(a, st);
        }
        return a.in_relu ? launch_tm<TM<3, 1, 16, 32, 16, false, true>>(a, st) : launch_tm<TM<3, 1, 16, 32, 16>>(a, st);
    }
    DMB_CHECK(!a.in_relu && !a.bias2, "conv_tm: ReLU on load / the fused tail exist for the 3x3 16 -> 32 layer only");
    if (a.ks == 4 && a.Cin == 8) return launch_tm<TM<4, 2, 8, 16, 64>>(a, st);
    if (a.ks == 4) return launch_tm<TM<4, 2, 16, 16, 32>>(a, st);
    if (a.ks == 3) return launch_tm<TM<3, 1, 16, 16, 16>>(a, st);
    return launch_tm<TM<1, 1, 32, 16, 16>>(a, st);
}

}  // namespace dmb
