// Tensor-core convolution for the wide (64-channel) configurations: implicit GEMM on tcgen05 with 3xTF32 operand
// splitting, so that the results stay within fp32 round-off of the reference (BASELINE configs[3]; the reference
// layers are HiddenStateExtractor/vq_vae.py:279-289 and the ResidualBlock convs at :203-209).
//
//   D[128 pixels x Cout] += A[128 pixels x 32 channels of one tap] * B[Cout x 32]^T        per k-block
//
// * activations are NHWC so that one filter tap of 32 input channels is a K-major operand row of 128 bytes: ONE
//   cp.async.bulk.tensor.4d per k-block over the tensor viewed as (C, W, H, B) with box (32, Wo*S, TH*S, 1) and
//   element strides (1, S, S, 1) lands the im2col tile in the 128-byte-swizzled canonical layout tcgen05 reads;
//   the zero padding is the TMA out-of-bounds fill;
// * weights are pre-split (hi = tf32(w), lo = tf32(w - hi)) and pre-swizzled by `pack_tc_weights`, one plain
//   cp.async.bulk per k-block;
// * warps 0-7 split the activation tile in shared memory (optional ReLU on load, hi in place, lo beside it) and
//   later run the epilogue; warp 8 lane 0 is the TMA producer; warp 9 lane 0 issues the MMAs: per K=8 step ONE
//   N = 2*Cout MMA  a_hi * [b_hi; b_lo]  (the hi and lo weight tiles are adjacent rows of one operand, so a_hi
//   crosses the shared-memory port once) and one N = Cout MMA  a_lo * b_hi, into TMEM accumulators laid out as
//   [main | small] column blocks, and commits stage release / accumulator-ready to mbarriers;
// * epilogue: tcgen05.ld (lane = pixel, column = output channel) -> bias, skip, ReLU -> NCHW (coalesced over the
//   128 consecutive pixels of the tile) or NHWC (16-byte stores) output.
#include "common.cuh"

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
// per-thread wait (producer / MMA lanes, and each splitter thread)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, FP32 accumulation.
// Converged-warp forms: the issuing warp stays converged and one elected lane executes the instruction.  Issued from a
// divergent `if (lane == 0)` region ptxas wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY loop.
__device__ __forceinline__ void tc_mma_tf32_elect(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        done = mbar_try(bar, parity) ? 1u : 0u;
    } while (!__all_sync(0xffffffffu, done != 0));
}
// 32 lanes x 32 consecutive columns: thread t of the warp gets lane (base lane + t), v[j] = column (base col + j)
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
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format TF32 (2) [7,10)/[10,13), K-major both,
// N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int TC_SPLIT_WARPS = 8;
constexpr int TC_THREADS = (TC_SPLIT_WARPS + 2) * 32;
constexpr int TC_A_BYTES = 128 * 128;        // 128 pixels x 32 channels fp32

template <int COUT, int NSTAGE, int NACC>
struct TcCfg {
    static constexpr int B_BYTES = 2 * COUT * 128;               // hi + lo
    static constexpr int STAGE = 2 * TC_A_BYTES + B_BYTES;       // A, A_lo, B_hi, B_lo
    static constexpr int BAR_BYTES = 256;
    static constexpr size_t SMEM = (size_t)NSTAGE * STAGE + BAR_BYTES + 1024;
    // accumulators in TMEM: NACC pairs (main = a_hi*b_hi, small = a_hi*b_lo + a_lo*b_hi); MMA k of every k-block goes
    // to pair k % NACC.  The tensor core truncates when it aligns and adds, so the error of a long
    // accumulation chain is a bias that grows with its length and with the magnitude of the running sum; short
    // chains summed in fp32 round-to-nearest by the epilogue keep the result at FFMA-chain accuracy.
    static constexpr int ACC_COLS = NACC * 2 * COUT;             // accumulator k: columns [2k*Cout, +Cout) main, next Cout small
    static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    static_assert(NACC == 1 || NACC == 2 || NACC == 4, "NACC must divide the four MMAs of a k-block");
    static_assert(ACC_COLS <= 512, "accumulators do not fit tensor memory");
    static_assert(COUT == 32 || COUT == 64, "Cout must be 32 or 64");
};

struct TcKArgs {
    const float* wtc;      // [k-block][hi|lo][Cout][32] swizzled
    const float* bias;     // [Cout]
    float* y;
    const float* skip;
    int Ho, Wo, TH, tiles_per_img, nkb, nhalf, ks, stride, pad;
    int in_relu, out_relu, out_nhwc, skip_nhwc;
};

template <int COUT, int NSTAGE, int NACC>
__global__ void __launch_bounds__(TC_THREADS, NSTAGE == 2 ? 2 : 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcKArgs a) {
    using C = TcCfg<COUT, NSTAGE, NACC>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + (uint32_t)NSTAGE * C::STAGE;
    // barrier map: full[s] = bars + 8s, split[s] = +64 + 8s, empty[s] = +128 + 8s, accum = +192, tmem slot = +200
    const uint32_t bar_full = bars, bar_split = bars + 64, bar_empty = bars + 128, bar_accum = bars + 192;
    const uint32_t tmem_slot = bars + 200;
    static_assert(NSTAGE <= 8, "barrier map holds eight stages");

    // (the shuffle tells ptxas that the warp index is warp-uniform: role branches become uniform branches and the MMA
    // warp's descriptor arithmetic stays in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    constexpr int W_TMA = TC_SPLIT_WARPS, W_MMA = TC_SPLIT_WARPS + 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_full + 8u * s, 1u);
            mbar_init(bar_split + 8u * s, (uint32_t)TC_SPLIT_WARPS);
            mbar_init(bar_empty + 8u * s, 1u);
        }
        mbar_init(bar_accum, 1u);
        fence_barrier_init();
    }
    if (warp == W_TMA && lane == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmap) : "memory");
    if (warp == W_MMA) tmem_alloc(tmem_slot, (uint32_t)C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + (size_t)NSTAGE * C::STAGE + 200);
    pdl_wait();

    const int tile = blockIdx.x;
    const int b = tile / a.tiles_per_img;
    const int oy0 = (tile - b * a.tiles_per_img) * a.TH;
    const int nkb = a.nkb;

    if (warp == W_TMA) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NSTAGE, it = kb / NSTAGE;
                mbar_wait(bar_empty + 8u * s, (uint32_t)((it & 1) ^ 1));
                const int tap = kb / a.nhalf, half = kb - tap * a.nhalf;
                const int kh = tap / a.ks, kw = tap - kh * a.ks;
                const uint32_t st = base + (uint32_t)s * C::STAGE;
                mbar_expect_tx(bar_full + 8u * s, (uint32_t)(TC_A_BYTES + C::B_BYTES));
                tma_load_4d(st, &tmap, bar_full + 8u * s, half * 32, kw - a.pad, a.stride * oy0 + kh - a.pad, b);
                bulk_load_1d(st + 2u * TC_A_BYTES, a.wtc + (size_t)kb * (C::B_BYTES / 4), (uint32_t)C::B_BYTES,
                             bar_full + 8u * s);
            }
        }
    } else if (warp == W_MMA) {
        {   // whole warp, converged; one elected lane per tcgen05 instruction
            constexpr uint32_t idesc2 = make_idesc_tf32(128, 2 * COUT), idesc1 = make_idesc_tf32(128, COUT);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NSTAGE, it = kb / NSTAGE;
                mbar_wait_warp(bar_full + 8u * s, (uint32_t)(it & 1));
                mbar_wait_warp(bar_split + 8u * s, (uint32_t)(it & 1));
                tc_fence_after();
                const uint32_t st = base + (uint32_t)s * C::STAGE;
                const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + TC_A_BYTES);
                const uint64_t b_hl = make_desc_sw128(st + 2u * TC_A_BYTES);   // rows [0,Cout) hi, [Cout,2Cout) lo
#pragma unroll
                for (int k = 0; k < 4; ++k) {                 // 4 x (K = 8 tf32 = 32 bytes) inside the swizzle atom
                    const uint64_t o = (uint64_t)(k * 2);
                    const uint32_t d = tmem_base + (uint32_t)((k % NACC) * 2 * COUT);
                    tc_mma_tf32_elect(d, a_hi + o, b_hl + o, idesc2, (kb != 0 || k >= NACC) ? 1u : 0u);   // [hi*hi | hi*lo]
                    tc_mma_tf32_elect(d + COUT, a_lo + o, b_hl + o, idesc1, 1u);                           // small += lo*hi
                }
                tc_commit_elect(bar_empty + 8u * s);          // frees the stage once these MMAs have read it
            }
            tc_commit_elect(bar_accum);
        }
    } else {
        // ---- splitter: hi = relu?(a) rounded to TF32 (round half away, on the bit pattern) in place, lo = a - hi
        // (exact; at most 13 significant bits, of which the tensor core keeps 11: 2^-21 relative, sign-symmetric)
        const int tid = threadIdx.x;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NSTAGE, it = kb / NSTAGE;
            mbar_wait(bar_full + 8u * s, (uint32_t)(it & 1));
            float* A = reinterpret_cast<float*>(base_ptr + (size_t)s * C::STAGE);
            float* Alo = A + TC_A_BYTES / 4;
            constexpr int PER = TC_A_BYTES / 16 / (TC_SPLIT_WARPS * 32);
            float4 v[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = *reinterpret_cast<const float4*>(A + (i * TC_SPLIT_WARPS * 32 + tid) * 4);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                float h[4], l[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (a.in_relu) x[u] = fmaxf(x[u], 0.f);
                    h[u] = __uint_as_float((__float_as_uint(x[u]) + 0x1000u) & 0xffffe000u);
                    l[u] = x[u] - h[u];
                }
                const int e = (i * TC_SPLIT_WARPS * 32 + tid) * 4;
                *reinterpret_cast<float4*>(A + e) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(Alo + e) = make_float4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_split + 8u * s);
        }
        // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (its pixels) and columns 32 (w / 4) .. +31 (+64 ...)
        mbar_wait(bar_accum, 0u);
        tc_fence_after();
        const int m = (warp & 3) * 32 + lane;                   // pixel of the tile == TMEM lane
        const int HoWo = a.Ho * a.Wo;
        const int64_t pix = (int64_t)oy0 * a.Wo + m;            // pixel inside the image (tiles are full-width rows)
#pragma unroll 1
        for (int c0 = (warp >> 2) * 32; c0 < COUT; c0 += 32 * (TC_SPLIT_WARPS / 4)) {
            uint32_t v[32];
            float o[32];
            const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
            // small cross terms first (their own round-off is 2^-11 of a result ulp), then the main partial sums
            tmem_ld32(lane_base + (uint32_t)COUT, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
#pragma unroll
            for (int q = 1; q < NACC; ++q) {
                tmem_ld32(lane_base + (uint32_t)(q * 2 * COUT + COUT), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] += __uint_as_float(v[j]);
            }
#pragma unroll
            for (int q = 0; q < NACC; ++q) {
                tmem_ld32(lane_base + (uint32_t)(q * 2 * COUT), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] += __uint_as_float(v[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] += __ldg(a.bias + c0 + j);
            if (a.skip) {
                if (a.skip_nhwc) {
                    const float4* sp = reinterpret_cast<const float4*>(a.skip + ((int64_t)b * HoWo + pix) * COUT + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = __ldg(sp + j);
                        o[4 * j] += t.x; o[4 * j + 1] += t.y; o[4 * j + 2] += t.z; o[4 * j + 3] += t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] += __ldg(a.skip + ((int64_t)b * COUT + c0 + j) * HoWo + pix);
                }
            }
            if (a.out_relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
            }
            if (a.out_nhwc) {
                float4* yp = reinterpret_cast<float4*>(a.y + ((int64_t)b * HoWo + pix) * COUT + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) yp[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) a.y[((int64_t)b * COUT + c0 + j) * HoWo + pix] = o[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)C::TMEM_COLS);
    }
}

// ---- NCHW -> NHWC --------------------------------------------------------------------------------------------
// (B, C, HW) -> (B, HW, C); 32 x 32 tiles through shared memory, both sides coalesced
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
    pdl_wait();
    __shared__ float t[32][33];
    const int b = blockIdx.x, c0 = blockIdx.y * 32, p0 = blockIdx.z * 32;     // batch on x: no 65535 limit
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xb = x + (int64_t)b * C * HW;
    float* yb = y + (int64_t)b * C * HW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i;
        t[ty + 8 * i][tx] = xb[(int64_t)c * HW + p0 + tx];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + ty + 8 * i;
        yb[(int64_t)p * C + c0 + tx] = t[tx][ty + 8 * i];
    }
}

// ---- weights: [Cin][k][k][Cout] fp32 -> per k-block (tap, 32-channel half): hi tile, lo tile, each
// [Cout rows][32 floats] with the 16-byte chunk index XOR-ed by (row & 7) (the 128-byte swizzle)
__global__ void __launch_bounds__(256) pack_tc_kernel(const float* __restrict__ w, float* __restrict__ out, int cin,
                                                      int cout, int ks) {
    pdl_wait();
    const int nhalf = cin / 32;
    const int64_t total = (int64_t)ks * ks * cin * cout;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(id % 32);
        int64_t t = id / 32;
        const int co = (int)(t % cout); t /= cout;
        const int kb = (int)t;
        const int tap = kb / nhalf, half = kb - tap * nhalf;
        const int ci = half * 32 + j;
        const float v = w[((int64_t)ci * ks * ks + tap) * cout + co];
        const float hi = tf32_rna(v);
        const float lo = tf32_rna(v - hi);
        const int64_t tile = (int64_t)kb * 2 * cout * 32;
        const int pos = co * 32 + (((j >> 2) ^ (co & 7)) << 2) + (j & 3);
        out[tile + pos] = hi;
        out[tile + (int64_t)cout * 32 + pos] = lo;
    }
}

PFN_cuTensorMapEncodeTiled tc_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }();
    return fn;
}

// Variant choice (0 = automatic).  Two CTAs per SM hide the per-k-block handshake latency better than four stages in
// one CTA, but need <= 256 TMEM columns and <= 113 KB each: that is Cout = 32 with four accumulator pairs, or Cout = 64
// with two -- the latter only for the 1x1 layers, whose accumulation chains are too short to need four.
int tc_env(const char* name) {
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}
int tc_stages() { static const int n = tc_env("DMB_TC_STAGES"); return n; }
int tc_nacc() { static const int n = tc_env("DMB_TC_NACC"); return n; }

template <int COUT, int NSTAGE, int NACC>
int launch_tc(const ConvTcArgs& a, const CUtensorMap& map, const TcKArgs& k, int64_t tiles, cudaStream_t st) {
    using C = TcCfg<COUT, NSTAGE, NACC>;
    auto kern = conv_tc_kernel<COUT, NSTAGE, NACC>;
    static bool configured[64] = {false};
    int dev = 0;
    DMB_CUDA(cudaGetDevice(&dev));
    DMB_CHECK(dev >= 0 && dev < 64, "conv_tc: device index %d out of range", dev);
    if (!configured[dev]) {
        DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[dev] = true;
    }
    DMB_LAUNCH((kern), dim3((unsigned)tiles), TC_THREADS, C::SMEM, st, map, k);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    (void)a;
    return 0;
}

template <int COUT>
int launch_tc_n(const ConvTcArgs& a, const CUtensorMap& map, const TcKArgs& k, int64_t tiles, cudaStream_t st) {
    int ns = tc_stages(), na = tc_nacc();
    if (ns == 0) ns = (COUT == 32 || k.nkb <= 4) ? 2 : 4;
    if (na == 0) na = (COUT == 64 && ns == 2) ? 2 : 4;
    switch (ns * 10 + na) {
        case 21: return launch_tc<COUT, 2, 1>(a, map, k, tiles, st);
        case 22: return launch_tc<COUT, 2, 2>(a, map, k, tiles, st);
        case 24: return launch_tc<COUT, 2, 4>(a, map, k, tiles, st);
        case 31: return launch_tc<COUT, 3, 1>(a, map, k, tiles, st);
        case 32: return launch_tc<COUT, 3, 2>(a, map, k, tiles, st);
        case 34: return launch_tc<COUT, 3, 4>(a, map, k, tiles, st);
        case 41: return launch_tc<COUT, 4, 1>(a, map, k, tiles, st);
        case 42: return launch_tc<COUT, 4, 2>(a, map, k, tiles, st);
        default: return launch_tc<COUT, 4, 4>(a, map, k, tiles, st);
    }
}

}  // namespace

bool conv_tc_supported(int cin, int cout, int ks, int stride, int H, int W) {
    if (!((ks == 1 && stride == 1) || (ks == 3 && stride == 1) || (ks == 4 && stride == 2))) return false;
    if (cin % 32 != 0 || cin <= 0) return false;
    if (!(cout == 32 || cout == 64)) return false;
    if (H % stride || W % stride) return false;
    const int Ho = H / stride, Wo = W / stride;
    if (!(Wo == 8 || Wo == 16 || Wo == 32 || Wo == 64 || Wo == 128)) return false;
    const int TH = 128 / Wo;
    if (Ho % TH) return false;
    if (Wo * stride > 256 || TH * stride > 256) return false;
    return true;
}

int64_t conv_tc_weight_floats(int cin, int cout, int ks) { return 2ll * ks * ks * cin * cout; }

int pack_tc_weights(const float* w_packed, float* out, int cin, int cout, int ks, cudaStream_t st) {
    const int64_t total = (int64_t)ks * ks * cin * cout;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    DMB_LAUNCH((pack_tc_kernel), blocks, 256, 0, st, w_packed, out, cin, cout, ks);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int nchw_to_nhwc(const float* x, float* y, int64_t B, int C, int HW, cudaStream_t st) {
    DMB_CHECK(C % 32 == 0 && HW % 32 == 0, "nchw_to_nhwc: C=%d and HW=%d must be multiples of 32", C, HW);
    DMB_CHECK(B > 0 && B < (1ll << 31) && HW / 32 < 65536, "nchw_to_nhwc: batch %lld / map size %d out of range", (long long)B, HW);
    DMB_LAUNCH((nchw_to_nhwc_kernel), dim3((unsigned)B, C / 32, HW / 32), 256, 0, st, x, y, C, HW);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int conv_tc(const ConvTcArgs& a, cudaStream_t st) {
    DMB_CHECK(conv_tc_supported(a.Cin, a.Cout, a.ks, a.stride, a.H, a.W), "conv_tc: unsupported layer %dx%d s%d %d->%d @%dx%d",
              a.ks, a.ks, a.stride, a.Cin, a.Cout, a.H, a.W);
    DMB_CHECK(!(reinterpret_cast<uintptr_t>(a.x) & 15) && !(reinterpret_cast<uintptr_t>(a.wtc) & 15) &&
              !(reinterpret_cast<uintptr_t>(a.y) & 15) && !(reinterpret_cast<uintptr_t>(a.skip) & 15),
              "conv_tc: pointers must be 16-byte aligned");
    PFN_cuTensorMapEncodeTiled enc = tc_encoder();
    DMB_CHECK(enc != nullptr, "conv_tc: cuTensorMapEncodeTiled is not available from this driver");
    const int S = a.stride, Ho = a.H / S, Wo = a.W / S, TH = 128 / Wo;
    CUtensorMap map;
    const cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)a.Cin * 4, (cuuint64_t)a.W * a.Cin * 4, (cuuint64_t)a.H * a.W * a.Cin * 4};
    const cuuint32_t box[4] = {32u, (cuuint32_t)(Wo * S), (cuuint32_t)(TH * S), 1u};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)S, (cuuint32_t)S, 1u};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.x), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DMB_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    TcKArgs k{};
    k.wtc = a.wtc; k.bias = a.bias; k.y = a.y; k.skip = a.skip;
    k.Ho = Ho; k.Wo = Wo; k.TH = TH; k.tiles_per_img = Ho / TH;
    k.nhalf = a.Cin / 32; k.nkb = a.ks * a.ks * k.nhalf; k.ks = a.ks; k.stride = S; k.pad = (a.ks == 1) ? 0 : 1;
    k.in_relu = a.in_relu; k.out_relu = a.out_relu; k.out_nhwc = a.out_nhwc; k.skip_nhwc = a.skip_nhwc;

    const int64_t tiles = (int64_t)a.B * k.tiles_per_img;
    DMB_CHECK(tiles > 0 && tiles < (1ll << 31), "conv_tc: grid %lld out of range", (long long)tiles);
    switch (a.Cout) {
        case 32: return launch_tc_n<32>(a, map, k, tiles, st);
        default: return launch_tc_n<64>(a, map, k, tiles, st);
    }
}

}  // namespace dmb
