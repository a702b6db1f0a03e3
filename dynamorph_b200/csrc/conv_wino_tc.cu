// Winograd F(2x2, 3x3) on the tensor cores for the 16-channel 3x3 layers of the default configuration in EVAL mode
// (HiddenStateExtractor/vq_vae.py:203-209 ResidualBlock conv3x3 16->32 behind a ReLU, :288 enc.10 16->16; 16x16 maps).
//
// These layers are FP32-FMA bound as direct convolutions on the CUDA cores and too thin (16 input channels, 16-32
// outputs) for an im2col tcgen05 kernel.  In the Winograd domain the work is 16 independent small GEMMs, one per
// transform point xi:
//      M_xi[tile, co] = sum_ci V_xi[tile, ci] * U_xi[co, ci],      V = B^T d B,  U = G g G^T,  Y = A^T M A
// with 2.25x fewer multiplies, and a GEMM is something the tensor cores take: one CTA works on 128 tiles (= two
// patches) at a time, the 16 accumulators M_xi (128 lanes x Cout columns each) fill tensor memory, operands are split
// 3xTF32 (hi/lo; a_lo*u_hi + a_hi*u_lo + a_hi*u_hi) so that the result stays at fp32 round-off.
//   * operand rows are 128 bytes = TWO transform points x 16 input channels, K-major, 128-byte swizzle (the layout
//     conv_tc.cu uses): K steps 0-1 of a row belong to the even xi, 2-3 to the odd one;
//   * 16 transform warps: thread = (one tile, four input channels); its four 4x4 input windows stay in registers
//     while the transform is produced two xi at a time (the transformed inputs of 128 tiles x 16 xi, hi + lo, would be
//     256 KB; one operand tile pair is 32 KB and there are two of them), rounded/split on the bit pattern and stored
//     with 16-byte STS whose lane mapping (a quarter-warp holds tiles j and j+4) is conflict free under the swizzle;
//   * a 17th warp's lane 0 issues the 12 MMAs of a half-chunk when the 16 warps have arrived on the buffer's `full`
//     mbarrier and commits to its `empty` mbarrier: the MMAs of half-chunk h run while h + 1 is transformed; the next
//     pair of patches streams in by cp.async meanwhile; the U tiles (all xi, hi + lo, pre-split and pre-swizzled by
//     pack_wino_tc_weights) stay resident in shared memory for the life of the persistent CTA;
//   * epilogue on all 16 warps: warp = (lane quarter, 8-channel group): tcgen05.ld of its 16 xi x 8 columns, output
//     transform in registers, bias, ReLU, float2 stores to NCHW.
#include "common.cuh"

#include <mutex>

namespace dmb {
namespace {

constexpr int WT_THREADS = 512;
constexpr int WT_CIN = 16;
constexpr int WT_RP = 18;                       // padded row pitch (zero border all round)
constexpr int WT_PLANE = WT_RP * WT_RP + 1;     // 325: odd, so that the 16 channel-lanes of a warp hit 16 banks
constexpr int WT_RAW = 2 * WT_CIN * WT_PLANE;   // two patches

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
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
// The issuing warp stays CONVERGED and one elected lane executes the instruction: issued from a divergent
// `if (lane == 0)` region ptxas wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY loop (~100 clk per MMA, which
// made the issue thread the bottleneck of this kernel: 96 small MMAs per pair of patches).
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
// warp-converged wait: the loop exits on a vote, so the code after it is still uniform for ptxas
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!__all_sync(0xffffffffu, done != 0));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// weights [32 ci][16 co] and bias [16] of the 1x1 convolution fused behind the 3x3 (ResidualBlock tail, vq_vae.py:207-209)
__constant__ float c_w2[32 * 16 + 16];

struct WinoTcArgs {
    const float* x;     // (B, 16, 16, 16) NCHW
    const float* u;     // pack_wino_tc_weights output
    const float* bias;  // [COUT]
    float* y;           // (B, COUT, 16, 16)
    float* y2;          // FUSE: (B, 16, 16, 16) = x + conv1x1(y) + bias2, the residual block's output (y itself is not written)
    int B;
    int in_relu, out_relu;
    int dbg;            // DMB_WINO_DBG bit mask (timing experiments only; results are wrong when set)
};

template <int COUT, int NVB = 2>
struct WtCfg {
    static constexpr int U_FLOATS = 2 * 8 * COUT * 32;          // hi, lo: 8 xi-pairs x COUT rows x 32
    static constexpr int V_FLOATS = NVB * 2 * 128 * 32;         // NVB buffers x (hi, lo) x one xi-pair x 128 tiles x 32
    static constexpr int TMEM_COLS = 16 * COUT;                  // 512 / 256
    static constexpr size_t SMEM = (size_t)(U_FLOATS + V_FLOATS + 2 * WT_RAW) * 4 + 128 + 1024;
    static_assert(SMEM <= 227 * 1024, "does not fit shared memory");
    static_assert(NVB == 2 || NVB == 3, "two or three operand buffers");
    static_assert(COUT == 32 || COUT == 16, "Cout must be 16 or 32");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// barrier among the 16 transform / epilogue warps only (the MMA warp never joins it)
__device__ __forceinline__ void work_sync() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }

template <int COUT, bool FUSE, int NVB>
__global__ void __launch_bounds__(WT_THREADS + 32, 1) conv_wino_tc_kernel(const WinoTcArgs a) {
    static_assert(!FUSE || COUT == 32, "the fused 1x1 tail takes 32 channels in");
    static_assert(!FUSE || NVB == 2, "the fused tail parks 64 KB of mid channels in exactly two operand buffers");
    using C = WtCfg<COUT, NVB>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    float* Us = reinterpret_cast<float*>(bp);                                  // [hi|lo][8 pairs][COUT][32]
    float* Vs = Us + C::U_FLOATS;                                              // [2 buffers][hi|lo][128][32]
    float* raw = Vs + C::V_FLOATS;                                             // [2 buffers][2 patches][16 ci][325]
    uint64_t* bar_mem = reinterpret_cast<uint64_t*>(raw + 2 * WT_RAW + 1);
    bar_mem = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bar_mem) + 7) & ~(uintptr_t)7);
    uint32_t* slot_mem = reinterpret_cast<uint32_t*>(bar_mem + 8);
    const uint32_t us_u = base, vs_u = base + C::U_FLOATS * 4u;
    const uint32_t raw_u = vs_u + C::V_FLOATS * 4u;
    // full[b]: the 16 transform warps have written operand buffer b; empty[b]: the MMAs that read it are complete
    const uint32_t bar_full = smem_u32(bar_mem), bar_empty = bar_full + 32u, slot = smem_u32(slot_mem);

    // (the shuffle tells ptxas that the warp index is warp-uniform: role branches become uniform branches and the MMA
    // warp's descriptor arithmetic stays in uniform registers instead of vector registers + R2UR per instruction)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    constexpr int W_MMA = WT_THREADS / 32;
    if (tid == 0) {
        for (int b = 0; b < NVB; ++b) { mbar_init(bar_full + 8u * b, 16u); mbar_init(bar_empty + 8u * b, 1u); }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(slot, (uint32_t)C::TMEM_COLS);
    for (int i = tid; i < 2 * WT_RAW; i += WT_THREADS + 32) raw[i] = 0.f;      // borders stay zero for good
    pdl_wait();
    for (int i = tid; i < C::U_FLOATS / 4; i += WT_THREADS + 32)
        reinterpret_cast<float4*>(Us)[i] = __ldg(reinterpret_cast<const float4*>(a.u) + i);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot_mem);
    const int npairs = (a.B + 1) / 2;

    if (warp == W_MMA) {
        // ===== MMA issuer warp (converged; one elected lane per instruction): 12 MMAs per half-chunk
        // (2 xi x 2 K-steps x 3 split products)
        {
            constexpr uint32_t idesc = make_idesc_tf32(128, COUT);
            uint32_t pf[NVB] = {};
            for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const int vb = h % NVB;
                    mbar_wait_warp(bar_full + 8u * vb, pf[vb]); pf[vb] ^= 1u;
                    tc_fence_after();
                    const uint64_t a_hi = make_desc_sw128(vs_u + (uint32_t)vb * 2u * 128u * 128u);
                    const uint64_t a_lo = make_desc_sw128(vs_u + (uint32_t)(vb * 2 + 1) * 128u * 128u);
                    const uint64_t u_hi = make_desc_sw128(us_u + (uint32_t)h * COUT * 128u);
                    const uint64_t u_lo = make_desc_sw128(us_u + (uint32_t)(8 + h) * COUT * 128u);
                    if (!(a.dbg & 1))
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const uint32_t dcol = tmem_base + (uint32_t)((2 * h + s2) * COUT);      // xi = 2 h + s2
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t o = (uint64_t)(2 * (s2 * 2 + kk));
                            tc_mma_tf32_elect(dcol, a_lo + o, u_hi + o, idesc, kk ? 1u : 0u);
                            tc_mma_tf32_elect(dcol, a_hi + o, u_lo + o, idesc, 1u);
                            tc_mma_tf32_elect(dcol, a_hi + o, u_hi + o, idesc, 1u);
                        }
                    }
                    tc_commit_elect(bar_empty + 8u * vb);
                }
            }
        }
    } else {
        // ===== transform / epilogue warps
        // transform role: thread = (one tile, four channels 4q..4q+3); a quarter-warp holds tiles (j, j+4), whose
        // swizzle phases differ in bit 2, so its eight 16-byte stores hit eight different bank groups
        const int q = tid & 3, e = (tid >> 2) & 7;
        const int m = warp * 8 + (e >> 1) + (e & 1) * 4;
        const int tpl = m >> 6, tt = m & 63, tty = tt >> 3, ttx = tt & 7;
        // epilogue role: warp = (TMEM lane quarter, 8-channel group)
        constexpr int NCG = COUT / 8;
        const int quarter = warp & 3, cg = warp >> 2;
        uint32_t pe[NVB] = {};

        // two patches: global NCHW -> the interior of zero-bordered planes (odd plane pitch: the transform reads them
        // with few bank conflicts).  The next pair is fetched with four 128-bit loads per thread that stay in flight
        // during this pair's transform and are scattered into the other staging buffer afterwards (4-byte cp.async
        // took 2-3 k clk of LSU time per pair: 8192 LDGSTS.32).  float4 id = k * 512 + tid: plane id >> 6.
        float4 nx[4];
        auto fetch = [&](int pair) {
            const float4* src = reinterpret_cast<const float4*>(a.x + (size_t)pair * (2 * WT_CIN * 256)) + tid;
            const bool second = (int64_t)pair * 2 + 1 < a.B;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                nx[k] = (k < 2 || second) ? __ldg(src + k * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto scatter = [&](int buf) {
            // float4 id -> plane (id >> 6), row (id >> 2) & 15, columns 4 (id & 3) .. + 3
            float* dst = raw + buf * WT_RAW + (tid >> 6) * WT_PLANE + (((tid >> 2) & 15) + 1) * WT_RP + (tid & 3) * 4 + 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float* p = dst + k * 8 * WT_PLANE;
                p[0] = nx[k].x; p[1] = nx[k].y; p[2] = nx[k].z; p[3] = nx[k].w;
            }
        };

        int pair = blockIdx.x;
        if (pair < npairs) { fetch(pair); scatter(0); }
        work_sync();                              // staging buffer 0 is complete
        for (int it = 0; pair < npairs; pair += gridDim.x, ++it) {
            const int buf = it & 1;
            const int nxt = pair + gridDim.x;
            if (nxt < npairs && !(a.dbg & 8)) fetch(nxt);
            float d[4][16];
#pragma unroll
            for (int cI = 0; cI < 4; ++cI) {
                const float* src = raw + buf * WT_RAW + (tpl * WT_CIN + 4 * q + cI) * WT_PLANE + (2 * tty) * WT_RP + 2 * ttx;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float v = src[rr * WT_RP + cc];
                        d[cI][rr * 4 + cc] = a.in_relu ? fmaxf(v, 0.f) : v;
                    }
            }

            // eight half-chunks: h = 2 * (row of V) + (xi pair inside the row); operand buffers alternate, so the MMAs
            // of half-chunk h run while half-chunk h + 1 is transformed
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const int chunk = h >> 1, pr = h & 1, vb = h % NVB;
                if (h >= NVB) {                   // the MMAs that read this buffer NVB half-chunks ago are done
                    mbar_wait(bar_empty + 8u * vb, pe[vb]); pe[vb] ^= 1u;
                }
                float* Vb = Vs + vb * (2 * 128 * 32);
                float hi[2][4], lo[2][4];
                if (!(a.dbg & 2)) {
#pragma unroll
                for (int cI = 0; cI < 4; ++cI) {
                    float u[4], v[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (chunk == 0) u[cc] = d[cI][0 + cc] - d[cI][8 + cc];
                        else if (chunk == 1) u[cc] = d[cI][4 + cc] + d[cI][8 + cc];
                        else if (chunk == 2) u[cc] = d[cI][8 + cc] - d[cI][4 + cc];
                        else u[cc] = d[cI][4 + cc] - d[cI][12 + cc];
                    }
                    v[0] = u[0] - u[2]; v[1] = u[1] + u[2]; v[2] = u[2] - u[1]; v[3] = u[1] - u[3];
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const float val = v[pr * 2 + s2];
                        hi[s2][cI] = __uint_as_float((__float_as_uint(val) + 0x1000u) & 0xffffe000u);
                        lo[s2][cI] = val - hi[s2][cI];
                    }
                }
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int off = m * 32 + (((s2 * 4 + q) ^ (m & 7)) << 2);
                    *reinterpret_cast<float4*>(Vb + off) = make_float4(hi[s2][0], hi[s2][1], hi[s2][2], hi[s2][3]);
                    *reinterpret_cast<float4*>(Vb + 128 * 32 + off) = make_float4(lo[s2][0], lo[s2][1], lo[s2][2], lo[s2][3]);
                }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8u * vb);
            }
#pragma unroll
            for (int b = 0; b < NVB; ++b) {           // the last commit on every buffer: M is complete
                mbar_wait(bar_empty + 8u * b, pe[b]); pe[b] ^= 1u;
            }
            tc_fence_after();

            // ---- epilogue: Y = A^T M A for tile (quarter * 32 + lane), channels cg * 8 .. + 7
            if (cg < NCG && !(a.dbg & 4)) {
                const int em = quarter * 32 + lane;
                const int pl = em >> 6, t = em & 63, ty = t >> 3, tx = t & 7;
                const int64_t p = (int64_t)pair * 2 + pl;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cg * 8);
                float r0[4][8], r1[4][8];      // A^T applied down the rows of M: r0 = m0 + m1 + m2, r1 = m1 - m2 - m3
#pragma unroll
                for (int jx = 0; jx < 4; ++jx) {
                    float q0[8], q1[8], q2[8], q3[8];
                    tmem_ld8(taddr + (uint32_t)((0 + jx) * COUT), q0);
                    tmem_ld8(taddr + (uint32_t)((4 + jx) * COUT), q1);
                    tmem_ld8(taddr + (uint32_t)((8 + jx) * COUT), q2);
                    tmem_ld8(taddr + (uint32_t)((12 + jx) * COUT), q3);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        r0[jx][c] = (q0[c] + q1[c]) + q2[c];
                        r1[jx][c] = (q1[c] - q2[c]) - q3[c];
                    }
                }
                if constexpr (FUSE) {
                    // the 32 mid channels of a pixel are spread over four warps: park them in the (idle) operand buffers
                    // as [pixel][32] rows, 16-byte chunk index XOR-ed with bits 1-3 of the pixel so that the eight tiles
                    // of a quarter-warp (two pixels apart) store to eight different bank groups
                    float yv[4][8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float b = __ldg(a.bias + cg * 8 + c);
                        yv[0][c] = ((r0[0][c] + r0[1][c]) + r0[2][c]) + b; yv[1][c] = ((r0[1][c] - r0[2][c]) - r0[3][c]) + b;
                        yv[2][c] = ((r1[0][c] + r1[1][c]) + r1[2][c]) + b; yv[3][c] = ((r1[1][c] - r1[2][c]) - r1[3][c]) + b;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int gpx = pl * 256 + (2 * ty + (i >> 1)) * 16 + 2 * tx + (i & 1);
                        float* row = Vs + gpx * 32;
                        const int sw = (gpx >> 1) & 7;
#pragma unroll
                        for (int hq = 0; hq < 2; ++hq) {
                            float4 v = make_float4(yv[i][4 * hq], yv[i][4 * hq + 1], yv[i][4 * hq + 2], yv[i][4 * hq + 3]);
                            if (a.out_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            *reinterpret_cast<float4*>(row + (((2 * cg + hq) ^ sw) << 2)) = v;
                        }
                    }
                } else if (p < a.B) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int co = cg * 8 + c;
                        const float b = __ldg(a.bias + co);
                        float y00 = ((r0[0][c] + r0[1][c]) + r0[2][c]) + b, y01 = ((r0[1][c] - r0[2][c]) - r0[3][c]) + b;
                        float y10 = ((r1[0][c] + r1[1][c]) + r1[2][c]) + b, y11 = ((r1[1][c] - r1[2][c]) - r1[3][c]) + b;
                        if (a.out_relu) { y00 = fmaxf(y00, 0.f); y01 = fmaxf(y01, 0.f); y10 = fmaxf(y10, 0.f); y11 = fmaxf(y11, 0.f); }
                        float* dst = a.y + (((size_t)p * COUT + co) * 16 + 2 * ty) * 16 + 2 * tx;
                        *reinterpret_cast<float2*>(dst) = make_float2(y00, y01);
                        *reinterpret_cast<float2*>(dst + 16) = make_float2(y10, y11);
                    }
                }
            }
            if constexpr (FUSE) {
                // ---- fused tail: out = x + conv1x1(mid) + bias2; thread = one pixel, 16 output channels; weights
                // from the constant bank (same for every thread: uniform operands), x from the staging planes
                work_sync();
                const int gpx = tid, pl = gpx >> 8, yy = (gpx >> 4) & 15, xx = gpx & 15;
                const int64_t p = (int64_t)pair * 2 + pl;
                const float* row = Vs + gpx * 32;
                const int sw = (gpx >> 1) & 7;
                float o[16];
#pragma unroll
                for (int co = 0; co < 16; ++co) o[co] = c_w2[512 + co];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const float4 v = *reinterpret_cast<const float4*>(row + ((ch ^ sw) << 2));
                    const float mv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int co = 0; co < 16; ++co) o[co] = fmaf(mv[u], c_w2[(ch * 4 + u) * 16 + co], o[co]);
                }
                if (p < a.B) {
                    const float* xs = raw + buf * WT_RAW + (pl * WT_CIN) * WT_PLANE + (yy + 1) * WT_RP + xx + 1;
                    float* dst = a.y2 + (size_t)p * 16 * 256 + yy * 16 + xx;
#pragma unroll
                    for (int co = 0; co < 16; ++co) dst[co * 256] = o[co] + xs[co * WT_PLANE];
                }
            }
            if (nxt < npairs) scatter(buf ^ 1);   // (its previous contents were consumed one pair ago)
            tc_fence_before();
            work_sync();                      // tensor memory is free and the other staging buffer complete
            tc_fence_after();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)C::TMEM_COLS);
    }
}

// U = G g G^T per (co, ci), G = [1 0 0; .5 .5 .5; .5 -.5 .5; 0 0 1], from the packed [ci][ky][kx][co] weights; split
// hi = tf32(u), lo = tf32(u - hi); stored as [hi|lo][xi pair][co][32 = (xi & 1) * 16 + ci] with the 128-byte swizzle
__global__ void __launch_bounds__(256) pack_wino_tc_kernel(const float* __restrict__ w, float* __restrict__ out, int cout) {
    pdl_wait();
    const double G[4][3] = {{1.0, 0.0, 0.0}, {0.5, 0.5, 0.5}, {0.5, -0.5, 0.5}, {0.0, 0.0, 1.0}};
    const int total = 16 * cout * WT_CIN;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
        const int c = id % WT_CIN;
        const int co = (id / WT_CIN) % cout;
        const int xi = id / (WT_CIN * cout);
        const int i = xi >> 2, j = xi & 3;
        double s = 0.0;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
                s += G[i][ky] * (double)w[((size_t)(c * 3 + ky) * 3 + kx) * cout + co] * G[j][kx];
        const float v = (float)s;
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        uint32_t lb;
        asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lb) : "f"(v - hi));
        const int col = (xi & 1) * 16 + c;
        const int off = (xi >> 1) * cout * 32 + co * 32 + ((((col >> 2) ^ (co & 7)) << 2) | (col & 3));
        out[off] = hi;
        out[8 * cout * 32 + off] = __uint_as_float(lb);
    }
}

struct WinoPool {           // the constant bank of the fused tail is one region per device (see conv_tma.cuh: PoolState)
    std::mutex mu;
    cudaEvent_t ev = nullptr;
    cudaStream_t last = nullptr;
    bool used = false;
};
WinoPool g_wpool[64];

template <int COUT, bool FUSE, int NVB>
int launch_wino_tc(const ConvWinoArgs& a, cudaStream_t st) {
    using C = WtCfg<COUT, NVB>;
    auto kern = conv_wino_tc_kernel<COUT, FUSE, NVB>;
    int dev = 0;
    DMB_CUDA(cudaGetDevice(&dev));
    DMB_CHECK(dev >= 0 && dev < 64, "conv_wino_tc: device index %d out of range", dev);
    static bool configured[64] = {false};
    if (!configured[dev]) {
        DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[dev] = true;
    }
    int sms = 148;
    DMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int npairs = (a.B + 1) / 2;
    const int grid = std::min(npairs, sms);
    WinoTcArgs k{};
    k.x = a.x; k.u = a.u; k.bias = a.bias; k.y = a.y; k.y2 = a.y2; k.B = a.B; k.in_relu = a.in_relu; k.out_relu = a.out_relu;
    { const char* e = getenv("DMB_WINO_DBG"); k.dbg = e ? atoi(e) : 0; }
    if constexpr (FUSE) {
        WinoPool& ps = g_wpool[dev];
        std::lock_guard<std::mutex> lock(ps.mu);
        if (!ps.ev) DMB_CUDA(cudaEventCreateWithFlags(&ps.ev, cudaEventDisableTiming));
        if (ps.used && ps.last != st) DMB_CUDA(cudaStreamWaitEvent(st, ps.ev, 0));
        DMB_CUDA(cudaMemcpyToSymbolAsync(c_w2, a.w2, 512 * 4, 0, cudaMemcpyDeviceToDevice, st));
        DMB_CUDA(cudaMemcpyToSymbolAsync(c_w2, a.bias2, 16 * 4, 512 * 4, cudaMemcpyDeviceToDevice, st));
        DMB_LAUNCH((kern), grid, WT_THREADS + 32, C::SMEM, st, k);
        DMB_CUDA(cudaGetLastError());
        DMB_CUDA(cudaEventRecord(ps.ev, st));
        ps.last = st; ps.used = true;
    } else {
        DMB_LAUNCH((kern), grid, WT_THREADS + 32, C::SMEM, st, k);
        DMB_CUDA(cudaGetLastError());
    }
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace

bool conv_wino_supported(int cin, int cout, int ks, int stride, int H, int W) {
    return ks == 3 && stride == 1 && cin == WT_CIN && (cout == 32 || cout == 16) && H == 16 && W == 16;
}

int64_t conv_wino_weight_floats(int cin, int cout) { return 2ll * 16 * cin * cout; }

int pack_wino_weights(const float* w_packed, float* out, int cin, int cout, cudaStream_t st) {
    DMB_CHECK(cin == WT_CIN, "pack_wino_weights: Cin must be %d", WT_CIN);
    const int total = 16 * cout * cin;
    DMB_LAUNCH((pack_wino_tc_kernel), (total + 255) / 256, 256, 0, st, w_packed, out, cout);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int conv_wino(const ConvWinoArgs& a, cudaStream_t st) {
    DMB_CHECK(a.B > 0, "conv_wino: empty batch");
    DMB_CHECK(!(reinterpret_cast<uintptr_t>(a.u) & 15) && !(reinterpret_cast<uintptr_t>(a.y) & 7),
              "conv_wino: u must be 16-byte and y 8-byte aligned");
    if (a.w2) {
        DMB_CHECK(a.Cout == 32 && a.bias2 && a.y2, "conv_wino: the fused 1x1 tail needs 32 mid channels, bias2 and y2");
        return launch_wino_tc<32, true, 2>(a, st);
    }
    if (a.Cout == 32) return launch_wino_tc<32, false, 2>(a, st);
    if (a.Cout == 16) {
        // a third operand buffer fits next to the smaller U tiles here, but measured slower (0.198 vs 0.181 ms per 8192
        // patches): kept behind DMB_WINO_NVB=3 as an experiment
        const char* e = getenv("DMB_WINO_NVB");
        return (e && e[0] == '3') ? launch_wino_tc<16, false, 3>(a, st) : launch_wino_tc<16, false, 2>(a, st);
    }
    DMB_CHECK(false, "conv_wino: Cout %d not in {16, 32}", a.Cout);
}

}  // namespace dmb
