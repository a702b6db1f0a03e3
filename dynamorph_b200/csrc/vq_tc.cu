// VectorQuantizer forward with the code search on the tensor cores (HiddenStateExtractor/vq_vae.py:52-84, :90-116).
//
// The reference's argmin runs over the direct-form distances  sum_d (z_d - e_kd)^2  in torch's summation order, and the
// indices have to come out bit-identical.  Evaluating that form for every (position, code) pair is what makes the
// CUDA-core kernel (vq.cu) FADD/FMUL-bound: 3 non-fusable operations per (position, code, channel).  Here the search
// is split in two:
//   1. scores  S[p,k] = z_p . e_k  for a tile of 128 positions against the whole codebook as ONE single-pass TF32
//      tcgen05 GEMM (A = z tile, B = codebook, both K-major 128-byte-swizzled in shared memory, accumulator
//      128 lanes x K columns in tensor memory); the approximate distance is  a_k = |e_k|^2 - 2 S[p,k];
//   2. every code whose approximate distance lies within a rigorous error bound of the smallest one is a candidate
//      (typically one or two of K), and only the candidates get the exact reference-order distance on the CUDA
//      cores (same cascade sum as vq.cu); first index wins ties, exactly as torch.argmax(-distances).
// Error bound: TF32 keeps 10 explicit mantissa bits (truncated by the tensor core), so |S - z.e| <= 2^-9 (1+eps)
// sum_d |z_d e_kd| + accumulation error <= 0.002 |z| |e_k|; a candidate test against the minimum doubles that twice:
// margin = 0.0085 |z| max_k|e_k| + 2e-5 (|z| + max_k|e_k|)^2 (the second term covers the fp32 round-off of |e_k|^2 and
// of the reference's own distance evaluation).  A position with no candidate (NaN input) or more than CAND_MAX
// candidates (degenerate codebook) falls back to the exhaustive exact search.
// The kernel is persistent (one CTA keeps the codebook in shared memory and walks tiles); one thread = one latent
// position = one TMEM lane; thread 0 issues the MMAs; gather, straight-through output, loss partial and histogram are
// fused exactly as in vq.cu.
#include "common.cuh"

namespace dmb {
namespace {

constexpr int VT = 128;          // threads == positions per tile == TMEM lanes
constexpr int CAND_MAX = 8;

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
// converged-warp forms (one elected lane executes; see conv_tc.cu)
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
// K-major operand, 128-byte swizzle (see conv_tc.cu)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// float4 chunk `q` (channels 4q..4q+3) of row `r` in a [rows][DP] operand stored as DP/32 tiles of [rows][32] with the
// 16-byte chunk index XOR-ed by (row & 7)
__device__ __forceinline__ int sw_off(int r, int q, int rows) {
    return (q >> 3) * rows * 32 + r * 32 + (((q & 7) ^ (r & 7)) << 2);
}

// torch's multi_row_sum order (see vq.cu: dist_cascade), codebook row read from the swizzled operand tile
template <int D>
__device__ __forceinline__ float dist_cascade_sw(const float (&z)[D], const float* __restrict__ cb, int k, int rows) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int i = 0;
#pragma unroll
    for (int c0 = 0; c0 + 16 <= D; c0 += 16) {
#pragma unroll
        for (int c = c0; c < c0 + 16; c += 4) {
            const float4 ev = *reinterpret_cast<const float4*>(cb + sw_off(k, c >> 2, rows));
            float d;
            d = __fsub_rn(z[c], ev.x);     acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
            d = __fsub_rn(z[c + 1], ev.y); acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
            d = __fsub_rn(z[c + 2], ev.z); acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
            d = __fsub_rn(z[c + 3], ev.w); acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
        }
        i = c0 + 16;
        acc1 = __fadd_rn(acc1, acc0); acc0 = 0.f;
        if ((i & (15 << 4)) == 0) {
            acc2 = __fadd_rn(acc2, acc1); acc1 = 0.f;
            if ((i & (15 << 8)) == 0) { acc3 = __fadd_rn(acc3, acc2); acc2 = 0.f; }
        }
    }
    acc0 = __fadd_rn(acc0, acc1);
    acc0 = __fadd_rn(acc0, acc2);
    acc0 = __fadd_rn(acc0, acc3);
    return acc0;
}

struct VqTcGeom {
    int Kp;          // codes padded to a multiple of 16 (operand rows)
    int tmem_cols;   // power of two >= max(32, Kp rounded up to 32)
    int64_t ntiles;
    int dbg;         // DMB_VQ_TC_DBG=1: write the candidate count instead of the index (scripts/dbg_vq_tc.py)
};

template <int D>
__global__ void __launch_bounds__(VT, D == 16 ? 8 : (D == 32 ? 4 : 1)) vq_tc_kernel(const VqArgs a, const VqTcGeom g) {
    static_assert(D == 16 || D == 32 || D == 64, "embedding_dim must be 16, 32 or 64");
    constexpr int DP = D < 32 ? 32 : D;      // channels padded to whole 128-byte operand rows
    constexpr int NH = DP / 32;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const int Kp = g.Kp;
    // carve: codebook tiles [NH][Kp][32] | z tiles [NH][128][32] | n2[Kp] | cand[128][CAND_MAX] u16 | hist[K] | misc
    float* cbs = reinterpret_cast<float*>(bp);
    const uint32_t cbs_u = base;
    const size_t cb_bytes = (size_t)NH * Kp * 128;
    const size_t cb_pad = (cb_bytes + 1023) & ~(size_t)1023;
    float* zs = reinterpret_cast<float*>(bp + cb_pad);
    const uint32_t zs_u = base + (uint32_t)cb_pad;
    float* n2 = reinterpret_cast<float*>(bp + cb_pad + (size_t)NH * VT * 128);
    const int K32 = (a.K + 31) & ~31;        // scan granularity (>= Kp - 16)
    unsigned short* cand = reinterpret_cast<unsigned short*>(n2 + (K32 > Kp ? K32 : Kp));
    int* hist = reinterpret_cast<int*>(cand + VT * CAND_MAX);
    double* red = reinterpret_cast<double*>(hist + ((a.K + 1) & ~1));
    float* redf = reinterpret_cast<float*>(red + 4);
    uint64_t* bar_mem = reinterpret_cast<uint64_t*>(redf + 4);
    uint32_t* slot_mem = reinterpret_cast<uint32_t*>(bar_mem + 1);
    const uint32_t bar = smem_u32(bar_mem), slot = smem_u32(slot_mem);

    // (the shuffle makes the warp index provably warp-uniform for ptxas: `if (warp == 0)` is then a uniform branch)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) { mbar_init(bar, 1u); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, (uint32_t)g.tmem_cols);
    pdl_wait();

    // ---- codebook -> swizzled operand tiles (zero padding for channels >= D and codes >= K), |e_k|^2, max |e_k|
#pragma unroll 8
    for (int i = tid; i < Kp * (DP / 4); i += VT) {
        const int k = i / (DP / 4), q = i - k * (DP / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < a.K && q * 4 < D) v = __ldg(reinterpret_cast<const float4*>(a.codebook + (size_t)k * D) + q);
        *reinterpret_cast<float4*>(cbs + sw_off(k, q, Kp)) = v;
    }
    for (int i = tid; i < a.K; i += VT) hist[i] = 0;
    __syncthreads();
    float emax2 = 0.f;
    for (int k = tid; k < (K32 > Kp ? K32 : Kp); k += VT) {
        float s = 3.0e38f;
        if (k < a.K) {
            s = 0.f;
#pragma unroll
            for (int q = 0; q < D / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(cbs + sw_off(k, q, Kp));
                s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
            }
        }
        n2[k] = s;
        if (k < a.K) emax2 = fmaxf(emax2, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax2 = fmaxf(emax2, __shfl_xor_sync(0xffffffffu, emax2, o));
    if ((tid & 31) == 0) redf[warp] = emax2;
    fence_proxy_async();                     // codebook tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot_mem);
    const float emax = sqrtf(fmaxf(fmaxf(redf[0], redf[1]), fmaxf(redf[2], redf[3])));

    const int64_t total = a.B * a.P;
    double lsum_cta = 0.0;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const int64_t n = tile * VT + tid;
        const bool live = n < total;
        const int64_t b = live ? n / a.P : 0;
        const int pos = live ? (int)(n - b * a.P) : 0;
        const size_t zbase = (size_t)b * D * a.P + pos;
        float z[D];
        if (live) {
            if (a.pre_b) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const size_t t = (a.pre_per_sample ? (size_t)b * D : 0) + c;
                    float va = __ldg(a.pre_a + zbase + (size_t)c * a.P);
                    float vb = __ldg(a.pre_b + zbase + (size_t)c * a.P);
                    if (a.pre_sa) va = fmaf(va, a.pre_sa[t], a.pre_ta[t]);
                    if (a.pre_sb) vb = fmaf(vb, a.pre_sb[t], a.pre_tb[t]);
                    z[c] = va + vb;
                    if (a.z_before_out) a.z_before_out[zbase + (size_t)c * a.P] = z[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < D; ++c) z[c] = __ldg(a.z + zbase + (size_t)c * a.P);
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) z[c] = 0.f;
        }
        float zz = 0.f;
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q * 4 < D) v = make_float4(z[(q * 4) % D], z[(q * 4 + 1) % D], z[(q * 4 + 2) % D], z[(q * 4 + 3) % D]);
            *reinterpret_cast<float4*>(zs + sw_off(tid, q, VT)) = v;
            zz = fmaf(v.x, v.x, zz); zz = fmaf(v.y, v.y, zz); zz = fmaf(v.z, v.z, zz); zz = fmaf(v.w, v.w, zz);
        }
        const float znorm = sqrtf(zz);
        const float margin = 0.0085f * znorm * emax + 2e-5f * (znorm + emax) * (znorm + emax);
        fence_proxy_async();
        tc_fence_before();                   // last tile's tcgen05.ld are complete before the accumulator is overwritten
        __syncthreads();
        if (warp == 0) {                     // converged warp; one elected lane per tcgen05 instruction
            tc_fence_after();
            for (int k0 = 0; k0 < Kp; k0 += 256) {
                const int nn = min(256, Kp - k0);
                const uint32_t idesc = make_idesc_tf32(128, nn);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const uint64_t ad = make_desc_sw128(zs_u + (uint32_t)h * VT * 128u);
                    const uint64_t bd = make_desc_sw128(cbs_u + (uint32_t)(h * Kp + k0) * 128u);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_tf32_elect(tmem_base + (uint32_t)k0, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc,
                                          (h | k) != 0 ? 1u : 0u);
                }
            }
            tc_commit_elect(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();

        // ---- pass 1: smallest approximate distance; pass 2: candidates within the error margin
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        // (codes K..K32-1 carry |e|^2 = +huge in n2, so the tail of the last 32-column chunk never wins; columns the
        // MMA did not write may hold anything: NaN loses both fminf and the <= test)
        float m0 = 3.0e38f, m1 = 3.0e38f, m2 = 3.0e38f, m3 = 3.0e38f;
        for (int c0 = 0; c0 < a.K; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 nv = *reinterpret_cast<const float4*>(n2 + c0 + j);
                m0 = fminf(m0, fmaf(-2.f, __uint_as_float(v[j]), nv.x));
                m1 = fminf(m1, fmaf(-2.f, __uint_as_float(v[j + 1]), nv.y));
                m2 = fminf(m2, fmaf(-2.f, __uint_as_float(v[j + 2]), nv.z));
                m3 = fminf(m3, fmaf(-2.f, __uint_as_float(v[j + 3]), nv.w));
            }
        }
        const float amin = fminf(fminf(m0, m1), fminf(m2, m3));
        const float thr = amin + margin;
        int ncand = 0;
        for (int c0 = 0; c0 < a.K; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_base + (uint32_t)c0, v);
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 nv = *reinterpret_cast<const float4*>(n2 + c0 + j);
                mask |= (fmaf(-2.f, __uint_as_float(v[j]), nv.x) <= thr ? 1u : 0u) << j;
                mask |= (fmaf(-2.f, __uint_as_float(v[j + 1]), nv.y) <= thr ? 1u : 0u) << (j + 1);
                mask |= (fmaf(-2.f, __uint_as_float(v[j + 2]), nv.z) <= thr ? 1u : 0u) << (j + 2);
                mask |= (fmaf(-2.f, __uint_as_float(v[j + 3]), nv.w) <= thr ? 1u : 0u) << (j + 3);
            }
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                if (ncand < CAND_MAX) cand[tid * CAND_MAX + ncand] = (unsigned short)(c0 + j);
                ++ncand;
            }
        }
        // ---- exact reference-order distance of the candidates (increasing index: first index wins ties)
        float best = 0.f;
        int bi = -1;
        if (ncand == 0 || ncand > CAND_MAX) {
            for (int k = 0; k < a.K; ++k) {
                const float d = dist_cascade_sw<D>(z, cbs, k, Kp);
                if (bi < 0 || d < best) { best = d; bi = k; }
            }
        } else {
            for (int r = 0; r < ncand; ++r) {
                const int k = cand[tid * CAND_MAX + r];
                const float d = dist_cascade_sw<D>(z, cbs, k, Kp);
                if (bi < 0 || d < best) { best = d; bi = k; }
            }
        }
        // ---- gather + straight-through value + loss partial
        if (live) {
            double lsum = 0.0;
#pragma unroll
            for (int q = 0; q < D / 4; ++q) {
                const float4 e = *reinterpret_cast<const float4*>(cbs + sw_off(bi, q, Kp));
                const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = q * 4 + u;
                    const float diff = __fsub_rn(ev[u], z[c]);
                    if (a.z_st) a.z_st[zbase + (size_t)c * a.P] = __fadd_rn(z[c], diff);
                    lsum += (double)diff * (double)diff;
                }
            }
            if (a.idx) a.idx[n] = g.dbg == 1 ? ncand : bi;
            if (a.stats) { atomicAdd(&hist[bi], 1); lsum_cta += lsum; }
        }
    }

    if (a.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum_cta += __shfl_xor_sync(0xffffffffu, lsum_cta, o);
        if ((tid & 31) == 0) red[warp] = lsum_cta;
        __syncthreads();
        if (tid == 0) {
            atomicAdd(a.stats + 0, red[0] + red[1] + red[2] + red[3]);
            int64_t mine = 0;
            for (int64_t tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
                const int64_t rem = total - tile * VT;
                mine += rem < VT ? rem : VT;
            }
            atomicAdd(a.stats + 1, (double)mine);
        }
        for (int i = tid; i < a.K; i += VT)
            if (hist[i]) atomicAdd(a.stats + 2 + i, (double)hist[i]);   // integer-valued: order-free
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
    }
}

template <int D>
int launch_vq_tc(const VqArgs& a, cudaStream_t st) {
    constexpr int DP = D < 32 ? 32 : D;
    constexpr int NH = DP / 32;
    VqTcGeom g{};
    g.Kp = (a.K + 15) & ~15;
    const int c32 = (g.Kp + 31) & ~31;
    g.tmem_cols = c32 <= 32 ? 32 : c32 <= 64 ? 64 : c32 <= 128 ? 128 : c32 <= 256 ? 256 : 512;
    const int64_t total = a.B * a.P;
    g.ntiles = (total + VT - 1) / VT;
    { const char* e = getenv("DMB_VQ_TC_DBG"); g.dbg = e ? atoi(e) : 0; }
    size_t smem = ((size_t)NH * g.Kp * 128 + 1023) & ~(size_t)1023;
    smem += (size_t)NH * VT * 128;                       // z tiles
    smem += (size_t)(g.Kp + 32) * 4 + VT * CAND_MAX * 2 + (size_t)((a.K + 1) & ~1) * 4 + 4 * 8 + 4 * 4 + 16;
    smem += 1024 + 64;                                   // alignment slack
    auto kern = vq_tc_kernel<D>;
    int dev = 0;
    DMB_CUDA(cudaGetDevice(&dev));
    static size_t configured[64] = {0};
    DMB_CHECK(dev >= 0 && dev < 64, "vq_tc: device index %d out of range", dev);
    if (smem > 227 * 1024) return 1;
    if (smem > configured[dev]) {
        DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    // resident CTAs per SM by hand (the occupancy API answers 1 for a kernel that allocates tensor memory): shared
    // memory (+1 KB the driver reserves per CTA), registers, and one accumulator of tmem_cols columns per CTA
    cudaFuncAttributes fa;
    DMB_CUDA(cudaFuncGetAttributes(&fa, kern));
    int per_sm = (int)((227 * 1024) / (smem + fa.sharedSizeBytes + 1024));
    per_sm = std::min(per_sm, 65536 / std::max(1, fa.numRegs * VT));
    per_sm = std::min(per_sm, 512 / g.tmem_cols);
    per_sm = std::min(per_sm, 8);
    if (per_sm < 1) return 1;
    int sms = 148;
    DMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t grid = std::min<int64_t>(g.ntiles, (int64_t)sms * per_sm);
    if (g.dbg == 3) fprintf(stderr, "vq_tc<%d>: K=%d smem=%zu per_sm=%d tmem_cols=%d grid=%lld tiles=%lld\n", D, a.K, smem,
                            per_sm, g.tmem_cols, (long long)grid, (long long)g.ntiles);
    DMB_LAUNCH((kern), (unsigned)grid, VT, smem, st, a, g);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace

// returns 1 when the shape is not served (the caller runs the CUDA-core kernel)
int vq_forward_tc(const VqArgs& a, cudaStream_t st) {
    if (a.K < 16 || a.K > 512) return 1;
    if ((reinterpret_cast<uintptr_t>(a.codebook) & 15) != 0) return 1;
    switch (a.D) {
        case 16: return launch_vq_tc<16>(a, st);
        case 32: return launch_vq_tc<32>(a, st);
        case 64: return launch_vq_tc<64>(a, st);
        default: return 1;
    }
}

}  // namespace dmb
