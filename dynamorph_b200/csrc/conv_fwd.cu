// Direct convolution forward on CUDA cores (fp32, exact), NCHW planar.
//
// Covers every Conv2d on the path: 1x1, 3x3 (pad 1) and 4x4 stride 2 (pad 1)
// (reference layers: HiddenStateExtractor/vq_vae.py:203-209, :276-289, :298).
//
// Why CUDA cores and not tcgen05: the reduction is C*kh*kw = 32..256 with 8..32 output
// channels and the parity bar is fp32 1e-4 / bit-exact indices (BASELINE.json north_star);
// see DESIGN.md "Roofline".  The kernel is FP32-FMA bound, so the design goal is FFMA
// issue share: each thread owns PW=8 consecutive output pixels x CO_T output channels
// (64 accumulators), activations come from a shared-memory tile with conflict-free 128-bit
// loads, weights from shared memory as warp-uniform 128-bit broadcasts.
//
// Shared-memory tile, per (local patch, input channel, input row): the RAW input row, shifted right by
// one 16-byte granule (4 floats) so that the left zero-padding column sits in a granule of its own and
// every data granule keeps its 16-byte alignment.  Rows are fetched global->shared with cp.async
// (16 B each, zero-fill for rows outside the image): no register staging, all copies of a chunk in
// flight at once.  Inside a row, granules are dealt round-robin to NGS sub-planes (NGS = granules a
// thread advances per strip: 4 for stride 2, 2 for stride 1) so the i-th 128-bit load of neighbouring
// threads hits neighbouring granules -> no bank conflicts although each thread walks 16-24 contiguous
// floats.  The producer's BatchNorm affine + ReLU (BATCH / PER_SAMPLE modes) is applied by an in-place
// pass over the landed tile; zero padding is untouched by it, as torch pads after the activation.
#include "common.cuh"

namespace dmb {

namespace {

constexpr int PW = 8;        // output pixels per thread along x
constexpr int NG = PW / 4;   // granules per strip

struct ConvK {
    ConvFwdArgs a;
    int TR, NP, CIC, nbands, SPR, RIN, SUBW;
    int row_stride, ci_stride, patch_stride, w_floats, tile_floats, stage_floats;
    int threads;
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// 16-byte global->shared async copy (LDGSTS); src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int KS, int STRIDE, int CO_T>
__global__ void __launch_bounds__(128, 4) conv_fwd_kernel(const ConvK k) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    extern __shared__ __align__(16) float smem[];
    const ConvFwdArgs& a = k.a;
    constexpr int PAD = (KS == 1) ? 0 : 1;
    constexpr int PADL = PAD ? 4 : 0;                       // floats of left shift inside a tile row
    constexpr int NGS = (STRIDE == 2) ? 4 : 2;               // sub-planes == granules per strip
    constexpr int NV = (KS == 1) ? 2 : ((STRIDE == 2) ? 6 : 4);   // float4 loads per tile row per thread

    const int tid = threadIdx.x;
    const int band = blockIdx.x % k.nbands;
    const int64_t b0 = (int64_t)(blockIdx.x / k.nbands) * k.NP;

    // thread -> (channel group, local patch, row, strip)
    const int strips_per_patch = k.TR * k.SPR;
    const int strips = k.NP * strips_per_patch;
    const int cg = tid / strips;
    const int srem = tid - cg * strips;
    const int pl = srem / strips_per_patch;
    const int prem = srem - pl * strips_per_patch;
    const int row = prem / k.SPR;
    const int sx = prem - row * k.SPR;
    const int oy = band * k.TR + row;
    const int64_t b = b0 + pl;
    const bool live = b < a.B;

    float acc[CO_T][PW];
#pragma unroll
    for (int c = 0; c < CO_T; ++c)
#pragma unroll
        for (int p = 0; p < PW; ++p) acc[c][p] = 0.f;

    const int in_row0 = band * k.TR * STRIDE - PAD;   // input row held in tile row 0
    const int W4 = a.W >> 2;
    // loader role of this thread: column slot ld_q (one 16-byte granule of the input row), first tile
    // row ld_r0, row step ld_rstep; the granule's place inside a tile row is loop-invariant.
    const int ld_q = tid % W4;
    const int ld_r0 = tid / W4;
    const int ld_rstep = (int)blockDim.x / W4;
    auto gslot = [&](int g) { return (g % NGS) * k.SUBW + (g / NGS) * 4; };
    const int so_dst = gslot(ld_q + PADL / 4);
    const int so_left = gslot(0), so_right = gslot(W4 + 1);
    const bool transform = (a.in_scale != nullptr) || a.in_relu || (a.x2 != nullptr);

    // Two-stage software pipeline over input-channel chunks: the cp.async copies of chunk c+1 are in flight
    // while chunk c is being multiplied (shared memory holds two {weights, tile} stages).
    auto issue_loads = [&](int c0, int stage) {
        float* wsS = smem + stage * k.stage_floats;
        float* tileS = wsS + ((k.w_floats + 3) & ~3);
            // ---- weights chunk: contiguous [CIC][KS][KS][Cout]
            if constexpr (CO_T % 4 == 0) {
                const float* src = a.w + (size_t)c0 * KS * KS * a.Cout;
                for (int i = tid; i < (k.w_floats >> 2); i += blockDim.x) cp_async16(wsS + 4 * i, src + 4 * i, 16);
            } else {
                const float* src = a.w + (size_t)c0 * KS * KS * a.Cout;
                for (int i = tid; i < k.w_floats; i += blockDim.x) wsS[i] = __ldg(src + i);
            }
            // ---- activation tile: one cp.async per (row, granule).  A thread keeps its granule column and walks
            // the rows of each (patch, channel) plane with strength-reduced pointers.
            if (ld_r0 < ld_rstep) {
                for (int lp = 0; lp < k.NP; ++lp) {
                    const int64_t bb = b0 + lp;
                    const bool pvalid = bb < a.B;
                    for (int cil = 0; cil < k.CIC; ++cil) {
                        const float* plane = a.x + ((size_t)(pvalid ? bb : 0) * a.Cin + (c0 + cil)) * a.H * a.W + 4 * ld_q;
                        float* dplane = tileS + lp * k.patch_stride + cil * k.ci_stride;
                        for (int r = ld_r0; r < k.RIN; r += ld_rstep) {
                            const int iy = in_row0 + r;
                            const bool valid = pvalid && iy >= 0 && iy < a.H;
                            float* rowp = dplane + r * k.row_stride;
                            cp_async16(rowp + so_dst, valid ? plane + (size_t)iy * a.W : a.x, valid ? 16 : 0);
                            if constexpr (PAD == 1) {
                                if (ld_q == 0) {
                                    *reinterpret_cast<float4*>(rowp + so_left) = make_float4(0.f, 0.f, 0.f, 0.f);
                                    *reinterpret_cast<float4*>(rowp + so_right) = make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
                        }
                    }
                }
            }
        cp_async_commit();
    };

    const int nchunks = a.Cin / k.CIC;
    issue_loads(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * k.CIC;
        float* ws = smem + (ch & 1) * k.stage_floats;
        float* tile = ws + ((k.w_floats + 3) & ~3);
        if (ch + 1 < nchunks) {
            issue_loads(c0 + k.CIC, (ch + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (transform) {
            // in-place BN affine + ReLU over the granules this thread copied (valid rows only)
            if (ld_r0 < ld_rstep) {
                for (int lp = 0; lp < k.NP; ++lp) {
                    const int64_t bb = b0 + lp;
                    if (bb >= a.B) break;
                    for (int cil = 0; cil < k.CIC; ++cil) {
                        const size_t ai = (a.in_per_sample ? (size_t)bb * a.Cin : 0) + c0 + cil;
                        float sc = 1.f, bc = 0.f, sh = 0.f;
                        if (a.in_scale) { sc = __ldg(a.in_scale + ai); sh = __ldg(a.in_shift + ai); }
                        if (a.x2) bc = __ldg(a.in_b + ai);
                        const float* plane2 = a.x2 ? a.x2 + ((size_t)bb * a.Cin + (c0 + cil)) * a.H * a.W + 4 * ld_q : nullptr;
                        float* dplane = tile + lp * k.patch_stride + cil * k.ci_stride + so_dst;
                        for (int r = ld_r0; r < k.RIN; r += ld_rstep) {
                            const int iy = in_row0 + r;
                            if (iy < 0 || iy >= a.H) continue;
                            float4* gp = reinterpret_cast<float4*>(dplane + r * k.row_stride);
                            float4 v = *gp;
                            if (plane2) {
                                const float4 u = __ldg(reinterpret_cast<const float4*>(plane2 + (size_t)iy * a.W));
                                v.x = fmaf(v.x, sc, fmaf(u.x, bc, sh)); v.y = fmaf(v.y, sc, fmaf(u.y, bc, sh));
                                v.z = fmaf(v.z, sc, fmaf(u.z, bc, sh)); v.w = fmaf(v.w, sc, fmaf(u.w, bc, sh));
                            } else if (a.in_scale) {
                                v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh);
                                v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
                            }
                            if (a.in_relu) {
                                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f);
                                v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                            }
                            *gp = v;
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---- FMA core
        const float* tp = tile + pl * k.patch_stride + (row * STRIDE) * k.row_stride;
        const float* wp = ws + cg * CO_T;
        for (int cil = 0; cil < k.CIC; ++cil) {
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const float* rp = tp + cil * k.ci_stride + ky * k.row_stride;
                float av[NV * 4];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const float4 t4 = lds4(rp + (i % NGS) * k.SUBW + (sx + i / NGS) * 4);
                    av[4 * i + 0] = t4.x; av[4 * i + 1] = t4.y;
                    av[4 * i + 2] = t4.z; av[4 * i + 3] = t4.w;
                }
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    float wv[CO_T];
                    const float* wrow = wp + ((cil * KS + ky) * KS + kx) * a.Cout;
                    if constexpr (CO_T % 4 == 0) {
#pragma unroll
                        for (int c = 0; c < CO_T; c += 4) {
                            const float4 t4 = lds4(wrow + c);
                            wv[c] = t4.x; wv[c + 1] = t4.y; wv[c + 2] = t4.z; wv[c + 3] = t4.w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < CO_T; ++c) wv[c] = wrow[c];
                    }
                    // window index of tap kx for output pixel p (window starts at tile position STRIDE*ox0;
                    // input column c lives at position c + PADL, and c = STRIDE*ox + kx - PAD)
#pragma unroll
                    for (int c = 0; c < CO_T; ++c)
#pragma unroll
                        for (int p = 0; p < PW; ++p)
                            acc[c][p] = fmaf(wv[c], av[STRIDE * p + kx + PADL - PAD], acc[c][p]);
                }
            }
        }
        __syncthreads();     // everyone is done with this stage before chunk ch+2 overwrites it
    }

    // ---- epilogue: bias (+ border classes), skip, ReLU, store, statistics
    const int rc = (oy == 0) ? 0 : ((oy == a.Ho - 1) ? 2 : 1);
    float ssum[CO_T], ssq[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
        const int co = cg * CO_T + c;
        float bmid, bl, br;
        if (a.bias_classes) {
            bl = __ldg(a.bias + (rc * 3 + 0) * a.Cout + co);
            bmid = __ldg(a.bias + (rc * 3 + 1) * a.Cout + co);
            br = __ldg(a.bias + (rc * 3 + 2) * a.Cout + co);
        } else {
            bl = bmid = br = __ldg(a.bias + co);
        }
        float o[PW];
#pragma unroll
        for (int p = 0; p < PW; ++p) o[p] = acc[c][p] + bmid;
        if (sx == 0) o[0] = acc[c][0] + bl;
        if (sx == k.SPR - 1) o[PW - 1] = acc[c][PW - 1] + br;
        const size_t off = (((size_t)b * a.Cout + co) * a.Ho + oy) * a.Wo + sx * PW;
        if (a.mask_src && live) {
            float ms = 1.f, mt = 0.f;
            if (a.mask_s) {
                const size_t mi = (a.mask_per_sample ? (size_t)b * a.Cout : 0) + co;
                ms = __ldg(a.mask_s + mi); mt = __ldg(a.mask_t + mi);
            }
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mask_src + off) + i);
                if (!(fmaf(m4.x, ms, mt) > 0.f)) o[4 * i] = 0.f;
                if (!(fmaf(m4.y, ms, mt) > 0.f)) o[4 * i + 1] = 0.f;
                if (!(fmaf(m4.z, ms, mt) > 0.f)) o[4 * i + 2] = 0.f;
                if (!(fmaf(m4.w, ms, mt) > 0.f)) o[4 * i + 3] = 0.f;
            }
        }
        if (a.skip && live) {
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.skip + off) + i);
                o[4 * i] += s4.x; o[4 * i + 1] += s4.y; o[4 * i + 2] += s4.z; o[4 * i + 3] += s4.w;
            }
        }
        if (a.out_relu) {
#pragma unroll
            for (int p = 0; p < PW; ++p) o[p] = fmaxf(o[p], 0.f);
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < NG; ++i)
                reinterpret_cast<float4*>(a.y + off)[i] =
                    make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
        float s = 0.f, q = 0.f;
        if (a.stats) {
            if (a.stat_src) {
                if (live) {
#pragma unroll
                    for (int i = 0; i < NG; ++i) {
                        const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.stat_src + off) + i);
                        s += o[4 * i] + o[4 * i + 1] + o[4 * i + 2] + o[4 * i + 3];
                        q = fmaf(o[4 * i], y4.x, q); q = fmaf(o[4 * i + 1], y4.y, q);
                        q = fmaf(o[4 * i + 2], y4.z, q); q = fmaf(o[4 * i + 3], y4.w, q);
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < PW; ++p) { s += o[p]; q = fmaf(o[p], o[p], q); }
            }
        }
        ssum[c] = s; ssq[c] = q;
    }

    if (a.stats) {
        // deterministic two-level reduction: thread partials -> smem -> one warp per (patch, channel)
        __syncthreads();
        float2* sp = reinterpret_cast<float2*>(smem);   // [NP][Cout][strips_per_patch]
#pragma unroll
        for (int c = 0; c < CO_T; ++c)
            sp[((size_t)pl * a.Cout + cg * CO_T + c) * strips_per_patch + prem] = make_float2(ssum[c], ssq[c]);
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
        for (int pc = warp; pc < k.NP * a.Cout; pc += nwarps) {
            const int lp = pc / a.Cout, co = pc - lp * a.Cout;
            double s = 0.0, q = 0.0;
            for (int i = lane; i < strips_per_patch; i += 32) {
                const float2 v = sp[(size_t)pc * strips_per_patch + i];
                s += (double)v.x; q += (double)v.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (lane == 0 && b0 + lp < a.B) {
                double* dst = a.stats + ((((size_t)(b0 + lp)) * k.nbands + band) * a.Cout + co) * 2;
                dst[0] = s; dst[1] = q;
            }
        }
    }
}

int plan(const ConvFwdArgs& a, int co_t, ConvK& k) {
    k.a = a;
    const int KS = a.ks, S = a.stride;
    const int ncg = a.Cout / co_t;
    k.SPR = a.Wo / PW;
    const int strips_per_patch_full = a.Ho * k.SPR;
    const int tpp = strips_per_patch_full * ncg;
    // 128 threads per CTA, unless that leaves most of the GPU idle (training batches: a few hundred patches of a
    // 16x16 map are only ~128 such CTAs, one warp per scheduler and nothing to hide latency with): then smaller CTAs
    int target = 128;
    {
        auto ctas = [&](int tgt) -> int64_t {
            if (tpp >= tgt) {
                int tr = 1;
                for (int t = 1; t <= a.Ho; ++t)
                    if (a.Ho % t == 0 && t * k.SPR * ncg <= tgt) tr = t;
                return (int64_t)a.B * (a.Ho / tr);
            }
            const int np = tgt / tpp;
            return ((int64_t)a.B + np - 1) / np;
        };
        // (3x3 only: the 1x1 data-gradient convs got slower with smaller CTAs, their weights are re-read per CTA)
        while (KS == 3 && target > 32 && ctas(target) < 4 * 148 && a.W / 4 <= target / 2) target >>= 1;
    }
    if (tpp >= target) {
        k.NP = 1;
        k.TR = 1;
        for (int tr = 1; tr <= a.Ho; ++tr)
            if (a.Ho % tr == 0 && tr * k.SPR * ncg <= target) k.TR = tr;
    } else {
        k.TR = a.Ho;
        k.NP = target / tpp;
        if (k.NP > a.B) k.NP = (int)a.B;
        if (k.NP < 1) k.NP = 1;
    }
    k.nbands = a.Ho / k.TR;
    k.threads = k.NP * k.TR * k.SPR * ncg;
    if (k.threads > 128 || k.threads < 1) return -1;
    k.RIN = (k.TR - 1) * S + KS;
    const int ngs = (S == 2) ? 4 : 2;
    const int granules = a.W / 4 + ((KS == 1) ? 0 : 2);     // data granules + one zero granule per side
    k.SUBW = ((granules + ngs - 1) / ngs) * 4;
    k.row_stride = ngs * k.SUBW;
    // Narrow maps put several output rows into one 8-lane shared-memory phase; pad the row so
    // that consecutive output rows start SPR granules apart (mod 32 banks) -> conflict-free.
    if (k.SPR < 8) {
        const int want = (k.SPR * 4) % 32;
        int rs = k.row_stride;
        for (int i = 0; i < 8 && (S * rs) % 32 != want; ++i) rs += 4;
        if ((S * rs) % 32 == want) k.row_stride = rs;
    }
    k.ci_stride = k.RIN * k.row_stride;
    // input-channel chunk: two pipeline stages of (weights + tile) under 54 KB so 4 CTAs share an SM;
    // prefer at least two chunks so that loads overlap the FMA core
    const int budget = 54 * 1024 / 4;
    k.CIC = 1;
    for (int c = 1; c <= a.Cin; ++c) {
        if (a.Cin % c) continue;
        if (a.Cin >= 2 && c > a.Cin / 2) break;
        const int fl = ((c * KS * KS * a.Cout + 3) & ~3) + k.NP * c * k.ci_stride;
        if (2 * fl <= budget) k.CIC = c;
    }
    k.patch_stride = k.CIC * k.ci_stride;
    k.w_floats = k.CIC * KS * KS * a.Cout;
    k.tile_floats = k.NP * k.patch_stride;
    k.stage_floats = ((k.w_floats + 3) & ~3) + k.tile_floats;
    return 0;
}

template <int KS, int STRIDE, int CO_T>
int launch(const ConvK& k, cudaStream_t st) {
    const ConvFwdArgs& a = k.a;
    size_t smem = (size_t)2 * k.stage_floats * sizeof(float);
    const size_t stats_smem = a.stats ? (size_t)k.NP * a.Cout * k.TR * k.SPR * sizeof(float2) : 0;
    if (stats_smem > smem) smem = stats_smem;
    auto kern = conv_fwd_kernel<KS, STRIDE, CO_T>;
    if (smem > 48 * 1024) {
        static size_t configured[64] = {0};   // per instantiation, per device
        int dev = 0;
        DMB_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    const int64_t groups = (a.B + k.NP - 1) / k.NP;
    const int64_t grid = groups * k.nbands;
    DMB_CHECK(grid > 0 && grid < (1ll << 31), "conv_fwd: grid %lld out of range", (long long)grid);
    DMB_LAUNCH((kern), (unsigned)grid, k.threads, smem, st, k);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int pick_co_t(int Cout) { return (Cout % 8 == 0) ? 8 : ((Cout % 4 == 0) ? 4 : ((Cout % 2 == 0) ? 2 : 0)); }

}  // namespace

int conv_fwd_bands(int ks, int stride, int Cin, int Cout, int Ho, int Wo, bool plain, int64_t B) {
    if (plain) {       // plain forward calls go to the TMA kernel when the geometry has an instantiation
        const int nb = conv_tma_bands(ks, stride, Cin, Cout, Ho * stride, Wo * stride, B);
        if (nb > 0) return nb;
    }
    ConvFwdArgs a{};
    a.ks = ks; a.stride = stride; a.Cin = Cin; a.Cout = Cout; a.Ho = Ho; a.Wo = Wo;
    a.H = (stride == 2) ? Ho * 2 : Ho; a.W = (stride == 2) ? Wo * 2 : Wo; a.B = (int)B;
    ConvK k;
    const int co_t = pick_co_t(Cout);
    if (!co_t || Wo % PW || plan(a, co_t, k)) return -1;
    return k.nbands;
}

int conv_fwd(const ConvFwdArgs& a, cudaStream_t st) {
    DMB_CHECK(a.B > 0, "conv_fwd: empty batch");
    DMB_CHECK((a.ks == 1 && a.stride == 1) || (a.ks == 3 && a.stride == 1) || (a.ks == 4 && a.stride == 2),
              "conv_fwd: unsupported kernel %dx%d stride %d", a.ks, a.ks, a.stride);
    DMB_CHECK(a.Ho * a.stride == a.H && a.Wo * a.stride == a.W, "conv_fwd: geometry mismatch");
    {
        const int r = conv_tma(a, st);      // 1 = not taken (unsupported shape or a data-gradient call)
        if (r <= 0) return r;
    }
    DMB_CHECK(!a.out_nhwc, "conv_fwd: channel-last output exists only for the plain 4x4 s2 2->32 @128 head (TMA kernel)");
    DMB_CHECK(a.Wo % PW == 0, "conv_fwd: output width %d must be a multiple of %d", a.Wo, PW);
    const int co_t = pick_co_t(a.Cout);
    DMB_CHECK(co_t != 0, "conv_fwd: Cout=%d must be even", a.Cout);
    ConvK k;
    DMB_CHECK(plan(a, co_t, k) == 0, "conv_fwd: no launch plan for Cout=%d Ho=%d Wo=%d", a.Cout, a.Ho, a.Wo);
    DMB_CHECK((size_t)2 * k.stage_floats * 4 <= 200 * 1024,
              "conv_fwd: tile does not fit shared memory (Cin=%d Cout=%d W=%d)", a.Cin, a.Cout, a.W);
    DMB_CHECK(a.W / 4 <= k.threads, "conv_fwd: input width %d too large for a %d-thread CTA", a.W, k.threads);
#define DMB_DISPATCH(KS, S)                                         \
    if (a.ks == KS && a.stride == S) {                              \
        if (co_t == 8) return launch<KS, S, 8>(k, st);              \
        if (co_t == 4) return launch<KS, S, 4>(k, st);              \
        return launch<KS, S, 2>(k, st);                             \
    }
    DMB_DISPATCH(1, 1)
    DMB_DISPATCH(3, 1)
    DMB_DISPATCH(4, 2)
#undef DMB_DISPATCH
    return -1;
}

}  // namespace dmb
