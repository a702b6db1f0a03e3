// Weight gradients of the default-width layers, TMA-fed with compile-time geometry (sm_100a).
// (autograd of F.conv2d / F.conv_transpose2d w.r.t. weight and bias; reference call site run_training.py:406.)
//
//   dW[ci][ky][kx][co] = sum_{b,oy,ox} gy[b][co][oy][ox] * act[b][ci][S*oy - P + ky][S*ox - P + kx]
//
// Same ownership and arithmetic as wgrad.cu (a thread owns the taps of TCI input channels x TCO output channels in
// registers for the whole persistent CTA and walks strips of four output pixels; splits of the pixel range are
// folded through shared memory; one partial per CTA; wgrad_reduce folds the partials in fixed order).  What changed
// is everything around the FMA loop: ncu on wgrad.cu showed FFMA at 30-40 % of the issued instructions -- the
// per-thread global->shared staging with run-time geometry (four div/mod per 16 bytes) was 40 % of the instruction
// stream and ran serialised with the FMA loop.  Here
//   * the RAW tiles of a work item (activation band with halo, output-gradient band, and the raw conv output when a
//     BatchNorm backward is folded in) are three cp.async.bulk.tensor.4d (UTMALDG) issued by one thread; rows above /
//     below the image are the TMA out-of-bounds fill;
//   * a short in-shared-memory pass applies the producer's BatchNorm affine + ReLU / the BatchNorm-backward
//     combination  g*A[c] + y*Bc[c] + Cc[c]  and writes the bank-conflict-free padded layout the FMA loop reads
//     (channel stride = 4 * odd floats, so the 8 channels a warp touches hit 8 different bank groups and every
//     128-bit load is one wavefront; lanes that own different output-channel groups of the same channel broadcast);
//   * the TMA of item i+1 is issued as soon as that pass has consumed the raw tiles and lands under the FMA loop of
//     item i.
#include "common.cuh"

#include <stdlib.h>

#include <cuda.h>
#include <cudaTypedefs.h>

namespace dmb {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
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

constexpr int WT_THREADS = 256;
constexpr int r32(int v) { return (v + 31) & ~31; }

template <int KS_, int S_, int CIN_, int ONES_, int COUT_, int W_, int TCI_, int TCO_, int TRO_>
struct WG {
    static constexpr int KS = KS_, S = S_, CIN = CIN_, ONES = ONES_, COUT = COUT_, W = W_, H = W_, TCI = TCI_, TCO = TCO_;
    static constexpr int TRO = TRO_;
    static constexpr int CINE = CIN + ONES;
    static constexpr int WO = W / S, HO = H / S;
    static constexpr int PAD = (KS == 1) ? 0 : 1;
    static constexpr int PADL = PAD ? 4 : 0;
    static constexpr int NT = KS * KS;
    static constexpr int NBANDS = HO / TRO;
    static constexpr int RIN = (TRO - 1) * S + KS;
    static constexpr int RINP = RIN | 1;                          // odd
    static constexpr int P = W + (PAD ? 12 : 4);                  // row pitch: an odd number of 16-byte granules
    static constexpr int XST = RINP * P;                          // channel stride: 4 * odd floats
    static constexpr int GP = WO + 4;
    static constexpr int TROP = TRO | 1;
    static constexpr int GST = TROP * GP;
    static constexpr int NCG = COUT / TCO, NCIG = CINE / TCI;
    static constexpr int OWNERS = NCG * NCIG;
    static constexpr int W4 = W / 4, WO4 = WO / 4;
    static constexpr int NSTRIPS = TRO * WO4;
    // pixel splits: as many as fit 256 threads, but never more than there are strips, and a divisor of the strip count
    static constexpr int pick_nsplit() {
        int n = WT_THREADS / OWNERS;
        if (n > NSTRIPS) n = NSTRIPS;
        while (NSTRIPS % n) --n;
        return n;
    }
    static constexpr int NSPLIT = pick_nsplit();
    static constexpr int NV = (KS == 1) ? 1 : ((S == 2) ? 4 : 3);
    static constexpr int XR = r32(CIN * RIN * W);                 // raw tiles (floats), 128-byte multiples
    static constexpr int GR = r32(COUT * TRO * WO);
    static constexpr int TX = r32(CINE * XST);
    static constexpr int TG = r32(COUT * GST + 4 * NCG);
    static constexpr int PER = TCI * TCO * NT + TCO;              // floats a thread hands to the final fold
    static constexpr int RED = NSPLIT * OWNERS * PER;
    static constexpr int OUT_FLOATS = CINE * NT * COUT + COUT;
    // shared-memory body (floats) without / with the raw conv-output tile of a folded BatchNorm backward; the fold
    // buffer of the last phase overlays it
    static constexpr int BODY1 = (XR + GR + TX + TG) > RED ? (XR + GR + TX + TG) : RED;
    static constexpr int BODY2 = (XR + 2 * GR + TX + TG) > RED ? (XR + 2 * GR + TX + TG) : RED;
    static constexpr size_t smem_bytes(bool dual) { return 1024 + (size_t)(dual ? BODY2 : BODY1) * 4 + 64; }
    static_assert(W % 4 == 0 && WO % 4 == 0 && HO % TRO == 0, "geometry");
    static_assert((P / 4) % 2 == 1 && (GP / 4) % 2 == 1, "row pitches must be an odd number of granules");
    static_assert(COUT % TCO == 0 && CINE % TCI == 0 && OWNERS * NSPLIT <= WT_THREADS, "ownership");
    static_assert((CIN * RIN * W) % 4 == 0 && (COUT * TRO * WO) % 4 == 0, "tiles are whole float4s");
};

struct WtArgs {
    WgradArgs a;
    int64_t work;       // B * NBANDS
    int dual;           // BatchNorm backward folded in: g' = g*ga + y*gb + gc
};

template <class C>
__global__ void __launch_bounds__(WT_THREADS, 2)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                 const __grid_constant__ CUtensorMap map_y, const WtArgs k) {
    constexpr int KS = C::KS, S = C::S, TCI = C::TCI, TCO = C::TCO, NT = C::NT, NV = C::NV, PAD = C::PAD, PADL = C::PADL;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
    const WgradArgs& a = k.a;
    const bool dual = k.dual != 0;
    float* raw_x = sm;
    float* raw_g = raw_x + C::XR;
    float* raw_y = raw_g + C::GR;
    float* tx = raw_g + C::GR * (dual ? 2 : 1);
    float* tg = tx + C::TX;
    const int body = dual ? C::BODY2 : C::BODY1;
    const uint32_t bar = base + (uint32_t)body * 4u;

    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(bar, 1u); fence_barrier_init(); }
    // the padding of the transformed tiles is written once: the per-item pass only touches image columns
    for (int i = tid; i < C::TX + C::TG; i += WT_THREADS) tx[i] = 0.f;
    __syncthreads();
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output

    const uint32_t tx_bytes = (uint32_t)(C::CIN * C::RIN * C::W + C::COUT * C::TRO * C::WO * (dual ? 2 : 1)) * 4u;
    auto issue = [&](int64_t wk) {
        const int b = (int)(wk / C::NBANDS);
        const int band = (int)(wk - (int64_t)b * C::NBANDS);
        mbar_expect_tx(bar, tx_bytes);
        tma_load_4d(smem_u32(raw_x), &map_x, bar, 0, band * C::TRO * S - PAD, b, 0);
        tma_load_4d(smem_u32(raw_g), &map_g, bar, 0, band * C::TRO, b, 0);
        if (dual) tma_load_4d(smem_u32(raw_y), &map_y, bar, 0, band * C::TRO, b, 0);
    };
    if (tid == 0 && (int64_t)blockIdx.x < k.work) issue(blockIdx.x);

    const int owner = tid % C::OWNERS;
    const int split = tid / C::OWNERS;
    const int cg = owner % C::NCG;
    const int cig = owner / C::NCG;
    const bool worker = split < C::NSPLIT;

    float acc[TCI][TCO][NT];
    float dbacc[TCO];
#pragma unroll
    for (int c = 0; c < TCO; ++c) {
        dbacc[c] = 0.f;
#pragma unroll
        for (int j = 0; j < TCI; ++j)
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[j][c][t] = 0.f;
    }

    uint32_t phase = 0;
    for (int64_t wk = blockIdx.x; wk < k.work; wk += gridDim.x) {
        const int b = (int)(wk / C::NBANDS);
        const int band = (int)(wk - (int64_t)b * C::NBANDS);
        const int in_row0 = band * C::TRO * S - PAD;
        mbar_wait_warp(bar, phase);
        phase ^= 1u;
        // ---- activation band: relu?(x * xs[c] + xt[c]) inside the image, zero outside
        {
            constexpr int TOTAL = C::CIN * C::RIN * C::W4;
            for (int e = tid; e < TOTAL; e += WT_THREADS) {
                const int q = e % C::W4;
                const int t = e / C::W4;
                const int r = t % C::RIN;
                const int c = t / C::RIN;
                float4 v = *reinterpret_cast<const float4*>(raw_x + (size_t)e * 4);
                const int iy = in_row0 + r;
                if (iy >= 0 && iy < C::H) {
                    if (a.xs) {
                        const size_t ai = (a.x_per_sample ? (size_t)b * C::CIN : 0) + c;
                        const float sc = __ldg(a.xs + ai), sh = __ldg(a.xt + ai);
                        v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
                    }
                    if (a.x_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                } else {
                    v = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                *reinterpret_cast<float4*>(tx + c * C::XST + r * C::P + PADL + 4 * q) = v;
            }
            if constexpr (C::ONES != 0) {                   // the constant-one channel of the composite head
                constexpr int TOT1 = C::RIN * C::W4;
                for (int e = tid; e < TOT1; e += WT_THREADS) {
                    const int q = e % C::W4, r = e / C::W4;
                    const int iy = in_row0 + r;
                    const float o = (iy >= 0 && iy < C::H) ? 1.f : 0.f;
                    *reinterpret_cast<float4*>(tx + C::CIN * C::XST + r * C::P + PADL + 4 * q) = make_float4(o, o, o, o);
                }
            }
        }
        // ---- output-gradient band, BatchNorm backward folded in
        {
            constexpr int TOTAL = C::COUT * C::TRO * C::WO4;
            for (int e = tid; e < TOTAL; e += WT_THREADS) {
                const int q = e % C::WO4;
                const int t = e / C::WO4;
                const int r = t % C::TRO;
                const int c = t / C::TRO;
                float4 v = *reinterpret_cast<const float4*>(raw_g + (size_t)e * 4);
                if (a.ga) {
                    const size_t ai = (a.g_per_sample ? (size_t)b * C::COUT : 0) + c;
                    const float sc = __ldg(a.ga + ai), sh = __ldg(a.gc + ai);
                    if (dual) {
                        const float4 u = *reinterpret_cast<const float4*>(raw_y + (size_t)e * 4);
                        const float bc = __ldg(a.gb + ai);
                        v.x = fmaf(v.x, sc, fmaf(u.x, bc, sh)); v.y = fmaf(v.y, sc, fmaf(u.y, bc, sh));
                        v.z = fmaf(v.z, sc, fmaf(u.z, bc, sh)); v.w = fmaf(v.w, sc, fmaf(u.w, bc, sh));
                    } else {
                        v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
                    }
                }
                if (a.g_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4*>(tg + c * C::GST + 4 * (c / TCO) + r * C::GP + 4 * q) = v;
            }
        }
        __syncthreads();                                   // transformed tiles complete, raw tiles consumed
        if (tid == 0 && wk + gridDim.x < k.work) issue(wk + gridDim.x);
        if (worker) {
            const float* gc = tg + (cg * TCO) * C::GST + 4 * cg;
            for (int sp = split; sp < C::NSTRIPS; sp += C::NSPLIT) {
                const int r = sp / C::WO4;
                const int ox0 = (sp % C::WO4) * 4;
                float gv[TCO][4];
#pragma unroll
                for (int c = 0; c < TCO; ++c) {
                    const float4 t4 = *reinterpret_cast<const float4*>(gc + c * C::GST + r * C::GP + ox0);
                    gv[c][0] = t4.x; gv[c][1] = t4.y; gv[c][2] = t4.z; gv[c][3] = t4.w;
                }
                if (cig == 0) {
#pragma unroll
                    for (int c = 0; c < TCO; ++c) dbacc[c] += (gv[c][0] + gv[c][1]) + (gv[c][2] + gv[c][3]);
                }
#pragma unroll
                for (int j = 0; j < TCI; ++j) {
                    const float* xc = tx + (cig * TCI + j) * C::XST + (r * S) * C::P + S * ox0;
#pragma unroll
                    for (int ky = 0; ky < KS; ++ky) {
                        float xv[NV * 4];
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            const float4 t4 = *reinterpret_cast<const float4*>(xc + ky * C::P + 4 * i);
                            xv[4 * i] = t4.x; xv[4 * i + 1] = t4.y; xv[4 * i + 2] = t4.z; xv[4 * i + 3] = t4.w;
                        }
#pragma unroll
                        for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                            for (int c = 0; c < TCO; ++c)
#pragma unroll
                                for (int p = 0; p < 4; ++p)
                                    acc[j][c][ky * KS + kx] = fmaf(gv[c][p], xv[S * p + kx + PADL - PAD], acc[j][c][ky * KS + kx]);
                    }
                }
            }
        }
        __syncthreads();                                   // everyone is done with the transformed tiles
    }

    // ---- fold the pixel-splits (fixed order) and write this CTA's partial
    float* red = sm;                         // [NSPLIT][OWNERS][PER]  (no TMA is in flight: every issued item was waited for)
    constexpr int PER = C::PER;
    if (worker) {
        float* dst = red + ((size_t)split * C::OWNERS + owner) * PER;
#pragma unroll
        for (int j = 0; j < TCI; ++j)
#pragma unroll
            for (int c = 0; c < TCO; ++c)
#pragma unroll
                for (int t = 0; t < NT; ++t) dst[(j * TCO + c) * NT + t] = acc[j][c][t];
#pragma unroll
        for (int c = 0; c < TCO; ++c) dst[TCI * TCO * NT + c] = dbacc[c];
    }
    __syncthreads();
    float* out = a.partials + (size_t)blockIdx.x * C::OUT_FLOATS;
    for (int e = tid; e < C::OWNERS * PER; e += WT_THREADS) {
        const int o = e / PER, jj = e - o * PER;
        float s = 0.f;
#pragma unroll
        for (int sp = 0; sp < C::NSPLIT; ++sp) s += red[((size_t)sp * C::OWNERS + o) * PER + jj];
        const int ocg = o % C::NCG, ocig = o / C::NCG;
        if (jj < TCI * TCO * NT) {
            const int j = jj / (TCO * NT), rem = jj - j * (TCO * NT);
            const int c = rem / NT, t = rem - c * NT;
            out[((size_t)(ocig * TCI + j) * NT + t) * C::COUT + ocg * TCO + c] = s;      // packed [ci][ky][kx][co]
        } else if (ocig == 0) {
            out[(size_t)C::CINE * NT * C::COUT + ocg * TCO + (jj - TCI * TCO * NT)] = s;   // db[co]
        }
    }
}

PFN_cuTensorMapEncodeTiled wt_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }();
    return fn;
}

// (Wd, Hd, B, Cd) view of an NCHW tensor, box = (Wd, rows, 1, Cd)
int make_map(CUtensorMap* map, const float* p, int Wd, int Hd, int64_t B, int Cd, int rows) {
    PFN_cuTensorMapEncodeTiled enc = wt_encoder();
    DMB_CHECK(enc != nullptr, "wgrad_tma: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[4] = {(cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)B, (cuuint64_t)Cd};
    const cuuint64_t gstr[3] = {(cuuint64_t)Wd * 4, (cuuint64_t)Wd * Hd * Cd * 4, (cuuint64_t)Wd * Hd * 4};
    const cuuint32_t box[4] = {(cuuint32_t)Wd, (cuuint32_t)rows, 1u, (cuuint32_t)Cd};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DMB_CHECK(r == CUDA_SUCCESS, "wgrad_tma: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

constexpr int WT_CTAS = 148 * 2;

template <class C>
int launch_wt(const WgradArgs& a, int* ncta_out, int* out_floats, bool query, cudaStream_t st) {
    WtArgs k{};
    k.a = a;
    k.work = (int64_t)a.B * C::NBANDS;
    k.dual = (a.ga && a.y) ? 1 : 0;
    int ncta = WT_CTAS;
    if (k.work < ncta) ncta = (int)k.work;
    *ncta_out = ncta;
    *out_floats = C::OUT_FLOATS;
    if (query) return 0;
    CUtensorMap mx, mg, my;
    DMB_TRY(make_map(&mx, a.x, C::W, C::H, a.B, C::CIN, C::RIN));
    DMB_TRY(make_map(&mg, a.g, C::WO, C::HO, a.B, C::COUT, C::TRO));
    DMB_TRY(make_map(&my, k.dual ? a.y : a.g, C::WO, C::HO, a.B, C::COUT, C::TRO));
    auto kern = wgrad_tma_kernel<C>;
    const size_t smem = C::smem_bytes(true);
    int dev = 0;
    DMB_CUDA(cudaGetDevice(&dev));
    DMB_CHECK(dev >= 0 && dev < 64, "wgrad_tma: device index %d out of range", dev);
    static bool configured[64] = {false};      // per instantiation and device
    if (!configured[dev]) {
        DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    DMB_LAUNCH((kern), ncta, WT_THREADS, C::smem_bytes(k.dual != 0), st, mx, mg, my, k);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

bool wt_enabled() {
    const char* e = getenv("DMB_WGRAD_TMA");
    return !(e && e[0] == '0');
}

// dispatch on the layer shape; returns 1 when there is no instantiation
template <class F>
int dispatch(const WgradArgs& a, F&& f) {
    if (a.H != a.W || a.Ho != a.Wo) return 1;
    const int ones = a.ones_channel ? 1 : 0;
#define WT_CASE(KS, S, CIN, ONES, COUT, WW, TCI, TCO, TRO)                                                     \
    if (a.ks == KS && a.stride == S && a.Cin == CIN && ones == ONES && a.Cout == COUT && a.W == WW)          \
        return f(WG<KS, S, CIN, ONES, COUT, WW, TCI, TCO, TRO>{});
    // the default architecture (num_hiddens 16, num_residual_hiddens 32) on 128 x 128 patches
    WT_CASE(4, 2, 2, 1, 8, 128, 1, 4, 4)        // composite head (virtual constant-one channel)
    WT_CASE(4, 2, 8, 0, 16, 64, 1, 4, 4)        // enc.4
    WT_CASE(4, 2, 16, 0, 16, 32, 1, 4, 4)       // enc.7
    WT_CASE(3, 1, 16, 0, 16, 16, 1, 4, 8)       // enc.10
    WT_CASE(3, 1, 16, 0, 32, 16, 1, 4, 8)       // residual 3x3
    WT_CASE(1, 1, 32, 0, 16, 16, 4, 4, 8)       // residual 1x1
    WT_CASE(4, 2, 8, 0, 16, 32, 1, 4, 8)        // dec.0 ConvTranspose (roles swapped)
    WT_CASE(4, 2, 4, 0, 8, 64, 1, 4, 8)         // dec.2
    WT_CASE(4, 2, 4, 0, 4, 128, 1, 4, 4)        // dec.4
#undef WT_CASE
    return 1;
}

}  // namespace

// 1: not taken (unsupported shape / dual-tensor activation transform / disabled)
int wgrad_tma_plan(const WgradArgs& a, int* ncta, int* out_floats) {
    if (!wt_enabled() || a.x2) return 1;
    return dispatch(a, [&](auto c) {
        using C = decltype(c);
        return launch_wt<C>(a, ncta, out_floats, true, nullptr);
    });
}

int wgrad_tma(const WgradArgs& a, int* ncta, int* out_floats, cudaStream_t st) {
    if (!wt_enabled() || a.x2) return 1;
    DMB_CHECK(a.partials != nullptr, "wgrad_tma: no partial buffer");
    DMB_CHECK(!(reinterpret_cast<uintptr_t>(a.x) & 15) && !(reinterpret_cast<uintptr_t>(a.g) & 15) &&
              !(a.y && (reinterpret_cast<uintptr_t>(a.y) & 15)), "wgrad_tma: tensors must be 16-byte aligned");
    return dispatch(a, [&](auto c) {
        using C = decltype(c);
        return launch_wt<C>(a, ncta, out_floats, false, st);
    });
}

}  // namespace dmb
