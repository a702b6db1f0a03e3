// ConvTranspose2d(k=4, stride=2, padding=1) forward on CUDA cores (fp32), NCHW planar.
// Reference layers: HiddenStateExtractor/vq_vae.py:292-296, vae.py:411-414.
//
// Sub-pixel form: output pixel (2m+py, 2n+px) receives input (iy, ix) through tap
// ky = oy + 1 - 2*iy, kx likewise, i.e. exactly two taps per axis:
//   py=0: (ky=1, iy=m) (ky=3, iy=m-1)      py=1: (ky=0, iy=m+1) (ky=2, iy=m)
// A thread owns 4 consecutive input columns of one input row = a 2x8 output block for CO_T
// output channels (64 accumulators at CO_T=4) and walks the 3x6 input neighbourhood per
// input channel; weights are warp-uniform 128-bit broadcasts from shared memory.
#include "common.cuh"

namespace dmb {
namespace {

constexpr int PWI = 4;   // input pixels per thread along x

struct ConvTK {
    ConvTFwdArgs a;
    int TR, NP, CIC, nbands, SPR, RIN, row_stride, ci_stride, patch_stride, w_floats, tile_floats, threads;
};

template <int CO_T>
__global__ void __launch_bounds__(256) convt_fwd_kernel(const ConvTK k) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    extern __shared__ __align__(16) float smem[];
    const ConvTFwdArgs& a = k.a;
    float* ws = smem;
    float* tile = smem + ((k.w_floats + 3) & ~3);
    const int tid = threadIdx.x;
    const int band = blockIdx.x % k.nbands;
    const int64_t b0 = (int64_t)(blockIdx.x / k.nbands) * k.NP;
    const int strips_per_patch = k.TR * k.SPR;
    const int strips = k.NP * strips_per_patch;
    const int cg = tid / strips;
    const int srem = tid - cg * strips;
    const int pl = srem / strips_per_patch;
    const int prem = srem - pl * strips_per_patch;
    const int row = prem / k.SPR;
    const int sx = prem - row * k.SPR;
    const int m = band * k.TR + row;           // input row
    const int64_t b = b0 + pl;
    const bool live = b < a.B;
    const int Ho = 2 * a.H, Wo = 2 * a.W;

    float acc[CO_T][2][2 * PWI];
#pragma unroll
    for (int c = 0; c < CO_T; ++c)
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int p = 0; p < 2 * PWI; ++p) acc[c][py][p] = 0.f;

    const int in_row0 = band * k.TR - 1;
    const int W4 = a.W >> 2;
    for (int c0 = 0; c0 < a.Cin; c0 += k.CIC) {
        __syncthreads();
        {
            const float* src = a.w + (size_t)c0 * 16 * a.Cout;
            for (int i = tid; i < k.w_floats; i += blockDim.x) ws[i] = __ldg(src + i);
        }
        {
            const int total = k.NP * k.CIC * k.RIN * W4;
            for (int e = tid; e < total; e += blockDim.x) {
                int q = e % W4;
                int t = e / W4;
                int r = t % k.RIN; t /= k.RIN;
                int cil = t % k.CIC;
                int lp = t / k.CIC;
                const int64_t bb = b0 + lp;
                const int ci = c0 + cil;
                const int iy = in_row0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bb < a.B && iy >= 0 && iy < a.H) {
                    v = __ldg(reinterpret_cast<const float4*>(
                            a.x + (((size_t)bb * a.Cin + ci) * a.H + iy) * a.W) + q);
                    const size_t ai = (a.in_per_sample ? (size_t)bb * a.Cin : 0) + ci;
                    if (a.x2) {
                        const float4 u = __ldg(reinterpret_cast<const float4*>(
                            a.x2 + (((size_t)bb * a.Cin + ci) * a.H + iy) * a.W) + q);
                        const float s = __ldg(a.in_scale + ai), bc = __ldg(a.in_b + ai), sh = __ldg(a.in_shift + ai);
                        v.x = fmaf(v.x, s, fmaf(u.x, bc, sh)); v.y = fmaf(v.y, s, fmaf(u.y, bc, sh));
                        v.z = fmaf(v.z, s, fmaf(u.z, bc, sh)); v.w = fmaf(v.w, s, fmaf(u.w, bc, sh));
                    } else if (a.in_scale) {
                        const float s = __ldg(a.in_scale + ai), sh = __ldg(a.in_shift + ai);
                        v.x = fmaf(v.x, s, sh); v.y = fmaf(v.y, s, sh);
                        v.z = fmaf(v.z, s, sh); v.w = fmaf(v.w, s, sh);
                    }
                    if (a.in_relu) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f);
                        v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                }
                float* rowp = tile + lp * k.patch_stride + cil * k.ci_stride + r * k.row_stride;
                rowp[4 * q + 1] = v.x; rowp[4 * q + 2] = v.y; rowp[4 * q + 3] = v.z; rowp[4 * q + 4] = v.w;
                if (q == 0) rowp[0] = 0.f;
                if (q == W4 - 1) rowp[a.W + 1] = 0.f;
            }
        }
        __syncthreads();

        const float* tp = tile + pl * k.patch_stride + row * k.row_stride + sx * PWI;
        const float* wp = ws + cg * CO_T;
        for (int cil = 0; cil < k.CIC; ++cil) {
            float av[3][8];
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) {
                const float* rp = tp + cil * k.ci_stride + dr * k.row_stride;
                const float4 u = *reinterpret_cast<const float4*>(rp);
                const float4 w4 = *reinterpret_cast<const float4*>(rp + 4);
                av[dr][0] = u.x; av[dr][1] = u.y; av[dr][2] = u.z; av[dr][3] = u.w;
                av[dr][4] = w4.x; av[dr][5] = w4.y; av[dr][6] = w4.z; av[dr][7] = w4.w;
            }
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const int py = (ky & 1) ? 0 : 1;
                const int dr = (ky == 1 || ky == 2) ? 1 : (ky == 3 ? 0 : 2);
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const int px = (kx & 1) ? 0 : 1;
                    const int dx = (kx == 1 || kx == 2) ? 0 : (kx == 3 ? -1 : 1);
                    float wv[CO_T];
                    const float* wrow = wp + ((cil * 4 + ky) * 4 + kx) * a.Cout;
                    if constexpr (CO_T == 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(wrow);
                        wv[0] = t4.x; wv[1] = t4.y; wv[2] = t4.z; wv[3] = t4.w;
                    } else {
#pragma unroll
                        for (int c = 0; c < CO_T; ++c) wv[c] = wrow[c];
                    }
#pragma unroll
                    for (int c = 0; c < CO_T; ++c)
#pragma unroll
                        for (int nl = 0; nl < PWI; ++nl)
                            acc[c][py][2 * nl + px] = fmaf(wv[c], av[dr][nl + 1 + dx], acc[c][py][2 * nl + px]);
                }
            }
        }
    }

    float ssum[CO_T], ssq[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
        const int co = cg * CO_T + c;
        const float bv = __ldg(a.bias + co);
        float s = 0.f, q = 0.f;
        float ms = 1.f, mt = 0.f;
        if (a.mask_src && a.mask_s) {
            const size_t mi = (a.mask_per_sample ? (size_t)b * a.Cout : 0) + co;
            if (live) { ms = __ldg(a.mask_s + mi); mt = __ldg(a.mask_t + mi); }
        }
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            float o[2 * PWI];
            const size_t off = (((size_t)b * a.Cout + co) * Ho + 2 * m + py) * Wo + sx * 2 * PWI;
#pragma unroll
            for (int p = 0; p < 2 * PWI; ++p) o[p] = acc[c][py][p] + bv;
            if (a.mask_src && live) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mask_src + off) + i);
                    if (!(fmaf(m4.x, ms, mt) > 0.f)) o[4 * i] = 0.f;
                    if (!(fmaf(m4.y, ms, mt) > 0.f)) o[4 * i + 1] = 0.f;
                    if (!(fmaf(m4.z, ms, mt) > 0.f)) o[4 * i + 2] = 0.f;
                    if (!(fmaf(m4.w, ms, mt) > 0.f)) o[4 * i + 3] = 0.f;
                }
            }
#pragma unroll
            for (int p = 0; p < 2 * PWI; ++p)
                if (a.out_relu) o[p] = fmaxf(o[p], 0.f);
            if (a.stats) {
                if (a.stat_src) {
                    if (live) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.stat_src + off) + i);
                            s += o[4 * i] + o[4 * i + 1] + o[4 * i + 2] + o[4 * i + 3];
                            q = fmaf(o[4 * i], y4.x, q); q = fmaf(o[4 * i + 1], y4.y, q);
                            q = fmaf(o[4 * i + 2], y4.z, q); q = fmaf(o[4 * i + 3], y4.w, q);
                        }
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < 2 * PWI; ++p) { s += o[p]; q = fmaf(o[p], o[p], q); }
                }
            }
            if (live) {
                float4* dst = reinterpret_cast<float4*>(a.y + off);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
        ssum[c] = s; ssq[c] = q;
    }
    if (a.stats) {
        __syncthreads();
        float2* sp = reinterpret_cast<float2*>(smem);
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

int pick_co_t(int Cout) { return (Cout % 4 == 0) ? 4 : ((Cout % 2 == 0) ? 2 : 0); }

int plan(const ConvTFwdArgs& a, int co_t, ConvTK& k) {
    k.a = a;
    const int ncg = a.Cout / co_t;
    k.SPR = a.W / PWI;
    const int tpp = a.H * k.SPR * ncg;
    const int target = 128;
    if (tpp >= target) {
        k.NP = 1; k.TR = 1;
        for (int tr = 1; tr <= a.H; ++tr)
            if (a.H % tr == 0 && tr * k.SPR * ncg <= target) k.TR = tr;
    } else {
        k.TR = a.H;
        k.NP = target / tpp;
        if (k.NP > a.B) k.NP = (int)a.B;
        if (k.NP < 1) k.NP = 1;
    }
    k.nbands = a.H / k.TR;
    k.threads = k.NP * k.TR * k.SPR * ncg;
    if (k.threads > 256 || k.threads < 1) return -1;
    k.RIN = k.TR + 2;
    k.row_stride = a.W + 8;              // padded row (W+2) + overrun of the second float4
    if (k.SPR < 8) {                     // rows start SPR granules apart (mod 32 banks)
        const int want = (k.SPR * 4) % 32;
        int rs = k.row_stride;
        for (int i = 0; i < 8 && rs % 32 != want; ++i) rs += 4;
        if (rs % 32 == want) k.row_stride = rs;
    }
    k.ci_stride = k.RIN * k.row_stride;
    const int budget = 54 * 1024 / 4;
    k.CIC = 1;
    for (int c = 1; c <= a.Cin; ++c) {
        if (a.Cin % c) continue;
        if (c * 16 * a.Cout + k.NP * c * k.ci_stride <= budget) k.CIC = c;
    }
    k.patch_stride = k.CIC * k.ci_stride;
    k.w_floats = k.CIC * 16 * a.Cout;
    k.tile_floats = k.NP * k.patch_stride;
    return 0;
}

template <int CO_T>
int launch(const ConvTK& k, cudaStream_t st) {
    const ConvTFwdArgs& a = k.a;
    size_t smem = (size_t)(((k.w_floats + 3) & ~3) + k.tile_floats) * sizeof(float);
    const size_t stats_smem = a.stats ? (size_t)k.NP * a.Cout * k.TR * k.SPR * sizeof(float2) : 0;
    if (stats_smem > smem) smem = stats_smem;
    auto kern = convt_fwd_kernel<CO_T>;
    if (smem > 48 * 1024) {
        static size_t configured[64] = {0};
        int dev = 0;
        DMB_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || smem > configured[dev]) {
            DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) configured[dev] = smem;
        }
    }
    const int64_t grid = ((a.B + k.NP - 1) / k.NP) * k.nbands;
    DMB_CHECK(grid > 0 && grid < (1ll << 31), "convt_fwd: grid out of range");
    DMB_LAUNCH((kern), (unsigned)grid, k.threads, smem, st, k);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace

int convt_fwd_bands(int Cin, int Cout, int H, int W) {
    ConvTFwdArgs a{};
    a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W; a.B = 1 << 20;
    ConvTK k;
    const int co_t = pick_co_t(Cout);
    if (!co_t || W % PWI || plan(a, co_t, k)) return -1;
    return k.nbands;
}

int convt_fwd(const ConvTFwdArgs& a, cudaStream_t st) {
    DMB_CHECK(a.B > 0, "convt_fwd: empty batch");
    DMB_CHECK(a.W % PWI == 0, "convt_fwd: input width %d must be a multiple of %d", a.W, PWI);
    const int co_t = pick_co_t(a.Cout);
    DMB_CHECK(co_t != 0, "convt_fwd: Cout=%d must be even", a.Cout);
    ConvTK k;
    DMB_CHECK(plan(a, co_t, k) == 0, "convt_fwd: no launch plan (Cout=%d H=%d W=%d)", a.Cout, a.H, a.W);
    DMB_CHECK((size_t)(k.w_floats + k.tile_floats) * 4 <= 200 * 1024, "convt_fwd: tile does not fit shared memory");
    return co_t == 4 ? launch<4>(k, st) : launch<2>(k, st);
}

}  // namespace dmb
