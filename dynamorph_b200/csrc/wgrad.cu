// Weight gradients of the convolutions (autograd of F.conv2d / F.conv_transpose2d w.r.t. weight and
// bias; reference call site: run_training.py:406 `total_loss.backward()`).
//
//   dW[co][ci][ky][kx] = sum_{b,oy,ox} gy[b][co][oy][ox] * act[b][ci][S*oy - P + ky][S*ox - P + kx]
//
// Persistent CTAs walk (patch, row band) work items.  A thread OWNS the taps of one input channel for
// TCO output channels (TCO*KS*KS accumulators, kept in registers for the whole kernel) and visits the
// band's pixels in strips of four; several pixel-splits of the same owners run side by side and are
// folded through shared memory at the end.  Each CTA writes ONE partial; wgrad_reduce sums the partials
// in a fixed order (deterministic) and scatters into the torch weight layout.
//
// gy and act are formed on load, so the backward pass never materialises BatchNorm/ReLU outputs:
//   gy  = g*ga[c] + y*gb[c] + gc[c]      (BatchNorm backward folded in; y = the raw conv output)
//   act = relu?(x*xs[c] + xt[c])         (BatchNorm + ReLU of the producer layer, recomputed)
// ConvTranspose2d weight gradients use the same kernel with the roles of the two tensors swapped.
#include "common.cuh"

#include <stdlib.h>

namespace dmb {
namespace {


struct WgK {
    WgradArgs a;
    int TRO, nbands, SPW, RIN, RSW, ci_stride, g_stride, cin_eff, ncg, owners, nsplit, threads, ci_per, slices, tco;
    int x_floats, g_floats, out_floats;
    int64_t work;
};

template <int KS, int STRIDE, int TCO>
__global__ void __launch_bounds__(256, (TCO * KS * KS >= 64) ? 2 : 3) wgrad_kernel(const WgK k) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    extern __shared__ __align__(16) float smem[];
    const WgradArgs& a = k.a;
    constexpr int PAD = (KS == 1) ? 0 : 1;
    constexpr int PADL = PAD ? 4 : 0;
    constexpr int NV = (KS == 1) ? 1 : ((STRIDE == 2) ? 4 : 3);
    constexpr int NT = KS * KS;
    float* xs = smem;                       // [cin_eff][RIN][RSW]   (ci_stride padded)
    float* gs = smem + k.x_floats;          // [Cout][TRO][Wo]       (g_stride padded)

    const int tid = threadIdx.x;
    const int owner = tid % k.owners;
    const int split = tid / k.owners;
    const int cg = owner % k.ncg;
    const int ci = owner / k.ncg;                 // channel inside this CTA's slice
    const int ci0 = blockIdx.y * k.ci_per;        // first input channel of the slice
    const bool worker = split < k.nsplit;

    float acc[TCO][NT];
    float dbacc[TCO];
#pragma unroll
    for (int c = 0; c < TCO; ++c) {
        dbacc[c] = 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[c][t] = 0.f;
    }

    const int W4 = a.W >> 2, Wo4 = a.Wo >> 2;
    for (int64_t wk = blockIdx.x; wk < k.work; wk += gridDim.x) {
        const int64_t b = wk / k.nbands;
        const int band = (int)(wk - b * k.nbands);
        const int oy0 = band * k.TRO;
        const int in_row0 = oy0 * STRIDE - PAD;
        __syncthreads();
        // ---- activation band (with halo), transform applied.  Both staging loops issue the global loads of LU
        // elements before touching any of them (the loops were one dependent load -> transform -> store per trip and
        // ncu showed the kernel stalled on long-scoreboard 4-6 warps per issue).
        constexpr int LU = (TCO * NT >= 32) ? 2 : 4;      // live accumulators leave room for two float4 pairs only
        {
            const int total = k.ci_per * k.RIN * W4;
            for (int e0 = tid; e0 < total; e0 += LU * blockDim.x) {
                float4 v[LU], u[LU];
                int q_[LU], r_[LU], cl_[LU];
                bool in_img[LU], real_c[LU];
#pragma unroll
                for (int j = 0; j < LU; ++j) {
                    const int e = e0 + j * blockDim.x;
                    const int ee = e < total ? e : 0;
                    q_[j] = ee % W4;
                    const int t = ee / W4;
                    r_[j] = t % k.RIN;
                    cl_[j] = t / k.RIN;
                    const int c = ci0 + cl_[j];
                    const int iy = in_row0 + r_[j];
                    in_img[j] = e < total && iy >= 0 && iy < a.H;
                    real_c[j] = c < a.Cin;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    u[j] = v[j];
                    if (in_img[j] && real_c[j]) {
                        const size_t gi = (((size_t)b * a.Cin + c) * a.H + iy) * a.W;
                        v[j] = __ldg(reinterpret_cast<const float4*>(a.x + gi) + q_[j]);
                        if (a.xs && a.x2) u[j] = __ldg(reinterpret_cast<const float4*>(a.x2 + gi) + q_[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < LU; ++j) {
                    if (e0 + j * blockDim.x >= total) break;
                    const int c = ci0 + cl_[j];
                    float4 w = v[j];
                    if (in_img[j]) {
                        if (real_c[j]) {
                            if (a.xs) {
                                const size_t ai = (a.x_per_sample ? (size_t)b * a.Cin : 0) + c;
                                const float sc = __ldg(a.xs + ai), sh = __ldg(a.xt + ai);
                                if (a.x2) {
                                    const float bc = __ldg(a.xb + ai);
                                    w.x = fmaf(w.x, sc, fmaf(u[j].x, bc, sh)); w.y = fmaf(w.y, sc, fmaf(u[j].y, bc, sh));
                                    w.z = fmaf(w.z, sc, fmaf(u[j].z, bc, sh)); w.w = fmaf(w.w, sc, fmaf(u[j].w, bc, sh));
                                } else {
                                    w.x = fmaf(w.x, sc, sh); w.y = fmaf(w.y, sc, sh); w.z = fmaf(w.z, sc, sh); w.w = fmaf(w.w, sc, sh);
                                }
                            }
                            if (a.x_relu) {
                                w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
                            }
                        } else {
                            w = make_float4(1.f, 1.f, 1.f, 1.f);     // the constant-one channel of the composite head
                        }
                    }
                    float* rowp = xs + cl_[j] * k.ci_stride + r_[j] * k.RSW;
                    *reinterpret_cast<float4*>(rowp + PADL + 4 * q_[j]) = w;
                    if (PAD) {
                        if (q_[j] == 0) *reinterpret_cast<float4*>(rowp) = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (q_[j] == W4 - 1) *reinterpret_cast<float4*>(rowp + PADL + a.W) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
        // ---- output-gradient band, BatchNorm backward folded in
        {
            const int total = a.Cout * k.TRO * Wo4;
            for (int e0 = tid; e0 < total; e0 += LU * blockDim.x) {
                float4 v[LU], u[LU];
                int q_[LU], r_[LU], c_[LU];
#pragma unroll
                for (int j = 0; j < LU; ++j) {
                    const int e = e0 + j * blockDim.x;
                    const int ee = e < total ? e : 0;
                    q_[j] = ee % Wo4;
                    const int t = ee / Wo4;
                    r_[j] = t % k.TRO;
                    c_[j] = t / k.TRO;
                    const size_t gi = (((size_t)b * a.Cout + c_[j]) * a.Ho + oy0 + r_[j]) * a.Wo;
                    v[j] = __ldg(reinterpret_cast<const float4*>(a.g + gi) + q_[j]);
                    u[j] = (a.ga && a.y) ? __ldg(reinterpret_cast<const float4*>(a.y + gi) + q_[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < LU; ++j) {
                    if (e0 + j * blockDim.x >= total) break;
                    float4 w = v[j];
                    if (a.ga) {
                        const size_t ai = (a.g_per_sample ? (size_t)b * a.Cout : 0) + c_[j];
                        const float sc = __ldg(a.ga + ai), sh = __ldg(a.gc + ai);
                        if (a.y) {
                            const float bc = __ldg(a.gb + ai);
                            w.x = fmaf(w.x, sc, fmaf(u[j].x, bc, sh)); w.y = fmaf(w.y, sc, fmaf(u[j].y, bc, sh));
                            w.z = fmaf(w.z, sc, fmaf(u[j].z, bc, sh)); w.w = fmaf(w.w, sc, fmaf(u[j].w, bc, sh));
                        } else {
                            w.x = fmaf(w.x, sc, sh); w.y = fmaf(w.y, sc, sh); w.z = fmaf(w.z, sc, sh); w.w = fmaf(w.w, sc, sh);
                        }
                    }
                    if (a.g_relu) {
                        w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(gs + c_[j] * k.g_stride + r_[j] * a.Wo + 4 * q_[j]) = w;
                }
            }
        }
        __syncthreads();
        if (!worker) continue;
        const float* xc = xs + ci * k.ci_stride;
        const float* gc = gs + (cg * TCO) * k.g_stride;
        const int nstrips = k.TRO * Wo4;
        for (int sp = split; sp < nstrips; sp += k.nsplit) {
            const int r = sp / Wo4;
            const int ox0 = (sp - r * Wo4) * 4;
            float gv[TCO][4];
#pragma unroll
            for (int c = 0; c < TCO; ++c) {
                const float4 t4 = *reinterpret_cast<const float4*>(gc + c * k.g_stride + r * a.Wo + ox0);
                gv[c][0] = t4.x; gv[c][1] = t4.y; gv[c][2] = t4.z; gv[c][3] = t4.w;
            }
            if (ci == 0 && ci0 == 0) {
#pragma unroll
                for (int c = 0; c < TCO; ++c) dbacc[c] += (gv[c][0] + gv[c][1]) + (gv[c][2] + gv[c][3]);
            }
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const float* rp = xc + (r * STRIDE + ky) * k.RSW + STRIDE * ox0;
                float xv[NV * 4];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const float4 t4 = *reinterpret_cast<const float4*>(rp + 4 * i);
                    xv[4 * i] = t4.x; xv[4 * i + 1] = t4.y; xv[4 * i + 2] = t4.z; xv[4 * i + 3] = t4.w;
                }
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                    for (int c = 0; c < TCO; ++c)
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            acc[c][ky * KS + kx] = fmaf(gv[c][p], xv[STRIDE * p + kx + PADL - PAD], acc[c][ky * KS + kx]);
            }
        }
    }

    // ---- fold the pixel-splits (fixed order) and write this CTA's partial
    __syncthreads();
    float* red = smem;                       // [nsplit][owners][TCO*NT + TCO]
    constexpr int PER = TCO * NT + TCO;
    if (worker) {
        float* dst = red + ((size_t)split * k.owners + owner) * PER;
#pragma unroll
        for (int c = 0; c < TCO; ++c) {
#pragma unroll
            for (int t = 0; t < NT; ++t) dst[c * NT + t] = acc[c][t];
            dst[TCO * NT + c] = dbacc[c];
        }
    }
    __syncthreads();
    float* out = a.partials + (size_t)blockIdx.x * k.out_floats;
    for (int e = tid; e < k.owners * PER; e += blockDim.x) {
        const int o = e / PER, j = e - o * PER;
        float s = 0.f;
        for (int sp = 0; sp < k.nsplit; ++sp) s += red[((size_t)sp * k.owners + o) * PER + j];
        const int ocg = o % k.ncg, oci = ci0 + o / k.ncg;
        if (j < TCO * NT) {
            const int c = j / NT, t = j - c * NT;
            out[((size_t)oci * NT + t) * a.Cout + ocg * TCO + c] = s;      // packed [ci][ky][kx][co]
        } else if (oci == 0) {
            out[(size_t)k.cin_eff * NT * a.Cout + ocg * TCO + (j - TCO * NT)] = s;   // db[co]
        }
    }
}

int plan(const WgradArgs& a, WgK& k) {
    k.a = a;
    const int KS = a.ks, S = a.stride;
    k.cin_eff = a.Cin + (a.ones_channel ? 1 : 0);
    k.tco = (a.Cout % 4 == 0) ? 4 : ((a.Cout % 2 == 0) ? 2 : 0);
    if (!k.tco) return -1;
    k.ncg = a.Cout / k.tco;
    k.slices = 1;
    while (k.ncg * ((k.cin_eff + k.slices - 1) / k.slices) > 256) ++k.slices;
    while (k.cin_eff % k.slices) ++k.slices;
    k.ci_per = k.cin_eff / k.slices;
    k.owners = k.ncg * k.ci_per;
    if (k.owners > 256) return -2;
    k.nsplit = 256 / k.owners;
    k.threads = ((k.owners * k.nsplit + 31) / 32) * 32;
    if (k.threads > 256) k.threads = 256;
    if (k.owners * k.nsplit > k.threads) k.nsplit = k.threads / k.owners;
    const int PADL = (KS == 1) ? 0 : 4;
    k.RSW = a.W + ((KS == 1) ? 0 : 8);
    // band height: largest divisor of Ho whose tiles fit ~64 KB
    k.TRO = 1;
    for (int tr = 1; tr <= a.Ho; ++tr) {
        if (a.Ho % tr) continue;
        const int rin = (tr - 1) * S + KS;
        int cs = rin * k.RSW; cs += (36 - cs % 32) % 32;
        int gsd = tr * a.Wo; gsd += (36 - gsd % 32) % 32;
        if ((size_t)(k.ci_per * cs + a.Cout * gsd) * 4 <= 64 * 1024) k.TRO = tr;
    }
    (void)PADL;
    k.nbands = a.Ho / k.TRO;
    k.RIN = (k.TRO - 1) * S + KS;
    k.ci_stride = k.RIN * k.RSW; k.ci_stride += (36 - k.ci_stride % 32) % 32;   // == 4 (mod 32): channels hit distinct banks
    k.g_stride = k.TRO * a.Wo; k.g_stride += (36 - k.g_stride % 32) % 32;
    k.x_floats = k.ci_per * k.ci_stride;
    k.g_floats = a.Cout * k.g_stride;
    k.out_floats = k.cin_eff * KS * KS * a.Cout + a.Cout;
    k.SPW = a.Wo / 4;
    k.work = (int64_t)a.B * k.nbands;
    return 0;
}

template <int KS, int STRIDE, int TCO>
int launch(const WgK& k, int ncta, cudaStream_t st) {
    const int PER = TCO * KS * KS + TCO;
    size_t smem = (size_t)(k.x_floats + k.g_floats) * 4;
    const size_t red = (size_t)k.nsplit * k.owners * PER * 4;
    if (red > smem) smem = red;
    auto kern = wgrad_kernel<KS, STRIDE, TCO>;
    if (smem > 48 * 1024) DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMB_LAUNCH((kern), dim3(ncta, k.slices), k.threads, smem, st, k);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

// sum partials over CTAs and scatter to the torch layout.  Block = 32 consecutive elements x 32 CTA groups
// (coalesced 128-byte rows, <= 10 independent loads per thread); groups are folded in fixed order.
__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(const float* __restrict__ partials, int ncta, int out_floats,
                                                            int cin_eff, int cout, int ks, float* __restrict__ dw,
                                                            float* __restrict__ db, float* __restrict__ packed_out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (e < out_floats) {
        const int per = (ncta + 31) / 32;
        const int c0 = grp * per, c1 = min(ncta, c0 + per);
        for (int c = c0; c < c1; ++c) s += partials[(size_t)c * out_floats + e];
    }
    red[grp][lane] = s;
    __syncthreads();
    if (grp != 0 || e >= out_floats) return;
    s = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) s += red[g][lane];
    if (packed_out) { packed_out[e] = s; return; }
    const int nt = ks * ks;
    const int nw = cin_eff * nt * cout;
    if (e >= nw) { if (db) db[e - nw] = s; return; }
    const int co = e % cout;
    const int t = e / cout;
    const int tap = t % nt;
    const int ci = t / nt;
    dw[((size_t)co * cin_eff + ci) * nt + tap] = s;      // (Cout, Cin, kh, kw) of the wgrad problem
}

// the same for every queued layer in ONE launch: a block finds its job from its index (jobs hold their first block)
__global__ void __launch_bounds__(1024) wgrad_reduce_multi_kernel(const WgReduceQueue q) {
    pdl_wait();
    __shared__ float red[32][33];
    int ji = 0;
    while (ji + 1 < q.n && (int)blockIdx.x >= q.j[ji + 1].block0) ++ji;
    const WgReduceJob& jb = q.j[ji];
    const float* __restrict__ partials = jb.partials;
    const int ncta = jb.ncta, out_floats = jb.out_floats;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int e = ((int)blockIdx.x - jb.block0) * 32 + lane;
    float s = 0.f;
    if (e < out_floats) {
        const int per = (ncta + 31) / 32;
        const int c0 = grp * per, c1 = min(ncta, c0 + per);
        for (int c = c0; c < c1; ++c) s += partials[(size_t)c * out_floats + e];
    }
    red[grp][lane] = s;
    __syncthreads();
    if (grp != 0 || e >= out_floats) return;
    s = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) s += red[g][lane];
    if (jb.packed_out) { jb.packed_out[e] = s; return; }
    const int nt = jb.ks * jb.ks;
    const int nw = jb.cin_eff * nt * jb.cout;
    if (e >= nw) { if (jb.db) jb.db[e - nw] = s; return; }
    const int co = e % jb.cout;
    const int t = e / jb.cout;
    const int tap = t % nt;
    const int ci = t / nt;
    jb.dw[((size_t)co * jb.cin_eff + ci) * nt + tap] = s;
}

// composite head: chain rule from the effective (ni+1)-channel 4x4 conv to enc.0 / enc.1 parameters
__global__ void composite_chain_kernel(const float* __restrict__ dweff, const float* __restrict__ w0,
                                       const float* __restrict__ b0, const float* __restrict__ w1, int ni, int cm,
                                       float* __restrict__ dw0, float* __restrict__ db0, float* __restrict__ dw1,
                                       float* __restrict__ db1) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // dweff: packed [ni+1][16][cm] then db[cm];  w0 (cm,ni)  b0 (cm)  w1 (cm,cm,4,4)
    const int n_w1 = cm * cm * 16, n_w0 = cm * ni;
    const int total = n_w1 + n_w0 + cm + cm;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        if (e < n_w1) {
            const int tap = e % 16;
            int t = e / 16;
            const int mch = t % cm, co = t / cm;
            double s = (double)b0[mch] * (double)dweff[((size_t)ni * 16 + tap) * cm + co];
            for (int c = 0; c < ni; ++c) s += (double)w0[mch * ni + c] * (double)dweff[((size_t)c * 16 + tap) * cm + co];
            dw1[e] = (float)s;
        } else if (e < n_w1 + n_w0) {
            const int j = e - n_w1;
            const int c = j % ni, mch = j / ni;
            double s = 0.0;
            for (int co = 0; co < cm; ++co)
                for (int tap = 0; tap < 16; ++tap)
                    s += (double)w1[((size_t)co * cm + mch) * 16 + tap] * (double)dweff[((size_t)c * 16 + tap) * cm + co];
            dw0[j] = (float)s;
        } else if (e < n_w1 + n_w0 + cm) {
            const int mch = e - n_w1 - n_w0;
            double s = 0.0;
            for (int co = 0; co < cm; ++co)
                for (int tap = 0; tap < 16; ++tap)
                    s += (double)w1[((size_t)co * cm + mch) * 16 + tap] * (double)dweff[((size_t)ni * 16 + tap) * cm + co];
            db0[mch] = (float)s;
        } else {
            const int co = e - n_w1 - n_w0 - cm;
            db1[co] = dweff[(size_t)(ni + 1) * 16 * cm + co];
        }
    }
}

}  // namespace

int wgrad_partial_floats(const WgradArgs& a, int* ncta) {
    {
        int n = 0, of = 0;
        if (wgrad_tma_plan(a, &n, &of) == 0) { if (ncta) *ncta = n; return of; }
    }
    WgK k;
    if (plan(a, k)) return -1;
    int n = 148 * 2;
    if (k.work < n) n = (int)k.work;
    if (ncta) *ncta = n;
    return k.out_floats;
}

// dw: torch-layout weight gradient (Cout_w, Cin_w, k, k) of the *wgrad problem* (for a ConvTranspose2d the
// caller swaps the tensors so that this IS the (Cin_T, Cout_T, k, k) layout); db may be nullptr.
namespace {
// fold now, or append to the queue
int reduce_or_queue(const float* partials, int ncta, int of, int cin_eff, int cout, int ks, float* dw, float* db,
                    float* packed_out, cudaStream_t st, WgReduceQueue* q) {
    if (q) {
        DMB_CHECK(q->n < WG_REDUCE_MAX, "wgrad: reduce queue full");
        WgReduceJob& j = q->j[q->n++];
        j = WgReduceJob{partials, ncta, of, cin_eff, cout, ks, dw, db, packed_out, q->blocks};
        q->blocks += (of + 31) / 32;
        return 0;
    }
    DMB_LAUNCH((wgrad_reduce_kernel), (of + 31) / 32, 1024, 0, st, partials, ncta, of, cin_eff, cout, ks, dw, db, packed_out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}
}  // namespace

int wgrad_reduce_flush(WgReduceQueue& q, cudaStream_t st) {
    if (q.n == 0) return 0;
    DMB_LAUNCH((wgrad_reduce_multi_kernel), q.blocks, 1024, 0, st, q);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    q.n = 0; q.blocks = 0;
    return 0;
}

int wgrad(const WgradArgs& a, float* dw, float* db, float* packed_out, cudaStream_t st, WgReduceQueue* q) {
    DMB_CHECK((a.ks == 1 && a.stride == 1) || (a.ks == 3 && a.stride == 1) || (a.ks == 4 && a.stride == 2),
              "wgrad: unsupported kernel %d stride %d", a.ks, a.stride);
    DMB_CHECK(a.Wo % 4 == 0 && a.W % 4 == 0, "wgrad: widths must be multiples of 4");
    {
        int n = 0, of = 0;
        const int rc = wgrad_tma(a, &n, &of, st);
        if (rc < 0) return rc;
        if (rc == 0) {
            const int cin_eff = a.Cin + (a.ones_channel ? 1 : 0);
            return reduce_or_queue(a.partials, n, of, cin_eff, a.Cout, a.ks, dw, db, packed_out, st, q);
        }
    }
    WgK k;
    const int rc = plan(a, k);
    DMB_CHECK(rc == 0, "wgrad: no plan (Cout=%d Cin=%d): %d", a.Cout, a.Cin, rc);
    int ncta = 148 * 2;
    if (k.work < ncta) ncta = (int)k.work;
    if (k.tco == 4) {
        if (a.ks == 1) DMB_TRY((launch<1, 1, 4>(k, ncta, st)));
        else if (a.ks == 3) DMB_TRY((launch<3, 1, 4>(k, ncta, st)));
        else DMB_TRY((launch<4, 2, 4>(k, ncta, st)));
    } else {
        if (a.ks == 1) DMB_TRY((launch<1, 1, 2>(k, ncta, st)));
        else if (a.ks == 3) DMB_TRY((launch<3, 1, 2>(k, ncta, st)));
        else DMB_TRY((launch<4, 2, 2>(k, ncta, st)));
    }
    return reduce_or_queue(a.partials, ncta, k.out_floats, k.cin_eff, a.Cout, a.ks, dw, db, packed_out, st, q);
}

int composite_chain(const float* dweff, const float* w0, const float* b0, const float* w1, int ni, int cm,
                    float* dw0, float* db0, float* dw1, float* db1, cudaStream_t st) {
    const int total = cm * cm * 16 + cm * ni + 2 * cm;
    DMB_LAUNCH((composite_chain_kernel), (total + 127) / 128, 128, 0, st, dweff, w0, b0, w1, ni, cm, dw0, db0, dw1, db1);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace dmb
