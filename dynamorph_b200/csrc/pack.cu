// Weight packing: torch layouts -> kernel layout, in one launch.
//   Conv2d          (Cout,Cin,kh,kw) -> [Cin][kh][kw][Cout]
//   ConvTranspose2d (Cin,Cout,kh,kw) -> [Cin][kh][kw][Cout]
//   z16 head: Conv2d(ni->h/2, 1x1) followed by Conv2d(h/2->h/2, 4x4, s2, p1)
//             (vq_vae.py:277-278, no nonlinearity between) is composed into one 4x4 conv over the
//             ni input channels: W[c][ky][kx][co] = sum_m W1[co][m][ky][kx] * W0[m][c].  The 1x1's
//             bias reaches an output only through taps that fall inside the image (zero padding
//             is applied AFTER the 1x1), hence 3x3 border classes of the effective bias.
//   DMB_BN_EVAL: BatchNorm folded in (w *= g/sqrt(rv+eps); b = (b-rm)*g/sqrt(rv+eps) + beta).
// All arithmetic in double, rounded once.
#include "model.h"

namespace dmb {
namespace {

constexpr int MAX_PACK = 32;
struct PackL {
    int cin, cout, ks, transposed, composite, cmid, bias_classes, fold;
    int64_t w_off, b_off, w0_off, b0_off, pw_off, pb_off;
    int64_t g_off, beta_off, rm_off, rv_off;
    int64_t start;     // first packed-element id of this layer (weights, bias, data-gradient weights)
    int64_t nw, nb, nd, pdw_off;
    int stride;
};
struct PackPlan { int n; int64_t total; float eps; PackL l[MAX_PACK]; int nbn; int64_t bn_start; int64_t zero_off; int nzero;
                  int64_t bg_off[MAX_PACK], bb_off[MAX_PACK], bpg_off[MAX_PACK], bpb_off[MAX_PACK]; int bc[MAX_PACK]; };

__device__ double bn_scale(const PackL& l, const float* params, const float* bnbuf, int co, float eps, double* shift_mean) {
    if (!l.fold) { *shift_mean = 0.0; return 1.0; }
    const double s = (double)params[l.g_off + co] / sqrt((double)bnbuf[l.rv_off + co] + (double)eps);
    *shift_mean = (double)bnbuf[l.rm_off + co];
    return s;
}

__global__ void pack_kernel(const PackPlan p, const float* __restrict__ params,
                            const float* __restrict__ bnbuf, float* __restrict__ packed) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // (32-bit index arithmetic: the plan holds a few hundred thousand elements at most -- checked by the host -- and
    // 64-bit div / mod by run-time values were most of this kernel's 12 us at the start of every training step)
    const int total = (int)p.total;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
        if (id >= p.bn_start) {      // gamma / beta copies, then the zero vector
            int r = id - (int)p.bn_start;
            int nbn_el = 0;
            for (int i = 0; i < p.nbn; ++i) nbn_el += 2 * p.bc[i];
            if (r >= nbn_el) { packed[p.zero_off + (r - nbn_el)] = 0.f; continue; }
            for (int i = 0; i < p.nbn; ++i) {
                if (r < 2 * p.bc[i]) {
                    if (r < p.bc[i]) packed[p.bpg_off[i] + r] = params[p.bg_off[i] + r];
                    else packed[p.bpb_off[i] + r - p.bc[i]] = params[p.bb_off[i] + r - p.bc[i]];
                    break;
                }
                r -= 2 * p.bc[i];
            }
            continue;
        }
        int li = 0;
        while (li + 1 < p.n && id >= p.l[li + 1].start) ++li;
        const PackL& l = p.l[li];
        const int e = id - (int)l.start;
        const int K = l.ks;
        if (e < (int)l.nw) {
            const int co = e % l.cout;
            int t = e / l.cout;
            const int kx = t % K; t /= K;
            const int ky = t % K;
            const int ci = t / K;
            double mean;
            const double s = bn_scale(l, params, bnbuf, co, p.eps, &mean);
            double w;
            if (l.composite) {
                w = 0.0;
                for (int mch = 0; mch < l.cmid; ++mch)
                    w += (double)params[l.w_off + ((int64_t)(co * l.cmid + mch) * K + ky) * K + kx] *
                         (double)params[l.w0_off + (int64_t)mch * l.cin + ci];
            } else if (l.transposed) {
                w = params[l.w_off + ((int64_t)(ci * l.cout + co) * K + ky) * K + kx];
            } else {
                w = params[l.w_off + ((int64_t)(co * l.cin + ci) * K + ky) * K + kx];
            }
            packed[l.pw_off + e] = (float)(w * s);
        } else if (e >= (int)(l.nw + l.nb)) {
            // data-gradient weights [cout][k][k][cin]: conv stride 1 flips the taps (correlation ->
            // convolution); stride-2 conv / ConvTranspose2d keep them (they swap kernels instead)
            const int ed = e - (int)l.nw - (int)l.nb;
            const int ci = ed % l.cin;
            int t = ed / l.cin;
            int kx = t % K; t /= K;
            int ky = t % K;
            const int co = t / K;
            if (!l.transposed && l.stride == 1) { ky = K - 1 - ky; kx = K - 1 - kx; }
            const float w = l.transposed ? params[l.w_off + ((int64_t)(ci * l.cout + co) * K + ky) * K + kx]
                                         : params[l.w_off + ((int64_t)(co * l.cin + ci) * K + ky) * K + kx];
            packed[l.pdw_off + ed] = w;
        } else {
            const int eb = e - (int)l.nw;
            const int co = eb % l.cout;
            const int cls = eb / l.cout;             // 0 unless bias_classes
            double mean;
            const double s = bn_scale(l, params, bnbuf, co, p.eps, &mean);
            double b = params[l.b_off + co];
            if (l.composite) {
                const int rc = cls / 3, cc = cls % 3;
                for (int ky = 0; ky < K; ++ky) {
                    if ((rc == 0 && ky == 0) || (rc == 2 && ky == K - 1)) continue;
                    for (int kx = 0; kx < K; ++kx) {
                        if ((cc == 0 && kx == 0) || (cc == 2 && kx == K - 1)) continue;
                        for (int mch = 0; mch < l.cmid; ++mch)
                            b += (double)params[l.w_off + ((int64_t)(co * l.cmid + mch) * K + ky) * K + kx] *
                                 (double)params[l.b0_off + mch];
                    }
                }
            }
            if (l.fold) b = (b - mean) * s + (double)params[l.beta_off + co];
            packed[l.pb_off + eb] = (float)b;
        }
    }
}

}  // namespace

int pack_weights(const Layout& L, const float* params, const float* bnbuf, int bn_mode, float* packed, cudaStream_t st) {
    DMB_CHECK((int)L.convs.size() <= MAX_PACK && (int)L.bns.size() <= MAX_PACK, "pack: too many layers");
    PackPlan p{};
    p.n = (int)L.convs.size();
    p.eps = L.m.bn_eps;
    int64_t start = 0;
    for (int i = 0; i < p.n; ++i) {
        const ConvL& c = L.convs[i];
        PackL& l = p.l[i];
        l.cin = c.cin; l.cout = c.cout; l.ks = c.ks; l.transposed = c.transposed; l.composite = c.composite;
        l.cmid = c.cmid; l.bias_classes = c.bias_classes;
        l.fold = (bn_mode == DMB_BN_EVAL && c.bn >= 0) ? 1 : 0;
        l.w_off = c.w_off; l.b_off = c.b_off; l.w0_off = c.w0_off; l.b0_off = c.b0_off;
        l.pw_off = c.pw_off; l.pb_off = c.pb_off;
        if (c.bn >= 0) {
            const BnL& b = L.bns[c.bn];
            l.g_off = b.g_off; l.beta_off = b.b_off; l.rm_off = b.rm_off; l.rv_off = b.rv_off;
        }
        l.start = start;
        l.nw = (int64_t)c.cin * c.ks * c.ks * c.cout;
        l.nb = (int64_t)(c.bias_classes ? 9 : 1) * c.cout;
        l.nd = (c.pdw_off >= 0) ? l.nw : 0;
        l.pdw_off = c.pdw_off;
        l.stride = c.stride;
        start += l.nw + l.nb + l.nd;
    }
    p.bn_start = start;
    p.nbn = (int)L.bns.size();
    for (int i = 0; i < p.nbn; ++i) {
        p.bg_off[i] = L.bns[i].g_off; p.bb_off[i] = L.bns[i].b_off;
        p.bpg_off[i] = L.bns[i].pg_off; p.bpb_off[i] = L.bns[i].pb_off; p.bc[i] = L.bns[i].c;
        start += 2 * L.bns[i].c;
    }
    p.zero_off = L.pzero_off;
    p.nzero = L.max_c;
    start += L.max_c;
    p.total = start;
    DMB_CHECK(p.total < (1ll << 30), "pack: plan too large for 32-bit indexing");
    const int threads = 256;
    int blocks = (int)((p.total + threads - 1) / threads);
    if (blocks > 148 * 8) blocks = 148 * 8;
    DMB_LAUNCH((pack_kernel), blocks, threads, 0, st, p, params, bnbuf, packed);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    if (bn_mode == DMB_BN_EVAL) {                // Winograd-domain tensor-core tiles for conv_wino_tc.cu
        for (const ConvL& c : L.convs)
            if (c.pwn_off >= 0) DMB_TRY(pack_wino_weights(packed + c.pw_off, packed + c.pwn_off, c.cin, c.cout, st));
    }
    {                                            // [b_hi; b_lo] operand images for conv_tm.cu (every mode: folded or raw)
        std::vector<TmPackJob> jobs;
        for (const ConvL& c : L.convs)
            if (c.ptm_off >= 0) {
                jobs.push_back(TmPackJob{packed + c.pw_off, packed + c.ptm_off, c.cin, c.cout, c.ks});
                if (c.ptm_tail >= 0) {
                    const ConvL& t = L.convs[c.ptm_tail];
                    jobs.push_back(TmPackJob{packed + t.pw_off, packed + c.ptm_off + conv_tm_weight_floats(c.cin, c.cout, c.ks),
                                             t.cin, t.cout, t.ks});
                }
            }
        for (const ConvL& c : L.convs)          // ConvTranspose2d layers in the transposed form
            if (c.pctm_off >= 0) jobs.push_back(TmPackJob{packed + c.pw_off, packed + c.pctm_off, c.cin, c.cout, 4, 1});
        if (bn_mode != DMB_BN_EVAL)             // training: images of the data-gradient weights (channels swapped)
            for (const ConvL& c : L.convs)
                if (c.pdtm_off >= 0) {
                    if (!c.transposed && c.stride == 2)      // data gradient = transposed convolution cout -> cin
                        jobs.push_back(TmPackJob{packed + c.pdw_off, packed + c.pdtm_off, c.cout, c.cin, 4, 1});
                    else
                        jobs.push_back(TmPackJob{packed + c.pdw_off, packed + c.pdtm_off, c.cout, c.cin, c.ks, 0});
                }
        if (!jobs.empty()) DMB_TRY(pack_tm_weights_multi(jobs.data(), (int)jobs.size(), st));
    }
    if (L.tc && bn_mode == DMB_BN_EVAL) {        // split, swizzled tiles of the folded weights for conv_tc.cu
        for (const ConvL& c : L.convs)
            if (c.ptc_off >= 0) DMB_TRY(pack_tc_weights(packed + c.pw_off, packed + c.ptc_off, c.cin, c.cout, c.ks, st));
    }
    return 0;
}

}  // namespace dmb
