// Host-side description of the two reference architectures: where every state_dict tensor
// sits in the flat `params` / `bnbuf` buffers, where its kernel-layout copy sits in `packed`,
// and how a batch's scratch is carved out of the caller's workspace.
#pragma once
#include "common.cuh"

namespace dmb {

struct BnL {
    int c;
    int64_t g_off, b_off;        // gamma, beta in params
    int64_t rm_off, rv_off;      // running stats in bnbuf
    int64_t pg_off, pb_off;      // copies of gamma, beta in packed
};

struct ConvL {
    int cin, cout, ks, stride;
    int transposed;              // ConvTranspose2d: torch weight is (Cin, Cout, kh, kw)
    int composite;               // z16 enc.0 (1x1) folded into enc.1 (4x4): cin = num_inputs
    int cmid;                    // composite: channels between the two convs
    int64_t w_off, b_off;        // params
    int64_t w0_off, b0_off;      // composite: the 1x1 conv's weight / bias in params
    int bn;                      // BatchNorm that follows (index into bns) or -1
    int64_t pw_off, pb_off;      // packed weight [cin][ks][ks][cout], bias [cout] or [9][cout]
    int64_t pdw_off;             // packed weight of the data-gradient conv [cout][ks][ks][cin] (-1: none)
    int bias_classes;
    int64_t ptc_off;             // tensor-core tiles of the EVAL-folded weights (conv_tc.cu) or -1
    int64_t pwn_off;             // Winograd-domain tensor-core tiles of the EVAL-folded weights (conv_wino_tc.cu) or -1
    int64_t ptm_off;             // [b_hi; b_lo] operand image of the EVAL-folded weights for conv_tm.cu or -1
    int ptm_tail;                // conv whose image follows this one's (the 1x1 of a fused residual layer) or -1
    int64_t pdtm_off;            // conv_tm.cu image of the data-gradient weights (training step) or -1
    int64_t pctm_off;            // ConvTranspose2d: conv_tm.cu image of the transposed form (3x3 on the input grid) or -1
};

struct Entry { std::string key; int which; int64_t off, numel; };

struct ResL { int a, b; };       // conv3x3 (h->rh), conv1x1 (rh->h)

struct Layout {
    dmb_model m;
    std::vector<Entry> entries;
    std::vector<ConvL> convs;
    std::vector<BnL> bns;
    int64_t n_params = 0, n_bnbuf = 0, n_packed = 0;
    int64_t codebook_off = 0;
    int64_t pzero_off = 0;       // zeros[max channels] in packed (bias of the data-gradient convs)
    int max_c = 0;
    // encoder
    int e1 = -1, e2 = -1, e3 = -1, e4 = -1;   // z32 uses e1, e2 only
    std::vector<ResL> enc_res, dec_res;
    // decoder
    int d0 = -1, d1 = -1, d2 = -1, d3 = -1;   // z16: three ConvT + conv1x1; z32: d0, d1 ConvT
    int D = 0, lh = 0, lw = 0;
    // every encoder layer after the head is a tensor-core shape (conv_tc_supported): the EVAL-mode encoder runs
    // them on tcgen05 over NHWC activations (64-wide configurations, BASELINE configs[3])
    bool tc = false;
};

int build_layout(const dmb_model* m, Layout& L);

// Every scratch buffer a forward (and the backward that follows) touches.
struct BnWs { double* part; float* scale; float* shift; float* mean; float* invstd; int nbands; int64_t count;
              double* gsum; };   // gsum: [C][2] folded sums that cross the ranks under synchronised BatchNorm
struct Workspace {
    // z16 encoder activations (raw conv outputs in BATCH/PER_SAMPLE mode, post-activation in EVAL)
    float *y1 = nullptr, *y2 = nullptr, *y3 = nullptr, *y4 = nullptr;
    float* y1t = nullptr;                  // NHWC copy of y1 feeding the tensor-core layers (Layout::tc, EVAL)
    std::vector<float*> era, erb, ehs;     // encoder residual: 3x3 out, 1x1 out, running sum
    std::vector<float*> dra, drb, dhs;     // decoder residual (z32)
    float *zb = nullptr, *za = nullptr;
    int32_t* idx = nullptr;
    float *t1 = nullptr, *t2 = nullptr, *t3 = nullptr;   // decoder activations
    float* dec = nullptr;
    std::vector<BnWs> bn;                  // one per BatchNorm, same order as Layout::bns
    // ---- backward (keep != 0)
    struct BnB { double* part; float *A, *Bc, *Cc; double* gsum; };
    std::vector<BnB> bnb;
    float *gd = nullptr, *g_t3 = nullptr, *g_t2 = nullptr, *g_t1 = nullptr, *g_za = nullptr, *g_zb = nullptr;
    std::vector<float*> g_era, g_eh;       // encoder residual: grad at BN_a output (masked), grad at layer input
    std::vector<float*> g_dra, g_dh;       // decoder residual (z32)
    float *g_y3 = nullptr, *g_y2 = nullptr, *g_y1 = nullptr;
    double* bias_part = nullptr;           // stats partials for bias grads of BN-less layers
    float* wg_part = nullptr;              // wgrad per-CTA partials
    float* dweff = nullptr;                // composite head: packed effective-weight gradient
    float* vq_part = nullptr;              // codebook-gradient partials [vq_part_rows][K][D]
    int vq_part_rows = 0;
    size_t wg_part_floats = 0;
    bool wg_queue = false;                 // wg_part holds one region per layer: the folds are queued and launched once
    double* vq_stats = nullptr;            // [2+K]
    double* recon_sum = nullptr;           // [1]
    float* scalars = nullptr;              // [8] vq loss, perplexity, ..., [4] time-matching loss
    float* tm_scratch = nullptr;           // time-matching pair sums + dloss/dsim (keep != 0)
    float* g_tm = nullptr;                 // time-matching gradient at the latent (B, D, lh, lw)
    float* g_tmp = nullptr;                // materialised BatchNorm-backward gradient feeding a TMA data-gradient conv
    size_t bytes = 0;
};
int carve_workspace(const Layout& L, int64_t B, int bn_mode, int keep, void* base, Workspace& w);

int pack_weights(const Layout& L, const float* params, const float* bnbuf, int bn_mode, float* packed, cudaStream_t st);

}  // namespace dmb
