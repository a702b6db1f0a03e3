// Model-level host code: parameter layout, workspace carving, and the layer schedule of the
// encoder / quantiser / decoder for both reference architectures, behind the C ABI.
//   z16: HiddenStateExtractor/vq_vae.py:276-298 (== vae.py:273-294)
//   z32: HiddenStateExtractor/vae.py:401-414
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "model.h"

namespace dmb {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("DMB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
// Tensor-memory-operand kernels (conv_tm.cu): smallest batch that takes them (persistent CTAs; tiles = batch x 2..8).
// DMB_TM=0 switches them off, DMB_TM_MIN_B overrides the threshold.
constexpr bool TM_DEFAULT_ON = true;
int64_t tm_min_batch() {
    const char* sw = getenv("DMB_TM");
    const bool on = sw ? (sw[0] != '0') : TM_DEFAULT_ON;
    if (!on) return INT64_MAX;
    const char* e = getenv("DMB_TM_MIN_B");
    return e ? atoll(e) : 256;
}
// ... and their train-mode-BatchNorm form: per-patch statistics (bulk encoding with the as-written process_VAE semantics)
// and whole-batch statistics (the training step, from the same batch size: 256 patches are 512 .. 8192 tiles for the 296
// persistent CTAs).  DMB_TM_BN=0 switches both off, DMB_TM_BN_BATCH=0 the training form only.
bool tm_bn_mode(int bn_mode, int64_t B) {
    const char* e = getenv("DMB_TM_BN");
    if (e && e[0] == '0') return false;
    if (bn_mode == DMB_BN_BATCH) {
        const char* eb = getenv("DMB_TM_BN_BATCH");
        if (eb && eb[0] == '0') return false;
    }
    return (bn_mode == DMB_BN_PER_SAMPLE || bn_mode == DMB_BN_BATCH) && B >= tm_min_batch();
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------------------
namespace {

struct Builder {
    Layout& L;
    int64_t p = 0, bb = 0, pk = 0;
    explicit Builder(Layout& l) : L(l) {}
    int64_t take_param(const std::string& key, int64_t n) {
        L.entries.push_back({key, 0, p, n});
        const int64_t o = p; p += n; return o;
    }
    int64_t take_buf(const std::string& key, int64_t n) {
        L.entries.push_back({key, 1, bb, n});
        const int64_t o = bb; bb += n; return o;
    }
    int64_t take_packed(int64_t n) { const int64_t o = pk; pk += (n + 3) & ~3ll; return o; }

    int conv(const std::string& key, int cin, int cout, int ks, int stride, bool transposed) {
        ConvL c{};
        c.cin = cin; c.cout = cout; c.ks = ks; c.stride = stride; c.transposed = transposed; c.bn = -1;
        c.ptc_off = -1; c.pwn_off = -1; c.ptm_off = -1; c.ptm_tail = -1; c.pdtm_off = -1; c.pctm_off = -1;
        c.w_off = take_param(key + ".weight", (int64_t)cin * cout * ks * ks);
        c.b_off = take_param(key + ".bias", cout);
        c.pw_off = take_packed((int64_t)cin * cout * ks * ks);
        c.pb_off = take_packed(cout);
        c.pdw_off = take_packed((int64_t)cin * cout * ks * ks);
        if (cin > L.max_c) L.max_c = cin;
        if (cout > L.max_c) L.max_c = cout;
        L.convs.push_back(c);
        return (int)L.convs.size() - 1;
    }
    // z16 head: conv1x1 (key0) + conv4x4 s2 (key1) -> one composite conv over `cin` channels
    int composite(const std::string& key0, const std::string& key1, int cin, int cmid) {
        ConvL c{};
        c.cin = cin; c.cout = cmid; c.cmid = cmid; c.ks = 4; c.stride = 2; c.composite = 1; c.bn = -1;
        c.bias_classes = 1;
        c.ptc_off = -1; c.pwn_off = -1; c.ptm_off = -1; c.ptm_tail = -1; c.pdtm_off = -1; c.pctm_off = -1;
        c.w0_off = take_param(key0 + ".weight", (int64_t)cmid * cin);
        c.b0_off = take_param(key0 + ".bias", cmid);
        c.w_off = take_param(key1 + ".weight", (int64_t)cmid * cmid * 16);
        c.b_off = take_param(key1 + ".bias", cmid);
        c.pw_off = take_packed((int64_t)cin * 16 * cmid);
        c.pb_off = take_packed(9 * cmid);
        c.pdw_off = -1;
        L.convs.push_back(c);
        return (int)L.convs.size() - 1;
    }
    void bn(const std::string& key, int c, int conv_idx) {
        BnL b{};
        b.c = c;
        b.g_off = take_param(key + ".weight", c);
        b.b_off = take_param(key + ".bias", c);
        b.rm_off = take_buf(key + ".running_mean", c);
        b.rv_off = take_buf(key + ".running_var", c);
        b.pg_off = take_packed(c);
        b.pb_off = take_packed(c);
        L.bns.push_back(b);
        L.convs[conv_idx].bn = (int)L.bns.size() - 1;
    }
    void res(const std::string& prefix, int h, int rh, int n, std::vector<ResL>& out) {
        for (int i = 0; i < n; ++i) {
            const std::string p = prefix + ".layers." + std::to_string(i);
            ResL r;
            r.a = conv(p + ".1", h, rh, 3, 1, false);
            bn(p + ".2", rh, r.a);
            r.b = conv(p + ".4", rh, h, 1, 1, false);
            bn(p + ".5", h, r.b);
            out.push_back(r);
        }
    }
};

}  // namespace

int build_layout(const dmb_model* m, Layout& L) {
    DMB_CHECK(m != nullptr, "null model descriptor");
    DMB_CHECK(m->arch == DMB_ARCH_Z16 || m->arch == DMB_ARCH_Z32, "unknown arch %d", m->arch);
    DMB_CHECK(m->num_inputs >= 1 && m->num_hiddens >= 8 && m->num_hiddens % 8 == 0,
              "num_hiddens=%d must be a positive multiple of 8", m->num_hiddens);
    DMB_CHECK(m->num_residual_hiddens >= 8 && m->num_residual_hiddens % 8 == 0,
              "num_residual_hiddens=%d must be a positive multiple of 8", m->num_residual_hiddens);
    DMB_CHECK(m->num_residual_layers >= 0 && m->num_residual_layers <= 6, "num_residual_layers out of range");
    DMB_CHECK(m->num_embeddings >= 1 && m->num_embeddings <= 1024, "num_embeddings=%d outside [1,1024]", m->num_embeddings);
    const int down = (m->arch == DMB_ARCH_Z16) ? 8 : 4;
    DMB_CHECK(m->height > 0 && m->width > 0 && m->height % down == 0 && m->width % (down * 8) == 0,
              "patch %dx%d: height must be a multiple of %d and width of %d", m->height, m->width, down, down * 8);
    DMB_CHECK(m->num_inputs % 2 == 0 || m->arch == DMB_ARCH_Z16 || true, "unused");
    L = Layout();
    L.m = *m;
    Builder B(L);
    const int ni = m->num_inputs, h = m->num_hiddens, h2 = h / 2, h4 = h / 4, rh = m->num_residual_hiddens;
    const int nl = m->num_residual_layers;
    if (m->arch == DMB_ARCH_Z16) {
        L.e1 = B.composite("enc.0", "enc.1", ni, h2); B.bn("enc.2", h2, L.e1);
        L.e2 = B.conv("enc.4", h2, h, 4, 2, false);   B.bn("enc.5", h, L.e2);
        L.e3 = B.conv("enc.7", h, h, 4, 2, false);    B.bn("enc.8", h, L.e3);
        L.e4 = B.conv("enc.10", h, h, 3, 1, false);   B.bn("enc.11", h, L.e4);
        B.res("enc.12", h, rh, nl, L.enc_res);
        L.codebook_off = B.take_param("vq.w.weight", (int64_t)m->num_embeddings * h);
        L.d0 = B.conv("dec.0", h, h2, 4, 2, true);
        L.d1 = B.conv("dec.2", h2, h4, 4, 2, true);
        L.d2 = B.conv("dec.4", h4, h4, 4, 2, true);
        L.d3 = B.conv("dec.6", h4, ni, 1, 1, false);
        L.lh = m->height / 8; L.lw = m->width / 8;
    } else {
        L.e1 = B.conv("enc.0", ni, h2, 4, 2, false);  B.bn("enc.1", h2, L.e1);
        L.e2 = B.conv("enc.3", h2, h, 4, 2, false);   B.bn("enc.4", h, L.e2);
        B.res("enc.5", h, rh, nl, L.enc_res);
        L.codebook_off = B.take_param("vq.w.weight", (int64_t)m->num_embeddings * h);
        B.res("dec.0", h, rh, nl, L.dec_res);
        L.d0 = B.conv("dec.1", h, h2, 4, 2, true);    B.bn("dec.2", h2, L.d0);
        L.d1 = B.conv("dec.4", h2, ni, 4, 2, true);
        L.lh = m->height / 4; L.lw = m->width / 4;
    }
    L.D = h;
    {   // tensor-core plan: all encoder layers behind the head must be tcgen05 shapes
        std::vector<std::pair<int, int>> wide;      // (conv index, input height == width scale)
        const int H = m->height, W = m->width;
        bool ok = true;
        auto want = [&](int ci, int hh, int ww) {
            const ConvL& c = L.convs[ci];
            ok = ok && conv_tc_supported(c.cin, c.cout, c.ks, c.stride, hh, ww);
            wide.push_back({ci, 0});
        };
        if (m->arch == DMB_ARCH_Z16) {
            want(L.e2, H / 2, W / 2); want(L.e3, H / 4, W / 4); want(L.e4, H / 8, W / 8);
        } else {
            want(L.e2, H / 2, W / 2);
        }
        for (const ResL& r : L.enc_res) { want(r.a, L.lh, L.lw); want(r.b, L.lh, L.lw); }
        ok = ok && ((int64_t)(H / 2) * (W / 2)) % 32 == 0 && L.convs[L.e1].cout % 32 == 0;
        if (ok) {
            L.tc = true;
            for (auto& wv : wide) {
                ConvL& c = L.convs[wv.first];
                c.ptc_off = B.take_packed(conv_tc_weight_floats(c.cin, c.cout, c.ks));
            }
        }
    }
    {   // Winograd tensor-core plan: the 16-channel 3x3 layers at a 16x16 latent (default configuration, EVAL mode)
        std::vector<int> cand;
        if (m->arch == DMB_ARCH_Z16) cand.push_back(L.e4);
        for (const ResL& r : L.enc_res) cand.push_back(r.a);
        for (int ci : cand) {
            ConvL& c = L.convs[ci];
            // (the kernel also takes 16 output channels, but there the direct CUDA-core kernel is still faster)
            if (c.cout == 32 && conv_wino_supported(c.cin, c.cout, c.ks, c.stride, L.lh, L.lw))
                c.pwn_off = B.take_packed(conv_wino_weight_floats(c.cin, c.cout));
        }
    }
    {   // tensor-memory-operand plan (conv_tm.cu): the thin encoder layers behind the head, EVAL mode
        const int H = m->height, W = m->width;
        auto want = [&](int ci, int hh, int ww) {
            ConvL& c = L.convs[ci];
            if (!c.transposed && !c.composite && conv_tm_supported(c.cin, c.cout, c.ks, c.stride, hh, ww))
                c.ptm_off = B.take_packed(conv_tm_weight_floats(c.cin, c.cout, c.ks));
        };
        if (m->arch == DMB_ARCH_Z16) { want(L.e2, H / 2, W / 2); want(L.e3, H / 4, W / 4); want(L.e4, H / 8, W / 8); }
        else want(L.e2, H / 2, W / 2);
        for (const ResL& r : L.enc_res) {
            ConvL& ca = L.convs[r.a];
            const ConvL& cb = L.convs[r.b];
            const bool fuse = conv_tm_supported(ca.cin, ca.cout, ca.ks, ca.stride, L.lh, L.lw) && ca.ks == 3 &&
                              ca.cout == 32 && conv_tm_supported(cb.cin, cb.cout, cb.ks, cb.stride, L.lh, L.lw);
            if (fuse) {      // one image: the 3x3's tiles, then the 1x1's (conv_tm.cu FUSE)
                ca.ptm_off = B.take_packed(conv_tm_weight_floats(ca.cin, ca.cout, ca.ks) +
                                           conv_tm_weight_floats(cb.cin, cb.cout, cb.ks));
                ca.ptm_tail = r.b;
            } else {
                want(r.a, L.lh, L.lw);
            }
            want(r.b, L.lh, L.lw);
        }
        // ... and of the training step's data-gradient convolutions that have a tensor-memory form: stride-1 layers at the
        // latent resolution (channels swapped, taps flipped: pdw_off) and the stride-2 convolutions that back-propagate
        // through a ConvTranspose2d
        auto want_dg = [&](int ci, int hh, int ww) {      // hh x ww: the layer's INPUT map
            if (ci < 0) return;
            ConvL& c = L.convs[ci];
            if (c.pdw_off < 0 || c.composite) return;
            if (!c.transposed && c.stride == 2) {
                // the data gradient of a stride-2 convolution is a transposed convolution cout -> cin over its output map
                if (c.ks == 4 && conv_tm_ct_supported(c.cout, c.cin, hh / 2, ww / 2, true))
                    c.pdtm_off = B.take_packed(conv_tm_weight_floats(c.cout, 4 * c.cin, 3));
                return;
            }
            bool ok;
            if (c.transposed) ok = conv_tm_dg_supported(c.cout, c.cin, 4, 2, 2 * hh, 2 * ww);
            else ok = c.stride == 1 && conv_tm_dg_supported(c.cout, c.cin, c.ks, 1, hh, ww);
            if (ok) c.pdtm_off = B.take_packed(conv_tm_weight_floats(c.cout, c.cin, c.ks));
        };
        auto want_ct = [&](int ci, int hh, int ww) {      // forward of a ConvTranspose2d in the transposed form
            if (ci < 0) return;
            ConvL& c = L.convs[ci];
            if (c.transposed && c.ks == 4 && c.bn < 0 && conv_tm_ct_supported(c.cin, c.cout, hh, ww, false))
                c.pctm_off = B.take_packed(conv_tm_weight_floats(c.cin, 4 * c.cout, 3));
        };
        if (m->arch == DMB_ARCH_Z16) {
            want_ct(L.d0, L.lh, L.lw);
            want_dg(L.e2, H / 2, W / 2); want_dg(L.e3, H / 4, W / 4);
            want_dg(L.e4, H / 8, W / 8);
            want_dg(L.d0, L.lh, L.lw); want_dg(L.d1, 2 * L.lh, 2 * L.lw); want_dg(L.d2, 4 * L.lh, 4 * L.lw);
        }
        for (const ResL& r : L.enc_res) { want_dg(r.a, L.lh, L.lw); want_dg(r.b, L.lh, L.lw); }
    }
    L.pzero_off = B.take_packed(L.max_c);
    L.n_params = B.p; L.n_bnbuf = B.bb; L.n_packed = B.pk;
    return 0;
}

// ---------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------
namespace {
struct Bump {
    char* base; size_t off = 0;
    explicit Bump(void* b) : base((char*)b) {}
    template <class T> T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};
}  // namespace

int carve_workspace(const Layout& L, int64_t B, int bn_mode, int keep, void* base, Workspace& w) {
    (void)keep;
    const dmb_model& m = L.m;
    Bump bp(base);
    const int h = m.num_hiddens, h2 = h / 2, h4 = h / 4, rh = m.num_residual_hiddens;
    const int64_t H = m.height, W = m.width;
    const int64_t lat = (int64_t)L.lh * L.lw;
    w = Workspace();
    if (m.arch == DMB_ARCH_Z16) {
        w.y1 = bp.take<float>(B * h2 * (H / 2) * (W / 2));
        w.y2 = bp.take<float>(B * h * (H / 4) * (W / 4));
        w.y3 = bp.take<float>(B * h * lat);
        w.y4 = bp.take<float>(B * h * lat);
    } else {
        w.y1 = bp.take<float>(B * h2 * (H / 2) * (W / 2));
        w.y2 = bp.take<float>(B * h * lat);
    }
    if (L.tc && bn_mode == DMB_BN_EVAL) w.y1t = bp.take<float>(B * h2 * (H / 2) * (W / 2));
    for (size_t i = 0; i < L.enc_res.size(); ++i) {
        w.era.push_back(bp.take<float>(B * rh * lat));
        w.erb.push_back(bp.take<float>(B * h * lat));
        w.ehs.push_back(bp.take<float>(B * h * lat));
    }
    w.zb = bp.take<float>(B * h * lat);
    w.za = bp.take<float>(B * h * lat);
    w.idx = bp.take<int32_t>(B * lat);
    for (size_t i = 0; i < L.dec_res.size(); ++i) {
        w.dra.push_back(bp.take<float>(B * rh * lat));
        w.drb.push_back(bp.take<float>(B * h * lat));
        w.dhs.push_back(bp.take<float>(B * h * lat));
    }
    if (keep) {
        if (m.arch == DMB_ARCH_Z16) {
            w.t1 = bp.take<float>(B * h2 * 4 * lat);
            w.t2 = bp.take<float>(B * h4 * 16 * lat);
            w.t3 = bp.take<float>(B * h4 * 64 * lat);
        } else {
            w.t1 = bp.take<float>(B * h2 * 4 * lat);
        }
        w.dec = bp.take<float>(B * m.num_inputs * H * W);
    }
    // BatchNorm scratch, one per BN layer, in Layout::bns order (conv order)
    w.bn.resize(L.bns.size());
    if (bn_mode != DMB_BN_EVAL) {
        for (size_t ci = 0; ci < L.convs.size(); ++ci) {
            const ConvL& c = L.convs[ci];
            if (c.bn < 0) continue;
            int64_t Ho, Wo;
            int nb;
            const bool enc_side = (int)ci == L.e1 || (int)ci == L.e2 || (int)ci == L.e3 || (int)ci == L.e4;
            if (c.transposed) {                      // z32 dec.1 (lat -> 2x)
                Ho = 2 * L.lh; Wo = 2 * L.lw;
                nb = convt_fwd_bands(c.cin, c.cout, L.lh, L.lw);
            } else if ((int)ci == L.e1) {
                Ho = H / 2; Wo = W / 2; nb = conv_fwd_bands(c.ks, c.stride, c.cin, c.cout, (int)Ho, (int)Wo, true, B);
            } else if ((int)ci == L.e2) {
                Ho = H / 4; Wo = W / 4; nb = conv_fwd_bands(c.ks, c.stride, c.cin, c.cout, (int)Ho, (int)Wo, true, B);
            } else {
                (void)enc_side;
                Ho = L.lh; Wo = L.lw; nb = conv_fwd_bands(c.ks, c.stride, c.cin, c.cout, (int)Ho, (int)Wo, true, B);
            }
            if (c.ptm_off >= 0 && tm_bn_mode(bn_mode, B) && !c.transposed &&
                conv_tm_supported(c.cin, c.cout, c.ks, c.stride, (int)Ho * c.stride, (int)Wo * c.stride))
                nb = conv_tm_bands(c.cin, c.cout, c.ks, c.stride, (int)Ho * c.stride, (int)Wo * c.stride);
            DMB_CHECK(nb > 0, "no launch plan for conv %zu", ci);
            BnWs& b = w.bn[c.bn];
            const int64_t rows = (bn_mode == DMB_BN_PER_SAMPLE) ? B : 1;
            b.nbands = nb;
            b.count = Ho * Wo;
            // (whole-batch statistics of the tensor-memory kernels: one row per warp of the persistent grid)
            const int64_t part_rows = (B * nb > TM_BATCH_ROWS_MAX) ? B * nb : TM_BATCH_ROWS_MAX;
            b.part = bp.take<double>(part_rows * c.cout * 2);
            b.scale = bp.take<float>(rows * c.cout);
            b.shift = bp.take<float>(rows * c.cout);
            b.mean = bp.take<float>(rows * c.cout);
            b.invstd = bp.take<float>(rows * c.cout);
            b.gsum = bp.take<double>(2 * c.cout);
        }
    }
    if (keep && bn_mode != DMB_BN_EVAL) {
        // ---- backward scratch
        const int64_t rows = (bn_mode == DMB_BN_PER_SAMPLE) ? B : 1;
        w.bnb.resize(L.bns.size());
        for (size_t ci = 0; ci < L.convs.size(); ++ci) {
            const ConvL& c = L.convs[ci];
            if (c.bn < 0) continue;
            Workspace::BnB& b = w.bnb[c.bn];
            // the gradient that feeds this BN comes from a kernel whose band count is at most the map height
            const int64_t hmax = (m.arch == DMB_ARCH_Z16 && (int)ci == L.e1) ? H / 2 :
                                 (((int)ci == L.e2 && m.arch == DMB_ARCH_Z16) ? H / 4 :
                                 ((m.arch == DMB_ARCH_Z32 && ((int)ci == L.e1 || (int)ci == L.d0)) ? H / 2 : L.lh));
            const int64_t prow = (B * hmax > 2 * TM_BATCH_ROWS_MAX) ? B * hmax : 2 * TM_BATCH_ROWS_MAX;   // (transposed form: 8 rows per CTA)
            b.part = bp.take<double>(prow * c.cout * 2);
            b.A = bp.take<float>(rows * c.cout);
            b.Bc = bp.take<float>(rows * c.cout);
            b.Cc = bp.take<float>(rows * c.cout);
            b.gsum = bp.take<double>(2 * c.cout);
        }
        w.gd = bp.take<float>(B * m.num_inputs * H * W);
        if (m.arch == DMB_ARCH_Z16) {
            w.g_t3 = bp.take<float>(B * h4 * 64 * lat);
            w.g_t2 = bp.take<float>(B * h4 * 16 * lat);
            w.g_t1 = bp.take<float>(B * h2 * 4 * lat);
            w.g_y3 = bp.take<float>(B * h * lat);
            w.g_y2 = bp.take<float>(B * h * (H / 4) * (W / 4));
            w.g_y1 = bp.take<float>(B * h2 * (H / 2) * (W / 2));
        } else {
            w.g_t1 = bp.take<float>(B * h2 * 4 * lat);
            w.g_y1 = bp.take<float>(B * h2 * (H / 2) * (W / 2));
        }
        w.g_za = bp.take<float>(B * h * lat);
        w.g_zb = bp.take<float>(B * h * lat);
        for (size_t i = 0; i < L.enc_res.size(); ++i) {
            w.g_era.push_back(bp.take<float>(B * rh * lat));
            w.g_eh.push_back(bp.take<float>(B * h * lat));
        }
        for (size_t i = 0; i < L.dec_res.size(); ++i) {
            w.g_dra.push_back(bp.take<float>(B * rh * lat));
            w.g_dh.push_back(bp.take<float>(B * h * lat));
        }
        {
            const int64_t prow = (B * (H / 2) > TM_BATCH_ROWS_MAX) ? B * (H / 2) : TM_BATCH_ROWS_MAX;
            w.bias_part = bp.take<double>(prow * (size_t)(L.max_c > 2 ? L.max_c : 2) * 2);
        }
        size_t maxw = 0;
        for (const ConvL& c : L.convs) {
            const size_t f = (size_t)(c.cin + 1) * c.ks * c.ks * c.cout + c.cout + (size_t)c.cin;
            if (f > maxw) maxw = f;
        }
        // Opt-in (DMB_WG_QUEUE=1): one region of per-CTA partials per layer (the default widths: ~30 MB) and the folds of a
        // step in TWO launches (everything before the head, then the head) instead of one per layer: 79 -> 69 launches per
        // step, but measured slower as a CUDA-graph replay (0.858 against 0.850 ms at batch 256; eager: 0.844 against 0.854)
        // -- the big fold sits in the tail where only the weight-gradient stream still works.
        size_t sumw = 0;
        for (const ConvL& c : L.convs) {
            const size_t f = (size_t)(c.cin + 1) * c.ks * c.ks * c.cout + c.cout + (size_t)c.cin;
            sumw += (f + 63) & ~(size_t)63;
        }
        w.wg_queue = false;
        {
            const char* e = getenv("DMB_WG_QUEUE");
            if (e && e[0] == '1')
                w.wg_queue = (int)L.convs.size() + 1 <= WG_REDUCE_MAX && sumw * 148 * 2 * sizeof(float) <= ((size_t)64 << 20);
        }
        w.wg_part_floats = (w.wg_queue ? sumw : maxw) * 148 * 2;
        w.wg_part = bp.take<float>(w.wg_part_floats);
        w.dweff = bp.take<float>((size_t)(m.num_inputs + 1) * 16 * h2 + h2 + 16);
        w.vq_part_rows = (int)((B * lat + 255) / 256 < 296 ? (B * lat + 255) / 256 : 296);
        if (w.vq_part_rows < 2) w.vq_part_rows = 2;
        w.vq_part = bp.take<float>((size_t)w.vq_part_rows * m.num_embeddings * h);
        w.tm_scratch = bp.take<float>(tm_scratch_floats(B, (int64_t)h * lat));
        w.g_tm = bp.take<float>(B * h * lat);
        w.g_tmp = bp.take<float>(B * (h > rh ? h : rh) * lat);
    }
    w.vq_stats = bp.take<double>(2 + m.num_embeddings);
    w.recon_sum = bp.take<double>(4);
    w.scalars = bp.take<float>(8);
    w.bytes = (bp.off + 255) & ~(size_t)255;
    return 0;
}

// ---------------------------------------------------------------------------------------
// schedules
// ---------------------------------------------------------------------------------------
namespace {

struct Act {             // an activation tensor plus the affine (BN) still to be applied to it
    const float* p = nullptr;
    const float* s = nullptr;
    const float* t = nullptr;
};

struct Ctx {
    const Layout& L;
    const float* packed;
    Workspace& w;
    int64_t B;
    int mode;
    float* bnbuf;      // running stats to update in BATCH mode (may be null)
    cudaStream_t st;
    const dmb_sync_bn* sync = nullptr;     // synchronised BatchNorm across data-parallel ranks (BATCH mode only)
    unsigned* fin_ticket = nullptr;        // zeroed counter: whole-batch BatchNorm finalised inside conv_tm.cu (training step)
    bool per_sample() const { return mode == DMB_BN_PER_SAMPLE; }
    // Opt-in (DMB_TM_FIN=1): the last CTA of a conv_tm launch finalises the BatchNorm it left sums for, instead of a
    // bn_finalize / bn_backward_finalize launch (68 instead of 79 launches per step).  Measured slower at batch 256: the
    // serial fold of 296 rows by one CTA (fence, ticket, one or two L2 round trips, the fold) adds 4-5 us to every
    // producing kernel, a separate finalize launch under programmatic dependent launch costs 3-4 (0.827 against 0.820 ms).
    bool fin_in_kernel() const {
        if (!fin_ticket || mode != DMB_BN_BATCH || synced()) return false;
        const char* e = getenv("DMB_TM_FIN");
        return e && e[0] == '1';
    }
    bool synced() const { return sync && sync->world > 1 && sync->allreduce && mode == DMB_BN_BATCH; }
    // fold the per-CTA partials of one BatchNorm into [C][2] doubles and sum those across the ranks
    int exchange(const double* partials, int64_t rows, int C, double* gsum) const {
        DMB_CHECK(gsum != nullptr, "synchronised BatchNorm needs a BATCH-mode workspace");
        DMB_TRY(fold_partials(partials, rows, C, gsum, st));
        const int r = sync->allreduce(sync->user, gsum, 2 * (int64_t)C, (void*)st);
        DMB_CHECK(r == 0, "the BatchNorm allreduce callback failed (%d)", r);
        return 0;
    }
};

// Smallest batch that takes the Winograd tensor-core kernel (persistent CTAs of two patches each: below a few hundred
// patches most SMs would idle through its prologue).  DMB_WINO=0 switches it off, DMB_WINO_MIN_B overrides the threshold.
int64_t wino_min_batch() {
    const char* off = getenv("DMB_WINO");
    if (off && off[0] == '0') return INT64_MAX;
    const char* e = getenv("DMB_WINO_MIN_B");
    return e ? atoll(e) : 512;
}

bool tm_dg_enabled() {      // data-gradient form of the tensor-memory kernels in the training step (DMB_TM_DG=0: off)
    const char* e = getenv("DMB_TM_DG");
    return !(e && e[0] == '0');
}

bool tm_ct_enabled() {      // transposed form of the tensor-memory kernels (DMB_TM_CT=0: off)
    const char* e = getenv("DMB_TM_CT");
    return !(e && e[0] == '0');
}

bool convt_small_on() {      // DMB_CONVT_SMALL=0: the generic transposed-convolution kernel (convt_fwd.cu)
    const char* e = getenv("DMB_CONVT_SMALL");
    return !(e && e[0] == '0');
}

bool tm_fuse_enabled() {
    const char* e = getenv("DMB_TM_FUSE");
    return !(e && e[0] == '0');
}

// conv (or convT) layer `ci`: in -> out, optional ReLU on load; in BN modes gathers statistics and
// finalises them, returning the affine that the consumer must apply.
int run_conv(Ctx& c, int ci, const Act& in, bool in_relu, int H, int W, float* out,
             const float* skip, bool out_relu, Act* result) {
    const ConvL& l = c.L.convs[ci];
    const bool bn_live = (l.bn >= 0 && c.mode != DMB_BN_EVAL);
    BnWs* bw = bn_live ? &c.w.bn[l.bn] : nullptr;
    int Ho, Wo;
    int stat_rows = 0;      // whole-batch statistics: rows of partials when that is not B * nbands
    bool fin_done = false;  // ... already finalised inside the producing kernel
    if (l.transposed && l.pctm_off >= 0 && !bn_live && !in.s && !in_relu && tm_ct_enabled() && c.B >= tm_min_batch() &&
        conv_tm_ct_supported(l.cin, l.cout, H, W, false)) {
        // ConvTranspose2d on the tensor cores: 3x3 on the input grid + pixel shuffle (conv_tm.cu, CT form)
        ConvTmArgs a{};
        a.x = in.p; a.wtm = c.packed + l.pctm_off; a.bias = c.packed + l.pb_off; a.y = out;
        a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.ks = 4; a.stride = 2;
        a.ct = 1; a.out_relu = out_relu;
        DMB_TRY(conv_tm(a, c.st));
        Ho = 2 * H; Wo = 2 * W;
    } else if (l.transposed && l.ks == 4 && !bn_live && !in.s && !in_relu && convt_small_on() &&
               convt_small_supported(l.cin, l.cout, H, W)) {
        // four output channels, plain input: the two-pixels-per-thread streaming kernel (dec_tail.cu)
        DMB_TRY(convt_small(in.p, c.packed + l.pw_off, c.packed + l.pb_off, out, c.B, l.cin, l.cout, H, W, out_relu ? 1 : 0, c.st));
        Ho = 2 * H; Wo = 2 * W;
    } else if (l.transposed) {
        ConvTFwdArgs a{};
        a.x = in.p; a.y = out; a.w = c.packed + l.pw_off; a.bias = c.packed + l.pb_off;
        a.in_scale = in.s; a.in_shift = in.t; a.in_per_sample = c.per_sample(); a.in_relu = in_relu;
        a.out_relu = out_relu && !bn_live;
        a.stats = bn_live ? bw->part : nullptr;
        a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout;
        DMB_TRY(convt_fwd(a, c.st));
        Ho = 2 * H; Wo = 2 * W;
    } else if (bn_live && l.ptm_off >= 0 && !skip && tm_bn_mode(c.mode, c.B) &&
               conv_tm_supported(l.cin, l.cout, l.ks, l.stride, H, W)) {
        // per-patch statistics: the same tensor-core kernel with the producer's affine + ReLU on load and the
        // statistics partials in the epilogue (conv_tm.cu, BN form)
        ConvTmArgs a{};
        a.x = in.p; a.wtm = c.packed + l.ptm_off; a.bias = c.packed + l.pb_off; a.y = out;
        a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.ks = l.ks; a.stride = l.stride;
        a.bn = 1; a.in_scale = in.s; a.in_shift = in.t; a.in_per_sample = c.per_sample(); a.in_relu = in_relu;
        a.stats = bw->part;
        a.stats_batch = c.per_sample() ? 0 : 1; a.stat_rows = &stat_rows;
        DMB_CHECK(bw->nbands == conv_tm_bands(l.cin, l.cout, l.ks, l.stride, H, W), "statistics layout of conv %d", ci);
        TmFinArgs fin{};
        if (c.fin_in_kernel()) {     // the last CTA finalises the BatchNorm: no bn_finalize launch below
            const BnL& b = c.L.bns[l.bn];
            fin.ticket = c.fin_ticket; fin.mode = 1;
            fin.cnt = (double)(H / l.stride) * (W / l.stride) * (double)c.B;
            fin.gamma = c.packed + b.pg_off; fin.beta = c.packed + b.pb_off;
            fin.eps = c.L.m.bn_eps; fin.momentum = c.L.m.bn_momentum;
            fin.scale = bw->scale; fin.shift = bw->shift;
            fin.running_mean = c.bnbuf ? c.bnbuf + b.rm_off : nullptr;
            fin.running_var = c.bnbuf ? c.bnbuf + b.rv_off : nullptr;
            fin.save_mean = bw->mean; fin.save_invstd = bw->invstd;
            a.fin = &fin;
            fin_done = true;
        }
        DMB_TRY(conv_tm(a, c.st));
        if (c.per_sample()) stat_rows = 0;
        Ho = H / l.stride; Wo = W / l.stride;
    } else if (c.mode == DMB_BN_EVAL && l.ptm_off >= 0 && !in.s && c.B >= tm_min_batch() &&
               conv_tm_supported(l.cin, l.cout, l.ks, l.stride, H, W)) {
        // thin layers of the default configuration: tcgen05 with the activation operand in tensor memory (conv_tm.cu)
        ConvTmArgs a{};
        a.x = in.p; a.wtm = c.packed + l.ptm_off; a.bias = c.packed + l.pb_off; a.y = out; a.skip = skip;
        a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.ks = l.ks; a.stride = l.stride;
        a.in_relu = in_relu; a.out_relu = out_relu;
        DMB_TRY(conv_tm(a, c.st));
        Ho = H / l.stride; Wo = W / l.stride;
    } else if (c.mode == DMB_BN_EVAL && l.pwn_off >= 0 && !skip && !in.s && c.B >= wino_min_batch() &&
               conv_wino_supported(l.cin, l.cout, l.ks, l.stride, H, W)) {
        // default-width 3x3 at the latent resolution: Winograd F(2x2,3x3) on the tensor cores (conv_wino_tc.cu)
        ConvWinoArgs a{};
        a.x = in.p; a.u = c.packed + l.pwn_off; a.bias = c.packed + l.pb_off; a.y = out;
        a.B = (int)c.B; a.Cout = l.cout; a.in_relu = in_relu; a.out_relu = out_relu;
        DMB_TRY(conv_wino(a, c.st));
        Ho = H; Wo = W;
    } else {
        ConvFwdArgs a{};
        a.x = in.p; a.y = out; a.w = c.packed + l.pw_off; a.bias = c.packed + l.pb_off;
        a.bias_classes = l.bias_classes;
        a.in_scale = in.s; a.in_shift = in.t; a.in_per_sample = c.per_sample(); a.in_relu = in_relu;
        a.skip = bn_live ? nullptr : skip;
        a.out_relu = out_relu && !bn_live;
        a.stats = bn_live ? bw->part : nullptr;
        a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout;
        a.ks = l.ks; a.stride = l.stride;
        a.Ho = H / l.stride; a.Wo = W / l.stride;
        DMB_TRY(conv_fwd(a, c.st));
        Ho = a.Ho; Wo = a.Wo;
    }
    if (result) { result->p = out; result->s = nullptr; result->t = nullptr; }
    if (bn_live && fin_done) {
        if (result) { result->s = bw->scale; result->t = bw->shift; }
    } else if (bn_live) {
        const BnL& b = c.L.bns[l.bn];
        BnFinalizeArgs f{};
        f.partials = bw->part; f.B = (int)c.B; f.nbands = bw->nbands; f.C = l.cout; f.rows = stat_rows;
        f.count_per_sample = (int64_t)Ho * Wo;
        if (c.synced()) {          // statistics of the GLOBAL batch: one row of summed partials, global element count
            DMB_TRY(c.exchange(bw->part, stat_rows > 0 ? stat_rows : c.B * bw->nbands, l.cout, bw->gsum));
            f.rows = 0;
            f.partials = bw->gsum; f.B = 1; f.nbands = 1;
            f.count_per_sample = (int64_t)Ho * Wo * c.B * c.sync->world;
        }
        f.per_sample = c.per_sample();
        f.gamma = c.packed + b.pg_off; f.beta = c.packed + b.pb_off;
        f.eps = c.L.m.bn_eps; f.momentum = c.L.m.bn_momentum;
        f.scale = bw->scale; f.shift = bw->shift;
        f.running_mean = (c.bnbuf && !c.per_sample()) ? c.bnbuf + b.rm_off : nullptr;
        f.running_var = (c.bnbuf && !c.per_sample()) ? c.bnbuf + b.rv_off : nullptr;
        f.save_mean = bw->mean; f.save_invstd = bw->invstd;
        DMB_TRY(bn_finalize(f, c.st));
        if (result) { result->s = bw->scale; result->t = bw->shift; }
    }
    return 0;
}

// ResidualBlock (vq_vae.py:203-225).  `h` enters possibly with a pending affine; returns the block
// output.  If `fuse_last` is given, the last layer's merge is left to the consumer (VQ kernel).
struct Pending { Act a, b; bool valid = false; };

// EVAL-mode residual layer whose 3x3 takes the Winograd tensor-core kernel and whose 1x1 is the 32 -> 16 tail that kernel
// can fuse.  Opt-in (DMB_WINO_FUSE=1): measured on B200 the fused kernel takes as much longer as the separate HBM-bound
// 1x1 launch costs (3.59 M vs 3.62 M patches/s) -- its phases run one after another, so the extra 512 FFMA per pixel
// extend the critical path instead of filling idle issue slots.
bool wino_fused_ok(const Ctx& c, const ResL& r, const Act& h, int H, int W) {
    const ConvL& la = c.L.convs[r.a];
    const ConvL& lb = c.L.convs[r.b];
    const char* e = getenv("DMB_WINO_FUSE");
    if (!(e && e[0] == '1')) return false;
    return la.pwn_off >= 0 && !h.s && c.B >= wino_min_batch() && la.cout == 32 &&
           conv_wino_supported(la.cin, la.cout, la.ks, la.stride, H, W) &&
           lb.ks == 1 && lb.stride == 1 && lb.cin == 32 && lb.cout == 16 && !lb.transposed;
}
int run_res(Ctx& c, const std::vector<ResL>& res, std::vector<float*>& ra, std::vector<float*>& rb,
            std::vector<float*>& hs, Act h, int H, int W, float* final_out, Pending* fuse_last, Act* out) {
    const int64_t hw = (int64_t)H * W;
    for (size_t i = 0; i < res.size(); ++i) {
        const bool last = (i + 1 == res.size());
        float* dst = (last && final_out) ? final_out : hs[i];
        Act a1, b1;
        if (c.mode == DMB_BN_EVAL && c.L.convs[res[i].a].ptm_tail == res[i].b && !h.s && c.B >= tm_min_batch() && tm_fuse_enabled()) {
            // whole residual layer in ONE tensor-core kernel: 3x3 -> ReLU -> 1x1 -> + skip (conv_tm.cu, FUSE)
            const ConvL& la = c.L.convs[res[i].a];
            const ConvL& lb = c.L.convs[res[i].b];
            ConvTmArgs a{};
            a.x = h.p; a.wtm = c.packed + la.ptm_off; a.bias = c.packed + la.pb_off; a.y = dst; a.skip = nullptr;
            a.B = (int)c.B; a.Cin = la.cin; a.H = H; a.W = W; a.Cout = la.cout; a.ks = la.ks; a.stride = la.stride;
            a.in_relu = 1; a.out_relu = 1; a.bias2 = c.packed + lb.pb_off;
            DMB_TRY(conv_tm(a, c.st));
            h = Act(); h.p = dst;
        } else if (c.mode == DMB_BN_EVAL && wino_fused_ok(c, res[i], h, H, W)) {
            // conv3x3 (Winograd, tensor cores) -> ReLU -> conv1x1 + skip in ONE kernel (conv_wino_tc.cu, FUSE)
            const ConvL& la = c.L.convs[res[i].a];
            const ConvL& lb = c.L.convs[res[i].b];
            ConvWinoArgs a{};
            a.x = h.p; a.u = c.packed + la.pwn_off; a.bias = c.packed + la.pb_off; a.y = nullptr;
            a.B = (int)c.B; a.Cout = la.cout; a.in_relu = 1; a.out_relu = 1;
            a.w2 = c.packed + lb.pw_off; a.bias2 = c.packed + lb.pb_off; a.y2 = dst;
            DMB_TRY(conv_wino(a, c.st));
            h = Act(); h.p = dst;
        } else if (c.mode == DMB_BN_EVAL) {
            DMB_TRY(run_conv(c, res[i].a, h, true, H, W, ra[i], nullptr, true, &a1));
            DMB_TRY(run_conv(c, res[i].b, a1, false, H, W, dst, h.p, false, &b1));
            h = b1;
        } else {
            DMB_TRY(run_conv(c, res[i].a, h, true, H, W, ra[i], nullptr, false, &a1));
            DMB_TRY(run_conv(c, res[i].b, a1, true, H, W, rb[i], nullptr, false, &b1));
            if (last && fuse_last) {
                fuse_last->a = h; fuse_last->b = b1; fuse_last->valid = true;
                *out = Act();
                return 0;
            }
            AffineAddArgs aa{};
            aa.a = h.p; aa.sa = h.s; aa.ta = h.t; aa.b = b1.p; aa.sb = b1.s; aa.tb = b1.t;
            aa.per_sample = c.per_sample(); aa.out = dst; aa.B = c.B; aa.C = c.L.m.num_hiddens; aa.HW = (int)hw;
            DMB_TRY(affine_add(aa, c.st));
            h = Act(); h.p = dst;
        }
    }
    *out = h;
    return 0;
}

bool tc_enabled() {
    const char* e = getenv("DMB_TC");      // read per call: tests flip it to compare with the CUDA-core kernels
    return !(e && e[0] == '0');
}

// one tensor-core layer over NHWC activations (EVAL mode: BatchNorm folded into the packed weights / bias)
int run_conv_tc(Ctx& c, int ci, const float* in, bool in_relu, int H, int W, float* out, const float* skip_nhwc,
                bool out_relu, bool out_nhwc) {
    const ConvL& l = c.L.convs[ci];
    DMB_CHECK(l.ptc_off >= 0, "conv %d has no tensor-core weights", ci);
    ConvTcArgs a{};
    a.x = in; a.wtc = c.packed + l.ptc_off; a.bias = c.packed + l.pb_off; a.y = out; a.skip = skip_nhwc;
    a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.ks = l.ks; a.stride = l.stride;
    a.in_relu = in_relu; a.out_relu = out_relu; a.out_nhwc = out_nhwc; a.skip_nhwc = 1;
    return conv_tc(a, c.st);
}

// EVAL-mode encoder of the wide configurations: CUDA-core head (thin input), NCHW -> NHWC once, then every layer on
// tcgen05; the last layer writes NCHW (the layout of z_before and of the quantiser).
int run_encoder_tc(Ctx& c, const float* x, float* zb_out) {
    const Layout& L = c.L;
    const dmb_model& m = L.m;
    const int H = m.height, W = m.width;
    {   // head on the CUDA cores (2 input channels), written channel-last when the TMA kernel has that form
        const ConvL& l = L.convs[L.e1];
        ConvFwdArgs a{};
        a.x = x; a.y = c.w.y1t; a.w = c.packed + l.pw_off; a.bias = c.packed + l.pb_off; a.bias_classes = l.bias_classes;
        a.out_relu = 1; a.out_nhwc = 1;
        a.B = (int)c.B; a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.ks = l.ks; a.stride = l.stride;
        a.Ho = H / l.stride; a.Wo = W / l.stride;
        int r = 1;
        const char* e = getenv("DMB_HEAD_NHWC");
        if (!(e && e[0] == '0')) r = conv_tma(a, c.st);
        if (r < 0) return r;
        if (r == 1) {                        // no channel-last instantiation for this shape: NCHW + one transpose
            Act in; in.p = x;
            Act a1;
            DMB_TRY(run_conv(c, L.e1, in, false, H, W, c.w.y1, nullptr, true, &a1));
            DMB_TRY(nchw_to_nhwc(c.w.y1, c.w.y1t, c.B, l.cout, (H / 2) * (W / 2), c.st));
        }
    }
    const bool no_res = L.enc_res.empty();
    const float* h;
    if (m.arch == DMB_ARCH_Z16) {
        DMB_TRY(run_conv_tc(c, L.e2, c.w.y1t, false, H / 2, W / 2, c.w.y2, nullptr, true, true));
        DMB_TRY(run_conv_tc(c, L.e3, c.w.y2, false, H / 4, W / 4, c.w.y3, nullptr, true, true));
        float* y4 = no_res ? zb_out : c.w.y4;
        DMB_TRY(run_conv_tc(c, L.e4, c.w.y3, false, H / 8, W / 8, y4, nullptr, false, !no_res));
        h = y4;
    } else {
        float* y2 = no_res ? zb_out : c.w.y2;
        DMB_TRY(run_conv_tc(c, L.e2, c.w.y1t, false, H / 2, W / 2, y2, nullptr, false, !no_res));
        h = y2;
    }
    for (size_t i = 0; i < L.enc_res.size(); ++i) {
        const bool last = (i + 1 == L.enc_res.size());
        float* dst = last ? zb_out : c.w.ehs[i];
        DMB_TRY(run_conv_tc(c, L.enc_res[i].a, h, true, L.lh, L.lw, c.w.era[i], nullptr, true, true));
        DMB_TRY(run_conv_tc(c, L.enc_res[i].b, c.w.era[i], false, L.lh, L.lw, dst, h, false, !last));
        h = dst;
    }
    return 0;
}

// Sub-batch (patches) of the head -> enc.4 pair in EVAL mode, 0 = whole batch at once (the default).  Measured on B200
// (profiles/README.md, round 2): DMB_ENC_SUBBATCH=256 cuts the step's DRAM traffic from 625 to 443 KB per patch (the head
// output never reaches HBM) but the 64 short dependent launches per 8192 patches cost more than the traffic saves
// (4.94 -> 3.59 M patches/s), so it is opt-in.
int64_t enc_subbatch(int64_t B) {
    const char* e = getenv("DMB_ENC_SUBBATCH");
    const int64_t sb = e ? atoll(e) : 0;
    if (sb <= 0 || B < 4 * sb) return 0;
    return sb;
}

// encoder: x -> z_before (written to zb_out unless the final merge is handed to `fuse`)
int run_encoder(Ctx& c, const float* x, float* zb_out, Pending* fuse) {
    const Layout& L = c.L;
    const dmb_model& m = L.m;
    const int H = m.height, W = m.width;
    const bool ev = c.mode == DMB_BN_EVAL;
    if (ev && L.tc && c.w.y1t && zb_out && tc_enabled()) return run_encoder_tc(c, x, zb_out);
    Act in; in.p = x;
    Act a1, a2, a3, a4, out;
    if (m.arch == DMB_ARCH_Z16) {
        const int64_t SB = ev ? enc_subbatch(c.B) : 0;
        if (SB > 0) {
            // The head's output (h/2 x H/2 x W/2 floats per patch, as large as the input) is the biggest tensor of the
            // step and is read exactly once, by the next layer.  The two layers therefore walk the batch in sub-batches
            // whose head output fits the 126 MB L2 and alternate between two slots: the consumer reads it from L2 and
            // the slot is overwritten while its lines are still resident, so it never has to reach HBM.
            const size_t x_per = (size_t)m.num_inputs * H * W;
            const size_t y1_per = (size_t)(m.num_hiddens / 2) * (H / 2) * (W / 2);
            const size_t y2_per = (size_t)m.num_hiddens * (H / 4) * (W / 4);
            int slot = 0;
            for (int64_t b0 = 0; b0 < c.B; b0 += SB, slot ^= 1) {
                const int64_t n = (c.B - b0 < SB) ? c.B - b0 : SB;
                Ctx cs{c.L, c.packed, c.w, n, c.mode, c.bnbuf, c.st, c.sync};
                Act xin, o1, o2;
                xin.p = x + (size_t)b0 * x_per;
                float* y1s = c.w.y1 + (size_t)slot * SB * y1_per;
                DMB_TRY(run_conv(cs, L.e1, xin, false, H, W, y1s, nullptr, true, &o1));
                DMB_TRY(run_conv(cs, L.e2, o1, false, H / 2, W / 2, c.w.y2 + (size_t)b0 * y2_per, nullptr, true, &o2));
            }
            a2 = Act(); a2.p = c.w.y2;
        } else {
            DMB_TRY(run_conv(c, L.e1, in, false, H, W, c.w.y1, nullptr, ev, &a1));
            DMB_TRY(run_conv(c, L.e2, a1, !ev, H / 2, W / 2, c.w.y2, nullptr, ev, &a2));
        }
        DMB_TRY(run_conv(c, L.e3, a2, !ev, H / 4, W / 4, c.w.y3, nullptr, ev, &a3));
        float* y4 = (L.enc_res.empty() && zb_out && ev) ? zb_out : c.w.y4;
        DMB_TRY(run_conv(c, L.e4, a3, !ev, H / 8, W / 8, y4, nullptr, false, &a4));
    } else {
        DMB_TRY(run_conv(c, L.e1, in, false, H, W, c.w.y1, nullptr, ev, &a1));
        float* y2 = (L.enc_res.empty() && zb_out && ev) ? zb_out : c.w.y2;
        DMB_TRY(run_conv(c, L.e2, a1, !ev, H / 2, W / 2, y2, nullptr, false, &a4));
    }
    if (L.enc_res.empty()) {
        if (a4.s) {                       // materialise the pending BN affine (no residual block)
            AffineAddArgs aa{};
            aa.a = a4.p; aa.sa = a4.s; aa.ta = a4.t; aa.b = nullptr;
            aa.per_sample = c.per_sample(); aa.out = zb_out; aa.B = c.B; aa.C = m.num_hiddens;
            aa.HW = L.lh * L.lw;
            DMB_TRY(affine_add(aa, c.st));
        }
        return 0;
    }
    DMB_TRY(run_res(c, L.enc_res, c.w.era, c.w.erb, c.w.ehs, a4, L.lh, L.lw, zb_out, fuse, &out));
    return 0;
}

int run_vq(Ctx& c, const float* codebook, const float* z, const Pending* pre, float* zb_out,
           float* z_after, int32_t* idx, double* stats) {
    VqArgs a{};
    a.codebook = codebook; a.B = c.B; a.D = c.L.D; a.P = c.L.lh * c.L.lw; a.K = c.L.m.num_embeddings;
    a.z_st = z_after; a.idx = idx; a.stats = stats;
    if (pre && pre->valid) {
        a.pre_a = pre->a.p; a.pre_sa = pre->a.s; a.pre_ta = pre->a.t;
        a.pre_b = pre->b.p; a.pre_sb = pre->b.s; a.pre_tb = pre->b.t;
        a.pre_per_sample = c.per_sample();
        a.z_before_out = zb_out;
    } else {
        a.z = z;
    }
    return vq_forward(a, c.st);
}

// The fused decoder tail of the training step (dec_tail.cu); DMB_DEC_TAIL=0 keeps the separate launches.
bool dec_tail_on(const dmb_model& m) {
    const char* e = getenv("DMB_DEC_TAIL");
    if (e && e[0] == '0') return false;
    return m.arch == DMB_ARCH_Z16 && dec_tail_supported(m.num_hiddens / 4, m.num_inputs, m.height * m.width);
}

// ... with dec.4 in front of it as well (dec_tail2_forward); DMB_DEC_TAIL2=0 keeps dec.4 a separate launch.
bool dec_tail2_on(const Layout& L) {
    const char* e = getenv("DMB_DEC_TAIL2");
    if (e && e[0] == '0') return false;
    const dmb_model& m = L.m;
    if (!dec_tail_on(m)) return false;
    const ConvL& l4 = L.convs[L.d2];
    const ConvL& l6 = L.convs[L.d3];
    return l4.transposed && l4.ks == 4 && dec_tail2_supported(l4.cin, l4.cout, l6.cout, 4 * L.lh, 4 * L.lw);
}

int run_decoder(Ctx& c, const float* za, float* decoded, bool skip_tail = false, bool skip_d2 = false) {
    const Layout& L = c.L;
    const dmb_model& m = L.m;
    const bool ev = c.mode == DMB_BN_EVAL;
    Act in; in.p = za;
    Act a1, a2, a3, o;
    if (m.arch == DMB_ARCH_Z16) {
        DMB_CHECK(c.w.t1 && c.w.t2 && c.w.t3, "decoder needs a workspace carved with keep_activations=1");
        DMB_TRY(run_conv(c, L.d0, in, false, L.lh, L.lw, c.w.t1, nullptr, true, &a1));
        DMB_TRY(run_conv(c, L.d1, a1, false, 2 * L.lh, 2 * L.lw, c.w.t2, nullptr, true, &a2));
        if (!skip_d2) DMB_TRY(run_conv(c, L.d2, a2, false, 4 * L.lh, 4 * L.lw, c.w.t3, nullptr, true, &a3));
        if (!skip_tail) DMB_TRY(run_conv(c, L.d3, a3, false, 8 * L.lh, 8 * L.lw, decoded, nullptr, false, &o));
    } else {
        DMB_CHECK(c.w.t1, "decoder needs a workspace carved with keep_activations=1");
        Act h;
        DMB_TRY(run_res(c, L.dec_res, c.w.dra, c.w.drb, c.w.dhs, in, L.lh, L.lw, nullptr, nullptr, &h));
        DMB_TRY(run_conv(c, L.d0, h, false, L.lh, L.lw, c.w.t1, nullptr, ev, &a1));
        DMB_TRY(run_conv(c, L.d1, a1, !ev, 2 * L.lh, 2 * L.lw, decoded, nullptr, false, &o));
    }
    return 0;
}

__global__ void recon_loss_kernel(const float* __restrict__ dec, const float* __restrict__ x,
                                  const float* __restrict__ mask, int mask_c, const float* __restrict__ cvar,
                                  int64_t total4, int C, int hw4, double* out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ double red[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t plane = i / hw4;
        const int c = (int)(plane % C);
        const int64_t b = plane / C;
        const float4 d = __ldg(reinterpret_cast<const float4*>(dec) + i);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
        if (mask) {
            const int64_t mi = (mask_c == 1) ? (b * hw4 + (i - plane * hw4)) : i;
            mk = __ldg(reinterpret_cast<const float4*>(mask) + mi);
        }
        const float cv = __ldg(cvar + c);
        // mse_loss(decoded*mask, inputs*mask, 'none') / channel_var  (vq_vae.py:322)
        float e0 = d.x * mk.x - v.x * mk.x, e1 = d.y * mk.y - v.y * mk.y;
        float e2 = d.z * mk.z - v.z * mk.z, e3 = d.w * mk.w - v.w * mk.w;
        acc += (double)((e0 * e0) / cv) + (double)((e1 * e1) / cv) + (double)((e2 * e2) / cv) + (double)((e3 * e3) / cv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(out, s);
    }
}


// ---------------------------------------------------------------------------------------
// training step: backward schedule
// ---------------------------------------------------------------------------------------
__global__ void recon_grad_kernel(const float* __restrict__ dec, const float* __restrict__ x,
                                  const float* __restrict__ mask, int mask_c, const float* __restrict__ cvar,
                                  int64_t total4, int C, int hw4, float scale, float* __restrict__ gd) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // d/d dec of  scale * sum ((dec*m - x*m)^2 / cv)  =  scale * 2 (dec*m - x*m) m / cv
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t plane = i / hw4;
        const int c = (int)(plane % C);
        const int64_t b = plane / C;
        const float4 d = __ldg(reinterpret_cast<const float4*>(dec) + i);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
        if (mask) {
            const int64_t mi = (mask_c == 1) ? (b * hw4 + (i - plane * hw4)) : i;
            mk = __ldg(reinterpret_cast<const float4*>(mask) + mi);
        }
        const float k2 = 2.f * scale / __ldg(cvar + c);
        float4 o;
        o.x = k2 * (d.x * mk.x - v.x * mk.x) * mk.x; o.y = k2 * (d.y * mk.y - v.y * mk.y) * mk.y;
        o.z = k2 * (d.z * mk.z - v.z * mk.z) * mk.z; o.w = k2 * (d.w * mk.w - v.w * mk.w) * mk.w;
        reinterpret_cast<float4*>(gd)[i] = o;
    }
}

// vq2: [0] vq loss, [1] perplexity, [4] time-matching loss (when w_tm != 0).  n_out = 4 (legacy) or 8.
__global__ void train_losses_kernel(const float* vq2, const double* recon_sum, double n_recon, float w_r, float w_c,
                                    float w_tm, int has_tm, int n_out, float* out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const float recon = (float)(recon_sum[0] / n_recon);
    const float tm = has_tm ? vq2[4] : 0.f;
    out[0] = recon;
    out[1] = vq2[0];
    out[2] = w_r * recon + w_c * vq2[0] + (has_tm ? w_tm * tm : 0.f);
    out[3] = vq2[1];
    if (n_out > 4) { out[4] = tm; out[5] = 0.f; out[6] = 0.f; out[7] = 0.f; }
}

struct GradT {            // dL/d(conv output) = A*g + Bc*y + Cc  (A == nullptr: just g)
    const float* g = nullptr; const float* y = nullptr;
    const float* A = nullptr; const float* Bc = nullptr; const float* Cc = nullptr;
};

// Weight gradients run on a library-owned side stream, concurrently with the data-gradient chain on the caller's
// stream: at training batch sizes every backward kernel is a few hundred CTAs and latency-bound, and the weight
// gradient of a layer is off the critical path (only the data gradient feeds the next layer).  Fork / join through
// events, which a CUDA-graph capture of the caller's stream records as ordinary dependencies.  One side stream per
// host thread and device; DMB_TRAIN_OVERLAP=0 keeps everything on the caller's stream.
struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
int side_stream(SideStream** out) {
    static thread_local SideStream side[64];
    *out = nullptr;
    const char* e = getenv("DMB_TRAIN_OVERLAP");
    if (e && e[0] == '0') return 0;
    int dev = 0;
    DMB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return 0;
    SideStream& ss = side[dev];
    if (!ss.s) {
        DMB_CUDA(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
        DMB_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        DMB_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
    }
    *out = &ss;
    return 0;
}

struct Bwd {
    Ctx& c;
    const float* params;
    float* grads;
    cudaStream_t st;
    SideStream* side = nullptr;
    bool forked = false;
    int stat_rows = 0;      // rows of partial sums the LAST dgrad_layer call left when that is not B * nbands (conv_tm.cu)
    int fin_bn = -1;        // BatchNorm whose backward the LAST dgrad_layer call finalised inside its kernel (-1: none)
    WgReduceQueue rq;       // folds of the weight-gradient partials, launched once (flush_wgrad) when wg_queue
    size_t wg_off = 0;
    bool ps() const { return c.per_sample(); }
    // in-kernel BatchNorm-backward finalize for the sums a data-gradient kernel leaves in `stats` (a bnb[].part)
    bool fill_fin(double* stats, int64_t count, TmFinArgs& fin) {
        fin_bn = -1;
        if (!stats || !c.fin_in_kernel()) return false;
        int bn = -1;
        for (size_t j = 0; j < c.w.bnb.size(); ++j)
            if (c.w.bnb[j].part == stats) bn = (int)j;
        if (bn < 0) return false;          // (bias_part: a plain sum, folded by sum_partials)
        const BnL& b = c.L.bns[bn];
        Workspace::BnB& bb = c.w.bnb[bn];
        BnWs& bw = c.w.bn[bn];
        fin = TmFinArgs{};
        fin.ticket = c.fin_ticket; fin.mode = 2; fin.cnt = (double)count * (double)c.B;
        fin.gamma = c.packed + b.pg_off; fin.mean = bw.mean; fin.invstd = bw.invstd;
        fin.A = bb.A; fin.Bc = bb.Bc; fin.Cc = bb.Cc; fin.dgamma = grads + b.g_off; fin.dbeta = grads + b.b_off;
        fin_bn = bn;
        return true;
    }
    // scratch for the next weight-gradient launch (need floats) and the queue its fold goes to (nullptr: fold at once)
    int wgrad_scratch(size_t need, float** part, WgReduceQueue** q) {
        if (c.w.wg_queue) {
            DMB_CHECK(wg_off + need <= c.w.wg_part_floats, "wgrad scratch too small (queued folds)");
            *part = c.w.wg_part + wg_off;
            wg_off += (need + 63) & ~(size_t)63;
            *q = &rq;
        } else {
            DMB_CHECK(need <= c.w.wg_part_floats, "wgrad scratch too small");
            *part = c.w.wg_part;
            *q = nullptr;
        }
        return 0;
    }
    int flush_wgrad() {     // on the stream the weight gradients run on
        if (rq.n == 0) return 0;
        cudaStream_t sw = side ? side->s : st;
        if (side && !forked) DMB_TRY(wgrad_stream(&sw));
        return wgrad_reduce_flush(rq, sw);
    }
    int L_lat_h() const { return c.L.lh; }
    int L_lat_w() const { return c.L.lw; }
    // stream for the next weight-gradient launch: the side stream once it has seen everything issued so far
    int wgrad_stream(cudaStream_t* out) {
        if (!side) { *out = st; return 0; }
        DMB_CUDA(cudaEventRecord(side->fork, st));
        DMB_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
        forked = true;
        *out = side->s;
        return 0;
    }
    int join() {
        if (side && forked) {
            DMB_CUDA(cudaEventRecord(side->join, side->s));
            DMB_CUDA(cudaStreamWaitEvent(st, side->join, 0));
            forked = false;
        }
        return 0;
    }

    // weight / bias gradient of conv layer li (input activation xin, H x W).  For ConvTranspose2d the roles
    // are swapped: `G` must be a plain tensor (no BN after the decoder's ConvT in z16) of shape (Cout,2H,2W).
    int wgrad_layer(int li, const GradT& G, const Act& xin, bool x_relu, int H, int W, bool with_bias) {
        const ConvL& l = c.L.convs[li];
        WgradArgs a{};
        a.B = (int)c.B; a.ks = l.ks; a.stride = l.stride; a.partials = c.w.wg_part;
        if (l.transposed) {
            // roles swapped: "gy" := the ConvTranspose input activation relu?(xin*s+t), "act" := dL/d(output)
            a.g = xin.p; a.ga = xin.s; a.gc = xin.t; a.gb = nullptr; a.y = nullptr; a.g_per_sample = ps();
            a.g_relu = x_relu;
            a.Cout = l.cin; a.Ho = H; a.Wo = W;
            a.x = G.g; a.xs = G.A; a.xt = G.Cc; a.x2 = G.A ? G.y : nullptr; a.xb = G.Bc; a.x_per_sample = ps();
            a.Cin = l.cout; a.H = 2 * H; a.W = 2 * W;
        } else {
            a.g = G.g; a.y = G.y; a.ga = G.A; a.gb = G.Bc; a.gc = G.Cc; a.g_per_sample = ps();
            a.x = xin.p; a.xs = xin.s; a.xt = xin.t; a.x_per_sample = ps(); a.x_relu = x_relu;
            a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.Ho = H / l.stride; a.Wo = W / l.stride;
        }
        int ncta = 0;
        const int pf = wgrad_partial_floats(a, &ncta);
        DMB_CHECK(pf > 0, "no weight-gradient plan for layer %d", li);
        WgReduceQueue* q = nullptr;
        DMB_TRY(wgrad_scratch((size_t)pf * ncta, &a.partials, &q));
        float* db = (with_bias && !l.transposed) ? grads + l.b_off : nullptr;
        cudaStream_t sw;
        DMB_TRY(wgrad_stream(&sw));
        return wgrad(a, grads + l.w_off, db, nullptr, sw, q);
    }

    // data gradient of conv layer li: G (at the layer's output, Ho x Wo) -> gout (at its input, H x W), gated by
    // the ReLU of the tensor that fed the layer, plus optional skip and BatchNorm-backward sums.
    int dgrad_layer(int li, const GradT& G, int H, int W, float* gout, const Act* gate, const float* skip,
                    double* stats, const float* stat_src, int* nbands) {
        const ConvL& l = c.L.convs[li];
        DMB_CHECK(l.pdw_off >= 0, "layer %d has no data-gradient weights", li);
        const float* wd = c.packed + l.pdw_off;
        const float* zero = c.packed + c.L.pzero_off;
        const bool up = (!l.transposed && l.stride == 2);      // conv stride 2 -> transposed-conv kernel
        stat_rows = 0;
        fin_bn = -1;
        if (up && l.pdtm_off >= 0 && !ps() && tm_dg_enabled() && tm_ct_enabled() && c.B >= tm_min_batch() && !skip &&
            conv_tm_ct_supported(l.cout, l.cin, H / 2, W / 2, true)) {
            // stride-2 convolution: its data gradient is a transposed convolution over the output map -- tensor cores,
            // BatchNorm backward applied on load, gate and the next BatchNorm backward's sums in the epilogue
            ConvTmArgs a{};
            a.x = G.g; a.wtm = c.packed + l.pdtm_off; a.bias = zero; a.y = gout;
            a.B = (int)c.B; a.Cin = l.cout; a.Cout = l.cin; a.H = H / 2; a.W = W / 2; a.ks = 4; a.stride = 2;
            a.ct = 1; a.dg = 1;
            if (G.A) { a.x2 = G.y; a.in_a = G.A; a.in_b = G.Bc; a.in_c = G.Cc; }
            if (gate) { a.mask_src = gate->p; a.mask_s = gate->s; a.mask_t = gate->t; }
            a.stats = stats; a.stat_src = stat_src; a.stats_batch = 1; a.stat_rows = &stat_rows;
            if (nbands) *nbands = 0;
            TmFinArgs fin{};
            if (fill_fin(stats, (int64_t)H * W, fin)) a.fin = &fin;
            DMB_TRY(conv_tm(a, st));
            if (!stats) stat_rows = 0;
            return 0;
        }
        if (!up && l.pdtm_off >= 0 && !ps() && tm_dg_enabled() && c.B >= tm_min_batch() &&
            (!G.A || (c.w.g_tmp && H == L_lat_h() && W == L_lat_w() && !l.transposed))) {
            // tensor cores, activation operand in tensor memory (conv_tm.cu, data-gradient form).  A BatchNorm-backward
            // gradient  A*g + Bc*y + Cc  is materialised first (one small elementwise launch at the latent resolution).
            ConvTmArgs a{};
            a.x = G.g;
            a.Cin = l.cout; a.Cout = l.cin; a.B = (int)c.B;
            if (l.transposed) { a.ks = 4; a.stride = 2; a.H = 2 * H; a.W = 2 * W; }
            else { a.ks = l.ks; a.stride = 1; a.H = H; a.W = W; }
            if (G.A && conv_tm_dg_dual(a.Cin, a.Cout, a.ks, a.stride, a.H, a.W)) {
                a.x2 = G.y; a.in_a = G.A; a.in_b = G.Bc; a.in_c = G.Cc;       // BatchNorm backward applied on load
            } else if (G.A) {
                AffineAddArgs aa{};
                aa.a = G.g; aa.sa = G.A; aa.ta = G.Cc; aa.b = G.y; aa.sb = G.Bc; aa.tb = zero;
                aa.per_sample = 0; aa.out = c.w.g_tmp; aa.B = c.B; aa.C = a.Cin; aa.HW = a.H * a.W;
                DMB_TRY(affine_add(aa, st));
                a.x = c.w.g_tmp;
            }
            a.wtm = c.packed + l.pdtm_off; a.bias = zero; a.y = gout; a.skip = skip; a.dg = 1;
            if (gate) { a.mask_src = gate->p; a.mask_s = gate->s; a.mask_t = gate->t; }
            a.stats = stats; a.stat_src = stat_src; a.stats_batch = 1; a.stat_rows = &stat_rows;
            if (nbands) *nbands = 0;
            TmFinArgs fin{};
            if (fill_fin(stats, (int64_t)a.H / a.stride * (a.W / a.stride), fin)) a.fin = &fin;
            DMB_TRY(conv_tm(a, st));
            if (!stats) stat_rows = 0;
            return 0;
        }
        if (up) {
            ConvTFwdArgs a{};
            a.x = G.g; a.x2 = G.y; a.in_scale = G.A; a.in_b = G.Bc; a.in_shift = G.Cc; a.in_per_sample = ps();
            a.y = gout; a.w = wd; a.bias = zero;
            if (gate) { a.mask_src = gate->p; a.mask_s = gate->s; a.mask_t = gate->t; a.mask_per_sample = ps(); }
            a.stats = stats; a.stat_src = stat_src;
            a.B = (int)c.B; a.Cin = l.cout; a.H = H / 2; a.W = W / 2; a.Cout = l.cin;
            DMB_CHECK(skip == nullptr, "skip not supported on the transposed data-gradient path");
            if (nbands) *nbands = convt_fwd_bands(a.Cin, a.Cout, a.H, a.W);
            return convt_fwd(a, st);
        }
        ConvFwdArgs a{};
        a.x = G.g; a.x2 = G.y; a.in_scale = G.A; a.in_b = G.Bc; a.in_shift = G.Cc; a.in_per_sample = ps();
        a.y = gout; a.w = wd; a.bias = zero;
        if (gate) { a.mask_src = gate->p; a.mask_s = gate->s; a.mask_t = gate->t; a.mask_per_sample = ps(); }
        a.skip = skip; a.stats = stats; a.stat_src = stat_src;
        a.B = (int)c.B; a.Cin = l.cout; a.Cout = l.cin;
        if (l.transposed) {      // ConvTranspose2d -> stride-2 conv over the (2H x 2W) output gradient
            a.ks = 4; a.stride = 2; a.H = 2 * H; a.W = 2 * W; a.Ho = H; a.Wo = W;
        } else {
            a.ks = l.ks; a.stride = 1; a.H = H; a.W = W; a.Ho = H; a.Wo = W;
        }
        // The TMA kernel has no dual-tensor transform on load.  Where it has an instantiation for this geometry the
        // BatchNorm-backward gradient  A*g + Bc*y + Cc  is materialised first (one small elementwise launch; the
        // latent-resolution tensors are a few MB at training batch sizes) and fed to it as a plain input.
        if (G.A && !ps() && c.w.g_tmp && a.H == L_lat_h() && a.W == L_lat_w() &&
            conv_tma_bands(a.ks, a.stride, a.Cin, a.Cout, a.H, a.W, 0) > 0) {
            AffineAddArgs aa{};
            aa.a = G.g; aa.sa = G.A; aa.ta = G.Cc; aa.b = G.y; aa.sb = G.Bc; aa.tb = zero;
            aa.per_sample = 0; aa.out = c.w.g_tmp; aa.B = c.B; aa.C = a.Cin; aa.HW = a.H * a.W;
            DMB_TRY(affine_add(aa, st));
            a.x = c.w.g_tmp; a.x2 = nullptr; a.in_scale = nullptr; a.in_b = nullptr; a.in_shift = nullptr;
        }
        if (nbands) {
            const bool plain = !a.x2 && !a.in_b;
            *nbands = conv_fwd_bands(a.ks, a.stride, a.Cin, a.Cout, a.Ho, a.Wo, plain, c.B);
        }
        return conv_fwd(a, st);
    }

    // BatchNorm backward of layer li's BN from the partial sums the gradient producer left in bnb[].part
    int bn_bwd(int li, int nbands, int64_t count, GradT* G, const float* g, const float* y) {
        const ConvL& l = c.L.convs[li];
        const BnL& b = c.L.bns[l.bn];
        Workspace::BnB& bb = c.w.bnb[l.bn];
        BnWs& bw = c.w.bn[l.bn];
        if (fin_bn == l.bn) {      // the kernel that left these sums finalised them too
            fin_bn = -1;
            G->g = g; G->y = y; G->A = bb.A; G->Bc = bb.Bc; G->Cc = bb.Cc;
            return 0;
        }
        BnBwdArgs a{};
        a.partials = bb.part; a.B = (int)c.B; a.nbands = nbands; a.C = l.cout; a.count_per_sample = count;
        a.rows = stat_rows;      // (set by the dgrad_layer call that left these sums)
        if (c.synced()) {
            DMB_TRY(c.exchange(bb.part, stat_rows > 0 ? stat_rows : c.B * nbands, l.cout, bb.gsum));
            a.rows = 0;
            a.partials = bb.gsum; a.B = 1; a.nbands = 1; a.count_per_sample = count * c.B * c.sync->world;
            a.grad_div = c.sync->world;
        }
        a.per_sample = ps(); a.gamma = c.packed + b.pg_off; a.mean = bw.mean; a.invstd = bw.invstd;
        a.A = bb.A; a.Bc = bb.Bc; a.Cc = bb.Cc; a.dgamma = grads + b.g_off; a.dbeta = grads + b.b_off;
        DMB_TRY(bn_backward_finalize(a, st));
        G->g = g; G->y = y; G->A = bb.A; G->Bc = bb.Bc; G->Cc = bb.Cc;
        return 0;
    }
};

__global__ void channel_sum_kernel(const float* __restrict__ g, int64_t B, int C, int hw, float* __restrict__ out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // out[c] = sum over (b, pixel) of g[b][c][pixel]; one CTA per channel, fixed order
    __shared__ double red[8];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t b = 0; b < B; ++b) {
        const float* p = g + ((size_t)b * C + c) * hw;
        float part = 0.f;
        for (int i = threadIdx.x; i < hw; i += blockDim.x) part += p[i];
        s += (double)part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
        out[c] = (float)t;
    }
}

// ResidualBlock backward (vq_vae.py:203-225), last layer first.  On entry `g_cur` is the gradient at the
// block output and the BatchNorm-backward sums of the LAST layer's second BN are already in bnb[].part
// (nb_cur bands).  h0 = the block input (with a pending affine if it is a BN output); prev_bn / stat0 name the
// BatchNorm that produced h0 (-1: none) so that its sums are gathered while the input gradient is written.
int res_backward(Bwd& B, const std::vector<ResL>& res, std::vector<float*>& ra, std::vector<float*>& rb,
                 std::vector<float*>& hs, std::vector<float*>& g_ra, std::vector<float*>& g_h, const Act& h0,
                 int prev_bn0, const float* stat0, const float* g_cur, int nb_cur, int lh, int lw,
                 const float** g_out, int* nb_out) {
    Ctx& c = B.c;
    const Layout& L = c.L;
    Workspace& w = c.w;
    const int P = lh * lw;
    int nb = nb_cur;
    for (int i = (int)res.size() - 1; i >= 0; --i) {
        const int la = res[i].a, lb = res[i].b;
        GradT Gb, Ga;
        DMB_TRY(B.bn_bwd(lb, nb_cur, P, &Gb, g_cur, rb[i]));
        Act a_in; a_in.p = ra[i]; a_in.s = w.bn[L.convs[la].bn].scale; a_in.t = w.bn[L.convs[la].bn].shift;
        DMB_TRY(B.wgrad_layer(lb, Gb, a_in, true, lh, lw, true));
        DMB_TRY(B.dgrad_layer(lb, Gb, lh, lw, g_ra[i], &a_in, nullptr, w.bnb[L.convs[la].bn].part, ra[i], &nb));
        DMB_TRY(B.bn_bwd(la, nb, P, &Ga, g_ra[i], ra[i]));
        Act hin;
        if (i == 0) hin = h0; else hin.p = hs[i - 1];
        DMB_TRY(B.wgrad_layer(la, Ga, hin, true, lh, lw, true));
        const int prev_bn = (i == 0) ? prev_bn0 : L.convs[res[i - 1].b].bn;
        const float* stat_src = (i == 0) ? stat0 : rb[i - 1];
        double* part = prev_bn >= 0 ? w.bnb[prev_bn].part : nullptr;
        DMB_TRY(B.dgrad_layer(la, Ga, lh, lw, g_h[i], &hin, g_cur, part, prev_bn >= 0 ? stat_src : nullptr, &nb));
        g_cur = g_h[i];
        nb_cur = nb;
    }
    *g_out = g_cur;
    *nb_out = nb_cur;
    return 0;
}

int recon_grad(Ctx& c, const float* x, const float* mask, int mask_c, const float* cvar, const float* decoded,
               float scale) {
    const dmb_model& m = c.L.m;
    const int64_t nrec = c.B * (int64_t)m.num_inputs * m.height * m.width;
    const int64_t total4 = nrec / 4;
    int64_t blocks = (total4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    DMB_LAUNCH((recon_grad_kernel), (unsigned)blocks, 256, 0, c.st, decoded, x, mask, mask_c, cvar, total4, m.num_inputs, m.height * m.width / 4, scale / (float)nrec, c.w.gd);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int run_backward_z16(Ctx& c, const float* params, const float* x, const float* mask, int mask_c,
                     const float* cvar, const float* decoded, float grad_scale, float* grads, const float* g_tm) {
    const Layout& L = c.L;
    const dmb_model& m = L.m;
    Workspace& w = c.w;
    cudaStream_t st = c.st;
    Bwd B{c, params, grads, st};
    DMB_TRY(side_stream(&B.side));
    const int H = m.height, W = m.width, lh = L.lh, lw = L.lw;
    const int h = m.num_hiddens, h2 = h / 2, h4 = h / 4;
    int nb = 0;

    Act t3; t3.p = w.t3; Act t2; t2.p = w.t2; Act t1; t1.p = w.t1; Act za; za.p = w.za;
    GradT G;
    {   // the codebook gradient needs nothing but the forward's z_before and indices: first job of the side stream,
        // beside the decoder's data-gradient chain
        cudaStream_t sw;
        DMB_TRY(B.wgrad_stream(&sw));
        DMB_TRY(vq_codebook_grad_only(w.zb, params + L.codebook_off, w.idx, grad_scale * m.weight_commitment, c.B, h,
                                      lh * lw, m.num_embeddings, grads + L.codebook_off, w.vq_part, w.vq_part_rows, sw));
    }
    if (dec_tail_on(m)) {
        // 0-1. reconstruction-loss gradient, dec.6 (1x1) weight / bias / data gradients and dec.4's bias gradient in
        // one pass over the full-resolution tensors (the loss gradient itself is never stored)
        const ConvL& l = L.convs[L.d3];
        DecTailArgs t{};
        t.B = c.B; t.cm = l.cin; t.ni = l.cout; t.hw = H * W;
        t.t3 = w.t3; t.x = x; t.mask = mask; t.mask_c = mask_c; t.cvar = cvar; t.w = c.packed + l.pw_off;
        t.decoded = const_cast<float*>(decoded); t.g_t3 = w.g_t3;
        t.scale = grad_scale * m.weight_recon / (float)(c.B * (int64_t)m.num_inputs * H * W);
        DMB_CHECK(dec_tail_partial_doubles(c.B, H * W, t.cm, t.ni) <= (int64_t)c.B * (H / 2) * (L.max_c > 2 ? L.max_c : 2) * 2,
                  "decoder tail: partial scratch too small");
        t.partials = w.bias_part; t.ticket = reinterpret_cast<unsigned*>(w.recon_sum + 2);
        t.dw = grads + l.w_off; t.db = grads + l.b_off; t.db_prev = grads + L.convs[L.d2].b_off;
        DMB_TRY(dec_tail_backward(t, st));
    } else {
        // 0. reconstruction-loss gradient
        DMB_TRY(recon_grad(c, x, mask, mask_c, cvar, decoded, grad_scale * m.weight_recon));
        // 1. dec.6 conv1x1
        G.g = w.gd;
        DMB_TRY(B.wgrad_layer(L.d3, G, t3, false, H, W, true));
        DMB_TRY(B.dgrad_layer(L.d3, G, H, W, w.g_t3, &t3, nullptr, w.bias_part, nullptr, &nb));
        DMB_TRY(sum_partials(w.bias_part, (int)c.B, nb, h4, grads + L.convs[L.d2].b_off, st));
    }
    // 2-4. decoder (no BatchNorm): dec.4 / dec.2 / dec.0 ConvT
    G = GradT(); G.g = w.g_t3;
    DMB_TRY(B.wgrad_layer(L.d2, G, t2, false, H / 2, W / 2, false));
    DMB_TRY(B.dgrad_layer(L.d2, G, H / 2, W / 2, w.g_t2, &t2, nullptr, w.bias_part, nullptr, &nb));
    DMB_TRY(sum_partials(w.bias_part, B.stat_rows > 0 ? 1 : (int)c.B, B.stat_rows > 0 ? B.stat_rows : nb, h4,
                         grads + L.convs[L.d1].b_off, st));
    G = GradT(); G.g = w.g_t2;
    DMB_TRY(B.wgrad_layer(L.d1, G, t1, false, H / 4, W / 4, false));
    DMB_TRY(B.dgrad_layer(L.d1, G, H / 4, W / 4, w.g_t1, &t1, nullptr, w.bias_part, nullptr, &nb));
    DMB_TRY(sum_partials(w.bias_part, B.stat_rows > 0 ? 1 : (int)c.B, B.stat_rows > 0 ? B.stat_rows : nb, h2,
                         grads + L.convs[L.d0].b_off, st));
    G = GradT(); G.g = w.g_t1;
    DMB_TRY(B.wgrad_layer(L.d0, G, za, false, lh, lw, false));
    DMB_TRY(B.dgrad_layer(L.d0, G, lh, lw, w.g_za, nullptr, nullptr, nullptr, nullptr, nullptr));

    // 5. quantiser: straight-through + commitment term; codebook gradient; BN sums for the last residual layer
    const int nres = (int)L.enc_res.size();
    DMB_CHECK(nres >= 1, "training needs num_residual_layers >= 1");
    const int P = lh * lw;
    {
        const ConvL& lb = L.convs[L.enc_res[nres - 1].b];
        DMB_TRY(vq_backward_stats(w.zb, params + L.codebook_off, w.idx, w.g_za, g_tm, grad_scale * m.weight_commitment,
                                  m.commitment_cost, c.B, h, P, m.num_embeddings, w.g_zb, nullptr,
                                  w.bnb[lb.bn].part, w.erb[nres - 1], w.vq_part, w.vq_part_rows, st));
    }
    // 6. residual layers, last to first
    const float* g_cur = nullptr;
    int nb_cur = 0;
    Act y4a; y4a.p = w.y4; y4a.s = w.bn[L.convs[L.e4].bn].scale; y4a.t = w.bn[L.convs[L.e4].bn].shift;
    DMB_TRY(res_backward(B, L.enc_res, w.era, w.erb, w.ehs, w.g_era, w.g_eh, y4a, L.convs[L.e4].bn, w.y4,
                         w.g_zb, P / 128, lh, lw, &g_cur, &nb_cur));
    // 7. enc.10 (3x3) behind enc.11 BN
    GradT G4, G3, G2, G1;
    DMB_TRY(B.bn_bwd(L.e4, nb_cur, P, &G4, g_cur, w.y4));
    Act y3a; y3a.p = w.y3; y3a.s = w.bn[L.convs[L.e3].bn].scale; y3a.t = w.bn[L.convs[L.e3].bn].shift;
    DMB_TRY(B.wgrad_layer(L.e4, G4, y3a, true, lh, lw, true));
    DMB_TRY(B.dgrad_layer(L.e4, G4, lh, lw, w.g_y3, &y3a, nullptr, w.bnb[L.convs[L.e3].bn].part, w.y3, &nb));
    // 8. enc.7 (4x4 s2) behind enc.8 BN
    DMB_TRY(B.bn_bwd(L.e3, nb, P, &G3, w.g_y3, w.y3));
    Act y2a; y2a.p = w.y2; y2a.s = w.bn[L.convs[L.e2].bn].scale; y2a.t = w.bn[L.convs[L.e2].bn].shift;
    DMB_TRY(B.wgrad_layer(L.e3, G3, y2a, true, H / 4, W / 4, true));
    DMB_TRY(B.dgrad_layer(L.e3, G3, H / 4, W / 4, w.g_y2, &y2a, nullptr, w.bnb[L.convs[L.e2].bn].part, w.y2, &nb));
    // 9. enc.4 (4x4 s2) behind enc.5 BN
    DMB_TRY(B.bn_bwd(L.e2, nb, (int64_t)(H / 4) * (W / 4), &G2, w.g_y2, w.y2));
    Act y1a; y1a.p = w.y1; y1a.s = w.bn[L.convs[L.e1].bn].scale; y1a.t = w.bn[L.convs[L.e1].bn].shift;
    DMB_TRY(B.wgrad_layer(L.e2, G2, y1a, true, H / 2, W / 2, true));
    DMB_TRY(B.dgrad_layer(L.e2, G2, H / 2, W / 2, w.g_y1, &y1a, nullptr, w.bnb[L.convs[L.e1].bn].part, w.y1, &nb));
    // (fold every layer queued so far while the head's inputs are still being produced: the last fold is then the head's alone)
    DMB_TRY(B.flush_wgrad());
    // 10. composite head (enc.0 1x1 + enc.1 4x4 s2) behind enc.2 BN: gradient of the effective conv, then chain rule
    DMB_TRY(B.bn_bwd(L.e1, nb, (int64_t)(H / 2) * (W / 2), &G1, w.g_y1, w.y1));
    {
        const ConvL& l = L.convs[L.e1];
        WgradArgs a{};
        a.B = (int)c.B; a.ks = 4; a.stride = 2; a.partials = w.wg_part;
        a.g = G1.g; a.y = G1.y; a.ga = G1.A; a.gb = G1.Bc; a.gc = G1.Cc; a.g_per_sample = c.per_sample();
        a.x = x; a.ones_channel = 1;
        a.Cin = l.cin; a.H = H; a.W = W; a.Cout = l.cout; a.Ho = H / 2; a.Wo = W / 2;
        int ncta = 0;
        const int pf = wgrad_partial_floats(a, &ncta);
        DMB_CHECK(pf > 0, "no weight-gradient plan for the head");
        WgReduceQueue* q = nullptr;
        DMB_TRY(B.wgrad_scratch((size_t)pf * ncta, &a.partials, &q));
        cudaStream_t sw;
        DMB_TRY(B.wgrad_stream(&sw));
        DMB_TRY(wgrad(a, nullptr, nullptr, w.dweff, sw, q));
        DMB_TRY(B.flush_wgrad());        // every layer's partials -> gradients, one launch
        DMB_TRY(composite_chain(w.dweff, params + l.w0_off, params + l.b0_off, params + l.w_off, l.cin, l.cmid,
                                grads + l.w0_off, grads + l.b0_off, grads + l.w_off, grads + l.b_off, sw));
    }
    DMB_TRY(B.flush_wgrad());
    return B.join();
}

// vae.py:401-414.  enc: conv4s2 -> BN -> ReLU -> conv4s2 -> BN -> ResidualBlock;  dec: ResidualBlock -> ConvT ->
// BN -> ReLU -> ConvT.  total = recon + commitment (vae.py:440).
int run_backward_z32(Ctx& c, const float* params, const float* x, const float* mask, int mask_c,
                     const float* cvar, const float* decoded, float grad_scale, float* grads, const float* g_tm) {
    const Layout& L = c.L;
    const dmb_model& m = L.m;
    Workspace& w = c.w;
    cudaStream_t st = c.st;
    Bwd B{c, params, grads, st};
    DMB_TRY(side_stream(&B.side));
    const int H = m.height, W = m.width, lh = L.lh, lw = L.lw;
    const int h = m.num_hiddens;
    const int P = lh * lw;
    const int nres_e = (int)L.enc_res.size(), nres_d = (int)L.dec_res.size();
    DMB_CHECK(nres_e >= 1 && nres_d >= 1, "training needs num_residual_layers >= 1");
    int nb = 0;

    DMB_TRY(recon_grad(c, x, mask, mask_c, cvar, decoded, grad_scale));
    // dec.4 ConvT (h/2 -> ni), input relu(bn(t1)); its bias gradient is the per-channel sum of the loss gradient
    DMB_LAUNCH((channel_sum_kernel), m.num_inputs, 256, 0, st, w.gd, c.B, m.num_inputs, H * W, grads + L.convs[L.d1].b_off);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    const int bn_d = L.convs[L.d0].bn;
    Act t1a; t1a.p = w.t1; t1a.s = w.bn[bn_d].scale; t1a.t = w.bn[bn_d].shift;
    GradT G; G.g = w.gd;
    DMB_TRY(B.wgrad_layer(L.d1, G, t1a, true, H / 2, W / 2, false));
    DMB_TRY(B.dgrad_layer(L.d1, G, H / 2, W / 2, w.g_t1, &t1a, nullptr, w.bnb[bn_d].part, w.t1, &nb));
    // dec.1 ConvT (h -> h/2) behind dec.2 BN; input = decoder residual block output
    GradT Gd;
    DMB_TRY(B.bn_bwd(L.d0, nb, (int64_t)(H / 2) * (W / 2), &Gd, w.g_t1, w.t1));
    Act hD; hD.p = w.dhs[nres_d - 1];
    DMB_TRY(B.wgrad_layer(L.d0, Gd, hD, false, lh, lw, false));
    DMB_CUDA(cudaMemsetAsync(grads + L.convs[L.d0].b_off, 0, sizeof(float) * L.convs[L.d0].cout, st));   // exactly 0: BN follows
    {
        const ConvL& lb = L.convs[L.dec_res[nres_d - 1].b];
        DMB_TRY(B.dgrad_layer(L.d0, Gd, lh, lw, w.g_za, nullptr, nullptr, w.bnb[lb.bn].part, w.drb[nres_d - 1], &nb));
    }
    // decoder residual block (input: the quantised latent, no BN before it)
    Act zaA; zaA.p = w.za;
    const float* g_cur = nullptr;
    int nb_cur = 0;
    DMB_TRY(res_backward(B, L.dec_res, w.dra, w.drb, w.dhs, w.g_dra, w.g_dh, zaA, -1, nullptr, w.g_za, nb, lh, lw,
                         &g_cur, &nb_cur));
    // quantiser
    {
        const ConvL& lb = L.convs[L.enc_res[nres_e - 1].b];
        DMB_TRY(vq_backward_stats(w.zb, params + L.codebook_off, w.idx, g_cur, g_tm, grad_scale, m.commitment_cost, c.B, h, P,
                                  m.num_embeddings, w.g_zb, grads + L.codebook_off, w.bnb[lb.bn].part,
                                  w.erb[nres_e - 1], w.vq_part, w.vq_part_rows, st));
    }
    // encoder residual block (input: bn(y2), pending affine)
    const int bn2 = L.convs[L.e2].bn, bn1 = L.convs[L.e1].bn;
    Act y2a; y2a.p = w.y2; y2a.s = w.bn[bn2].scale; y2a.t = w.bn[bn2].shift;
    DMB_TRY(res_backward(B, L.enc_res, w.era, w.erb, w.ehs, w.g_era, w.g_eh, y2a, bn2, w.y2, w.g_zb, P / 128, lh, lw,
                         &g_cur, &nb_cur));
    // enc.3 conv4x4s2 (h/2 -> h) behind enc.4 BN; input relu(bn(y1))
    GradT G2, G1;
    DMB_TRY(B.bn_bwd(L.e2, nb_cur, P, &G2, g_cur, w.y2));
    Act y1a; y1a.p = w.y1; y1a.s = w.bn[bn1].scale; y1a.t = w.bn[bn1].shift;
    DMB_TRY(B.wgrad_layer(L.e2, G2, y1a, true, H / 2, W / 2, true));
    DMB_TRY(B.dgrad_layer(L.e2, G2, H / 2, W / 2, w.g_y1, &y1a, nullptr, w.bnb[bn1].part, w.y1, &nb));
    // enc.0 conv4x4s2 (ni -> h/2) behind enc.1 BN; input x
    DMB_TRY(B.bn_bwd(L.e1, nb, (int64_t)(H / 2) * (W / 2), &G1, w.g_y1, w.y1));
    Act xin; xin.p = x;
    DMB_TRY(B.wgrad_layer(L.e1, G1, xin, false, H, W, true));
    DMB_TRY(B.flush_wgrad());
    return B.join();
}

}  // namespace
}  // namespace dmb

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
using namespace dmb;

extern "C" {

int dmb_abi_version(void) { return DMB_ABI_VERSION; }
const char* dmb_last_error(void) { return dmb::g_err; }

int dmb_param_count(const dmb_model* m, int64_t* n_params, int64_t* n_bnbuf, int32_t* n_bn) {
    Layout L;
    DMB_TRY(build_layout(m, L));
    if (n_params) *n_params = L.n_params;
    if (n_bnbuf) *n_bnbuf = L.n_bnbuf;
    if (n_bn) *n_bn = (int32_t)L.bns.size();
    return 0;
}

int dmb_param_lookup(const dmb_model* m, const char* key_host, int32_t* which, int64_t* offset, int64_t* numel) {
    Layout L;
    DMB_TRY(build_layout(m, L));
    DMB_CHECK(key_host != nullptr, "null key");
    for (const Entry& e : L.entries)
        if (e.key == key_host) {
            if (which) *which = e.which;
            if (offset) *offset = e.off;
            if (numel) *numel = e.numel;
            return 0;
        }
    DMB_CHECK(false, "no such state_dict key: %s", key_host);
}

int dmb_latent_shape(const dmb_model* m, int32_t* d, int32_t* lh, int32_t* lw) {
    Layout L;
    DMB_TRY(build_layout(m, L));
    if (d) *d = L.D;
    if (lh) *lh = L.lh;
    if (lw) *lw = L.lw;
    return 0;
}

int dmb_packed_floats(const dmb_model* m, int64_t* n) {
    Layout L;
    DMB_TRY(build_layout(m, L));
    DMB_CHECK(n != nullptr, "null output");
    *n = L.n_packed;
    return 0;
}

int dmb_pack_weights(const dmb_model* m, const float* params, const float* bnbuf, int32_t bn_mode,
                     float* packed, void* stream) {
    Layout L;
    DMB_TRY(build_layout(m, L));
    DMB_CHECK(params && bnbuf && packed, "dmb_pack_weights: null pointer");
    DMB_CHECK(bn_mode >= 0 && bn_mode <= 2, "bad bn_mode %d", bn_mode);
    return pack_weights(L, params, bnbuf, bn_mode, packed, (cudaStream_t)stream);
}

int dmb_workspace_bytes(const dmb_model* m, int64_t batch, int32_t bn_mode, int32_t keep, size_t* bytes) {
    Layout L;
    DMB_TRY(build_layout(m, L));
    DMB_CHECK(batch > 0 && bytes, "dmb_workspace_bytes: bad arguments");
    Workspace w;
    DMB_TRY(carve_workspace(L, batch, bn_mode, keep, nullptr, w));
    *bytes = w.bytes;
    return 0;
}

static int prep(const dmb_model* m, int64_t batch, int32_t bn_mode, int keep, void* workspace,
                size_t workspace_bytes, Layout& L, Workspace& w) {
    DMB_TRY(build_layout(m, L));
    DMB_CHECK(batch > 0 && batch < (1 << 24), "batch %lld out of range", (long long)batch);
    DMB_CHECK(bn_mode >= 0 && bn_mode <= 2, "bad bn_mode %d", bn_mode);
    DMB_CHECK(workspace != nullptr, "null workspace");
    DMB_TRY(carve_workspace(L, batch, bn_mode, keep, workspace, w));
    DMB_CHECK(w.bytes <= workspace_bytes, "workspace too small: need %zu bytes, have %zu", w.bytes, workspace_bytes);
    return 0;
}

int dmb_encoder_forward(const dmb_model* m, const float* packed, const float* x, int64_t batch,
                        int32_t bn_mode, float* z_before, float* bnbuf_inout, void* workspace,
                        size_t workspace_bytes, void* stream) {
    Layout L; Workspace w;
    DMB_CHECK(packed && x && z_before, "dmb_encoder_forward: null pointer");
    // keep flag only affects the tail of the carve; probe which one the caller sized for
    DMB_TRY(prep(m, batch, bn_mode, 0, workspace, workspace_bytes, L, w));
    Ctx c{L, packed, w, batch, bn_mode, bnbuf_inout, (cudaStream_t)stream};
    return run_encoder(c, x, z_before, nullptr);
}

int dmb_encode(const dmb_model* m, const float* packed, const float* codebook, const float* x,
               int64_t batch, int32_t bn_mode, float* z_before, float* z_after, int32_t* idx,
               double* vq_stats, void* workspace, size_t workspace_bytes, void* stream) {
    Layout L; Workspace w;
    DMB_CHECK(packed && codebook && x, "dmb_encode: null pointer");
    DMB_CHECK(bn_mode != DMB_BN_BATCH, "dmb_encode: BATCH statistics make patches interdependent; "
              "use EVAL or PER_SAMPLE for bulk encoding");
    DMB_TRY(prep(m, batch, bn_mode, 0, workspace, workspace_bytes, L, w));
    Ctx c{L, packed, w, batch, bn_mode, nullptr, (cudaStream_t)stream};
    float* zb = z_before ? z_before : w.zb;
    Pending pend;
    DMB_TRY(run_encoder(c, x, zb, bn_mode == DMB_BN_EVAL ? nullptr : &pend));
    return run_vq(c, codebook, zb, &pend, zb, z_after, idx, vq_stats);
}

int dmb_decoder_forward(const dmb_model* m, const float* packed, const float* z_after, int64_t batch,
                        int32_t bn_mode, float* decoded, float* bnbuf_inout, void* workspace,
                        size_t workspace_bytes, void* stream) {
    Layout L; Workspace w;
    DMB_CHECK(packed && z_after && decoded, "dmb_decoder_forward: null pointer");
    DMB_TRY(prep(m, batch, bn_mode, 1, workspace, workspace_bytes, L, w));
    Ctx c{L, packed, w, batch, bn_mode, bnbuf_inout, (cudaStream_t)stream};
    return run_decoder(c, z_after, decoded);
}

int dmb_conv2d_forward(const float* x, const float* w_packed, const float* bias, float* y,
                       int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout, int32_t ksize,
                       int32_t stride, const float* in_scale, const float* in_shift,
                       int32_t in_per_sample, int32_t in_relu, const float* skip, int32_t out_relu,
                       void* stream) {
    DMB_CHECK(x && w_packed && bias && y, "dmb_conv2d_forward: null pointer");
    DMB_CHECK((in_scale == nullptr) == (in_shift == nullptr), "dmb_conv2d_forward: scale/shift must come together");
    DMB_CHECK(stride >= 1 && h % stride == 0 && w % stride == 0, "dmb_conv2d_forward: bad geometry");
    ConvFwdArgs a{};
    a.x = x; a.y = y; a.w = w_packed; a.bias = bias; a.in_scale = in_scale; a.in_shift = in_shift;
    a.in_per_sample = in_per_sample; a.in_relu = in_relu; a.skip = skip; a.out_relu = out_relu;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = ksize; a.stride = stride;
    a.Ho = h / stride; a.Wo = w / stride;
    return conv_fwd(a, (cudaStream_t)stream);
}

int dmb_conv2d_tc_scratch_floats(int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout, int32_t ksize,
                                 int64_t* floats) {
    DMB_CHECK(floats && batch > 0, "dmb_conv2d_tc_scratch_floats: bad arguments");
    *floats = ((batch * cin * h * w + 63) & ~63ll) + conv_tc_weight_floats(cin, cout, ksize);
    return 0;
}

int dmb_conv2d_tc(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                  int32_t h, int32_t w, int32_t cout, int32_t ksize, int32_t stride, int32_t in_relu,
                  const float* skip, int32_t out_relu, int32_t nhwc_io, float* scratch, void* stream) {
    DMB_CHECK(x && w_packed && bias && y && scratch, "dmb_conv2d_tc: null pointer");
    DMB_CHECK(conv_tc_supported(cin, cout, ksize, stride, h, w), "dmb_conv2d_tc: layer %dx%d s%d %d->%d @%dx%d is not "
              "a tensor-core shape (Cin %% 32, Cout in {32,64}, output width in {8..128})", ksize, ksize, stride, cin,
              cout, h, w);
    cudaStream_t st = (cudaStream_t)stream;
    float* xt = scratch;
    float* wtc = scratch + ((batch * cin * h * w + 63) & ~63ll);
    DMB_TRY(pack_tc_weights(w_packed, wtc, cin, cout, ksize, st));
    const float* xin = x;
    if (!nhwc_io) { DMB_TRY(nchw_to_nhwc(x, xt, batch, cin, h * w, st)); xin = xt; }
    ConvTcArgs a{};
    a.x = xin; a.wtc = wtc; a.bias = bias; a.y = y; a.skip = skip;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = ksize; a.stride = stride;
    a.in_relu = in_relu; a.out_relu = out_relu; a.out_nhwc = nhwc_io; a.skip_nhwc = nhwc_io;
    return conv_tc(a, st);
}

int dmb_conv2d_wino(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                    int32_t h, int32_t w, int32_t cout, int32_t in_relu, int32_t out_relu, const float* w2_packed,
                    const float* bias2, float* y2, float* scratch, void* stream) {
    DMB_CHECK(x && w_packed && bias && (y || y2) && scratch, "dmb_conv2d_wino: null pointer");
    DMB_CHECK(conv_wino_supported(cin, cout, 3, 1, h, w), "dmb_conv2d_wino: only 3x3 stride-1 layers with 16 input and "
              "16 or 32 output channels on 16x16 maps (got %d->%d @%dx%d)", cin, cout, h, w);
    DMB_CHECK((w2_packed == nullptr) == (y2 == nullptr) && (w2_packed == nullptr) == (bias2 == nullptr),
              "dmb_conv2d_wino: w2_packed, bias2 and y2 come together");
    cudaStream_t st = (cudaStream_t)stream;
    DMB_TRY(pack_wino_weights(w_packed, scratch, cin, cout, st));
    ConvWinoArgs a{};
    a.x = x; a.u = scratch; a.bias = bias; a.y = y; a.B = (int)batch; a.Cout = cout; a.in_relu = in_relu; a.out_relu = out_relu;
    a.w2 = w2_packed; a.bias2 = bias2; a.y2 = y2;
    return conv_wino(a, st);
}

int dmb_conv2d_tm_scratch_floats(int32_t cin, int32_t cout, int32_t ksize, int64_t* floats) {
    DMB_CHECK(floats != nullptr, "dmb_conv2d_tm_scratch_floats: null output");
    *floats = conv_tm_weight_floats(cin, cout, ksize);
    return 0;
}

int dmb_conv2d_tm(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                  int32_t h, int32_t w, int32_t cout, int32_t ksize, int32_t stride, int32_t in_relu, const float* skip,
                  int32_t out_relu, float* scratch, void* stream) {
    DMB_CHECK(x && w_packed && bias && y && scratch, "dmb_conv2d_tm: null pointer");
    DMB_CHECK(conv_tm_supported(cin, cout, ksize, stride, h, w), "dmb_conv2d_tm: layer %dx%d s%d %d->%d @%dx%d is not one "
              "of the thin encoder shapes this kernel is built for", ksize, ksize, stride, cin, cout, h, w);
    cudaStream_t st = (cudaStream_t)stream;
    DMB_TRY(pack_tm_weights(w_packed, scratch, cin, cout, ksize, st));
    ConvTmArgs a{};
    a.x = x; a.wtm = scratch; a.bias = bias; a.y = y; a.skip = skip;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = ksize; a.stride = stride;
    a.in_relu = in_relu; a.out_relu = out_relu;
    return conv_tm(a, st);
}

int dmb_conv2d_tm_bn(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                     int32_t h, int32_t w, int32_t cout, int32_t ksize, int32_t stride, const float* in_scale,
                     const float* in_shift, int32_t in_per_sample, int32_t in_relu, double* stats, int32_t* bands,
                     float* scratch, void* stream) {
    DMB_CHECK(x && w_packed && bias && y && scratch, "dmb_conv2d_tm_bn: null pointer");
    DMB_CHECK(conv_tm_supported(cin, cout, ksize, stride, h, w), "dmb_conv2d_tm_bn: layer %dx%d s%d %d->%d @%dx%d is not one "
              "of the thin encoder shapes this kernel is built for", ksize, ksize, stride, cin, cout, h, w);
    if (bands) *bands = conv_tm_bands(cin, cout, ksize, stride, h, w);
    cudaStream_t st = (cudaStream_t)stream;
    DMB_TRY(pack_tm_weights(w_packed, scratch, cin, cout, ksize, st));
    ConvTmArgs a{};
    a.x = x; a.wtm = scratch; a.bias = bias; a.y = y;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = ksize; a.stride = stride;
    a.bn = 1; a.in_scale = in_scale; a.in_shift = in_shift; a.in_per_sample = in_per_sample; a.in_relu = in_relu;
    a.stats = stats;
    return conv_tm(a, st);
}

int dmb_conv2d_tm_batch_stat_rows(int32_t* rows) {
    DMB_CHECK(rows != nullptr, "dmb_conv2d_tm_batch_stat_rows: null output");
    *rows = TM_BATCH_ROWS_MAX;
    return 0;
}

int dmb_conv2d_tm_dgrad(const float* gy, const float* w_packed, float* gx, int64_t batch, int32_t cin, int32_t h, int32_t w,
                        int32_t cout, int32_t ksize, int32_t stride, const float* y_raw, const float* ga, const float* gb,
                        const float* gc, const float* mask_src, const float* mask_scale,
                        const float* mask_shift, const float* skip, double* stats, const float* stat_src,
                        int32_t* stat_rows, float* scratch, void* stream) {
    DMB_CHECK(gy && w_packed && gx && scratch, "dmb_conv2d_tm_dgrad: null pointer");
    DMB_CHECK(conv_tm_dg_supported(cin, cout, ksize, stride, h, w), "dmb_conv2d_tm_dgrad: %dx%d s%d %d->%d @%dx%d is not one "
              "of the data-gradient shapes this kernel is built for", ksize, ksize, stride, cin, cout, h, w);
    DMB_CHECK(!stats || stat_rows, "dmb_conv2d_tm_dgrad: stats needs stat_rows");
    DMB_CHECK((ga != nullptr) == (y_raw != nullptr) && (ga != nullptr) == (gb != nullptr) && (ga != nullptr) == (gc != nullptr),
              "dmb_conv2d_tm_dgrad: y_raw / ga / gb / gc come together");
    DMB_CHECK(!ga || conv_tm_dg_dual(cin, cout, ksize, stride, h, w), "dmb_conv2d_tm_dgrad: this shape has no BatchNorm-"
              "backward transform on load");
    cudaStream_t st = (cudaStream_t)stream;
    DMB_TRY(pack_tm_weights(w_packed, scratch, cin, cout, ksize, st));
    const int64_t wf = (conv_tm_weight_floats(cin, cout, ksize) + 63) & ~63ll;
    float* zero = scratch + wf;
    DMB_CUDA(cudaMemsetAsync(zero, 0, sizeof(float) * cout, st));
    ConvTmArgs a{};
    a.x = gy; a.wtm = scratch; a.bias = zero; a.y = gx; a.skip = skip;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = ksize; a.stride = stride;
    a.dg = 1; a.mask_src = mask_src; a.mask_s = mask_scale; a.mask_t = mask_shift;
    a.stats = stats; a.stat_src = stat_src; a.stats_batch = 1; a.stat_rows = stat_rows;
    a.x2 = y_raw; a.in_a = ga; a.in_b = gb; a.in_c = gc;
    return conv_tm(a, st);
}

int dmb_conv_transpose2d_tm(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                            int32_t h, int32_t w, int32_t cout, int32_t out_relu, int32_t data_gradient, const float* y_raw,
                            const float* ga, const float* gb, const float* gc, const float* mask_src,
                            const float* mask_scale, const float* mask_shift, double* stats, const float* stat_src,
                            int32_t* stat_rows, float* scratch, void* stream) {
    DMB_CHECK(x && w_packed && y && scratch, "dmb_conv_transpose2d_tm: null pointer");
    DMB_CHECK(conv_tm_ct_supported(cin, cout, h, w, data_gradient != 0), "dmb_conv_transpose2d_tm: %d->%d @%dx%d%s is not one of "
              "the shapes this kernel is built for", cin, cout, h, w, data_gradient ? " (data gradient)" : "");
    DMB_CHECK(data_gradient || (!y_raw && !ga && !gb && !gc && !mask_src && !mask_scale && !mask_shift && !stats && !stat_src),
              "dmb_conv_transpose2d_tm: the BatchNorm-backward load, gate and sums belong to the data-gradient form");
    DMB_CHECK((ga != nullptr) == (y_raw != nullptr) && (ga != nullptr) == (gb != nullptr) && (ga != nullptr) == (gc != nullptr),
              "dmb_conv_transpose2d_tm: y_raw / ga / gb / gc come together");
    DMB_CHECK(!stats || stat_rows, "dmb_conv_transpose2d_tm: stats needs stat_rows");
    cudaStream_t st = (cudaStream_t)stream;
    DMB_TRY(pack_tm_weights_ct(w_packed, scratch, cin, cout, st));
    const int64_t wf = (conv_tm_weight_floats(cin, 4 * cout, 3) + 63) & ~63ll;
    float* zero = scratch + wf;
    if (!bias) DMB_CUDA(cudaMemsetAsync(zero, 0, sizeof(float) * cout, st));
    ConvTmArgs a{};
    a.x = x; a.wtm = scratch; a.bias = bias ? bias : zero; a.y = y;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = 4; a.stride = 2;
    a.ct = 1; a.dg = data_gradient ? 1 : 0; a.out_relu = out_relu;
    a.x2 = y_raw; a.in_a = ga; a.in_b = gb; a.in_c = gc;
    a.mask_src = mask_src; a.mask_s = mask_scale; a.mask_t = mask_shift;
    a.stats = stats; a.stat_src = stat_src; a.stats_batch = stats ? 1 : 0; a.stat_rows = stat_rows;
    return conv_tm(a, st);
}

int dmb_residual_layer_tm_scratch_floats(int64_t* floats) {
    DMB_CHECK(floats != nullptr, "dmb_residual_layer_tm_scratch_floats: null output");
    *floats = conv_tm_weight_floats(16, 32, 3) + conv_tm_weight_floats(32, 16, 1);
    return 0;
}

int dmb_residual_layer_tm(const float* x, const float* w1_packed, const float* bias1, const float* w2_packed,
                          const float* bias2, float* y, int64_t batch, float* scratch, void* stream) {
    DMB_CHECK(x && w1_packed && bias1 && w2_packed && bias2 && y && scratch, "dmb_residual_layer_tm: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DMB_TRY(pack_tm_weights(w1_packed, scratch, 16, 32, 3, st));
    DMB_TRY(pack_tm_weights(w2_packed, scratch + conv_tm_weight_floats(16, 32, 3), 32, 16, 1, st));
    ConvTmArgs a{};
    a.x = x; a.wtm = scratch; a.bias = bias1; a.y = y; a.skip = nullptr; a.bias2 = bias2;
    a.B = (int)batch; a.Cin = 16; a.H = 16; a.W = 16; a.Cout = 32; a.ks = 3; a.stride = 1;
    a.in_relu = 1; a.out_relu = 1;
    return conv_tm(a, st);
}

static void wgrad_args(WgradArgs& a, const float* x, const float* gy, int64_t batch, int32_t cin, int32_t h, int32_t w,
                       int32_t cout, int32_t ksize, int32_t stride) {
    a.x = x; a.g = gy; a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout; a.ks = ksize; a.stride = stride;
    a.Ho = h / stride; a.Wo = w / stride;
}

int dmb_conv2d_weight_grad_scratch_floats(int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout, int32_t ksize,
                                          int32_t stride, int64_t* floats) {
    DMB_CHECK(floats && batch > 0 && stride > 0, "dmb_conv2d_weight_grad_scratch_floats: bad arguments");
    WgradArgs a{};
    wgrad_args(a, nullptr, nullptr, batch, cin, h, w, cout, ksize, stride);
    int ncta = 0;
    const int pf = wgrad_partial_floats(a, &ncta);
    DMB_CHECK(pf > 0, "dmb_conv2d_weight_grad: unsupported layer %dx%d s%d %d->%d", ksize, ksize, stride, cin, cout);
    *floats = (int64_t)pf * ncta;
    return 0;
}

int dmb_conv2d_weight_grad(const float* x, const float* gy, float* dw, float* db, int64_t batch, int32_t cin, int32_t h,
                           int32_t w, int32_t cout, int32_t ksize, int32_t stride, const float* x_scale,
                           const float* x_shift, int32_t x_relu, const float* y_raw, const float* ga, const float* gb,
                           const float* gc, float* scratch, void* stream) {
    DMB_CHECK(x && gy && dw && scratch && batch > 0, "dmb_conv2d_weight_grad: null pointer / empty batch");
    DMB_CHECK((x_scale == nullptr) == (x_shift == nullptr), "dmb_conv2d_weight_grad: x_scale / x_shift come together");
    DMB_CHECK((ga == nullptr) == (gc == nullptr) && (y_raw == nullptr) == (gb == nullptr) && (ga || !y_raw),
              "dmb_conv2d_weight_grad: ga / gc come together, y_raw / gb come together and need ga");
    WgradArgs a{};
    wgrad_args(a, x, gy, batch, cin, h, w, cout, ksize, stride);
    a.xs = x_scale; a.xt = x_shift; a.x_relu = x_relu; a.y = y_raw; a.ga = ga; a.gb = gb; a.gc = gc;
    a.partials = scratch;
    return wgrad(a, dw, db, nullptr, (cudaStream_t)stream);
}

int dmb_conv_transpose2d_forward(const float* x, const float* w_packed, const float* bias, float* y,
                                 int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout,
                                 const float* in_scale, const float* in_shift, int32_t in_per_sample,
                                 int32_t in_relu, int32_t out_relu, void* stream) {
    DMB_CHECK(x && w_packed && bias && y, "dmb_conv_transpose2d_forward: null pointer");
    ConvTFwdArgs a{};
    a.x = x; a.y = y; a.w = w_packed; a.bias = bias; a.in_scale = in_scale; a.in_shift = in_shift;
    a.in_per_sample = in_per_sample; a.in_relu = in_relu; a.out_relu = out_relu;
    a.B = (int)batch; a.Cin = cin; a.H = h; a.W = w; a.Cout = cout;
    return convt_fwd(a, (cudaStream_t)stream);
}

// ---- stand-alone ResidualBlock.forward (vq_vae.py:212-225): the same run_res schedule over a layout that holds
// nothing but the block's layers
namespace {
struct ResPlan {
    Layout L; std::vector<ResL> res; Workspace w; float* packed = nullptr; size_t bytes = 0;
};
int res_plan(int h, int rh, int nl, int64_t B, int H, int W, int bn_mode, void* base, ResPlan& p) {
    DMB_CHECK(nl >= 0 && nl <= 16, "residual block: %d layers out of range", nl);
    DMB_CHECK(h >= 8 && h % 8 == 0 && rh >= 8 && rh % 8 == 0,
              "residual block: num_hiddens=%d and num_residual_hiddens=%d must be positive multiples of 8", h, rh);
    DMB_CHECK(W % 8 == 0, "residual block: map width %d must be a multiple of 8", W);
    DMB_CHECK(B > 0 && B < (1 << 24) && H > 0 && W > 0, "residual block: bad batch / map size");
    DMB_CHECK(bn_mode >= 0 && bn_mode <= 2, "bad bn_mode %d", bn_mode);
    p.L = Layout();
    p.L.m = dmb_model{};
    p.L.m.num_hiddens = h; p.L.m.num_residual_hiddens = rh; p.L.m.num_residual_layers = nl;
    p.L.m.bn_eps = 1e-5f; p.L.m.bn_momentum = 0.1f;
    Builder bd(p.L);
    bd.res("", h, rh, nl, p.res);
    p.L.pzero_off = bd.take_packed(p.L.max_c);
    p.L.n_params = bd.p; p.L.n_bnbuf = bd.bb; p.L.n_packed = bd.pk;
    p.L.D = h; p.L.lh = H; p.L.lw = W;
    Bump bp(base);
    const int64_t hw = (int64_t)H * W;
    p.packed = bp.take<float>(p.L.n_packed);
    p.w = Workspace();
    for (int i = 0; i < nl; ++i) {
        p.w.era.push_back(bp.take<float>(B * rh * hw));
        p.w.erb.push_back(bp.take<float>(B * h * hw));
        p.w.ehs.push_back(bp.take<float>(B * h * hw));
    }
    p.w.bn.resize(p.L.bns.size());
    if (bn_mode != DMB_BN_EVAL) {
        for (const ConvL& c : p.L.convs) {
            const int nb = conv_fwd_bands(c.ks, c.stride, c.cin, c.cout, H, W, true, B);
            DMB_CHECK(nb > 0, "no launch plan for a %dx%d %d->%d conv on %dx%d maps", c.ks, c.ks, c.cin, c.cout, H, W);
            BnWs& b = p.w.bn[c.bn];
            const int64_t rows = (bn_mode == DMB_BN_PER_SAMPLE) ? B : 1;
            b.nbands = nb; b.count = hw;
            b.part = bp.take<double>(B * nb * c.cout * 2);
            b.scale = bp.take<float>(rows * c.cout); b.shift = bp.take<float>(rows * c.cout);
            b.mean = bp.take<float>(rows * c.cout); b.invstd = bp.take<float>(rows * c.cout);
            b.gsum = nullptr;
        }
    }
    p.bytes = (bp.off + 255) & ~(size_t)255;
    return 0;
}
}  // namespace

int dmb_residual_block_sizes(int32_t num_hiddens, int32_t num_residual_hiddens, int32_t num_residual_layers,
                             int64_t batch, int32_t h, int32_t w, int32_t bn_mode, int64_t* n_params,
                             int64_t* n_bnbuf, size_t* workspace_bytes) {
    ResPlan p;
    DMB_TRY(res_plan(num_hiddens, num_residual_hiddens, num_residual_layers, batch, h, w, bn_mode, nullptr, p));
    if (n_params) *n_params = p.L.n_params;
    if (n_bnbuf) *n_bnbuf = p.L.n_bnbuf;
    if (workspace_bytes) *workspace_bytes = p.bytes;
    return 0;
}

int dmb_residual_block_forward(int32_t num_hiddens, int32_t num_residual_hiddens, int32_t num_residual_layers,
                               const float* params, const float* bnbuf, const float* x, int64_t batch, int32_t h,
                               int32_t w, int32_t bn_mode, float* y, float* bnbuf_inout, void* workspace,
                               size_t workspace_bytes, void* stream) {
    DMB_CHECK(x && y && workspace && (params || num_residual_layers == 0), "dmb_residual_block_forward: null pointer");
    DMB_CHECK(bnbuf || bn_mode != DMB_BN_EVAL || num_residual_layers == 0,
              "dmb_residual_block_forward: EVAL mode needs the running statistics");
    ResPlan p;
    DMB_TRY(res_plan(num_hiddens, num_residual_hiddens, num_residual_layers, batch, h, w, bn_mode, workspace, p));
    DMB_CHECK(p.bytes <= workspace_bytes, "workspace too small: need %zu bytes, have %zu", p.bytes, workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    if (num_residual_layers == 0) {
        DMB_CUDA(cudaMemcpyAsync(y, x, sizeof(float) * batch * num_hiddens * h * w, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    DMB_TRY(pack_weights(p.L, params, bnbuf, bn_mode, p.packed, st));
    Ctx c{p.L, p.packed, p.w, batch, bn_mode, bnbuf_inout, st};
    Act in; in.p = x;
    Act out;
    return run_res(c, p.res, p.w.era, p.w.erb, p.w.ehs, in, h, w, y, nullptr, &out);
}

static int train_forward_impl(const dmb_model* m, const float* packed, const float* params, const float* x,
                              const float* mask, int32_t mask_channels, const float* channel_var, int64_t batch,
                              const dmb_time_matching* tm, int n_losses, float* decoded, float* losses_out,
                              float* bnbuf_inout, void* workspace, size_t workspace_bytes, void* stream,
                              const dmb_sync_bn* sync = nullptr) {
    Layout L; Workspace w;
    DMB_CHECK(packed && params && x && channel_var && decoded && losses_out, "dmb_train_forward: null pointer");
    DMB_TRY(prep(m, batch, DMB_BN_BATCH, 1, workspace, workspace_bytes, L, w));
    cudaStream_t st = (cudaStream_t)stream;
    Ctx c{L, packed, w, batch, DMB_BN_BATCH, bnbuf_inout, st, sync};
    Pending pend;
    DMB_CUDA(cudaMemsetAsync(w.vq_stats, 0, sizeof(double) * (2 + m->num_embeddings), st));
    DMB_CUDA(cudaMemsetAsync(w.recon_sum, 0, sizeof(double) * 4, st));
    c.fin_ticket = reinterpret_cast<unsigned*>(w.recon_sum + 3);       // zeroed above; every user leaves it zero
    DMB_TRY(run_encoder(c, x, w.zb, &pend));
    DMB_TRY(run_vq(c, params + L.codebook_off, w.zb, &pend, w.zb, w.za, w.idx, w.vq_stats));
    DMB_TRY(dmb_vq_finalize(w.vq_stats, L.D, m->num_embeddings, m->commitment_cost, w.scalars, stream));
    if (dec_tail2_on(L)) {
        // dec.4 (ConvTranspose2d + ReLU), dec.6 (1x1) and the reconstruction loss in one pass: t3 is written once
        DMB_TRY(run_decoder(c, w.za, decoded, true, true));
        const ConvL& l4 = L.convs[L.d2];
        const ConvL& l6 = L.convs[L.d3];
        DecTail2Args t{};
        t.B = batch; t.ci = l4.cin; t.cm = l4.cout; t.ni = l6.cout; t.hi = 4 * L.lh; t.wi = 4 * L.lw;
        t.t2 = w.t2; t.w4 = packed + l4.pw_off; t.b4 = packed + l4.pb_off; t.t3 = w.t3;
        t.x = x; t.mask = mask; t.mask_c = mask_channels; t.cvar = channel_var;
        t.w6 = packed + l6.pw_off; t.b6 = packed + l6.pb_off; t.decoded = decoded; t.loss_sum = w.recon_sum;
        DMB_TRY(dec_tail2_forward(t, st));
    } else if (dec_tail_on(*m)) {
        // dec.6 (1x1) and the reconstruction loss in one pass over the full-resolution tensors
        DMB_TRY(run_decoder(c, w.za, decoded, true));
        const ConvL& l = L.convs[L.d3];
        DecTailArgs t{};
        t.B = batch; t.cm = l.cin; t.ni = l.cout; t.hw = m->height * m->width;
        t.t3 = w.t3; t.x = x; t.mask = mask; t.mask_c = mask_channels; t.cvar = channel_var;
        t.w = packed + l.pw_off; t.bias = packed + l.pb_off; t.decoded = decoded; t.loss_sum = w.recon_sum;
        DMB_TRY(dec_tail_forward(t, st));
    } else {
        DMB_TRY(run_decoder(c, w.za, decoded));
        DMB_TRY(dmb_recon_loss(decoded, x, mask, mask_channels, channel_var, batch, m->num_inputs,
                               m->height * m->width, w.recon_sum, stream));
    }
    const bool z32 = m->arch == DMB_ARCH_Z32;
    if (tm) {
        // pair similarities on z_before (vq_vae.py:325, vae.py:322) or, for VQ_VAE_z32, on z_after (vae.py:444)
        DMB_CHECK(w.tm_scratch, "time matching needs a workspace carved with keep_activations=1");
        DMB_TRY(tm_forward(z32 ? w.za : w.zb, batch, (int64_t)L.D * L.lh * L.lw, *tm, w.tm_scratch, w.scalars + 4, st));
    }
    DMB_LAUNCH((train_losses_kernel), 1, 1, 0, st, w.scalars, w.recon_sum, (double)batch * m->num_inputs * m->height * m->width, z32 ? 1.f : m->weight_recon, z32 ? 1.f : m->weight_commitment, tm ? tm->weight : 0.f, tm ? 1 : 0, n_losses, losses_out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int dmb_train_forward(const dmb_model* m, const float* packed, const float* params, const float* x,
                      const float* mask, int32_t mask_channels, const float* channel_var, int64_t batch,
                      float* decoded, float* losses_out, float* bnbuf_inout, void* workspace,
                      size_t workspace_bytes, void* stream) {
    return train_forward_impl(m, packed, params, x, mask, mask_channels, channel_var, batch, nullptr, 4, decoded,
                              losses_out, bnbuf_inout, workspace, workspace_bytes, stream);
}

int dmb_train_forward_tm(const dmb_model* m, const float* packed, const float* params, const float* x,
                         const float* mask, int32_t mask_channels, const float* channel_var, int64_t batch,
                         const dmb_time_matching* tm, float* decoded, float* losses_out, float* bnbuf_inout,
                         void* workspace, size_t workspace_bytes, void* stream) {
    return train_forward_impl(m, packed, params, x, mask, mask_channels, channel_var, batch, tm, 8, decoded,
                              losses_out, bnbuf_inout, workspace, workspace_bytes, stream);
}

int dmb_train_forward_sync(const dmb_model* m, const float* packed, const float* params, const float* x,
                           const float* mask, int32_t mask_channels, const float* channel_var, int64_t batch,
                           const dmb_time_matching* tm, const dmb_sync_bn* sync, float* decoded, float* losses_out,
                           float* bnbuf_inout, void* workspace, size_t workspace_bytes, void* stream) {
    return train_forward_impl(m, packed, params, x, mask, mask_channels, channel_var, batch, tm, 8, decoded,
                              losses_out, bnbuf_inout, workspace, workspace_bytes, stream, sync);
}

int dmb_train_backward_tm(const dmb_model* m, const float* packed, const float* params, const float* x,
                          const float* mask, int32_t mask_channels, const float* channel_var,
                          const float* decoded, int64_t batch, const dmb_time_matching* tm, float grad_scale,
                          float* grads, void* workspace, size_t workspace_bytes, void* stream) {
    return dmb_train_backward_sync(m, packed, params, x, mask, mask_channels, channel_var, decoded, batch, tm, nullptr,
                                   grad_scale, grads, workspace, workspace_bytes, stream);
}

int dmb_train_backward_sync(const dmb_model* m, const float* packed, const float* params, const float* x,
                            const float* mask, int32_t mask_channels, const float* channel_var,
                            const float* decoded, int64_t batch, const dmb_time_matching* tm, const dmb_sync_bn* sync,
                            float grad_scale, float* grads, void* workspace, size_t workspace_bytes, void* stream) {
    Layout L; Workspace w;
    DMB_CHECK(packed && params && x && channel_var && decoded && grads, "dmb_train_backward: null pointer");
    DMB_TRY(prep(m, batch, DMB_BN_BATCH, 1, workspace, workspace_bytes, L, w));
    Ctx c{L, packed, w, batch, DMB_BN_BATCH, nullptr, (cudaStream_t)stream, sync};
    c.fin_ticket = reinterpret_cast<unsigned*>(w.recon_sum + 3);       // zeroed by the forward pass, left zero by every user
    const float* g_tm = nullptr;
    if (tm) {
        // d(weight * tm_loss)/dz; z_after is a straight-through copy of z_before, so either source feeds dL/dz_before
        const bool z32 = m->arch == DMB_ARCH_Z32;
        DMB_TRY(tm_backward(z32 ? w.za : w.zb, batch, (int64_t)L.D * L.lh * L.lw, w.tm_scratch,
                            grad_scale * tm->weight, w.g_tm, 0, c.st));
        g_tm = w.g_tm;
    }
    if (m->arch == DMB_ARCH_Z32)
        return run_backward_z32(c, params, x, mask, mask_channels, channel_var, decoded, grad_scale, grads, g_tm);
    return run_backward_z16(c, params, x, mask, mask_channels, channel_var, decoded, grad_scale, grads, g_tm);
}

int dmb_train_backward(const dmb_model* m, const float* packed, const float* params, const float* x,
                       const float* mask, int32_t mask_channels, const float* channel_var,
                       const float* decoded, int64_t batch, float grad_scale, float* grads, void* workspace,
                       size_t workspace_bytes, void* stream) {
    return dmb_train_backward_tm(m, packed, params, x, mask, mask_channels, channel_var, decoded, batch, nullptr,
                                 grad_scale, grads, workspace, workspace_bytes, stream);
}

long long dmb_launch_count(int reset) {
    const long long v = dmb::g_launches.load();
    if (reset) dmb::g_launches.store(0);
    return v;
}

int dmb_recon_loss(const float* decoded, const float* x, const float* mask, int32_t mask_channels,
                   const float* channel_var, int64_t batch, int32_t channels, int32_t hw,
                   double* sum_out, void* stream) {
    DMB_CHECK(decoded && x && channel_var && sum_out, "dmb_recon_loss: null pointer");
    DMB_CHECK(hw % 4 == 0, "dmb_recon_loss: H*W must be a multiple of 4");
    DMB_CHECK(!mask || mask_channels == 1 || mask_channels == channels, "dmb_recon_loss: mask channels %d", mask_channels);
    const int64_t total4 = batch * channels * (hw / 4);
    int64_t blocks = (total4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    DMB_LAUNCH((recon_loss_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, decoded, x, mask, mask_channels, channel_var, total4, channels, hw / 4, sum_out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // extern "C"
