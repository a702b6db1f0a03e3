// Shared declarations for the dynamorph_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/dynamorph_b200.h"

namespace dmb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;   // kernels launched by this library (dmb_launch_count)
#define DMB_LAUNCHED(n) (::dmb::g_launches.fetch_add(n, std::memory_order_relaxed))

#define DMB_CHECK(cond, ...)                      \
    do {                                          \
        if (!(cond)) {                            \
            ::dmb::set_error(__VA_ARGS__);        \
            return -1;                            \
        }                                         \
    } while (0)

#define DMB_CUDA(expr)                                                                    \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::dmb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                             __FILE__, __LINE__);                                         \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

#define DMB_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != 0) return _r;    \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the library starts with pdl_wait() (griddepcontrol.wait: returns
// once the preceding kernel in the stream has completed and flushed; a no-op for an ordinary launch) and is launched
// through DMB_LAUNCH with cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs can be scheduled while the
// previous kernel drains: the steps are chains of short dependent kernels (85 launches in a 1.3 ms training step).
// Stream capture records these as programmatic edges of the CUDA graph.  DMB_PDL=0 launches the ordinary way.
// ------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define DMB_LAUNCH(kern, grid, block, smem, stream, ...)                                                   \
    do {                                                                                                   \
        cudaError_t _le = ::dmb::launch_k(kern, dim3(grid), dim3(block), (size_t)(smem),                   \
                                          (cudaStream_t)(stream), __VA_ARGS__);                            \
        if (_le != cudaSuccess) {                                                                          \
            ::dmb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_le), __FILE__, __LINE__); \
            return -2;                                                                                     \
        }                                                                                                  \
    } while (0)

// ------------------------------------------------------------------------------------
// direct convolution, forward (conv_fwd.cu)
// ------------------------------------------------------------------------------------
struct ConvFwdArgs {
    const float* x;         // (B, Cin, H, W)
    float* y;               // (B, Cout, Ho, Wo)
    const float* w;         // packed [Cin][KS][KS][Cout]
    const float* bias;      // [Cout] or [9][Cout] when bias_classes (row class*3 + col class)
    int bias_classes;       // 0 / 1
    // transform applied to x on load:  v = relu?(v * in_scale[c] + x2 * in_b[c] + in_shift[c])
    const float* in_scale;  // nullptr = identity; [Cin] or [B][Cin]
    const float* in_shift;
    int in_per_sample;      // stride B over the affine table
    int in_relu;
    const float* x2;        // optional second input tensor (BatchNorm backward: the raw activation)
    const float* in_b;      // its per-channel coefficient table (same indexing as in_scale)
    // epilogue:  o = acc + bias;  o *= [mask_src*mask_s+mask_t > 0];  o += skip;  o = relu?(o)
    const float* mask_src;  // (B, Cout, Ho, Wo) or nullptr: ReLU gate of the layer back-propagated into
    const float* mask_s;    // its affine (nullptr = identity), [Cout] or [B][Cout] (mask_per_sample)
    const float* mask_t;
    int mask_per_sample;
    const float* skip;      // (B, Cout, Ho, Wo) added to the output, or nullptr
    int out_relu;
    double* stats;          // [B][nbands][Cout][2] per-CTA partial sums or nullptr
    const float* stat_src;  // nullptr: (sum o, sum o^2); else (sum o, sum o*stat_src)  (BatchNorm backward)
    int B, Cin, H, W, Cout, Ho, Wo;
    int ks, stride;         // (1,1) (3,1) (4,2)
    int out_nhwc;           // y is (B, Ho, Wo, Cout): only the head shape of the 64-wide encoder (4x4 s2, 2 -> 32 @128),
                            // plain call (no skip / statistics / transform); anything else is an error
};
int conv_fwd(const ConvFwdArgs& a, cudaStream_t st);
// number of (sum, sumsq) partial rows per sample the kernel will write for this geometry; `plain` = a forward
// call without the data-gradient extras (x2 / mask_src / stat_src), which the TMA kernel serves when it can
int conv_fwd_bands(int ks, int stride, int Cin, int Cout, int Ho, int Wo, bool plain, int64_t B);
// TMA-fed specialisation (conv_tma.cu): returns 1 when it does not take the call
int conv_tma(const ConvFwdArgs& a, cudaStream_t st);
int conv_tma_bands(int ks, int stride, int Cin, int Cout, int H, int W, int64_t B);

// tensor-core (tcgen05, 3xTF32) convolution over NHWC activations (conv_tc.cu); EVAL-mode wide configurations
struct ConvTcArgs {
    const float* x;         // (B, H, W, Cin)  NHWC
    const float* wtc;       // pack_tc_weights output: [tap][Cin/32][hi|lo][Cout][32], 128-byte swizzled rows
    const float* bias;      // [Cout]
    float* y;               // (B, Ho, Wo, Cout) when out_nhwc, else (B, Cout, Ho, Wo)
    const float* skip;      // optional, added before the output ReLU; layout per skip_nhwc
    int B, Cin, H, W, Cout, ks, stride;   // (1,1) (3,1) (4,2); pad = ks > 1
    int in_relu, out_relu, out_nhwc, skip_nhwc;
};
bool conv_tc_supported(int cin, int cout, int ks, int stride, int H, int W);
int64_t conv_tc_weight_floats(int cin, int cout, int ks);
// w_packed = [Cin][ks][ks][Cout] (the layout every other conv kernel reads) -> tensor-core tiles
int pack_tc_weights(const float* w_packed, float* out, int cin, int cout, int ks, cudaStream_t st);
int nchw_to_nhwc(const float* x, float* y, int64_t B, int C, int HW, cudaStream_t st);
int conv_tc(const ConvTcArgs& a, cudaStream_t st);

// Winograd F(2x2,3x3) on the tensor cores for the 16-channel 3x3 layers at 16x16 (conv_wino_tc.cu); EVAL mode, no skip /
// statistics / BatchNorm transform
struct ConvWinoArgs {
    const float* x;         // (B, 16, 16, 16) NCHW
    const float* u;         // pack_wino_weights output: [hi|lo][xi pair][Cout][32], 128-byte swizzled rows
    const float* bias;      // [Cout]
    float* y;               // (B, Cout, 16, 16)
    int B, Cout;
    int in_relu, out_relu;
    // optional fused ResidualBlock tail (Cout == 32): y2 = x + conv1x1(relu?(y)) + bias2, y itself is not written
    const float* w2;        // [32][16] (the packed [Cin][1][1][Cout] layout of the 1x1) or nullptr
    const float* bias2;     // [16]
    float* y2;              // (B, 16, 16, 16)
};
bool conv_wino_supported(int cin, int cout, int ks, int stride, int H, int W);
int64_t conv_wino_weight_floats(int cin, int cout);
int pack_wino_weights(const float* w_packed, float* out, int cin, int cout, cudaStream_t st);
int conv_wino(const ConvWinoArgs& a, cudaStream_t st);

// Whole-batch statistics of conv_tm.cu finalised INSIDE the producing kernel: the last CTA to finish (ticket counter)
// folds the per-CTA rows in a fixed order and does what bn_finalize (mode 1) / bn_backward_finalize (mode 2) would have
// done in a launch of their own.  ticket: one zeroed unsigned in device memory, left zero again.
struct TmFinArgs {
    unsigned* ticket;
    int mode;                // 0 none, 1 BatchNorm forward (scale / shift / running statistics), 2 BatchNorm backward (A, Bc, Cc, dgamma, dbeta)
    double cnt;              // elements per channel (pixels x batch)
    const float* gamma; const float* beta; float eps, momentum;
    float* scale; float* shift; float* running_mean; float* running_var; float* save_mean; float* save_invstd;   // mode 1
    const float* mean; const float* invstd; float* A; float* Bc; float* Cc; float* dgamma; float* dbeta;         // mode 2
};

// thin-channel convolutions on tcgen05 with the activation operand in tensor memory (conv_tm.cu); EVAL mode, NCHW in/out
struct ConvTmArgs {
    const float* x;         // (B, Cin, H, W) NCHW
    const float* wtm;       // pack_tm_weights output
    const float* bias;      // [Cout]
    float* y;               // (B, Cout, Ho, Wo)
    const float* skip;      // optional (B, Cout, Ho, Wo), added before the output ReLU
    int B, Cin, H, W, Cout, ks, stride;
    int in_relu, out_relu;
    // fused ResidualBlock layer (3x3 16 -> 32 only): bias2 != nullptr -> y (B, 16, Ho, Wo) = x + conv1x1(relu(conv3x3(
    // relu?(x)) + bias)) + bias2; the weight image then holds the 1x1's image (pack_tm_weights(w2, ., 32, 16, 1)) right
    // behind the 3x3's; out_relu here names the ReLU BETWEEN the two convolutions and must be set, skip must be null
    const float* bias2;
    // train-mode BatchNorm around the layer (bn != 0; DMB_BN_PER_SAMPLE / DMB_BN_BATCH): the producer's pending affine
    // relu?(x * in_scale[c] + in_shift[c]) ([Cin] or [B][Cin] tables, nullptr = identity) is applied on load, y receives
    // the raw output and `stats` one (sum, sum of squares) double pair per (patch, warp of 32 pixels, channel), conv_tm_bands() rows
    // per patch -- the layout bn_finalize reads.  skip / out_relu / bias2 must be unset.
    int bn;
    const float* in_scale;
    const float* in_shift;
    int in_per_sample;
    double* stats;
    // whole-batch statistics (DMB_BN_BATCH): every warp of the persistent CTAs adds up the partials of its tiles and
    // `stats` receives *stat_rows = grid <= TM_BATCH_ROWS_MAX rows of [Cout][2], one per CTA (fixed order: deterministic)
    int stats_batch;
    int* stat_rows;
    // data-gradient form (dg != 0; training step, whole-batch statistics; shapes: conv_tm_dg_supported): plain input x (the
    // gradient at the layer's output), wtm = image of the data-gradient weights, epilogue
    //   o = acc + bias;  o *= [mask_src * mask_s[c] + mask_t[c] > 0];  o += skip;  store;  stats: (sum o, sum o * stat_src)
    // or (sum o, sum o^2) without stat_src -- the contract of conv_fwd's data-gradient extras
    int dg;
    const float* mask_src;
    const float* mask_s;
    const float* mask_t;
    const float* stat_src;
    // ... with the BatchNorm backward in front of the layer applied on load (shapes: conv_tm_dg_dual):
    //   x := in_a[c] * x + in_b[c] * x2 + in_c[c]   ([Cin] tables; x2 = the raw activation, same shape as x)
    const float* x2;
    const float* in_a;
    const float* in_b;
    const float* in_c;
    // transposed form (ct != 0): ConvTranspose2d(4x4, stride 2, padding 1) Cin -> Cout, x (B, Cin, H, W) -> y (B, Cout, 2H,
    // 2W); wtm = pack_tm_weights_ct image; plain (bias + optional ReLU on store) or with dg = 1 (gate / sums / dual load;
    // statistics rows: one per CTA, [Cout][2]).  Shapes: conv_tm_ct_supported.
    int ct;
    const TmFinArgs* fin;   // whole-batch statistics finalised in the kernel (nullptr: rows only)
};
constexpr int TM_MAX_SMS = 192;
constexpr int TM_BATCH_ROWS_MAX = 4 * 2 * TM_MAX_SMS;
int conv_tm_bands(int cin, int cout, int ks, int stride, int H, int W);     // statistics rows per patch (0: unsupported)
bool conv_tm_supported(int cin, int cout, int ks, int stride, int H, int W);
bool conv_tm_dg_supported(int cin, int cout, int ks, int stride, int H, int W);
bool conv_tm_dg_dual(int cin, int cout, int ks, int stride, int H, int W);
bool conv_tm_ct_supported(int cin, int cout, int H, int W, bool dg);
int pack_tm_weights_ct(const float* w_packed, float* out, int cin, int cout, cudaStream_t st);   // image: conv_tm_weight_floats(cin, 4*cout, 3)
int64_t conv_tm_weight_floats(int cin, int cout, int ks);
int pack_tm_weights(const float* w_packed, float* out, int cin, int cout, int ks, cudaStream_t st);
constexpr int TM_PACK_MAX = 32;
struct TmPackJob { const float* w; float* out; int cin, cout, ks; int ct; };      // ct: transposed form (w = [Cin][4][4][Cout])
int pack_tm_weights_multi(const TmPackJob* jobs, int n, cudaStream_t st);      // all layers of a model in one launch
int conv_tm(const ConvTmArgs& a, cudaStream_t st);

// transposed 4x4 stride-2 pad-1 convolution, forward (convt_fwd.cu)
struct ConvTFwdArgs {
    const float* x;         // (B, Cin, H, W)
    float* y;               // (B, Cout, 2H, 2W)
    const float* w;         // packed [Cin][4][4][Cout]
    const float* bias;      // [Cout]
    const float* in_scale;  // as ConvFwdArgs
    const float* in_shift;
    int in_per_sample;
    int in_relu;
    const float* x2;
    const float* in_b;
    const float* mask_src;  // (B, Cout, 2H, 2W)
    const float* mask_s;
    const float* mask_t;
    int mask_per_sample;
    int out_relu;
    double* stats;          // [B][nbands][Cout][2] or nullptr
    const float* stat_src;
    int B, Cin, H, W, Cout;
};
int convt_fwd(const ConvTFwdArgs& a, cudaStream_t st);
int convt_fwd_bands(int Cin, int Cout, int H, int W);

// ------------------------------------------------------------------------------------
// weight gradients (wgrad.cu)
// ------------------------------------------------------------------------------------
struct WgradArgs {
    const float* g;         // (B, Cout, Ho, Wo) gradient tensor;  gy = g*ga[c] + y*gb[c] + gc[c]
    const float* y;         // optional raw conv output (BatchNorm backward) or nullptr
    const float* ga; const float* gb; const float* gc;   // nullptr = identity
    int g_per_sample;
    int g_relu;             // relu(gy) after the affine (ConvTranspose2d whose input is relu(bn(.)))
    const float* x;         // (B, Cin, H, W) conv input before the producer's BN/ReLU
    const float* xs; const float* xt;                    // act = relu?(x*xs[c] + x2*xb[c] + xt[c]); nullptr = identity
    const float* x2; const float* xb;                    // optional second tensor (ConvTranspose2d followed by BN)
    int x_per_sample;
    int x_relu;
    int ones_channel;       // append a constant-one input channel (composite head bias chain)
    int B, Cin, H, W, Cout, Ho, Wo, ks, stride;
    float* partials;        // scratch: wgrad_partial_floats() floats per CTA x *ncta
};
int wgrad_partial_floats(const WgradArgs& a, int* ncta);
// With a queue the fold of the per-CTA partials is NOT launched: the job is appended and wgrad_reduce_flush() folds every
// queued layer in ONE launch (the partial regions must then stay untouched until the flush: one region per layer).
struct WgReduceJob { const float* partials; int ncta, out_floats, cin_eff, cout, ks; float* dw; float* db; float* packed_out; int block0; };
constexpr int WG_REDUCE_MAX = 24;
struct WgReduceQueue { int n = 0; int blocks = 0; WgReduceJob j[WG_REDUCE_MAX]; };
int wgrad(const WgradArgs& a, float* dw, float* db, float* packed_out, cudaStream_t st, WgReduceQueue* q = nullptr);
int wgrad_reduce_flush(WgReduceQueue& q, cudaStream_t st);
// TMA-fed specialisation for the default-width layers (wgrad_tma.cu): both return 1 when they do not take the call;
// they write the same per-CTA partial layout, which wgrad() folds with the same reduction kernel
int wgrad_tma_plan(const WgradArgs& a, int* ncta, int* out_floats);
int wgrad_tma(const WgradArgs& a, int* ncta, int* out_floats, cudaStream_t st);
int composite_chain(const float* dweff, const float* w0, const float* b0, const float* w1, int ni, int cm,
                    float* dw0, float* db0, float* dw1, float* db1, cudaStream_t st);

// ------------------------------------------------------------------------------------
// batch-norm bookkeeping (bn.cu)
// ------------------------------------------------------------------------------------
// Reduce per-CTA partials into per-channel (or per-sample-per-channel) scale/shift:
//   scale = gamma / sqrt(var + eps), shift = beta - mean * scale
// BATCH mode (per_sample = 0) also updates running stats and writes mean / invstd.
struct BnFinalizeArgs {
    const double* partials;  // [B][nbands][C][2]
    int B, nbands, C;
    int rows;                // BATCH mode: number of partial rows when it is not B * nbands (0 = B * nbands)
    int64_t count_per_sample;  // Ho*Wo
    int per_sample;
    const float* gamma;
    const float* beta;
    float eps, momentum;
    float* scale;            // [C] or [B][C]
    float* shift;
    float* running_mean;     // may be nullptr
    float* running_var;
    float* save_mean;        // [C] (BATCH) or [B][C]; may be nullptr
    float* save_invstd;
};
int bn_finalize(const BnFinalizeArgs& a, cudaStream_t st);

// BatchNorm backward bookkeeping: from per-CTA partial (sum g, sum g*y) build dgamma, dbeta and the
// coefficients of  dL/dy = A*g + Bc*y + Cc  that the dgrad / wgrad kernels apply on load.
struct BnBwdArgs {
    const double* partials;  // [B][nbands][C][2]
    int B, nbands, C;
    int rows;                // BATCH mode: number of partial rows when it is not B * nbands (0 = B * nbands)
    int64_t count_per_sample;
    int per_sample;
    const float* gamma;
    const float* mean;       // saved by the forward finalize ([C] or [B][C])
    const float* invstd;
    float* A; float* Bc; float* Cc;     // [C] or [B][C]
    float* dgamma; float* dbeta;        // [C]; written (BATCH) or accumulated over samples (PER_SAMPLE)
    int grad_div;                       // > 1: dgamma / dbeta are divided by it (synchronised BatchNorm: the sums are global)
};
int bn_backward_finalize(const BnBwdArgs& a, cudaStream_t st);
// out[c] = sum over (b, band) of partials[b][band][c][0]   (bias gradient of a layer without BatchNorm)
int sum_partials(const double* partials, int B, int nbands, int C, float* out, cudaStream_t st);
// out[c][0..1] = sum over n rows of partials[row][c][0..1], in double, fixed order (synchronised BatchNorm: the 2*C
// doubles that cross the ranks)
int fold_partials(const double* partials, int64_t n, int C, double* out, cudaStream_t st);

// out = (a*sa+ta) + (b*sb+tb), all (B, C, HW); the affine tables follow the per_sample flag.
struct AffineAddArgs {
    const float* a; const float* sa; const float* ta;   // sa == nullptr: identity
    const float* b; const float* sb; const float* tb;
    int per_sample;
    float* out;
    int64_t B; int C; int HW;
};
int affine_add(const AffineAddArgs& a, cudaStream_t st);

// The z16 decoder's tail in the training step (dec_tail.cu): conv1x1 CM -> NI at full resolution + the reconstruction
// loss in one streaming kernel; loss gradient + the 1x1's data / weight / bias gradients + the previous layer's bias
// gradient in another.
struct DecTailArgs {
    int64_t B; int cm, ni, hw;
    const float* t3;       // (B, CM, H, W) post-ReLU input of the 1x1
    const float* x;        // (B, NI, H, W) the batch
    const float* mask; int mask_c; const float* cvar;
    const float* w;        // packed [CM][NI]
    const float* bias;     // [NI] (forward)
    float* decoded;        // forward: out, backward: in
    double* loss_sum;      // forward: += sum of the weighted squared error
    float* g_t3;           // backward: gradient at t3 (ReLU gate applied)
    float scale;           // backward: grad_scale * weight_recon / numel(decoded)
    double* partials;      // backward: dec_tail_partial_doubles() scratch
    unsigned* ticket;      // backward: zero on entry (reset on exit)
    float* dw; float* db; float* db_prev;      // (NI, CM), (NI), (CM)
};
// ... with dec.4 (ConvTranspose2d ci -> cm, 4x4 s2 p1, + ReLU) in front of the forward tail: t2 -> t3, decoded, loss
struct DecTail2Args {
    int64_t B; int ci, cm, ni, hi, wi;      // hi x wi: the maps of t2 (t3 / decoded / x are 2hi x 2wi)
    const float* t2; const float* w4; const float* b4; float* t3;
    const float* x; const float* mask; int mask_c; const float* cvar;
    const float* w6; const float* b6; float* decoded; double* loss_sum;
};
bool dec_tail2_supported(int ci, int cm, int ni, int hi, int wi);
int dec_tail2_forward(const DecTail2Args& a, cudaStream_t st);
// ConvTranspose2d(ci -> 4, 4x4 s2 p1) + bias + optional ReLU on its own, plain input (dec_tail.cu: convt_small_kernel)
bool convt_small_supported(int ci, int co, int hi, int wi);
int convt_small(const float* x, const float* w_packed, const float* bias, float* y, int64_t B, int ci, int co, int hi, int wi,
                int relu, cudaStream_t st);
bool dec_tail_supported(int cm, int ni, int hw);
int64_t dec_tail_partial_doubles(int64_t B, int hw, int cm, int ni);
int dec_tail_forward(const DecTailArgs& a, cudaStream_t st);
int dec_tail_backward(const DecTailArgs& a, cudaStream_t st);

// ------------------------------------------------------------------------------------
// vector quantiser (vq.cu)
// ------------------------------------------------------------------------------------
struct VqArgs {
    const float* z;          // (B, D, P) ; if pre_b != nullptr, z = pre_a + pre_b*sb+tb is formed first
    const float* pre_a; const float* pre_sa; const float* pre_ta;
    const float* pre_b; const float* pre_sb; const float* pre_tb;
    int pre_per_sample;
    float* z_before_out;     // written when the pre-transform is used (or nullptr)
    const float* codebook;   // (K, D)
    int64_t B; int D; int P; int K;
    float* z_st;             // may be nullptr
    int32_t* idx;            // may be nullptr
    double* stats;           // [2+K] or nullptr
};
int vq_forward(const VqArgs& a, cudaStream_t st);
// tensor-core code search + exact refinement of the candidates (vq_tc.cu); returns 1 when it does not take the call
// (K outside [16, 512], D not in {16, 32, 64}); vq_forward tries it first unless DMB_VQ_TC=0
int vq_forward_tc(const VqArgs& a, cudaStream_t st);
// gradient of the quantiser; optional per-CTA BatchNorm-backward sums [B*p/128][d][2] against stat_src
// time-matching loss (matching.cu)
size_t tm_scratch_floats(int64_t B, int64_t L);
int tm_forward(const float* z, int64_t B, int64_t L, const dmb_time_matching& tm, float* scratch, float* loss_out,
               cudaStream_t st);
int tm_backward(const float* z, int64_t B, int64_t L, const float* scratch, float scale, float* g, int accumulate,
                cudaStream_t st);

// g_extra (optional, same shape as z): added to the straight-through gradient (time-matching term)
int vq_backward_stats(const float* z, const float* codebook, const int32_t* idx, const float* g_zst, const float* g_extra,
                      float g_loss_scale, float beta, int64_t batch, int d, int p, int k, float* grad_z,
                      float* grad_codebook, double* stats, const float* stat_src, float* scratch, int scratch_rows,
                      cudaStream_t st);
int vq_codebook_grad_only(const float* z, const float* codebook, const int32_t* idx, float g_loss_scale, int64_t batch,
                          int d, int p, int k, float* grad_codebook, float* scratch, int scratch_rows, cudaStream_t st);

}  // namespace dmb
