/*
 * dynamorph_b200 -- C ABI of the B200-native VQ-VAE latent-encoding path.
 *
 * The reference (mehta-lab/dynamorph) has no FFI: its boundary for this path is the
 * Python class API (HiddenStateExtractor/vq_vae.py, HiddenStateExtractor/vae.py,
 * pipeline/patch_VAE.py:process_VAE, run_training.py:run_one_batch).  This header is the
 * C-ABI a maintainer binds underneath that API (ctypes stub in INTEGRATION.md); the Python
 * mirror in dynamorph_b200/ is exactly such a binding.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; dmb_last_error() gives the text
 *     (thread-local).  No exceptions cross the ABI, nothing is allocated, no implicit sync.
 *   - every pointer is a caller-owned DEVICE pointer unless its name ends in _host;
 *     every launch goes to the caller's `stream` (a cudaStream_t passed as void*).
 *   - tensors are float32, dense NCHW (the reference's layout); indices are int32.
 *   - `params`  : all trainable tensors, concatenated in reference state_dict order, each
 *                 in its torch layout (Conv2d (Cout,Cin,kh,kw); ConvTranspose2d
 *                 (Cin,Cout,kh,kw); BatchNorm weight/bias (C); Embedding (K,D)).
 *     `bnbuf`   : for every BatchNorm in state_dict order: running_mean[C], running_var[C].
 *     dmb_param_lookup() exposes the offsets by state_dict key.
 */
#ifndef DYNAMORPH_B200_H
#define DYNAMORPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_ABI_VERSION 1

/* architecture: HiddenStateExtractor/vq_vae.py:276-298 == vae.py:273-294 (z16); vae.py:401-414 (z32) */
enum { DMB_ARCH_Z16 = 0, DMB_ARCH_Z32 = 1 };

/* BatchNorm statistics mode:
 *   EVAL       running statistics (model.eval()); folded into the conv weights at pack time
 *   BATCH      statistics of the current call's batch (train mode; run_training.py:404)
 *   PER_SAMPLE train-mode statistics with batch 1, i.e. per patch -- what
 *              pipeline/patch_VAE.py:445-449 computes as written                              */
enum { DMB_BN_EVAL = 0, DMB_BN_BATCH = 1, DMB_BN_PER_SAMPLE = 2 };

/* Constructor arguments of VQ_VAE (vq_vae.py:232-245) that shape the computation. */
typedef struct dmb_model {
    int32_t arch;
    int32_t num_inputs;
    int32_t num_hiddens;
    int32_t num_residual_hiddens;
    int32_t num_residual_layers;
    int32_t num_embeddings;
    int32_t height;            /* input patch height (128 in the reference pipeline) */
    int32_t width;
    float   commitment_cost;
    float   weight_recon;      /* ignored (1) for z32: vae.py:440 */
    float   weight_commitment; /* ignored (1) for z32 */
    float   bn_eps;            /* 1e-5 */
    float   bn_momentum;       /* 0.1  */
} dmb_model;

int         dmb_abi_version(void);
const char* dmb_last_error(void);

/* ---- parameter layout -------------------------------------------------------------- */
/* Totals: trainable floats in `params`, floats in `bnbuf`, number of BatchNorm layers.  */
int dmb_param_count(const dmb_model* m, int64_t* n_params, int64_t* n_bnbuf, int32_t* n_bn);
/* Offset/numel of one state_dict entry (e.g. "enc.4.weight", "enc.5.running_var",
 * "vq.w.weight").  *which = 0 if it lives in `params`, 1 if in `bnbuf`.                  */
int dmb_param_lookup(const dmb_model* m, const char* key_host, int32_t* which,
                     int64_t* offset, int64_t* numel);
/* Latent geometry: D = num_hiddens, (lh, lw) = 16x16 (z16) or 32x32 (z32) for 128x128.  */
int dmb_latent_shape(const dmb_model* m, int32_t* d, int32_t* lh, int32_t* lw);

/* ---- weight packing ---------------------------------------------------------------- */
/* Kernel-layout copy of the weights ([Cin][kh][kw][Cout], conv1x1(2->h/2) composed into
 * the first 4x4 conv, BN folded in for DMB_BN_EVAL).  Re-run whenever params change.     */
int dmb_packed_floats(const dmb_model* m, int64_t* n);
int dmb_pack_weights(const dmb_model* m, const float* params, const float* bnbuf,
                     int32_t bn_mode, float* packed, void* stream);

/* ---- encode path: model.enc + model.vq (pipeline/patch_VAE.py:448-449) ------------- */
/* Bytes of scratch for a batch of B patches.  keep_activations != 0 reserves room for
 * everything the backward pass re-reads.                                                 */
int dmb_workspace_bytes(const dmb_model* m, int64_t batch, int32_t bn_mode,
                        int32_t keep_activations, size_t* bytes);

/* z_before = enc(x).  x: (B, num_inputs, H, W); z_before: (B, D, lh, lw).
 * BATCH mode also updates running_mean / running_var in `bnbuf_inout` (may be NULL).     */
int dmb_encoder_forward(const dmb_model* m, const float* packed, const float* x, int64_t batch,
                        int32_t bn_mode, float* z_before, float* bnbuf_inout,
                        void* workspace, size_t workspace_bytes, void* stream);

/* VectorQuantizer.forward (vq_vae.py:52-84) on z: (B, D, lh, lw), codebook (K, D).
 *   z_st      : z + (q - z)                       (may be NULL)
 *   idx       : int32 (B, lh, lw) first-argmin    (may be NULL)
 *   stats     : double[2+K] accumulators, zeroed by the caller's previous dmb_vq_reset():
 *               [0] = sum (q - z)^2, [1] = positions, [2..] = code histogram (may be NULL) */
int dmb_vq_forward(const float* z, const float* codebook, int64_t batch, int32_t d,
                   int32_t positions_per_patch, int32_t k, float* z_st, int32_t* idx,
                   double* stats, void* stream);
int dmb_vq_reset(double* stats, int32_t k, void* stream);
/* loss = (1 + commitment_cost) * mean((q - z)^2); perplexity = exp(-sum p log(p + 1e-10)).
 * out2: float[2] = {loss, perplexity}.                                                   */
int dmb_vq_finalize(const double* stats, int32_t d, int32_t k, float commitment_cost,
                    float* out2, void* stream);
/* VectorQuantizer.decode_inputs (vq_vae.py:105-116): idx (B, lh, lw) -> (B, D, lh, lw).  */
int dmb_vq_gather(const int32_t* idx, const float* codebook, int64_t batch, int32_t d,
                  int32_t positions_per_patch, int32_t k, float* q, void* stream);

/* Gradient of VectorQuantizer.forward (autograd of vq_vae.py:65-76).  g_zst: gradient w.r.t. the
 * straight-through output (may be NULL = 0); g_loss_dev: device scalar gradient of the loss output
 * (NULL = 1), multiplied by g_loss_scale.  grad_z / grad_codebook (zeroed inside) may be NULL.   */
int dmb_vq_backward(const float* z, const float* codebook, const int32_t* idx, const float* g_zst,
                    const float* g_loss_dev, float g_loss_scale, float commitment_cost, int64_t batch,
                    int32_t d, int32_t positions_per_patch, int32_t k, float* grad_z,
                    float* grad_codebook, void* stream);

/* enc + vq in one call: the process_VAE hot path.  Any of z_before / z_after / idx /
 * vq_stats may be NULL.                                                                  */
int dmb_encode(const dmb_model* m, const float* packed, const float* codebook, const float* x,
               int64_t batch, int32_t bn_mode, float* z_before, float* z_after, int32_t* idx,
               double* vq_stats, void* workspace, size_t workspace_bytes, void* stream);

/* ---- decoder + losses: model.dec and the loss of VQ_VAE.forward (vq_vae.py:319-323) -- */
int dmb_decoder_forward(const dmb_model* m, const float* packed, const float* z_after,
                        int64_t batch, int32_t bn_mode, float* decoded, float* bnbuf_inout,
                        void* workspace, size_t workspace_bytes, void* stream);
/* sum over all elements of ((decoded*mask - x*mask)^2 / channel_var[c]) -> *sum_out (double,
 * accumulated; zero it first).  mask: (B,1,H,W) or (B,C,H,W) or NULL (ones).             */
int dmb_recon_loss(const float* decoded, const float* x, const float* mask, int32_t mask_channels,
                   const float* channel_var, int64_t batch, int32_t channels, int32_t hw,
                   double* sum_out, void* stream);

/* ---- ResidualBlock.forward stand-alone (vq_vae.py:180-225) ------------------------------ */
/* y = x; for every layer: y = y + BN(Conv1x1(ReLU(BN(Conv3x3(ReLU(y)))))) on x: (B, num_hiddens, h, w).
 * `params` / `bnbuf` hold the block's own tensors in its state_dict order (layers.{i}.{1,2,4,5}: conv weight, bias,
 * BatchNorm weight, bias / running_mean, running_var); dmb_residual_block_sizes() gives their lengths and the bytes
 * of `workspace` (packed weights + intermediate maps).  BATCH mode updates the running statistics in bnbuf_inout.   */
int dmb_residual_block_sizes(int32_t num_hiddens, int32_t num_residual_hiddens, int32_t num_residual_layers,
                             int64_t batch, int32_t h, int32_t w, int32_t bn_mode, int64_t* n_params,
                             int64_t* n_bnbuf, size_t* workspace_bytes);
int dmb_residual_block_forward(int32_t num_hiddens, int32_t num_residual_hiddens, int32_t num_residual_layers,
                               const float* params, const float* bnbuf, const float* x, int64_t batch, int32_t h,
                               int32_t w, int32_t bn_mode, float* y, float* bnbuf_inout, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ---- single layers (unit tests, micro-benchmarks) --------------------------------------- */
/* nn.Conv2d forward, kernel/stride in {(1,1), (3,1) pad 1, (4,2) pad 1}; w_packed is
 * [Cin][k][k][Cout].  Optional on-load transform relu?(x*in_scale[c]+in_shift[c]) (tables [Cin] or
 * [B][Cin]), optional skip tensor added to the output, optional ReLU on the output.            */
int dmb_conv2d_forward(const float* x, const float* w_packed, const float* bias, float* y,
                       int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout, int32_t ksize,
                       int32_t stride, const float* in_scale, const float* in_shift,
                       int32_t in_per_sample, int32_t in_relu, const float* skip, int32_t out_relu,
                       void* stream);
/* nn.Conv2d forward on the tensor cores (tcgen05, 3xTF32 operand split, fp32 accumulation in TMEM) for the wide
 * layers of BASELINE configs[3]: Cin a multiple of 32, Cout 32 or 64, output width a power of two in [8, 128],
 * same kernel/stride set as dmb_conv2d_forward (reference: vq_vae.py:279-289, :203-209).  nhwc_io = 0: x, skip, y
 * are NCHW (x is transposed into `scratch` first); nhwc_io = 1: all three are NHWC.  `scratch` holds
 * dmb_conv2d_tc_scratch_floats() floats (the transposed input and the split, swizzled weight tiles).            */
int dmb_conv2d_tc_scratch_floats(int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout, int32_t ksize,
                                 int64_t* floats);
int dmb_conv2d_tc(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                  int32_t h, int32_t w, int32_t cout, int32_t ksize, int32_t stride, int32_t in_relu,
                  const float* skip, int32_t out_relu, int32_t nhwc_io, float* scratch, void* stream);
/* nn.Conv2d(16 -> 16|32, 3x3, padding 1) on 16x16 maps as Winograd F(2x2,3x3) on the tensor cores (16 batched TF32x3
 * GEMMs, accumulators in tensor memory): the latent-resolution 3x3 layers of the default configuration in eval mode
 * (reference: vq_vae.py:203-209, :288).  x, y NCHW; w_packed [Cin][3][3][Cout]; optional ReLU on load / on store.
 * `scratch` holds 2*16*Cin*Cout floats (the transformed, split, swizzled weights).  With w2_packed / bias2 / y2 given
 * (Cout = 32) the rest of a ResidualBlock layer is fused behind it: y2 = x + conv1x1(relu?(y)) + bias2 with w2_packed
 * [32][16], and y itself is not written (vq_vae.py:203-209, eval mode with BatchNorm folded).                      */
int dmb_conv2d_wino(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                    int32_t h, int32_t w, int32_t cout, int32_t in_relu, int32_t out_relu, const float* w2_packed,
                    const float* bias2, float* y2, float* scratch, void* stream);
/* nn.Conv2d forward for the THIN layers of the default configuration (8 -> 16 4x4 s2 @64, 16 -> 16 4x4 s2 @32,
 * 16 -> 16|32 3x3 @16, 32 -> 16 1x1 @16; vq_vae.py:280-289, :203-209) on the tensor cores with the activation operand
 * in tensor memory: each thread builds the im2col row of its output pixel in registers, splits it hi/lo (3xTF32) and
 * writes it with tcgen05.st; tcgen05.mma reads A from TMEM and [b_hi; b_lo] from shared memory.  x, skip, y NCHW;
 * w_packed [Cin][k][k][Cout]; `scratch` holds dmb_conv2d_tm_scratch_floats() floats (the split, swizzled weight image). */
int dmb_conv2d_tm_scratch_floats(int32_t cin, int32_t cout, int32_t ksize, int64_t* floats);
int dmb_conv2d_tm(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                  int32_t h, int32_t w, int32_t cout, int32_t ksize, int32_t stride, int32_t in_relu,
                  const float* skip, int32_t out_relu, float* scratch, void* stream);
/* The same layers with a train-mode BatchNorm on either side (DMB_BN_PER_SAMPLE / DMB_BN_BATCH): the producer's pending
 * affine relu?(x * in_scale[c] + in_shift[c]) ([Cin] tables, or [B][Cin] with in_per_sample) is applied on load (NULL =
 * identity), y receives the raw convolution output, and `stats` (may be NULL) one (sum, sum of squares) double pair per
 * (patch, warp of 32 output pixels, channel): [B][*bands][Cout][2], the layout bn_finalize reads.                    */
int dmb_conv2d_tm_bn(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                     int32_t h, int32_t w, int32_t cout, int32_t ksize, int32_t stride, const float* in_scale,
                     const float* in_shift, int32_t in_per_sample, int32_t in_relu, double* stats, int32_t* bands,
                     float* scratch, void* stream);
/* Data gradient of a training step on the same kernel (whole-batch BatchNorm; autograd of F.conv2d w.r.t. its input,
 * run_training.py:406): gx = conv(gy) with the data-gradient weights w_packed [Cin][k][k][Cout] (Cin = channels of gy;
 * stride-1 layers: the layer's weight with channels swapped and taps flipped; a ConvTranspose2d back-propagates through
 * the 4x4 stride-2 convolution of its own taps), then the epilogue of the CUDA-core data-gradient kernels:
 *   gx *= [mask_src * mask_scale[c] + mask_shift[c] > 0]   (ReLU gate of the producer; NULL = none, scale NULL = identity)
 *   gx += skip                                              (NULL = none)
 *   stats (may be NULL): per-CTA partial (sum gx, sum gx * stat_src) -- (sum gx, sum gx^2) without stat_src -- as
 *   *stat_rows rows of [Cout][2] doubles (at most dmb_conv2d_tm_batch_stat_rows() rows), the input of the next
 *   BatchNorm backward / bias gradient.
 * y_raw / ga / gb / gc (all or none; stride-1 shapes except 3x3 32 -> 16): the BatchNorm backward in front of the layer
 * applied on load, gy' = ga[c] * gy + gb[c] * y_raw + gc[c] (both tensors staged by TMA side by side).
 * Shapes (cin -> cout of THIS convolution): 1x1 16 -> 32 @16, 3x3 32 -> 16 @16, 3x3 16 -> 16 @16, 4x4 s2 8 -> 16 @64 and
 * @32, 4x4 s2 16 -> 16 @32.  scratch: dmb_conv2d_tm_scratch_floats(cin, cout, k) rounded up to 64, + cout floats.          */
int dmb_conv2d_tm_batch_stat_rows(int32_t* rows);
int dmb_conv2d_tm_dgrad(const float* gy, const float* w_packed, float* gx, int64_t batch, int32_t cin, int32_t h, int32_t w,
                        int32_t cout, int32_t ksize, int32_t stride, const float* y_raw, const float* ga, const float* gb,
                        const float* gc, const float* mask_src, const float* mask_scale,
                        const float* mask_shift, const float* skip, double* stats, const float* stat_src,
                        int32_t* stat_rows, float* scratch, void* stream);
/* nn.ConvTranspose2d(cin, cout, 4, stride 2, padding 1) (vq_vae.py:292-296) on the same kernel: the layer is the 3x3
 * convolution it amounts to on the INPUT grid with 4*cout phase channels (unused taps zero) followed by a pixel shuffle
 * in the epilogue.  x (B, cin, h, w) -> y (B, cout, 2h, 2w); w_packed [cin][4][4][cout]; bias [cout] (NULL = none).
 *   data_gradient = 0: bias, optional ReLU on store.  Shape: 16 -> 8 @16 (the default decoder's first layer).
 *   data_gradient = 1: the data gradient of a stride-2 nn.Conv2d(cout -> cin, 4, 2, 1) of the training step
 *     (w_packed = that layer's weight as [its Cout][4][4][its Cin]), with the extras of dmb_conv2d_tm_dgrad: BatchNorm
 *     backward on load (y_raw / ga / gb / gc), ReLU gate, per-CTA sums (one row per CTA, at most
 *     dmb_conv2d_tm_batch_stat_rows()).  Shapes: 16 -> 16 @16 (enc.7) and 16 -> 8 @32 (enc.4).
 * scratch: dmb_conv2d_tm_scratch_floats(cin, 4*cout, 3) rounded up to 64, + cout floats.                            */
int dmb_conv_transpose2d_tm(const float* x, const float* w_packed, const float* bias, float* y, int64_t batch, int32_t cin,
                            int32_t h, int32_t w, int32_t cout, int32_t out_relu, int32_t data_gradient, const float* y_raw,
                            const float* ga, const float* gb, const float* gc, const float* mask_src,
                            const float* mask_scale, const float* mask_shift, double* stats, const float* stat_src,
                            int32_t* stat_rows, float* scratch, void* stream);
/* One whole ResidualBlock layer of the default configuration at the 16x16 latent (vq_vae.py:203-209, :222-225, eval mode
 * with BatchNorm folded): y = x + conv1x1(relu(conv3x3(relu(x)) + bias1)) + bias2 in ONE tensor-core kernel; the 1x1 is a
 * second GEMM whose activation operand is written to tensor memory by the first one's epilogue.  x, y (B,16,16,16);
 * w1_packed [16][3][3][32]; w2_packed [32][1][1][16]; scratch: dmb_residual_layer_tm_scratch_floats() floats.        */
int dmb_residual_layer_tm_scratch_floats(int64_t* floats);
int dmb_residual_layer_tm(const float* x, const float* w1_packed, const float* bias1, const float* w2_packed,
                          const float* bias2, float* y, int64_t batch, float* scratch, void* stream);
/* nn.ConvTranspose2d(k=4, stride=2, padding=1) forward; w_packed is [Cin][4][4][Cout].        */
int dmb_conv_transpose2d_forward(const float* x, const float* w_packed, const float* bias, float* y,
                                 int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout,
                                 const float* in_scale, const float* in_shift, int32_t in_per_sample,
                                 int32_t in_relu, int32_t out_relu, void* stream);
/* Weight and bias gradient of one Conv2d (autograd of F.conv2d w.r.t. weight / bias, run_training.py:406):
 *   dw[co][ci][ky][kx] = sum_{b,oy,ox} gy'[b][co][oy][ox] * act[b][ci][S*oy - P + ky][S*ox - P + kx],  db[co] = sum gy'
 * with the transforms the training step folds into the loads:  act = relu?(x * x_scale[c] + x_shift[c])  (the producer's
 * BatchNorm + ReLU; NULL = identity) and  gy' = gy * ga[c] + y_raw * gb[c] + gc[c]  (the BatchNorm backward of the layer's
 * own BatchNorm; ga NULL = plain gy, y_raw/gb NULL = affine only).  ksize/stride: (1,1) (3,1) (4,2), padding 1 unless 1x1.
 * dw (Cout, Cin, k, k) and db (Cout, may be NULL) in the torch layout.  The default-width layer shapes on 128x128 patches
 * run the TMA-fed kernel (csrc/wgrad_tma.cu), everything else the generic one (csrc/wgrad.cu); both are deterministic.
 * scratch: dmb_conv2d_weight_grad_scratch_floats() floats.                                                              */
int dmb_conv2d_weight_grad_scratch_floats(int64_t batch, int32_t cin, int32_t h, int32_t w, int32_t cout, int32_t ksize,
                                          int32_t stride, int64_t* floats);
int dmb_conv2d_weight_grad(const float* x, const float* gy, float* dw, float* db, int64_t batch, int32_t cin, int32_t h,
                           int32_t w, int32_t cout, int32_t ksize, int32_t stride, const float* x_scale,
                           const float* x_shift, int32_t x_relu, const float* y_raw, const float* ga, const float* gb,
                           const float* gc, float* scratch, void* stream);
/* Launch a register-resident FMA loop (16 chains x iters per thread) and report the FLOPs it
 * performs in *flops_out_host; the caller times it with CUDA events to get the FP32 roof.    */
int dmb_bench_fp32_fma(int32_t blocks, int32_t threads, int32_t iters, float* scratch,
                       double* flops_out_host, void* stream);
/* Same for the conv kernels' register-tile shape (8x8 accumulators, acc[c][p] += w[c]*a[p+kx]; order 0 = pixel
 * loop innermost, 1 = channel loop innermost): the practical FFMA ceiling of that inner loop.     */
int dmb_bench_fma_tile(int32_t order, int32_t blocks, int32_t iters, float* scratch,
                       double* flops_out_host, void* stream);
/* ... and with the packed FFMA2 (fma.rn.f32x2, new on sm_100) form of the same tile.             */
int dmb_bench_fma2_tile(int32_t order, int32_t blocks, int32_t iters, float* scratch,
                        double* flops_out_host, void* stream);
/* The conv inner loop with its weight operand in shared memory (0), __constant__ memory (1) or the kernel
 * parameter block (2): measures what the register-bank conflicts of the shared-memory form cost.   */
int dmb_bench_fma_conv(int32_t variant, int32_t blocks, int32_t iters, float* scratch,
                       double* flops_out_host, void* stream);
/* Number of kernels this library has launched in the process (reset != 0 zeroes it).        */
long long dmb_launch_count(int reset);

/* ---- PCA projection of the latents (run_dim_reduction.py:53-92 process_PCA -> sklearn PCA.transform) ----------- */
/* out (n, n_components) = (x (n, latent_len) - mean) @ components (n_components, latent_len)^T, times
 * inv_scale[j] (= 1/sqrt(explained_variance_[j]) for a whitening PCA; NULL otherwise).  fp32.                      */
int dmb_pca_transform(const float* x, int64_t n, int32_t latent_len, const float* mean,
                      const float* components, int32_t n_components, const float* inv_scale, float* out,
                      void* stream);

/* The same projection on the tensor cores (tcgen05, 3xTF32 operand split, fp32 accumulation): x is read as the NHWC
 * input of a 1x1 convolution with latent_len input and 64 output channels per pass, mean folded into the bias.
 * `scratch`: dmb_pca_transform_scratch_floats(n, latent_len) floats, 16-byte aligned.  Rows past the last multiple of
 * 128 and shapes the tensor-core kernel does not take (latent_len % 32 != 0) run on the kernel above.               */
int dmb_pca_transform_scratch_floats(int64_t n, int32_t latent_len, int64_t* floats);
int dmb_pca_transform_tc(const float* x, int64_t n, int32_t latent_len, const float* mean,
                         const float* components, int32_t n_components, const float* inv_scale, float* out,
                         float* scratch, void* stream);

/* ---- training augmentation (run_training.py:396-403) ------------------------------------ */
/* out[b] = rot90(flip(x[b], dims=(flip,)), k=rot, dims=[1, 2]) for every sample in one launch; ops_dev holds one
 * byte per sample: flip in {0 none, 1 H, 2 W} | rot in {0..3} << 2 (drawn on the host in the reference's order).  */
int dmb_augment_batch(const float* x, const uint8_t* ops_dev, int64_t batch, int32_t channels,
                      int32_t height, int32_t width, float* out, void* stream);

/* ---- time-matching loss (vq_vae.py:324-332; vae.py:321-336, :442-457) --------------------- */
typedef struct dmb_time_matching {
    const float* mat;     /* (B, B) device tensor: pair weights (variant 0) or pair classes 0 / 1 / 2 (variant 1) */
    int32_t variant;      /* 0: VQ_VAE  loss = sum(sim * mat);  1: VQ_VAE_z16 / z32  weighted, hinged, mean        */
    float w_a, w_t, w_n;  /* variant 1: weight of class 2 / 1 / 0 pairs                                            */
    float margin;         /* variant 1: class-0 terms become max(sim * w_n + margin, 0)                             */
    float weight;         /* weight_matching: total_loss += weight * time_matching_loss                            */
} dmb_time_matching;
/* sim[i][j] = mean_l (z[i][l] - z[j][l])^2 over z (B, L); loss_out[0] <- loss.  `scratch` needs
 * dmb_time_matching_scratch_floats() floats (16-byte aligned) and keeps dloss/dsim for the backward.  */
int dmb_time_matching_scratch_floats(int64_t batch, int64_t latent_len, size_t* floats);
int dmb_time_matching_forward(const float* z, int64_t batch, int64_t latent_len, const dmb_time_matching* tm,
                              float* scratch, float* loss_out, void* stream);
/* grad_z (B, L) <- (accumulate != 0: +=) scale * dloss/dz, from the scratch of the forward call.      */
int dmb_time_matching_backward(const float* z, int64_t batch, int64_t latent_len, const float* scratch,
                               float scale, float* grad_z, int32_t accumulate, void* stream);

/* ---- training step (run_training.py:404-408) ---------------------------------------- */
/* Full forward in BATCH mode keeping activations, producing decoded and
 * losses_out float[4] = {recon_loss, commitment_loss, total_loss, perplexity}.           */
int dmb_train_forward(const dmb_model* m, const float* packed, const float* params,
                      const float* x, const float* mask, int32_t mask_channels,
                      const float* channel_var, int64_t batch, float* decoded,
                      float* losses_out, float* bnbuf_inout, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Gradient of total_loss w.r.t. every trainable tensor, written (not accumulated) to
 * `grads` in the layout of `params`.  Uses the workspace left by dmb_train_forward.       */
int dmb_train_backward(const dmb_model* m, const float* packed, const float* params,
                       const float* x, const float* mask, int32_t mask_channels,
                       const float* channel_var, const float* decoded, int64_t batch, float grad_scale,
                       float* grads, void* workspace, size_t workspace_bytes, void* stream);
/* The same two calls with the optional time-matching term (tm == NULL: identical to the above).  The pair
 * similarities are taken on z_before (VQ_VAE, VQ_VAE_z16) or on the quantised z_after (VQ_VAE_z32), as in the
 * reference; losses_out is float[8] = {recon, commitment, total, perplexity, time_matching, 0, 0, 0}.            */
int dmb_train_forward_tm(const dmb_model* m, const float* packed, const float* params,
                         const float* x, const float* mask, int32_t mask_channels,
                         const float* channel_var, int64_t batch, const dmb_time_matching* tm, float* decoded,
                         float* losses_out, float* bnbuf_inout, void* workspace,
                         size_t workspace_bytes, void* stream);
int dmb_train_backward_tm(const dmb_model* m, const float* packed, const float* params,
                          const float* x, const float* mask, int32_t mask_channels,
                          const float* channel_var, const float* decoded, int64_t batch,
                          const dmb_time_matching* tm, float grad_scale,
                          float* grads, void* workspace, size_t workspace_bytes, void* stream);
/* The same two calls for data-parallel training with SYNCHRONISED BatchNorm: every BatchNorm's per-channel sums
 * (forward: sum y, sum y^2; backward: sum g, sum g*y) are folded on the device into `2 * channels` doubles and handed
 * to `allreduce` (which must sum them in place across the ranks, ordered on `stream` -- e.g. ncclAllReduce /
 * torch.distributed.all_reduce), then finalised with the GLOBAL element count, so that `world` ranks with `batch`
 * patches each compute exactly the step a single process computes on the concatenated batch (the reference's
 * single-GPU semantics, run_training.py:404).  BatchNorm weight / bias gradients come out pre-divided by `world`
 * so that the usual allreduce-sum of `grads` followed by Adam's 1/world yields the global-batch gradient.
 * sync == NULL or sync->world <= 1: identical to the _tm calls.                                                     */
typedef int (*dmb_allreduce_fn)(void* user, double* sums_dev, int64_t n, void* stream);
typedef struct dmb_sync_bn {
    dmb_allreduce_fn allreduce;
    void* user;
    int32_t world;
} dmb_sync_bn;
int dmb_train_forward_sync(const dmb_model* m, const float* packed, const float* params,
                           const float* x, const float* mask, int32_t mask_channels,
                           const float* channel_var, int64_t batch, const dmb_time_matching* tm,
                           const dmb_sync_bn* sync, float* decoded, float* losses_out, float* bnbuf_inout,
                           void* workspace, size_t workspace_bytes, void* stream);
int dmb_train_backward_sync(const dmb_model* m, const float* packed, const float* params,
                            const float* x, const float* mask, int32_t mask_channels,
                            const float* channel_var, const float* decoded, int64_t batch,
                            const dmb_time_matching* tm, const dmb_sync_bn* sync, float grad_scale,
                            float* grads, void* workspace, size_t workspace_bytes, void* stream);
/* torch.optim.Adam (betas, eps, no weight decay), bias-corrected, step is 1-based.
 * grad_scale multiplies the gradient first (1/world_size after an allreduce-sum).        */
int dmb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  int64_t n, float lr, float beta1, float beta2, float eps, int32_t step,
                  float grad_scale, void* stream);

/* Same update with the step counter on the device: step_dev (int32, starts at 0) is incremented and the
 * bias corrections written to bc_dev (float[2]) by a one-thread kernel first, so a captured CUDA graph
 * can be replayed unchanged every step.                                                          */
int dmb_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, int32_t* step_dev, float* bc_dev,
                      float grad_scale, void* stream);
/* nn.BatchNorm2d.num_batches_tracked += 1 of a train-mode forward (torch/nn/modules/batchnorm.py; the reference's
 * model(batch) at run_training.py:404) for the n BatchNorm layers of the model, kept as n consecutive int64.        */
int dmb_bn_count_batch(int64_t* num_batches_tracked, int32_t n, void* stream);

/* ---- input staging (pipeline/train_utils.py:252-274) -------------------------------- */
/* zscore_patch: per patch and channel (x - mean) / (std + DBL_EPSILON), float64 or
 * float32 or uint16 input, float32 output.  in_dtype: 0 = f32, 1 = f64, 2 = u16.         */
int dmb_zscore_patch(const void* raw, int32_t in_dtype, int64_t planes, int32_t hw,
                     float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNAMORPH_B200_H */
