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
// direct convolution, forward (conv_fwd.cu)
// ------------------------------------------------------------------------------------
struct ConvFwdArgs {
    const float* x;         // (B, Cin, H, W)
    float* y;               // (B, Cout, Ho, Wo)
    const float* w;         // packed [Cin][KS][KS][Cout]
    const float* bias;      // [Cout] or [9][Cout] when bias_classes (row class*3 + col class)
    int bias_classes;       // 0 / 1
    // transform applied to x on load:  v = relu?(v * in_scale[c] + in_shift[c])
    const float* in_scale;  // nullptr = identity; [Cin] or [B][Cin]
    const float* in_shift;
    int in_per_sample;      // stride B over the affine table
    int in_relu;
    // epilogue
    const float* skip;      // (B, Cout, Ho, Wo) added to the output, or nullptr
    int out_relu;
    double* stats;          // [B][nbands][Cout][2] per-CTA (sum, sum of squares) or nullptr
    int B, Cin, H, W, Cout, Ho, Wo;
    int ks, stride;         // (1,1) (3,1) (4,2)
};
int conv_fwd(const ConvFwdArgs& a, cudaStream_t st);
// number of (sum, sumsq) partial rows per sample the kernel will write for this geometry
int conv_fwd_bands(int ks, int stride, int Cin, int Cout, int Ho, int Wo);

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
    int out_relu;
    double* stats;          // [B][nbands][Cout][2] or nullptr
    int B, Cin, H, W, Cout;
};
int convt_fwd(const ConvTFwdArgs& a, cudaStream_t st);
int convt_fwd_bands(int Cin, int Cout, int H, int W);

// ------------------------------------------------------------------------------------
// batch-norm bookkeeping (bn.cu)
// ------------------------------------------------------------------------------------
// Reduce per-CTA partials into per-channel (or per-sample-per-channel) scale/shift:
//   scale = gamma / sqrt(var + eps), shift = beta - mean * scale
// BATCH mode (per_sample = 0) also updates running stats and writes mean / invstd.
struct BnFinalizeArgs {
    const double* partials;  // [B][nbands][C][2]
    int B, nbands, C;
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

// out = (a*sa+ta) + (b*sb+tb), all (B, C, HW); the affine tables follow the per_sample flag.
struct AffineAddArgs {
    const float* a; const float* sa; const float* ta;   // sa == nullptr: identity
    const float* b; const float* sb; const float* tb;
    int per_sample;
    float* out;
    int64_t B; int C; int HW;
};
int affine_add(const AffineAddArgs& a, cudaStream_t st);

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

}  // namespace dmb
