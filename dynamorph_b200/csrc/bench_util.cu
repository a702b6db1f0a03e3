// Measurement helper: FP32 FMA throughput of the device (the roof that bounds the conv kernels).
#include "common.cuh"

namespace dmb {
namespace {
__global__ void fma_peak_kernel(float* out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;   // never true; keeps the loop alive
}
}  // namespace
}  // namespace dmb

extern "C" int dmb_bench_fp32_fma(int32_t blocks, int32_t threads, int32_t iters, float* scratch,
                                  double* flops_out_host, void* stream) {
    DMB_CHECK(scratch && flops_out_host, "dmb_bench_fp32_fma: null pointer");
    dmb::fma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 0.001f);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    *flops_out_host = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    return 0;
}

// Register-tile FMA ceilings: the conv inner loop is acc[c][p] += w[c] * a[p + kx] over an 8x8 accumulator tile.
// order 0: pixel loop innermost (weight operand reused), order 1: channel loop innermost (activation reused).
namespace dmb {
namespace {
template <int ORDER>
__global__ void __launch_bounds__(128, 4) fma_tile_kernel(float* out, int iters, const float* __restrict__ src) {
    float acc[8][8], a[12], w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = src[threadIdx.x + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = src[64 + threadIdx.x + i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            if (ORDER == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int p = 0; p < 8; ++p) acc[c][p] = fmaf(w[c], a[p + kx], acc[c][p]);
            } else {
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c][p] = fmaf(w[c], a[p + kx], acc[c][p]);
            }
        }
        // perturb the operands a little so nothing folds away (cheap relative to 256 FMAs)
        a[it & 7] += 1e-9f;
        w[it & 7] -= 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) s += acc[c][p];
    if (s == 123.456f) out[0] = s;
}
}  // namespace
}  // namespace dmb

extern "C" int dmb_bench_fma_tile(int32_t order, int32_t blocks, int32_t iters, float* scratch,
                                  double* flops_out_host, void* stream) {
    DMB_CHECK(scratch && flops_out_host, "dmb_bench_fma_tile: null pointer");
    if (order == 0) dmb::fma_tile_kernel<0><<<blocks, 128, 0, (cudaStream_t)stream>>>(scratch, iters, scratch);
    else dmb::fma_tile_kernel<1><<<blocks, 128, 0, (cudaStream_t)stream>>>(scratch, iters, scratch);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    *flops_out_host = 2.0 * 256.0 * (double)iters * (double)blocks * 128.0;
    return 0;
}

// FFMA2 (packed fp32x2, new on sm_100): same 8x8 tile, channels paired -> 32 float2 accumulators,
// weights as natural float2 pairs, activations duplicated into both halves.
namespace dmb {
namespace {
template <int ORDER>
__global__ void __launch_bounds__(128, 4) fma2_tile_kernel(float* out, int iters, const float* __restrict__ src) {
    float2 acc[4][8], ad[12], w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 12; ++i) { const float v = src[threadIdx.x + i]; ad[i] = make_float2(v, v); }
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = make_float2(src[64 + threadIdx.x + 2 * i], src[65 + threadIdx.x + 2 * i]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            if (ORDER == 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int p = 0; p < 8; ++p) acc[c][p] = __ffma2_rn(w[c], ad[p + kx], acc[c][p]);
            } else {
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c][p] = __ffma2_rn(w[c], ad[p + kx], acc[c][p]);
            }
        }
        ad[it & 7].x += 1e-9f; ad[it & 7].y += 1e-9f;
        w[it & 3].x -= 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) s += acc[c][p].x + acc[c][p].y;
    if (s == 123.456f) out[0] = s;
}
}  // namespace
}  // namespace dmb

extern "C" int dmb_bench_fma2_tile(int32_t order, int32_t blocks, int32_t iters, float* scratch,
                                   double* flops_out_host, void* stream) {
    DMB_CHECK(scratch && flops_out_host, "dmb_bench_fma2_tile: null pointer");
    if (order == 0) dmb::fma2_tile_kernel<0><<<blocks, 128, 0, (cudaStream_t)stream>>>(scratch, iters, scratch);
    else dmb::fma2_tile_kernel<1><<<blocks, 128, 0, (cudaStream_t)stream>>>(scratch, iters, scratch);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    *flops_out_host = 2.0 * 256.0 * (double)iters * (double)blocks * 128.0;
    return 0;
}

// Where should the weight operand of the conv inner loop live?  Same per-thread work as conv_tma's 4x4 stride-2
// core (acc[8][8] += w[c] * a[2p + kx + 3], 24 activation registers refreshed from shared memory per (ci, ky)),
// weights read (0) from shared memory with 128-bit broadcast loads, (1) from __constant__ memory, (2) from the
// kernel parameter block.  Variants 1/2 let ptxas feed the FFMA from a uniform register / constant operand, which
// removes one vector-register read per FFMA (register-bank conflicts are what caps variant 0).
namespace dmb {
namespace {
constexpr int WTAB = 4096;                       // 16 input channels x 16 taps x 8 channels ... x2 channel groups
__constant__ float c_wtab[WTAB];
struct WParam { float w[WTAB]; };

template <int VARIANT>
__global__ void __launch_bounds__(128, 4) fma_conv_kernel(float* out, int iters, const float* __restrict__ src,
                                                          const __grid_constant__ WParam wp) {
    __shared__ __align__(16) float s_w[WTAB];
    __shared__ __align__(16) float s_a[128 * 24 + 64];
    for (int i = threadIdx.x; i < WTAB; i += 128) s_w[i] = src[i & 1023];
    for (int i = threadIdx.x; i < 128 * 24 + 64; i += 128) s_a[i] = src[i & 2047];
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = 0.f;
    const float* ap = s_a + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5) * 768;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                float av[24];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const float4 t = *reinterpret_cast<const float4*>(ap + ((ci + ky + it) & 3) * 8 + i * 128);
                    av[4 * i] = t.x; av[4 * i + 1] = t.y; av[4 * i + 2] = t.z; av[4 * i + 3] = t.w;
                }
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    float wv[8];
                    const int wi = ((ci * 4 + ky) * 4 + kx) * 16 + ((threadIdx.x >> 6) & 1) * 8;
                    if (VARIANT == 0) {
                        const float4 t0 = *reinterpret_cast<const float4*>(s_w + wi);
                        const float4 t1 = *reinterpret_cast<const float4*>(s_w + wi + 4);
                        wv[0] = t0.x; wv[1] = t0.y; wv[2] = t0.z; wv[3] = t0.w;
                        wv[4] = t1.x; wv[5] = t1.y; wv[6] = t1.z; wv[7] = t1.w;
                    } else if (VARIANT == 1) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) wv[c] = c_wtab[((ci * 4 + ky) * 4 + kx) * 16 + c];
                    } else {
#pragma unroll
                        for (int c = 0; c < 8; ++c) wv[c] = wp.w[((ci * 4 + ky) * 4 + kx) * 16 + c];
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c)
#pragma unroll
                        for (int p = 0; p < 8; ++p) acc[c][p] = fmaf(wv[c], av[2 * p + kx + 3], acc[c][p]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) s += acc[c][p];
    if (s == 123.456f) out[0] = s;
}
}  // namespace
}  // namespace dmb

extern "C" int dmb_bench_fma_conv(int32_t variant, int32_t blocks, int32_t iters, float* scratch,
                                  double* flops_out_host, void* stream) {
    DMB_CHECK(scratch && flops_out_host, "dmb_bench_fma_conv: null pointer");
    static dmb::WParam wp;      // zero weights are fine: only the instruction stream matters
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 0) dmb::fma_conv_kernel<0><<<blocks, 128, 0, st>>>(scratch, iters, scratch, wp);
    else if (variant == 1) dmb::fma_conv_kernel<1><<<blocks, 128, 0, st>>>(scratch, iters, scratch, wp);
    else dmb::fma_conv_kernel<2><<<blocks, 128, 0, st>>>(scratch, iters, scratch, wp);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    *flops_out_host = 2.0 * 64.0 * 4 * 4 * 8 * (double)iters * (double)blocks * 128.0;
    return 0;
}
