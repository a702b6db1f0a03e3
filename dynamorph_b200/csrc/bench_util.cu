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
