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
