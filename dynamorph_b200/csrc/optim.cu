// Fused flat Adam (torch.optim.Adam semantics as run_training.py:485 constructs it:
// betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) and the per-patch z-score of
// pipeline/train_utils.py:252-274.
#include <float.h>

#include "common.cuh"

namespace dmb {
namespace {

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt, float gscale) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // torch (_single_tensor_adam): exp_avg.lerp_(grad, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2);
    // denom = sqrt(exp_avg_sq)/sqrt(bc2) + eps; param.addcdiv_(exp_avg, denom, value=-lr/bc1)
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = m[i] + (1.f - b1) * (gi - m[i]);
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

// device-resident step counter variant (CUDA-graph replay: nothing in the launch changes between steps)
__global__ void adam_tick_kernel(int* step, float* bc, float b1, float b2) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const int t = ++step[0];
    bc[0] = (float)(1.0 - pow((double)b1, (double)t));
    bc[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}
// BatchNorm2d.num_batches_tracked += 1 for every BatchNorm of the model (one int64 each, flat)
__global__ void bn_count_kernel(long long* nbt, int n) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nbt[i] += 1;
}
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                const float* __restrict__ bc, float gscale) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const float step_size = lr / bc[0];
    const float bc2_sqrt = bc[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = m[i] + (1.f - b1) * (gi - m[i]);
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
}

// one CTA per (patch, channel) plane: mean / population std in double, output float32
template <typename T>
__global__ void zscore_kernel(const T* __restrict__ raw, int hw, float* __restrict__ out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ double red[2][32];
    const T* src = raw + (size_t)blockIdx.x * hw;
    float* dst = out + (size_t)blockIdx.x * hw;
    double s = 0.0;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) s += (double)src[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[0][w];
    const double mean = tot / hw;
    double q = 0.0;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) { const double d = (double)src[i] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = q;
    __syncthreads();
    double tq = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tq += red[1][w];
    const double denom = sqrt(tq / hw) + DBL_EPSILON;     // np.std (ddof 0) + np.finfo(float).eps
    for (int i = threadIdx.x; i < hw; i += blockDim.x) dst[i] = (float)(((double)src[i] - mean) / denom);
}

}  // namespace
}  // namespace dmb

extern "C" {

int dmb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale, void* stream) {
    DMB_CHECK(params && grads && exp_avg && exp_avg_sq, "dmb_adam_step: null pointer");
    DMB_CHECK(step >= 1, "dmb_adam_step: step is 1-based");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    DMB_LAUNCH((dmb::adam_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int dmb_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, int32_t* step_dev, float* bc_dev,
                      float grad_scale, void* stream) {
    DMB_CHECK(params && grads && exp_avg && exp_avg_sq && step_dev && bc_dev, "dmb_adam_step_dev: null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DMB_LAUNCH((dmb::adam_tick_kernel), 1, 1, 0, st, step_dev, bc_dev, beta1, beta2);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    DMB_LAUNCH((dmb::adam_dev_kernel), (unsigned)blocks, 256, 0, st, params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc_dev, grad_scale);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int dmb_bn_count_batch(int64_t* num_batches_tracked, int32_t n, void* stream) {
    DMB_CHECK(num_batches_tracked || n == 0, "dmb_bn_count_batch: null pointer");
    if (n <= 0) return 0;
    DMB_LAUNCH((dmb::bn_count_kernel), (n + 63) / 64, 64, 0, (cudaStream_t)stream, reinterpret_cast<long long*>(num_batches_tracked), n);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int dmb_zscore_patch(const void* raw, int32_t in_dtype, int64_t planes, int32_t hw, float* out, void* stream) {
    DMB_CHECK(raw && out, "dmb_zscore_patch: null pointer");
    DMB_CHECK(planes >= 0 && planes < (1ll << 31) && hw > 0, "dmb_zscore_patch: bad shape");
    if (planes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (in_dtype) {
        case 0: DMB_LAUNCH((dmb::zscore_kernel<float>), (unsigned)planes, 256, 0, st, (const float*)raw, hw, out); break;
        case 1: DMB_LAUNCH((dmb::zscore_kernel<double>), (unsigned)planes, 256, 0, st, (const double*)raw, hw, out); break;
        case 2: DMB_LAUNCH((dmb::zscore_kernel<uint16_t>), (unsigned)planes, 256, 0, st, (const uint16_t*)raw, hw, out); break;
        default: DMB_CHECK(false, "dmb_zscore_patch: in_dtype %d not in {0:f32,1:f64,2:u16}", in_dtype);
    }
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // extern "C"
