// BatchNorm bookkeeping kernels: statistics finalisation and the residual merge.
// Reference semantics: nn.BatchNorm2d (eps 1e-5, momentum 0.1, biased variance to normalise,
// unbiased to track) as used at HiddenStateExtractor/vq_vae.py:205,208,279-288.
#include "common.cuh"

namespace dmb {
namespace {

// Fixed-order block reduction of two doubles (256 threads): lane tree, then the eight warp sums in order.
__device__ __forceinline__ void block_sum2(double& s, double& q, double (*red)[2]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) { red[warp][0] = s; red[warp][1] = q; }
    __syncthreads();
    s = 0.0; q = 0.0;
    for (int w = 0; w < nw; ++w) { s += red[w][0]; q += red[w][1]; }
}

// BATCH mode: one CTA per channel, every thread sums B*nbands/256 partials with independent loads (the one-warp
// version below spent ~10 us per launch walking up to 2048 partials as a latency chain).
__global__ void __launch_bounds__(256) bn_finalize_batch_kernel(const BnFinalizeArgs a) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ double red[8][2];
    const int c = blockIdx.x;
    const int n = a.rows > 0 ? a.rows : a.B * a.nbands;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double2 v = *reinterpret_cast<const double2*>(a.partials + ((size_t)i * a.C + c) * 2);
        s += v.x; q += v.y;
    }
    block_sum2(s, q, red);
    if (threadIdx.x) return;
    const double cnt = (double)a.count_per_sample * a.B;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = a.gamma[c] * invstd;
    a.scale[c] = sc;
    a.shift[c] = a.beta[c] - (float)mean * sc;
    if (a.save_mean) { a.save_mean[c] = (float)mean; a.save_invstd[c] = invstd; }
    if (a.running_mean) {
        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (float)mean;
        a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
    }
}

// PER_SAMPLE mode: one THREAD per (sample, channel) -- there are only nbands (1..8) partials to add, and B*C outputs
// (131 k for an 8192-patch chunk); a warp per output left 31 lanes idle and cost 40-80 us per launch.
__global__ void __launch_bounds__(256) bn_finalize_sample_kernel(const BnFinalizeArgs a) {
    pdl_wait();
    const int64_t o = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (o >= (int64_t)a.B * a.C) return;
    const int c = (int)(o % a.C);
    const int64_t bs = o / a.C;
    double s = 0.0, q = 0.0;
    for (int band = 0; band < a.nbands; ++band) {
        const double2 v = *reinterpret_cast<const double2*>(a.partials + ((bs * a.nbands + band) * a.C + c) * 2);
        s += v.x; q += v.y;
    }
    const double cnt = (double)a.count_per_sample;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = a.gamma[c] * invstd;
    a.scale[o] = sc;
    a.shift[o] = a.beta[c] - (float)mean * sc;
    if (a.save_mean) { a.save_mean[o] = (float)mean; a.save_invstd[o] = invstd; }
}

// one warp per output (channel, or sample*channel): fixed-order sum of the per-CTA partials
__global__ void bn_finalize_kernel(const BnFinalizeArgs a) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nout = a.per_sample ? a.B * a.C : a.C;
    if (warp >= nout) return;
    const int c = warp % a.C;
    const int bs = a.per_sample ? warp / a.C : 0;
    const int nb = a.per_sample ? 1 : a.B;
    const int n = nb * a.nbands;
    double s = 0.0, q = 0.0;
    for (int i = lane; i < n; i += 32) {
        const int bb = bs + i / a.nbands, band = i % a.nbands;
        const double* p = a.partials + (((size_t)bb * a.nbands + band) * a.C + c) * 2;
        s += p[0]; q += p[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane) return;
    const double cnt = (double)a.count_per_sample * nb;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = a.gamma[c] * invstd;
    a.scale[warp] = sc;
    a.shift[warp] = a.beta[c] - (float)mean * sc;
    if (a.save_mean) { a.save_mean[warp] = (float)mean; a.save_invstd[warp] = invstd; }
    if (!a.per_sample && a.running_mean) {
        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (float)mean;
        a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unbiased;
    }
}

__global__ void affine_add_kernel(const AffineAddArgs a, int64_t total4) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const int hw4 = a.HW >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t plane = i / hw4;            // b*C + c
        const int c = (int)(plane % a.C);
        const int64_t t = a.per_sample ? plane : c;
        float4 va = __ldg(reinterpret_cast<const float4*>(a.a) + i);
        float4 vb = a.b ? __ldg(reinterpret_cast<const float4*>(a.b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.sa) {
            const float s = a.sa[t], sh = a.ta[t];
            va.x = fmaf(va.x, s, sh); va.y = fmaf(va.y, s, sh); va.z = fmaf(va.z, s, sh); va.w = fmaf(va.w, s, sh);
        }
        if (a.b && a.sb) {
            const float s = a.sb[t], sh = a.tb[t];
            vb.x = fmaf(vb.x, s, sh); vb.y = fmaf(vb.y, s, sh); vb.z = fmaf(vb.z, s, sh); vb.w = fmaf(vb.w, s, sh);
        }
        reinterpret_cast<float4*>(a.out)[i] = make_float4(va.x + vb.x, va.y + vb.y, va.z + vb.z, va.w + vb.w);
    }
}

}  // namespace

int bn_finalize(const BnFinalizeArgs& a, cudaStream_t st) {
    DMB_CHECK(a.rows == 0 || !a.per_sample, "bn_finalize: a row count belongs to whole-batch statistics");
    if (!a.per_sample && (a.rows > 0 || (int64_t)a.B * a.nbands >= 64)) {
        DMB_LAUNCH((bn_finalize_batch_kernel), a.C, 256, 0, st, a);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
        return 0;
    }
    const int64_t nout = a.per_sample ? (int64_t)a.B * a.C : a.C;
    if (a.per_sample && a.nbands <= 16) {
        DMB_LAUNCH((bn_finalize_sample_kernel), (unsigned)((nout + 255) / 256), 256, 0, st, a);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
        return 0;
    }
    const int threads = 128;
    const int64_t blocks = (nout * 32 + threads - 1) / threads;
    DMB_LAUNCH((bn_finalize_kernel), (unsigned)blocks, threads, 0, st, a);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int affine_add(const AffineAddArgs& a, cudaStream_t st) {
    DMB_CHECK(a.HW % 4 == 0, "affine_add: HW must be a multiple of 4");
    const int64_t total4 = a.B * a.C * (a.HW >> 2);
    int64_t blocks = (total4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    DMB_LAUNCH((affine_add_kernel), (unsigned)blocks, 256, 0, st, a, total4);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace dmb

// ---------------------------------------------------------------------------------------
// backward bookkeeping
// ---------------------------------------------------------------------------------------
namespace dmb {
namespace {

__global__ void bn_bwd_finalize_kernel(const BnBwdArgs a) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nout = a.per_sample ? a.B * a.C : a.C;
    if (warp >= nout) return;
    const int c = warp % a.C;
    const int bs = a.per_sample ? warp / a.C : 0;
    const int nb = a.per_sample ? 1 : a.B;
    const int n = nb * a.nbands;
    double sg = 0.0, sgy = 0.0;
    for (int i = lane; i < n; i += 32) {
        const int bb = bs + i / a.nbands, band = i % a.nbands;
        const double* p = a.partials + (((size_t)bb * a.nbands + band) * a.C + c) * 2;
        sg += p[0]; sgy += p[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sgy += __shfl_xor_sync(0xffffffffu, sgy, o);
    }
    if (lane) return;
    const double N = (double)a.count_per_sample * nb;
    const double mu = a.mean[warp], is = a.invstd[warp], g = a.gamma[c];
    const double dbeta = sg;
    const double dgamma = is * (sgy - mu * sg);          // sum g * xhat
    // dL/dy = g*is*[ gr - dbeta/N - xhat*dgamma/N ],  xhat = (y - mu)*is
    a.A[warp] = (float)(g * is);
    a.Bc[warp] = (float)(-g * is * is * dgamma / N);
    a.Cc[warp] = (float)(-g * is * dbeta / N + g * is * is * mu * dgamma / N);
    const double div = a.grad_div > 1 ? (double)a.grad_div : 1.0;
    if (a.per_sample) {
        atomicAdd(a.dgamma + c, (float)dgamma);
        atomicAdd(a.dbeta + c, (float)dbeta);
    } else {
        a.dgamma[c] = (float)(dgamma / div);
        a.dbeta[c] = (float)(dbeta / div);
    }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_batch_kernel(const BnBwdArgs a) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ double red[8][2];
    const int c = blockIdx.x;
    const int n = a.rows > 0 ? a.rows : a.B * a.nbands;
    double sg = 0.0, sgy = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double2 v = *reinterpret_cast<const double2*>(a.partials + ((size_t)i * a.C + c) * 2);
        sg += v.x; sgy += v.y;
    }
    block_sum2(sg, sgy, red);
    if (threadIdx.x) return;
    const double N = (double)a.count_per_sample * a.B;
    const double mu = a.mean[c], is = a.invstd[c], g = a.gamma[c];
    const double dbeta = sg;
    const double dgamma = is * (sgy - mu * sg);
    a.A[c] = (float)(g * is);
    a.Bc[c] = (float)(-g * is * is * dgamma / N);
    a.Cc[c] = (float)(-g * is * dbeta / N + g * is * is * mu * dgamma / N);
    const double div = a.grad_div > 1 ? (double)a.grad_div : 1.0;
    a.dgamma[c] = (float)(dgamma / div);
    a.dbeta[c] = (float)(dbeta / div);
}

__global__ void __launch_bounds__(256) fold_partials_kernel(const double* partials, int64_t n, int C, double* out) {
    pdl_wait();
    __shared__ double red[8][2];
    const int c = blockIdx.x;
    double s = 0.0, q = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const double2 v = *reinterpret_cast<const double2*>(partials + ((size_t)i * C + c) * 2);
        s += v.x; q += v.y;
    }
    block_sum2(s, q, red);
    if (threadIdx.x == 0) { out[2 * c] = s; out[2 * c + 1] = q; }
}

__global__ void __launch_bounds__(256) sum_partials_block_kernel(const double* partials, int n, int C, float* out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ double red[8][2];
    const int c = blockIdx.x;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partials[((size_t)i * C + c) * 2];
    block_sum2(s, q, red);
    if (threadIdx.x == 0) out[c] = (float)s;
}

__global__ void sum_partials_kernel(const double* partials, int n, int C, float* out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= C) return;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += partials[((size_t)i * C + warp) * 2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[warp] = (float)s;
}

}  // namespace

int bn_backward_finalize(const BnBwdArgs& a, cudaStream_t st) {
    DMB_CHECK(a.rows == 0 || !a.per_sample, "bn_backward_finalize: a row count belongs to whole-batch statistics");
    if (!a.per_sample && (a.rows > 0 || (int64_t)a.B * a.nbands >= 64)) {
        DMB_LAUNCH((bn_bwd_finalize_batch_kernel), a.C, 256, 0, st, a);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
        return 0;
    }
    const int64_t nout = a.per_sample ? (int64_t)a.B * a.C : a.C;
    const int threads = 128;
    DMB_LAUNCH((bn_bwd_finalize_kernel), (unsigned)((nout * 32 + threads - 1) / threads), threads, 0, st, a);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int fold_partials(const double* partials, int64_t n, int C, double* out, cudaStream_t st) {
    DMB_LAUNCH((fold_partials_kernel), C, 256, 0, st, partials, n, C, out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int sum_partials(const double* partials, int B, int nbands, int C, float* out, cudaStream_t st) {
    if ((int64_t)B * nbands >= 64) DMB_LAUNCH((sum_partials_block_kernel), C, 256, 0, st, partials, B * nbands, C, out);
    else DMB_LAUNCH((sum_partials_kernel), (C * 32 + 127) / 128, 128, 0, st, partials, B * nbands, C, out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace dmb
