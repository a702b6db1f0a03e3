// PCA projection of the latent vectors: sklearn's PCA.transform as used by
// /root/reference/run_dim_reduction.py:53-92 (process_PCA):
//     out[n][j] = sum_l (x[n][l] - mean[l]) * components[j][l]        (optionally / sqrt(explained_variance[j]) when whitening)
// x = (N, L) latent vectors exactly as process_VAE writes them (L = D*h*w = 4096 for the default model), components
// (k, L).  Two forms: dmb_pca_transform_tc runs it on the tensor cores as a 3xTF32 1x1 convolution (conv_tc.cu; fp32
// round-off level, a single-pass TF32 product would cost three digits); dmb_pca_transform is the fp32 CUDA-core GEMM
// that also serves row tails and odd shapes: CTA tile 128 samples x 64 components x 16 latent columns, thread tile 8 x 8 (64 FFMA accumulators), both
// operand tiles staged K-major in shared memory, register-prefetched so the next tile's global loads overlap the FMAs.
#include <stdlib.h>

#include "common.cuh"

namespace dmb {
namespace {

constexpr int BM = 128, BN = 64, BK = 16;

__global__ void __launch_bounds__(128) pca_kernel(const float* __restrict__ x, int64_t n, int l,
                                                  const float* __restrict__ mean, const float* __restrict__ comp, int k,
                                                  const float* __restrict__ inv_scale, float* __restrict__ out) {
    pdl_wait();
    __shared__ __align__(16) float As[BK][BM + 4];      // [latent column][sample]
    __shared__ __align__(16) float Bs[BK][BN + 4];      // [latent column][component]
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tm = (tid >> 3) * 8;          // 16 x 8 threads: rows tm..tm+7, components tn..tn+7
    const int tn = (tid & 7) * 8;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loader roles: A tile = 128 rows x 16 columns = 512 float4 -> 4 per thread; B tile = 64 x 16 = 256 float4 -> 2
    float4 ra[4], rb[2];
    auto load = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = tid + i * 128;
            const int row = f >> 2, c4 = (f & 3) * 4;
            const int64_t gr = m0 + row;
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < n && k0 + c4 < l) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(x + gr * l + k0 + c4));
                const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + k0 + c4));
                ra[i] = make_float4(v.x - mu.x, v.y - mu.y, v.z - mu.z, v.w - mu.w);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 128;
            const int row = f >> 2, c4 = (f & 3) * 4;
            rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < k && k0 + c4 < l)
                rb[i] = __ldg(reinterpret_cast<const float4*>(comp + (size_t)(n0 + row) * l + k0 + c4));
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = tid + i * 128;
            const int row = f >> 2, c4 = (f & 3) * 4;
            As[c4][row] = ra[i].x; As[c4 + 1][row] = ra[i].y; As[c4 + 2][row] = ra[i].z; As[c4 + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 128;
            const int row = f >> 2, c4 = (f & 3) * 4;
            Bs[c4][row] = rb[i].x; Bs[c4 + 1][row] = rb[i].y; Bs[c4 + 2][row] = rb[i].z; Bs[c4 + 3][row] = rb[i].w;
        }
    };
    load(0);
    for (int k0 = 0; k0 < l; k0 += BK) {
        __syncthreads();
        stash();
        __syncthreads();
        if (k0 + BK < l) load(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tm]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tm + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tn + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gr = m0 + tm + i;
        if (gr >= n) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gc = n0 + tn + j;
            if (gc < k) out[gr * k + gc] = inv_scale ? acc[i][j] * __ldg(inv_scale + gc) : acc[i][j];
        }
    }
}

// ---- tensor-core form (tcgen05, 3xTF32): the projection is a 1x1 convolution over "pixels" = samples with
// Cin = latent_len and Cout = 64 components per pass -- exactly the shape conv_tc.cu runs for the 64-wide encoder
// (x viewed as NHWC (1, n/128, 128, latent_len)).  The mean is folded into the bias (-mean . c_j, summed in double) and
// the whitening factor into the weights.
__global__ void pca_pack_kernel(const float* __restrict__ comp, const float* __restrict__ mean,
                                const float* __restrict__ inv_scale, int k, int l, int j0, float* __restrict__ w_packed,
                                float* __restrict__ bias) {
    pdl_wait();
    // w_packed [l][64] (the [Cin][1][1][Cout] layout pack_tc_weights reads): component j0 + j in column j, zero beyond k
    const int total = l * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int li = i >> 6, j = i & 63;
        float v = 0.f;
        if (j0 + j < k) v = __ldg(comp + (size_t)(j0 + j) * l + li) * (inv_scale ? __ldg(inv_scale + j0 + j) : 1.f);
        w_packed[i] = v;
    }
    // bias: one warp per component, double accumulation, fixed order
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp < 64) {
        double acc = 0.0;
        if (j0 + warp < k)
            for (int li = lane; li < l; li += 32)
                acc += (double)__ldg(mean + li) * (double)__ldg(comp + (size_t)(j0 + warp) * l + li);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) bias[warp] = (j0 + warp < k) ? (float)(-acc * (inv_scale ? (double)inv_scale[j0 + warp] : 1.0)) : 0.f;
    }
}

__global__ void pca_compact_kernel(const float* __restrict__ y, int64_t rows, int k, int j0, float* __restrict__ out) {
    pdl_wait();
    const int cols = (k - j0 < 64) ? k - j0 : 64;
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int j = (int)(i - r * cols);
        out[r * k + j0 + j] = y[r * 64 + j];
    }
}

bool pca_tc_enabled() {
    const char* e = getenv("DMB_PCA_TC");
    return !(e && e[0] == '0');
}

int launch_pca_fma(const float* x, int64_t n, int l, const float* mean, const float* comp, int k, const float* inv_scale,
                   float* out, cudaStream_t st) {
    const int64_t gx = (n + BM - 1) / BM;
    DMB_CHECK(gx < (1ll << 31), "dmb_pca_transform: too many samples");
    DMB_LAUNCH((pca_kernel), dim3((unsigned)gx, (unsigned)((k + BN - 1) / BN)), 128, 0, st, x, n, l, mean, comp, k,
               inv_scale, out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace
}  // namespace dmb

using namespace dmb;

extern "C" int dmb_pca_transform_scratch_floats(int64_t n, int32_t latent_len, int64_t* floats) {
    DMB_CHECK(floats && n >= 0 && latent_len > 0, "dmb_pca_transform_scratch_floats: bad arguments");
    const int64_t rows = n / 128 * 128;
    *floats = (int64_t)latent_len * 64 + 64 + conv_tc_weight_floats(latent_len, 64, 1) + rows * 64 + 64;
    return 0;
}

// sklearn PCA.transform on the tensor cores; `scratch` holds dmb_pca_transform_scratch_floats() floats (16-byte aligned).
// Rows beyond the last multiple of 128 (and every shape the tensor-core kernel does not take) run on the FFMA kernel.
extern "C" int dmb_pca_transform_tc(const float* x, int64_t n, int32_t latent_len, const float* mean,
                                    const float* components, int32_t n_components, const float* inv_scale, float* out,
                                    float* scratch, void* stream) {
    DMB_CHECK(x && mean && components && out, "dmb_pca_transform_tc: null pointer");
    DMB_CHECK(latent_len > 0 && latent_len % 4 == 0, "dmb_pca_transform_tc: latent length %d must be a multiple of 4", latent_len);
    DMB_CHECK(n_components > 0, "dmb_pca_transform_tc: no components");
    DMB_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)mean & 15) == 0 && ((uintptr_t)components & 15) == 0,
              "dmb_pca_transform_tc: x, mean and components must be 16-byte aligned");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = n / 128 * 128;
    const bool tc = scratch && rows > 0 && pca_tc_enabled() && ((uintptr_t)scratch & 15) == 0 &&
                    conv_tc_supported(latent_len, 64, 1, 1, (int)(rows / 128), 128) && rows / 128 < (1 << 24);
    if (!tc) return launch_pca_fma(x, n, latent_len, mean, components, n_components, inv_scale, out, st);
    float* w_packed = scratch;
    float* bias = w_packed + (int64_t)latent_len * 64;
    float* wtc = bias + 64;
    float* y = wtc + conv_tc_weight_floats(latent_len, 64, 1);
    for (int j0 = 0; j0 < n_components; j0 += 64) {
        int blocks = (latent_len * 64 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (blocks < 8) blocks = 8;                         // 64 warps for the bias rows
        DMB_LAUNCH((pca_pack_kernel), blocks, 256, 0, st, components, mean, inv_scale, n_components, latent_len, j0,
                   w_packed, bias);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
        DMB_TRY(pack_tc_weights(w_packed, wtc, latent_len, 64, 1, st));
        ConvTcArgs a{};
        a.x = x; a.wtc = wtc; a.bias = bias; a.y = y; a.skip = nullptr;
        a.B = 1; a.Cin = latent_len; a.H = (int)(rows / 128); a.W = 128; a.Cout = 64; a.ks = 1; a.stride = 1;
        a.in_relu = 0; a.out_relu = 0; a.out_nhwc = 1; a.skip_nhwc = 1;
        DMB_TRY(conv_tc(a, st));
        int64_t cb = (rows * 64 + 255) / 256;
        if (cb > 148 * 16) cb = 148 * 16;
        DMB_LAUNCH((pca_compact_kernel), (unsigned)cb, 256, 0, st, y, rows, n_components, j0, out);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
    }
    if (rows < n)
        DMB_TRY(launch_pca_fma(x + rows * latent_len, n - rows, latent_len, mean, components, n_components, inv_scale,
                               out + rows * n_components, st));
    return 0;
}

extern "C" int dmb_pca_transform(const float* x, int64_t n, int32_t latent_len, const float* mean,
                                 const float* components, int32_t n_components, const float* inv_scale, float* out,
                                 void* stream) {
    DMB_CHECK(x && mean && components && out, "dmb_pca_transform: null pointer");
    DMB_CHECK(latent_len > 0 && latent_len % 4 == 0, "dmb_pca_transform: latent length %d must be a multiple of 4", latent_len);
    DMB_CHECK(n_components > 0, "dmb_pca_transform: no components");
    DMB_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)mean & 15) == 0 && ((uintptr_t)components & 15) == 0,
              "dmb_pca_transform: x, mean and components must be 16-byte aligned");
    if (n == 0) return 0;
    return launch_pca_fma(x, n, latent_len, mean, components, n_components, inv_scale, out, (cudaStream_t)stream);
}
