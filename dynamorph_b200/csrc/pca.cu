// PCA projection of the latent vectors: sklearn's PCA.transform as used by
// /root/reference/run_dim_reduction.py:53-92 (process_PCA):
//     out[n][j] = sum_l (x[n][l] - mean[l]) * components[j][l]        (optionally / sqrt(explained_variance[j]) when whitening)
// x = (N, L) latent vectors exactly as process_VAE writes them (L = D*h*w = 4096 for the default model), components
// (k, L).  fp32 on CUDA cores (the projection feeds clustering; a single-pass TF32 tensor-core product would cost three
// digits): CTA tile 128 samples x 64 components x 16 latent columns, thread tile 8 x 8 (64 FFMA accumulators), both
// operand tiles staged K-major in shared memory, register-prefetched so the next tile's global loads overlap the FMAs.
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

}  // namespace
}  // namespace dmb

extern "C" int dmb_pca_transform(const float* x, int64_t n, int32_t latent_len, const float* mean,
                                 const float* components, int32_t n_components, const float* inv_scale, float* out,
                                 void* stream) {
    DMB_CHECK(x && mean && components && out, "dmb_pca_transform: null pointer");
    DMB_CHECK(latent_len > 0 && latent_len % 4 == 0, "dmb_pca_transform: latent length %d must be a multiple of 4", latent_len);
    DMB_CHECK(n_components > 0, "dmb_pca_transform: no components");
    DMB_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)mean & 15) == 0 && ((uintptr_t)components & 15) == 0,
              "dmb_pca_transform: x, mean and components must be 16-byte aligned");
    if (n == 0) return 0;
    const int64_t gx = (n + dmb::BM - 1) / dmb::BM;
    DMB_CHECK(gx < (1ll << 31), "dmb_pca_transform: too many samples");
    DMB_LAUNCH((dmb::pca_kernel), dim3((unsigned)gx, (unsigned)((n_components + dmb::BN - 1) / dmb::BN)), 128, 0, stream,
               x, n, latent_len, mean, components, n_components, inv_scale, out);
    DMB_LAUNCHED(1);
    return 0;
}
