// Time-matching loss of VQ_VAE.forward.
//
//   sim[i][j] = mean_l (z[i][l] - z[j][l])^2                        z = latents flattened to (B, L)
//   VQ_VAE      (HiddenStateExtractor/vq_vae.py:324-332):  loss = sum_ij sim[i][j] * mat[i][j]
//   VQ_VAE_z16 / VQ_VAE_z32 (vae.py:321-336, :442-457):    w = {2: w_a, 1: w_t, 0: w_n}[mat];  t = sim * w;
//                                                          t[mat == 0] = max(t + margin, 0);    loss = mean_ij t
//
// The reference materialises the (B, B, L) difference tensor (1 GiB at B = 256, L = 4096).  Here the pair sums are
// tiled: a CTA owns a 32x32 block of (i, j) pairs and one slice of L, stages 32+32 latent rows of 64 columns in shared
// memory, every thread accumulates a 2x2 block of pairs; the L slices are folded in a fixed order (deterministic).
// The backward is  dL/dz[i][l] = (2/L) * sum_j (G[i][j] + G[j][i]) * (z[i][l] - z[j][l]),  G = dloss/dsim.
#include "common.cuh"

namespace dmb {
namespace {

constexpr int TP = 32;        // pairs tile edge
constexpr int TL = 64;        // latent columns per stage

__global__ void __launch_bounds__(256) tm_pair_kernel(const float* __restrict__ z, int B, int64_t L, int64_t lslice,
                                                      float* __restrict__ part) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ float zi[TP][TL + 1], zj[TP][TL + 1];
    const int nt = (B + TP - 1) / TP;
    const int ti = blockIdx.x / nt, tj = blockIdx.x % nt;
    const int64_t l0 = (int64_t)blockIdx.y * lslice;
    const int64_t l1 = (l0 + lslice < L) ? l0 + lslice : L;
    const int tid = threadIdx.x;
    const int pi = (tid >> 4) * 2, pj = (tid & 15) * 2;     // this thread's 2x2 pairs
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int64_t lc = l0; lc < l1; lc += TL) {
        __syncthreads();
        for (int e = tid; e < TP * TL; e += 256) {
            const int r = e / TL, cidx = e % TL;
            const int64_t l = lc + cidx;
            const int gi = ti * TP + r, gj = tj * TP + r;
            zi[r][cidx] = (gi < B && l < l1) ? __ldg(z + (size_t)gi * L + l) : 0.f;
            zj[r][cidx] = (gj < B && l < l1) ? __ldg(z + (size_t)gj * L + l) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int cidx = 0; cidx < TL; ++cidx) {
            const float a0 = zi[pi][cidx], a1 = zi[pi + 1][cidx];
            const float b0 = zj[pj][cidx], b1 = zj[pj + 1][cidx];
            float d;
            d = b0 - a0; acc[0][0] = fmaf(d, d, acc[0][0]);
            d = b1 - a0; acc[0][1] = fmaf(d, d, acc[0][1]);
            d = b0 - a1; acc[1][0] = fmaf(d, d, acc[1][0]);
            d = b1 - a1; acc[1][1] = fmaf(d, d, acc[1][1]);
        }
    }
    float* dst = part + (size_t)blockIdx.y * B * B;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int gi = ti * TP + pi + u, gj = tj * TP + pj + v;
            if (gi < B && gj < B) dst[(size_t)gi * B + gj] = acc[u][v];
        }
}

// sim -> per-pair loss term and G = weight * dloss/dsim; per-CTA partial loss sums (double), folded by tm_fold_kernel
__global__ void __launch_bounds__(256) tm_loss_kernel(const float* __restrict__ part, int nsplit, int B, int64_t L,
                                                      const float* __restrict__ mat, int variant, float w_a, float w_t,
                                                      float w_n, float margin, float* __restrict__ G,
                                                      double* __restrict__ loss_part) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ double red[8];
    const int64_t n = (int64_t)B * B;
    const float inv_l = 1.f / (float)L;
    const float inv_n = 1.f / (float)n;
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * n + e];
        const float sim = s * inv_l;
        const float m = __ldg(mat + e);
        float term, g;
        if (variant == 0) {
            term = sim * m; g = m;
        } else {
            const float w = (m == 2.f) ? w_a : ((m == 1.f) ? w_t : ((m == 0.f) ? w_n : m));
            term = sim * w; g = w;
            if (m == 0.f) {
                const float h = term + margin;
                if (h > 0.f) term = h; else { term = 0.f; g = 0.f; }
            }
            g *= inv_n;
        }
        if (G) G[e] = g;
        acc += (double)term;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        loss_part[blockIdx.x] = t;
    }
}

__global__ void tm_fold_kernel(const double* __restrict__ loss_part, int n, int variant, double inv_pairs,
                               float* __restrict__ loss_out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += loss_part[i];
    loss_out[0] = (float)(variant == 0 ? t : t * inv_pairs);
}

// g[i][l] (+)= scale * (2/L) * sum_j (G[i][j] + G[j][i]) * (z[i][l] - z[j][l]);  CTA = 8 rows i x 256 columns l
__global__ void __launch_bounds__(256) tm_grad_kernel(const float* __restrict__ z, const float* __restrict__ G, int B,
                                                      int64_t L, float scale, float* __restrict__ g, int accumulate) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    constexpr int TI = 8;
    extern __shared__ float S[];                 // [TI][B]
    const int i0 = blockIdx.y * TI;
    const int64_t l = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (int e = threadIdx.x; e < TI * B; e += 256) {
        const int r = e / B, j = e - r * B;
        const int i = i0 + r;
        S[e] = (i < B) ? (__ldg(G + (size_t)i * B + j) + __ldg(G + (size_t)j * B + i)) : 0.f;
    }
    __syncthreads();
    if (l >= L) return;
    float zi[TI], acc[TI];
#pragma unroll
    for (int r = 0; r < TI; ++r) {
        zi[r] = (i0 + r < B) ? __ldg(z + (size_t)(i0 + r) * L + l) : 0.f;
        acc[r] = 0.f;
    }
#pragma unroll 4
    for (int j = 0; j < B; ++j) {
        const float zj = __ldg(z + (size_t)j * L + l);
#pragma unroll
        for (int r = 0; r < TI; ++r) acc[r] = fmaf(S[r * B + j], zi[r] - zj, acc[r]);
    }
    const float k = scale * 2.f / (float)L;
#pragma unroll
    for (int r = 0; r < TI; ++r) {
        if (i0 + r >= B) break;
        float* dst = g + (size_t)(i0 + r) * L + l;
        *dst = accumulate ? (*dst + k * acc[r]) : k * acc[r];
    }
}

}  // namespace

int tm_splits(int64_t B, int64_t L) {
    const int64_t nt = (B + TP - 1) / TP;
    int64_t s = (2 * 148 + nt * nt - 1) / (nt * nt);
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    const int64_t chunks = (L + TL - 1) / TL;
    if (s > chunks) s = chunks;
    return (int)s;
}
// scratch floats: pair partials + G; doubles: loss partials
size_t tm_scratch_floats(int64_t B, int64_t L) { return (size_t)(tm_splits(B, L) + 1) * B * B + 2 * 296 + 8; }

// loss_out[0] <- loss.  scratch: tm_scratch_floats(B, L) floats (16-byte aligned); G is kept in it for the backward.
int tm_forward(const float* z, int64_t B, int64_t L, const dmb_time_matching& tm, float* scratch, float* loss_out,
               cudaStream_t st) {
    DMB_CHECK(z && tm.mat && scratch && loss_out, "time matching: null pointer");
    DMB_CHECK(B >= 1 && B <= 46340 && L >= 1, "time matching: bad shape B=%lld L=%lld", (long long)B, (long long)L);
    const int ns = tm_splits(B, L);
    const int64_t chunks = (L + TL - 1) / TL;
    const int64_t lslice = ((chunks + ns - 1) / ns) * TL;
    float* part = scratch;
    float* G = scratch + (size_t)ns * B * B;
    size_t lp_off = (size_t)(ns + 1) * B * B;
    lp_off += lp_off & 1;                       // 8-byte alignment of the double partials
    double* lp = reinterpret_cast<double*>(scratch + lp_off);
    const int nt = (int)((B + TP - 1) / TP);
    DMB_LAUNCH((tm_pair_kernel), dim3(nt * nt, ns), 256, 0, st, z, (int)B, L, lslice, part);
    DMB_CUDA(cudaGetLastError());
    int nblk = (int)((B * B + 1023) / 1024);
    if (nblk > 296) nblk = 296;
    DMB_LAUNCH((tm_loss_kernel), nblk, 256, 0, st, part, ns, (int)B, L, tm.mat, tm.variant, tm.w_a, tm.w_t, tm.w_n, tm.margin, G, lp);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCH((tm_fold_kernel), 1, 1, 0, st, lp, nblk, tm.variant, 1.0 / ((double)B * (double)B), loss_out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(3);
    return 0;
}

const float* tm_G(const float* scratch, int64_t B, int64_t L) { return scratch + (size_t)tm_splits(B, L) * B * B; }

// g (B, L) <- (or +=) scale * d loss / d z, using the G left in scratch by tm_forward
int tm_backward(const float* z, int64_t B, int64_t L, const float* scratch, float scale, float* g, int accumulate,
                cudaStream_t st) {
    DMB_CHECK(z && scratch && g, "time matching backward: null pointer");
    const size_t smem = (size_t)8 * B * sizeof(float);
    DMB_CHECK(smem <= 200 * 1024, "time matching backward: batch %lld too large", (long long)B);
    if (smem > 48 * 1024)
        DMB_CUDA(cudaFuncSetAttribute(tm_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMB_LAUNCH((tm_grad_kernel), dim3((unsigned)((L + 255) / 256), (unsigned)((B + 7) / 8)), 256, smem, st, z, tm_G(scratch, B, L), (int)B, L, scale, g, accumulate);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace dmb

extern "C" {

int dmb_time_matching_scratch_floats(int64_t batch, int64_t latent_len, size_t* floats) {
    DMB_CHECK(floats && batch >= 1 && latent_len >= 1, "dmb_time_matching_scratch_floats: bad arguments");
    *floats = dmb::tm_scratch_floats(batch, latent_len);
    return 0;
}

int dmb_time_matching_forward(const float* z, int64_t batch, int64_t latent_len, const dmb_time_matching* tm,
                              float* scratch, float* loss_out, void* stream) {
    DMB_CHECK(tm, "dmb_time_matching_forward: null descriptor");
    return dmb::tm_forward(z, batch, latent_len, *tm, scratch, loss_out, (cudaStream_t)stream);
}

int dmb_time_matching_backward(const float* z, int64_t batch, int64_t latent_len, const float* scratch, float scale,
                               float* grad_z, int32_t accumulate, void* stream) {
    return dmb::tm_backward(z, batch, latent_len, scratch, scale, grad_z, accumulate, (cudaStream_t)stream);
}

}  // extern "C"
