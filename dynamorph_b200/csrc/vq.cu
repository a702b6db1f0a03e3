// Fused VectorQuantizer forward (HiddenStateExtractor/vq_vae.py:52-84, :90-116).
//
// One thread per latent position.  The codebook lives in shared memory, the position's D
// channel values in registers; squared distances use the reference's direct-difference
// form and reproduce torch's CPU summation order for the channel reduction (cascade sum:
// 16-element sequential runs folded level by level, no FMA contraction), so that on
// identical inputs the argmin is bit-identical to torch.argmax(-distances) -- first index
// wins ties.  The same pass gathers the code vector, forms z + (q - z), accumulates
// sum (q - z)^2 and the code histogram: the reference's (B,K,D,H,W) broadcast is never
// materialised.  HBM traffic is the algorithmic minimum: read z once, write z_st + idx.
#include "common.cuh"

namespace dmb {
namespace {

constexpr int VQ_THREADS = 128;
constexpr int KU = 4;   // codes processed together (independent add chains)

// torch's multi_row_sum (aten/src/ATen/native/cpu/SumKernel.cpp) for one output element:
// level_step = 16 for every size < 2^20.
template <int D>
__device__ __forceinline__ void dist_cascade(const float (&z)[D], const float* __restrict__ e, float& out) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int i = 0;
#pragma unroll
    for (int c0 = 0; c0 + 16 <= D; c0 += 16) {
#pragma unroll
        for (int c = c0; c < c0 + 16; c += 4) {
            const float4 ev = *reinterpret_cast<const float4*>(e + c);
            float d;
            d = __fsub_rn(z[c], ev.x);     acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
            d = __fsub_rn(z[c + 1], ev.y); acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
            d = __fsub_rn(z[c + 2], ev.z); acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
            d = __fsub_rn(z[c + 3], ev.w); acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
        }
        i = c0 + 16;
        acc1 = __fadd_rn(acc1, acc0); acc0 = 0.f;
        if ((i & (15 << 4)) == 0) {
            acc2 = __fadd_rn(acc2, acc1); acc1 = 0.f;
            if ((i & (15 << 8)) == 0) { acc3 = __fadd_rn(acc3, acc2); acc2 = 0.f; }
        }
    }
#pragma unroll
    for (int c = (D / 16) * 16; c < D; ++c) {
        const float d = __fsub_rn(z[c], e[c]);
        acc0 = __fadd_rn(acc0, __fmul_rn(d, d));
    }
    acc0 = __fadd_rn(acc0, acc1);
    acc0 = __fadd_rn(acc0, acc2);
    acc0 = __fadd_rn(acc0, acc3);
    out = acc0;
}

template <int D>
__global__ void __launch_bounds__(VQ_THREADS) vq_kernel(const VqArgs a, int kchunk) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    extern __shared__ __align__(16) float cb[];          // [kchunk][D]
    __shared__ int hist[1024];
    __shared__ double red[VQ_THREADS / 32];
    const int tid = threadIdx.x;
    const int64_t n = (int64_t)blockIdx.x * VQ_THREADS + tid;
    const int64_t total = a.B * a.P;
    const bool live = n < total;
    const int64_t b = live ? n / a.P : 0;
    const int pos = live ? (int)(n - b * a.P) : 0;

    float z[D];
    if (live) {
        const size_t base = (size_t)b * D * a.P + pos;
        if (a.pre_b) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const size_t t = (a.pre_per_sample ? (size_t)b * D : 0) + c;
                float va = __ldg(a.pre_a + base + (size_t)c * a.P);
                float vb = __ldg(a.pre_b + base + (size_t)c * a.P);
                if (a.pre_sa) va = fmaf(va, a.pre_sa[t], a.pre_ta[t]);
                if (a.pre_sb) vb = fmaf(vb, a.pre_sb[t], a.pre_tb[t]);
                z[c] = va + vb;
                if (a.z_before_out) a.z_before_out[base + (size_t)c * a.P] = z[c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) z[c] = __ldg(a.z + base + (size_t)c * a.P);
        }
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) z[c] = 0.f;
    }

    float best = 0.f;
    int bi = -1;
    for (int k0 = 0; k0 < a.K; k0 += kchunk) {
        const int kc = min(kchunk, a.K - k0);
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(a.codebook + (size_t)k0 * D);
            float4* dst = reinterpret_cast<float4*>(cb);
            for (int i = tid; i < kc * D / 4; i += VQ_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        int k = 0;
        for (; k + KU <= kc; k += KU) {
            float d[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) dist_cascade<D>(z, cb + (k + u) * D, d[u]);
#pragma unroll
            for (int u = 0; u < KU; ++u)
                if (bi < 0 || d[u] < best) { best = d[u]; bi = k0 + k + u; }
        }
        for (; k < kc; ++k) {
            float d;
            dist_cascade<D>(z, cb + k * D, d);
            if (bi < 0 || d < best) { best = d; bi = k0 + k; }
        }
    }

    // gather + straight-through value + loss partial
    double lsum = 0.0;
    if (live) {
        const float* e = a.codebook + (size_t)bi * D;
        const size_t base = (size_t)b * D * a.P + pos;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float q = __ldg(e + c);
            const float diff = __fsub_rn(q, z[c]);
            if (a.z_st) a.z_st[base + (size_t)c * a.P] = __fadd_rn(z[c], diff);
            lsum += (double)diff * (double)diff;
        }
        if (a.idx) a.idx[n] = bi;
    }
    if (a.stats) {
        for (int i = tid; i < a.K; i += VQ_THREADS) hist[i] = 0;
        __syncthreads();
        if (live) atomicAdd(&hist[bi], 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((tid & 31) == 0) red[tid >> 5] = lsum;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < VQ_THREADS / 32; ++w) s += red[w];
            atomicAdd(a.stats + 0, s);
            const int64_t rem = total - (int64_t)blockIdx.x * VQ_THREADS;
            atomicAdd(a.stats + 1, (double)(rem < VQ_THREADS ? rem : VQ_THREADS));
        }
        for (int i = tid; i < a.K; i += VQ_THREADS)
            if (hist[i]) atomicAdd(a.stats + 2 + i, (double)hist[i]);   // integer-valued: order-free
    }
}

__global__ void vq_finalize_kernel(const double* stats, int d, int k, float beta, float* out2) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // loss = mse(q, z.detach()) + beta * mse(q.detach(), z) (vq_vae.py:74-76); perplexity :79-82
    __shared__ double red[32];
    const double npos = stats[1];
    double ent = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const float p = (float)(stats[2 + i] / npos);
        ent += (double)(p * logf(p + 1e-10f));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ent;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        const float mse = (float)(stats[0] / (npos * d));
        out2[0] = mse + beta * mse;
        out2[1] = expf(-(float)s);
    }
}

__global__ void vq_gather_kernel(const int32_t* idx, const float* cb, int64_t total, int d, int p, int k, float* q) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= total) return;
    const int64_t b = n / p;
    const int pos = (int)(n - b * p);
    int i = idx[n];
    i = i < 0 ? 0 : (i >= k ? k - 1 : i);
    for (int c = 0; c < d; ++c) q[((size_t)b * d + c) * p + pos] = __ldg(cb + (size_t)i * d + c);
}

template <int D>
int launch_vq(const VqArgs& a, cudaStream_t st) {
    const int64_t total = a.B * a.P;
    const int64_t blocks = (total + VQ_THREADS - 1) / VQ_THREADS;
    DMB_CHECK(blocks < (1ll << 31), "vq: too many positions");
    int kchunk = a.K;
    const int max_codes = (96 * 1024) / (D * 4);
    if (kchunk > max_codes) kchunk = max_codes;
    const size_t smem = (size_t)kchunk * D * 4;
    auto kern = vq_kernel<D>;
    if (smem > 40 * 1024) DMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DMB_LAUNCH((kern), (unsigned)blocks, VQ_THREADS, smem, st, a, kchunk);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace

int vq_forward(const VqArgs& a, cudaStream_t st) {
    DMB_CHECK(a.B > 0 && a.P > 0, "vq: empty input");
    DMB_CHECK(a.K >= 1 && a.K <= 1024, "vq: num_embeddings %d outside [1, 1024]", a.K);
    {
        const char* e = getenv("DMB_VQ_TC");           // read per call: tests compare both kernels
        if (!(e && e[0] == '0')) {
            const int r = vq_forward_tc(a, st);
            if (r <= 0) return r;
        }
    }
    switch (a.D) {
        case 8: return launch_vq<8>(a, st);
        case 16: return launch_vq<16>(a, st);
        case 32: return launch_vq<32>(a, st);
        case 64: return launch_vq<64>(a, st);
        case 128: return launch_vq<128>(a, st);
        default: break;
    }
    DMB_CHECK(false, "vq: embedding_dim %d not in {8,16,32,64,128}", a.D);
}

}  // namespace dmb

// ---------------------------------------------------------------------------------------
// C ABI (include/dynamorph_b200.h)
// ---------------------------------------------------------------------------------------
extern "C" {

int dmb_vq_reset(double* stats, int32_t k, void* stream) {
    DMB_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (2 + (size_t)k), (cudaStream_t)stream));
    return 0;
}

int dmb_vq_forward(const float* z, const float* codebook, int64_t batch, int32_t d,
                   int32_t positions_per_patch, int32_t k, float* z_st, int32_t* idx,
                   double* stats, void* stream) {
    DMB_CHECK(z && codebook, "dmb_vq_forward: null input");
    dmb::VqArgs a{};
    a.z = z; a.codebook = codebook; a.B = batch; a.D = d; a.P = positions_per_patch; a.K = k;
    a.z_st = z_st; a.idx = idx; a.stats = stats;
    return dmb::vq_forward(a, (cudaStream_t)stream);
}

int dmb_vq_finalize(const double* stats, int32_t d, int32_t k, float commitment_cost,
                    float* out2, void* stream) {
    DMB_CHECK(stats && out2, "dmb_vq_finalize: null pointer");
    DMB_LAUNCH((dmb::vq_finalize_kernel), 1, 256, 0, (cudaStream_t)stream, stats, d, k, commitment_cost, out2);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

int dmb_vq_gather(const int32_t* idx, const float* codebook, int64_t batch, int32_t d,
                  int32_t positions_per_patch, int32_t k, float* q, void* stream) {
    DMB_CHECK(idx && codebook && q, "dmb_vq_gather: null pointer");
    const int64_t total = batch * positions_per_patch;
    if (total == 0) return 0;
    DMB_LAUNCH((dmb::vq_gather_kernel), (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, idx, codebook, total, d, positions_per_patch, k, q);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// VectorQuantizer backward (autograd of vq_vae.py:65-76):
//   dL/dz      = g_zst + g_loss * beta * 2 (z - q) / N            (commitment term + straight-through)
//   dL/dE[k]   = g_loss * sum_{p: idx[p]=k} 2 (q[p] - z[p]) / N   (codebook term; N = B*D*P)
// ---------------------------------------------------------------------------------------
namespace dmb {
namespace {

constexpr int VQB_MAXD = 128;

// grad_z (+ optional BatchNorm-backward sums).  No atomics: the codebook gradient has its own kernel.
__global__ void __launch_bounds__(128) vq_backward_kernel(
        const float* __restrict__ z, const float* __restrict__ cb, const int32_t* __restrict__ idx,
        const float* __restrict__ g_zst, const float* __restrict__ g_extra, const float* __restrict__ g_loss,
        float g_loss_scale, float beta, int64_t total, int d, int p, float* __restrict__ grad_z,
        double* __restrict__ stats, const float* __restrict__ stat_src) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // stats (optional): per-CTA (sum g, sum g*stat_src) per channel, layout [block][d][2] -- the BatchNorm
    // backward sums of the residual layer that produced z.  Needs p % 128 == 0 (a CTA stays in one patch).
    __shared__ float red[VQB_MAXD][4][2];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < total;
    const float gl = (g_loss ? __ldg(g_loss) : 1.f) * g_loss_scale;
    const float coef = gl * 2.f / (float)((double)total * d);
    const int64_t b = live ? n / p : 0;
    const int pos = live ? (int)(n - b * p) : 0;
    const int k = live ? idx[n] : 0;
    const size_t base = (size_t)b * d * p + pos;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // eight channels at a time: all of their loads are issued before the first use (the one-channel loop was a chain of
    // DRAM round trips: 28 us for 4 MB at batch 256)
    for (int c0 = 0; c0 < d; c0 += 8) {
        float zv[8], q[8], g[8], sv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const bool on = live && c < d;
            const size_t o = base + (size_t)c * p;
            zv[j] = on ? __ldg(z + o) : 0.f;
            q[j] = on ? __ldg(cb + (size_t)k * d + c) : 0.f;
            g[j] = (on && g_zst) ? __ldg(g_zst + o) : 0.f;
            if (on && g_extra) g[j] += __ldg(g_extra + o);
            sv[j] = (on && stats) ? __ldg(stat_src + o) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            if (c >= d) break;
            const float gz = live ? g[j] + coef * beta * (zv[j] - q[j]) : 0.f;
            if (live && grad_z) grad_z[base + (size_t)c * p] = gz;
            if (stats) {
                float s = gz, sy = gz * sv[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                    sy += __shfl_xor_sync(0xffffffffu, sy, o);
                }
                if (lane == 0) { red[c][warp][0] = s; red[c][warp][1] = sy; }
            }
        }
    }
    if (stats) {
        __syncthreads();
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            double* dst = stats + ((size_t)blockIdx.x * d + c) * 2;
            dst[0] = (double)red[c][0][0] + (double)red[c][1][0] + (double)red[c][2][0] + (double)red[c][3][0];
            dst[1] = (double)red[c][0][1] + (double)red[c][1][1] + (double)red[c][2][1] + (double)red[c][3][1];
        }
    }
}

// Codebook gradient  dL/dE[k] = coef * sum_{p: idx[p]=k} (E[k] - z[p])   (embedding scatter-add).
// Persistent CTAs walk tiles of 256 positions: coef*(q - z) is formed with coalesced reads into shared
// memory, then added into a per-CTA [K][D] accumulator.  Deterministic, no atomics inside the CTA: thread (c, j) OWNS
// the accumulators (k, c) with k % pg == j, scans the tile's codes in ascending position order and adds the
// positions that are its own -- run-to-run bit-identical sums (an earlier shared-memory atomicAdd version made the
// codebook gradient differ in the last bit between runs, which Adam amplifies on zero-mean gradients).
//   partial != nullptr : each CTA writes its accumulator to partial[blockIdx.x]; vq_codebook_fold_kernel sums the
//                        rows in a fixed order
//   partial == nullptr : accumulators are pushed into grad_cb (pre-zeroed) with global atomics (stand-alone op)
constexpr int ZT_PITCH = 257;                // [d][257]: lanes over channels hit distinct banks
__global__ void __launch_bounds__(256) vq_codebook_scatter_kernel(
        const float* __restrict__ z, const float* __restrict__ cb, const int32_t* __restrict__ idx,
        const float* __restrict__ g_loss, float g_loss_scale, int64_t total, int d, int p, int k,
        float* __restrict__ partial, float* __restrict__ grad_cb) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    extern __shared__ float sm[];
    float* acc = sm;                         // [k][d]
    float* zt = sm + (size_t)k * d;          // [d][ZT_PITCH]
    int* it = reinterpret_cast<int*>(zt + (((size_t)d * ZT_PITCH + 3) & ~(size_t)3));   // [256], 16-byte aligned
    int* own = it + 256;                     // [256] owner class (code % pg) of each position, -1 = none
    const int tid = threadIdx.x;
    const float gl = (g_loss ? __ldg(g_loss) : 1.f) * g_loss_scale;
    const float coef = gl * 2.f / (float)((double)total * d);
    for (int i = tid; i < k * d; i += 256) acc[i] = 0.f;
    const int64_t ntiles = (total + 255) / 256;
    const int pg = 256 / d;                  // code classes (k % pg) handled in parallel (d <= 128)
    const int c = tid % d, j = tid / d;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        __syncthreads();
        const int64_t n = t * 256 + tid;
        if (n < total) {
            const int64_t b = n / p;
            const size_t base = (size_t)b * d * p + (n - b * p);
            const int kk = __ldg(idx + n);
            for (int cc = 0; cc < d; ++cc)
                zt[cc * ZT_PITCH + tid] = coef * (__ldg(cb + (size_t)kk * d + cc) - __ldg(z + base + (size_t)cc * p));
            it[tid] = kk;
            own[tid] = kk % pg;
        } else {
            it[tid] = -1;
            own[tid] = -1;
        }
        __syncthreads();
        if (j < pg) {
            for (int pos4 = 0; pos4 < 256; pos4 += 4) {
                const int4 o4 = *reinterpret_cast<const int4*>(own + pos4);
                const int os[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (os[u] == j) acc[it[pos4 + u] * d + c] += zt[c * ZT_PITCH + pos4 + u];
            }
        }
    }
    __syncthreads();
    if (partial) {
        float* dst = partial + (size_t)blockIdx.x * k * d;
        for (int i = tid; i < k * d; i += 256) dst[i] = acc[i];
    } else {
        for (int i = tid; i < k * d; i += 256)
            if (acc[i] != 0.f) atomicAdd(grad_cb + i, acc[i]);
    }
}

// grad_cb[e] = sum over the per-CTA partial rows, fixed order: block = 32 elements x 8 row groups
__global__ void __launch_bounds__(256) vq_codebook_fold_kernel(const float* __restrict__ partial, int nparts, int n,
                                                               float* __restrict__ grad_cb) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (e < n) {
        const int per = (nparts + 7) / 8;
        const int r0 = grp * per, r1 = min(nparts, r0 + per);
        for (int i = r0; i < r1; ++i) s += partial[(size_t)i * n + e];
    }
    red[grp][lane] = s;
    __syncthreads();
    if (grp != 0 || e >= n) return;
    s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += red[g][lane];
    grad_cb[e] = s;
}

int vq_codebook_grad(const float* z, const float* cb, const int32_t* idx, const float* g_loss, float g_loss_scale,
                     int64_t total, int d, int p, int k, float* grad_cb, float* scratch, int scratch_rows,
                     cudaStream_t st) {
    // scratch: scratch_rows x k x d floats, or nullptr -> global atomics straight into grad_cb
    const size_t smem = ((size_t)k * d + (size_t)d * ZT_PITCH + 512 + 4) * sizeof(float);
    DMB_CHECK(smem <= 220 * 1024, "vq codebook gradient: K=%d D=%d does not fit shared memory", k, d);
    if (smem > 48 * 1024)
        DMB_CUDA(cudaFuncSetAttribute(vq_codebook_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (total + 255) / 256;
    int grid = (int)(ntiles < 296 ? ntiles : 296);
    if (scratch && scratch_rows > 0) {
        if (grid > scratch_rows) grid = scratch_rows;
        DMB_LAUNCH((vq_codebook_scatter_kernel), grid, 256, smem, st, z, cb, idx, g_loss, g_loss_scale, total, d, p, k, scratch, nullptr);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
        DMB_LAUNCH((vq_codebook_fold_kernel), (k * d + 31) / 32, 256, 0, st, scratch, grid, k * d, grad_cb);
    } else {
        DMB_CUDA(cudaMemsetAsync(grad_cb, 0, sizeof(float) * (size_t)k * d, st));
        DMB_LAUNCH((vq_codebook_scatter_kernel), grid, 256, smem, st, z, cb, idx, g_loss, g_loss_scale, total, d, p, k, nullptr, grad_cb);
    }
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace
}  // namespace dmb

namespace dmb {
int vq_backward_stats(const float* z, const float* codebook, const int32_t* idx, const float* g_zst,
                      const float* g_extra, float g_loss_scale, float beta, int64_t batch, int d, int p, int k, float* grad_z,
                      float* grad_codebook, double* stats, const float* stat_src, float* scratch, int scratch_rows,
                      cudaStream_t st) {
    const int64_t total = batch * p;
    DMB_CHECK(!stats || p % 128 == 0, "vq backward: positions per patch (%d) must be a multiple of 128", p);
    DMB_CHECK(d <= VQB_MAXD, "vq backward: embedding_dim %d > %d", d, VQB_MAXD);
    DMB_LAUNCH((vq_backward_kernel), (unsigned)((total + 127) / 128), 128, 0, st, z, codebook, idx, g_zst, g_extra, nullptr, g_loss_scale, beta, total, d, p, grad_z, stats, stat_src);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    if (grad_codebook)
        DMB_TRY(vq_codebook_grad(z, codebook, idx, nullptr, g_loss_scale, total, d, p, k, grad_codebook, scratch,
                                 scratch_rows, st));
    return 0;
}
// the codebook gradient alone (the training step runs it beside the data-gradient chain, on the weight-gradient stream)
int vq_codebook_grad_only(const float* z, const float* codebook, const int32_t* idx, float g_loss_scale, int64_t batch,
                          int d, int p, int k, float* grad_codebook, float* scratch, int scratch_rows, cudaStream_t st) {
    DMB_CHECK(d <= VQB_MAXD, "vq backward: embedding_dim %d > %d", d, VQB_MAXD);
    return vq_codebook_grad(z, codebook, idx, nullptr, g_loss_scale, batch * p, d, p, k, grad_codebook, scratch,
                            scratch_rows, st);
}
}  // namespace dmb

extern "C" int dmb_vq_backward(const float* z, const float* codebook, const int32_t* idx,
                               const float* g_zst, const float* g_loss_dev, float g_loss_scale,
                               float commitment_cost, int64_t batch, int32_t d,
                               int32_t positions_per_patch, int32_t k, float* grad_z,
                               float* grad_codebook, void* stream) {
    DMB_CHECK(z && codebook && idx, "dmb_vq_backward: null pointer");
    const int64_t total = batch * positions_per_patch;
    DMB_CHECK(d <= dmb::VQB_MAXD, "dmb_vq_backward: embedding_dim %d > %d", d, dmb::VQB_MAXD);
    if (total == 0) {
        if (grad_codebook)
            DMB_CUDA(cudaMemsetAsync(grad_codebook, 0, sizeof(float) * (size_t)k * d, (cudaStream_t)stream));
        return 0;
    }
    if (grad_z) {
        DMB_LAUNCH((dmb::vq_backward_kernel), (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, z, codebook, idx, g_zst, nullptr, g_loss_dev, g_loss_scale, commitment_cost, total, d, positions_per_patch, grad_z, nullptr, nullptr);
        DMB_CUDA(cudaGetLastError());
        DMB_LAUNCHED(1);
    }
    if (grad_codebook)
        DMB_TRY(dmb::vq_codebook_grad(z, codebook, idx, g_loss_dev, g_loss_scale, total, d, positions_per_patch, k,
                                      grad_codebook, nullptr, 0, (cudaStream_t)stream));
    return 0;
}
