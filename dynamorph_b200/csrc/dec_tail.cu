// The tail of the z16 decoder in the training step, fused into two streaming kernels
// (reference: HiddenStateExtractor/vq_vae.py:298 `nn.Conv2d(num_hiddens//4, num_inputs, 1)`, :322 the masked,
// channel_var-weighted MSE, and their autograd in run_training.py:406 `total_loss.backward()`).
//
// The 1x1 convolution CM -> NI at full resolution, the reconstruction loss and their backward are pure HBM/L2
// streaming work (a handful of FMAs per element).  As separate launches the step read and wrote the full-resolution
// tensors seven times (conv1x1, recon_loss | recon_grad, weight gradient, data gradient, bias-partial fold):
//   forward :  t3 (CM planes, post-ReLU) , x  ->  decoded , sum of the weighted squared error
//   backward:  decoded , x , t3  ->  g_t3 = relu'(t3) * W^T gd      (gd = d loss / d decoded, never stored)
//                                   dW[o][c] = sum gd[o] * t3[c] , db[o] = sum gd[o] , db_prev[c] = sum g_t3[c]
// One thread owns four consecutive pixels of every channel (float4 per plane, coalesced across the warp).  Sums: float
// per thread over its few quads, double across the warp / CTA, one row of partials per CTA, folded in fixed order by
// the last CTA to finish (ticket counter; it resets the counter for the next step), so the result is deterministic.
#include "common.cuh"

namespace dmb {
namespace {

struct TailArgs {
    const float* t3; const float* x; const float* mask; int mask_c; const float* cvar;
    const float* w;        // [CM][NI]  (the packed [Cin][1][1][Cout] layout)
    const float* bias;     // [NI]
    float* decoded;        // forward: out; backward: in
    double* loss_sum;      // forward: += sum of the weighted squared error
    float* g_t3;           // backward: out
    float scale;           // backward: grad_scale * weight_recon / numel(decoded)
    double* partials;      // backward: [grid][NI*CM + NI + CM]
    unsigned* ticket;      // backward: zero on entry, zero on exit
    float* dw; float* db; float* db_prev;
    int64_t quads;         // B * HW / 4
    int hw4;
};

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }

template <int CM, int NI>
__global__ void __launch_bounds__(256) dec_tail_forward_kernel(const TailArgs a) {
    pdl_wait();
    __shared__ double red[8];
    float w[CM][NI], bs[NI], cv[NI];
#pragma unroll
    for (int o = 0; o < NI; ++o) {
        bs[o] = __ldg(a.bias + o); cv[o] = __ldg(a.cvar + o);
#pragma unroll
        for (int c = 0; c < CM; ++c) w[c][o] = __ldg(a.w + c * NI + o);
    }
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.quads; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / a.hw4;
        const int64_t q = i - b * a.hw4;
        float4 d[NI];
#pragma unroll
        for (int o = 0; o < NI; ++o) d[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            const float4 t = ld4(a.t3, (b * CM + c) * a.hw4 + q);
#pragma unroll
            for (int o = 0; o < NI; ++o) {
                d[o].x = fmaf(w[c][o], t.x, d[o].x); d[o].y = fmaf(w[c][o], t.y, d[o].y);
                d[o].z = fmaf(w[c][o], t.z, d[o].z); d[o].w = fmaf(w[c][o], t.w, d[o].w);
            }
        }
        float part = 0.f;
#pragma unroll
        for (int o = 0; o < NI; ++o) {
            d[o].x += bs[o]; d[o].y += bs[o]; d[o].z += bs[o]; d[o].w += bs[o];
            const int64_t pi = (b * NI + o) * a.hw4 + q;
            reinterpret_cast<float4*>(a.decoded)[pi] = d[o];
            const float4 v = ld4(a.x, pi);
            float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
            if (a.mask) mk = ld4(a.mask, (a.mask_c == 1) ? (b * a.hw4 + q) : pi);
            // mse_loss(decoded*mask, inputs*mask, 'none') / channel_var  (vq_vae.py:322)
            const float e0 = d[o].x * mk.x - v.x * mk.x, e1 = d[o].y * mk.y - v.y * mk.y;
            const float e2 = d[o].z * mk.z - v.z * mk.z, e3 = d[o].w * mk.w - v.w * mk.w;
            part += ((e0 * e0) / cv[o] + (e1 * e1) / cv[o]) + ((e2 * e2) / cv[o] + (e3 * e3) / cv[o]);
        }
        acc += (double)part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int wi = 0; wi < 8; ++wi) s += red[wi];
        atomicAdd(a.loss_sum, s);
    }
}

template <int CM, int NI>
__global__ void __launch_bounds__(256, (CM <= 4) ? 4 : 1) dec_tail_backward_kernel(const TailArgs a) {
    pdl_wait();
    constexpr int NV = NI * CM + NI + CM;
    __shared__ double red[8][NV];
    __shared__ bool last;
    float w[CM][NI], k2[NI];
#pragma unroll
    for (int o = 0; o < NI; ++o) {
        k2[o] = 2.f * a.scale / __ldg(a.cvar + o);
#pragma unroll
        for (int c = 0; c < CM; ++c) w[c][o] = __ldg(a.w + c * NI + o);
    }
    // per-thread partial sums in float (a thread adds a few dozen products; the cross-thread tree below is double)
    float sums[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) sums[j] = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.quads; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / a.hw4;
        const int64_t q = i - b * a.hw4;
        // every load of the iteration is issued before the first use (one DRAM round trip per quad, not two)
        float4 d[NI], v[NI], mk[NI], t[CM];
#pragma unroll
        for (int o = 0; o < NI; ++o) {
            const int64_t pi = (b * NI + o) * a.hw4 + q;
            d[o] = ld4(a.decoded, pi); v[o] = ld4(a.x, pi);
            mk[o] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (a.mask) mk[o] = ld4(a.mask, (a.mask_c == 1) ? (b * a.hw4 + q) : pi);
        }
#pragma unroll
        for (int c = 0; c < CM; ++c) t[c] = ld4(a.t3, (b * CM + c) * a.hw4 + q);      // post-ReLU: relu'(.) = (t > 0)
        float4 gd[NI];
#pragma unroll
        for (int o = 0; o < NI; ++o) {
            // d/d dec of  scale * sum ((dec*m - x*m)^2 / cv)  =  scale * 2 (dec*m - x*m) m / cv
            gd[o].x = k2[o] * (d[o].x * mk[o].x - v[o].x * mk[o].x) * mk[o].x;
            gd[o].y = k2[o] * (d[o].y * mk[o].y - v[o].y * mk[o].y) * mk[o].y;
            gd[o].z = k2[o] * (d[o].z * mk[o].z - v[o].z * mk[o].z) * mk[o].z;
            gd[o].w = k2[o] * (d[o].w * mk[o].w - v[o].w * mk[o].w) * mk[o].w;
            sums[NI * CM + o] += (gd[o].x + gd[o].y) + (gd[o].z + gd[o].w);
        }
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            const int64_t pi = (b * CM + c) * a.hw4 + q;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < NI; ++o) {
                g.x = fmaf(w[c][o], gd[o].x, g.x); g.y = fmaf(w[c][o], gd[o].y, g.y);
                g.z = fmaf(w[c][o], gd[o].z, g.z); g.w = fmaf(w[c][o], gd[o].w, g.w);
                sums[o * CM + c] += fmaf(gd[o].x, t[c].x, gd[o].y * t[c].y) + fmaf(gd[o].z, t[c].z, gd[o].w * t[c].w);
            }
            if (!(t[c].x > 0.f)) g.x = 0.f;
            if (!(t[c].y > 0.f)) g.y = 0.f;
            if (!(t[c].z > 0.f)) g.z = 0.f;
            if (!(t[c].w > 0.f)) g.w = 0.f;
            reinterpret_cast<float4*>(a.g_t3)[pi] = g;
            sums[NI * CM + NI + c] += (g.x + g.y) + (g.z + g.w);
        }
    }
    // warp tree (fixed order) -> eight warp rows -> one row per CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double s = (double)sums[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int wi = 0; wi < 8; ++wi) s += red[wi][threadIdx.x];
        a.partials[(size_t)blockIdx.x * NV + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA folds the rows in index order: 256 threads = NV columns x (256 / NV) row groups, then the groups
    constexpr int GROUPS = 256 / NV;
    __shared__ double fold[GROUPS][NV];
    const int col = threadIdx.x % NV, grp = threadIdx.x / NV;
    if (grp < GROUPS) {
        const int per = ((int)gridDim.x + GROUPS - 1) / GROUPS;
        const int r0 = grp * per, r1 = min((int)gridDim.x, r0 + per);
        double s = 0.0;
        for (int r = r0; r < r1; r += 8) {           // eight independent loads in flight, added in row order
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (r + j < r1) ? __ldcg(a.partials + (size_t)(r + j) * NV + col) : 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j];
        }
        fold[grp][col] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int g = 0; g < GROUPS; ++g) s += fold[g][threadIdx.x];
        const int j = threadIdx.x;
        if (j < NI * CM) a.dw[j] = (float)s;                              // torch layout (NI, CM, 1, 1)
        else if (j < NI * CM + NI) a.db[j - NI * CM] = (float)s;
        else a.db_prev[j - NI * CM - NI] = (float)s;
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
}

// CTAs of a launch: the backward kernel writes one partial row per CTA, which ONE CTA folds at the end
int64_t tail_blocks(int64_t quads, bool backward) {
    int64_t blocks = (quads + 255) / 256;
    const int64_t cap = backward ? 148 * 4 : 148 * 8;
    return blocks > cap ? cap : blocks;
}

template <int CM, int NI>
int launch_tail(const TailArgs& a, bool backward, cudaStream_t st) {
    const int64_t blocks = tail_blocks(a.quads, backward);
    if (backward) DMB_LAUNCH((dec_tail_backward_kernel<CM, NI>), (unsigned)blocks, 256, 0, st, a);
    else DMB_LAUNCH((dec_tail_forward_kernel<CM, NI>), (unsigned)blocks, 256, 0, st, a);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

}  // namespace

bool dec_tail_supported(int cm, int ni, int hw) { return (cm == 4 || cm == 16) && ni == 2 && hw % 4 == 0; }

int64_t dec_tail_partial_doubles(int64_t B, int hw, int cm, int ni) {
    return tail_blocks(B * (int64_t)(hw / 4), true) * (ni * cm + ni + cm);
}

int dec_tail_forward(const DecTailArgs& d, cudaStream_t st) {
    DMB_CHECK(dec_tail_supported(d.cm, d.ni, d.hw), "dec_tail: unsupported shape %d -> %d", d.cm, d.ni);
    TailArgs a{};
    a.t3 = d.t3; a.x = d.x; a.mask = d.mask; a.mask_c = d.mask_c; a.cvar = d.cvar; a.w = d.w; a.bias = d.bias;
    a.decoded = d.decoded; a.loss_sum = d.loss_sum; a.quads = d.B * (int64_t)(d.hw / 4); a.hw4 = d.hw / 4;
    return d.cm == 4 ? launch_tail<4, 2>(a, false, st) : launch_tail<16, 2>(a, false, st);
}

int dec_tail_backward(const DecTailArgs& d, cudaStream_t st) {
    DMB_CHECK(dec_tail_supported(d.cm, d.ni, d.hw), "dec_tail: unsupported shape %d -> %d", d.cm, d.ni);
    TailArgs a{};
    a.t3 = d.t3; a.x = d.x; a.mask = d.mask; a.mask_c = d.mask_c; a.cvar = d.cvar; a.w = d.w;
    a.decoded = d.decoded; a.g_t3 = d.g_t3; a.scale = d.scale; a.partials = d.partials; a.ticket = d.ticket;
    a.dw = d.dw; a.db = d.db; a.db_prev = d.db_prev; a.quads = d.B * (int64_t)(d.hw / 4); a.hw4 = d.hw / 4;
    return d.cm == 4 ? launch_tail<4, 2>(a, true, st) : launch_tail<16, 2>(a, true, st);
}

}  // namespace dmb
