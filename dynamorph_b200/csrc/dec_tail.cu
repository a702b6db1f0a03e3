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

// dec.4 (ConvTranspose2d CI -> CM, 4x4 stride 2 padding 1, + ReLU) fused in front of the forward tail: t2 -> t3 (stored:
// the backward pass needs it), decoded and the loss in ONE pass -- t3 (the largest tensor of the step, 262 KB per patch)
// is written once and never read back by the forward (vq_vae.py:296-298, :322).
//   out[co][2y+py][2x+px] = b[co] + sum_ci sum_{dy in {py-1, py}} sum_{dx in {px-1, px}} in[ci][y+dy][x+dx] * w[ci][ky][kx][co]
//   with ky = py + 1 - 2 dy, kx = px + 1 - 2 dx.
// A thread owns two adjacent input pixels and output row parity py: four consecutive output pixels for all CM channels
// (float4 stores, a warp = 64 consecutive x = 512 contiguous bytes per channel); py is warp-uniform, so the weight reads
// are 16-byte shared-memory broadcasts.
struct Tail2Args {
    const float* t2;       // (B, CI, Hi, Wi) post-ReLU input of dec.4
    const float* w4;       // dec.4 packed [CI][4][4][CM]
    const float* b4;       // [CM]
    float* t3;             // (B, CM, 2Hi, 2Wi) out (post-ReLU)
    const float* x; const float* mask; int mask_c; const float* cvar;
    const float* w6;       // dec.6 packed [CM][NI]
    const float* b6;       // [NI]
    float* decoded;
    double* loss_sum;
    int64_t items;         // B * Hi * 2 * Wi
    int hi, wi;            // powers of two: item -> (b, y, py, x) by shifts (64-bit div / mod by run-time values cost more
    int lh, lw;            // instructions than the 144 FMAs of an item)
};

template <int CI, int CM, int NI>
__global__ void __launch_bounds__(256, 3) dec_tail2_forward_kernel(const Tail2Args a) {
    static_assert(CM == 4, "one float4 of output channels per tap");
    // A thread owns TWO adjacent input pixels (x0, x0 + 1) and one output row parity: four consecutive output pixels
    // per channel = one float4 store; the 2 x 4 input window is one 8-byte and two 4-byte loads per row and channel, and
    // every weight read serves both pixels (the one-pixel form spent 518 instructions per item for 144 FMAs).
    pdl_wait();
    __shared__ float4 w4s[CI * 16];
    __shared__ double red[8];
    for (int i = threadIdx.x; i < CI * 16; i += 256) w4s[i] = __ldg(reinterpret_cast<const float4*>(a.w4) + i);
    float w6[CM][NI], b6[NI], cv[NI];
#pragma unroll
    for (int o = 0; o < NI; ++o) {
        b6[o] = __ldg(a.b6 + o); cv[o] = __ldg(a.cvar + o);
#pragma unroll
        for (int c = 0; c < CM; ++c) w6[c][o] = __ldg(a.w6 + c * NI + o);
    }
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b4));
    __syncthreads();
    const int wi = a.wi, hi = a.hi, wo = 2 * wi, wh = wi >> 1;
    const size_t in_plane = (size_t)hi * wi, out_plane = 4 * in_plane;
    const int64_t items = a.items >> 1;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < items; i += (int64_t)gridDim.x * 256) {
        const int x0 = 2 * (int)(i & (wh - 1));
        int64_t r = i >> (a.lw - 1);
        const int py = (int)(r & 1); r >>= 1;
        const int y = (int)(r & (hi - 1));
        const int64_t b = r >> a.lh;
        // rows y + py - 1, y + py; columns x0 - 1 .. x0 + 2 (zero outside the map)
        float v[CI][2][4];
        const float* ip = a.t2 + (size_t)b * CI * in_plane;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int yy = y + py - 1 + d;
            const bool rok = (unsigned)yy < (unsigned)hi;
            const bool lok = rok && x0 > 0, hok = rok && x0 + 2 < wi;
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) {
                const float* rp = ip + ci * in_plane + (size_t)(rok ? yy : 0) * wi + x0;
                const float2 m = rok ? __ldg(reinterpret_cast<const float2*>(rp)) : make_float2(0.f, 0.f);
                v[ci][d][0] = lok ? __ldg(rp - 1) : 0.f;
                v[ci][d][1] = m.x; v[ci][d][2] = m.y;
                v[ci][d][3] = hok ? __ldg(rp + 2) : 0.f;
            }
        }
        const size_t opix = (size_t)(2 * y + py) * wo + 2 * x0;
        float4 xv[NI], mk[NI];
#pragma unroll
        for (int o = 0; o < NI; ++o) {
            const size_t pi = ((size_t)b * NI + o) * out_plane + opix;
            xv[o] = __ldg(reinterpret_cast<const float4*>(a.x + pi));
            mk[o] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (a.mask) mk[o] = __ldg(reinterpret_cast<const float4*>(a.mask + ((a.mask_c == 1) ? (size_t)b * out_plane + opix : pi)));
        }
        // o[q]: output pixel 2 x0 + q (q = 2 j + px for input pixel x0 + j), four channels each
        float4 o[4] = {b4, b4, b4, b4};
#pragma unroll
        for (int ci = 0; ci < CI; ++ci)
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int ky = 3 - py - 2 * d;                  // dy = py - 1 + d  ->  ky = py + 1 - 2 dy
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    // px = 0: dx = -1 + e, kx = 3 - 2e;  px = 1: dx = e, kx = 2 - 2e
                    const float4 wa = w4s[(ci * 4 + ky) * 4 + (3 - 2 * e)];
                    const float4 wb = w4s[(ci * 4 + ky) * 4 + (2 - 2 * e)];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float va = v[ci][d][j + e], vb = v[ci][d][j + e + 1];
                        float4& oa = o[2 * j];
                        float4& ob = o[2 * j + 1];
                        oa.x = fmaf(va, wa.x, oa.x); oa.y = fmaf(va, wa.y, oa.y); oa.z = fmaf(va, wa.z, oa.z); oa.w = fmaf(va, wa.w, oa.w);
                        ob.x = fmaf(vb, wb.x, ob.x); ob.y = fmaf(vb, wb.y, ob.y); ob.z = fmaf(vb, wb.z, ob.z); ob.w = fmaf(vb, wb.w, ob.w);
                    }
                }
            }
        float t[CM][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            t[0][q] = fmaxf(o[q].x, 0.f); t[1][q] = fmaxf(o[q].y, 0.f); t[2][q] = fmaxf(o[q].z, 0.f); t[3][q] = fmaxf(o[q].w, 0.f);
        }
        float* tp = a.t3 + (size_t)b * CM * out_plane + opix;
#pragma unroll
        for (int c = 0; c < CM; ++c) *reinterpret_cast<float4*>(tp + c * out_plane) = make_float4(t[c][0], t[c][1], t[c][2], t[c][3]);
        float part = 0.f;
#pragma unroll
        for (int oc = 0; oc < NI; ++oc) {
            float dv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CM; ++c) s = fmaf(w6[c][oc], t[c][q], s);
                dv[q] = s + b6[oc];
            }
            const size_t pi = ((size_t)b * NI + oc) * out_plane + opix;
            *reinterpret_cast<float4*>(a.decoded + pi) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            const float e0 = dv[0] * mk[oc].x - xv[oc].x * mk[oc].x, e1 = dv[1] * mk[oc].y - xv[oc].y * mk[oc].y;
            const float e2 = dv[2] * mk[oc].z - xv[oc].z * mk[oc].z, e3 = dv[3] * mk[oc].w - xv[oc].w * mk[oc].w;
            part += ((e0 * e0) / cv[oc] + (e1 * e1) / cv[oc]) + ((e2 * e2) / cv[oc] + (e3 * e3) / cv[oc]);
        }
        acc += (double)part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int wq = 0; wq < 8; ++wq) s += red[wq];
        atomicAdd(a.loss_sum, s);
    }
}

// The same transposed convolution on its own (ConvTranspose2d CI -> 4, 4x4 stride 2 padding 1, bias, optional ReLU): the
// decoder layers with four output channels whose input has no pending transform (dec.2 of the default model: 8 -> 4 @32,
// and dec.4 when the fused tail is off).  Same ownership as dec_tail2_forward (two input pixels x one output row parity
// per thread, float4 stores); the input channels are walked four at a time.
struct ConvtSmallArgs {
    const float* x; const float* w; const float* bias; float* y;
    int64_t items;         // B * Hi * 2 * Wi
    int hi, wi, lh, lw, relu;
};

template <int CI>
__global__ void __launch_bounds__(256, 3) convt_small_kernel(const ConvtSmallArgs a) {
    static_assert(CI % 4 == 0, "input channels are walked four at a time");
    pdl_wait();
    __shared__ float4 ws[CI * 16];
    for (int i = threadIdx.x; i < CI * 16; i += 256) ws[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias));
    __syncthreads();
    const int wi = a.wi, hi = a.hi, wo = 2 * wi, wh = wi >> 1;
    const size_t in_plane = (size_t)hi * wi, out_plane = 4 * in_plane;
    const int64_t items = a.items >> 1;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < items; i += (int64_t)gridDim.x * 256) {
        const int x0 = 2 * (int)(i & (wh - 1));
        int64_t r = i >> (a.lw - 1);
        const int py = (int)(r & 1); r >>= 1;
        const int y = (int)(r & (hi - 1));
        const int64_t b = r >> a.lh;
        const float* ip = a.x + (size_t)b * CI * in_plane;
        float4 o[4] = {b4, b4, b4, b4};
#pragma unroll
        for (int cb = 0; cb < CI; cb += 4) {
            float v[4][2][4];
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int yy = y + py - 1 + d;
                const bool rok = (unsigned)yy < (unsigned)hi;
                const bool lok = rok && x0 > 0, hok = rok && x0 + 2 < wi;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const float* rp = ip + (cb + ci) * in_plane + (size_t)(rok ? yy : 0) * wi + x0;
                    const float2 m = rok ? __ldg(reinterpret_cast<const float2*>(rp)) : make_float2(0.f, 0.f);
                    v[ci][d][0] = lok ? __ldg(rp - 1) : 0.f;
                    v[ci][d][1] = m.x; v[ci][d][2] = m.y;
                    v[ci][d][3] = hok ? __ldg(rp + 2) : 0.f;
                }
            }
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const int ky = 3 - py - 2 * d;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float4 wa = ws[((cb + ci) * 4 + ky) * 4 + (3 - 2 * e)];
                        const float4 wb = ws[((cb + ci) * 4 + ky) * 4 + (2 - 2 * e)];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float va = v[ci][d][j + e], vb = v[ci][d][j + e + 1];
                            float4& oa = o[2 * j];
                            float4& ob = o[2 * j + 1];
                            oa.x = fmaf(va, wa.x, oa.x); oa.y = fmaf(va, wa.y, oa.y); oa.z = fmaf(va, wa.z, oa.z); oa.w = fmaf(va, wa.w, oa.w);
                            ob.x = fmaf(vb, wb.x, ob.x); ob.y = fmaf(vb, wb.y, ob.y); ob.z = fmaf(vb, wb.z, ob.z); ob.w = fmaf(vb, wb.w, ob.w);
                        }
                    }
                }
        }
        if (a.relu) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o[q].x = fmaxf(o[q].x, 0.f); o[q].y = fmaxf(o[q].y, 0.f); o[q].z = fmaxf(o[q].z, 0.f); o[q].w = fmaxf(o[q].w, 0.f);
            }
        }
        float* yp = a.y + (size_t)b * 4 * out_plane + (size_t)(2 * y + py) * wo + 2 * x0;
        *reinterpret_cast<float4*>(yp) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
        *reinterpret_cast<float4*>(yp + out_plane) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
        *reinterpret_cast<float4*>(yp + 2 * out_plane) = make_float4(o[0].z, o[1].z, o[2].z, o[3].z);
        *reinterpret_cast<float4*>(yp + 3 * out_plane) = make_float4(o[0].w, o[1].w, o[2].w, o[3].w);
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

bool dec_tail2_supported(int ci, int cm, int ni, int hi, int wi) {
    return ci == 4 && cm == 4 && ni == 2 && wi >= 32 && (wi & (wi - 1)) == 0 && hi > 0 && (hi & (hi - 1)) == 0;
}

int dec_tail2_forward(const DecTail2Args& d, cudaStream_t st) {
    DMB_CHECK(dec_tail2_supported(d.ci, d.cm, d.ni, d.hi, d.wi), "dec_tail2: unsupported shape %d -> %d -> %d @%dx%d", d.ci, d.cm,
              d.ni, d.hi, d.wi);
    DMB_CHECK(!(reinterpret_cast<uintptr_t>(d.w4) & 15) && !(reinterpret_cast<uintptr_t>(d.b4) & 15), "dec_tail2: weights must be 16-byte aligned");
    Tail2Args a{};
    a.t2 = d.t2; a.w4 = d.w4; a.b4 = d.b4; a.t3 = d.t3; a.x = d.x; a.mask = d.mask; a.mask_c = d.mask_c; a.cvar = d.cvar;
    a.w6 = d.w6; a.b6 = d.b6; a.decoded = d.decoded; a.loss_sum = d.loss_sum;
    a.items = d.B * (int64_t)d.hi * 2 * d.wi; a.hi = d.hi; a.wi = d.wi;
    for (a.lh = 0; (1 << a.lh) < d.hi; ++a.lh) {}
    for (a.lw = 0; (1 << a.lw) < d.wi; ++a.lw) {}
    int64_t blocks = (a.items / 2 + 255) / 256;
    if (blocks > 148 * 3) blocks = 148 * 3;       // one resident wave (3 CTAs per SM), grid-stride
    DMB_LAUNCH((dec_tail2_forward_kernel<4, 4, 2>), (unsigned)blocks, 256, 0, st, a);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}

bool convt_small_supported(int ci, int co, int hi, int wi) {
    return (ci == 4 || ci == 8) && co == 4 && wi >= 32 && (wi & (wi - 1)) == 0 && hi > 0 && (hi & (hi - 1)) == 0;
}

int convt_small(const float* x, const float* w_packed, const float* bias, float* y, int64_t B, int ci, int co, int hi, int wi,
                int relu, cudaStream_t st) {
    DMB_CHECK(convt_small_supported(ci, co, hi, wi), "convt_small: unsupported shape %d -> %d @%dx%d", ci, co, hi, wi);
    DMB_CHECK(!(reinterpret_cast<uintptr_t>(w_packed) & 15) && !(reinterpret_cast<uintptr_t>(bias) & 15) &&
              !(reinterpret_cast<uintptr_t>(x) & 7) && !(reinterpret_cast<uintptr_t>(y) & 15), "convt_small: alignment");
    ConvtSmallArgs a{};
    a.x = x; a.w = w_packed; a.bias = bias; a.y = y; a.items = B * (int64_t)hi * 2 * wi; a.hi = hi; a.wi = wi; a.relu = relu;
    for (a.lh = 0; (1 << a.lh) < hi; ++a.lh) {}
    for (a.lw = 0; (1 << a.lw) < wi; ++a.lw) {}
    int64_t blocks = (a.items / 2 + 255) / 256;
    if (blocks > 148 * 3) blocks = 148 * 3;
    if (blocks < 1) blocks = 1;
    if (ci == 4) DMB_LAUNCH((convt_small_kernel<4>), (unsigned)blocks, 256, 0, st, a);
    else DMB_LAUNCH((convt_small_kernel<8>), (unsigned)blocks, 256, 0, st, a);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
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
