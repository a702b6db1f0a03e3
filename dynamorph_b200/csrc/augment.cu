// Training augmentation of run_training.run_one_batch (/root/reference/run_training.py:396-403) as ONE launch:
// per sample, flip over {none, H, W} then rot90 by k quarter turns over the (H, W) plane.  The reference does this in a
// Python loop (two tiny kernels per sample); the random draws stay on the host, in the reference's np.random order,
// and arrive here as one byte per sample (flip | rot << 2).  Pure permutation: bit-exact.
#include "common.cuh"

namespace dmb {
namespace {

__global__ void __launch_bounds__(256) augment_kernel(const float* __restrict__ x, const uint8_t* __restrict__ ops,
                                                      int64_t B, int C, int H, float* __restrict__ out) {
    pdl_wait();      // programmatic dependent launch: everything below may read the previous kernel's output
    // square planes (H == W): rot90 keeps the shape.  One thread per output element, rows of the OUTPUT are contiguous.
    const int64_t total = B * C * H * H;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int j = (int)(e % H);
        const int i = (int)((e / H) % H);
        const int64_t plane = e / ((int64_t)H * H);          // b * C + c
        const int64_t b = plane / C;
        const int op = ops[b];
        const int flip = op & 3, rot = (op >> 2) & 3;
        // torch.rot90(T, k, [1, 2]):  k=1: out[i][j] = T[j][W-1-i];  k=2: T[H-1-i][W-1-j];  k=3: T[H-1-j][i]
        int ti, tj;
        if (rot == 0) { ti = i; tj = j; }
        else if (rot == 1) { ti = j; tj = H - 1 - i; }
        else if (rot == 2) { ti = H - 1 - i; tj = H - 1 - j; }
        else { ti = H - 1 - j; tj = i; }
        // T = torch.flip(img, dims=(flip,)): 1 flips rows, 2 flips columns
        const int si = (flip == 1) ? H - 1 - ti : ti;
        const int sj = (flip == 2) ? H - 1 - tj : tj;
        out[e] = __ldg(x + (plane * H + si) * H + sj);
    }
}

}  // namespace
}  // namespace dmb

extern "C" int dmb_augment_batch(const float* x, const uint8_t* ops_dev, int64_t batch, int32_t channels,
                                 int32_t height, int32_t width, float* out, void* stream) {
    DMB_CHECK(x && ops_dev && out, "dmb_augment_batch: null pointer");
    DMB_CHECK(x != out, "dmb_augment_batch: in-place augmentation is not supported");
    DMB_CHECK(height == width, "dmb_augment_batch: rot90 needs square patches (%d x %d)", height, width);
    if (batch == 0) return 0;
    const int64_t total = batch * channels * height * width;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    DMB_LAUNCH((dmb::augment_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, x, ops_dev, batch, channels, height, out);
    DMB_CUDA(cudaGetLastError());
    DMB_LAUNCHED(1);
    return 0;
}
