// Training step behind the C ABI: forward keeping activations, backward, (Adam is in optim.cu).
#include "model.h"

extern "C" {

int dmb_train_forward(const dmb_model* m, const float* packed, const float* params, const float* x,
                      const float* mask, int32_t mask_channels, const float* channel_var, int64_t batch,
                      float* decoded, float* losses_out, float* bnbuf_inout, void* workspace,
                      size_t workspace_bytes, void* stream) {
    DMB_CHECK(false, "dmb_train_forward: not built yet");
}

int dmb_train_backward(const dmb_model* m, const float* packed, const float* params, const float* x,
                       const float* mask, int32_t mask_channels, const float* channel_var, int64_t batch,
                       float grad_scale, float* grads, void* workspace, size_t workspace_bytes, void* stream) {
    DMB_CHECK(false, "dmb_train_backward: not built yet");
}

}  // extern "C"
