// Entry points of the TMA-fed convolution: tries the per-family dispatch units (conv_tma_default.cu, conv_tma_heavy.cu;
// the kernel itself is conv_tma.cuh).  Split so that the two families compile in parallel.
#include "common.cuh"

#include <stdlib.h>

namespace dmb {

int conv_tma_default(const ConvFwdArgs& a, cudaStream_t st);
int conv_tma_heavy(const ConvFwdArgs& a, cudaStream_t st);
int conv_tma_bands_default(int ks, int stride, int Cin, int Cout, int H, int W, int64_t B);
int conv_tma_bands_heavy(int ks, int stride, int Cin, int Cout, int H, int W, int64_t B);

static bool tma_disabled() {          // A/B switch: DMB_CONV_TMA=0 -> generic kernel only
    const char* e = getenv("DMB_CONV_TMA");
    return e && e[0] == '0';
}

// Band count (BatchNorm partial rows per sample) of the TMA kernel for this geometry and batch size (B == 0: any), or 0
// if it has no instantiation.
int conv_tma_bands(int ks, int stride, int Cin, int Cout, int H, int W, int64_t B) {
    if (tma_disabled()) return 0;
    const int nb = conv_tma_bands_default(ks, stride, Cin, Cout, H, W, B);
    return nb ? nb : conv_tma_bands_heavy(ks, stride, Cin, Cout, H, W, B);
}

// Returns 1 if the call was not taken (caller falls back to conv_fwd's generic kernel), 0 on success, <0 on error.
int conv_tma(const ConvFwdArgs& a, cudaStream_t st) {
    if (tma_disabled()) return 1;
    const int r = conv_tma_default(a, st);
    return r == 1 ? conv_tma_heavy(a, st) : r;
}

}  // namespace dmb
