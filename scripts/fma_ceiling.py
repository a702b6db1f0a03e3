import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200._lib import call, ptr
dev = torch.device("cuda:0")
scratch = torch.rand(4096, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
fl = C.c_double()
ms = t(lambda: call("dmb_bench_fp32_fma", 148 * 8, 256, 20000, ptr(scratch), C.byref(fl), st))
print(f"register-chain FMA peak: {fl.value/ms/1e9:.1f} TFLOP/s")
for order in (0, 1):
    for blocks in (148 * 4, 148 * 8):
        ms = t(lambda: call("dmb_bench_fma_tile", order, blocks, 4000, ptr(scratch), C.byref(fl), st))
        print(f"8x8 tile order={order} blocks={blocks}: {fl.value/ms/1e9:.1f} TFLOP/s")
for order in (0, 1):
    ms = t(lambda: call("dmb_bench_fma2_tile", order, 148 * 4, 4000, ptr(scratch), C.byref(fl), st))
    print(f"8x8 tile FFMA2 order={order}: {fl.value/ms/1e9:.1f} TFLOP/s")
for variant, name in ((0, "weights in shared memory"), (1, "weights in __constant__"), (2, "weights in kernel params")):
    ms = t(lambda: call("dmb_bench_fma_conv", variant, 148 * 4, 40, ptr(scratch), C.byref(fl), st))
    print(f"conv core, {name}: {fl.value/ms/1e9:.1f} TFLOP/s")
