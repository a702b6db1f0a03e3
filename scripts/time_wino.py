"""Winograd-on-tensor-cores 3x3 (csrc/conv_wino_tc.cu) vs the direct CUDA-core TMA kernel, default-config shapes."""
import ctypes as C, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamorph_b200._lib import call, ptr

def timed(fn, n=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for cout, in_relu, out_relu in ((32, 1, 1), (16, 0, 0)):
    g = torch.Generator(device="cuda").manual_seed(cout)
    x = torch.randn(B, 16, 16, 16, device="cuda", generator=g)
    w = torch.randn(cout, 16, 3, 3, device="cuda", generator=g) / 12.0
    bias = torch.randn(cout, device="cuda", generator=g)
    wp = w.permute(1, 2, 3, 0).contiguous()
    scratch = torch.empty(2 * 16 * 16 * cout, device="cuda")
    y1 = torch.empty(B, cout, 16, 16, device="cuda"); y2 = torch.empty_like(y1)
    f1 = lambda: call("dmb_conv2d_wino", ptr(x), ptr(wp), ptr(bias), ptr(y1), B, 16, 16, 16, cout, in_relu, out_relu, None, None, None, ptr(scratch), st)
    f2 = lambda: call("dmb_conv2d_forward", ptr(x), ptr(wp), ptr(bias), ptr(y2), B, 16, 16, 16, cout, 3, 1, None, None, 0, in_relu, None, out_relu, st)
    t1, t2 = timed(f1), timed(f2)
    nb = 16
    xr = x[:nb].double().cpu(); xr = xr.relu() if in_relu else xr
    ref = F.conv2d(xr, w.double().cpu(), bias.double().cpu(), padding=1); ref = ref.relu() if out_relu else ref
    e1 = float((y1[:nb].double().cpu() - ref).abs().max() / ref.abs().max())
    e2 = float((y2[:nb].double().cpu() - ref).abs().max() / ref.abs().max())
    print(f"3x3 16->{cout} B={B}: winograd-tc {t1:.3f} ms err {e1:.2e} | direct {t2:.3f} ms err {e2:.2e} | x{t2 / t1:.2f}")
