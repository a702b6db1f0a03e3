"""Layer-level check + timing of the tensor-memory-operand convolution (csrc/conv_tm.cu) against an fp64 convolution and
the CUDA-core TMA kernel, on the default configuration's thin encoder shapes.  Usage: python scripts/tm_layers.py [B]"""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamorph_b200._lib import call, ptr, load  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
load()
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
LAYERS = [("enc.4 4x4s2 8->16 @64", 8, 64, 16, 4, 2, 0, 1, False),
          ("enc.7 4x4s2 16->16 @32", 16, 32, 16, 4, 2, 0, 1, False),
          ("enc.10 3x3 16->16 @16", 16, 16, 16, 3, 1, 0, 0, False),
          ("res 3x3 16->32 @16 relu-in", 16, 16, 32, 3, 1, 1, 1, False),
          ("res 1x1 32->16 @16 +skip", 32, 16, 16, 1, 1, 0, 0, True)]


def time_it(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, cin, H, cout, ks, s, in_relu, out_relu, with_skip in LAYERS:
    g = torch.Generator(device=dev).manual_seed(cin * 100 + cout)
    x = torch.randn(B, cin, H, H, device=dev, generator=g)
    w = torch.randn(cout, cin, ks, ks, device=dev, generator=g) * 0.1
    b = torch.randn(cout, device=dev, generator=g) * 0.1
    skip = torch.randn(B, cout, H // s, H // s, device=dev, generator=g) if with_skip else None
    wp = w.permute(1, 2, 3, 0).contiguous().reshape(-1)           # [Cin][k][k][Cout]
    y_tm = torch.empty(B, cout, H // s, H // s, device=dev)
    y_cc = torch.empty_like(y_tm)
    n = C.c_int64()
    call("dmb_conv2d_tm_scratch_floats", cin, cout, ks, C.byref(n))
    scratch = torch.zeros(n.value, device=dev)
    f_tm = lambda: call("dmb_conv2d_tm", ptr(x), ptr(wp), ptr(b), ptr(y_tm), B, cin, H, H, cout, ks, s, in_relu, ptr(skip),
                        out_relu, ptr(scratch), st)
    f_cc = lambda: call("dmb_conv2d_forward", ptr(x), ptr(wp), ptr(b), ptr(y_cc), B, cin, H, H, cout, ks, s, None, None, 0,
                        in_relu, ptr(skip), out_relu, st)
    ms_tm, ms_cc = time_it(f_tm), time_it(f_cc)
    nb = min(B, 64)                                                # fp64 reference on a slice (first + last patches)
    sel = torch.cat([torch.arange(nb // 2), torch.arange(B - nb // 2, B)]).to(dev)
    xd = x[sel].double()
    if in_relu:
        xd = xd.relu()
    ref = F.conv2d(xd, w.double(), b.double(), stride=s, padding=0 if ks == 1 else 1)
    if with_skip:
        ref = ref + skip[sel].double()
    if out_relu:
        ref = ref.relu()
    scale = float(ref.abs().max())
    e_tm = float((y_tm[sel].double() - ref).abs().max()) / scale
    e_cc = float((y_cc[sel].double() - ref).abs().max()) / scale
    macs = (H // s) ** 2 * cout * cin * ks * ks
    print(f"{name:30s} B={B}: tmem {ms_tm:7.3f} ms ({2 * macs * B / ms_tm / 1e9:6.1f} TFLOP/s fp32-equiv) err {e_tm:.2e} | "
          f"cuda-core {ms_cc:7.3f} ms ({2 * macs * B / ms_cc / 1e9:6.1f}) err {e_cc:.2e} | speed-up {ms_cc / ms_tm:.2f}x",
          flush=True)

# ---- fused residual layer: y = x + conv1x1(relu(conv3x3(relu(x)) + b1)) + b2 in one kernel, against the two-kernel chain
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn(B, 16, 16, 16, device=dev, generator=g)
w1 = torch.randn(32, 16, 3, 3, device=dev, generator=g) * 0.1
b1 = torch.randn(32, device=dev, generator=g) * 0.1
w2 = torch.randn(16, 32, 1, 1, device=dev, generator=g) * 0.1
b2 = torch.randn(16, device=dev, generator=g) * 0.1
w1p, w2p = w1.permute(1, 2, 3, 0).contiguous().reshape(-1), w2.permute(1, 2, 3, 0).contiguous().reshape(-1)
n = C.c_int64()
call("dmb_residual_layer_tm_scratch_floats", C.byref(n))
scratch = torch.zeros(n.value, device=dev)
y_f = torch.empty(B, 16, 16, 16, device=dev)
f_fused = lambda: call("dmb_residual_layer_tm", ptr(x), ptr(w1p), ptr(b1), ptr(w2p), ptr(b2), ptr(y_f), B, ptr(scratch), st)
ms_f = time_it(f_fused)
sel = torch.cat([torch.arange(32), torch.arange(B - 32, B)]).to(dev)
xd = x[sel].double()
ref = xd + F.conv2d(F.conv2d(xd.relu(), w1.double(), b1.double(), padding=1).relu(), w2.double(), b2.double())
err = float((y_f[sel].double() - ref).abs().max() / ref.abs().max())
print(f"fused residual layer (3x3 16->32, ReLU, 1x1 32->16, +skip) B={B}: {ms_f:.3f} ms, err {err:.2e}", flush=True)
