"""Profile target: one thin encoder layer through dmb_conv2d_tm (csrc/conv_tm.cu).
    python scripts/prof_tm.py e2 8192        # e2 = enc.4, e3 = enc.7, e4 = enc.10, ra = res 3x3, rb = res 1x1"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200._lib import call, ptr

LAYERS = {"e2": (8, 64, 16, 4, 2), "e3": (16, 32, 16, 4, 2), "e4": (16, 16, 16, 3, 1), "ra": (16, 16, 32, 3, 1),
          "rb": (32, 16, 16, 1, 1)}
which = sys.argv[1] if len(sys.argv) > 1 else "e2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cin, H, cout, ks, s = LAYERS[which]
dev = torch.device("cuda:0")
x = torch.randn(B, cin, H, H, device=dev)
w = torch.randn(cin * ks * ks * cout, device=dev) * 0.05
b = torch.zeros(cout, device=dev)
y = torch.empty(B, cout, H // s, H // s, device=dev)
n = C.c_int64()
call("dmb_conv2d_tm_scratch_floats", cin, cout, ks, C.byref(n))
scratch = torch.zeros(n.value, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fn = lambda: call("dmb_conv2d_tm", ptr(x), ptr(w), ptr(b), ptr(y), B, cin, H, H, cout, ks, s, 0, None, 1, ptr(scratch), st)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
macs = (H // s) ** 2 * cout * cin * ks * ks
print(f"{which} B={B}: {ms:.3f} ms  {2*macs*B/ms/1e9:.1f} TFLOP/s fp32-equivalent")
