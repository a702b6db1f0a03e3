"""Profile target: single conv layer launches through the layer-level C ABI.
    python scripts/prof_layer.py e2 2048"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200._lib import call, ptr

LAYERS = {"e1": (2, 128, 8, 4, 2), "e2": (8, 64, 16, 4, 2), "e3": (16, 32, 16, 4, 2), "e4": (16, 16, 16, 3, 1),
          "ra": (16, 16, 32, 3, 1), "rb": (32, 16, 16, 1, 1)}
which = sys.argv[1] if len(sys.argv) > 1 else "e2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cin, H, cout, ks, s = LAYERS[which]
dev = torch.device("cuda:0")
x = torch.randn(B, cin, H, H, device=dev)
w = torch.randn(cin * ks * ks * cout, device=dev) * 0.05
b = torch.zeros(9 * cout, device=dev)
y = torch.empty(B, cout, H // s, H // s, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
fn = lambda: call("dmb_conv2d_forward", ptr(x), ptr(w), ptr(b), ptr(y), B, cin, H, H, cout, ks, s, None, None, 0, 0, None, 1, st)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
macs = (H // s) ** 2 * cout * cin * ks * ks
print(f"{which} B={B}: {ms:.3f} ms  {2*macs*B/ms/1e9:.1f} TFLOP/s")
torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
