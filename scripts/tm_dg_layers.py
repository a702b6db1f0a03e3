"""Time the data-gradient / transposed forms of the tensor-memory kernels through the layer-level C ABI at a training
batch size (each call also re-packs the weight image: one small launch, the same for every variant).
    python scripts/tm_dg_layers.py [batch]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200._lib import call, ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(0)
R = lambda *s: torch.randn(*s, device="cuda", generator=g)
flush = torch.empty(64 << 20, device="cuda")


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rows_max = C.c_int32(); call("dmb_conv2d_tm_batch_stat_rows", C.byref(rows_max))
for (ks, s, cin, cout, W) in [(1, 1, 16, 32, 16), (3, 1, 32, 16, 16), (3, 1, 16, 16, 16), (4, 2, 8, 16, 64), (4, 2, 8, 16, 32)]:
    gy, wp = R(B, cin, W, W), R(cin, ks, ks, cout)
    n = C.c_int64(); call("dmb_conv2d_tm_scratch_floats", cin, cout, ks, C.byref(n))
    scr = torch.zeros(((n.value + 63) // 64) * 64 + cout, device="cuda")
    Ho = W // s
    gx, m, src, skip = torch.empty(B, cout, Ho, Ho, device="cuda"), R(B, cout, Ho, Ho), R(B, cout, Ho, Ho), R(B, cout, Ho, Ho)
    ms, mt = torch.rand(cout, device="cuda") + .5, R(cout)
    yr, ga, gb, gc = R(B, cin, W, W), R(cin), R(cin), R(cin)
    stats = torch.zeros(rows_max.value * cout * 2, dtype=torch.float64, device="cuda"); rows = C.c_int32()
    dual = s == 1 and not (ks == 3 and cin == 32)
    def run(d, gate, stat, same):
        call("dmb_conv2d_tm_dgrad", ptr(gy), ptr(wp), ptr(gx), B, cin, W, W, cout, ks, s, ptr(yr) if d else None,
             ptr(ga) if d else None, ptr(gb) if d else None, ptr(gc) if d else None, ptr(m) if gate else None,
             ptr(ms) if gate else None, ptr(mt) if gate else None, None, ptr(stats) if stat else None,
             (ptr(m) if same else ptr(src)) if stat else None, C.byref(rows), ptr(scr), st())
    line = f"dg {ks}x{ks} s{s} {cin}->{cout} @{W}: plain {timeit(lambda: run(0, 0, 0, 0)):6.1f}  gate {timeit(lambda: run(0, 1, 0, 0)):6.1f}  gate+sums(same src) {timeit(lambda: run(0, 1, 1, 1)):6.1f}  gate+sums(other src) {timeit(lambda: run(0, 1, 1, 0)):6.1f}"
    if dual:
        line += f"  dual {timeit(lambda: run(1, 0, 0, 0)):6.1f}  dual+gate+sums {timeit(lambda: run(1, 1, 1, 1)):6.1f}"
    print(line + " us")
for (cin, cout, W, d) in [(16, 8, 16, 0), (16, 16, 16, 1), (16, 8, 32, 1)]:
    x, wp, bias = R(B, cin, W, W), R(cin, 4, 4, cout), R(cout)
    n = C.c_int64(); call("dmb_conv2d_tm_scratch_floats", cin, 4 * cout, 3, C.byref(n))
    scr = torch.zeros(((n.value + 63) // 64) * 64 + cout, device="cuda")
    y, m = torch.empty(B, cout, 2 * W, 2 * W, device="cuda"), R(B, cout, 2 * W, 2 * W)
    ms, mt = torch.rand(cout, device="cuda") + .5, R(cout)
    yr, ga, gb, gc = R(B, cin, W, W), R(cin), R(cin), R(cin)
    stats = torch.zeros(2 * rows_max.value * cout * 2, dtype=torch.float64, device="cuda"); rows = C.c_int32()
    def run(dual, gate, stat):
        call("dmb_conv_transpose2d_tm", ptr(x), ptr(wp), None if d else ptr(bias), ptr(y), B, cin, W, W, cout, 0 if d else 1, d,
             ptr(yr) if dual else None, ptr(ga) if dual else None, ptr(gb) if dual else None, ptr(gc) if dual else None,
             ptr(m) if gate else None, ptr(ms) if gate else None, ptr(mt) if gate else None, ptr(stats) if stat else None,
             ptr(m) if stat else None, C.byref(rows), ptr(scr), st())
    line = f"ct {cin}->{cout} @{W} {'dg' if d else 'fwd'}: plain {timeit(lambda: run(0, 0, 0)):6.1f}"
    if d:
        line += f"  gate {timeit(lambda: run(0, 1, 0)):6.1f}  gate+sums {timeit(lambda: run(0, 1, 1)):6.1f}  dual {timeit(lambda: run(1, 0, 0)):6.1f}  all {timeit(lambda: run(1, 1, 1)):6.1f}"
    print(line + " us")
