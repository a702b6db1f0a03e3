"""Accuracy and speed of the tensor-core conv (csrc/conv_tc.cu) against the CUDA-core TMA conv, per wide layer.
usage: python scripts/tc_layers.py [batch]     (env DMB_TC_NACC / DMB_TC_STAGES select kernel variants)"""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamorph_b200._lib import call, ptr  # noqa: E402

SHAPES = [(4, 2, 32, 64, 64), (4, 2, 64, 64, 32), (3, 1, 64, 64, 16), (3, 1, 64, 32, 16), (1, 1, 32, 64, 16),
          (3, 1, 64, 64, 32), (1, 1, 64, 64, 32)]


def timed(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    print("NACC", os.environ.get("DMB_TC_NACC", "4"), "STAGES", os.environ.get("DMB_TC_STAGES", "4"), "batch", B)
    for ks, stride, cin, cout, W in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(ks * 100 + cin + cout)
        x = torch.randn(B, cin, W, W, device="cuda", generator=g)
        w = torch.randn(cout, cin, ks, ks, device="cuda", generator=g) * (cin * ks * ks) ** -0.5
        bias = torch.randn(cout, device="cuda", generator=g)
        wp = w.permute(1, 2, 3, 0).contiguous()
        Ho = W // stride
        n = C.c_int64(0)
        call("dmb_conv2d_tc_scratch_floats", B, cin, W, W, cout, ks, C.byref(n))
        scratch = torch.empty(n.value, device="cuda")
        xh = x.permute(0, 2, 3, 1).contiguous()
        y_tc = torch.empty(B, Ho, Ho, cout, device="cuda")
        y_cc = torch.empty(B, cout, Ho, Ho, device="cuda")

        def tc():
            call("dmb_conv2d_tc", ptr(xh), ptr(wp), ptr(bias), ptr(y_tc), B, cin, W, W, cout, ks, stride, 0, None, 0, 1,
                 ptr(scratch), st)

        def cc():
            call("dmb_conv2d_forward", ptr(x), ptr(wp), ptr(bias), ptr(y_cc), B, cin, W, W, cout, ks, stride, None, None,
                 0, 0, None, 0, st)

        t_tc, t_cc = timed(tc), timed(cc)
        nb = min(B, 16)
        ref = F.conv2d(x[:nb].double().cpu(), w.double().cpu(), bias.double().cpu(), stride=stride,
                       padding=0 if ks == 1 else 1)
        e_tc = float((y_tc[:nb].permute(0, 3, 1, 2).double().cpu() - ref).abs().max() / ref.abs().max())
        e_cc = float((y_cc[:nb].double().cpu() - ref).abs().max() / ref.abs().max())
        flops = 2.0 * B * Ho * Ho * cout * cin * ks * ks
        print(f"k{ks}s{stride} {cin:3d}->{cout:3d} @{W:3d}: tc {t_tc:7.3f} ms {flops / t_tc / 1e9:7.1f} TFLOP/s err {e_tc:.2e} | "
              f"cuda-core {t_cc:7.3f} ms {flops / t_cc / 1e9:7.1f} TFLOP/s err {e_cc:.2e} | x{t_cc / t_tc:.2f}")


if __name__ == "__main__":
    main()
