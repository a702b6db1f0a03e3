"""VectorQuantizer forward: tensor-core search (vq_tc.cu) vs the exhaustive CUDA-core kernel (vq.cu), same inputs.
usage: python scripts/time_vq.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamorph_b200._lib import call, ptr  # noqa: E402


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


st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for B, D, P, K in ((8192, 16, 256, 64), (1024, 64, 256, 512), (1024, 64, 1024, 512), (4096, 32, 256, 128)):
    g = torch.Generator(device="cuda").manual_seed(B + D + K)
    z = torch.randn(B, D, P, device="cuda", generator=g)
    pick = torch.randint(0, B * P, (K,), device="cuda", generator=g)
    cb = (z.permute(0, 2, 1).reshape(-1, D)[pick] + 0.05 * torch.randn(K, D, device="cuda", generator=g)).contiguous()
    out = {}
    for tc in ("1", "0"):
        os.environ["DMB_VQ_TC"] = tc
        zst = torch.empty_like(z)
        idx = torch.empty(B, P, dtype=torch.int32, device="cuda")
        stats = torch.zeros(2 + K, dtype=torch.float64, device="cuda")

        def run():
            call("dmb_vq_forward", ptr(z), ptr(cb), B, D, P, K, ptr(zst), ptr(idx), ptr(stats), st)

        ms = timed(run)
        stats.zero_()
        run()
        torch.cuda.synchronize()
        out[tc] = (ms, zst.clone(), idx.clone(), stats.clone())
    same = torch.equal(out["1"][2], out["0"][2]) and torch.equal(out["1"][1], out["0"][1])
    hist_same = torch.equal(out["1"][3][1:], out["0"][3][1:])
    loss_rel = float((out["1"][3][0] - out["0"][3][0]).abs() / out["0"][3][0])
    gb = B * P * (2 * D * 4 + 4) / 1e9
    print(f"B={B} D={D} P={P} K={K}: tensor-core {out['1'][0]:.3f} ms ({gb / out['1'][0] * 1e3:.0f} GB/s)  cuda-core "
          f"{out['0'][0]:.3f} ms  x{out['0'][0] / out['1'][0]:.2f}  identical idx/z_st: {same}  hist: {hist_same}  loss rel diff {loss_rel:.1e}")
