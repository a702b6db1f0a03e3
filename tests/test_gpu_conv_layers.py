"""-m gpu: every convolution shape the TMA kernel (csrc/conv_tma.cu) is instantiated for, through the layer-level
C ABI (dmb_conv2d_forward), against torch's CPU float64 conv2d on the same seeded inputs (the arithmetic the
reference's nn.Conv2d performs, HiddenStateExtractor/vq_vae.py:203-209, :276-289).  Tolerance 1e-5 of max|y|
(fp32 accumulation-order noise is ~1e-6; the path-level bar in BASELINE.json is 1e-4)."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (ksize, stride, cin, cout, width) == DMB_TMA_SHAPES in csrc/conv_tma.cu, plus two shapes only the generic kernel serves
SHAPES = [(4, 2, 2, 8, 128), (4, 2, 8, 16, 64), (4, 2, 16, 16, 32), (3, 1, 16, 16, 16), (3, 1, 16, 32, 16),
          (1, 1, 32, 16, 16), (3, 1, 16, 32, 32), (1, 1, 32, 16, 32),
          (4, 2, 2, 32, 128), (4, 2, 32, 64, 64), (4, 2, 64, 64, 32), (3, 1, 64, 64, 16), (3, 1, 64, 32, 16),
          (1, 1, 32, 64, 16), (3, 1, 64, 64, 32), (1, 1, 64, 64, 32),
          (3, 1, 8, 8, 64), (4, 2, 4, 8, 32)]


def run_layer(x, w, bias, ks, stride, in_scale=None, in_shift=None, per_sample=False, in_relu=False, skip=None,
              out_relu=False):
    from dynamorph_b200._lib import call, ptr
    B, cin, H, W = x.shape
    cout = w.shape[0]
    wp = w.permute(1, 2, 3, 0).contiguous()          # [Cin][k][k][Cout]
    y = torch.empty(B, cout, H // stride, W // stride, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call("dmb_conv2d_forward", ptr(x), ptr(wp), ptr(bias), ptr(y), B, cin, H, W, cout, ks, stride,
         ptr(in_scale) if in_scale is not None else None, ptr(in_shift) if in_shift is not None else None,
         int(per_sample), int(in_relu), ptr(skip) if skip is not None else None, int(out_relu), st)
    torch.cuda.synchronize()
    return y


def reference(x, w, bias, ks, stride, in_scale=None, in_shift=None, per_sample=False, in_relu=False, skip=None,
              out_relu=False):
    x = x.double().cpu()
    if in_scale is not None:
        s, t = in_scale.double().cpu(), in_shift.double().cpu()
        shape = (x.shape[0], x.shape[1], 1, 1) if per_sample else (1, x.shape[1], 1, 1)
        x = x * s.view(shape) + t.view(shape)
    if in_relu:
        x = x.relu()
    y = F.conv2d(x, w.double().cpu(), bias.double().cpu(), stride=stride, padding=0 if ks == 1 else 1)
    if skip is not None:
        y = y + skip.double().cpu()
    return y.relu() if out_relu else y


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("B", [1, 5])
@pytest.mark.parametrize("weights", ["smem", "const"])
def test_conv_layer_matches_torch(shape, B, weights, monkeypatch):
    # "const" forces the constant-pool weight form of csrc/conv_tma.cu (taken by default only for large launches)
    monkeypatch.setenv("DMB_CONV_WEIGHTS", weights)
    ks, stride, cin, cout, W = shape
    g = torch.Generator(device="cuda").manual_seed(ks * 1000 + cin * 10 + cout + B)
    x = torch.randn(B, cin, W, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, ks, ks, device="cuda", generator=g) * (cin * ks * ks) ** -0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    # plain
    y = run_layer(x, w, bias, ks, stride)
    ref = reference(x, w, bias, ks, stride)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"plain: {err:.3e}"
    # producer BatchNorm affine (per sample) + ReLU on load, skip tensor, ReLU on store
    sc = torch.rand(B, cin, device="cuda", generator=g) + 0.5
    sh = torch.randn(B, cin, device="cuda", generator=g) * 0.3
    skip = torch.randn(B, cout, W // stride, W // stride, device="cuda", generator=g)
    kw = dict(in_scale=sc, in_shift=sh, per_sample=True, in_relu=True, skip=skip, out_relu=True)
    y = run_layer(x, w, bias, ks, stride, **kw)
    ref = reference(x, w, bias, ks, stride, **kw)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"transform+skip+relu: {err:.3e}"
    # ReLU on load only (eval-mode residual block), per-channel tables
    kw = dict(in_relu=True)
    y = run_layer(x, w, bias, ks, stride, **kw)
    ref = reference(x, w, bias, ks, stride, **kw)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"relu on load: {err:.3e}"
