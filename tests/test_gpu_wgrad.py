"""Layer-level parity of the weight-gradient kernels (csrc/wgrad_tma.cu: TMA-fed, compile-time geometry, the default-width
layer shapes; csrc/wgrad.cu: generic) through the C ABI against float64 autograd of F.conv2d (the reference's
`total_loss.backward()`, run_training.py:406), with the transforms the training step folds into the loads."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (ks, stride, cin, cout, W): the shapes wgrad_tma.cu instantiates (ConvTranspose2d layers appear with swapped roles)
TMA_SHAPES = [(4, 2, 8, 16, 64), (4, 2, 16, 16, 32), (3, 1, 16, 16, 16), (3, 1, 16, 32, 16), (1, 1, 32, 16, 16),
              (4, 2, 8, 16, 32), (4, 2, 4, 8, 64), (4, 2, 4, 4, 128)]
GENERIC_SHAPES = [(3, 1, 8, 8, 32), (4, 2, 2, 8, 128), (1, 1, 4, 2, 64)]
TOL = 2e-5          # of max |dw|: fp32 sums of up to B*Ho*Wo products per element


def _stream():
    from dynamorph_b200._lib import STREAM
    return STREAM


def _run(shape, B, transforms, seed=0):
    from dynamorph_b200._lib import call, ptr
    ks, stride, cin, cout, W = shape
    g = torch.Generator(device="cuda").manual_seed(seed)
    Wo = W // stride
    x = torch.randn(B, cin, W, W, device="cuda", generator=g)
    gy = torch.randn(B, cout, Wo, Wo, device="cuda", generator=g)
    xs = xt = yr = ga = gb = gc = None
    relu = 0
    if transforms:
        xs = torch.rand(cin, device="cuda", generator=g) + 0.5
        xt = torch.randn(cin, device="cuda", generator=g) * 0.5
        relu = 1
        ga = torch.rand(cout, device="cuda", generator=g) + 0.5
        gc = torch.randn(cout, device="cuda", generator=g) * 0.1
        if transforms == 2:
            yr = torch.randn(B, cout, Wo, Wo, device="cuda", generator=g)
            gb = torch.randn(cout, device="cuda", generator=g) * 0.3
    n = C.c_int64()
    call("dmb_conv2d_weight_grad_scratch_floats", B, cin, W, W, cout, ks, stride, C.byref(n))
    scratch = torch.empty(n.value, device="cuda")
    dw = torch.full((cout, cin, ks, ks), float("nan"), device="cuda")
    db = torch.full((cout,), float("nan"), device="cuda")
    call("dmb_conv2d_weight_grad", ptr(x), ptr(gy), ptr(dw), ptr(db), B, cin, W, W, cout, ks, stride, ptr(xs), ptr(xt),
         relu, ptr(yr), ptr(ga), ptr(gb), ptr(gc), ptr(scratch), _stream())
    torch.cuda.synchronize()
    # float64 reference through autograd
    act = x.double().cpu()
    if xs is not None:
        act = (act * xs.double().cpu().view(1, -1, 1, 1) + xt.double().cpu().view(1, -1, 1, 1)).relu()
    gp = gy.double().cpu()
    if ga is not None:
        gp = gp * ga.double().cpu().view(1, -1, 1, 1) + gc.double().cpu().view(1, -1, 1, 1)
        if yr is not None:
            gp = gp + yr.double().cpu() * gb.double().cpu().view(1, -1, 1, 1)
    w = torch.zeros(cout, cin, ks, ks, dtype=torch.float64, requires_grad=True)
    bias = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(act, w, bias, stride=stride, padding=0 if ks == 1 else 1)
    (y * gp).sum().backward()
    return dw.cpu().double(), db.cpu().double(), w.grad, bias.grad


def _check(dw, db, rw, rb):
    assert torch.isfinite(dw).all() and torch.isfinite(db).all()
    assert float((dw - rw).abs().max() / rw.abs().max()) < TOL
    assert float((db - rb).abs().max() / rb.abs().max().clamp_min(1e-30)) < TOL


@pytest.mark.parametrize("shape", TMA_SHAPES + GENERIC_SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("transforms", [0, 1, 2], ids=["plain", "affine", "bn_backward"])
def test_weight_grad_matches_float64(shape, transforms):
    _check(*_run(shape, 5, transforms, seed=3))


@pytest.mark.parametrize("shape", TMA_SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("B", [1, 37, 300])
def test_weight_grad_batch_sizes(shape, B):
    """One patch (fewer work items than CTAs), a ragged count, and more items than persistent CTAs (every CTA wraps)."""
    _check(*_run(shape, B, 2, seed=B))


def test_tma_and_generic_kernels_agree():
    """The two kernels own the same accumulators and walk the pixels in the same order inside a work item; they differ
    in band height (items per patch), so they agree to fp32 round-off, not bit for bit."""
    shape = (3, 1, 16, 32, 16)
    dw1, db1, rw, _ = _run(shape, 64, 2, seed=9)
    os.environ["DMB_WGRAD_TMA"] = "0"
    try:
        dw0, db0, _, _ = _run(shape, 64, 2, seed=9)
    finally:
        del os.environ["DMB_WGRAD_TMA"]
    assert float((dw1 - dw0).abs().max() / rw.abs().max()) < TOL
    assert float((db1 - db0).abs().max() / db0.abs().max()) < TOL


def test_weight_grad_is_deterministic():
    shape = (4, 2, 8, 16, 64)
    a = _run(shape, 40, 2, seed=1)
    b = _run(shape, 40, 2, seed=1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
