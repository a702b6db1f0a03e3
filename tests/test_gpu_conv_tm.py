"""-m gpu: the thin-channel tensor-core convolutions with the activation operand in tensor memory (csrc/conv_tm.cu:
tcgen05.mma with A in TMEM, 3xTF32 split in registers) through the layer-level C ABI (dmb_conv2d_tm, dmb_conv2d_tm_bn,
dmb_residual_layer_tm), against torch's CPU float64 conv2d on the same seeded inputs -- the arithmetic of the reference's
nn.Conv2d layers at the default widths (HiddenStateExtractor/vq_vae.py:280-289, :203-209, :222-225).  Tolerance 2e-6 of
max|y| (measured 1.3e-7 .. 4.9e-7; the path-level bar in BASELINE.json is 1e-4)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (ksize, stride, cin, cout, input width): every layer of the default VQ_VAE / VQ_VAE_z16 encoder behind the head
SHAPES = [(4, 2, 8, 16, 64), (4, 2, 16, 16, 32), (3, 1, 16, 16, 16), (3, 1, 16, 32, 16), (1, 1, 32, 16, 16)]
TOL = 2e-6


def _inputs(shape, B, seed=0):
    ks, stride, cin, cout, W = shape
    g = torch.Generator(device="cuda").manual_seed(ks * 1000 + cin * 10 + cout + B + seed)
    x = torch.randn(B, cin, W, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, ks, ks, device="cuda", generator=g) * (cin * ks * ks) ** -0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    return x, w, bias, g


def _scratch(cin, cout, ks):
    from dynamorph_b200._lib import call
    n = C.c_int64()
    call("dmb_conv2d_tm_scratch_floats", cin, cout, ks, C.byref(n))
    return torch.zeros(n.value, device="cuda")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ref(x, w, bias, ks, stride, in_relu=False, skip=None, out_relu=False):
    x = x.double().cpu()
    if in_relu:
        x = x.relu()
    y = F.conv2d(x, w.double().cpu(), bias.double().cpu(), stride=stride, padding=0 if ks == 1 else 1)
    if skip is not None:
        y = y + skip.double().cpu()
    return y.relu() if out_relu else y


def _err(y, ref):
    return float((y.double().cpu() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("B", [1, 3, 300])
def test_tm_conv_matches_torch(shape, B):
    """Plain, and with everything the eval-mode schedule asks of the layer (ReLU on load for the residual 3x3, skip for
    the 1x1, ReLU on store).  B = 1 and 3: fewer tiles than persistent CTAs; B = 300: every CTA loops and wraps."""
    from dynamorph_b200._lib import call, ptr
    ks, stride, cin, cout, W = shape
    x, w, bias, g = _inputs(shape, B)
    wp = w.permute(1, 2, 3, 0).contiguous()
    scratch = _scratch(cin, cout, ks)
    y = torch.empty(B, cout, W // stride, W // stride, device="cuda")
    call("dmb_conv2d_tm", ptr(x), ptr(wp), ptr(bias), ptr(y), B, cin, W, W, cout, ks, stride, 0, None, 0, ptr(scratch), _stream())
    assert _err(y, _ref(x, w, bias, ks, stride)) < TOL
    in_relu = int(ks == 3 and cout == 32)
    skip = torch.randn(B, cout, W // stride, W // stride, device="cuda", generator=g) if ks == 1 else None
    call("dmb_conv2d_tm", ptr(x), ptr(wp), ptr(bias), ptr(y), B, cin, W, W, cout, ks, stride, in_relu, ptr(skip), 1,
         ptr(scratch), _stream())
    assert _err(y, _ref(x, w, bias, ks, stride, bool(in_relu), skip, True)) < TOL


@pytest.mark.parametrize("B", [1, 5, 300])
def test_tm_fused_residual_layer(B):
    """y = x + conv1x1(relu(conv3x3(relu(x)) + b1)) + b2 in ONE kernel (two chained GEMMs, the second one's A operand
    written to tensor memory by the first one's epilogue) == the two-convolution chain in float64."""
    from dynamorph_b200._lib import call, ptr
    g = torch.Generator(device="cuda").manual_seed(40 + B)
    x = torch.randn(B, 16, 16, 16, device="cuda", generator=g)
    w1 = torch.randn(32, 16, 3, 3, device="cuda", generator=g) / 12
    b1 = torch.randn(32, device="cuda", generator=g) * 0.1
    w2 = torch.randn(16, 32, 1, 1, device="cuda", generator=g) / 6
    b2 = torch.randn(16, device="cuda", generator=g) * 0.1
    n = C.c_int64()
    call("dmb_residual_layer_tm_scratch_floats", C.byref(n))
    scratch = torch.zeros(n.value, device="cuda")
    y = torch.empty_like(x)
    call("dmb_residual_layer_tm", ptr(x), ptr(w1.permute(1, 2, 3, 0).contiguous()), ptr(b1),
         ptr(w2.permute(1, 2, 3, 0).contiguous()), ptr(b2), ptr(y), B, ptr(scratch), _stream())
    xd = x.double().cpu()
    ref = xd + F.conv2d(F.conv2d(xd.relu(), w1.double().cpu(), b1.double().cpu(), padding=1).relu(), w2.double().cpu(),
                        b2.double().cpu())
    assert _err(y, ref) < TOL


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("per_sample,in_relu", [(True, True), (False, True), (True, False)])
def test_tm_conv_batchnorm_form(shape, per_sample, in_relu):
    """Train-mode BatchNorm around the layer: relu?(x * scale + shift) applied on load (per patch and channel, or per
    channel; zero padding must stay zero although shift != 0), raw output stored, (sum, sum of squares) partials per
    (patch, warp of 32 output pixels, channel) whose fold equals the float64 sums of the output."""
    from dynamorph_b200._lib import call, ptr
    ks, stride, cin, cout, W = shape
    B = 7
    x, w, bias, g = _inputs(shape, B, seed=5)
    rows = B if per_sample else 1
    sc = torch.rand(rows, cin, device="cuda", generator=g) + 0.5
    sh = torch.randn(rows, cin, device="cuda", generator=g) * 0.5
    wp = w.permute(1, 2, 3, 0).contiguous()
    scratch = _scratch(cin, cout, ks)
    Ho = W // stride
    y = torch.empty(B, cout, Ho, Ho, device="cuda")
    bands = C.c_int32()
    stats = torch.full((B * (Ho * Ho // 32) * cout * 2,), float("nan"), dtype=torch.float64, device="cuda")
    call("dmb_conv2d_tm_bn", ptr(x), ptr(wp), ptr(bias), ptr(y), B, cin, W, W, cout, ks, stride, ptr(sc), ptr(sh),
         int(per_sample), int(in_relu), ptr(stats), C.byref(bands), ptr(scratch), _stream())
    torch.cuda.synchronize()
    assert bands.value == Ho * Ho // 32             # one partial row per warp (32 output pixels)
    xt = x.double().cpu() * sc.double().cpu().reshape(rows, cin, 1, 1) + sh.double().cpu().reshape(rows, cin, 1, 1)
    if in_relu:
        xt = xt.relu()
    ref = F.conv2d(xt, w.double().cpu(), bias.double().cpu(), stride=stride, padding=0 if ks == 1 else 1)
    assert _err(y, ref) < TOL
    part = stats[:B * bands.value * cout * 2].reshape(B, bands.value, cout, 2).cpu()
    got = part.sum(1)                                            # per (patch, channel)
    yd = y.double().cpu()
    assert torch.allclose(got[..., 0], yd.sum((2, 3)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(got[..., 1], (yd * yd).sum((2, 3)), rtol=1e-5, atol=1e-4)


def test_tm_rejects_other_shapes():
    from dynamorph_b200._lib import DmbError, call, ptr
    x = torch.zeros(1, 8, 32, 32, device="cuda")
    w = torch.zeros(8 * 16 * 16, device="cuda")
    b = torch.zeros(16, device="cuda")
    y = torch.zeros(1, 16, 16, 16, device="cuda")
    with pytest.raises(DmbError, match="thin encoder shapes"):
        call("dmb_conv2d_tm", ptr(x), ptr(w), ptr(b), ptr(y), 1, 8, 32, 32, 16, 4, 2, 0, None, 0, ptr(w), _stream())


# (ksize, stride, cin, cout, input width) of the data-gradient convolutions: residual 1x1 / 3x3, enc.10, and the stride-2
# convolutions that back-propagate through the decoder's ConvTranspose2d layers
DG_SHAPES = [(1, 1, 16, 32, 16), (3, 1, 32, 16, 16), (3, 1, 16, 16, 16), (4, 2, 8, 16, 64), (4, 2, 8, 16, 32),
             (4, 2, 16, 16, 32)]


@pytest.mark.parametrize("shape", DG_SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("B,extras", [(2, False), (5, True), (300, True)])
def test_tm_conv_data_gradient_form(shape, B, extras):
    """The training step's data-gradient form: plain gradient in, then gate by the producer's ReLU
    ([mask_src * s + t > 0]), skip gradient, and per-CTA partial sums (sum gx, sum gx * stat_src) whose fold equals the
    float64 sums -- the epilogue contract of the CUDA-core data-gradient kernels (csrc/conv_fwd.cu)."""
    from dynamorph_b200._lib import call, ptr
    ks, stride, cin, cout, W = shape
    gy, w, _, g = _inputs(shape, B, seed=11)
    wp = w.permute(1, 2, 3, 0).contiguous()
    n = C.c_int64()
    call("dmb_conv2d_tm_scratch_floats", cin, cout, ks, C.byref(n))
    scratch = torch.zeros(((n.value + 63) // 64) * 64 + cout, device="cuda")
    Ho = W // stride
    gx = torch.empty(B, cout, Ho, Ho, device="cuda")
    zero = torch.zeros(cout, device="cuda")
    ref = _ref(gy, w, zero, ks, stride)
    rows_max = C.c_int32()
    call("dmb_conv2d_tm_batch_stat_rows", C.byref(rows_max))
    if not extras:
        call("dmb_conv2d_tm_dgrad", ptr(gy), ptr(wp), ptr(gx), B, cin, W, W, cout, ks, stride, None, None, None, None,
             None, None, None, None, None, None, None, ptr(scratch), _stream())
        assert _err(gx, ref) < TOL
        return
    dual = stride == 1 and not (ks == 3 and cin == 32)       # BatchNorm backward applied on load
    y_raw = ga = gb = gc = None
    if dual:
        y_raw = torch.randn(B, cin, W, W, device="cuda", generator=g)
        ga = torch.rand(cin, device="cuda", generator=g) + 0.5
        gb = torch.randn(cin, device="cuda", generator=g) * 0.3
        gc = torch.randn(cin, device="cuda", generator=g) * 0.3
        gt = (gy.double() * ga.double().reshape(1, -1, 1, 1) + y_raw.double() * gb.double().reshape(1, -1, 1, 1)
              + gc.double().reshape(1, -1, 1, 1))
        ref = _ref(gt, w, zero, ks, stride)
    mask_src = torch.randn(B, cout, Ho, Ho, device="cuda", generator=g)
    ms = torch.rand(cout, device="cuda", generator=g) + 0.5
    mt = torch.randn(cout, device="cuda", generator=g) * 0.3
    skip = torch.randn(B, cout, Ho, Ho, device="cuda", generator=g)
    src = torch.randn(B, cout, Ho, Ho, device="cuda", generator=g)
    gate = ((mask_src * ms.reshape(1, -1, 1, 1) + mt.reshape(1, -1, 1, 1)) > 0).double().cpu()
    ref = ref * gate + skip.double().cpu()
    for stat_src in (src, None):
        stats = torch.full((rows_max.value * cout * 2,), float("nan"), dtype=torch.float64, device="cuda")
        rows = C.c_int32()
        call("dmb_conv2d_tm_dgrad", ptr(gy), ptr(wp), ptr(gx), B, cin, W, W, cout, ks, stride, ptr(y_raw), ptr(ga), ptr(gb),
             ptr(gc), ptr(mask_src), ptr(ms), ptr(mt), ptr(skip), ptr(stats), ptr(stat_src), C.byref(rows), ptr(scratch),
             _stream())
        torch.cuda.synchronize()
        assert _err(gx, ref) < TOL
        assert 0 < rows.value <= rows_max.value                   # one row per persistent CTA
        got = stats[:rows.value * cout * 2].reshape(rows.value, cout, 2).sum(0).cpu()
        gd = gx.double().cpu()
        other = src.double().cpu() if stat_src is not None else gd
        assert torch.allclose(got[:, 0], gd.sum((0, 2, 3)), rtol=1e-5, atol=1e-3)
        assert torch.allclose(got[:, 1], (gd * other).sum((0, 2, 3)), rtol=1e-5, atol=1e-3)


# transposed form: (cin, cout, input width, data_gradient)
CT_SHAPES = [(16, 8, 16, False), (16, 16, 16, True), (16, 8, 32, True)]


@pytest.mark.parametrize("shape", CT_SHAPES, ids=lambda s: "%dto%d_w%d_%s" % (s[0], s[1], s[2], "dg" if s[3] else "fwd"))
@pytest.mark.parametrize("B", [1, 5, 300])
def test_tm_conv_transpose_form(shape, B):
    """nn.ConvTranspose2d(4, 2, 1) as a 3x3 convolution on the input grid + pixel shuffle: the forward layer (bias, ReLU on
    store) and the data gradient of a stride-2 convolution (BatchNorm backward on load, gate, per-CTA sums) against
    torch's float64 conv_transpose2d."""
    from dynamorph_b200._lib import call, ptr
    cin, cout, W, dgrad = shape
    g = torch.Generator(device="cuda").manual_seed(cin * 100 + cout + W + B)
    x = torch.randn(B, cin, W, W, device="cuda", generator=g)
    w = torch.randn(cin, cout, 4, 4, device="cuda", generator=g) * (cin * 4) ** -0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    wp = w.permute(0, 2, 3, 1).contiguous()                     # [cin][4][4][cout]
    n = C.c_int64()
    call("dmb_conv2d_tm_scratch_floats", cin, 4 * cout, 3, C.byref(n))
    scratch = torch.zeros(((n.value + 63) // 64) * 64 + cout, device="cuda")
    y = torch.full((B, cout, 2 * W, 2 * W), float("nan"), device="cuda")
    if not dgrad:
        for relu in (0, 1):
            call("dmb_conv_transpose2d_tm", ptr(x), ptr(wp), ptr(bias), ptr(y), B, cin, W, W, cout, relu, 0, None, None, None,
                 None, None, None, None, None, None, None, ptr(scratch), _stream())
            ref = F.conv_transpose2d(x.double().cpu(), w.double().cpu(), bias.double().cpu(), stride=2, padding=1)
            assert _err(y, ref.relu() if relu else ref) < TOL
        return
    y_raw = torch.randn(B, cin, W, W, device="cuda", generator=g)
    ga = torch.rand(cin, device="cuda", generator=g) + 0.5
    gb = torch.randn(cin, device="cuda", generator=g) * 0.3
    gc = torch.randn(cin, device="cuda", generator=g) * 0.3
    mask_src = torch.randn(B, cout, 2 * W, 2 * W, device="cuda", generator=g)
    ms = torch.rand(cout, device="cuda", generator=g) + 0.5
    mt = torch.randn(cout, device="cuda", generator=g) * 0.3
    src = torch.randn(B, cout, 2 * W, 2 * W, device="cuda", generator=g)
    # plain data gradient first
    call("dmb_conv_transpose2d_tm", ptr(x), ptr(wp), None, ptr(y), B, cin, W, W, cout, 0, 1, None, None, None, None, None,
         None, None, None, None, None, ptr(scratch), _stream())
    assert _err(y, F.conv_transpose2d(x.double().cpu(), w.double().cpu(), None, stride=2, padding=1)) < TOL
    xt = (x.double() * ga.double().reshape(1, -1, 1, 1) + y_raw.double() * gb.double().reshape(1, -1, 1, 1)
          + gc.double().reshape(1, -1, 1, 1)).cpu()
    gate = ((mask_src * ms.reshape(1, -1, 1, 1) + mt.reshape(1, -1, 1, 1)) > 0).double().cpu()
    ref = F.conv_transpose2d(xt, w.double().cpu(), None, stride=2, padding=1) * gate
    rows_max = C.c_int32()
    call("dmb_conv2d_tm_batch_stat_rows", C.byref(rows_max))
    for stat_src in (mask_src, src, None):
        stats = torch.full((2 * rows_max.value * cout * 2,), float("nan"), dtype=torch.float64, device="cuda")
        rows = C.c_int32()
        call("dmb_conv_transpose2d_tm", ptr(x), ptr(wp), None, ptr(y), B, cin, W, W, cout, 0, 1, ptr(y_raw), ptr(ga), ptr(gb),
             ptr(gc), ptr(mask_src), ptr(ms), ptr(mt), ptr(stats), ptr(stat_src), C.byref(rows), ptr(scratch), _stream())
        torch.cuda.synchronize()
        assert _err(y, ref) < TOL
        assert 0 < rows.value <= rows_max.value                   # one row per persistent CTA
        got = stats[:rows.value * cout * 2].reshape(rows.value, cout, 2).sum(0).cpu()
        gd = y.double().cpu()
        other = stat_src.double().cpu() if stat_src is not None else gd
        assert torch.allclose(got[:, 0], gd.sum((0, 2, 3)), rtol=1e-5, atol=1e-3)
        assert torch.allclose(got[:, 1], (gd * other).sum((0, 2, 3)), rtol=1e-5, atol=1e-3)
