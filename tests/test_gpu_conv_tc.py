"""-m gpu: the tensor-core convolution (csrc/conv_tc.cu: tcgen05 implicit GEMM, 3xTF32 operand split, fp32 accumulation
in TMEM) through the layer-level C ABI (dmb_conv2d_tc), against torch's CPU float64 conv2d on the same seeded inputs
(the arithmetic of the reference's nn.Conv2d layers at the 64-wide widths, HiddenStateExtractor/vq_vae.py:279-289,
:203-209; vae.py:401-407).  Tolerance 1e-5 of max|y|: the split keeps 21+ mantissa bits per operand, so the error is
fp32-round-off sized (the path-level bar in BASELINE.json is 1e-4)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (ksize, stride, cin, cout, input width): every wide layer of VQ_VAE(64, 32, K) and VQ_VAE_z32(64, 64, K)
SHAPES = [(4, 2, 32, 64, 64), (4, 2, 64, 64, 32), (3, 1, 64, 64, 16), (3, 1, 64, 32, 16), (1, 1, 32, 64, 16),
          (3, 1, 64, 64, 32), (1, 1, 64, 64, 32)]


def run_tc(x, w, bias, ks, stride, in_relu=False, skip=None, out_relu=False, nhwc=False):
    from dynamorph_b200._lib import call, ptr
    B, cin, H, W = x.shape
    cout = w.shape[0]
    wp = w.permute(1, 2, 3, 0).contiguous()          # [Cin][k][k][Cout]
    n = C.c_int64(0)
    call("dmb_conv2d_tc_scratch_floats", B, cin, H, W, cout, ks, C.byref(n))
    scratch = torch.empty(n.value, device=x.device)
    Ho, Wo = H // stride, W // stride
    if nhwc:
        xin = x.permute(0, 2, 3, 1).contiguous()
        sk = skip.permute(0, 2, 3, 1).contiguous() if skip is not None else None
        y = torch.empty(B, Ho, Wo, cout, device=x.device)
    else:
        xin, sk = x, skip
        y = torch.empty(B, cout, Ho, Wo, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call("dmb_conv2d_tc", ptr(xin), ptr(wp), ptr(bias), ptr(y), B, cin, H, W, cout, ks, stride, int(in_relu),
         ptr(sk) if sk is not None else None, int(out_relu), int(nhwc), ptr(scratch), st)
    torch.cuda.synchronize()
    return y.permute(0, 3, 1, 2) if nhwc else y


def reference(x, w, bias, ks, stride, in_relu=False, skip=None, out_relu=False):
    x = x.double().cpu()
    if in_relu:
        x = x.relu()
    y = F.conv2d(x, w.double().cpu(), bias.double().cpu(), stride=stride, padding=0 if ks == 1 else 1)
    if skip is not None:
        y = y + skip.double().cpu()
    return y.relu() if out_relu else y


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "k%ds%d_%dto%d_w%d" % s)
@pytest.mark.parametrize("B", [1, 5])
@pytest.mark.parametrize("nhwc", [False, True], ids=["nchw", "nhwc"])
def test_tc_conv_matches_torch(shape, B, nhwc):
    ks, stride, cin, cout, W = shape
    g = torch.Generator(device="cuda").manual_seed(ks * 1000 + cin * 10 + cout + B)
    x = torch.randn(B, cin, W, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, ks, ks, device="cuda", generator=g) * (cin * ks * ks) ** -0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    y = run_tc(x, w, bias, ks, stride, nhwc=nhwc)
    ref = reference(x, w, bias, ks, stride)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"plain: {err:.3e}"
    skip = torch.randn(B, cout, W // stride, W // stride, device="cuda", generator=g)
    kw = dict(in_relu=True, skip=skip, out_relu=True)
    y = run_tc(x, w, bias, ks, stride, nhwc=nhwc, **kw)
    ref = reference(x, w, bias, ks, stride, **kw)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, f"relu on load + skip + relu: {err:.3e}"


def test_tc_conv_error_is_fp32_sized():
    """The 3xTF32 split must not be a disguised single-pass TF32 (error ~1e-3): compare with an fp32 FFMA conv."""
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(8, 64, 32, 32, device="cuda", generator=g)
    w = torch.randn(64, 64, 4, 4, device="cuda", generator=g) / 32.0
    bias = torch.zeros(64, device="cuda")
    y = run_tc(x, w, bias, 4, 2)
    ref = reference(x, w, bias, 4, 2)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, f"{err:.3e}"


def test_tc_conv_rejects_thin_layers():
    from dynamorph_b200._lib import call, ptr
    x = torch.zeros(1, 16, 16, 16, device="cuda")
    with pytest.raises(RuntimeError):
        call("dmb_conv2d_tc", ptr(x), ptr(x), ptr(x), ptr(x), 1, 16, 16, 16, 16, 3, 1, 0, None, 0, 0, ptr(x),
             C.c_void_p(0))


@pytest.mark.parametrize("arch,rh", [("z16", 32), ("z32", 64)], ids=["VQ_VAE_64_512", "VQ_VAE_z32_64_64_512"])
def test_wide_encoder_on_tensor_cores_matches_oracle(arch, rh, monkeypatch):
    """BASELINE configs[3] through the drop-in classes in eval mode (the tcgen05 encoder, csrc/model.cu:
    run_encoder_tc) against the CPU oracle on seeded inputs; and against the CUDA-core schedule (DMB_TC=0)."""
    import gpu_util as U
    from oracle import vqvae_oracle as O
    st = O.default_state(arch, num_hiddens=64, num_residual_hiddens=rh, num_embeddings=512, seed=3)
    st = O.calibrate_state(st, O.synthetic_patches(8, 5), seed=1)
    x = O.synthetic_patches(6, 11)
    with torch.no_grad():
        zb_ref = O.encoder(x, st, O.EVAL)
    idx_ref = O.vq_indices(zb_ref, st["vq.w.weight"])
    m = U.model_from_state(st).eval()
    from dynamorph_b200._lib import load
    lib = load()
    m.encode_latents(x.cuda(), "eval")             # first call packs the weights
    lib.dmb_launch_count(1)
    zb, za, idx = m.encode_latents(x.cuda(), "eval")
    n_tc = lib.dmb_launch_count(1)
    assert U.rel(zb, zb_ref) < U.REL_TOL
    flips = U.check_indices(idx, zb_ref, st["vq.w.weight"], idx_ref, arch)
    if flips == 0:
        assert U.rel(za, O.vq_gather(idx_ref, st["vq.w.weight"])) < U.REL_TOL
    monkeypatch.setenv("DMB_TC", "0")
    zb0, za0, idx0 = m.encode_latents(x.cuda(), "eval")
    n_cc = lib.dmb_launch_count(1)
    assert n_tc in (n_cc, n_cc + 1), (n_tc, n_cc)  # same layer count (+ a transpose when the head has no NHWC form)
    assert U.rel(zb, zb0) < 1e-5
    # ragged batch: one patch, odd count
    for nb in (1, 3):
        zb1, _, idx1 = m.encode_latents(x[:nb].cuda(), "eval")
        monkeypatch.delenv("DMB_TC")
        zb2, _, idx2 = m.encode_latents(x[:nb].cuda(), "eval")
        monkeypatch.setenv("DMB_TC", "0")
        assert torch.equal(zb2, zb[:nb]) and torch.equal(idx2, idx[:nb])
        assert U.rel(zb1, zb2) < 1e-5


def run_wino(x, w, bias, in_relu=False, out_relu=False):
    from dynamorph_b200._lib import call, ptr
    B, cin, H, W = x.shape
    cout = w.shape[0]
    wp = w.permute(1, 2, 3, 0).contiguous()          # [Cin][3][3][Cout]
    scratch = torch.empty(2 * 16 * cin * cout, device=x.device)
    y = torch.empty(B, cout, H, W, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call("dmb_conv2d_wino", ptr(x), ptr(wp), ptr(bias), ptr(y), B, cin, H, W, cout, int(in_relu), int(out_relu),
         None, None, None, ptr(scratch), st)
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("cout", [32, 16])
@pytest.mark.parametrize("B", [1, 2, 5, 301])
def test_winograd_tc_conv_matches_torch(cout, B):
    """csrc/conv_wino_tc.cu: Winograd F(2x2,3x3) as 16 TF32x3 GEMMs in tensor memory, against an fp64 convolution
    (the 16-channel 3x3 layers of the default configuration, vq_vae.py:203-209, :288)."""
    g = torch.Generator(device="cuda").manual_seed(100 * cout + B)
    x = torch.randn(B, 16, 16, 16, device="cuda", generator=g)
    w = torch.randn(cout, 16, 3, 3, device="cuda", generator=g) / 12.0
    bias = torch.randn(cout, device="cuda", generator=g)
    for in_relu, out_relu in ((False, False), (True, True)):
        y = run_wino(x, w, bias, in_relu, out_relu)
        ref = reference(x, w, bias, 3, 1, in_relu=in_relu, out_relu=out_relu)
        err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
        assert err < 5e-6, f"relu {in_relu}/{out_relu}: {err:.3e}"


def test_default_encoder_with_winograd_tc_matches_golden(monkeypatch):
    """The eval-mode encoder of the DEFAULT configuration with its three latent-resolution 3x3 layers on the Winograd
    tensor-core kernel (taken for batches >= 512 by default; forced here) against the fixture written by the
    unmodified reference, and against the direct CUDA-core schedule."""
    import gpu_util as U
    from conftest import Golden
    g = Golden("vqvae_default")
    st = g.state()
    m = U.model_from_state(st).eval()
    x = g.t("x_eval").cuda()
    monkeypatch.setenv("DMB_WINO_MIN_B", "1")
    zb, za, idx = m.encode_latents(x, "eval")
    assert U.rel(zb, g.t("eval/z_before")) < U.REL_TOL
    flips = U.check_indices(idx, g.t("eval/z_before"), st["vq.w.weight"], g["eval/idx"], "winograd-tc")
    if flips == 0:
        assert U.rel(za, g.t("eval/z_after")) < U.REL_TOL
    monkeypatch.setenv("DMB_WINO", "0")
    zb0, _, idx0 = m.encode_latents(x, "eval")
    assert not torch.equal(zb, zb0), "the Winograd kernel was not taken"
    assert U.rel(zb, zb0) < 1e-5
    # odd batch (the kernel works on pairs of patches)
    monkeypatch.delenv("DMB_WINO")
    zb3, _, idx3 = m.encode_latents(x[:3], "eval")
    assert torch.equal(zb3, zb[:3]) and torch.equal(idx3, idx[:3])


@pytest.mark.parametrize("D,K,B,P", [(16, 16, 3, 100), (32, 100, 5, 77), (64, 250, 2, 129), (16, 512, 1, 256),
                                      (64, 48, 7, 33)])
def test_vq_tensor_core_search_is_bit_identical_to_exhaustive(D, K, B, P, monkeypatch):
    """csrc/vq_tc.cu against csrc/vq.cu (the exhaustive direct-form search in torch's summation order, itself checked
    against the reference in test_gpu_encode.py) on ragged sizes: codes not a multiple of 16, positions not a multiple
    of the 128-position tile, near-duplicate codes, a collapsed codebook (overflow path) and NaN input (fallback)."""
    from dynamorph_b200._lib import call, ptr
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(D * 1000 + K)
    z = torch.randn(B, D, P, device="cuda", generator=g)
    flat = z.permute(0, 2, 1).reshape(-1, D)
    cb = flat[torch.randint(0, B * P, (K,), device="cuda", generator=g)].clone()
    cb[1::3] += 1e-6 * torch.randn_like(cb[1::3])                    # near-duplicates of sampled positions
    cb[2::5] = cb[0]                                                 # exact duplicates: first index must win
    cases = {"sampled": cb.contiguous(), "collapsed": cb[:1].expand(K, D).contiguous()}
    zn = z.clone()
    zn[0, :, 0] = float("nan")
    for name, codebook in cases.items():
        for zin in (z, zn):
            out = {}
            for tc in ("1", "0"):
                monkeypatch.setenv("DMB_VQ_TC", tc)
                zst = torch.zeros_like(zin)
                idx = torch.full((B, P), -7, dtype=torch.int32, device="cuda")
                stats = torch.zeros(2 + K, dtype=torch.float64, device="cuda")
                call("dmb_vq_forward", ptr(zin), ptr(codebook), B, D, P, K, ptr(zst), ptr(idx), ptr(stats), st)
                torch.cuda.synchronize()
                out[tc] = (zst, idx, stats)
            assert torch.equal(out["1"][1], out["0"][1]), name
            assert torch.equal(out["1"][0].nan_to_num(7.0), out["0"][0].nan_to_num(7.0)), name
            assert torch.equal(out["1"][2][1:], out["0"][2][1:]), name           # position count + histogram
            a, b = out["1"][2][0], out["0"][2][0]
            assert (torch.isnan(a) and torch.isnan(b)) or abs(float(a - b)) <= 1e-9 * abs(float(b)), name


@pytest.mark.parametrize("B", [1, 3, 300])
def test_winograd_tc_fused_residual_tail_matches_torch(B):
    """conv3x3(ReLU(x)) -> ReLU -> conv1x1 + x in one kernel (the eval-mode ResidualBlock layer, vq_vae.py:203-209 with
    BatchNorm folded) against fp64 torch."""
    from dynamorph_b200._lib import call, ptr
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.randn(B, 16, 16, 16, device="cuda", generator=g)
    w3 = torch.randn(32, 16, 3, 3, device="cuda", generator=g) / 12.0
    b3 = torch.randn(32, device="cuda", generator=g)
    w1 = torch.randn(16, 32, 1, 1, device="cuda", generator=g) / 6.0
    b1 = torch.randn(16, device="cuda", generator=g)
    w3p = w3.permute(1, 2, 3, 0).contiguous()
    w1p = w1.permute(1, 2, 3, 0).contiguous()          # [32][1][1][16]
    scratch = torch.empty(2 * 16 * 16 * 32, device="cuda")
    y2 = torch.empty(B, 16, 16, 16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call("dmb_conv2d_wino", ptr(x), ptr(w3p), ptr(b3), None, B, 16, 16, 16, 32, 1, 1, ptr(w1p), ptr(b1), ptr(y2),
         ptr(scratch), st)
    torch.cuda.synchronize()
    xd = x.double().cpu()
    mid = F.conv2d(xd.relu(), w3.double().cpu(), b3.double().cpu(), padding=1).relu()
    ref = xd + F.conv2d(mid, w1.double().cpu(), b1.double().cpu())
    err = float((y2.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, f"{err:.3e}"


def test_default_encoder_with_fused_winograd_tail(monkeypatch):
    """DMB_WINO_FUSE=1: the whole eval-mode residual layer in one kernel, through the model path."""
    import gpu_util as U
    from conftest import Golden
    g = Golden("vqvae_default")
    st = g.state()
    m = U.model_from_state(st).eval()
    x = g.t("x_eval").cuda()
    monkeypatch.setenv("DMB_WINO_MIN_B", "1")
    zb0, _, idx0 = m.encode_latents(x, "eval")
    monkeypatch.setenv("DMB_WINO_FUSE", "1")
    zb, _, idx = m.encode_latents(x, "eval")
    assert U.rel(zb, g.t("eval/z_before")) < U.REL_TOL
    U.check_indices(idx, g.t("eval/z_before"), st["vq.w.weight"], g["eval/idx"], "winograd-tc fused")
    assert not torch.equal(zb, zb0) and U.rel(zb, zb0) < 1e-5
