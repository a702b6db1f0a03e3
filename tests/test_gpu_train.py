"""-m gpu: training step (forward with batch statistics, backward, Adam) against fixtures produced by the
unmodified reference module + torch.optim.Adam driven as run_training.run_one_batch drives them."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu

Z16_CASES = ["vqvae_default", "z16_masked", "vqvae_heavy", "z32_default"]


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


def _mask(g):
    return g.t("mask_train").cuda() if g.has("mask_train") else None


@pytest.mark.parametrize("name", Z16_CASES)
def test_train_forward_matches_reference(name, U):
    g = Golden(name)
    st = g.state()
    m = U.model_from_state(st).train()
    x = g.t("x_train").cuda()
    dec, d = m(x, batch_mask=_mask(g))
    assert U.rel(dec, g.t("train/decoded")) < U.REL_TOL
    for k in ("recon_loss", "commitment_loss", "total_loss", "perplexity"):
        ref = float(g["train/loss/" + k])
        assert abs(float(d[k]) - ref) <= U.REL_TOL * abs(ref), k
    assert d["total_loss"].requires_grad
    sd = m.state_dict()
    for k, v in g.group("train/after_fwd").items():
        assert U.rel(sd[k], v) < U.REL_TOL, k


STRICT = 2e-4          # of the tensor's max-abs (1e-4-level, fp32)


def _check_grads(model, grads_ref, noise, envelope=None, what=""):
    """Every gradient element within STRICT * max|reference tensor| of the reference, plus -- element by element --
    the gate-flip envelope of the ReLU inputs that sit within 2e-6 of zero in the reference forward
    (oracle.relu_gate_envelopes: exactly what flipping those gates adds or removes; zero where none can reach).
    Conv biases that feed a train-mode BatchNorm have an exactly-zero gradient that autograd reports as rounding
    noise: those are only required to be negligible.  Returns the worst error in units of the strict bound."""
    named = dict(model.named_parameters())
    scale = max(float(v.abs().max()) for k, v in grads_ref.items() if k not in noise)
    worst = 0.0
    for k, ref in grads_ref.items():
        got = named[k].grad
        assert got is not None, k
        got = got.detach().cpu()
        if k in noise:
            assert float(got.abs().max()) <= 1e-5 * scale, (k, float(got.abs().max()))
            continue
        allowed = STRICT * float(ref.abs().max().clamp(min=1e-30))
        slack = (got - ref).abs() - (1.05 * envelope[k] if envelope is not None else 0.0)
        e = float(slack.max()) / allowed
        worst = max(worst, e)
        assert e < 1.0, (what, k, e, "x the strict bound after the per-element gate-flip envelope")
    assert named["channel_var"].grad is None
    return worst


@pytest.mark.parametrize("name", Z16_CASES)
def test_gradients_match_reference(name, U):
    """All 43 parameter gradients against the reference module's autograd (fixtures)."""
    g = Golden(name)
    st = g.state()
    m = U.model_from_state(st).train()
    x = g.t("x_train").cuda()
    m.zero_grad()
    mask = g.t("mask_train") if g.has("mask_train") else None
    _, d = m(x, batch_mask=_mask(g))
    d["total_loss"].backward()
    env, ties = O.relu_gate_envelopes(g.t("x_train"), st, O.BATCH, batch_mask=mask,
                                      commitment_cost=float(g["hp/commitment_cost"]))
    worst = _check_grads(m, g.group("train/grad"), set(O.bias_feeds_train_bn(st)), env, name)
    print(f"{name}: relu near-ties {ties}, worst grad err {worst:.2f} x strict")


def test_gradients_heavy_over_seeds(U):
    """Several seeded batches of the 64-wide configuration: each within the strict bound plus its own per-element
    gate-flip envelope (most batches have no near-tie at all and the envelope is identically zero)."""
    g = Golden("vqvae_heavy")
    st = g.state()
    m = U.model_from_state(st).train()
    noise = set(O.bias_feeds_train_bn(st))
    for seed in range(5):
        x = O.synthetic_patches(2, 5000 + seed)
        m.load_state_dict(st)
        m.zero_grad()
        _, d = m(x.cuda())
        d["total_loss"].backward()
        _, _, grads, _ = O.loss_and_grads(x, st, O.BATCH)
        env, ties = O.relu_gate_envelopes(x, st, O.BATCH)
        worst = _check_grads(m, grads, noise, env, f"seed {seed}")
        print(f"seed {seed}: near-ties {ties}, worst {worst:.2f} x strict")


def _check_after_steps(g, st, model, steps, lr):
    """Several Adam steps on a 3-4 patch batch are chaotic at the parameter level (the update is
    ~lr*sign(g) for every element, so elements whose gradient hovers around zero diverge by O(lr) per step);
    what must hold: counters, running statistics, and every element within Adam's reach of the reference."""
    sd = model.state_dict()
    for k, ref in g.group("train/after_steps").items():
        got = sd[k].detach().cpu()
        if k.endswith("num_batches_tracked"):
            assert int(got) == int(ref), k
            continue
        diff = (got.double() - ref.double()).abs()
        if "running" in k:
            assert float(diff.max() / ref.abs().max().clamp(min=1e-30)) < 1e-3, k
        elif k == "channel_var":
            assert float(diff.max()) == 0.0
        else:
            assert float(diff.max()) <= 2.01 * steps * lr, k


@pytest.mark.parametrize("name", ["vqvae_default", "z16_masked", "z32_default"])
@pytest.mark.parametrize("opt_kind", ["torch_adam", "fused_adam"])
def test_first_adam_step_elementwise(name, opt_kind, U):
    """One step from the fixture's state: expected parameters = torch.optim.Adam applied (on the CPU) to the
    reference's own gradients.  Where the gradient is significant the match must be tight."""
    from dynamorph_b200.run_training import run_one_batch
    from dynamorph_b200.optim import FusedAdam
    g = Golden(name)
    st = g.state()
    lr = float(g["train/lr"])
    gref = g.group("train/grad")
    exp = {k: torch.nn.Parameter(st[k].clone()) for k in gref}
    for k in exp:
        exp[k].grad = gref[k].clone()
    torch.optim.Adam(list(exp.values()), lr=lr, betas=(.9, .999)).step()
    m = U.model_from_state(st).train()
    opt = torch.optim.Adam(m.parameters(), lr=lr, betas=(.9, .999)) if opt_kind == "torch_adam" \
        else FusedAdam(m, lr=lr, betas=(.9, .999))
    m, tl = run_one_batch(m, g.t("x_train").cuda(), {}, model_kwargs={"batch_mask": _mask(g)}, optimizer=opt,
                          transform=None, training=True)
    assert abs(tl["total_loss"][0] - float(g["train/curve/total_loss"][0])) <= 1e-4 * abs(tl["total_loss"][0])
    sd = m.state_dict()
    noise = set(O.bias_feeds_train_bn(st))
    for k, e in exp.items():
        diff = (sd[k].detach().cpu() - e.detach()).abs()
        assert float(diff.max()) <= 2.01 * lr, k
        if k in noise:
            continue
        sig = gref[k].abs() > 1e-3 * gref[k].abs().max()
        assert float(diff[sig].max()) <= 2e-3 * lr + 1e-7, (k, float(diff[sig].max()))
    assert all(p.grad is None for p in m.parameters())       # model.zero_grad() after the step


@pytest.mark.parametrize("name", ["vqvae_default", "z16_masked", "z32_default"])
@pytest.mark.parametrize("opt_kind", ["torch_adam", "fused_adam"])
def test_run_one_batch_adam_steps(name, opt_kind, U):
    from dynamorph_b200.run_training import run_one_batch
    from dynamorph_b200.optim import FusedAdam
    g = Golden(name)
    st = g.state()
    m = U.model_from_state(st).train()
    x = g.t("x_train").cuda()
    steps, lr = int(g["train/steps"]), float(g["train/lr"])
    opt = torch.optim.Adam(m.parameters(), lr=lr, betas=(.9, .999)) if opt_kind == "torch_adam" \
        else FusedAdam(m, lr=lr, betas=(.9, .999))
    tl = {}
    for _ in range(steps):
        m, tl = run_one_batch(m, x.clone(), tl, model_kwargs={"batch_mask": _mask(g)}, optimizer=opt,
                              transform=None, training=True)
    assert set(tl) == {"recon_loss", "commitment_loss", "time_matching_loss", "total_loss", "perplexity"}
    assert np.allclose(tl["total_loss"], g["train/curve/total_loss"], rtol=5e-4)
    assert np.allclose(tl["recon_loss"], g["train/curve/recon_loss"], rtol=5e-4)
    _check_after_steps(g, st, m, steps, lr)


def test_validation_batch_does_not_step(U):
    from dynamorph_b200.run_training import run_one_batch
    g = Golden("vqvae_default")
    m = U.model_from_state(g.state()).train()
    before = m._engine.flat_params.clone()
    m, vl = run_one_batch(m, g.t("x_train").cuda(), {}, model_kwargs={}, optimizer=None, transform=None, training=False)
    assert torch.equal(before, m._engine.flat_params)
    assert len(vl["total_loss"]) == 1


def test_fused_adam_matches_torch_adam():
    import ctypes as C
    from dynamorph_b200._lib import call, ptr
    torch.manual_seed(0)
    n = 10007
    p = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=3e-3, betas=(.9, .999))
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n, device="cuda") * (10.0 ** torch.randint(-4, 2, (n,), device="cuda").float())
        ref.grad = g.clone()
        opt.step()
        call("dmb_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), n, 3e-3, 0.9, 0.999, 1e-8, step, 1.0,
             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert float((p - ref.detach()).abs().max()) < 2e-6


def test_zscore_patch_device():
    from dynamorph_b200.pipeline.train_utils import zscore_patch, zscore_patch_device
    z = Golden("zscore_patch")
    raw = np.squeeze(z["raw"])
    out = zscore_patch(raw)
    assert out.dtype == np.float32 and out.shape == raw.shape
    ref = z["z"].astype(np.float32)
    assert np.allclose(out, ref, rtol=0, atol=2e-6 * np.abs(ref).max())
    u16 = torch.from_numpy(raw.astype(np.uint16)).cuda()
    o2 = zscore_patch_device(u16).cpu().numpy()
    r2 = O.zscore_patch(raw.astype(np.uint16).astype(np.float64)).astype(np.float32)
    assert np.allclose(o2, r2, atol=2e-6 * np.abs(r2).max())


@pytest.mark.parametrize("name", ["vqvae_default", "z32_default"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_trainer_matches_run_one_batch(name, use_graph, U):
    """The CUDA-graph trainer and the autograd path are the same arithmetic: identical losses and parameters."""
    from dynamorph_b200.run_training import run_one_batch
    from dynamorph_b200.optim import FusedAdam
    from dynamorph_b200.trainer import FusedTrainer
    g = Golden(name)
    st = g.state()
    x = g.t("x_train").cuda()
    lr = 1e-3
    m1 = U.model_from_state(st).train()
    opt = FusedAdam(m1, lr=lr)
    tl = {}
    for _ in range(3):
        m1, tl = run_one_batch(m1, x.clone(), tl, model_kwargs={}, optimizer=opt, transform=None, training=True)
    m2 = U.model_from_state(st).train()
    tr = FusedTrainer(m2, lr=lr, use_graph=use_graph)
    curve = []
    for _ in range(3):
        curve.append(tr.step(x).tolist()[2])
    assert np.allclose(curve, tl["total_loss"], rtol=1e-5)
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    for k in sd1:
        if sd1[k].dtype.is_floating_point:
            assert float((sd1[k] - sd2[k]).abs().max()) <= 1e-6 + 1e-5 * float(sd1[k].abs().max()), k
        else:
            assert int(sd1[k]) == int(sd2[k]), k


@pytest.mark.parametrize("shape", [(37, 2, 128, 128), (5, 3, 16, 16)])
def test_device_augmentation_matches_reference_loop(shape):
    """dmb_augment_batch (one launch) == the reference's per-sample flip / rot90 loop (run_training.py:396-403) with
    the same np.random stream: bit-identical, and the RNG is left in the same state."""
    from dynamorph_b200.run_training import augment_batch
    x = torch.randn(*shape)
    np.random.seed(1234)
    ref = O.augment_batch(x.clone(), np.random)    # the reference's per-sample loop on the module-level stream
    after_ref = np.random.randint(1 << 30)
    np.random.seed(1234)
    xg = x.cuda()
    got = augment_batch(xg)
    after_got = np.random.randint(1 << 30)
    assert got.data_ptr() == xg.data_ptr()          # transformed in place, like the reference
    assert torch.equal(got.cpu(), ref)
    assert after_ref == after_got


@pytest.mark.parametrize("name", ["vqvae_default", "z16_masked"])
def test_eager_api_graph_replay_is_the_eager_step(name, U, monkeypatch):
    """`model(batch)` + `total_loss.backward()` replays captured CUDA graphs from the second call of a batch geometry on
    (dynamorph_b200/autograd.py).  Same kernels, same order: five run_one_batch steps end with bit-identical parameters,
    running statistics and loss curves with and without the graphs (DMB_EAGER_GRAPH=0)."""
    from dynamorph_b200.run_training import run_one_batch
    from dynamorph_b200.optim import FusedAdam
    g = Golden(name)
    x = g.t("x_train").cuda()
    results = []
    for flag in ("1", "0"):
        monkeypatch.setenv("DMB_EAGER_GRAPH", flag)
        m = U.model_from_state(g.state()).train()
        opt = FusedAdam(m, lr=1e-3)
        tl = {}
        for _ in range(5):
            m, tl = run_one_batch(m, x.clone(), tl, model_kwargs={"batch_mask": _mask(g)}, optimizer=opt, transform=None,
                                  training=True)
        cache = m.__dict__.get("_eager_cache")
        assert (cache is not None and cache.tr is not None) == (flag == "1")      # the graphs were (not) built
        if cache is not None:                          # copying / pickling a model never copies captured graphs
            import copy, pickle
            for c2 in (copy.deepcopy(cache), pickle.loads(pickle.dumps(cache))):
                assert c2.tr is None and not c2.seen
        results.append((tl, {k: v.detach().clone() for k, v in m.state_dict().items()}))
    (tl1, sd1), (tl0, sd0) = results
    assert tl1["total_loss"] == tl0["total_loss"] and tl1["perplexity"] == tl0["perplexity"]
    for k in sd0:
        assert torch.equal(sd1[k], sd0[k]), k


def test_eager_api_backward_after_another_forward_raises(U):
    """The activations of a forward live in one workspace per batch geometry: backward() after ANOTHER forward of the same
    geometry must fail loudly, graph replay or not (third call on: replayed graphs)."""
    g = Golden("vqvae_default")
    m = U.model_from_state(g.state()).train()
    x = g.t("x_train").cuda()
    for _ in range(2):
        m(x)[1]["total_loss"].backward()
        m.zero_grad()
    _, d1 = m(x)
    _, d2 = m(x)
    with pytest.raises(RuntimeError, match="reused by another call"):
        d1["total_loss"].backward()
    d2["total_loss"].backward()
    assert all(p.grad is not None for p in m.parameters() if p.requires_grad)
