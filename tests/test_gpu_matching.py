"""-m gpu: the time-matching term of VQ_VAE.forward (reference: HiddenStateExtractor/vq_vae.py:324-332 for VQ_VAE,
vae.py:321-336 / :442-457 for VQ_VAE_z16 / VQ_VAE_z32) -- stand-alone op, eval forward and the fused training step
(losses and all parameter gradients) against the oracle on the same seeded inputs.  Tolerances: losses 1e-4 relative,
gradients 2e-4 of the tensor's max-abs (BASELINE.json north_star / SURVEY.md section 8d)."""
import pytest
import torch

from conftest import Golden
from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu

#        fixture          class name     variant   matching weight (the reference default 0.005 makes the term invisible)
CASES = [("vqvae_default", "VQ_VAE", "sum", 0.05),
         ("vqvae_default", "VQ_VAE_z16", "hinge", 2.0),
         ("z32_default", "VQ_VAE_z32", "hinge", 2.0)]


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


def _model(U, st, cls_name, weight):
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
    from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16, VQ_VAE_z32
    cls = {"VQ_VAE": VQ_VAE, "VQ_VAE_z16": VQ_VAE_z16, "VQ_VAE_z32": VQ_VAE_z32}[cls_name]
    return U.model_from_state(st, cls=cls, weight_matching=weight)


def _mat(B, seed):
    """Pair classes 0 / 1 / 2, asymmetric, every class present even for B = 2 (the z32 fixture's batch)."""
    i = torch.arange(B).view(-1, 1)
    j = torch.arange(B).view(1, -1)
    return ((i + 2 * j + 1 + seed) % 3).float()


def _oracle_kw(m, variant, weight):
    kw = dict(weight_matching=weight, tm_variant=variant, weight_recon=getattr(m, "weight_recon", 1.0),
              weight_commitment=getattr(m, "weight_commitment", 1.0))
    if variant == "hinge":
        kw.update(w_a=m.w_a, w_t=m.w_t, w_n=m.w_n, margin=m.margin)
    return kw


@pytest.mark.parametrize("variant", ["sum", "hinge"])
@pytest.mark.parametrize("B,L", [(5, 100), (33, 4096), (64, 1024)])
def test_time_matching_op(variant, B, L):
    """Stand-alone op: loss and dloss/dz against the oracle's direct (B, B, L) formulation."""
    from dynamorph_b200.matching import time_matching_loss

    class M:                       # the attributes dynamorph_b200.matching.descriptor reads
        weight_matching = 1.0
    m = M()
    if variant == "hinge":
        m.w_a, m.w_t, m.w_n, m.margin = 1.1, 0.1, -0.5, 0.5
    g = torch.Generator().manual_seed(B * 7 + L)
    z = torch.randn(B, L, generator=g) * 0.7
    mat = _mat(B, B + L)
    zr = z.clone().requires_grad_(True)
    kw = dict(w_a=1.1, w_t=0.1, w_n=-0.5, margin=0.5) if variant == "hinge" else {}
    ref = O.time_matching_loss(zr, mat, variant, **kw)
    ref.backward()
    zg = z.cuda().requires_grad_(True)
    out = time_matching_loss(m, zg, mat.cuda())
    out.backward()
    assert abs(float(out) - float(ref)) <= 1e-4 * abs(float(ref)) + 1e-7
    err = float((zg.grad.cpu() - zr.grad).abs().max() / zr.grad.abs().max().clamp(min=1e-30))
    assert err < 2e-4, err


@pytest.mark.parametrize("name,cls_name,variant,weight", CASES)
def test_eval_forward_with_time_matching(name, cls_name, variant, weight, U):
    g = Golden(name)
    st = g.state()
    m = _model(U, st, cls_name, weight).eval()
    x = g.t("x_train")
    mat = _mat(x.shape[0], 3)
    with torch.no_grad():
        _, d = m(x.cuda(), time_matching_mat=mat.cuda())
        _, ref = O.forward(x, st, O.EVAL, time_matching_mat=mat, **_oracle_kw(m, variant, weight))
    for k in ("recon_loss", "commitment_loss", "time_matching_loss", "total_loss"):
        assert abs(float(d[k]) - float(ref[k])) <= U.REL_TOL * abs(float(ref[k])), (k, float(d[k]), float(ref[k]))
    assert float(ref["time_matching_loss"]) != 0.0
    assert list(d) == list(m(x.cuda())[1])                 # same keys, same order as without the term


@pytest.mark.parametrize("name,cls_name,variant,weight", CASES)
def test_train_step_with_time_matching(name, cls_name, variant, weight, U):
    from test_gpu_train import _check_grads
    g = Golden(name)
    st = g.state()
    m = _model(U, st, cls_name, weight).train()
    x = g.t("x_train")
    mat = _mat(x.shape[0], 0)
    kw = _oracle_kw(m, variant, weight)
    _, ref, grads_ref, _ = O.loss_and_grads(x, st, O.BATCH, time_matching_mat=mat, **kw)
    _, base, grads_base, _ = O.loss_and_grads(x, st, O.BATCH, **{k: v for k, v in kw.items()
                                                                 if k in ("weight_recon", "weight_commitment")})
    m.zero_grad()
    _, d = m(x.cuda(), time_matching_mat=mat.cuda())
    d["total_loss"].backward()
    for k in ("recon_loss", "commitment_loss", "time_matching_loss", "total_loss"):
        assert abs(float(d[k]) - float(ref[k])) <= U.REL_TOL * abs(float(ref[k])), (k, float(d[k]), float(ref[k]))
    # the term must matter in this test: it changes the encoder gradients visibly
    key = "enc.0.weight"
    assert float((grads_ref[key] - grads_base[key]).abs().max()) > 1e-3 * float(grads_base[key].abs().max())
    env, ties = O.relu_gate_envelopes(x, st, O.BATCH, time_matching_mat=mat, **kw)
    worst = _check_grads(m, grads_ref, set(O.bias_feeds_train_bn(st)), env, f"{name}/{cls_name}")
    print(f"{name}/{cls_name}: relu near-ties {ties}, worst grad err {worst:.2f} x strict")
    with pytest.raises(AssertionError):
        m(x.cuda(), time_matching_mat=mat[:-1].cuda())     # vq_vae.py:329 `assert sim_mat.shape == time_matching_mat.shape`


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_trainer_with_time_matching(use_graph, U):
    """FusedTrainer.step(x, time_matching_mat=...) (C-ABI calls on flat buffers, optionally one CUDA graph) walks the
    same loss curve as run_one_batch + FusedAdam on the autograd path."""
    import numpy as np
    from dynamorph_b200.optim import FusedAdam
    from dynamorph_b200.run_training import run_one_batch
    from dynamorph_b200.trainer import FusedTrainer
    g = Golden("vqvae_default")
    st = g.state()
    x = g.t("x_train").cuda()
    mat = _mat(x.shape[0], 1).cuda()
    m1 = _model(U, st, "VQ_VAE_z16", 2.0).train()
    opt = FusedAdam(m1, lr=1e-3)
    tl = {}
    for _ in range(3):
        m1, tl = run_one_batch(m1, x.clone(), tl, model_kwargs={"time_matching_mat": mat}, optimizer=opt,
                               transform=None, training=True)
    m2 = _model(U, st, "VQ_VAE_z16", 2.0).train()
    tr = FusedTrainer(m2, lr=1e-3, use_graph=use_graph)
    curve, tms = [], []
    for _ in range(3):
        l = tr.step(x, time_matching_mat=mat).tolist()
        curve.append(l[2]); tms.append(l[4])
    assert np.allclose(curve, tl["total_loss"], rtol=1e-5)
    assert np.allclose(tms, tl["time_matching_loss"], rtol=1e-5) and tms[0] != 0.0
