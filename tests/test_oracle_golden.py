"""CPU: the oracle restatement against fixtures produced by the unmodified reference
modules (oracle/gen_golden.py).  Same torch build => expected bit-exact; tolerance kept
at 1e-6 relative so a different oneDNN dispatch on another host cannot flake."""
import numpy as np
import pytest
import torch

from conftest import Golden, rel_err
from oracle import vqvae_oracle as O

TOL = 1e-6


def test_eval_encode(golden_case):
    g = golden_case
    st = g.state()
    x = g.t("x_eval")
    with torch.no_grad():
        zb = O.encoder(x, st, O.EVAL)
        za, loss, ppl, idx = O.vq_forward(zb, st["vq.w.weight"], float(g["hp/commitment_cost"]))
    assert rel_err(zb, g.t("eval/z_before")) < TOL
    assert rel_err(za, g.t("eval/z_after")) < TOL
    assert np.array_equal(idx.numpy().astype(np.int32), g["eval/idx"])
    assert rel_err(loss, g.t("eval/vq_loss")) < TOL
    assert rel_err(ppl, g.t("eval/perplexity")) < TOL
    assert rel_err(O.vq_gather(idx, st["vq.w.weight"]), g.t("eval/decode_inputs")) == 0.0


def test_eval_forward_losses(golden_case):
    g = golden_case
    st = g.state()
    with torch.no_grad():
        dec, losses = O.forward(g.t("x_eval"), st, O.EVAL, float(g["hp/commitment_cost"]))
    assert rel_err(dec, g.t("eval/decoded")) < TOL
    for k in ("recon_loss", "commitment_loss", "total_loss", "perplexity"):
        assert abs(float(losses[k]) - float(g["eval/loss/" + k])) <= TOL * abs(float(g["eval/loss/" + k]))


def test_per_sample_encode(golden_case):
    g = golden_case
    st = g.state()
    with torch.no_grad():
        zb = O.encoder(g.t("x_eval"), st, O.PER_SAMPLE)
        za = O.vq_forward(zb, st["vq.w.weight"], 0.25)[0]
    assert rel_err(zb, g.t("per_sample/z_before")) < TOL
    assert rel_err(za, g.t("per_sample/z_after")) < TOL


def test_train_grads_and_running_stats(golden_case):
    g = golden_case
    st = g.state()
    mask = g.t("mask_train") if g.has("mask_train") else None
    dec, losses, grads, nr = O.loss_and_grads(g.t("x_train"), st, O.BATCH,
                                              commitment_cost=float(g["hp/commitment_cost"]),
                                              batch_mask=mask)
    assert rel_err(dec, g.t("train/decoded")) < TOL
    for k in ("recon_loss", "commitment_loss", "total_loss"):
        assert abs(float(losses[k]) - float(g["train/loss/" + k])) <= TOL * abs(float(g["train/loss/" + k]))
    gg = g.group("train/grad")
    assert set(gg) == set(grads)
    for k, v in gg.items():
        assert rel_err(grads[k], v) < 1e-5, k
    for k, v in g.group("train/after_fwd").items():
        assert rel_err(nr[k], v) < TOL, k


def test_adam_steps(golden_case):
    g = golden_case
    st = g.state()
    mask = g.t("mask_train") if g.has("mask_train") else None
    opt = {"m": {}, "v": {}}
    curve = []
    for s in range(int(g["train/steps"])):
        l = O.train_step(g.t("x_train"), st, opt, s + 1, float(g["train/lr"]), O.BATCH,
                         commitment_cost=float(g["hp/commitment_cost"]), batch_mask=mask)
        curve.append(float(l["total_loss"]))
    assert np.allclose(curve, g["train/curve/total_loss"], rtol=1e-5)
    noise = set(O.bias_feeds_train_bn(st))
    assert "enc.1.bias" in noise or "enc.0.bias" in noise
    for k, v in g.group("train/after_steps").items():
        if k.endswith("num_batches_tracked"):
            assert int(st[k]) == int(v)
        elif k in noise:   # zero-gradient parameters: bounded by Adam's max step, not comparable
            assert float((st[k] - v).abs().max()) <= 2.01 * len(curve) * float(g["train/lr"]), k
        else:
            assert rel_err(st[k], v) < 2e-5, k


def test_zscore_patch():
    z = np.load(Golden("zscore_patch").z.fid.name) if False else Golden("zscore_patch")
    out = O.zscore_patch(np.squeeze(z["raw"]))
    assert np.allclose(out, z["z"], rtol=1e-12, atol=1e-12)


def test_vq_edge_ties():
    g = Golden("vq_edge")
    cb, z = g.t("codebook"), g.t("z")
    zst, loss, ppl, idx = O.vq_forward(z, cb, 0.25)
    assert np.array_equal(idx.numpy().astype(np.int32), g["idx"])
    assert int(idx[0, 0, 0]) == 3          # exact tie 3 / 7 / 50 -> lowest index
    assert torch.equal(zst, g.t("z_st"))
    assert rel_err(loss, g.t("loss")) < TOL and rel_err(ppl, g.t("perplexity")) < TOL


def test_process_vae_arrays_layout():
    g = Golden("vqvae_default")
    st = g.state()
    raw = np.load(Golden("zscore_patch").z.fid.name)["raw"] if False else Golden("zscore_patch")["raw"]
    zb, za = O.process_vae_arrays(raw, st, O.PER_SAMPLE)
    assert zb.shape == (3, 16 * 16 * 16) and za.dtype == np.float32
    # NCHW flattening: index = c*256 + h*16 + w  (pipeline/patch_VAE.py:454)
    x = torch.from_numpy(O.zscore_patch(np.squeeze(raw))).float()
    with torch.no_grad():
        z0 = O.encoder(x[:1], st, O.BATCH)
    assert np.array_equal(zb[0], z0.numpy().reshape(-1))
