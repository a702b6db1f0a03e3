"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

The reference holds no golden vectors of its own, so these
fixtures -- outputs of ``HiddenStateExtractor.vq_vae.VQ_VAE``,
``HiddenStateExtractor.vae.VQ_VAE_z16 / VQ_VAE_z32``, ``torch.optim.Adam`` driven
as ``run_training.run_one_batch`` drives it, and ``pipeline.train_utils.zscore_patch``
-- are the pins both the oracle and the CUDA path are tested against.
Test infrastructure; never imported by the product.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DYNAMORPH_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import vqvae_oracle as O  # noqa: E402  (input / weight recipes only)


def _stub_missing():
    """pipeline.patch_VAE / run_training import plotting + hdf5 libraries that the
    image lacks; they are not on the arithmetic path."""
    for name in ("matplotlib", "matplotlib.pyplot", "h5py", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.use = lambda *a, **k: None
                sys.modules[name] = m
    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # pipeline/train_utils.py:32 predates NumPy 2


def _np(t):
    return t.detach().cpu().numpy().copy()


def make_case(name, ctor, ctor_kw, n_eval, n_train, steps, lr, seed, extra=None):
    torch.manual_seed(seed)
    model = ctor(device="cpu", **ctor_kw)
    calib = O.synthetic_patches(32, 100 + seed)
    state = O.calibrate_state({k: v.detach().clone() for k, v in model.state_dict().items()},
                              calib, seed=seed)
    if extra and extra.get("channel_var") is not None:   # vq_vae_supp.py:22 values
        state["channel_var"] = torch.tensor(extra["channel_var"], dtype=torch.float32).reshape(1, -1, 1, 1)
    out = {}
    for k, v in state.items():
        out["state/" + k] = _np(v)
    hp = dict(commitment_cost=model.commitment_cost)
    out["hp/commitment_cost"] = np.float64(model.commitment_cost)

    x = O.synthetic_patches(n_eval, 200 + seed)
    out["x_eval"] = _np(x).astype(np.float16)
    assert np.array_equal(out["x_eval"].astype(np.float32), _np(x))

    # ---- eval-mode encode (enc -> vq), batched
    model.load_state_dict(state)
    model.eval()
    with torch.no_grad():
        zb = model.enc(x)
        za, vql, ppl = model.vq(zb)
        idx = model.vq.encode_inputs(zb)
        dq = model.vq.decode_inputs(idx)
    out["eval/z_before"] = _np(zb)
    out["eval/z_after"] = _np(za)
    out["eval/idx"] = _np(idx).astype(np.int32)
    out["eval/vq_loss"] = _np(vql)
    out["eval/perplexity"] = _np(ppl)
    out["eval/decode_inputs"] = _np(dq)
    with torch.no_grad():
        dec, losses = model(x)
    out["eval/decoded"] = _np(dec)
    for k, v in losses.items():
        out["eval/loss/" + k] = np.float64(float(v))

    # ---- as-written process_VAE inner loop: batch 1, train-mode BN (patch_VAE.py:445-452)
    model.load_state_dict(state)
    model.train()
    zbs, zas = [], []
    for i in range(n_eval):
        s = x[i:i + 1]
        z_b = model.enc(s)
        z_a, _, _ = model.vq(z_b)
        zbs.append(z_b.cpu().data.numpy())
        zas.append(z_a.cpu().data.numpy())
    out["per_sample/z_before"] = np.concatenate(zbs, 0)
    out["per_sample/z_after"] = np.concatenate(zas, 0)

    # ---- training: forward/backward/Adam as run_training.run_one_batch (:404-408) drives it
    _stub_missing()
    import run_training as RT
    xt = O.synthetic_patches(n_train, 300 + seed)
    out["x_train"] = _np(xt).astype(np.float16)
    mask = None
    kw = {}
    if extra and extra.get("mask"):
        g = torch.Generator().manual_seed(5)
        mask = (torch.rand(n_train, 1, 128, 128, generator=g) > 0.3).float() * 0.5 + 0.5
        out["mask_train"] = _np(mask).astype(np.float16)
        mask = torch.from_numpy(out["mask_train"].astype(np.float32))
    model.load_state_dict(state)
    model.train()
    model.zero_grad()
    dec, losses = model(xt, batch_mask=mask)
    losses["total_loss"].backward()
    out["train/decoded"] = _np(dec)
    for k, v in losses.items():
        out["train/loss/" + k] = np.float64(float(v))
    for k, p in model.named_parameters():
        if p.requires_grad:
            out["train/grad/" + k] = _np(p.grad)
    for k, v in model.state_dict().items():
        if "running" in k:
            out["train/after_fwd/" + k] = _np(v)

    model.load_state_dict(state)
    model.train()
    model.zero_grad()
    opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(.9, .999))
    tl = {}
    for _ in range(steps):
        model, tl = RT.run_one_batch(model, xt.clone(), tl, model_kwargs={"batch_mask": mask},
                                     optimizer=opt, transform=None, training=True)
    out["train/steps"] = np.int64(steps)
    out["train/lr"] = np.float64(lr)
    for k, v in tl.items():
        out["train/curve/" + k] = np.asarray(v, np.float64)
    for k, v in model.state_dict().items():
        out[f"train/after_steps/{k}"] = _np(v)

    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB",
          "perplexity(eval)=%.2f" % float(ppl))


def make_pipeline_case():
    """zscore_patch on raw float64 stacks with a singleton Z axis, as
    pipeline/patch_VAE.py:413-419 receives them."""
    _stub_missing()
    from pipeline.train_utils import zscore_patch
    rng = np.random.RandomState(11)
    raw = rng.rand(3, 2, 1, 128, 128) * np.array([40000.0, 300.0]).reshape(1, 2, 1, 1, 1) + 2000.0
    raw[1, 1] = 7.0  # constant channel: std 0 -> eps in the denominator
    z = zscore_patch(np.squeeze(raw))
    path = os.path.join(ROOT, "tests", "golden", "zscore_patch.npz")
    np.savez_compressed(path, raw=raw.astype(np.float64), z=z)
    print("zscore ->", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


def make_vq_edge_case():
    """VectorQuantizer on hand-made exact ties / near ties (first index wins)."""
    from HiddenStateExtractor.vq_vae import VectorQuantizer
    torch.manual_seed(3)
    D, K = 16, 64
    vq = VectorQuantizer(D, K, 0.25, device="cpu")
    with torch.no_grad():
        vq.w.weight[7] = vq.w.weight[3]            # duplicate rows: exact tie -> 3 wins
        vq.w.weight[50] = vq.w.weight[3]
        vq.w.weight[20] = vq.w.weight[21] * (1 + 2 ** -22)   # near tie
    z = torch.randn(2, D, 16, 16)
    with torch.no_grad():
        z[0, :, 0, 0] = vq.w.weight[3]
        z[0, :, 0, 1] = vq.w.weight[7] + 1e-3
        z[0, :, 0, 2] = vq.w.weight[21]
        z[0, :, 0, 3] = 0.5 * (vq.w.weight[1] + vq.w.weight[2])  # equidistant in exact arithmetic
        zst, loss, ppl = vq(z)
        idx = vq.encode_inputs(z)
    path = os.path.join(ROOT, "tests", "golden", "vq_edge.npz")
    np.savez_compressed(path, codebook=_np(vq.w.weight), z=_np(z), z_st=_np(zst), loss=_np(loss),
                        perplexity=_np(ppl), idx=_np(idx).astype(np.int32))
    print("vq_edge ->", path)


def main():
    from HiddenStateExtractor import vq_vae as R, vae as RV
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    make_case("vqvae_default", R.VQ_VAE, {}, n_eval=4, n_train=4, steps=3, lr=1e-3, seed=0)
    make_case("z16_masked", RV.VQ_VAE_z16, {}, n_eval=2, n_train=3, steps=2, lr=1e-3, seed=1,
              extra={"mask": True, "channel_var": [0.0475, 0.0394]})
    make_case("z32_default", RV.VQ_VAE_z32, {}, n_eval=2, n_train=2, steps=2, lr=1e-3, seed=2)
    make_case("vqvae_heavy", R.VQ_VAE, dict(num_hiddens=64, num_embeddings=512),
              n_eval=2, n_train=2, steps=1, lr=1e-3, seed=3)
    make_pipeline_case()
    make_vq_edge_case()


if __name__ == "__main__":
    main()
