"""Helpers for the -m gpu parity tests."""
import numpy as np
import torch

from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16, VQ_VAE_z32
from oracle import vqvae_oracle as O

REL_TOL = 1e-4          # BASELINE.json north_star: latents / reconstructions / losses, fp32
NEAR_TIE = 1e-6         # relative best/second-best gap below which an index flip is a documented near-tie


def model_from_state(state, cls=None, **kw):
    arch = O.arch_of(state)
    K, D = state["vq.w.weight"].shape
    n_res = O.num_residual_layers(state, "enc.12" if arch == "z16" else "enc.5")
    rh = state[("enc.12" if arch == "z16" else "enc.5") + ".layers.0.1.weight"].shape[0]
    if cls is None:
        cls = VQ_VAE if arch == "z16" else VQ_VAE_z32
    m = cls(num_inputs=state["channel_var"].numel(), num_hiddens=D, num_residual_hiddens=rh,
            num_residual_layers=n_res, num_embeddings=K, **kw)
    m.load_state_dict(state)
    return m.to("cuda")


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def check_indices(idx_gpu, z_ref, codebook, idx_ref, what=""):
    """Bit-exact except positions whose REFERENCE best/second-best gap is < NEAR_TIE relative."""
    idx_gpu = torch.as_tensor(idx_gpu).cpu().long()
    idx_ref = torch.as_tensor(idx_ref).cpu().long()
    bad = idx_gpu != idx_ref
    n_bad = int(bad.sum())
    if n_bad == 0:
        return 0
    gap = O.best_second_gap(z_ref, codebook)
    worst = float(gap[bad].max())
    assert worst < NEAR_TIE, f"{what}: {n_bad} index mismatches, largest reference gap {worst:.3e} >= {NEAR_TIE}"
    return n_bad
