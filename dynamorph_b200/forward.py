"""VQ_VAE.forward (reference: HiddenStateExtractor/vq_vae.py:300-338, vae.py:297-346, :417-466)
expressed as C-ABI calls."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import call, ptr
from .engine import _require_cuda, _stream, vq_forward


def recon_loss(model, decoded, inputs, batch_mask):
    """mean( mse(decoded*mask, inputs*mask, 'none') / channel_var )  (vq_vae.py:320-322)."""
    B, Cc, H, W = inputs.shape
    acc = torch.zeros(1, dtype=torch.float64, device=inputs.device)
    mc = 0
    if batch_mask is not None:
        batch_mask = _require_cuda(batch_mask, "batch_mask")
        mc = batch_mask.shape[1]
        if batch_mask.shape[0] != B or batch_mask.shape[2:] != inputs.shape[2:] or mc not in (1, Cc):
            batch_mask = batch_mask.expand(B, Cc, H, W).contiguous()
            mc = Cc
    cv = model.channel_var.data.reshape(-1).contiguous()
    call("dmb_recon_loss", ptr(decoded), ptr(inputs), ptr(batch_mask), mc, ptr(cv), B, Cc, H * W, ptr(acc), _stream())
    return (acc[0] / float(inputs.numel())).float()


def model_forward(model, inputs, time_matching_mat=None, batch_mask=None):
    inputs = _require_cuda(inputs, "inputs")
    if torch.is_grad_enabled() and model.training and any(p.requires_grad for p in model.parameters()):
        from .autograd import train_forward
        return train_forward(model, inputs, time_matching_mat, batch_mask)
    eng = model._engine
    z_before = eng.encoder_forward(inputs)
    z_after, c_loss, perplexity = vq_forward(z_before, eng.codebook(), model.commitment_cost)
    decoded = eng.decoder_forward(z_after)
    if eng.module.training:
        pass
    rl = recon_loss(model, decoded, inputs, batch_mask)
    if model._arch == _lib.ARCH_Z32:
        total = rl + c_loss
    else:
        total = model.weight_recon * rl + model.weight_commitment * c_loss
    tm = 0.
    if time_matching_mat is not None:
        from .matching import time_matching_loss
        src = z_after if model._arch == _lib.ARCH_Z32 else z_before
        tm = time_matching_loss(model, src, time_matching_mat)
        total = total + model.weight_matching * tm
    out = {'recon_loss': rl, 'commitment_loss': c_loss, 'time_matching_loss': tm}
    if getattr(model, "_total_last", False):
        out['perplexity'] = perplexity
        out['total_loss'] = total
    else:
        out['total_loss'] = total
        out['perplexity'] = perplexity
    return decoded, out
