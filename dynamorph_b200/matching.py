"""Time-matching loss of VQ_VAE.forward (reference: HiddenStateExtractor/vq_vae.py:324-332 for VQ_VAE;
vae.py:321-336 / :442-457 for VQ_VAE_z16 / VQ_VAE_z32) as C-ABI calls (csrc/matching.cu).  The reference's
(B, B, L) broadcast is never formed."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import DmbTimeMatching, call, ptr
from .engine import _require_cuda, _stream


def descriptor(model, mat: torch.Tensor):
    """-> (struct dmb_time_matching, the float32 device copy of `mat` that it points to)."""
    mat = _require_cuda(mat.detach(), "time_matching_mat").float().contiguous()
    variant = 1 if hasattr(model, "w_a") else 0          # VQ_VAE_z16 / z32 carry w_a, w_t, w_n, margin
    d = DmbTimeMatching(mat.data_ptr(), variant, float(getattr(model, "w_a", 0.)), float(getattr(model, "w_t", 0.)),
                        float(getattr(model, "w_n", 0.)), float(getattr(model, "margin", 0.)),
                        float(model.weight_matching))
    return d, mat


class TimeMatchingFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, mat, model):
        zf = _require_cuda(z.detach(), "latents").reshape(z.shape[0], -1).contiguous()
        B, L = zf.shape
        if tuple(mat.shape) != (B, B):
            raise AssertionError("sim_mat.shape == time_matching_mat.shape")      # vq_vae.py:329
        d, matf = descriptor(model, mat)
        n = C.c_size_t()
        call("dmb_time_matching_scratch_floats", B, L, C.byref(n))
        scratch = torch.empty(n.value, dtype=torch.float32, device=zf.device)
        out = torch.empty(1, dtype=torch.float32, device=zf.device)
        call("dmb_time_matching_forward", ptr(zf), B, L, C.byref(d), ptr(scratch), ptr(out), _stream())
        ctx.save_for_backward(zf, scratch)
        ctx.shape = z.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        zf, scratch = ctx.saved_tensors
        B, L = zf.shape
        gz = torch.empty_like(zf)
        call("dmb_time_matching_backward", ptr(zf), B, L, ptr(scratch), 1.0, ptr(gz), 0, _stream())
        return (gz * g).view(ctx.shape), None, None


def time_matching_loss(model, z: torch.Tensor, mat: torch.Tensor) -> torch.Tensor:
    return TimeMatchingFunction.apply(z, mat, model)
