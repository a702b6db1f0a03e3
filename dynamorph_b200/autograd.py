"""torch.autograd glue: the forward/backward of each Function is a C-ABI call; autograd only
routes gradients between them."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import call, ptr
from .engine import _require_cuda, _stream


class VQFunction(torch.autograd.Function):
    """VectorQuantizer.forward with the reference's gradients (straight-through + both MSE terms)."""

    @staticmethod
    def forward(ctx, inputs, codebook, commitment_cost):
        z = _require_cuda(inputs.detach(), "VectorQuantizer input")
        cb = _require_cuda(codebook.detach(), "codebook")
        B, D, H, W = z.shape
        K = cb.shape[0]
        z_st = torch.empty_like(z)
        idx = torch.empty(B, H, W, dtype=torch.int32, device=z.device)
        stats = torch.zeros(2 + K, dtype=torch.float64, device=z.device)
        call("dmb_vq_forward", ptr(z), ptr(cb), B, D, H * W, K, ptr(z_st), ptr(idx), ptr(stats), _stream())
        out2 = torch.empty(2, dtype=torch.float32, device=z.device)
        call("dmb_vq_finalize", ptr(stats), D, K, float(commitment_cost), ptr(out2), _stream())
        ctx.save_for_backward(z, cb, idx)
        ctx.beta = float(commitment_cost)
        loss, ppl = out2[0], out2[1]
        ctx.mark_non_differentiable(ppl)
        return z_st, loss, ppl

    @staticmethod
    def backward(ctx, g_zst, g_loss, _g_ppl):
        z, cb, idx = ctx.saved_tensors
        B, D, H, W = z.shape
        K = cb.shape[0]
        gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        gcb = torch.empty_like(cb) if ctx.needs_input_grad[1] else None
        gzst = g_zst.contiguous() if g_zst is not None else None
        gl = g_loss.contiguous().float() if g_loss is not None else None
        scale = 1.0 if gl is not None else 0.0
        call("dmb_vq_backward", ptr(z), ptr(cb), ptr(idx), ptr(gzst), ptr(gl), scale, ctx.beta,
             B, D, H * W, K, ptr(gz), ptr(gcb), _stream())
        return gz, gcb, None
