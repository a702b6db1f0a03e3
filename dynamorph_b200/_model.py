"""Shared machinery of the drop-in module classes (see HiddenStateExtractor/vq_vae.py, vae.py).

The classes own the same `nn.Conv2d / nn.BatchNorm2d / nn.ConvTranspose2d / nn.Embedding`
containers, constructed in the reference's order, so `state_dict()` keys, default
initialisation and RNG consumption are identical to the reference
(/root/reference/HiddenStateExtractor/vq_vae.py:276-298, vae.py:273-294, :401-414) and a
`model.pt` written by either side loads in the other.  Their `forward`s never touch
`torch.nn.functional`: they call the engine, which calls the C ABI."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib, engine as _engine

CHANNEL_VAR = np.array([1., 1.])


class VectorQuantizer(nn.Module):
    """Drop-in for HiddenStateExtractor/vq_vae.py:25-116 (= vae.py:12-103)."""

    def __init__(self, embedding_dim=128, num_embeddings=128, commitment_cost=0.25, device='cuda:0'):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.commitment_cost = commitment_cost
        self.device = device
        self.w = nn.Embedding(num_embeddings, embedding_dim)

    def forward(self, inputs):
        """-> (inputs + (quantized - inputs), loss, perplexity); one fused kernel."""
        if torch.is_grad_enabled() and (inputs.requires_grad or self.w.weight.requires_grad):
            from .autograd import VQFunction
            return VQFunction.apply(inputs, self.w.weight, float(self.commitment_cost))
        return _engine.vq_forward(inputs, self.w.weight.data, self.commitment_cost)

    @property
    def embeddings(self):
        return self.w.weight

    def encode_inputs(self, inputs):
        """-> LongTensor (B, H, W) of nearest-code indices (first index on ties)."""
        return _engine.vq_indices(inputs.detach(), self.w.weight.data).long()

    def decode_inputs(self, encoding_indices):
        """indices (B, H, W) -> quantized encodings (B, D, H, W)."""
        return _engine.vq_gather(encoding_indices, self.w.weight.data)


class ResidualBlock(nn.Module):
    """Drop-in for HiddenStateExtractor/vq_vae.py:180-225: same `.layers` containers/keys."""

    def __init__(self, num_hiddens=128, num_residual_hiddens=512, num_residual_layers=2):
        super().__init__()
        self.num_hiddens = num_hiddens
        self.num_residual_layers = num_residual_layers
        self.num_residual_hiddens = num_residual_hiddens
        self.layers = []
        for _ in range(self.num_residual_layers):
            self.layers.append(nn.Sequential(
                nn.ReLU(),
                nn.Conv2d(self.num_hiddens, self.num_residual_hiddens, 3, padding=1),
                nn.BatchNorm2d(self.num_residual_hiddens),
                nn.ReLU(),
                nn.Conv2d(self.num_residual_hiddens, self.num_hiddens, 1),
                nn.BatchNorm2d(self.num_hiddens)))
        self.layers = nn.ModuleList(self.layers)

    def forward(self, x, bn_mode=None):
        """output = x; output = output + layers[i](output) for every layer (vq_vae.py:212-225) as ONE C-ABI call
        (dmb_residual_block_forward: the schedule `model.enc` / `model.dec` run for their block).  BatchNorm follows
        `self.training` (batch statistics + running-stat update, or running statistics); forward only -- gradients
        flow through the block as part of `model(...)`, the way the reference trains it."""
        import ctypes as C
        from ._lib import BN_BATCH, BN_EVAL, BN_MODES, call, ptr
        x = _engine._require_cuda(x.detach(), "ResidualBlock input")
        B, Cx, H, W = x.shape
        if Cx != self.num_hiddens:
            raise RuntimeError(f"expected {self.num_hiddens} channels, got {Cx}")
        y = torch.empty_like(x)
        if B == 0:
            return y
        mode = BN_MODES[bn_mode] if bn_mode is not None else (BN_BATCH if self.training else BN_EVAL)
        dev = x.device
        convs_bns = [(l[1], l[2], l[4], l[5]) for l in self.layers]
        flat = [t.detach().reshape(-1).to(dev, torch.float32)
                for ca, ba, cb, bb in convs_bns for m in (ca, ba, cb, bb) for t in (m.weight, m.bias)]
        stats = [t.detach().reshape(-1).to(dev, torch.float32)
                 for ca, ba, cb, bb in convs_bns for m in (ba, bb) for t in (m.running_mean, m.running_var)]
        params = torch.cat(flat) if flat else torch.empty(0, device=dev)
        bnbuf = torch.cat(stats) if stats else torch.empty(0, device=dev)
        n_p, n_b, nbytes = C.c_int64(), C.c_int64(), C.c_size_t()
        args = (self.num_hiddens, self.num_residual_hiddens, self.num_residual_layers)
        call("dmb_residual_block_sizes", *args, B, H, W, mode, C.byref(n_p), C.byref(n_b), C.byref(nbytes))
        if n_p.value != params.numel() or n_b.value != bnbuf.numel():
            raise RuntimeError(f"ResidualBlock parameter layout mismatch: python {params.numel()}/{bnbuf.numel()} vs "
                               f"library {n_p.value}/{n_b.value}")
        ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        upd = ptr(bnbuf) if mode == BN_BATCH else None
        call("dmb_residual_block_forward", *args, ptr(params), ptr(bnbuf), ptr(x), B, H, W, mode, ptr(y), upd, ptr(ws),
             nbytes.value, _engine._stream())
        if mode == BN_BATCH:                       # hand the updated running statistics back to the BatchNorm modules
            off = 0
            with torch.no_grad():
                for ca, ba, cb, bb in convs_bns:
                    for m in (ba, bb):
                        c = m.num_features
                        m.running_mean.copy_(bnbuf[off:off + c]); m.running_var.copy_(bnbuf[off + c:off + 2 * c])
                        m.num_batches_tracked += 1
                        off += 2 * c
        return y


class _Stage(nn.Sequential):
    """`model.enc` / `model.dec`: an nn.Sequential (for the state_dict keys) whose call runs the
    fused CUDA schedule instead of the children."""

    def _bind(self, owner, which):
        object.__setattr__(self, "_owner", owner)
        object.__setattr__(self, "_which", which)

    def forward(self, x, bn_mode=None):
        owner = self._owner
        eng = owner._engine
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad and owner.training and getattr(owner, "_stage_autograd", False):
            from .autograd import stage_apply
            return stage_apply(owner, self._which, x)
        if self._which == "enc":
            return eng.encoder_forward(x.detach(), bn_mode)
        return eng.decoder_forward(x.detach(), bn_mode)


class VQVAEBase(nn.Module):
    """Common body of VQ_VAE / VQ_VAE_z16 / VQ_VAE_z32."""

    _arch = _lib.ARCH_Z16

    def _finish_init(self):
        object.__setattr__(self, "_engine", _engine.Engine(self, self._arch))
        self.enc._bind(self, "enc")
        self.dec._bind(self, "dec")

    # -- conveniences named by the north star (absent from the reference class) ---------------
    def encode(self, inputs, bn_mode=None):
        """inputs (B,C,H,W) -> LongTensor code indices (B,h,w): enc -> vq.encode_inputs."""
        mode = bn_mode or ("eval" if not self.training else None)
        if mode is None:
            return self.vq.encode_inputs(self.enc(inputs))
        return self._engine.encode(inputs, mode)[2].long()

    def decode(self, encoding_indices, bn_mode=None):
        """code indices -> reconstruction: vq.decode_inputs -> dec."""
        return self.dec(self.vq.decode_inputs(encoding_indices), bn_mode)

    def encode_latents(self, inputs, bn_mode="eval"):
        """Batched process_VAE body: -> (z_before, z_after, idx int32)."""
        return self._engine.encode(inputs, bn_mode)

    def predict(self, inputs):
        """Prediction fn, same as forward pass (vq_vae.py:340-342)."""
        return self.forward(inputs)
