"""Drop-in for /root/reference/HiddenStateExtractor/vq_vae.py (VQ-VAE classes only).

`VQ_VAE`, `VectorQuantizer`, `ResidualBlock` keep the reference's constructor arguments,
attributes (`.enc`, `.vq`, `.dec`, `.w`, `.layers`), return values and state_dict keys; the
arithmetic runs in sm_100a CUDA kernels.  VAE / IWAE / AAE are out of scope."""
from __future__ import annotations

import torch as t
import torch.nn as nn

from .._model import CHANNEL_VAR, ResidualBlock, VectorQuantizer, VQVAEBase, _Stage
from .. import _lib

__all__ = ["VQ_VAE", "VectorQuantizer", "ResidualBlock", "CHANNEL_VAR"]


class VQ_VAE(VQVAEBase):
    """Vector-Quantized VAE (reference: vq_vae.py:228-342)."""

    _arch = _lib.ARCH_Z16

    def __init__(self,
                 num_inputs=2,
                 num_hiddens=16,
                 num_residual_hiddens=32,
                 num_residual_layers=2,
                 num_embeddings=64,
                 commitment_cost=0.25,
                 channel_var=CHANNEL_VAR,
                 weight_recon=1.,
                 weight_commitment=1.,
                 weight_matching=0.005,
                 device="cuda:0",
                 **kwargs):
        # plot scripts pass alpha=..., gpu=... (plot_scripts/plottings.py:400), which modern
        # nn.Module.__init__ rejects; they never influenced the computation.
        kwargs.pop("alpha", None)
        kwargs.pop("gpu", None)
        super().__init__(**kwargs)
        self.num_inputs = num_inputs
        self.num_hiddens = num_hiddens
        self.num_residual_layers = num_residual_layers
        self.num_residual_hiddens = num_residual_hiddens
        self.num_embeddings = num_embeddings
        self.commitment_cost = commitment_cost
        self.channel_var = nn.Parameter(
            t.from_numpy(channel_var).float().reshape((1, num_inputs, 1, 1)), requires_grad=False)
        self.weight_recon = weight_recon
        self.weight_commitment = weight_commitment
        self.weight_matching = weight_matching
        h = self.num_hiddens
        self.enc = _Stage(
            nn.Conv2d(self.num_inputs, h // 2, 1),
            nn.Conv2d(h // 2, h // 2, 4, stride=2, padding=1),
            nn.BatchNorm2d(h // 2),
            nn.ReLU(),
            nn.Conv2d(h // 2, h, 4, stride=2, padding=1),
            nn.BatchNorm2d(h),
            nn.ReLU(),
            nn.Conv2d(h, h, 4, stride=2, padding=1),
            nn.BatchNorm2d(h),
            nn.ReLU(),
            nn.Conv2d(h, h, 3, padding=1),
            nn.BatchNorm2d(h),
            ResidualBlock(h, self.num_residual_hiddens, self.num_residual_layers))
        self.vq = VectorQuantizer(h, self.num_embeddings, commitment_cost=self.commitment_cost, device=device)
        self.dec = _Stage(
            nn.ConvTranspose2d(h, h // 2, 4, stride=2, padding=1),
            nn.ReLU(),
            nn.ConvTranspose2d(h // 2, h // 4, 4, stride=2, padding=1),
            nn.ReLU(),
            nn.ConvTranspose2d(h // 4, h // 4, 4, stride=2, padding=1),
            nn.ReLU(),
            nn.Conv2d(h // 4, self.num_inputs, 1))
        self._tm_variant = "sum"
        self._finish_init()

    def forward(self, inputs, time_matching_mat=None, batch_mask=None):
        """-> (decoded, {'recon_loss','commitment_loss','time_matching_loss','total_loss','perplexity'})"""
        from ..forward import model_forward
        return model_forward(self, inputs, time_matching_mat, batch_mask)
