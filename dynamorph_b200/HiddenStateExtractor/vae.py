"""Drop-in for /root/reference/HiddenStateExtractor/vae.py (VQ-VAE classes only):
`VQ_VAE_z16` (vae.py:216-346) and `VQ_VAE_z32` (vae.py:348-474) -- the classes
pipeline.patch_VAE.process_VAE and run_training actually instantiate."""
from __future__ import annotations

import numpy as np
import torch as t
import torch.nn as nn

from .._model import CHANNEL_VAR, ResidualBlock, VectorQuantizer, VQVAEBase, _Stage
from .. import _lib

__all__ = ["VQ_VAE_z16", "VQ_VAE_z32", "VectorQuantizer", "ResidualBlock", "CHANNEL_VAR"]


class VQ_VAE_z16(VQVAEBase):
    """Reduced VQ-VAE with a 16 x 16 x num_hiddens latent (reference: vae.py:216-346)."""

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
                 w_a=1.1,
                 w_t=0.1,
                 w_n=-0.5,
                 margin=0.5,
                 **kwargs):
        kwargs.pop("alpha", None)
        kwargs.pop("gpu", None)     # pipeline/patch_VAE.py:431 passes gpu=True
        super().__init__(**kwargs)
        self.num_inputs = num_inputs
        self.num_hiddens = num_hiddens
        self.num_residual_layers = num_residual_layers
        self.num_residual_hiddens = num_residual_hiddens
        self.num_embeddings = num_embeddings
        self.commitment_cost = commitment_cost
        self.channel_var = nn.Parameter(
            t.from_numpy(np.asarray(channel_var, dtype=np.float64)).float().reshape((1, num_inputs, 1, 1)),
            requires_grad=False)
        self.weight_recon = weight_recon
        self.weight_commitment = weight_commitment
        self.weight_matching = weight_matching
        self.w_a, self.w_t, self.w_n, self.margin = w_a, w_t, w_n, margin
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
        self._tm_variant = "hinge"
        self._total_last = True      # key order of the returned dict (vae.py:337-342)
        self._finish_init()

    def forward(self, inputs, time_matching_mat=None, batch_mask=None):
        from ..forward import model_forward
        return model_forward(self, inputs, time_matching_mat, batch_mask)


class VQ_VAE_z32(VQVAEBase):
    """VQ-VAE with a 32 x 32 x num_hiddens latent (reference: vae.py:348-474)."""

    _arch = _lib.ARCH_Z32

    def __init__(self,
                 num_inputs=2,
                 num_hiddens=16,
                 num_residual_hiddens=32,
                 num_residual_layers=2,
                 num_embeddings=64,
                 commitment_cost=0.25,
                 channel_var=np.ones(2),
                 weight_matching=0.005,
                 w_a=1.1,
                 w_t=0.1,
                 w_n=-0.5,
                 margin=0.5,
                 extra_loss=None,
                 device="cuda:0",
                 **kwargs):
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
            t.from_numpy(np.asarray(channel_var, dtype=np.float64)).float().reshape((1, num_inputs, 1, 1)),
            requires_grad=False)
        self.weight_matching = weight_matching
        self.w_a, self.w_t, self.w_n, self.margin = w_a, w_t, w_n, margin
        h = self.num_hiddens
        self.enc = _Stage(
            nn.Conv2d(self.num_inputs, h // 2, 4, stride=2, padding=1),
            nn.BatchNorm2d(h // 2),
            nn.ReLU(),
            nn.Conv2d(h // 2, h, 4, stride=2, padding=1),
            nn.BatchNorm2d(h),
            ResidualBlock(h, self.num_residual_hiddens, self.num_residual_layers))
        self.vq = VectorQuantizer(h, self.num_embeddings, commitment_cost=self.commitment_cost, device=device)
        self.dec = _Stage(
            ResidualBlock(h, self.num_residual_hiddens, self.num_residual_layers),
            nn.ConvTranspose2d(h, h // 2, 4, stride=2, padding=1),
            nn.BatchNorm2d(h // 2),
            nn.ReLU(),
            nn.ConvTranspose2d(h // 2, self.num_inputs, 4, stride=2, padding=1))
        if extra_loss is not None:
            raise NotImplementedError("extra_loss (triplet miners, vae.py:463-469) is outside the VQ-VAE hot path")
        self.extra_loss = None
        self._tm_variant = "hinge"
        self._total_last = True
        self._finish_init()

    def forward(self, inputs, labels=None, time_matching_mat=None, batch_mask=None):
        from ..forward import model_forward
        return model_forward(self, inputs, time_matching_mat, batch_mask)
