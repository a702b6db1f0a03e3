"""Synthetic inputs and non-degenerate synthetic weights for benchmarks / smoke runs (SURVEY.md
section 8d recipe), produced with the CUDA path itself -- no reference or oracle involved."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def synthetic_patches(n: int, seed: int, device, channels: int = 2, size: int = 128) -> torch.Tensor:
    """zscore_patch-like data: N(0,1) noise, 3x3 box low-pass, re-standardised per patch/channel."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, channels, size, size, generator=g, device=device)
    x = F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), 3, stride=1)
    x = (x - x.mean((2, 3), keepdim=True)) / x.std((2, 3), keepdim=True, unbiased=False)
    return x.contiguous()


@torch.no_grad()
def calibrate(model, calib: torch.Tensor, seed: int = 0, jitter: float = 0.05):
    """Default torch init collapses the codebook (hard part 7): perturb the BN
    affine, set running statistics from one train-mode pass (momentum 1) and draw the codebook from
    encoder outputs so that many codes are in use."""
    g = torch.Generator().manual_seed(seed + 1000)
    dev = calib.device
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.copy_((torch.rand(m.num_features, generator=g) + 0.5).to(dev))
            m.bias.copy_((torch.randn(m.num_features, generator=g) * 0.1).to(dev))
    was_training = model.training
    model.train()
    model._bn_momentum = 1.0
    try:
        zb = model.enc(calib)
        if hasattr(model, "_arch") and model._arch == 1:
            model.dec(zb)
    finally:
        model._bn_momentum = 0.1
    model.eval()
    z = model.enc(calib)
    K, D = model.vq.w.weight.shape
    vecs = z.permute(0, 2, 3, 1).reshape(-1, D)
    pick = torch.randperm(vecs.shape[0], generator=g)[:K].to(dev)
    model.vq.w.weight.copy_(vecs[pick] + jitter * torch.randn(K, D, generator=g).to(dev))
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.num_batches_tracked.zero_()
    model.train(was_training)
    return model
