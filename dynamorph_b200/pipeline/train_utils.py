"""Drop-in for the parts of /root/reference/pipeline/train_utils.py on the VQ-VAE path:
`zscore_patch` (:252-274, on the GPU) and `EarlyStopping` (:8-60)."""
from __future__ import annotations

import numpy as np
import torch

from .._lib import STREAM, call, ptr

_DTYPES = {torch.float32: 0, torch.float64: 1, torch.uint16: 2}


def zscore_patch_device(imgs: torch.Tensor) -> torch.Tensor:
    """(N, C, H, W) float32 / float64 / uint16 CUDA tensor -> float32, per patch & channel
    (x - mean) / (std + eps) with the statistics in float64 (one CTA per plane)."""
    if not imgs.is_cuda:
        raise RuntimeError("dynamorph_b200: zscore_patch_device needs a CUDA tensor")
    if imgs.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported dtype {imgs.dtype}; use float32, float64 or uint16")
    x = imgs.contiguous()
    n, c, h, w = x.shape
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=x.device)
    call("dmb_zscore_patch", ptr(x), _DTYPES[x.dtype], n * c, h * w, ptr(out), STREAM)
    return out


def zscore_patch(imgs, device="cuda:0", chunk=4096):
    """numpy (N, C, H, W) -> numpy float32 of the same shape, computed on the GPU in chunks.
    (The reference returns float64 and its caller immediately casts to float32, patch_VAE.py:418-419.)"""
    imgs = np.asarray(imgs)
    if imgs.ndim != 4:
        raise ValueError("zscore_patch expects (N, C, H, W)")
    out = np.empty(imgs.shape, dtype=np.float32)
    for a in range(0, imgs.shape[0], chunk):
        part = np.ascontiguousarray(imgs[a:a + chunk])
        if part.dtype not in (np.float32, np.float64, np.uint16):
            part = part.astype(np.float64)
        d = torch.from_numpy(part).to(device)
        out[a:a + chunk] = zscore_patch_device(d).cpu().numpy()
    return out


class EarlyStopping:
    """Stop training once the validation loss has failed to improve by `delta` for `patience` consecutive calls, and
    checkpoint `model.state_dict()` to `path` at every improvement -- the contract of the reference's class
    (train_utils.py:8-60: same constructor, `counter` / `best_score` / `early_stop` / `val_loss_min` attributes,
    `__call__(val_loss, model)`, `save_checkpoint`).  `patience=None` never stops (run_training.train allows it)."""

    def __init__(self, patience=7, verbose=False, delta=0, path='checkpoint.pt', trace_func=print):
        self.patience, self.verbose, self.delta = patience, verbose, delta
        self.path, self.trace_func = path, trace_func
        self.counter = 0
        self.early_stop = False
        self.val_loss_min = np.inf          # loss of the last checkpoint written
        self._lowest = None                 # lowest validation loss accepted as an improvement

    @property
    def best_score(self):
        """The reference tracks the negated loss."""
        return None if self._lowest is None else -self._lowest

    def __call__(self, val_loss, model):
        stalled = self._lowest is not None and val_loss > self._lowest - self.delta
        if not stalled:                     # first call, an improvement of at least delta (or a NaN, as in the reference)
            self._lowest = val_loss
            self.counter = 0
            self.save_checkpoint(val_loss, model)
            return
        self.counter += 1
        self.trace_func(f'EarlyStopping counter: {self.counter} out of {self.patience}')
        self.early_stop = self.patience is not None and self.counter >= self.patience

    def save_checkpoint(self, val_loss, model):
        if self.verbose:
            self.trace_func(f'Validation loss decreased ({self.val_loss_min:.6f} --> {val_loss:.6f}).  Saving model ...')
        # the parameters are views of one flat device buffer: save independent tensors, written atomically
        state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        tmp = f"{self.path}.tmp"
        torch.save(state, tmp)
        import os
        os.replace(tmp, self.path)
        self.val_loss_min = val_loss
