"""Drop-in for the parts of /root/reference/pipeline/train_utils.py on the VQ-VAE path:
`zscore_patch` (:252-274, on the GPU) and `EarlyStopping` (:8-60)."""
from __future__ import annotations

import numpy as np
import torch

from .._lib import call, ptr

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
    import ctypes as C
    call("dmb_zscore_patch", ptr(x), _DTYPES[x.dtype], n * c, h * w, ptr(out),
         C.c_void_p(torch.cuda.current_stream().cuda_stream))
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
    """Early stops the training if validation loss doesn't improve after a given patience; saves
    `model.state_dict()` to `path` on every improvement (reference: train_utils.py:8-60)."""

    def __init__(self, patience=7, verbose=False, delta=0, path='checkpoint.pt', trace_func=print):
        self.patience = patience
        self.verbose = verbose
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.val_loss_min = np.inf
        self.delta = delta
        self.path = path
        self.trace_func = trace_func

    def __call__(self, val_loss, model):
        score = -val_loss
        if self.best_score is None:
            self.best_score = score
            self.save_checkpoint(val_loss, model)
        elif score < self.best_score + self.delta:
            self.counter += 1
            self.trace_func(f'EarlyStopping counter: {self.counter} out of {self.patience}')
            if self.patience is not None and self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.save_checkpoint(val_loss, model)
            self.counter = 0

    def save_checkpoint(self, val_loss, model):
        if self.verbose:
            self.trace_func(f'Validation loss decreased ({self.val_loss_min:.6f} --> {val_loss:.6f}).  Saving model ...')
        # clone: the parameters are views of one flat buffer; save them as independent tensors
        torch.save({k: v.detach().clone() for k, v in model.state_dict().items()}, self.path)
        self.val_loss_min = val_loss
