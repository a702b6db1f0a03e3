"""Drop-in for the PCA projection step of /root/reference/run_dim_reduction.py: `process_PCA` (:53-92) applies a
pre-fit sklearn PCA (`<weights_dir>/pca_model.pkl`) to the latent vectors `process_VAE` wrote and saves the top
components.  The projection `(X - mean_) @ components_.T` (sklearn PCA.transform; / sqrt(explained_variance_) when the
model whitens) runs on the GPU (csrc/pca.cu) in pinned, double-buffered chunks; fitting, UMAP and the plots of the
reference script are outside the hot path."""
from __future__ import annotations

import ctypes as C
import os
import pickle

import numpy as np
import torch

from ._lib import call, ptr


def pca_transform(pca, dats: np.ndarray, device="cuda:0", chunk: int = 65536) -> np.ndarray:
    """sklearn-compatible `pca.transform(dats)` -> (N, n_components) in the dtype of `pca.components_` (what sklearn
    returns), computed on `device` in float32: the 4096-long dot products are accumulated in fp32, so values agree
    with sklearn's float64 result to ~1e-6 relative, not bit for bit (tests/test_gpu_pca.py states the tolerance)."""
    dats = np.asarray(dats)
    if dats.ndim != 2 or dats.shape[1] != pca.components_.shape[1]:
        raise ValueError(f"X has {dats.shape[1] if dats.ndim == 2 else '?'} features, but PCA is expecting "
                         f"{pca.components_.shape[1]} features as input.")
    dev = torch.device(device)
    n, l = dats.shape
    k = pca.components_.shape[0]
    comp = torch.as_tensor(np.ascontiguousarray(pca.components_, dtype=np.float32), device=dev)
    mean = torch.as_tensor(np.ascontiguousarray(pca.mean_ if pca.mean_ is not None else np.zeros(l), dtype=np.float32),
                           device=dev)
    inv = None
    if getattr(pca, "whiten", False):
        inv = torch.as_tensor((1.0 / np.sqrt(pca.explained_variance_)).astype(np.float32), device=dev)
    out = np.empty((n, k), dtype=np.float32)
    copy, comp_s = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    bufs = [torch.empty(min(chunk, max(n, 1)), l, dtype=torch.float32, device=dev) for _ in range(2)]
    outs = [torch.empty(min(chunk, max(n, 1)), k, dtype=torch.float32, device=dev) for _ in range(2)]
    nf = C.c_int64()
    call("dmb_pca_transform_scratch_floats", min(chunk, max(n, 1)), l, C.byref(nf))
    scratch = torch.empty(nf.value, dtype=torch.float32, device=dev)      # tensor-core form: packed weights + (rows, 64) tile
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    pending = []
    for i, a in enumerate(range(0, n, chunk)):
        j = i & 1
        m = min(chunk, n - a)
        part = torch.from_numpy(np.ascontiguousarray(dats[a:a + m], dtype=np.float32))
        with torch.cuda.stream(copy):
            if i >= 2:
                copy.wait_event(ev_done[j])
            bufs[j][:m].copy_(part, non_blocking=True)
            ev_in[j].record(copy)
        comp_s.wait_event(ev_in[j])
        if i >= 2:                                   # outs[j] of chunk i-2 must be on the host before it is reused
            pa, pm, pj = pending.pop(0)
            out[pa:pa + pm] = outs[pj][:pm].cpu().numpy()
        call("dmb_pca_transform_tc", ptr(bufs[j]), m, l, ptr(mean), ptr(comp), k, ptr(inv), ptr(outs[j]), ptr(scratch),
             C.c_void_p(comp_s.cuda_stream))
        ev_done[j].record(comp_s)
        pending.append((a, m, j))
    for pa, pm, pj in pending:
        out[pa:pa + pm] = outs[pj][:pm].cpu().numpy()
    want = np.asarray(pca.components_).dtype
    return out if want == np.float32 else out.astype(want)


def process_PCA(input_dir, output_dir, weights_dir, prefix, suffix='_after', device="cuda:0"):
    """Loads `<input_dir>/<prefix>_latent_space_<suffix>.pkl`, applies the PCA of `<weights_dir>/pca_model.pkl`, writes
    `<output_dir>/<prefix>_latent_space_<suffix>_PCAed.pkl` (pickle protocol 4) -- file naming exactly as
    run_dim_reduction.py:83-84."""
    os.mkdir(output_dir) if not os.path.exists(output_dir) else None
    model_path = os.path.join(weights_dir, 'pca_model.pkl')
    try:
        with open(model_path, 'rb') as pretrained_model:
            pca = pickle.load(pretrained_model)
    except Exception as ex:
        print(ex)
        raise ValueError("Error in loading pre-saved PCA weights")
    input_fname = '{}_latent_space_{}.pkl'.format(prefix, suffix)
    output_fname = '{}_latent_space_{}_PCAed.pkl'.format(prefix, suffix)
    with open(os.path.join(input_dir, input_fname), 'rb') as latent:
        dats = pickle.load(latent)
    dats_ = pca_transform(pca, dats, device=device)
    output_file = os.path.join(output_dir, output_fname)
    print(f"\tSaving PCA-transformed latent space to {output_file}")
    with open(output_file, 'wb') as f:
        pickle.dump(dats_, f, protocol=4)
