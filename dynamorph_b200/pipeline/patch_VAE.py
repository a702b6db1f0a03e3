"""Drop-in for `process_VAE` of /root/reference/pipeline/patch_VAE.py:343-509 (VAE branch).

Same inputs (`<raw>/<well>_file_paths.pkl`, `<well>_static_patches.pkl`, the `latent_encoding` config
section) and the same outputs (`<raw>/<basename(weights)>/<well>_latent_space.pkl` and
`_latent_space_after.pkl`: float32 (N, D*h*w), NCHW-flattened, pickle protocol 4), but the per-patch
python loop (batch 1, two blocking D2H copies per patch, autograd on) is replaced by the pipelined bulk
encoder: z-score on the GPU, chunked H2D / encode / D2H on three streams.

BatchNorm semantics: the reference never calls `model.eval()`, so as written it normalises every patch
with that patch's own statistics (train-mode BN, batch 1).  `bn_mode="per_sample"` (default) reproduces
exactly that; `bn_mode="eval"` uses the trained running statistics."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from ..HiddenStateExtractor import vae
from ..bulk import BulkEncoder
from .train_utils import zscore_patch_device


def encode_patches(model, dataset: np.ndarray, device, bn_mode: str = "per_sample", chunk: int = 4096,
                   zscore: bool = True):
    """(N, C, H, W) raw patches on the host -> (z_before, z_after) float32 (N, D*h*w) on the host."""
    n = dataset.shape[0]
    enc = BulkEncoder(model, chunk=chunk, bn_mode=bn_mode, device=device, outputs=("z_before", "z_after"),
                      zscore=zscore)
    out = enc.allocate_outputs(n, dataset.shape[1], dataset.shape[2], dataset.shape[3], pin=True)
    if not zscore:
        enc.encode(torch.from_numpy(np.ascontiguousarray(dataset, dtype=np.float32)), out)
    else:
        # z-score needs the raw (float64 / uint16) values: they are shipped as they are and normalised on the device
        # in front of the encoder, chunk i+1 in flight while chunk i is encoded (no float32 host copy)
        part = np.ascontiguousarray(dataset)
        if part.dtype not in (np.float32, np.float64, np.uint16):
            part = part.astype(np.float64)
        enc.encode(torch.from_numpy(part), out)
    torch.cuda.synchronize(device)
    return out["z_before"].numpy(), out["z_after"].numpy()


def encode_patches_to_shards(model, dataset: np.ndarray, device, out_dir: str, well: str, bn_mode: str = "per_sample",
                             chunk: int = 4096, rows_per_shard: int = 65536, block_rows: int = 65536, rank: int = 0,
                             world: int = 1, first_row: int = 0):
    """Streaming form of `encode_patches` for runs whose (N, D*h*w) float32 outputs do not fit in host memory: the raw
    patches are walked in blocks, each block's latents go straight from the pinned D2H buffers into the sharded store
    (latent_shards.py: `<well>_latent_space.*.npy`, `<well>_latent_space_after.*.npy` + per-rank manifests; call
    `merge_manifests` once every rank is done).  Rows keep the reference's layout (patch_VAE.py:454-461)."""
    from ..latent_shards import ShardedLatentWriter
    enc = BulkEncoder(model, chunk=chunk, bn_mode=bn_mode, device=device, outputs=("z_before", "z_after"), zscore=True)
    part = dataset if dataset.dtype in (np.float32, np.float64, np.uint16) else dataset.astype(np.float64)
    writers = {}

    def source():
        for a in range(0, part.shape[0], block_rows):
            yield torch.from_numpy(np.ascontiguousarray(part[a:a + block_rows]))

    def sink(_first, out):
        for kind, key in (("latent_space", "z_before"), ("latent_space_after", "z_after")):
            if kind not in writers:
                writers[kind] = ShardedLatentWriter(out_dir, well, kind, out[key].shape[1], rank=rank, world=world,
                                                    first_row=first_row, rows_per_shard=rows_per_shard)
            writers[kind].append(out[key].numpy())

    n = enc.encode_stream(source(), sink)
    for w in writers.values():
        w.close()
    return n


def _contrast_u8(img, tol=1):
    """SingleCellPatch/extract_patches.py:314-334 `im_adjust`: stretch the [tol, 100-tol] percentile range to 8 bit."""
    lo, hi = np.percentile(img, [tol, 100 - tol])
    return np.clip((img.astype(np.float32) - lo) / (hi - lo) * 255., 0, 255).astype(np.uint8)


def save_reconstructions(model, dataset, output_dir, device, count=20):
    """patch_VAE.py:464-489: `count` randomly chosen patches (np.random.seed(0)) and their reconstructions as
    `recon_<i>.jpg`, a 2 x 2 panel phase | phase_recon / retardance | retardance_recon, contrast-adjusted like the
    reference.  The panels are written with OpenCV (the reference renders them through matplotlib, absent here);
    the model call is the reference's `model(sample)` in train mode."""
    import cv2
    np.random.seed(0)
    picks = np.random.randint(0, len(dataset), (count,))
    was_training = model.training
    model.train()
    with torch.no_grad():
        for i in picks:
            sample = zscore_patch_device(torch.from_numpy(np.ascontiguousarray(dataset[i:i + 1])).to(device))
            output = model(sample)[0]
            src, rec = sample[0].cpu().numpy(), output[0].cpu().numpy()
            rows = []
            for c, name in enumerate(('phase', 'retard')[:src.shape[0]]):
                tiles = []
                for img, label in ((src[c], name), (rec[c], name + '_recon')):
                    tile = cv2.resize(_contrast_u8(img), None, fx=3, fy=3, interpolation=cv2.INTER_NEAREST)
                    tile = cv2.copyMakeBorder(tile, 24, 4, 4, 4, cv2.BORDER_CONSTANT, value=255)
                    cv2.putText(tile, label, (6, 17), cv2.FONT_HERSHEY_SIMPLEX, 0.5, 0, 1, cv2.LINE_AA)
                    tiles.append(tile)
                rows.append(np.hstack(tiles))
            cv2.imwrite(os.path.join(output_dir, 'recon_%d.jpg' % i), np.vstack(rows))
    model.train(was_training)


def process_VAE(raw_folder: str, supp_folder: str, sites: list, config_, gpu: int = 0, bn_mode: str = "per_sample",
                shard_rows: int = 0, **kwargs):
    """Wrapper method for VAE encoding: loads the prepared dataset of one well and encodes its static
    patches with the trained VQ-VAE; writes the two latent-space pickles (reference: patch_VAE.py:343-462).
    shard_rows > 0 additionally writes the same rows as `.npy` shards of that many rows plus a manifest
    (latent_shards.py), the form multi-million-patch runs use instead of one pickle."""
    cfg = config_.latent_encoding
    channels = cfg.channels
    network = cfg.network
    weights_dir = cfg.weights
    save_output = getattr(cfg, "save_output", False)
    assert len(channels) > 0, "At least one channel must be specified"

    model_path = os.path.join(weights_dir, 'model.pt')
    model_name = os.path.basename(weights_dir)
    output_dir = os.path.join(raw_folder, model_name)
    os.makedirs(output_dir, exist_ok=True)

    assert len(set(site[:2] for site in sites)) == 1, "Sites should be from a single well/condition"
    well = sites[0][:2]

    print(f"\tloading file paths {os.path.join(raw_folder, '%s_file_paths.pkl' % well)}")
    with open(os.path.join(raw_folder, '%s_file_paths.pkl' % well), 'rb') as f:
        fs = pickle.load(f)
    print(f"\tloading static patches {os.path.join(raw_folder, '%s_static_patches.pkl' % well)}")
    with open(os.path.join(raw_folder, '%s_static_patches.pkl' % well), 'rb') as f:
        dataset = pickle.load(f)
    dataset = np.squeeze(dataset)
    assert dataset.ndim == 4, "dataset tensor dimension can only be 4, not {}".format(dataset.ndim)
    assert len(fs) == dataset.shape[0], "file paths and patches disagree"
    device = torch.device('cuda:%d' % gpu)
    torch.cuda.set_device(device)       # streams, events and library state of this call belong to `gpu`
    print('Encoding images using gpu {}...'.format(gpu))
    if 'VAE' not in network:
        raise ValueError('Network {} is not available'.format(network))
    network_cls = getattr(vae, network)
    model = network_cls(num_inputs=dataset.shape[1],
                        num_hiddens=cfg.num_hiddens,
                        num_residual_hiddens=cfg.num_residual_hiddens,
                        num_residual_layers=2,
                        num_embeddings=cfg.num_embeddings,
                        commitment_cost=getattr(cfg, "commitment_cost", 0.25),
                        gpu=True)
    model = model.to(device)
    try:
        model.load_state_dict(torch.load(model_path, map_location=device))
    except Exception as ex:
        print(ex)
        raise ValueError("Error in loading model weights for VQ-VAE")

    # the reference keys results by file name and re-stacks them in `fs` order (patch_VAE.py:450-455);
    # encoding in dataset order is the same thing as long as names are unique
    order = {f: i for i, f in enumerate(fs)}
    take = np.asarray([order[f] for f in fs])
    if shard_rows > 0 and kwargs.get("stream", False):
        # multi-million-patch form: no (N, D*h*w) array is ever held; shards + manifest only (no single pickle)
        assert np.array_equal(take, np.arange(len(fs))), "streaming needs unique file names (dataset order == fs order)"
        from ..latent_shards import merge_manifests
        encode_patches_to_shards(model, dataset, device, output_dir, well, bn_mode=bn_mode, rows_per_shard=shard_rows,
                                 block_rows=kwargs.get("block_rows", 65536))
        for kind in ("latent_space", "latent_space_after"):
            merge_manifests(output_dir, well, kind)
        return output_dir
    z_b, z_a = encode_patches(model, dataset, device, bn_mode=bn_mode)
    dats = np.ascontiguousarray(z_b[take]).reshape((len(fs), -1))
    print(f"\tsaving {os.path.join(output_dir, '%s_latent_space.pkl' % well)}")
    with open(os.path.join(output_dir, '%s_latent_space.pkl' % well), 'wb') as f:
        pickle.dump(dats, f, protocol=4)
    dats = np.ascontiguousarray(z_a[take]).reshape((len(fs), -1))
    print(f"\tsaving {os.path.join(output_dir, '%s_latent_space_after.pkl' % well)}")
    with open(os.path.join(output_dir, '%s_latent_space_after.pkl' % well), 'wb') as f:
        pickle.dump(dats, f, protocol=4)
    if shard_rows > 0:
        from ..latent_shards import ShardedLatentWriter, merge_manifests
        for kind, arr in (("latent_space", z_b), ("latent_space_after", z_a)):
            with ShardedLatentWriter(output_dir, well, kind, int(np.prod(arr.shape[1:])), rows_per_shard=shard_rows) as w:
                w.append(np.ascontiguousarray(arr[take]).reshape((len(fs), -1)))
            merge_manifests(output_dir, well, kind)
    if save_output:
        save_reconstructions(model, dataset, output_dir, device)
    return output_dir
