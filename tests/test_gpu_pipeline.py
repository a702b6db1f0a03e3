"""-m gpu: the process_VAE drop-in end to end (pickles in -> pickles out) against the oracle's restatement
of pipeline/patch_VAE.py:413-462, and the host-buffer bulk encoder."""
import os
import pickle
import types

import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu


def _raw_patches(n, seed):
    rng = np.random.RandomState(seed)
    base = O.synthetic_patches(n, seed).numpy().astype(np.float64)
    raw = base * np.array([900.0, 12.0]).reshape(1, 2, 1, 1) + np.array([32000.0, 40.0]).reshape(1, 2, 1, 1)
    return raw.reshape(n, 2, 1, 128, 128) + rng.rand(n, 2, 1, 128, 128) * 1e-3     # singleton Z axis as on disk


@pytest.mark.parametrize("bn_mode", ["per_sample", "eval"])
def test_process_vae_roundtrip(tmp_path, bn_mode):
    import gpu_util as U
    from dynamorph_b200.pipeline.patch_VAE import process_VAE
    g = Golden("vqvae_default")
    st = g.state()
    raw_dir = tmp_path / "raw"
    wdir = tmp_path / "weights" / "my_model"
    os.makedirs(raw_dir); os.makedirs(wdir)
    torch.save(st, wdir / "model.pt")
    n = 11
    raw = _raw_patches(n, 3)
    fs = [f"/data/C5-Site_{i % 3}/patch_{i}.h5" for i in range(n)]
    pickle.dump(fs, open(raw_dir / "C5_file_paths.pkl", "wb"))
    pickle.dump(raw, open(raw_dir / "C5_static_patches.pkl", "wb"), protocol=4)
    cfg = types.SimpleNamespace(latent_encoding=types.SimpleNamespace(
        weights=str(wdir), channels=[0, 1], num_hiddens=16, num_residual_hiddens=32, num_embeddings=64,
        commitment_cost=0.25, network="VQ_VAE_z16", save_output=False, channel_mean=None, channel_std=None))
    out_dir = process_VAE(str(raw_dir), None, ["C5-Site_0", "C5-Site_1"], cfg, gpu=0, bn_mode=bn_mode)
    assert out_dir == str(raw_dir / "my_model")
    zb = pickle.load(open(raw_dir / "my_model" / "C5_latent_space.pkl", "rb"))
    za = pickle.load(open(raw_dir / "my_model" / "C5_latent_space_after.pkl", "rb"))
    assert zb.dtype == np.float32 and zb.shape == (n, 16 * 16 * 16) and za.shape == zb.shape
    rb, ra = O.process_vae_arrays(raw, st, O.PER_SAMPLE if bn_mode == "per_sample" else O.EVAL)
    assert U.rel(zb, rb) < U.REL_TOL
    # z_after rows equal codebook rows gathered by index: compare where indices agree
    ref_idx = O.vq_indices(torch.from_numpy(rb).reshape(n, 16, 16, 16), st["vq.w.weight"])
    my_idx = O.vq_indices(torch.from_numpy(zb).reshape(n, 16, 16, 16), st["vq.w.weight"])
    U.check_indices(my_idx, torch.from_numpy(rb).reshape(n, 16, 16, 16), st["vq.w.weight"], ref_idx, "process_VAE")
    same = (ref_idx == my_idx).reshape(n, 1, 16, 16).expand(n, 16, 16, 16).reshape(n, -1).numpy()
    assert np.allclose(za[same], ra[same], rtol=1e-4, atol=1e-6)


def test_bulk_encoder_matches_device_path():
    import gpu_util as U
    from dynamorph_b200.bulk import BulkEncoder
    g = Golden("vqvae_default")
    m = U.model_from_state(g.state()).eval()
    x = O.synthetic_patches(37, 77)
    zb, za, idx = m.encode_latents(x.cuda(), "eval")
    for chunk in (8, 16, 64):
        out = BulkEncoder(m, chunk=chunk, bn_mode="eval").encode(x.pin_memory())
        torch.cuda.synchronize()
        assert torch.equal(out["z_before"], zb.reshape(37, -1).cpu())
        assert torch.equal(out["z_after"], za.reshape(37, -1).cpu())
        assert torch.equal(out["idx"], idx.reshape(37, -1).cpu())


@pytest.mark.parametrize("dtype", [np.uint16, np.float64, np.float32])
def test_bulk_encoder_raw_input_zscore_on_device(dtype):
    """zscore=True: raw host patches (uint16 camera counts / float64 as pickled) -> device z-score -> encode, equal
    to the oracle's restatement of patch_VAE.py:413-449 (zscore_patch in float64 on the host, float32 cast)."""
    import gpu_util as U
    from dynamorph_b200.bulk import BulkEncoder
    g = Golden("vqvae_default")
    st = g.state()
    m = U.model_from_state(st).eval()
    n = 21
    raw = _raw_patches(n, 9).reshape(n, 2, 128, 128)
    raw = np.clip(raw, 0, 65535).astype(dtype)
    out = BulkEncoder(m, chunk=8, bn_mode="per_sample", zscore=True).encode(torch.from_numpy(raw))
    torch.cuda.synchronize()
    rb, ra = O.process_vae_arrays(raw.astype(np.float64).reshape(n, 2, 1, 128, 128), st, O.PER_SAMPLE)
    assert U.rel(out["z_before"], rb) < U.REL_TOL
    with pytest.raises(ValueError):
        BulkEncoder(m, chunk=8).encode(torch.from_numpy(raw.astype(np.float64)))       # float32 only without zscore


def test_checkpoint_roundtrip_with_reference_keys(tmp_path):
    """model.pt written by the drop-in loads into a fresh drop-in and keeps the reference's keys."""
    import gpu_util as U
    g = Golden("vqvae_default")
    st = g.state()
    m = U.model_from_state(st)
    from dynamorph_b200.pipeline.train_utils import EarlyStopping
    es = EarlyStopping(patience=2, path=str(tmp_path / "model.pt"))
    es(1.0, m)
    sd = torch.load(tmp_path / "model.pt")
    assert list(sd) == list(st)
    for k in st:
        assert torch.equal(sd[k].cpu(), st[k]), k


def test_process_vae_streams_into_shards(tmp_path):
    """process_VAE(..., shard_rows=4, stream=True): blocks of patches go from the bulk encoder's D2H buffers straight into
    the sharded store; the rows equal the single-pickle output of the default path bit for bit."""
    from dynamorph_b200.latent_shards import open_latents
    from dynamorph_b200.pipeline.patch_VAE import process_VAE
    g = Golden("vqvae_default")
    raw_dir = tmp_path / "raw"
    wdir = tmp_path / "weights" / "my_model"
    os.makedirs(raw_dir); os.makedirs(wdir)
    torch.save(g.state(), wdir / "model.pt")
    n = 13
    raw = _raw_patches(n, 5)
    fs = [f"/data/C5-Site_0/patch_{i}.h5" for i in range(n)]
    pickle.dump(fs, open(raw_dir / "C5_file_paths.pkl", "wb"))
    pickle.dump(raw, open(raw_dir / "C5_static_patches.pkl", "wb"), protocol=4)
    cfg = types.SimpleNamespace(latent_encoding=types.SimpleNamespace(
        weights=str(wdir), channels=[0, 1], num_hiddens=16, num_residual_hiddens=32, num_embeddings=64,
        commitment_cost=0.25, network="VQ_VAE_z16", save_output=False, channel_mean=None, channel_std=None))
    out_dir = process_VAE(str(raw_dir), None, ["C5-Site_0"], cfg, gpu=0)
    zb = pickle.load(open(os.path.join(out_dir, "C5_latent_space.pkl"), "rb"))
    za = pickle.load(open(os.path.join(out_dir, "C5_latent_space_after.pkl"), "rb"))
    process_VAE(str(raw_dir), None, ["C5-Site_0"], cfg, gpu=0, shard_rows=4, stream=True, block_rows=5)
    vb = open_latents(out_dir, "C5", "latent_space")
    va = open_latents(out_dir, "C5", "latent_space_after")
    assert vb.shape == zb.shape and len(vb.manifest["shards"]) == 4
    assert np.array_equal(vb.to_array(), zb) and np.array_equal(va.to_array(), za)

