"""CPU: host-side logic of the drop-in package that needs no GPU."""
import os
import textwrap

import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as O


def test_drop_in_state_dict_is_seed_for_seed_identical_to_the_oracle_layout():
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
    from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16, VQ_VAE_z32
    for cls, arch in ((VQ_VAE, "z16"), (VQ_VAE_z16, "z16"), (VQ_VAE_z32, "z32")):
        m = cls()
        ref = O.default_state(arch)
        sd = m.state_dict()
        assert list(sd) == list(ref)
        for k in sd:
            assert tuple(sd[k].shape) == tuple(ref[k].shape), k
        assert not m.channel_var.requires_grad
    m = VQ_VAE(num_hiddens=64, num_embeddings=512, alpha=0.002, gpu=False)     # plot scripts pass alpha/gpu
    assert m.vq.w.weight.shape == (512, 64)


def test_cpu_model_fails_loudly():
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
    m = VQ_VAE()
    with pytest.raises(RuntimeError, match="CUDA"):
        m.enc(torch.zeros(1, 2, 128, 128))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.vq(torch.zeros(1, 16, 16, 16))
    with pytest.raises(RuntimeError):
        m.enc[12](torch.zeros(1, 16, 16, 16))        # ResidualBlock has no stand-alone path


def test_config_reader(tmp_path):
    from dynamorph_b200.configs.config_reader import YamlReader
    p = tmp_path / "c.yml"
    p.write_text(textwrap.dedent("""
        latent_encoding:
          raw_dirs: ['/data/raw']
          weights: /models/vq
          network: VQ_VAE_z16
          num_hiddens: 16
          num_residual_hiddens: 32
          num_embeddings: 64
          commitment_cost: 0.25
          channels: [0, 1]
          gpu_ids: [0, 1]
          unknown_key: 3
        training:
          learn_rate: 0.0001
          batch_size: 64
    """))
    c = YamlReader().read_config(str(p))
    assert c.latent_encoding.network == "VQ_VAE_z16" and c.latent_encoding.gpu_ids == [0, 1]
    assert c.training.batch_size == 64 and c.latent_encoding.unknown_key == 3


def test_index_packing_dtypes():
    from dynamorph_b200.dist import index_dtype, pack_indices
    assert index_dtype(64) == torch.uint8 and index_dtype(256) == torch.uint8 and index_dtype(512) == torch.int16
    idx = torch.tensor([[0, 255], [17, 3]])
    assert pack_indices(idx, 256).dtype == torch.uint8 and torch.equal(pack_indices(idx, 256).long(), idx)


def test_early_stopping_counts_and_saves(tmp_path):
    from dynamorph_b200.pipeline.train_utils import EarlyStopping
    net = torch.nn.Linear(2, 2)
    es = EarlyStopping(patience=2, path=str(tmp_path / "m.pt"))
    es(1.0, net); es(0.9, net)
    assert es.counter == 0 and os.path.exists(tmp_path / "m.pt")
    es(0.95, net)
    assert es.counter == 1 and not es.early_stop
    es(0.96, net)
    assert es.early_stop


def test_run_vae_cli_rejects_other_methods(tmp_path):
    from dynamorph_b200 import run_VAE
    import types
    cfg = types.SimpleNamespace(latent_encoding=types.SimpleNamespace(weights="w", gpu_ids=[0], fov=None))
    with pytest.raises(ValueError):
        run_VAE.main("assemble", str(tmp_path), None, cfg, "c.yml")
    with pytest.raises(AttributeError):
        run_VAE.main("process", None, None, cfg, "c.yml")


def test_augmentation_draws_follow_the_reference_rng_order():
    """draw_augmentation (the host half of the one-launch device augmentation) consumes np.random exactly like the
    per-sample loop of run_training.py:396-403: applying its (flip, rot) bytes reproduces the oracle's augment_batch
    on the same stream, and the CPU fallback of augment_batch is that loop itself."""
    from dynamorph_b200.run_training import augment_batch, draw_augmentation
    x = torch.randn(9, 2, 8, 8)
    ref = O.augment_batch(x, np.random.RandomState(77))
    np.random.seed(77)
    ops = draw_augmentation(9)
    nxt = np.random.randint(1 << 30)
    out = x.clone()
    for i, op in enumerate(ops):
        img = x[i]
        if op & 3:
            img = torch.flip(img, dims=(int(op & 3),))
        out[i] = torch.rot90(img, k=int(op >> 2), dims=[1, 2])
    assert torch.equal(out, ref)
    np.random.seed(77)
    assert torch.equal(augment_batch(x.clone()), ref)
    assert np.random.randint(1 << 30) == nxt


def test_time_matching_descriptor_variants():
    """VQ_VAE carries no w_a / margin (sum variant, vq_vae.py:324-332); VQ_VAE_z16 / z32 do (hinged mean,
    vae.py:321-336).  The descriptor needs a CUDA matrix: CPU input fails loudly."""
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
    from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
    from dynamorph_b200.matching import descriptor
    assert not hasattr(VQ_VAE(), "w_a") and hasattr(VQ_VAE_z16(), "w_a")
    with pytest.raises(RuntimeError, match="CUDA"):
        descriptor(VQ_VAE_z16(), torch.zeros(4, 4))
