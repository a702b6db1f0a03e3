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
    assert es.best_score == -0.9 and es.val_loss_min == 0.9        # the reference tracks the negated loss
    never = EarlyStopping(patience=None, path=str(tmp_path / "n.pt"))
    for v in (1.0, 2.0, 3.0, 4.0):
        never(v, net)
    assert never.counter == 3 and not never.early_stop
    es2 = EarlyStopping(patience=1, delta=0.1, path=str(tmp_path / "d.pt"))
    es2(1.0, net); es2(0.95, net)                                  # better, but by less than delta
    assert es2.early_stop and es2.val_loss_min == 1.0


def test_training_data_assembly_helpers():
    """run_training.py:97-159 / :299-321 / :335-355 restated: trajectories stay contiguous, the relation matrix follows
    the permutation, offsets shift ids, and the batch relation block is the dense slice."""
    from scipy.sparse import csr_matrix
    from torch.utils.data import TensorDataset
    from dynamorph_b200.run_training import concat_relations, get_relation_tensor, reorder_with_trajectories, zscore
    rel_a = {(0, 1): 2, (1, 0): 2, (1, 2): 2, (2, 1): 2, (0, 2): 1, (2, 0): 1}
    rel_b = {(0, 1): 2, (1, 0): 2}
    merged, labels = concat_relations([rel_a, rel_b], [np.arange(4), np.arange(3)], [0, 4])
    assert merged[(4, 5)] == 2 and merged[(0, 2)] == 1 and len(merged) == 8
    assert labels.tolist() == [0, 1, 2, 3, 4, 5, 6]
    ds = TensorDataset(torch.arange(7.).reshape(7, 1, 1, 1))
    out, mat, order = reorder_with_trajectories(ds, merged, seed=5)
    assert sorted(order) == list(range(7))
    pos = {v: i for i, v in enumerate(order)}
    assert max(pos[0], pos[1], pos[2]) - min(pos[0], pos[1], pos[2]) == 2       # trajectory {0,1,2} is contiguous
    assert abs(pos[4] - pos[5]) == 1
    assert out.tensors[0].reshape(-1).tolist() == [float(v) for v in order]
    dense = np.asarray(mat.todense())
    for (a, b), v in merged.items():
        assert dense[pos[a], pos[b]] == v
    assert dense.sum() == sum(merged.values())
    again = reorder_with_trajectories(ds, merged, seed=5)[2]
    assert again == order
    blk = get_relation_tensor(mat, [pos[0], pos[1], pos[6]], device=None)
    assert blk.dtype == torch.float32 and blk.tolist() == [[0., 2., 0.], [2., 0., 0.], [0., 0., 0.]]
    assert get_relation_tensor(None, [0, 1]) is None
    x = np.random.RandomState(0).rand(5, 2, 4, 4) * 3 + 1
    z = zscore(x)
    assert np.allclose(z.mean(axis=(0, 2, 3)), 0, atol=1e-12) and np.allclose(z.std(axis=(0, 2, 3)), 1, atol=1e-9)
    z2 = zscore(x, channel_mean=[1., 2.], channel_std=[2., 4.])
    assert np.allclose(z2[:, 1], (x[:, 1] - 2.) / (4. + np.finfo(float).eps))


def test_run_vae_cli_rejects_other_methods(tmp_path):
    from dynamorph_b200 import run_VAE
    import types
    cfg = types.SimpleNamespace(latent_encoding=types.SimpleNamespace(weights="w", gpu_ids=[0], fov=None))
    with pytest.raises(ValueError):
        run_VAE.main("assemble", str(tmp_path), None, cfg)
    with pytest.raises(AttributeError):
        run_VAE.main("process", None, None, cfg)
    cfg.latent_encoding.weights = None
    with pytest.raises(AttributeError):
        run_VAE.main("process", str(tmp_path), None, cfg)


def test_get_im_sites_follows_the_reference(tmp_path):
    """SingleCellPatch/extract_patches.py:337-350: stems of the .npy files, `_NN` maps excluded; assembled-only
    directories fall back to one pseudo-site per well."""
    from dynamorph_b200.run_VAE import get_im_sites
    for f in ("B2-Site_0.npy", "B2-Site_0_NN.npy", "B2-Site_3.npy", "C5-Site_1.npy", "notes.txt"):
        (tmp_path / f).write_bytes(b"")
    assert get_im_sites(str(tmp_path)) == ["B2-Site_0", "B2-Site_3", "C5-Site_1"]
    only = tmp_path / "assembled"
    only.mkdir()
    (only / "D4_static_patches.pkl").write_bytes(b"")
    assert get_im_sites(str(only)) == ["D4-Site_0"]


def test_augmentation_draws_follow_the_reference_rng_order():
    """draw_augmentation (the host half of the one-launch device augmentation) consumes np.random exactly like the
    per-sample loop of run_training.py:396-403: applying its (flip, rot) bytes reproduces the oracle's augment_batch
    on the same stream.  augment_batch itself is GPU-only (no CPU fallback): CPU input fails loudly."""
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
    assert torch.equal(O.augment_batch(x, np.random), ref)      # the module-level stream is the same generator
    assert np.random.randint(1 << 30) == nxt
    with pytest.raises(RuntimeError, match="GPU"):
        augment_batch(x.clone())


def test_time_matching_descriptor_variants():
    """VQ_VAE carries no w_a / margin (sum variant, vq_vae.py:324-332); VQ_VAE_z16 / z32 do (hinged mean,
    vae.py:321-336).  The descriptor needs a CUDA matrix: CPU input fails loudly."""
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
    from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
    from dynamorph_b200.matching import descriptor
    assert not hasattr(VQ_VAE(), "w_a") and hasattr(VQ_VAE_z16(), "w_a")
    with pytest.raises(RuntimeError, match="CUDA"):
        descriptor(VQ_VAE_z16(), torch.zeros(4, 4))
