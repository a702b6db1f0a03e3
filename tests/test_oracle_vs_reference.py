"""CPU, build container only: the oracle restatement against the UNMODIFIED reference
modules imported from /root/reference (absent on the GPU box -> skipped there; the
committed fixtures in tests/golden carry the same pins)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import REFERENCE, rel_err
from oracle import vqvae_oracle as O

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "HiddenStateExtractor")),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REFERENCE)
    for name in ("matplotlib", "matplotlib.pyplot", "h5py", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub = types.ModuleType(name)
                stub.use = lambda *a, **k: None
                sys.modules[name] = stub
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    from HiddenStateExtractor import vq_vae, vae
    import run_training
    from pipeline import train_utils
    yield types.SimpleNamespace(vq_vae=vq_vae, vae=vae, rt=run_training, tu=train_utils)
    sys.path.remove(REFERENCE)


def _mk(ref, which, **kw):
    torch.manual_seed(4)
    cls = {"vqvae": ref.vq_vae.VQ_VAE, "z16": ref.vae.VQ_VAE_z16, "z32": ref.vae.VQ_VAE_z32}[which]
    m = cls(device="cpu", **kw)
    st = O.calibrate_state({k: v.detach().clone() for k, v in m.state_dict().items()},
                           O.synthetic_patches(16, 9), seed=4)
    m.load_state_dict(st)
    return m, st


CASES = [("vqvae", {}), ("z16", {}), ("z32", {}),
         ("vqvae", dict(num_hiddens=32, num_residual_hiddens=16, num_residual_layers=1, num_embeddings=40)),
         ("z32", dict(num_hiddens=32, num_embeddings=128))]


@pytest.mark.parametrize("which,kw", CASES)
def test_state_dict_keys_and_default_state(ref, which, kw):
    m, st = _mk(ref, which, **kw)
    mine = O.default_state("z32" if which == "z32" else "z16", **kw)
    assert list(mine) == list(m.state_dict())
    for k, v in m.state_dict().items():
        assert tuple(mine[k].shape) == tuple(v.shape), k
    assert O.trainable_keys(st) == [k for k, p in m.named_parameters() if p.requires_grad]


@pytest.mark.parametrize("which,kw", CASES)
def test_forward_modes(ref, which, kw):
    m, st = _mk(ref, which, **kw)
    x = O.synthetic_patches(3, 21)
    m.train()
    dec, d = m(x)
    dec2, d2 = O.forward(x, st, O.BATCH)
    assert torch.equal(dec, dec2)
    for k in ("recon_loss", "commitment_loss", "total_loss", "perplexity"):
        assert float(d[k]) == float(d2[k]), k
    m.load_state_dict(st)
    m.eval()
    with torch.no_grad():
        zb = m.enc(x)
        za, l, p = m.vq(zb)
        assert torch.equal(zb, O.encoder(x, st, O.EVAL))
        za2, l2, p2, idx2 = O.vq_forward(zb, st["vq.w.weight"], m.commitment_cost)
        assert torch.equal(za, za2) and float(l) == float(l2) and float(p) == float(p2)
        assert torch.equal(m.vq.encode_inputs(zb), idx2)
        assert torch.equal(m.vq.decode_inputs(idx2), O.vq_gather(idx2, st["vq.w.weight"]))
        assert torch.equal(m.dec(za), O.decoder(za, st, O.EVAL))
    m.train()
    zbs = torch.cat([m.enc(x[i:i + 1]) for i in range(3)]).detach()
    with torch.no_grad():
        assert torch.equal(zbs, O.encoder(x, st, O.PER_SAMPLE))


@pytest.mark.parametrize("which", ["vqvae", "z16", "z32"])
def test_time_matching_and_mask(ref, which):
    m, st = _mk(ref, which)
    x = O.synthetic_patches(4, 22)
    g = torch.Generator().manual_seed(0)
    mat = torch.randint(0, 3, (4, 4), generator=g).float()
    mat = torch.triu(mat) + torch.triu(mat, 1).T
    mask = torch.rand(4, 1, 128, 128, generator=g)
    m.train()
    dec, d = m(x, time_matching_mat=mat, batch_mask=mask)
    kw = {} if which == "vqvae" else dict(w_a=m.w_a, w_t=m.w_t, w_n=m.w_n, margin=m.margin)
    dec2, d2 = O.forward(x, st, O.BATCH, weight_matching=m.weight_matching, time_matching_mat=mat,
                         batch_mask=mask, tm_variant="sum" if which == "vqvae" else "hinge", **kw)
    for k in ("recon_loss", "time_matching_loss", "total_loss"):
        assert abs(float(d[k]) - float(d2[k])) <= 1e-6 * abs(float(d[k])), k


def test_run_one_batch_train_steps(ref):
    m, st = _mk(ref, "z16")
    x = O.synthetic_patches(4, 23)
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, betas=(.9, .999))
    tl = {}
    state = {k: v.clone() for k, v in st.items()}
    ost = {"m": {}, "v": {}}
    for s in range(3):
        m, tl = ref.rt.run_one_batch(m, x.clone(), tl, model_kwargs={}, optimizer=opt,
                                     transform=None, training=True)
        l = O.train_step(x, state, ost, s + 1, 1e-3)
        assert abs(tl["total_loss"][-1] - float(l["total_loss"])) < 1e-5 * abs(tl["total_loss"][-1])
    noise = set(O.bias_feeds_train_bn(st))
    for k, v in m.state_dict().items():
        if k in noise or k.endswith("num_batches_tracked"):
            continue
        assert rel_err(state[k], v) < 2e-5, k


def test_augmentation_matches_run_one_batch(ref):
    """run_training.py:396-403 consumes np.random in a fixed order; the restatement must too."""
    class Probe(torch.nn.Module):
        def forward(self, batch):
            self.seen = batch.clone()
            return None, {}
    x = O.synthetic_patches(5, 24)
    np.random.seed(77)
    p = Probe()
    ref.rt.run_one_batch(p, x.clone(), {}, model_kwargs={}, optimizer=None, transform=True, training=False)
    mine = O.augment_batch(x, np.random.RandomState(77))
    assert torch.equal(p.seen, mine)


def test_zscore_patch_and_process_loop(ref):
    rng = np.random.RandomState(5)
    raw = rng.rand(3, 2, 1, 128, 128) * 1000 + 10
    assert np.array_equal(ref.tu.zscore_patch(np.squeeze(raw)), O.zscore_patch(np.squeeze(raw))) or \
        np.allclose(ref.tu.zscore_patch(np.squeeze(raw)), O.zscore_patch(np.squeeze(raw)), rtol=1e-13, atol=1e-13)
    m, st = _mk(ref, "z16")
    m.train()
    # the loop of pipeline/patch_VAE.py:418-419,445-462 against the unmodified model
    data = torch.from_numpy(ref.tu.zscore_patch(np.squeeze(raw))).float()
    zb, za = [], []
    for i in range(3):
        s = data[i:i + 1].reshape([-1, 2, 128, 128])
        z_b = m.enc(s)
        z_a, _, _ = m.vq(z_b)
        zb.append(z_b.cpu().data.numpy())
        za.append(z_a.cpu().data.numpy())
    zb = np.stack(zb, 0).reshape((3, -1))
    za = np.stack(za, 0).reshape((3, -1))
    ob, oa = O.process_vae_arrays(raw, st, O.PER_SAMPLE)
    assert np.array_equal(zb, ob) and np.array_equal(za, oa)


def test_training_glue_matches_reference(ref, tmp_path):
    """The host glue of the drop-in trainer (dynamorph_b200/run_training.py, pipeline/train_utils.py) against the
    reference's own functions on the same inputs and seeds: trajectory reordering (run_training.py:97-159), relation
    merging (:299-321), the batch relation block (:335-355), dataset z-score (train_utils.py:228-250) and
    EarlyStopping's decisions (train_utils.py:8-60)."""
    from torch.utils.data import TensorDataset
    from dynamorph_b200 import run_training as mine
    from dynamorph_b200.pipeline.train_utils import EarlyStopping
    rng = np.random.RandomState(3)
    n = 40
    rel = {}
    for a in range(0, 30, 5):                      # six trajectories of five frames, rest singletons
        for i in range(a, a + 4):
            rel[(i, i + 1)] = 2; rel[(i + 1, i)] = 2
        for i in range(a, a + 5):
            for j in range(a, a + 5):
                if abs(i - j) > 1:
                    rel[(i, j)] = 1
    ds = TensorDataset(torch.from_numpy(rng.rand(n, 2, 4, 4).astype(np.float32)))
    r_ds, r_mat, r_order = ref.rt.reorder_with_trajectories(ds, rel, seed=123)
    m_ds, m_mat, m_order = mine.reorder_with_trajectories(ds, rel, seed=123)
    assert [int(v) for v in r_order] == [int(v) for v in m_order]
    assert torch.equal(r_ds.tensors[0], m_ds.tensors[0])
    assert (r_mat != m_mat).nnz == 0
    ids = [3, 17, 4, 29, 30, 5]
    assert torch.equal(ref.rt.get_relation_tensor(r_mat, ids, device=None),
                       mine.get_relation_tensor(m_mat, ids, device=None))
    r_rel, r_lab = ref.rt.concat_relations([rel, {(0, 1): 2}], [np.arange(n), np.arange(2)], [0, n])
    m_rel, m_lab = mine.concat_relations([rel, {(0, 1): 2}], [np.arange(n), np.arange(2)], [0, n])
    assert r_rel == m_rel and np.array_equal(r_lab, m_lab)
    x = rng.rand(6, 2, 8, 8) * 5
    assert np.array_equal(ref.tu.zscore(x), mine.zscore(x))
    assert np.array_equal(ref.tu.zscore(x, [1., 2.], [3., 4.]), mine.zscore(x, [1., 2.], [3., 4.]))
    net = torch.nn.Linear(2, 2)
    losses = [1.0, 0.8, 0.85, 0.79, 0.795, 0.9, 0.91, 0.5]
    a = ref.tu.EarlyStopping(patience=3, delta=0.005, path=str(tmp_path / "a.pt"), trace_func=lambda *_: None)
    b = EarlyStopping(patience=3, delta=0.005, path=str(tmp_path / "b.pt"), trace_func=lambda *_: None)
    for v in losses:
        a(v, net); b(v, net)
        assert (a.counter, a.early_stop, a.best_score, a.val_loss_min) == (b.counter, b.early_stop, b.best_score,
                                                                           b.val_loss_min)
