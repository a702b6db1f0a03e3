"""-m gpu: the CUDA path (through the drop-in classes -> C ABI) against the golden fixtures
written by the unmodified reference modules and against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


def test_library_loaded():
    from dynamorph_b200 import _lib
    lib = _lib.load()
    assert lib.dmb_abi_version() == 1
    maps = open("/proc/self/maps").read()
    assert "libdynamorph_b200.so" in maps


def test_vq_op_bit_exact_on_reference_latents(golden_case, U):
    """Same z as the reference -> identical indices and straight-through values (no tolerance)."""
    g = golden_case
    st = g.state()
    m = U.model_from_state(st)
    z = g.t("eval/z_before").cuda()
    z_st, loss, ppl = m.vq(z)
    idx = m.vq.encode_inputs(z)
    assert idx.dtype == torch.int64 and idx.shape == z.shape[:1] + z.shape[2:]
    assert np.array_equal(idx.cpu().numpy().astype(np.int32), g["eval/idx"])
    assert torch.equal(z_st.cpu(), g.t("eval/z_after"))
    assert U.rel(loss, g.t("eval/vq_loss")) < U.REL_TOL
    assert U.rel(ppl, g.t("eval/perplexity")) < U.REL_TOL
    q = m.vq.decode_inputs(idx)
    assert torch.equal(q.cpu(), g.t("eval/decode_inputs"))


def test_vq_ties_first_index(U):
    g = Golden("vq_edge")
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VectorQuantizer
    vq = VectorQuantizer(16, 64, 0.25).cuda()
    with torch.no_grad():
        vq.w.weight.copy_(g.t("codebook"))
    z = g.t("z").cuda()
    z_st, loss, ppl = vq(z)
    idx = vq.encode_inputs(z)
    assert np.array_equal(idx.cpu().numpy().astype(np.int32), g["idx"])
    assert int(idx[0, 0, 0]) == 3
    assert torch.equal(z_st.cpu(), g.t("z_st"))
    assert U.rel(loss, g.t("loss")) < U.REL_TOL and U.rel(ppl, g.t("perplexity")) < U.REL_TOL


@pytest.mark.parametrize("D,K", [(16, 64), (64, 512), (32, 7)])
def test_vq_search_exact_on_adversarial_codebooks(D, K, U):
    """The argmin must reproduce torch.argmax(-distances) of vq_vae.py:65-66 bit for bit on the hard cases: exactly
    duplicated codebook rows (lowest index wins), rows that differ by 1e-7, latents sitting on a code, latents
    equidistant from two codes.  (A two-phase search -- |e|^2 - 2 z.e scan + exact re-check of the candidates inside
    an error margin -- passed this test too but was not faster under SIMT than the direct form, and was dropped.)"""
    from dynamorph_b200.engine import vq_indices
    g = torch.Generator().manual_seed(D * 1000 + K)
    cb = torch.randn(K, D, generator=g)
    if K >= 32:
        cb[10:16] = cb[10]                                              # exact duplicates: lowest index must win
        cb[20:28] = cb[20] + torch.randn(8, D, generator=g) * 1e-7      # near duplicates: > 4 candidates in the margin
    B, H = 6, 16
    z = torch.randn(B, D, H, H, generator=g) * 1.3
    flat = z.permute(0, 2, 3, 1).reshape(-1, D)
    n = flat.shape[0]
    pick = torch.randint(0, K, (n,), generator=g)
    near = cb[pick] + torch.randn(n, D, generator=g) * 1e-6
    sel = torch.rand(n, generator=g)
    flat = torch.where((sel < 0.4).unsqueeze(1), near, flat)             # 40 % of the positions hug a code
    mid = 0.5 * (cb[pick] + cb[(pick + 1) % K])
    flat = torch.where((sel > 0.9).unsqueeze(1), mid, flat)              # 10 % sit between two codes
    z = flat.reshape(B, H, H, D).permute(0, 3, 1, 2).contiguous()
    ref = O.vq_indices(z, cb)
    got = vq_indices(z.cuda(), cb.cuda()).cpu().long()
    assert torch.equal(got, ref), int((got != ref).sum())


def test_eval_encode(golden_case, U):
    g = golden_case
    st = g.state()
    m = U.model_from_state(st).eval()
    x = g.t("x_eval").cuda()
    with torch.no_grad():
        zb = m.enc(x)
        za, loss, ppl = m.vq(zb)
        idx = m.vq.encode_inputs(zb)
    assert U.rel(zb, g.t("eval/z_before")) < U.REL_TOL
    flips = U.check_indices(idx, g.t("eval/z_before"), st["vq.w.weight"], g["eval/idx"], g.name)
    if flips == 0:
        assert U.rel(za, g.t("eval/z_after")) < U.REL_TOL
        assert U.rel(loss, g.t("eval/vq_loss")) < U.REL_TOL
        assert U.rel(ppl, g.t("eval/perplexity")) < U.REL_TOL
    # fused enc+vq entry point gives the same answer as the two calls
    zb2, za2, idx2 = m.encode_latents(x, "eval")
    assert torch.equal(zb2, zb) and torch.equal(za2, za) and torch.equal(idx2.long(), idx)
    assert torch.equal(m.encode(x), idx)


def test_per_sample_encode(golden_case, U):
    """As-written process_VAE semantics: train-mode BN with batch 1 (patch_VAE.py:445-449)."""
    g = golden_case
    st = g.state()
    m = U.model_from_state(st)
    x = g.t("x_eval").cuda()
    zb, za, idx = m.encode_latents(x, "per_sample")
    assert U.rel(zb, g.t("per_sample/z_before")) < U.REL_TOL
    idx_ref = O.vq_indices(g.t("per_sample/z_before"), st["vq.w.weight"])
    flips = U.check_indices(idx, g.t("per_sample/z_before"), st["vq.w.weight"], idx_ref, g.name)
    if flips == 0:
        assert U.rel(za, g.t("per_sample/z_after")) < U.REL_TOL
    # the literal loop (train mode, batch 1) through model.enc / model.vq agrees too
    m.train()
    with torch.no_grad():
        zb1 = torch.cat([m.enc(x[i:i + 1]) for i in range(x.shape[0])])
    assert U.rel(zb1, g.t("per_sample/z_before")) < U.REL_TOL
    # batch composition must not matter in per_sample mode: shard == whole, bit for bit
    zb_a, _, idx_a = m.encode_latents(x[:1], "per_sample")
    assert torch.equal(zb_a, zb[:1]) and torch.equal(idx_a, idx[:1])


def test_batch_mode_encoder_and_running_stats(golden_case, U):
    g = golden_case
    st = g.state()
    m = U.model_from_state(st).train()
    x = g.t("x_train").cuda()
    with torch.no_grad():
        zb = m.enc(x)
    nr = {}
    with torch.no_grad():
        ref = O.encoder(g.t("x_train"), st, O.BATCH, nr)
    assert U.rel(zb, ref) < U.REL_TOL
    sd = m.state_dict()
    for k, v in nr.items():
        if k.startswith("enc."):
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), k
            else:
                assert U.rel(sd[k], v) < U.REL_TOL, k


def test_decoder_eval(golden_case, U):
    g = golden_case
    st = g.state()
    m = U.model_from_state(st).eval()
    with torch.no_grad():
        dec = m.dec(g.t("eval/z_after").cuda())
    assert U.rel(dec, g.t("eval/decoded")) < U.REL_TOL


def test_forward_eval_losses(golden_case, U):
    g = golden_case
    st = g.state()
    m = U.model_from_state(st).eval()
    x = g.t("x_eval").cuda()
    with torch.no_grad():
        dec, d = m(x)
    idx = m.vq.encode_inputs(m.enc(x))
    flips = U.check_indices(idx, g.t("eval/z_before"), st["vq.w.weight"], g["eval/idx"], g.name)
    if flips:
        pytest.skip("near-tie flip changes downstream values")
    assert U.rel(dec, g.t("eval/decoded")) < U.REL_TOL
    for k in ("recon_loss", "commitment_loss", "total_loss", "perplexity"):
        assert abs(float(d[k]) - float(g["eval/loss/" + k])) <= U.REL_TOL * abs(float(g["eval/loss/" + k])), k
    assert set(d) == {"recon_loss", "commitment_loss", "time_matching_loss", "total_loss", "perplexity"}


def test_cpu_tensors_are_rejected(U, golden_default):
    m = U.model_from_state(golden_default.state()).eval()
    with pytest.raises(RuntimeError):
        m.enc(golden_default.t("x_eval"))


@pytest.mark.parametrize("B", [1, 3, 37])
def test_ragged_batches_match_oracle(B, U, golden_default):
    st = golden_default.state()
    m = U.model_from_state(st).eval()
    x = O.synthetic_patches(B, 900 + B)
    with torch.no_grad():
        ref = O.encoder(x, st, O.EVAL)
        zb, za, idx = m.encode_latents(x.cuda(), "eval")
    assert U.rel(zb, ref) < U.REL_TOL
    U.check_indices(idx, ref, st["vq.w.weight"], O.vq_indices(ref, st["vq.w.weight"]), f"B={B}")


def test_empty_and_single_patch_batches(U):
    """Edge cases of the batch axis: B = 0 returns empty tensors (torch semantics of the reference modules; no kernel
    launches), B = 1 and an odd B equal the corresponding rows of a larger batch bit for bit in the patch-independent
    BatchNorm modes (ragged tails of a chunked bulk encode)."""
    from dynamorph_b200.bulk import BulkEncoder
    g = Golden("vqvae_default")
    m = U.model_from_state(g.state()).eval()
    x = O.synthetic_patches(9, 5).cuda()
    for mode in ("eval", "per_sample"):
        zb, za, idx = m.encode_latents(x, mode)
        e = m.encode_latents(x[:0], mode)
        assert e[0].shape == (0, 16, 16, 16) and e[1].shape == (0, 16, 16, 16) and e[2].shape[0] == 0
        for n in (1, 7):
            zb1, za1, idx1 = m.encode_latents(x[:n].contiguous(), mode)
            assert torch.equal(zb1, zb[:n]) and torch.equal(za1, za[:n]) and torch.equal(idx1, idx[:n])
    with torch.no_grad():
        assert m.enc(x[:0]).shape == (0, 16, 16, 16)
        assert m.dec(torch.empty(0, 16, 16, 16, device="cuda")).shape == (0, 2, 128, 128)
    out = BulkEncoder(m, chunk=4).encode(torch.empty(0, 2, 128, 128))
    assert out["z_before"].shape == (0, 4096) and out["idx"].shape == (0, 256)
