"""CPU, world_size 2 over gloo: the host-side multi-rank logic (shard ranges, ragged index gather, the single
flat-gradient allreduce of a data-parallel step).  The arithmetic inside each rank is the oracle's here -- the
CUDA path is covered by the -m gpu tests; what is under test is the sharding / collective plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import Golden
from oracle import vqvae_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    for r in res:
        if isinstance(r, Exception) or (isinstance(r, tuple) and r and r[0] == "error"):
            raise AssertionError(r)
    return sorted(res, key=lambda t: t[0])


def _entry(fn, rank, world, port, q, *args):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("error", rank, traceback.format_exc()))


def test_shard_range_partitions():
    from dynamorph_b200.dist import shard_range
    for n in (0, 1, 7, 8, 9, 4_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def _encode_shard(rank, world, n):
    from dynamorph_b200.dist import gather_code_indices, shard_range
    g = Golden("vqvae_default")
    st = g.state()
    x = O.synthetic_patches(n, 31)
    a, b = shard_range(n, rank, world)
    with torch.no_grad():
        zb = O.encoder(x[a:b], st, O.EVAL)
        idx = O.vq_indices(zb, st["vq.w.weight"])
    full = gather_code_indices(idx, n, 64)
    return full.numpy()


def test_sharded_encode_equals_single_rank():
    """Encode is embarrassingly parallel: ragged 2-rank shards gathered == one-rank result, bit for bit."""
    n = 5
    res = _run(_encode_shard, 2, n)
    g = Golden("vqvae_default")
    st = g.state()
    with torch.no_grad():
        ref = O.vq_indices(O.encoder(O.synthetic_patches(n, 31), st, O.EVAL), st["vq.w.weight"])
    for rank, full in res:
        assert full.dtype == np.uint8 and full.shape == (n, 16, 16)
        assert np.array_equal(full.astype(np.int64), ref.numpy())


def _dp_step(rank, world):
    from dynamorph_b200.dist import allreduce_flat, shard_range
    g = Golden("vqvae_default")
    st = g.state()
    x = g.t("x_train")
    a, b = shard_range(x.shape[0], rank, world)
    _, losses, grads, _ = O.loss_and_grads(x[a:b], st, O.BATCH)
    keys = O.trainable_keys(st)
    flat = torch.cat([grads[k].reshape(-1) for k in keys])
    allreduce_flat(flat, average=True)
    return flat.numpy()


def test_dp_gradient_allreduce_matches_rank_emulation():
    """DP oracle: R independent reference passes on the shards, gradients averaged."""
    res = _run(_dp_step, 2)
    g = Golden("vqvae_default")
    st = g.state()
    x = g.t("x_train")
    keys = O.trainable_keys(st)
    parts = []
    for r in range(2):
        a, b = (0, 2) if r == 0 else (2, 4)
        _, _, grads, _ = O.loss_and_grads(x[a:b], st, O.BATCH)
        parts.append(torch.cat([grads[k].reshape(-1) for k in keys]))
    ref = (parts[0] + parts[1]) / 2
    for rank, flat in res:
        assert np.allclose(flat, ref.numpy(), rtol=1e-4, atol=1e-5 * float(ref.abs().max()))
    assert np.array_equal(res[0][1], res[1][1])     # replicas stay identical


def _shard_writer(rank, world, out_dir, n, width):
    """Each rank writes its patch range of a deterministic latent matrix into the sharded store; rank 0 merges."""
    from dynamorph_b200.dist import shard_range
    from dynamorph_b200.latent_shards import ShardedLatentWriter, merge_manifests
    z = np.arange(n * width, dtype=np.float32).reshape(n, width)
    a, b = shard_range(n, rank, world)
    with ShardedLatentWriter(out_dir, "A1", "latent_space", width, rank=rank, world=world, first_row=a,
                             rows_per_shard=5) as w:
        for s in range(a, b, 3):                      # ragged appends crossing shard boundaries
            w.append(z[s:min(b, s + 3)])
    dist.barrier()                                    # every rank's manifest is on disk
    if rank == 0:
        merge_manifests(out_dir, "A1", "latent_space")
    dist.barrier()
    return b - a


def test_sharded_latent_store_across_ranks(tmp_path):
    """process_VAE at scale: ranks own contiguous patch ranges, write their own shards with no
    data-path collective, and the merged manifest reads back as the reference's (N, D*h*w) matrix."""
    from dynamorph_b200.latent_shards import open_latents
    n, width = 23, 12
    res = _run(_shard_writer, 2, str(tmp_path), n, width)
    assert sum(r[1] for r in res) == n
    v = open_latents(str(tmp_path), "A1", "latent_space")
    assert v.shape == (n, width)
    assert np.array_equal(v.to_array(), np.arange(n * width, dtype=np.float32).reshape(n, width))
