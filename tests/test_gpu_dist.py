"""-m gpu, two GPUs, NCCL: the data-parallel training step on the CUDA path (BASELINE configs[4]).

  * per-rank BatchNorm statistics (DDP semantics): `FusedTrainer` on two ranks -- CUDA-graph replay of forward/backward,
    ONE NCCL allreduce of the flat gradient buffer, Adam with 1/world folded in -- against the oracle's recipe: two
    independent reference passes on the shards, gradients averaged, one Adam step.
  * `sync_bn=True`: against ONE reference pass over the concatenated batch (the reference's single-GPU semantics).

Skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import Golden
from oracle import vqvae_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]

LR = 1e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(fn, rank, world, port, q, *args):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        out = fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception:  # pragma: no cover
        import traceback
        q.put(("error", rank, traceback.format_exc()))


def _run(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    for r in res:
        if isinstance(r, tuple) and r and r[0] == "error":
            raise AssertionError(r[2])
    def revive(o):
        if isinstance(o, np.ndarray):
            return torch.from_numpy(o)
        if isinstance(o, dict):
            return {k: revive(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(revive(v) for v in o)
        return o
    return [revive(r[1]) for r in sorted(res, key=lambda t: t[0])]


def _batch(case, per_rank, world):
    """Deterministic global batch; rank r trains on rows [r*per_rank, (r+1)*per_rank)."""
    return O.synthetic_patches(per_rank * world, 900 + per_rank)


def _train_rank(rank, world, case, per_rank, steps, sync_bn, use_graph):
    import gpu_util as U
    from dynamorph_b200.trainer import FusedTrainer
    st = Golden(case).state()
    m = U.model_from_state(st).to(f"cuda:{rank}").train()
    tr = FusedTrainer(m, lr=LR, use_graph=use_graph, sync_bn=sync_bn)
    assert tr.world == world
    x = _batch(case, per_rank, world)[rank * per_rank:(rank + 1) * per_rank].cuda(rank)
    losses = []
    for _ in range(steps):
        losses.append(tr.step(x).cpu().tolist())
    # numpy, not tensors: a tensor put on a multiprocessing queue travels as a shared-memory handle that dies with the rank
    sd = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    return losses, sd, tr.grad.cpu().numpy().copy()


def _flat(grads, keys):
    return torch.cat([grads[k].reshape(-1) for k in keys])


@pytest.mark.parametrize("case,per_rank,use_graph", [("vqvae_default", 4, True), ("vqvae_default", 4, False),
                                                     ("z32_default", 3, True)])
def test_two_rank_step_per_rank_statistics(case, per_rank, use_graph):
    world = 2
    res = _run(_train_rank, world, case, per_rank, 1, False, use_graph)
    st = Golden(case).state()
    x = _batch(case, per_rank, world)
    keys = O.trainable_keys(st)
    noise = set(O.bias_feeds_train_bn(st))
    per = [O.loss_and_grads(x[r * per_rank:(r + 1) * per_rank], st, O.BATCH) for r in range(world)]
    g_ref = {k: sum(p[2][k] for p in per) / world for k in keys}
    # the allreduced flat gradient buffer = SUM over ranks (1/world lives in the Adam kernel)
    flat_ref = _flat(g_ref, keys) * world
    for r in range(world):
        losses, sd, grad = res[r]
        for i, k in enumerate(("recon_loss", "commitment_loss", "total_loss", "perplexity")):
            ref = float(per[r][1][k])
            assert abs(losses[0][i] - ref) <= 1e-4 * abs(ref), (r, k)
        off = 0
        for k in keys:
            n = g_ref[k].numel()
            if k not in noise:
                got, ref = grad[off:off + n], flat_ref[off:off + n]
                assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), (r, k)
            off += n
        # running statistics are the rank's own (DDP semantics)
        for k, v in per[r][3].items():
            if not k.endswith("num_batches_tracked"):
                assert float((sd[k] - v).abs().max()) <= 1e-4 * float(v.abs().max()), (r, k)
    # one Adam step on the averaged gradient; replicas stay bit-identical
    exp = {k: torch.nn.Parameter(st[k].clone()) for k in keys}
    for k in keys:
        exp[k].grad = g_ref[k].clone()
    torch.optim.Adam(list(exp.values()), lr=LR, betas=(.9, .999)).step()
    for k in keys:
        assert torch.equal(res[0][1][k], res[1][1][k]), k
        diff = (res[0][1][k] - exp[k].detach()).abs()
        assert float(diff.max()) <= 2.01 * LR, k
        if k not in noise:
            sig = g_ref[k].abs() > 1e-2 * g_ref[k].abs().max()
            assert float(diff[sig].max()) <= 5e-3 * LR + 1e-7, (k, float(diff[sig].max()))


@pytest.mark.parametrize("case,per_rank", [("vqvae_default", 4), ("z32_default", 3)])
def test_two_rank_step_synchronised_batchnorm(case, per_rank):
    world = 2
    res = _run(_train_rank, world, case, per_rank, 1, True, True)
    st = Golden(case).state()
    x = _batch(case, per_rank, world)
    keys = O.trainable_keys(st)
    noise = set(O.bias_feeds_train_bn(st))
    _, loss_ref, g_ref, running = O.loss_and_grads(x, st, O.BATCH)          # ONE pass over the global batch
    # recon / commitment are means over equal shards: their rank average is the global value
    for i, k in enumerate(("recon_loss", "commitment_loss", "total_loss")):
        avg = sum(res[r][0][0][i] for r in range(world)) / world
        assert abs(avg - float(loss_ref[k])) <= 1e-4 * abs(float(loss_ref[k])), k
    flat_ref = _flat(g_ref, keys) * world
    for r in range(world):
        _, sd, grad = res[r]
        off = 0
        for k in keys:
            n = g_ref[k].numel()
            if k not in noise:
                got, ref = grad[off:off + n], flat_ref[off:off + n]
                assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), (r, k)
            off += n
        for k, v in running.items():          # running statistics of the GLOBAL batch on every rank
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v)
            else:
                assert float((sd[k] - v).abs().max()) <= 1e-4 * float(v.abs().max()), (r, k)
    for k in keys:
        assert torch.equal(res[0][1][k], res[1][1][k]), k


def _encode_rank(rank, world, n):
    """Bulk-encode sharding: rank r encodes its contiguous patch range; indices gathered at the end."""
    import gpu_util as U
    from dynamorph_b200.dist import gather_code_indices, shard_range
    st = Golden("vqvae_default").state()
    m = U.model_from_state(st).to(f"cuda:{rank}").eval()
    x = O.synthetic_patches(n, 55)
    a, b = shard_range(n, rank, world)
    _, _, idx = m.encode_latents(x[a:b].cuda(rank), "eval")
    return gather_code_indices(idx, n, 64).cpu().numpy().copy()


def test_two_rank_encode_sharding_and_index_gather():
    n = 21                                       # ragged: 11 + 10
    res = _run(_encode_rank, 2, n)
    import gpu_util as U
    st = Golden("vqvae_default").state()
    with torch.no_grad():
        zb = O.encoder(O.synthetic_patches(n, 55), st, O.EVAL)
        ref = O.vq_indices(zb, st["vq.w.weight"])
    assert torch.equal(res[0], res[1]) and res[0].dtype == torch.uint8 and res[0].shape == (n, 16, 16)
    U.check_indices(res[0], zb, st["vq.w.weight"], ref, "sharded encode")
