"""Multi-GPU plumbing (one process per GPU, torch.distributed): patch-range sharding for bulk encoding --
patches are independent in 'eval' / 'per_sample' BatchNorm modes, so there is NO data-path collective, only an
optional final gather of the compact code indices -- and the single flat-gradient allreduce of data-parallel
training.  Nothing here touches CUDA directly; it runs under gloo on CPU for tests."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced range [start, stop) of rank `rank`: the first n % world ranks hold one extra patch."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def index_dtype(num_embeddings: int) -> torch.dtype:
    """uint8 for K <= 256, else int16 (K <= 32768): 256-512 B per 16x16 patch instead of 2 KB of int64."""
    if num_embeddings <= 256:
        return torch.uint8
    if num_embeddings <= 32768:
        return torch.int16
    return torch.int32


def pack_indices(idx: torch.Tensor, num_embeddings: int) -> torch.Tensor:
    return idx.to(index_dtype(num_embeddings))


def gather_code_indices(idx_local: torch.Tensor, n_total: int, num_embeddings: int,
                        group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks contribute the indices of their shard (n_local, h, w); every rank returns (n_total, h, w) in
    global patch order.  Ragged shards are padded to the largest one for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    packed = pack_indices(idx_local, num_embeddings).contiguous()
    if world == 1:
        return packed
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    a, b = sizes[rank]
    if packed.shape[0] != b - a:
        raise ValueError(f"rank {rank} holds {packed.shape[0]} patches, expected {b - a}")
    pad = torch.zeros((biggest,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    pad[:b - a] = packed
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:sb - sa] for o, (sa, sb) in zip(out, sizes)], 0)


def allreduce_flat(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None, average: bool = False) -> torch.Tensor:
    """In-place sum (or mean) of one flat buffer across ranks: THE collective of a data-parallel step."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(dist.get_world_size(group))
    return flat


def broadcast_state(model, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Make every replica start from rank `src`'s parameters and BatchNorm buffers (two flat broadcasts)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    eng = model._engine
    eng.flatten()
    dist.broadcast(eng._flat, src, group=group)
    dist.broadcast(eng._flat_bn, src, group=group)
    dist.broadcast(eng._flat_nbt, src, group=group)
    eng.mark_params_written()


class ShardedEncoder:
    """process_VAE-equivalent bulk encoding of N patches over all ranks: rank r encodes shard_range(N, r, world)
    with its own BulkEncoder; latents stay sharded (4 M patches x 2 x 16 KB do not fit one host), indices can be
    gathered."""

    def __init__(self, model, chunk: int = 8192, bn_mode: str = "eval", group=None):
        from .bulk import BulkEncoder
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.model = model
        self.enc = BulkEncoder(model, chunk=chunk, bn_mode=bn_mode)

    def encode(self, x_host_all: torch.Tensor, gather_indices: bool = False):
        n = x_host_all.shape[0]
        a, b = shard_range(n, self.rank, self.world)
        out = self.enc.encode(x_host_all[a:b])
        torch.cuda.synchronize()
        out["range"] = (a, b)
        if gather_indices:
            lat = out["idx"].shape[1]
            side = int(round(lat ** 0.5))
            dev = self.model._engine.device
            idx = out["idx"].to(dev).view(b - a, side, lat // side)
            out["idx_all"] = gather_code_indices(idx, n, self.model.num_embeddings, self.group)
        return out


def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index`, BEFORE it allocates
    pinned host buffers: first-touch then places the staging memory of the bulk encoder on the NUMA node next to that
    GPU's PCIe root, so 4-8 ranks on one box do not pull their host->device traffic across the socket interconnect.
    Returns the core list, or None when NVML / affinity control is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        h = None
        if bus is not None:
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                    h = hi
                    break
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
