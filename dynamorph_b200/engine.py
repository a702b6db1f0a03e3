"""Host-side engine: owns the flat parameter buffers, the packed-weight cache and the scratch
workspace of one VQ-VAE module, and turns the reference's module calls (`model.enc(x)`,
`model.vq(z)`, `model.dec(z)`, `model(x)`) into calls of the C ABI (include/dynamorph_b200.h).

PyTorch is used for device memory, streams and autograd plumbing only; every arithmetic
operation on the path runs in libdynamorph_b200.so.  There is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import BN_BATCH, BN_EVAL, BN_PER_SAMPLE, BN_MODES, DmbModel, call, ptr


def _stream():
    """The stream argument of a C-ABI call: the current stream of the GPU that owns the call's pointers
    (resolved by `_lib.call`, which also makes that GPU current for the call)."""
    return _lib.STREAM


def _require_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"dynamorph_b200: {what} must be a CUDA tensor (there is no CPU path); got {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"dynamorph_b200: {what} must be float32, got {t.dtype}")
    return t.contiguous()


class Engine:
    """One per VQ-VAE module instance."""

    def __init__(self, module: torch.nn.Module, arch: int):
        self.module = module
        self.arch = arch
        self._flat: Optional[torch.Tensor] = None        # trainable parameters, state_dict order
        self._flat_bn: Optional[torch.Tensor] = None     # running_mean / running_var pairs
        self._flat_nbt: Optional[torch.Tensor] = None    # num_batches_tracked, one int64 per BN
        self._views: List[Tuple[torch.nn.Parameter, int, int]] = []
        self._packed: Dict[int, Tuple[tuple, torch.Tensor]] = {}
        self._ws: Optional[torch.Tensor] = None
        self._flat_dirty = 0                              # bumped when the flat buffer is written directly
        self._ws_uses = 0
        self.last_flat_grad = None
        self._bn_dirty = 0                                # bumped when the library updates running stats
        self._spec_cache: Dict[Tuple[int, int], DmbModel] = {}

    # ---------------------------------------------------------------- description
    def spec(self, height: int, width: int) -> DmbModel:
        key = (height, width)
        s = self._spec_cache.get(key)
        m = self.module
        if s is None:
            s = DmbModel()
            s.arch = self.arch
            s.num_inputs = m.num_inputs
            s.num_hiddens = m.num_hiddens
            s.num_residual_hiddens = m.num_residual_hiddens
            s.num_residual_layers = m.num_residual_layers
            s.num_embeddings = m.num_embeddings
            s.height, s.width = height, width
            s.bn_eps = 1e-5
            self._spec_cache[key] = s
        s.bn_momentum = float(getattr(m, "_bn_momentum", 0.1))
        s.commitment_cost = float(m.commitment_cost)
        s.weight_recon = float(getattr(m, "weight_recon", 1.0))
        s.weight_commitment = float(getattr(m, "weight_commitment", 1.0))
        return s

    def trainable(self) -> List[Tuple[str, torch.nn.Parameter]]:
        return [(k, p) for k, p in self.module.named_parameters() if k != "channel_var"]

    def bn_modules(self) -> List[torch.nn.BatchNorm2d]:
        return [m for m in self.module.modules() if isinstance(m, torch.nn.BatchNorm2d)]

    @property
    def device(self) -> torch.device:
        return self.module.channel_var.device

    # ---------------------------------------------------------------- flat buffers
    def flatten(self) -> None:
        """Make every parameter / BN buffer a view into one flat device buffer (state_dict order),
        which is what the C ABI's `params` / `bnbuf` arguments are.  Re-done lazily after
        `.to(device)` replaces the tensors."""
        params = self.trainable()
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("dynamorph_b200: the model must be on a CUDA device (model.to('cuda:N')); "
                               "there is no CPU path")
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            base = self._flat.data_ptr()
            for (_, p), (_, off, _) in zip(params, self._views):
                if p.data_ptr() != base + 4 * off:
                    ok = False
                    break
        if ok:
            bns = self.bn_modules()
            base = self._flat_bn.data_ptr() if self._flat_bn is not None and self._flat_bn.numel() else 0
            off = 0
            for bn in bns:
                c = bn.num_features
                if bn.running_mean.data_ptr() != base + 4 * off or bn.running_var.data_ptr() != base + 4 * (off + c):
                    ok = False
                    break
                off += 2 * c
        if ok:
            return
        n = sum(p.numel() for _, p in params)
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        views = []
        off = 0
        with torch.no_grad():
            for _, p in params:
                k = p.numel()
                flat[off:off + k].copy_(p.detach().reshape(-1).to(dev, torch.float32))
                p.data = flat[off:off + k].view(p.shape)
                views.append((p, off, k))
                off += k
            bns = self.bn_modules()
            nb = sum(2 * b.num_features for b in bns)
            fbn = torch.empty(nb, dtype=torch.float32, device=dev)
            nbt = torch.zeros(max(len(bns), 1), dtype=torch.int64, device=dev)
            off = 0
            for i, b in enumerate(bns):
                c = b.num_features
                fbn[off:off + c].copy_(b.running_mean.to(dev))
                fbn[off + c:off + 2 * c].copy_(b.running_var.to(dev))
                b.running_mean = fbn[off:off + c]
                b.running_var = fbn[off + c:off + 2 * c]
                nbt[i] = b.num_batches_tracked.to(dev)
                b.num_batches_tracked = nbt[i]
                off += 2 * c
        self._flat, self._flat_bn, self._flat_nbt, self._views = flat, fbn, nbt, views
        self._packed.clear()
        # the C side must agree on the layout
        s = self.spec(128, 128)
        n_p, n_b, n_bn = C.c_int64(), C.c_int64(), C.c_int32()
        call("dmb_param_count", C.byref(s), C.byref(n_p), C.byref(n_b), C.byref(n_bn))
        if n_p.value != n or n_b.value != nb or n_bn.value != len(bns):
            raise RuntimeError(f"parameter layout mismatch: python {n}/{nb}/{len(bns)} vs library "
                               f"{n_p.value}/{n_b.value}/{n_bn.value}")

    @property
    def flat_params(self) -> torch.Tensor:
        self.flatten()
        return self._flat

    @property
    def flat_bn(self) -> torch.Tensor:
        self.flatten()
        return self._flat_bn

    def codebook(self) -> torch.Tensor:
        self.flatten()
        return self.module.vq.w.weight.data

    def mark_params_written(self) -> None:
        self._flat_dirty += 1

    # ---------------------------------------------------------------- packed weights
    def packed(self, mode: int, height: int, width: int) -> torch.Tensor:
        self.flatten()
        key = (self._flat_dirty, tuple(p._version for p, _, _ in self._views))
        if mode == BN_EVAL:
            key = key + (self._bn_dirty,) + tuple(b.running_mean._version + b.running_var._version
                                                   for b in self.bn_modules())
        hit = self._packed.get(mode)
        if hit is not None and hit[0] == key:
            return hit[1]
        s = self.spec(height, width)
        n = C.c_int64()
        call("dmb_packed_floats", C.byref(s), C.byref(n))
        buf = hit[1] if hit is not None and hit[1].numel() == n.value else \
            torch.empty(n.value, dtype=torch.float32, device=self.device)
        call("dmb_pack_weights", C.byref(s), ptr(self._flat), ptr(self._flat_bn), mode, ptr(buf), _stream())
        self._packed[mode] = (key, buf)
        return buf

    # ---------------------------------------------------------------- workspace
    def workspace_token(self) -> int:
        """Changes whenever a call may have overwritten the activation workspace."""
        return self._ws_uses

    def workspace(self, s: DmbModel, batch: int, mode: int, keep: int) -> Tuple[torch.Tensor, int]:
        self._ws_uses += 1
        nbytes = C.c_size_t()
        call("dmb_workspace_bytes", C.byref(s), batch, mode, keep, C.byref(nbytes))
        need = nbytes.value
        if self._ws is None or self._ws.numel() < need or self._ws.device != self.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    def bn_mode(self, batch: int, override: Optional[str] = None) -> int:
        if override is not None:
            return BN_MODES[override]
        return BN_BATCH if self.module.training else BN_EVAL

    def _count_batch(self):
        """num_batches_tracked += 1 of every BatchNorm (one small launch of the library on the current stream, so that a
        captured training step holds no ATen kernel)."""
        t = self._flat_nbt
        if t is None or t.numel() == 0:
            return
        if t.is_cuda:
            call("dmb_bn_count_batch", ptr(t), t.numel(), _stream())
        else:
            t += 1

    def _bump_num_batches_tracked(self):
        self._bn_dirty += 1
        if self._flat_nbt is not None and len(self.bn_modules()):
            self._count_batch()

    # ---------------------------------------------------------------- ops
    def encoder_forward(self, x: torch.Tensor, bn_mode: Optional[str] = None) -> torch.Tensor:
        x = _require_cuda(x, "input")
        B, Cin, H, W = x.shape
        s = self.spec(H, W)
        mode = self.bn_mode(B, bn_mode)
        packed = self.packed(mode, H, W)
        d, lh, lw = C.c_int32(), C.c_int32(), C.c_int32()
        call("dmb_latent_shape", C.byref(s), C.byref(d), C.byref(lh), C.byref(lw))
        if Cin != s.num_inputs:
            raise RuntimeError(f"expected {s.num_inputs} input channels, got {Cin}")
        zb = torch.empty(B, d.value, lh.value, lw.value, dtype=torch.float32, device=x.device)
        if B == 0:                       # empty batch: empty latents, like the reference's modules
            return zb
        ws, n = self.workspace(s, B, mode, 0)
        upd = ptr(self._flat_bn) if mode == BN_BATCH else None
        call("dmb_encoder_forward", C.byref(s), ptr(packed), ptr(x), B, mode, ptr(zb), upd, ptr(ws), n, _stream())
        if mode == BN_BATCH:
            self._bump_num_batches_tracked()
        return zb

    def encode(self, x: torch.Tensor, bn_mode: str = "eval", want_stats: bool = False, out=None):
        """enc + vq fused (the process_VAE hot call).  Returns z_before, z_after, idx[, (loss, perplexity)].
        `out` = (z_before, z_after, idx) preallocated device tensors to write into."""
        x = _require_cuda(x, "input")
        B, Cin, H, W = x.shape
        s = self.spec(H, W)
        mode = BN_MODES[bn_mode]
        packed = self.packed(mode, H, W)
        d, lh, lw = C.c_int32(), C.c_int32(), C.c_int32()
        call("dmb_latent_shape", C.byref(s), C.byref(d), C.byref(lh), C.byref(lw))
        if out is not None:
            zb, za, idx = out
            for t_, nel in ((zb, B * d.value * lh.value * lw.value), (za, B * d.value * lh.value * lw.value),
                            (idx, B * lh.value * lw.value)):
                if t_ is not None and (not t_.is_cuda or not t_.is_contiguous() or t_.numel() < nel):
                    raise RuntimeError("encode(out=...): buffers must be contiguous CUDA tensors of sufficient size")
        else:
            zb = torch.empty(B, d.value, lh.value, lw.value, dtype=torch.float32, device=x.device)
            za = torch.empty_like(zb)
            idx = torch.empty(B, lh.value, lw.value, dtype=torch.int32, device=x.device)
        if B == 0 and not want_stats:
            return zb[:0], za[:0], idx[:0]
        if B == 0:
            raise RuntimeError("encode(want_stats=True) needs at least one patch")
        ws, n = self.workspace(s, B, mode, 0)
        stats = None
        if want_stats:
            stats = torch.zeros(2 + s.num_embeddings, dtype=torch.float64, device=x.device)
        call("dmb_encode", C.byref(s), ptr(packed), ptr(self.codebook()), ptr(x), B, mode, ptr(zb), ptr(za),
             ptr(idx), ptr(stats), ptr(ws), n, _stream())
        if want_stats:
            out2 = torch.empty(2, dtype=torch.float32, device=x.device)
            call("dmb_vq_finalize", ptr(stats), d.value, s.num_embeddings, float(s.commitment_cost), ptr(out2), _stream())
            return zb, za, idx, (out2[0], out2[1])
        return zb, za, idx

    def decoder_forward(self, z: torch.Tensor, bn_mode: Optional[str] = None) -> torch.Tensor:
        z = _require_cuda(z, "latent")
        B, D, lh, lw = z.shape
        up = 8 if self.arch == _lib.ARCH_Z16 else 4
        H, W = lh * up, lw * up
        s = self.spec(H, W)
        mode = self.bn_mode(B, bn_mode)
        packed = self.packed(mode, H, W)
        out = torch.empty(B, s.num_inputs, H, W, dtype=torch.float32, device=z.device)
        if B == 0:
            return out
        ws, n = self.workspace(s, B, mode, 1)
        upd = ptr(self._flat_bn) if mode == BN_BATCH else None
        call("dmb_decoder_forward", C.byref(s), ptr(packed), ptr(z), B, mode, ptr(out), upd, ptr(ws), n, _stream())
        if mode == BN_BATCH and self.arch == _lib.ARCH_Z32:
            pass  # num_batches_tracked of decoder BNs is bumped by the full-model call
        return out


# ---------------------------------------------------------------- stand-alone quantiser ops
def vq_forward(z: torch.Tensor, codebook: torch.Tensor, commitment_cost: float,
               want_indices: bool = False):
    z = _require_cuda(z, "VectorQuantizer input")
    cb = _require_cuda(codebook, "codebook")
    B, D, H, W = z.shape
    K = cb.shape[0]
    z_st = torch.empty_like(z)
    idx = torch.empty(B, H, W, dtype=torch.int32, device=z.device)
    stats = torch.zeros(2 + K, dtype=torch.float64, device=z.device)
    call("dmb_vq_forward", ptr(z), ptr(cb), B, D, H * W, K, ptr(z_st), ptr(idx), ptr(stats), _stream())
    out2 = torch.empty(2, dtype=torch.float32, device=z.device)
    call("dmb_vq_finalize", ptr(stats), D, K, float(commitment_cost), ptr(out2), _stream())
    if want_indices:
        return z_st, out2[0], out2[1], idx
    return z_st, out2[0], out2[1]


def vq_indices(z: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    z = _require_cuda(z, "VectorQuantizer input")
    cb = _require_cuda(codebook, "codebook")
    B, D, H, W = z.shape
    idx = torch.empty(B, H, W, dtype=torch.int32, device=z.device)
    call("dmb_vq_forward", ptr(z), ptr(cb), B, D, H * W, cb.shape[0], None, ptr(idx), None, _stream())
    return idx


def vq_gather(idx: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    cb = _require_cuda(codebook, "codebook")
    if not idx.is_cuda:
        raise RuntimeError("dynamorph_b200: indices must be a CUDA tensor")
    B, H, W = idx.shape
    i32 = idx.to(torch.int32).contiguous()
    q = torch.empty(B, cb.shape[1], H, W, dtype=torch.float32, device=cb.device)
    call("dmb_vq_gather", ptr(i32), ptr(cb), B, cb.shape[1], H * W, cb.shape[0], ptr(q), _stream())
    return q
