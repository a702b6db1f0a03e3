"""Bulk patch encoding from HOST buffers: the body of process_VAE
(/root/reference/pipeline/patch_VAE.py:443-462) as a three-stream pipeline -- pinned-host -> HBM copy
of chunk i+1, encode of chunk i, HBM -> pinned-host copy of chunk i-1 all overlap."""
from __future__ import annotations

from typing import Dict, Optional

import torch


class BulkEncoder:
    def __init__(self, model, chunk: int = 8192, bn_mode: str = "eval", device=None,
                 outputs=("z_before", "z_after", "idx"), zscore: bool = False):
        """zscore=True: `encode` takes RAW patches (float32 / float64 / uint16 on the host), ships them as they are
        and applies pipeline/train_utils.py:252-274 `zscore_patch` on the device in front of the encoder (uint16
        input halves the host->device bytes of the float32 path)."""
        if bn_mode not in ("eval", "per_sample"):
            raise ValueError("bulk encoding needs patch-independent statistics: bn_mode 'eval' or 'per_sample'")
        self.model = model
        self.engine = model._engine
        self.chunk = int(chunk)
        self.bn_mode = bn_mode
        self.device = torch.device(device) if device is not None else self.engine.device
        self.outputs = tuple(outputs)
        self.zscore = bool(zscore)
        self._raw = None
        self._bufs = None
        self._streams = None

    def _setup(self, C, H, W):
        from . import _lib
        import ctypes
        s = self.engine.spec(H, W)
        d, lh, lw = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _lib.call("dmb_latent_shape", ctypes.byref(s), ctypes.byref(d), ctypes.byref(lh), ctypes.byref(lw))
        key = (C, H, W)
        if self._bufs is not None and self._bufs["key"] == key:
            return
        dev, n = self.device, self.chunk
        self._bufs = {
            "key": key, "lat": (d.value, lh.value, lw.value),
            "x": [torch.empty(n, C, H, W, dtype=torch.float32, device=dev) for _ in range(2)],
            "zb": [torch.empty(n, d.value, lh.value, lw.value, dtype=torch.float32, device=dev) for _ in range(2)],
            "za": [torch.empty(n, d.value, lh.value, lw.value, dtype=torch.float32, device=dev) for _ in range(2)],
            "idx": [torch.empty(n, lh.value, lw.value, dtype=torch.int32, device=dev) for _ in range(2)],
        }
        self._streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        # per double-buffer slot: input copied / chunk encoded / outputs drained.  They persist across encode() calls so
        # that the first copies of a call overlap the last encode + drain of the previous one (back-to-back calls would
        # otherwise serialise on a whole-pipeline join: ~4 ms of a 43 ms 16384-patch call)
        self._ev_in = [torch.cuda.Event() for _ in range(2)]
        self._ev_cmp = [torch.cuda.Event() for _ in range(2)]
        self._ev_out = [torch.cuda.Event() for _ in range(2)]
        self._used = [False, False]

    def allocate_outputs(self, n: int, C: int = 2, H: int = 128, W: int = 128, pin: bool = True) -> Dict[str, torch.Tensor]:
        self._setup(C, H, W)
        d, lh, lw = self._bufs["lat"]
        out = {}
        if "z_before" in self.outputs:
            out["z_before"] = torch.empty(n, d * lh * lw, dtype=torch.float32, pin_memory=pin)
        if "z_after" in self.outputs:
            out["z_after"] = torch.empty(n, d * lh * lw, dtype=torch.float32, pin_memory=pin)
        if "idx" in self.outputs:
            out["idx"] = torch.empty(n, lh * lw, dtype=torch.int32, pin_memory=pin)
        return out

    @torch.no_grad()
    def encode(self, x_host: torch.Tensor, out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """x_host: (N, C, H, W) float32 on the host (pinned for full overlap); with zscore=True raw float32 /
        float64 / uint16.  Returns host tensors 'z_before' / 'z_after' (N, D*h*w) NCHW-flattened like
        patch_VAE.py:454-461, and 'idx'.  Asynchronous: the results are complete once the current stream has caught up
        (e.g. torch.cuda.synchronize()); x_host must hold its data when the call is made (the input copies do not wait
        for work queued on the current stream, so that back-to-back calls overlap)."""
        if x_host.is_cuda:
            raise ValueError("BulkEncoder.encode takes host tensors; use model.encode_latents for device data")
        N, C, H, W = x_host.shape
        self._setup(C, H, W)
        if not self.zscore and x_host.dtype != torch.float32:
            raise ValueError("BulkEncoder.encode takes float32 patches (or construct it with zscore=True for raw data)")
        if self.zscore:
            from .pipeline.train_utils import _DTYPES
            if x_host.dtype not in _DTYPES:
                raise ValueError(f"raw patches must be float32, float64 or uint16, not {x_host.dtype}")
            if self._raw is None or self._raw[0].dtype != x_host.dtype or self._raw[0].shape[1:] != (C, H, W):
                torch.cuda.synchronize(self.device)      # the old staging buffers may still be in use on the side streams
                self._raw = [torch.empty(self.chunk, C, H, W, dtype=x_host.dtype, device=self.device) for _ in range(2)]
        if out is None:
            out = self.allocate_outputs(N, C, H, W)
        b = self._bufs
        s_in, s_cmp, s_out = self._streams
        cur = torch.cuda.current_stream(self.device)
        ev_in, ev_cmp, ev_out, used = self._ev_in, self._ev_cmp, self._ev_out, self._used
        self.engine.packed(0 if self.bn_mode == "eval" else 2, H, W)   # pack on the current stream first
        # compute and drain follow the caller's stream (packed weights, anything the caller queued); the input copies only
        # depend on their staging slot being free (ev_cmp of its previous use, possibly from the previous call)
        s_cmp.wait_stream(cur)
        s_out.wait_stream(cur)
        for i, a in enumerate(range(0, N, self.chunk)):
            j = i & 1
            e = min(N, a + self.chunk)
            n = e - a
            with torch.cuda.stream(s_in):
                if used[j]:
                    s_in.wait_event(ev_cmp[j])          # x[j] consumed by the encode two chunks ago
                (self._raw if self.zscore else b["x"])[j][:n].copy_(x_host[a:e], non_blocking=True)
                ev_in[j].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[j])
                if used[j]:
                    s_cmp.wait_event(ev_out[j])         # outputs[j] drained to the host
                if self.zscore:
                    import ctypes
                    from ._lib import call, ptr
                    from .pipeline.train_utils import _DTYPES
                    call("dmb_zscore_patch", ptr(self._raw[j]), _DTYPES[x_host.dtype], n * C, H * W, ptr(b["x"][j]),
                         ctypes.c_void_p(s_cmp.cuda_stream))
                self.engine.encode(b["x"][j][:n], self.bn_mode, out=(b["zb"][j], b["za"][j], b["idx"][j]))
                ev_cmp[j].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[j])
                if "z_before" in out:
                    out["z_before"][a:e].copy_(b["zb"][j][:n].reshape(n, -1), non_blocking=True)
                if "z_after" in out:
                    out["z_after"][a:e].copy_(b["za"][j][:n].reshape(n, -1), non_blocking=True)
                if "idx" in out:
                    out["idx"][a:e].copy_(b["idx"][j][:n].reshape(n, -1), non_blocking=True)
                ev_out[j].record(s_out)
            used[j] = True
        for s in self._streams:
            cur.wait_stream(s)
        return out

    @torch.no_grad()
    def encode_stream(self, source, sink) -> int:
        """Encode an arbitrarily long run without ever holding its outputs: `source` yields host tensors
        (n_i, C, H, W) (slices of a memory-mapped array, successive pickles, ...), `sink(first_row, out)` receives the
        host result of each block (`out` as returned by `encode`; valid only during the call) -- e.g.
        `latent_shards.ShardedLatentWriter.append`.  Two pinned output sets alternate, so the sink of block i (disk)
        runs while block i + 1 is copied in and encoded.  Returns the number of rows encoded."""
        bufs = [None, None]
        pending = None          # (first_row, n, out dict, completion event)
        row = 0
        for i, block in enumerate(source):
            j = i & 1
            n, C, H, W = block.shape
            if bufs[j] is None or bufs[j]["_n"] < n:
                bufs[j] = self.allocate_outputs(n, C, H, W)
                bufs[j]["_n"] = n
            out = {k: v[:n] for k, v in bufs[j].items() if k != "_n"}
            self.encode(block, out)                                   # enqueues; returns before the GPU is done
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            if pending is not None:
                pending[3].synchronize()
                sink(pending[0], pending[2])
            pending = (row, n, out, ev)
            row += n
        if pending is not None:
            pending[3].synchronize()
            sink(pending[0], pending[2])
        return row

