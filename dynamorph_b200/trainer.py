"""Fused training step: pack -> forward -> backward -> (gradient allreduce) -> Adam as C-ABI calls on flat
buffers, replayed as a CUDA graph (reference step: /root/reference/run_training.py:404-408, Adam :485).

Data parallel: one process per GPU, identical replicas, ONE NCCL allreduce(sum) of the flat gradient
buffer per step (96 KB for the default model), 1/world folded into the Adam kernel.  BatchNorm statistics
are per rank (DDP semantics) unless `sync_bn=True`, which exchanges the per-channel
sums of every BatchNorm so that the step equals a single-process step on the global batch (the reference's
single-GPU semantics, run_training.py:404)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from ._lib import BN_BATCH, call, ptr
from .engine import _require_cuda, _stream


class _Plan:
    """Static buffers + captured graphs of one (batch shape, mask, time-matching) combination."""
    __slots__ = ("spec", "mc", "nws", "x", "mask", "cv", "decoded", "packed", "ws", "tm", "tm_mat", "graphs",
                 "fwd_graph")


class FusedTrainer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, use_graph=True, process_group=None,
                 sync_bn=False):
        self.model = model
        self.eng = model._engine
        self.eng.flatten()
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.sync_bn = bool(sync_bn) and self.world > 1
        flat = self.eng.flat_params
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.grad = torch.zeros_like(flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=flat.device)
        self.bc_dev = torch.zeros(2, dtype=torch.float32, device=flat.device)
        self.losses = torch.zeros(8, dtype=torch.float32, device=flat.device)
        self._plans: Dict[tuple, _Plan] = {}
        self._cur: Optional[_Plan] = None
        if self.sync_bn:
            from .sync_bn import SyncBNStep
            self._sync = SyncBNStep(self)

    # ------------------------------------------------------------------ pieces
    def _forward(self, st: _Plan, update_running: bool = True):
        eng = self.eng
        s, B = st.spec, st.x.shape[0]
        call("dmb_pack_weights", C.byref(s), ptr(eng._flat), ptr(eng._flat_bn), BN_BATCH, ptr(st.packed), _stream())
        tm = C.byref(st.tm) if st.tm is not None else None
        call("dmb_train_forward_tm", C.byref(s), ptr(st.packed), ptr(eng._flat), ptr(st.x), ptr(st.mask), st.mc,
             ptr(st.cv), B, tm, ptr(st.decoded), ptr(self.losses), ptr(eng._flat_bn) if update_running else None,
             ptr(st.ws), st.nws, _stream())
        if update_running:
            eng._count_batch()

    def _backward(self, st: _Plan):
        eng = self.eng
        s, B = st.spec, st.x.shape[0]
        tm = C.byref(st.tm) if st.tm is not None else None
        call("dmb_train_backward_tm", C.byref(s), ptr(st.packed), ptr(eng._flat), ptr(st.x), ptr(st.mask), st.mc,
             ptr(st.cv), ptr(st.decoded), B, tm, 1.0, ptr(self.grad), ptr(st.ws), st.nws, _stream())

    def _fwd_bwd(self, st: _Plan):
        self._forward(st)
        self._backward(st)

    def _adam(self):
        eng = self.eng
        call("dmb_adam_step_dev", ptr(eng._flat), ptr(self.grad), ptr(self.m), ptr(self.v), eng._flat.numel(),
             self.lr, self.betas[0], self.betas[1], self.eps, ptr(self.step_dev), ptr(self.bc_dev),
             1.0 / self.world, _stream())

    def _allreduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)

    def _plan(self, x, mask, tm_mat=None) -> _Plan:
        eng = self.eng
        B, Cin, H, W = x.shape
        mc = 0 if mask is None else mask.shape[1]
        key = (B, Cin, H, W, mc, eng._flat.data_ptr(), tm_mat is not None)
        st = self._plans.get(key)
        if st is not None:
            return st
        eng.flatten()
        s = eng.spec(H, W)
        n = C.c_int64()
        call("dmb_packed_floats", C.byref(s), C.byref(n))
        nbytes = C.c_size_t()
        call("dmb_workspace_bytes", C.byref(s), B, BN_BATCH, 1, C.byref(nbytes))
        dev = x.device
        st = _Plan()
        st.spec, st.mc, st.nws = s, mc, nbytes.value
        st.x = torch.empty_like(x)
        st.mask = None if mask is None else torch.empty_like(mask)
        st.cv = self.model.channel_var.data.reshape(-1).contiguous().clone()
        st.decoded = torch.empty_like(x)
        st.packed = torch.empty(n.value, dtype=torch.float32, device=dev)
        st.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        st.tm, st.tm_mat, st.graphs, st.fwd_graph = None, None, None, None
        if tm_mat is not None:
            from .matching import descriptor
            if tuple(tm_mat.shape) != (B, B):
                raise AssertionError("sim_mat.shape == time_matching_mat.shape")      # vq_vae.py:329
            # static (B, B) buffer the captured graph reads; step() refreshes its contents
            st.tm, st.tm_mat = descriptor(self.model, torch.empty(B, B, device=dev))
        if len(self._plans) >= 8:                 # ragged last batches etc.: a handful of shapes, never unbounded
            self._plans.pop(next(iter(self._plans)))
        self._plans[key] = st
        return st

    def _snapshot(self):
        t = (self.eng._flat, self.eng._flat_bn, self.eng._flat_nbt, self.m, self.v, self.step_dev)
        return t, tuple(a.clone() for a in t)

    def _capture(self, st: _Plan):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # warm-up off-graph (kernel attributes, lazy module loading); restore the state it touched
        live, snap = self._snapshot()
        with torch.cuda.stream(side):
            self._fwd_bwd(st)
            self._adam()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip(live, snap):
            dst.copy_(src)
        # (an explicit capture stream of THIS device: torch.cuda.graph's default one is created once, on whichever device
        # was current first, and a capture through it on another GPU records nothing)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, stream=side):
            self._fwd_bwd(st)
            if self.world == 1:
                self._adam()
        g2 = None
        if self.world > 1:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=side):
                self._adam()
        st.graphs = (g1, g2)

    def _load(self, st: _Plan, x, batch_mask, time_matching_mat):
        st.x.copy_(x)
        if batch_mask is not None:
            st.mask.copy_(batch_mask)
        if time_matching_mat is not None:
            st.tm_mat.copy_(_require_cuda(time_matching_mat, "time_matching_mat"))

    # ------------------------------------------------------------------ public
    def step(self, x: torch.Tensor, batch_mask: Optional[torch.Tensor] = None,
             time_matching_mat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on batch x.  Returns a device tensor
        [recon_loss, commitment_loss, total_loss, perplexity, time_matching_loss, 0, 0, 0] (no host sync)."""
        x = _require_cuda(x, "batch")
        if batch_mask is not None:
            batch_mask = _require_cuda(batch_mask, "batch_mask")
        with torch.cuda.device(x.device):          # graphs, side streams and events belong to the batch's GPU
            return self._step(x, batch_mask, time_matching_mat)

    def _step(self, x, batch_mask, time_matching_mat):
        st = self._cur = self._plan(x, batch_mask, time_matching_mat)
        self._load(st, x, batch_mask, time_matching_mat)
        if self.sync_bn:
            self._sync.step(st)
        elif self.use_graph:
            if st.graphs is None:
                self._capture(st)
            g1, g2 = st.graphs
            g1.replay()
            if g2 is not None:
                self._allreduce()
                g2.replay()
        else:
            self._fwd_bwd(st)
            self._allreduce()
            self._adam()
        self.eng.mark_params_written()
        self.eng._bn_dirty += 1
        return self.losses

    def forward_only(self, x: torch.Tensor, batch_mask: Optional[torch.Tensor] = None,
                     time_matching_mat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """A validation batch as the reference runs it (run_training.py:522-531: `run_one_batch(training=False)` on a
        model left in train mode): train-mode forward -- batch statistics, running statistics updated -- and the five
        losses; no backward, no optimiser step.  Returns the same device tensor as `step`."""
        x = _require_cuda(x, "batch")
        if batch_mask is not None:
            batch_mask = _require_cuda(batch_mask, "batch_mask")
        with torch.cuda.device(x.device):
            st = self._cur = self._plan(x, batch_mask, time_matching_mat)
            self._load(st, x, batch_mask, time_matching_mat)
            if self.use_graph and not self.sync_bn:
                if st.fwd_graph is None:
                    torch.cuda.synchronize()
                    live, snap = self._snapshot()
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        self._forward(st)
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    for dst, src in zip(live, snap):
                        dst.copy_(src)
                    st.fwd_graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(st.fwd_graph, stream=side):
                        self._forward(st)
                st.fwd_graph.replay()
            elif self.sync_bn:
                self._sync.forward(st)
            else:
                self._forward(st)
            self.eng._bn_dirty += 1
        return self.losses

    @property
    def decoded(self) -> torch.Tensor:
        return self._cur.decoded
