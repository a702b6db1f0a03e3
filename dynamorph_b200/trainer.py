"""Fused training step: pack -> forward -> backward -> (gradient allreduce) -> Adam as C-ABI calls on flat
buffers, replayed as a CUDA graph (reference step: /root/reference/run_training.py:404-408, Adam :485).

Data parallel: one process per GPU, identical replicas, ONE NCCL allreduce(sum) of the flat gradient
buffer per step (96 KB for the default model), 1/world folded into the Adam kernel.  BatchNorm statistics
are per rank (DDP semantics)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import BN_BATCH, call, ptr
from .engine import _require_cuda, _stream


class FusedTrainer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, use_graph=True, process_group=None):
        self.model = model
        self.eng = model._engine
        self.eng.flatten()
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        flat = self.eng.flat_params
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.grad = torch.zeros_like(flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=flat.device)
        self.bc_dev = torch.zeros(2, dtype=torch.float32, device=flat.device)
        self.losses = torch.zeros(8, dtype=torch.float32, device=flat.device)
        self._key = None
        self._graphs = None
        self._static = None

    # ------------------------------------------------------------------ pieces
    def _fwd_bwd(self):
        st = self._static
        eng = self.eng
        s, B = st["spec"], st["x"].shape[0]
        call("dmb_pack_weights", C.byref(s), ptr(eng._flat), ptr(eng._flat_bn), BN_BATCH, ptr(st["packed"]), _stream())
        tm = C.byref(st["tm"]) if st["tm"] is not None else None
        call("dmb_train_forward_tm", C.byref(s), ptr(st["packed"]), ptr(eng._flat), ptr(st["x"]), ptr(st["mask"]), st["mc"],
             ptr(st["cv"]), B, tm, ptr(st["decoded"]), ptr(self.losses), ptr(eng._flat_bn), ptr(st["ws"]), st["nws"],
             _stream())
        eng._flat_nbt += 1
        call("dmb_train_backward_tm", C.byref(s), ptr(st["packed"]), ptr(eng._flat), ptr(st["x"]), ptr(st["mask"]), st["mc"],
             ptr(st["cv"]), ptr(st["decoded"]), B, tm, 1.0, ptr(self.grad), ptr(st["ws"]), st["nws"], _stream())

    def _adam(self):
        eng = self.eng
        call("dmb_adam_step_dev", ptr(eng._flat), ptr(self.grad), ptr(self.m), ptr(self.v), eng._flat.numel(),
             self.lr, self.betas[0], self.betas[1], self.eps, ptr(self.step_dev), ptr(self.bc_dev),
             1.0 / self.world, _stream())

    def _allreduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)

    def _prepare(self, x, mask, tm_mat=None):
        eng = self.eng
        B, Cin, H, W = x.shape
        mc = 0 if mask is None else mask.shape[1]
        key = (B, Cin, H, W, mc, eng._flat.data_ptr(), tm_mat is not None)
        if key == self._key:
            return
        eng.flatten()
        s = eng.spec(H, W)
        n = C.c_int64()
        call("dmb_packed_floats", C.byref(s), C.byref(n))
        nbytes = C.c_size_t()
        call("dmb_workspace_bytes", C.byref(s), B, BN_BATCH, 1, C.byref(nbytes))
        dev = x.device
        self._static = {
            "spec": s, "mc": mc, "nws": nbytes.value,
            "x": torch.empty_like(x), "mask": None if mask is None else torch.empty_like(mask),
            "cv": self.model.channel_var.data.reshape(-1).contiguous().clone(),
            "decoded": torch.empty_like(x), "packed": torch.empty(n.value, dtype=torch.float32, device=dev),
            "ws": torch.empty(nbytes.value, dtype=torch.uint8, device=dev),
            "tm": None, "tm_mat": None,
        }
        if tm_mat is not None:
            from .matching import descriptor
            if tuple(tm_mat.shape) != (B, B):
                raise AssertionError("sim_mat.shape == time_matching_mat.shape")      # vq_vae.py:329
            # static (B, B) buffer the captured graph reads; step() refreshes its contents
            self._static["tm"], self._static["tm_mat"] = descriptor(self.model, torch.empty(B, B, device=dev))
        self._key = key
        self._graphs = None

    def _capture(self):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # warm-up off-graph (kernel attributes, lazy module loading); restore the state it touched
        snap = (self.eng._flat.clone(), self.eng._flat_bn.clone(), self.eng._flat_nbt.clone(), self.m.clone(),
                self.v.clone(), self.step_dev.clone())
        with torch.cuda.stream(side):
            self._fwd_bwd()
            self._adam()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for dst, src in zip((self.eng._flat, self.eng._flat_bn, self.eng._flat_nbt, self.m, self.v, self.step_dev), snap):
            dst.copy_(src)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._fwd_bwd()
            if self.world == 1:
                self._adam()
        g2 = None
        if self.world > 1:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                self._adam()
        self._graphs = (g1, g2)

    # ------------------------------------------------------------------ public
    def step(self, x: torch.Tensor, batch_mask: Optional[torch.Tensor] = None,
             time_matching_mat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on batch x.  Returns a device tensor
        [recon_loss, commitment_loss, total_loss, perplexity, time_matching_loss, 0, 0, 0] (no host sync)."""
        x = _require_cuda(x, "batch")
        if batch_mask is not None:
            batch_mask = _require_cuda(batch_mask, "batch_mask")
        self._prepare(x, batch_mask, time_matching_mat)
        st = self._static
        st["x"].copy_(x)
        if batch_mask is not None:
            st["mask"].copy_(batch_mask)
        if time_matching_mat is not None:
            st["tm_mat"].copy_(_require_cuda(time_matching_mat, "time_matching_mat"))
        if self.use_graph:
            if self._graphs is None:
                self._capture()
            g1, g2 = self._graphs
            g1.replay()
            if g2 is not None:
                self._allreduce()
                g2.replay()
        else:
            self._fwd_bwd()
            self._allreduce()
            self._adam()
        self.eng.mark_params_written()
        self.eng._bn_dirty += 1
        return self.losses

    @property
    def decoded(self) -> torch.Tensor:
        return self._static["decoded"]
