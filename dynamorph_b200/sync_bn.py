"""Synchronised BatchNorm for data-parallel training (optional `sync_bn=True` of
`trainer.FusedTrainer`): the reference trains on ONE GPU, so its BatchNorm statistics span the whole batch
(/root/reference/run_training.py:404).  With `world` ranks holding equal shards, the library folds every BatchNorm's
per-channel sums into 2*C doubles and calls back here to sum them across the ranks (dmb_train_forward_sync /
dmb_train_backward_sync, include/dynamorph_b200.h); one flat-gradient allreduce closes the step as usual.  The result
is the single-process step on the concatenated batch (tests/test_gpu_dist.py).

8 + 8 tiny allreduces per step for the default model (latency-bound, ~0.4 ms in total): this is the exact mode, the
per-rank-statistics graph replay stays the fast one.  The step runs eagerly -- the callbacks re-enter Python between
kernel launches."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import BN_BATCH, DmbSyncBN, SYNC_ALLREDUCE_FN, call, ptr
from .engine import _stream


class SyncBNStep:
    def __init__(self, trainer):
        self.tr = trainer
        self._ws = None
        self._error = None

        def allreduce(_user, sums_dev, n, stream):
            try:
                ws = self._ws
                off = sums_dev - ws.data_ptr()
                if off < 0 or off + 8 * n > ws.numel() or off % 8:
                    raise RuntimeError("BatchNorm sums lie outside the step's workspace")
                cur = torch.cuda.current_stream(ws.device).cuda_stream
                if (stream or 0) != cur:
                    raise RuntimeError("the library is not launching on torch's current stream")
                dist.all_reduce(ws[off:off + 8 * n].view(torch.float64), group=self.tr.pg)
                return 0
            except Exception as e:              # exceptions must not cross the C ABI
                self._error = e
                return -1

        self._cb = SYNC_ALLREDUCE_FN(allreduce)          # keep the trampoline alive
        self.desc = DmbSyncBN(self._cb, None, self.tr.world)

    def _call(self, name, *args):
        self._error = None
        try:
            call(name, *args)
        except Exception:
            if self._error is not None:
                raise self._error
            raise

    def forward(self, st, update_running=True):
        tr, eng = self.tr, self.tr.eng
        s, B = st.spec, st.x.shape[0]
        self._ws = st.ws
        call("dmb_pack_weights", C.byref(s), ptr(eng._flat), ptr(eng._flat_bn), BN_BATCH, ptr(st.packed), _stream())
        tm = C.byref(st.tm) if st.tm is not None else None
        self._call("dmb_train_forward_sync", C.byref(s), ptr(st.packed), ptr(eng._flat), ptr(st.x), ptr(st.mask), st.mc,
                   ptr(st.cv), B, tm, C.byref(self.desc), ptr(st.decoded), ptr(tr.losses),
                   ptr(eng._flat_bn) if update_running else None, ptr(st.ws), st.nws, _stream())
        if update_running:
            eng._count_batch()

    def backward(self, st):
        tr, eng = self.tr, self.tr.eng
        s, B = st.spec, st.x.shape[0]
        tm = C.byref(st.tm) if st.tm is not None else None
        self._call("dmb_train_backward_sync", C.byref(s), ptr(st.packed), ptr(eng._flat), ptr(st.x), ptr(st.mask),
                   st.mc, ptr(st.cv), ptr(st.decoded), B, tm, C.byref(self.desc), 1.0, ptr(tr.grad), ptr(st.ws), st.nws,
                   _stream())

    def step(self, st):
        self.forward(st)
        self.backward(st)
        self.tr._allreduce()
        self.tr._adam()
