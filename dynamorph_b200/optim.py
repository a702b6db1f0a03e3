"""Fused flat Adam: one kernel over the model's flat parameter buffer (torch.optim.Adam semantics as
constructed at /root/reference/run_training.py:485 -- betas (0.9, 0.999), eps 1e-8, no weight decay)."""
from __future__ import annotations

import torch

from ._lib import call, ptr
from .engine import _stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.engine = model._engine
        self.engine.flatten()
        params = [p for p, _, _ in self.engine._views if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._m = None
        self._v = None
        self._step = 0
        self.grad_scale = 1.0      # e.g. 1/world_size after an allreduce(sum)

    def _flat_grad(self) -> torch.Tensor:
        eng = self.engine
        eng.flatten()
        g = eng.last_flat_grad
        ok = g is not None and g.device == eng.flat_params.device
        if ok:
            base = g.data_ptr()
            for p, off, n in eng._views:
                if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                    ok = False
                    break
        if ok:
            return g
        # gradients were accumulated / produced elsewhere: gather them (slow path, still on device)
        g = torch.zeros_like(eng.flat_params)
        for p, off, n in eng._views:
            if p.grad is not None:
                g[off:off + n].copy_(p.grad.reshape(-1))
        return g

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        eng = self.engine
        g = self._flat_grad()
        flat = eng.flat_params
        if self._m is None or self._m.data_ptr() == 0 or self._m.numel() != flat.numel() or self._m.device != flat.device:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
        self._step += 1
        grp = self.param_groups[0]
        b1, b2 = grp["betas"]
        call("dmb_adam_step", ptr(flat), ptr(g), ptr(self._m), ptr(self._v), flat.numel(), float(grp["lr"]),
             float(b1), float(b2), float(grp["eps"]), self._step, float(self.grad_scale), _stream())
        eng.mark_params_written()
        return loss

    def state_dict(self):
        return {"step": self._step, "exp_avg": self._m, "exp_avg_sq": self._v,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self._step = int(sd["step"])
        self._m = sd["exp_avg"].clone() if sd["exp_avg"] is not None else None
        self._v = sd["exp_avg_sq"].clone() if sd["exp_avg_sq"] is not None else None
