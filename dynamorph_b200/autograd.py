"""torch.autograd glue: the forward/backward of each Function is a C-ABI call; autograd only
routes gradients between them."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import call, ptr
from .engine import _require_cuda, _stream


class VQFunction(torch.autograd.Function):
    """VectorQuantizer.forward with the reference's gradients (straight-through + both MSE terms)."""

    @staticmethod
    def forward(ctx, inputs, codebook, commitment_cost):
        z = _require_cuda(inputs.detach(), "VectorQuantizer input")
        cb = _require_cuda(codebook.detach(), "codebook")
        B, D, H, W = z.shape
        K = cb.shape[0]
        z_st = torch.empty_like(z)
        idx = torch.empty(B, H, W, dtype=torch.int32, device=z.device)
        stats = torch.zeros(2 + K, dtype=torch.float64, device=z.device)
        call("dmb_vq_forward", ptr(z), ptr(cb), B, D, H * W, K, ptr(z_st), ptr(idx), ptr(stats), _stream())
        out2 = torch.empty(2, dtype=torch.float32, device=z.device)
        call("dmb_vq_finalize", ptr(stats), D, K, float(commitment_cost), ptr(out2), _stream())
        ctx.save_for_backward(z, cb, idx)
        ctx.beta = float(commitment_cost)
        loss, ppl = out2[0], out2[1]
        ctx.mark_non_differentiable(ppl)
        return z_st, loss, ppl

    @staticmethod
    def backward(ctx, g_zst, g_loss, _g_ppl):
        z, cb, idx = ctx.saved_tensors
        B, D, H, W = z.shape
        K = cb.shape[0]
        gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        gcb = torch.empty_like(cb) if ctx.needs_input_grad[1] else None
        gzst = g_zst.contiguous() if g_zst is not None else None
        gl = g_loss.contiguous().float() if g_loss is not None else None
        scale = 1.0 if gl is not None else 0.0
        call("dmb_vq_backward", ptr(z), ptr(cb), ptr(idx), ptr(gzst), ptr(gl), scale, ctx.beta,
             B, D, H * W, K, ptr(gz), ptr(gcb), _stream())
        return gz, gcb, None



# ---------------------------------------------------------------------------------------------------------------------
# Graph replay behind the eager API.  `model(batch)` + `total_loss.backward()` is what the reference's run_one_batch does
# every step (run_training.py:404-406); the library's forward and backward are ~80 launches each way, issued from the
# host every call.  From the second call with the same batch geometry on, both are CUDA-graph replays over static buffers
# (the plan machinery of trainer.FusedTrainer): the inputs are copied in, the outputs cloned out, autograd still routes the
# gradient of total_loss.  DMB_EAGER_GRAPH=0 keeps every call eager.
# ---------------------------------------------------------------------------------------------------------------------
class _EagerCache:
    """Per-model cache of captured plans (lives in the module's __dict__, so it dies with the model); a pickled or
    deep-copied model gets an EMPTY one -- CUDA graphs and their static buffers are never copied."""

    def __init__(self):
        self.seen = {}          # geometry key -> None (seen once, not captured yet) | captured plan
        self.tr = None          # the hidden FusedTrainer that owns the static buffers

    def __deepcopy__(self, memo):
        return _EagerCache()

    def __reduce__(self):
        return (_EagerCache, ())


def _eager_graphs_enabled() -> bool:
    import os
    return os.environ.get("DMB_EAGER_GRAPH", "1") != "0"


def _eager_plan(model, x, mask, time_matching_mat):
    """The captured (forward graph, backward graph) pair for this batch geometry, or None (first call of a geometry,
    capture in progress elsewhere, disabled)."""
    if not _eager_graphs_enabled() or torch.cuda.is_current_stream_capturing():
        return None
    eng = model._engine
    key = (tuple(x.shape), None if mask is None else tuple(mask.shape), time_matching_mat is not None,
           eng._flat.data_ptr(), x.device.index)
    cache = model.__dict__.get("_eager_cache")
    if cache is None:
        cache = model.__dict__["_eager_cache"] = _EagerCache()
    seen = cache.seen
    if key not in seen:
        if len(seen) >= 8:
            seen.pop(next(iter(seen)))
        seen[key] = None                      # first call of this geometry runs eagerly (lazy initialisation, warm-up)
        return None
    if seen[key] is None:
        from .trainer import FusedTrainer
        tr = cache.tr
        if tr is None or tr.eng is not eng:
            tr = cache.tr = FusedTrainer(model, lr=0.0, use_graph=False)
            tr.world, tr.sync_bn = 1, False
        with torch.cuda.device(x.device):
            st = tr._plan(x, mask, time_matching_mat)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf, stream=side):
                tr._forward(st)
            with torch.cuda.graph(gb, stream=side):
                tr._backward(st)
            # (capture records, it does not run: the counters / flags _forward touched on the host are put back)
        seen[key] = {"tr": tr, "st": st, "gf": gf, "gb": gb, "n": 0}
    return seen[key]


class TrainStepFunction(torch.autograd.Function):
    """Whole-model forward (train-mode BatchNorm) with the gradient of `total_loss` w.r.t. every
    trainable tensor computed by the fused backward schedule (dmb_train_forward / dmb_train_backward)."""

    @staticmethod
    def forward(ctx, model, inputs, batch_mask, time_matching_mat, *params):
        import ctypes as C
        from ._lib import BN_BATCH
        eng = model._engine
        x = _require_cuda(inputs.detach(), "inputs")
        B, Cin, H, W = x.shape
        if batch_mask is not None:
            _m = _require_cuda(batch_mask.detach(), "batch_mask")
            if _m.shape[0] != B or tuple(_m.shape[2:]) != (H, W) or _m.shape[1] not in (1, Cin):
                batch_mask = _m.expand(B, Cin, H, W).contiguous()
        if time_matching_mat is not None and tuple(time_matching_mat.shape) != (B, B):
            raise AssertionError("sim_mat.shape == time_matching_mat.shape")      # vq_vae.py:329
        gp = _eager_plan(model, x, None if batch_mask is None else batch_mask.detach(), time_matching_mat)
        if gp is not None:
            tr, st = gp["tr"], gp["st"]
            with torch.cuda.device(x.device):
                tr._load(st, x, None if batch_mask is None else batch_mask.detach(), time_matching_mat)
                gp["gf"].replay()
            gp["n"] += 1
            eng._bn_dirty += 1
            ctx.graph_plan, ctx.graph_n = gp, gp["n"]
            ctx.model = model
            ctx.set_materialize_grads(False)
            decoded = st.decoded.clone()
            losses = tr.losses.clone()
            ctx.mark_non_differentiable(decoded)
            recon, commit, total, ppl, tml = losses[0], losses[1], losses[2], losses[3], losses[4]
            ctx.mark_non_differentiable(ppl)
            return decoded, recon, commit, total, ppl, tml
        ctx.graph_plan = None
        s = eng.spec(H, W)
        packed = eng.packed(BN_BATCH, H, W)
        ws, n = eng.workspace(s, B, BN_BATCH, 1)
        mask, mc = None, 0
        if batch_mask is not None:
            mask = _require_cuda(batch_mask.detach(), "batch_mask")
            mc = mask.shape[1]
            if mask.shape[0] != B or tuple(mask.shape[2:]) != (H, W) or mc not in (1, Cin):
                mask = mask.expand(B, Cin, H, W).contiguous()
                mc = Cin
        cv = model.channel_var.data.reshape(-1).contiguous()
        decoded = torch.empty_like(x)
        losses = torch.empty(8, dtype=torch.float32, device=x.device)
        tm, tm_mat = None, None
        if time_matching_mat is not None:
            from .matching import descriptor
            if tuple(time_matching_mat.shape) != (B, B):
                raise AssertionError("sim_mat.shape == time_matching_mat.shape")      # vq_vae.py:329
            tm, tm_mat = descriptor(model, time_matching_mat)
        call("dmb_train_forward_tm", C.byref(s), ptr(packed), ptr(eng.flat_params), ptr(x), ptr(mask), mc, ptr(cv), B,
             C.byref(tm) if tm is not None else None, ptr(decoded), ptr(losses), ptr(eng.flat_bn), ptr(ws), n, _stream())
        ctx.tm, ctx.tm_mat = tm, tm_mat
        eng._bump_num_batches_tracked()
        ctx.model, ctx.spec, ctx.mc, ctx.nws = model, s, mc, n
        ctx.save_for_backward(x, mask if mask is not None else torch.empty(0, device=x.device), cv, decoded, packed)
        ctx.ws_token = eng.workspace_token()
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(decoded)
        recon, commit, total, ppl, tml = losses[0], losses[1], losses[2], losses[3], losses[4]
        ctx.mark_non_differentiable(ppl)
        return decoded, recon, commit, total, ppl, tml

    @staticmethod
    def backward(ctx, g_dec, g_recon, g_commit, g_total, g_ppl, g_tm=None):
        import ctypes as C
        model = ctx.model
        eng = model._engine
        if g_recon is not None or g_commit is not None or g_tm is not None:
            raise NotImplementedError("dynamorph_b200: back-propagate through `total_loss` (recon_loss / "
                                      "commitment_loss are reported values of the fused step)")
        if g_total is None:
            return (None,) * (4 + len(eng.trainable()))
        if ctx.graph_plan is not None:
            gp = ctx.graph_plan
            if gp["n"] != ctx.graph_n:
                raise RuntimeError("dynamorph_b200: the activation workspace was reused by another call before "
                                   "backward(); call backward() right after the forward of the same batch")
            with torch.cuda.device(gp["st"].x.device):
                gp["gb"].replay()
            flat_g = gp["tr"].grad.clone()
            flat_g.mul_(g_total)
            eng.last_flat_grad = flat_g
            grads = tuple(flat_g[off:off + n].view(p.shape) if p.requires_grad else None for p, off, n in eng._views)
            return (None, None, None, None) + grads
        if eng.workspace_token() != ctx.ws_token:
            raise RuntimeError("dynamorph_b200: the activation workspace was reused by another call before "
                               "backward(); call backward() right after the forward of the same batch")
        x, mask, cv, decoded, packed = ctx.saved_tensors
        mask_t = mask if mask.numel() else None
        B = x.shape[0]
        flat_g = torch.empty_like(eng.flat_params)
        ws = eng._ws
        call("dmb_train_backward_tm", C.byref(ctx.spec), ptr(packed), ptr(eng.flat_params), ptr(x), ptr(mask_t), ctx.mc,
             ptr(cv), ptr(decoded), B, C.byref(ctx.tm) if ctx.tm is not None else None, 1.0, ptr(flat_g), ptr(ws),
             ctx.nws, _stream())
        flat_g.mul_(g_total)
        eng.last_flat_grad = flat_g
        grads = tuple(flat_g[off:off + n].view(p.shape) if p.requires_grad else None for p, off, n in eng._views)
        return (None, None, None, None) + grads


def train_forward(model, inputs, time_matching_mat=None, batch_mask=None):
    """VQ_VAE.forward in training mode with autograd attached (reference: vq_vae.py:300-338)."""
    eng = model._engine
    eng.flatten()
    params = [p for p, _, _ in eng._views]
    decoded, recon, commit, total, ppl, tml = TrainStepFunction.apply(model, inputs, batch_mask, time_matching_mat,
                                                                      *params)
    out = {'recon_loss': recon, 'commitment_loss': commit,
           'time_matching_loss': tml if time_matching_mat is not None else 0.}
    if getattr(model, "_total_last", False):
        out['perplexity'] = ppl
        out['total_loss'] = total
    else:
        out['total_loss'] = total
        out['perplexity'] = ppl
    return decoded, out
