"""CPU oracle for the DynaMorph VQ-VAE latent-encoding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``dynamorph_b200``)
may import this file; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / reported CPU baseline.

It is a *functional* restatement (plain functions over a ``state_dict``; no
``nn.Module``) of the reference's arithmetic, written against torch's CPU ops,
which is where the reference's arithmetic lives (``torch>=1.0.1``,
/root/reference/requirements/default.txt:12).  Every function cites the
reference lines it restates.

Pinning: the reference ships no tests, golden vectors or weights,
so parity is pinned by running the *unmodified reference modules* in the build
container (``tests/test_oracle_vs_reference.py``, skipped where /root/reference
is absent) and by golden fixtures those modules produced
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

BN_EPS = 1e-5          # nn.BatchNorm2d default used at vq_vae.py:205,208,279
BN_MOMENTUM = 0.1      # idem

EVAL = "eval"              # running statistics (model.eval())
BATCH = "batch"            # train-mode statistics over the call's batch (run_training.py:404)
PER_SAMPLE = "per_sample"  # train-mode statistics with batch 1 (pipeline/patch_VAE.py:445-449)


# --------------------------------------------------------------------------
# architecture description (key names probed from the reference state_dict)
# --------------------------------------------------------------------------
def arch_of(state: State) -> str:
    """'z16' (vq_vae.VQ_VAE == vae.VQ_VAE_z16) or 'z32' (vae.VQ_VAE_z32)."""
    return "z16" if "enc.10.weight" in state else "z32"


def num_residual_layers(state: State, prefix: str) -> int:
    n = 0
    while f"{prefix}.layers.{n}.1.weight" in state:
        n += 1
    return n


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
_RELU_TAPS = None   # when a list: every ReLU input of the forward pass is appended (near-tie analysis)
_RELU_GRAPH = None  # when a list: (input, output) of every ReLU WITH their autograd graph (gate-flip envelopes)


def _relu(t: Tensor) -> Tensor:
    if _RELU_TAPS is not None:
        _RELU_TAPS.append(t.detach())
    out = F.relu(t)
    if _RELU_GRAPH is not None and out.requires_grad:
        out.retain_grad()
        _RELU_GRAPH.append((t, out))
    return out


def relu_near_ties(x: Tensor, state: State, mode: str, tol: float = 2e-6) -> int:
    """Number of ReLU inputs of one forward pass with |pre-activation| < tol.  A different fp32
    summation order can flip the gate of such an element, which changes gradients by O(one term):
    the ReLU analogue of a VQ near-tie.  Tests relax the gradient tolerance when this is non-zero."""
    global _RELU_TAPS
    _RELU_TAPS = []
    try:
        with torch.no_grad():
            zb = encoder(x, state, mode)
            za = vq_forward(zb, state["vq.w.weight"], 0.25)[0]
            decoder(za, state, mode)
        return int(sum(int(((t.abs() < tol) & (t != 0)).sum()) for t in _RELU_TAPS))
    finally:
        _RELU_TAPS = None

def batchnorm(x: Tensor, state: State, key: str, mode: str,
              new_running: Optional[State] = None) -> Tensor:
    """nn.BatchNorm2d in one of the three modes the reference's callers produce
    (SURVEY.md §3.4).  ``new_running`` (if given) receives the running-stat
    update a train-mode call performs (momentum 0.1, unbiased variance)."""
    g, b = state[key + ".weight"], state[key + ".bias"]
    rm, rv = state[key + ".running_mean"], state[key + ".running_var"]
    if mode == EVAL:
        return F.batch_norm(x, rm, rv, g, b, False, BN_MOMENTUM, BN_EPS)
    if mode == BATCH:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, g, b, True, BN_MOMENTUM, BN_EPS)
        if new_running is not None:
            new_running[key + ".running_mean"] = rm2
            new_running[key + ".running_var"] = rv2
            new_running[key + ".num_batches_tracked"] = state[key + ".num_batches_tracked"] + 1
        return y
    raise ValueError(mode)  # PER_SAMPLE is resolved by the callers (whole net per sample)


def residual_block(x: Tensor, state: State, prefix: str, mode: str,
                   new_running: Optional[State] = None) -> Tensor:
    """vq_vae.py:203-209 (layer definition) and :222-225 (skip sum)."""
    out = x
    for i in range(num_residual_layers(state, prefix)):
        p = f"{prefix}.layers.{i}"
        h = _relu(out)
        h = F.conv2d(h, state[p + ".1.weight"], state[p + ".1.bias"], padding=1)
        h = batchnorm(h, state, p + ".2", mode, new_running)
        h = _relu(h)
        h = F.conv2d(h, state[p + ".4.weight"], state[p + ".4.bias"])
        h = batchnorm(h, state, p + ".5", mode, new_running)
        out = out + h
    return out


def encoder(x: Tensor, state: State, mode: str,
            new_running: Optional[State] = None) -> Tensor:
    """z16: vq_vae.py:276-289 (= vae.py:273-286).  z32: vae.py:401-407."""
    s = state
    if mode == PER_SAMPLE:
        # batch-of-one, train-mode BN: the loop of pipeline/patch_VAE.py:445-449, whole
        # network per sample (running stats mutate there but are never saved).
        return torch.cat([encoder(x[i:i + 1], s, BATCH) for i in range(x.shape[0])], 0)
    if arch_of(s) == "z16":
        h = F.conv2d(x, s["enc.0.weight"], s["enc.0.bias"])
        h = F.conv2d(h, s["enc.1.weight"], s["enc.1.bias"], stride=2, padding=1)
        h = _relu(batchnorm(h, s, "enc.2", mode, new_running))
        h = F.conv2d(h, s["enc.4.weight"], s["enc.4.bias"], stride=2, padding=1)
        h = _relu(batchnorm(h, s, "enc.5", mode, new_running))
        h = F.conv2d(h, s["enc.7.weight"], s["enc.7.bias"], stride=2, padding=1)
        h = _relu(batchnorm(h, s, "enc.8", mode, new_running))
        h = F.conv2d(h, s["enc.10.weight"], s["enc.10.bias"], padding=1)
        h = batchnorm(h, s, "enc.11", mode, new_running)
        return residual_block(h, s, "enc.12", mode, new_running)
    h = F.conv2d(x, s["enc.0.weight"], s["enc.0.bias"], stride=2, padding=1)
    h = _relu(batchnorm(h, s, "enc.1", mode, new_running))
    h = F.conv2d(h, s["enc.3.weight"], s["enc.3.bias"], stride=2, padding=1)
    h = batchnorm(h, s, "enc.4", mode, new_running)
    return residual_block(h, s, "enc.5", mode, new_running)


def decoder(zq: Tensor, state: State, mode: str,
            new_running: Optional[State] = None) -> Tensor:
    """z16: vq_vae.py:291-298.  z32: vae.py:409-414."""
    s = state
    if mode == PER_SAMPLE:
        return torch.cat([decoder(zq[i:i + 1], s, BATCH) for i in range(zq.shape[0])], 0)
    if arch_of(s) == "z16":
        h = _relu(F.conv_transpose2d(zq, s["dec.0.weight"], s["dec.0.bias"], stride=2, padding=1))
        h = _relu(F.conv_transpose2d(h, s["dec.2.weight"], s["dec.2.bias"], stride=2, padding=1))
        h = _relu(F.conv_transpose2d(h, s["dec.4.weight"], s["dec.4.bias"], stride=2, padding=1))
        return F.conv2d(h, s["dec.6.weight"], s["dec.6.bias"])
    h = residual_block(zq, s, "dec.0", mode, new_running)
    h = F.conv_transpose2d(h, s["dec.1.weight"], s["dec.1.bias"], stride=2, padding=1)
    h = _relu(batchnorm(h, s, "dec.2", mode, new_running))
    return F.conv_transpose2d(h, s["dec.4.weight"], s["dec.4.bias"], stride=2, padding=1)


def vq_distances(z: Tensor, codebook: Tensor) -> Tensor:
    """vq_vae.py:65 — direct-difference squared distances, (B,K,H,W)."""
    K, D = codebook.shape
    return torch.sum((z.unsqueeze(1) - codebook.reshape(1, K, D, 1, 1)) ** 2, 2)


def vq_indices(z: Tensor, codebook: Tensor, chunk: int = 16) -> Tensor:
    """vq_vae.py:90-103 (encode_inputs).  Chunked over the batch: positions are
    independent and the reference's broadcast is B*K*D*H*W*4 bytes."""
    out = []
    for i in range(0, z.shape[0], chunk):
        out.append(torch.argmax(-vq_distances(z[i:i + chunk], codebook), 1))
    return torch.cat(out, 0)


def vq_gather(idx: Tensor, codebook: Tensor) -> Tensor:
    """vq_vae.py:105-116 (decode_inputs): (B,H,W) -> (B,D,H,W)."""
    return F.embedding(idx, codebook).transpose(2, 3).transpose(1, 2)


def vq_forward(z: Tensor, codebook: Tensor, commitment_cost: float
               ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """vq_vae.py:52-84.  Returns (straight-through output, loss, perplexity, indices)."""
    K = codebook.shape[0]
    with torch.no_grad():      # argmax cuts the graph anyway; without this the (B,K,D,H,W) broadcast is kept for backward
        idx = vq_indices(z.detach(), codebook.detach())
    q = vq_gather(idx, codebook)
    assert q.shape == z.shape
    z_st = z + (q - z).detach()
    e_latent = F.mse_loss(q.detach(), z)
    q_latent = F.mse_loss(q, z.detach())
    loss = q_latent + commitment_cost * e_latent
    onehot = torch.zeros(idx.numel(), K, device=idx.device)
    onehot.scatter_(1, idx.flatten().unsqueeze(1), 1)
    p = torch.mean(onehot, 0)
    perplexity = torch.exp(-torch.sum(p * torch.log(p + 1e-10)))
    return z_st, loss, perplexity, idx


def best_second_gap(z: Tensor, codebook: Tensor, chunk: int = 16) -> Tensor:
    """Relative gap (d2 - d1)/d1 between the best and second-best code at every
    position: the yardstick for 'documented near-ties' (BASELINE.json north_star)."""
    out = []
    for i in range(0, z.shape[0], chunk):
        d = vq_distances(z[i:i + chunk], codebook)
        top2 = torch.topk(-d, 2, dim=1).values
        d1, d2 = -top2[:, 0], -top2[:, 1]
        out.append((d2 - d1) / torch.clamp(d1, min=1e-30))
    return torch.cat(out, 0)


def time_matching_loss(z_flat: Tensor, mat: Tensor, variant: str,
                       w_a: float = 1.1, w_t: float = 0.1, w_n: float = -0.5,
                       margin: float = 0.5) -> Tensor:
    """variant 'sum' = vq_vae.py:324-331; variant 'hinge' = vae.py:321-335 / :442-456."""
    L = z_flat.shape[1]
    sim = torch.pow(z_flat.reshape(1, -1, L) - z_flat.reshape(-1, 1, L), 2).mean(2)
    assert sim.shape == mat.shape
    if variant == "sum":
        return (sim * mat).sum()
    w = mat.clone()
    w[mat == 2] = w_a
    w[mat == 1] = w_t
    w[mat == 0] = w_n
    tm = sim * w
    tm = torch.where(mat == 0, torch.clamp(tm + margin, min=0), tm)
    return tm.mean()


def forward(x: Tensor, state: State, mode: str, commitment_cost: float = 0.25,
            weight_recon: float = 1.0, weight_commitment: float = 1.0,
            weight_matching: float = 0.005, time_matching_mat: Optional[Tensor] = None,
            batch_mask: Optional[Tensor] = None, tm_variant: Optional[str] = None,
            new_running: Optional[State] = None, **tm_kw
            ) -> Tuple[Tensor, Dict[str, Tensor]]:
    """VQ_VAE.forward, vq_vae.py:300-338 (z16 identical up to the matching term,
    vae.py:314-346; z32 fixes both weights to 1 and matches on z_after, vae.py:417-466)."""
    z_before = encoder(x, state, mode, new_running)
    z_after, c_loss, perplexity, _ = vq_forward(z_before, state["vq.w.weight"], commitment_cost)
    decoded = decoder(z_after, state, mode, new_running)
    if batch_mask is None:
        batch_mask = torch.ones_like(x)
    recon = torch.mean(F.mse_loss(decoded * batch_mask, x * batch_mask, reduction="none")
                       / state["channel_var"])
    z32 = arch_of(state) == "z32"
    total = recon + c_loss if z32 else weight_recon * recon + weight_commitment * c_loss
    tm = 0.0
    if time_matching_mat is not None:
        variant = tm_variant or ("hinge" if z32 else "sum")
        zsrc = z_after if z32 else z_before
        tm = time_matching_loss(zsrc.reshape(zsrc.shape[0], -1), time_matching_mat, variant, **tm_kw)
        total = total + weight_matching * tm
    return decoded, {"recon_loss": recon, "commitment_loss": c_loss,
                     "time_matching_loss": tm, "total_loss": total,
                     "perplexity": perplexity}


# --------------------------------------------------------------------------
# training step (run_training.py:404-408 with Adam of :485)
# --------------------------------------------------------------------------
def trainable_keys(state: State) -> List[str]:
    """Parameter order == nn.Module.parameters() order == state_dict order,
    minus buffers and the frozen channel_var (vq_vae.py:272)."""
    return [k for k in state
            if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))
            and k != "channel_var"]


def bias_feeds_train_bn(state: State) -> List[str]:
    """Conv biases immediately followed by a train-mode BatchNorm: their exact gradient is
    zero (the batch mean absorbs them), so autograd returns rounding noise and Adam turns
    that noise into +-lr steps.  No two implementations agree on these; tests bound them
    by steps*lr instead of comparing values."""
    keys = list(state)
    out = []
    for i, k in enumerate(keys):
        if k.endswith(".bias") and k[:-5] + ".weight" in state and state[k[:-5] + ".weight"].dim() == 4:
            nxt = keys[i + 1] if i + 1 < len(keys) else ""
            stem = nxt.rsplit(".", 1)[0]
            if stem + ".running_mean" in state and nxt.endswith(".weight") and state[nxt].dim() == 1:
                out.append(k)
    return out


def loss_and_grads(x: Tensor, state: State, mode: str = BATCH, **fw
                   ) -> Tuple[Tensor, Dict[str, Tensor], Dict[str, Tensor], State]:
    keys = trainable_keys(state)
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in keys else v.detach().clone())
            for k, v in state.items()}
    new_running: State = {}
    decoded, losses = forward(x, leaf, mode, new_running=new_running, **fw)
    grads = torch.autograd.grad(losses["total_loss"], [leaf[k] for k in keys], allow_unused=True)
    g = {k: (gi if gi is not None else torch.zeros_like(leaf[k])) for k, gi in zip(keys, grads)}
    losses = {k: (v.detach() if isinstance(v, Tensor) else v) for k, v in losses.items()}
    return decoded.detach(), losses, g, new_running


def relu_gate_envelopes(x: Tensor, state: State, mode: str = BATCH, tol: float = 2e-6, max_positions: int = 64,
                        **fw) -> Tuple[Dict[str, Tensor], int]:
    """Per-ELEMENT bound on how far a gradient may move when ReLU gates flip under a different fp32 summation order.

    A ReLU whose input p satisfies 0 < |pre_p| < tol can come out on the other side of zero in another implementation.
    Flipping that one gate adds or removes exactly the paths through p:  delta_p = dL/d(relu_out_p) * d(pre_p)/d(theta).
    Returns ({parameter key: sum_p |delta_p| elementwise}, number of near-tie positions): gradients must agree with the
    reference to the normal tolerance PLUS this envelope, element by element -- zero wherever no near-tie can reach."""
    global _RELU_GRAPH
    keys = trainable_keys(state)
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in keys else v.detach().clone())
            for k, v in state.items()}
    _RELU_GRAPH = []
    try:
        _, losses = forward(x, leaf, mode, **fw)
        taps = _RELU_GRAPH
    finally:
        _RELU_GRAPH = None
    losses["total_loss"].backward(retain_graph=True)
    env = {k: torch.zeros_like(leaf[k]) for k in keys}
    n = 0
    for pre, out in taps:
        near = ((pre.detach().abs() < tol) & (pre.detach() != 0)).nonzero(as_tuple=False)
        for pos in near:
            n += 1
            if n > max_positions:
                raise RuntimeError(f"more than {max_positions} ReLU near-ties; use a larger tolerance budget")
            pos = tuple(int(i) for i in pos)
            g_out = float(out.grad[pos]) if out.grad is not None else 0.0
            if g_out == 0.0:
                continue
            grads = torch.autograd.grad(pre[pos], [leaf[k] for k in keys], retain_graph=True, allow_unused=True)
            for k, g in zip(keys, grads):
                if g is not None:
                    env[k] += (g * g_out).abs()
    return env, n


def adam_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
                b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam (no amsgrad / weight decay), in place; ``step`` is 1-based."""
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def train_step(x: Tensor, state: State, opt: Dict[str, Dict[str, Tensor]], step: int,
               lr: float, mode: str = BATCH, **fw) -> Dict[str, Tensor]:
    """One forward/backward/Adam step; mutates ``state`` (params + running stats)
    and ``opt`` ({'m': {...}, 'v': {...}})."""
    _, losses, grads, new_running = loss_and_grads(x, state, mode, **fw)
    for k, g in grads.items():
        if k not in opt["m"]:
            opt["m"][k] = torch.zeros_like(state[k])
            opt["v"][k] = torch.zeros_like(state[k])
        p = state[k].detach()
        adam_update(p, g, opt["m"][k], opt["v"][k], step, lr)
        state[k] = p
    state.update(new_running)
    return losses


# --------------------------------------------------------------------------
# pipeline drivers
# --------------------------------------------------------------------------
def zscore_patch(imgs: np.ndarray) -> np.ndarray:
    """pipeline/train_utils.py:252-274: per patch, per channel (x-mean)/(std+eps)."""
    mean = imgs.mean(axis=(2, 3), keepdims=True)
    std = imgs.std(axis=(2, 3), keepdims=True)
    return (imgs - mean) / (std + np.finfo(float).eps)


def process_vae_arrays(patches: np.ndarray, state: State, mode: str = PER_SAMPLE,
                       commitment_cost: float = 0.25, batch: int = 1
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """pipeline/patch_VAE.py:418-419 and :445-462 without the pickle I/O:
    z-score -> float32 -> enc -> vq, one sample at a time as written
    (``batch`` > 1 is only meaningful for EVAL / PER_SAMPLE, where samples are
    independent).  Returns the two (N, D*H*W) float32 arrays the reference pickles."""
    data = torch.from_numpy(zscore_patch(np.squeeze(patches))).float()
    assert data.dim() == 4
    zb, za = [], []
    with torch.no_grad():
        for i in range(0, data.shape[0], batch):
            s = data[i:i + batch]
            z_b = encoder(s, state, mode)
            z_a = vq_forward(z_b, state["vq.w.weight"], commitment_cost)[0]
            zb.append(z_b.numpy())
            za.append(z_a.numpy())
    zb, za = np.concatenate(zb, 0), np.concatenate(za, 0)
    n = data.shape[0]
    return zb.reshape(n, -1), za.reshape(n, -1)


def augment_batch(batch: Tensor, rng: np.random.RandomState) -> Tensor:
    """run_training.py:396-403: per-sample flip over {none, H, W} then rot90 k in 0..3."""
    out = batch.clone()
    for i in range(len(batch)):
        img = batch[i]
        flip = rng.choice([0, 1, 2])
        if flip != 0:
            img = torch.flip(img, dims=(int(flip),))
        k = int(rng.choice([0, 1, 2, 3]))
        out[i] = torch.rot90(img, k=k, dims=[1, 2])
    return out


# --------------------------------------------------------------------------
# synthetic inputs / weights (SURVEY.md §8d recipe)
# --------------------------------------------------------------------------
def synthetic_patches(n: int, seed: int, channels: int = 2, size: int = 128) -> Tensor:
    """zscore_patch-like inputs: N(0,1) noise, 3x3 box low-pass, re-standardised
    per patch and channel.  Values are rounded through fp16 so fixtures store compactly
    and exactly."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, channels, size, size, generator=g)
    x = F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), 3, stride=1)
    x = (x - x.mean((2, 3), keepdim=True)) / x.std((2, 3), keepdim=True, unbiased=False)
    return x.half().float()


def default_state(arch: str = "z16", num_inputs: int = 2, num_hiddens: int = 16,
                  num_residual_hiddens: int = 32, num_residual_layers: int = 2,
                  num_embeddings: int = 64, seed: int = 0) -> State:
    """A state_dict with the reference's key names, shapes and default torch
    initialisers (kaiming_uniform(a=sqrt(5)) conv weights, U(+-1/sqrt(fan_in)) bias,
    BN 1/0, N(0,1) codebook).  NOT seed-for-seed identical to constructing the
    reference module (fixtures carry reference-made states for that)."""
    g = torch.Generator().manual_seed(seed)
    st: State = {}

    def conv(key, cout, cin, k, transposed=False):
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        fan_in = shape[1] * k * k
        bound = 1.0 / math.sqrt(fan_in)
        st[key + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        st[key + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(key, c):
        st[key + ".weight"] = torch.ones(c)
        st[key + ".bias"] = torch.zeros(c)
        st[key + ".running_mean"] = torch.zeros(c)
        st[key + ".running_var"] = torch.ones(c)
        st[key + ".num_batches_tracked"] = torch.tensor(0)

    def res(prefix, h, rh, n):
        for i in range(n):
            conv(f"{prefix}.layers.{i}.1", rh, h, 3)
            bn(f"{prefix}.layers.{i}.2", rh)
            conv(f"{prefix}.layers.{i}.4", h, rh, 1)
            bn(f"{prefix}.layers.{i}.5", h)

    h, h2, h4 = num_hiddens, num_hiddens // 2, num_hiddens // 4
    st["channel_var"] = torch.ones(1, num_inputs, 1, 1)
    if arch == "z16":
        conv("enc.0", h2, num_inputs, 1)
        conv("enc.1", h2, h2, 4); bn("enc.2", h2)
        conv("enc.4", h, h2, 4); bn("enc.5", h)
        conv("enc.7", h, h, 4); bn("enc.8", h)
        conv("enc.10", h, h, 3); bn("enc.11", h)
        res("enc.12", h, num_residual_hiddens, num_residual_layers)
        st["vq.w.weight"] = torch.randn(num_embeddings, h, generator=g)
        conv("dec.0", h2, h, 4, True)
        conv("dec.2", h4, h2, 4, True)
        conv("dec.4", h4, h4, 4, True)
        conv("dec.6", num_inputs, h4, 1)
    else:
        conv("enc.0", h2, num_inputs, 4); bn("enc.1", h2)
        conv("enc.3", h, h2, 4); bn("enc.4", h)
        res("enc.5", h, num_residual_hiddens, num_residual_layers)
        st["vq.w.weight"] = torch.randn(num_embeddings, h, generator=g)
        res("dec.0", h, num_residual_hiddens, num_residual_layers)
        conv("dec.1", h2, h, 4, True); bn("dec.2", h2)
        conv("dec.4", num_inputs, h2, 4, True)
    return st


def calibrate_state(state: State, calib: Tensor, seed: int = 0, codebook_mode: str = EVAL,
                    jitter: float = 0.05) -> State:
    """avoid the collapsed-codebook trap of default init.
    Perturb BN affine, set running stats from one train-mode pass over ``calib``
    (momentum 1), then draw codebook rows from encoder outputs (+ jitter)."""
    g = torch.Generator().manual_seed(seed + 1000)
    st = {k: v.clone() for k, v in state.items()}
    bn_keys = [k[:-len(".running_mean")] for k in st if k.endswith(".running_mean")]
    for k in bn_keys:
        c = st[k + ".weight"].numel()
        st[k + ".weight"] = torch.rand(c, generator=g) + 0.5
        st[k + ".bias"] = torch.randn(c, generator=g) * 0.1
    # one BATCH-mode pass; running <- batch statistics (momentum = 1 semantics)
    global BN_MOMENTUM
    keep = BN_MOMENTUM
    BN_MOMENTUM = 1.0
    try:
        nr: State = {}
        with torch.no_grad():
            zb = encoder(calib, st, BATCH, nr)
            if arch_of(st) == "z32":
                decoder(zb, st, BATCH, nr)
    finally:
        BN_MOMENTUM = keep
    for k, v in nr.items():
        st[k] = v if not k.endswith("num_batches_tracked") else torch.tensor(0)
    with torch.no_grad():
        z = encoder(calib, st, codebook_mode)
    K, D = st["vq.w.weight"].shape
    vecs = z.permute(0, 2, 3, 1).reshape(-1, D)
    pick = torch.randperm(vecs.shape[0], generator=g)[:K]
    st["vq.w.weight"] = vecs[pick] + jitter * torch.randn(K, D, generator=g)
    return st
