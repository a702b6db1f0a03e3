import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import Golden
import gpu_util as U
from oracle import vqvae_oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "vqvae_heavy"
g = Golden(name); st = g.state()
m = U.model_from_state(st).train()
x = g.t("x_train").cuda()
mask = g.t("mask_train").cuda() if g.has("mask_train") else None
_, d = m(x, batch_mask=mask)
d["total_loss"].backward()
noise = set(O.bias_feeds_train_bn(st))
named = dict(m.named_parameters())
for k, ref in g.group("train/grad").items():
    got = named[k].grad.detach().cpu()
    err = float((got - ref).abs().max() / ref.abs().max().clamp(min=1e-30))
    print(f"{k:32s} {'noise' if k in noise else '     '} relerr {err:.3e}  max|ref| {float(ref.abs().max()):.3e} max|got| {float(got.abs().max()):.3e}")
