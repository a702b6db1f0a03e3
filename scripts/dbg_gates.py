import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from conftest import Golden
import gpu_util as U
from oracle import vqvae_oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "vqvae_heavy"
g = Golden(name); st = g.state()
m = U.model_from_state(st).train()
for seed in range(6):
    x = O.synthetic_patches(2, 5000 + seed)
    m.load_state_dict(st); m.zero_grad()
    _, d = m(x.cuda()); d["total_loss"].backward()
    _, losses, grads, _ = O.loss_and_grads(x, st, O.BATCH)
    # count near-zero decoder pre-activations in the reference forward
    with torch.no_grad():
        zb = O.encoder(x, st, O.BATCH); za = O.vq_forward(zb, st["vq.w.weight"], 0.25)[0]
        p1 = F.conv_transpose2d(za, st["dec.0.weight"], st["dec.0.bias"], stride=2, padding=1)
        p2 = F.conv_transpose2d(F.relu(p1), st["dec.2.weight"], st["dec.2.bias"], stride=2, padding=1)
        p3 = F.conv_transpose2d(F.relu(p2), st["dec.4.weight"], st["dec.4.bias"], stride=2, padding=1)
    near = [int((p.abs() < 2e-6).sum()) for p in (p1, p2, p3)]
    named = dict(m.named_parameters())
    errs = {k: float((named[k].grad.cpu() - v).abs().max() / v.abs().max()) for k, v in grads.items() if k.startswith("dec") or k == "enc.4.weight"}
    print(seed, "near-zero preacts", near, {k: f"{e:.1e}" for k, e in errs.items()})
