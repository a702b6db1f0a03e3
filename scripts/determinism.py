"""Repeat forward+backward on identical inputs and compare every gradient bit for bit (run on the GPU box).
    python scripts/determinism.py [reps] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
from dynamorph_b200.synthetic import calibrate, synthetic_patches
from dynamorph_b200.trainer import FusedTrainer

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VQ_VAE().to(dev)
calibrate(m, synthetic_patches(32, 1, dev))
m.train()
tr = FusedTrainer(m, lr=0.0, use_graph=False)      # lr 0: parameters never move
x = synthetic_patches(B, 7, dev)
st = tr._plan(x, None)
tr._load(st, x, None, None)
ref = None
bad = {}
names = [(k, o, n) for k, o, n in tr.eng.param_slices()] if hasattr(tr.eng, "param_slices") else None
for r in range(reps):
    tr._fwd_bwd(st)
    torch.cuda.synchronize()
    g = tr.grad.clone(); l = tr.losses.clone(); d = st.decoded.clone()
    if ref is None:
        ref = (g, l, d)
        continue
    if not torch.equal(g, ref[0]):
        idx = (g != ref[0]).nonzero().flatten()
        bad[r] = (int(idx.numel()), int(idx[0]), float((g - ref[0]).abs().max()))
    if not torch.equal(d, ref[2]):
        bad[("decoded", r)] = float((d - ref[2]).abs().max())
    if not torch.equal(l, ref[1]):
        bad[("loss", r)] = (l - ref[1]).tolist()
print("TMA", os.environ.get("DMB_CONV_TMA", "1"), "reps", reps, "B", B, "mismatching reps:", bad if bad else "none")
