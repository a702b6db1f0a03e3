"""Eval-mode encode of N patches for several head -> enc.4 sub-batch sizes (csrc/model.cu:enc_subbatch; 0 = off).
    python scripts/subbatch_sweep.py [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
from dynamorph_b200.synthetic import calibrate, synthetic_patches

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VQ_VAE_z16().to(dev)
calibrate(model, synthetic_patches(64, 1, dev))
model.eval()
x = torch.cat([synthetic_patches(2048, 5 + i, dev) for i in range(n // 2048)])
ref = None
for sb in (0, 128, 256, 384, 512, 1024):
    os.environ["DMB_ENC_SUBBATCH"] = str(sb)
    for _ in range(3):
        out = model.encode_latents(x, "eval")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = model.encode_latents(x, "eval")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    same = "" if ref is None else f" identical to sub-batch 0: {bool(torch.equal(out[0], ref[0]) and torch.equal(out[2], ref[2]))}"
    if ref is None:
        ref = [t.clone() for t in out]
    print(f"sub-batch {sb:5d}: {ms:.3f} ms per {n} patches -> {n / ms * 1e3 / 1e6:.3f} M patches/s{same}", flush=True)
