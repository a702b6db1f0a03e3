"""Profile target: one eval-mode encode of N patches with a 64-wide model (BASELINE configs[3]).
    python scripts/prof_heavy.py [z16|z32] [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z32
from dynamorph_b200.synthetic import calibrate, synthetic_patches
which = sys.argv[1] if len(sys.argv) > 1 else "z32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = (VQ_VAE(num_hiddens=64, num_embeddings=512) if which == "z16" else
     VQ_VAE_z32(num_hiddens=64, num_residual_hiddens=64, num_embeddings=512)).to(dev)
calibrate(m, synthetic_patches(64, 1, dev))
m.eval()
x = torch.cat([synthetic_patches(256, 5 + i, dev) for i in range(n // 256)])
for _ in range(2):
    m.encode_latents(x, "eval")
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.encode_latents(x, "eval")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", which, n)
