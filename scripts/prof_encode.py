"""Profile target: one eval-mode + one per-sample encode of N patches between cudaProfilerStart/Stop.
    python scripts/prof_encode.py [N]            (plain run must exit 0 before running it under ncu)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
from dynamorph_b200.synthetic import calibrate, synthetic_patches

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["eval", "per_sample"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VQ_VAE_z16().to(dev)
calibrate(model, synthetic_patches(64, 1, dev))
model.eval()
x = torch.cat([synthetic_patches(min(n, 2048), 5 + i, dev) for i in range(max(1, n // 2048))])
for m in modes:
    for _ in range(2):
        model.encode_latents(x, m)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for m in modes:
    model.encode_latents(x, m)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", n, modes)
