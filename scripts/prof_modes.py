"""Kernel-level time table (torch profiler, CUDA activities) of ONE encode step per BatchNorm mode.
    python scripts/prof_modes.py [N] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
from dynamorph_b200.synthetic import calibrate, synthetic_patches

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = sys.argv[2] if len(sys.argv) > 2 else "per_sample"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VQ_VAE_z16().to(dev)
calibrate(model, synthetic_patches(64, 1, dev))
model.eval()
x = torch.cat([synthetic_patches(2048, 5 + i, dev) for i in range(n // 2048)])
for _ in range(3):
    model.encode_latents(x, mode)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    model.encode_latents(x, mode)
e1.record(); torch.cuda.synchronize()
print(f"{mode}: {e0.elapsed_time(e1) / 5:.3f} ms per {n} patches")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.encode_latents(x, mode); torch.cuda.synchronize()
for ev in sorted(prof.events(), key=lambda e: e.time_range.start):
    if ev.device_type.name == "CUDA" or "kernel" in ev.name.lower():
        print(f"{ev.cuda_time if hasattr(ev, 'cuda_time') else ev.device_time:9.1f} us  {ev.name[:150]}")
