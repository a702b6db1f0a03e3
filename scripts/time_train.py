"""Train-step timing (BASELINE.json configs[1]: VQ_VAE defaults, batch 256, fp32) + kernel breakdown."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
from dynamorph_b200.optim import FusedAdam
from dynamorph_b200.run_training import run_one_batch
from dynamorph_b200.synthetic import calibrate, synthetic_patches

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VQ_VAE_z16().to(dev)
calibrate(model, synthetic_patches(64, 1, dev))
model.train()
opt = FusedAdam(model, lr=1e-4)
x = synthetic_patches(B, 2, dev)
tl = {}
for _ in range(5):
    run_one_batch(model, x, tl, model_kwargs={}, optimizer=opt, transform=None, training=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 30
for _ in range(n):
    run_one_batch(model, x, tl, model_kwargs={}, optimizer=opt, transform=None, training=True)
torch.cuda.synchronize()
print(f"B={B}: run_one_batch wall {1e3*(time.perf_counter()-t0)/n:.3f} ms/step (includes the loss readback sync)")
# device-only time of forward+backward+adam without the host sync
def step():
    _, d = model(x)
    d["total_loss"].backward()
    opt.step(); model.zero_grad()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): step()
e1.record(); torch.cuda.synchronize()
print(f"B={B}: device {e0.elapsed_time(e1)/n:.3f} ms/step; loss {tl['total_loss'][0]:.4f} -> {tl['total_loss'][-1]:.4f}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))

# ---- fused trainer (C-ABI calls on flat buffers, CUDA graph replay)
from dynamorph_b200.trainer import FusedTrainer
for use_graph in (False, True):
    torch.manual_seed(0)
    model2 = VQ_VAE_z16().to(dev)
    calibrate(model2, synthetic_patches(64, 1, dev))
    model2.train()
    tr = FusedTrainer(model2, lr=1e-4, use_graph=use_graph)
    for _ in range(5): l = tr.step(x)
    torch.cuda.synchronize()
    first = l.tolist()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(50): l = tr.step(x)
    e1.record(); torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0) / 50
    print(f"FusedTrainer graph={use_graph}: device {e0.elapsed_time(e1)/50:.3f} ms/step, wall {wall:.3f} ms/step, "
          f"total_loss {first[2]:.4f} -> {l.tolist()[2]:.4f}")
