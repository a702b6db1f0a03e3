"""Profile target: ONE eager training step (pack + forward + backward + Adam) between cudaProfilerStart/Stop.
    python scripts/prof_train.py [batch]      (plain run must exit 0 before running it under ncu)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
from dynamorph_b200.synthetic import calibrate, synthetic_patches
from dynamorph_b200.trainer import FusedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VQ_VAE_z16().to(dev)
calibrate(m, synthetic_patches(64, 1, dev))
m.train()
tr = FusedTrainer(m, lr=1e-4, use_graph=False)
x = synthetic_patches(B, 7, dev)
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", B)
