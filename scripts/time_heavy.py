"""BASELINE.json configs[3]: VQ_VAE(num_hiddens=64, num_embeddings=512) and VQ_VAE_z32(64, 64, 512), encode + train
step at batch 1024 (timing only; parity of these configs is in tests/ on the golden fixtures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z32
from dynamorph_b200.synthetic import calibrate, synthetic_patches
from dynamorph_b200.trainer import FusedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x = torch.cat([synthetic_patches(256, 10 + i, dev) for i in range(B // 256)])
for name, ctor, flops in (("VQ_VAE(64,.,512)", lambda: VQ_VAE(num_hiddens=64, num_embeddings=512), 293601280),
                          ("VQ_VAE_z32(64,64,512)", lambda: VQ_VAE_z32(num_hiddens=64, num_residual_hiddens=64,
                                                                       num_embeddings=512), 310378496)):
    torch.manual_seed(0)
    m = ctor().to(dev)
    calibrate(m, synthetic_patches(64, 1, dev))
    m.eval()
    for mode in ("eval", "per_sample"):
        for _ in range(2):
            m.encode_latents(x, mode)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            m.encode_latents(x, mode)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} encode {mode:10s} B={B}: {ms:8.3f} ms  {B/ms*1e3:10.0f} patches/s  {flops*B/ms/1e9:6.1f} TFLOP/s (reference FLOP count)")
    m.train()
    tr = FusedTrainer(m, lr=1e-4, use_graph=True)
    for _ in range(3):
        tr.step(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        l = tr.step(x)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} train step B={B}: {e0.elapsed_time(e1)/5:8.3f} ms   losses {[round(v, 4) for v in l.tolist()]}")
    del m, tr
    torch.cuda.empty_cache()
