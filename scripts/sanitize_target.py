"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16, VQ_VAE_z32
from dynamorph_b200.synthetic import calibrate, synthetic_patches
from dynamorph_b200.run_training import run_one_batch
from dynamorph_b200.optim import FusedAdam

dev = torch.device("cuda:0")
torch.manual_seed(0)
for cls, kw in ((VQ_VAE_z16, {}), (VQ_VAE_z32, {}), (VQ_VAE, dict(num_hiddens=64, num_embeddings=512))):
    m = cls(**kw).to(dev)
    calibrate(m, synthetic_patches(8, 1, dev))
    x = synthetic_patches(3, 2, dev)
    m.eval()
    for weights in ("smem", "const"):
        os.environ["DMB_CONV_WEIGHTS"] = weights
        for mode in ("eval", "per_sample"):
            zb, za, idx = m.encode_latents(x, mode)
    os.environ.pop("DMB_CONV_WEIGHTS")
    m.train()
    opt = FusedAdam(m, lr=1e-4)
    mat = (torch.arange(3).view(-1, 1) + torch.arange(3).view(1, -1)).remainder(3).float().to(dev)
    kw2 = {"time_matching_mat": mat}
    tl = {}
    run_one_batch(m, x.clone(), tl, model_kwargs=kw2, optimizer=opt, transform=True, training=True)
    torch.cuda.synchronize()
    print(cls.__name__, {k: round(v[-1], 4) for k, v in tl.items()})
print("ok")
