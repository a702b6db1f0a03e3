"""Two-GPU debug: the same two training steps on cuda:0 and cuda:1 (cuda:0 current) through the three step paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import Golden
import gpu_util as U
from dynamorph_b200.trainer import FusedTrainer
from dynamorph_b200.optim import FusedAdam
from dynamorph_b200.run_training import run_one_batch

g = Golden("vqvae_default")
st = g.state()
xt = g.t("x_train")
torch.cuda.set_device(0)
for kind in ("nograph", "graph", "eager"):
    for order in ((0, 1), (1, 0)):
        outs = {}
        for dev in order:
            m = U.model_from_state(st).to(f"cuda:{dev}").train()
            x = xt.to(f"cuda:{dev}")
            if kind == "eager":
                opt = FusedAdam(m, lr=1e-3)
                tl = {}
                for _ in range(2):
                    run_one_batch(m, x.clone(), tl, model_kwargs={}, optimizer=opt, transform=None, training=True)
                outs[dev] = [round(v, 6) for v in tl["total_loss"]]
            else:
                tr = FusedTrainer(m, lr=1e-3, use_graph=(kind == "graph"))
                ls = []
                for _ in range(2):
                    ls.append(round(float(tr.step(x)[2]), 6))
                outs[dev] = ls
            torch.cuda.synchronize(dev)
        print(kind, "order", order, outs, "current device", torch.cuda.current_device(), flush=True)
