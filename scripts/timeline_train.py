"""Timeline of ONE eager FusedTrainer step (torch profiler, CUDA activities): stream, start offset, duration of every
kernel, so that the critical path / overlap of the weight-gradient side stream / gaps between launches are visible.
    python scripts/timeline_train.py [batch] [heavy]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
from dynamorph_b200.synthetic import calibrate, synthetic_patches
from dynamorph_b200.trainer import FusedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
heavy = len(sys.argv) > 2 and sys.argv[2] == "heavy"      # BASELINE configs[3]: num_hiddens=64, num_embeddings=512
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = (VQ_VAE_z16(num_hiddens=64, num_embeddings=512) if heavy else VQ_VAE_z16()).to(dev)
calibrate(m, synthetic_patches(64, 1, dev))
m.train()
x = synthetic_patches(B, 7, dev)
for use_graph in (True, False):
    tr = FusedTrainer(m, lr=1e-4, use_graph=use_graph)
    for _ in range(5):
        tr.step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10 if heavy else 50):
        tr.step(x)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} graph={use_graph}: {e0.elapsed_time(e1) / (10 if heavy else 50):.3f} ms/step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(x); torch.cuda.synchronize()
import json, tempfile
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
evs = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
evs.sort(key=lambda e: e["ts"])
t0 = evs[0]["ts"]
streams = {}
for e in evs:
    s = streams.setdefault(e["args"].get("stream"), len(streams))
    name = e["name"].replace("dmb::(anonymous namespace)::", "").replace("void ", "")[:70]
    print(f"s{s} {e['ts'] - t0:8.1f} +{e['dur']:7.1f}  grid {str(e['args'].get('grid')):18s} {name}")
print(f"span {evs[-1]['ts'] + evs[-1]['dur'] - t0:.1f} us, kernel-time sum {sum(e['dur'] for e in evs):.1f} us, {len(evs)} activities")
