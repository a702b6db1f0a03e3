import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import Golden
import gpu_util as U
g = Golden('vqvae_default')
m = U.model_from_state(g.state()).eval()
B = 4096
x = torch.randn(B, 2, 128, 128, device='cuda')
for mode in ['eval', 'per_sample']:
    for _ in range(3): m.encode_latents(x, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m.encode_latents(x, mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(mode, f'{ms:.3f} ms per {B} patches -> {B/ms*1e3/1e6:.3f} M patches/s')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.encode_latents(x, 'eval'); m.encode_latents(x, 'per_sample'); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20, max_name_column_width=80))
