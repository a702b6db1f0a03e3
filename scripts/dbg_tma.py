"""Bring-up aid for csrc/conv_tma.cu: run one conv shape per subprocess with DMB_TMA_DEBUG variants."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
CHILD = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import test_gpu_conv_layers as T
ks, s, ci, co, W = %s
B = %d
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B, ci, W, W, device="cuda", generator=g)
w = torch.randn(co, ci, ks, ks, device="cuda", generator=g) * (ci*ks*ks) ** -0.5
b = torch.randn(co, device="cuda", generator=g)
y = T.run_layer(x, w, b, ks, s)
ref = T.reference(x, w, b, ks, s)
print("err", float((y.double().cpu() - ref).abs().max() / ref.abs().max()))
'''
import test_gpu_conv_layers as T  # noqa
for sh in T.SHAPES[:16]:
    B, dbg = 3, 0
    env = dict(os.environ, DMB_TMA_DEBUG=str(dbg), CUDA_LAUNCH_BLOCKING="1")
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT, sh, B)], env=env, capture_output=True, text=True, timeout=120)
    tail = (r.stdout.strip().splitlines() or [""])[-1] + " | " + (r.stderr.strip().splitlines() or [""])[-1][:110]
    print(f"dbg={dbg} shape={sh} B={B} rc={r.returncode}: {tail}", flush=True)
