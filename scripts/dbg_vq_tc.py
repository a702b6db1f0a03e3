import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamorph_b200._lib import call, ptr
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for B, D, P, K in ((64, 16, 256, 64), (64, 64, 256, 512), (64, 32, 256, 128)):
    g = torch.Generator(device="cuda").manual_seed(B + D + K)
    z = torch.randn(B, D, P, device="cuda", generator=g)
    pick = torch.randint(0, B * P, (K,), device="cuda", generator=g)
    cb = (z.permute(0, 2, 1).reshape(-1, D)[pick] + 0.05 * torch.randn(K, D, device="cuda", generator=g)).contiguous()
    res = {}
    for dbg in ("0", "1"):
        os.environ["DMB_VQ_TC_DBG"] = dbg
        idx = torch.empty(B, P, dtype=torch.int32, device="cuda")
        call("dmb_vq_forward", ptr(z), ptr(cb), B, D, P, K, None, ptr(idx), None, st)
        torch.cuda.synchronize()
        res[dbg] = idx.clone()
    nc = res["1"].flatten()
    print(f"D={D} K={K}: ncand histogram", torch.bincount(nc.clamp(max=20)).tolist())
