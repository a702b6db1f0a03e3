"""Host->device / device->host bandwidth of pinned buffers: one stream vs two concurrent streams, chunk sizes of the
bulk encoder (scripts/pcie_bw.py).  Decides whether BulkEncoder should split its copies."""
import torch, time
dev = torch.device("cuda:0")
MB = 1 << 20
for size_mb in (64, 256, 512):
    n = size_mb * MB // 4
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    h2 = torch.empty(n // 4, dtype=torch.float32, pin_memory=True)
    d2 = torch.empty(n // 4, dtype=torch.float32, device=dev)
    s = [torch.cuda.Stream() for _ in range(3)]
    def t(fn, reps=8):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    def one():
        with torch.cuda.stream(s[0]): d.copy_(h, non_blocking=True)
    def two():
        half = n // 2
        with torch.cuda.stream(s[0]): d[:half].copy_(h[:half], non_blocking=True)
        with torch.cuda.stream(s[1]): d[half:].copy_(h[half:], non_blocking=True)
    def one_bidir():
        with torch.cuda.stream(s[0]): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s[2]): h2.copy_(d2, non_blocking=True)
    def two_bidir():
        half = n // 2
        with torch.cuda.stream(s[0]): d[:half].copy_(h[:half], non_blocking=True)
        with torch.cuda.stream(s[1]): d[half:].copy_(h[half:], non_blocking=True)
        with torch.cuda.stream(s[2]): h2.copy_(d2, non_blocking=True)
    gb = size_mb / 1024
    print(f"{size_mb} MB H2D: 1 stream {gb / t(one):.1f} GB/s, 2 streams {gb / t(two):.1f} GB/s; with a concurrent D2H of a quarter "
          f"the size: 1 stream {gb / t(one_bidir):.1f}, 2 streams {gb / t(two_bidir):.1f} GB/s (H2D bytes only)")
