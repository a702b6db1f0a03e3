#!/usr/bin/env python
"""Benchmark of the VQ-VAE latent-encoding hot path (BASELINE.json: "encoded patches/sec at 1/2/4/8
B200; train step ms; % of HBM roofline").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

A "step" is one pass of the hot path (model.enc + model.vq, i.e. the body of
pipeline/patch_VAE.py:process_VAE) over one chunk of synthetic z-scored 2x128x128 patches per GPU.
Prints ONE JSON line (rank 0).  See the module-level constants and DESIGN.md "Measurement"."""
from __future__ import annotations

import argparse
import json

import numpy as np
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
# stdout carries the ONE JSON line and nothing else: NCCL writes its version banner (printed at WARN too) and any
# warning to its debug file, stdout by default
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

# algorithmic figures per patch, defaults (SURVEY.md section 8d / BASELINE.md section 4)
ENC_BYTES = 131072 + 16384 + 16384          # x in, z_before out, z_after out
ENC_FLOPS_ALGO = 22151168                    # 2 * (enc MACs + VQ MACs), reference layer structure
ENC_FLOPS_EXEC = 15335424                    # direct-form FLOPs after folding enc.0 (1x1) into enc.1 (4x4), DESIGN.md section 4
# ... of which the CUDA cores still execute this much once the two residual 3x3 layers (2 x 1,179,648 MACs) run as
# Winograd GEMMs on the tensor cores and the quantiser's search (262,144 MACs) as a TF32 GEMM
ENC_FLOPS_CUDA_CORE = 2 * 1048576              # round 2: only the composite head (2 -> 8, 4x4 s2 @128) stays on the CUDA cores
# DRAM bytes per patch of each kernel (dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture of an
# 8192-patch eval encode step of the last build, profiles/r1_launches_encode_step_final.csv: launches in schedule order),
# for roofline.traffic
TRAFFIC_CSV = "r2_launches_encode_step.csv"


def ncu_dram_bytes_per_patch():
    import csv
    path = os.path.join(ROOT, "profiles", TRAFFIC_CSV)
    order = ["enc.0+enc.1 composite conv4x4s2", "enc.4 conv4x4s2", "enc.7 conv4x4s2", "enc.10 conv3x3",
             "res layer fused: conv3x3 -> ReLU -> conv1x1 -> +skip", "res layer fused: conv3x3 -> ReLU -> conv1x1 -> +skip",
             "vq fused (tensor-core search)"]
    try:
        rows = list(csv.reader(open(path)))
        h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr = rows[h]
        tot = {}
        for r in rows[h + 1:]:
            if len(r) == len(hdr) and r[hdr.index("Metric Name")].startswith("dram__bytes_"):
                tot[int(r[0])] = tot.get(int(r[0]), 0.0) + float(r[hdr.index("Metric Value")])
        return {order[i]: v / 8192 for i, v in sorted(tot.items()) if i < len(order)}
    except Exception:
        return {}


CHUNK = 16384                                # patches per step per GPU (2.1 GB of input >> 126 MB L2)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def count(self, t0, t1):
        return sum(1 for ts, _ in self.rows if t0 <= ts <= t1)

    def stop(self, windows):
        """windows: list of (t0, t1) perf_counter intervals during which the measured step was running."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for ts, r in self.rows if any(t0 <= ts <= t1 + 0.05 for t0, t1 in windows)]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        return world, rank, local, dist
    return 1, 0, 0, None


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port; the Python reference
# cannot travel to the GPU box), all host threads, bounded sample per step
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(n_patches: int, batch: int, repeats: int, warmup: int = 1):
    """Batched eval-mode enc+vq on the host (the fastest configuration of the reference's own code,
    BASELINE.md section 2).  Returns patches/s and the thread count."""
    from oracle import vqvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.calibrate_state(O.default_state("z16"), O.synthetic_patches(32, 1), seed=0)
    x = O.synthetic_patches(n_patches, 2)
    times = []
    with torch.no_grad():
        for r in range(warmup + repeats):
            t0 = time.perf_counter()
            for i in range(0, n_patches, batch):
                zb = O.encoder(x[i:i + batch], st, O.EVAL)
                O.vq_forward(zb, st["vq.w.weight"], 0.25)
            dt = time.perf_counter() - t0
            if r >= warmup:
                times.append(dt)
    return n_patches / min(times), torch.get_num_threads(), sum(times) / len(times)


def library_gpu_rate(dev, n_patches: int = 2048, batch: int = 256):
    """The reference's own layer structure executed by the library kernels on the SAME GPU (SURVEY.md section 8d:
    "the kernel set to beat on the same box"): the oracle port's torch ops (cuDNN convolutions, ATen batch-norm / ReLU,
    the (B,K,D,H,W) broadcast quantiser) on cuda, batched eval + no_grad, TF32 off (parity-comparable) and on (the
    PyTorch default for convolutions).  A reported baseline beside cpu_baseline, nothing of ours runs in it."""
    from oracle import vqvae_oracle as O
    st = O.calibrate_state(O.default_state("z16"), O.synthetic_patches(32, 1), seed=0)
    st = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in st.items()}
    x = O.synthetic_patches(n_patches, 2).to(dev)
    out = {}
    keep = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def step():
                with torch.no_grad():
                    for i in range(0, n_patches, batch):
                        zb = O.encoder(x[i:i + batch], st, O.EVAL)
                        O.vq_forward(zb, st["vq.w.weight"], 0.25)

            step()
            torch.cuda.synchronize()
            ms = time_events(step, 3)
            out[tag] = {"patches_per_s": n_patches / (ms * 1e-3), "ms_per_%d" % n_patches: ms}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = keep
    out["kind"] = "oracle port of the reference modules on cuda (torch %s / cuDNN), batch %d, eval + no_grad" % (
        torch.__version__, batch)
    return out


def raw_u16_like(x):
    """Synthetic RAW patches in camera-count range from z-scored ones (what `zscore_patch` receives in the pipeline):
    per-channel offset + gain, rounded to uint16.  Works on CPU and CUDA tensors."""
    return (x * 2000.0 + 30000.0).clamp_(0, 65535).to(torch.uint16)


def run_reference(args):
    """The reference's CPU implementation of the SAME work as our `e2e` headline: raw uint16 patches ->
    `zscore_patch` (pipeline/train_utils.py:252-274, float64 on the host) -> float32 -> enc + vq (batched, eval-mode BN,
    no_grad: the fastest configuration of the reference's own code).  Oracle port on all host threads; bounded sample."""
    world, rank, local, dist = dist_setup(args.gpus)
    if rank != 0:
        return
    sample, batch = 256, 256
    from oracle import vqvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.calibrate_state(O.default_state("z16"), O.synthetic_patches(32, 1), seed=0)
    raw = raw_u16_like(O.synthetic_patches(sample, 2)).numpy()

    def step():
        with torch.no_grad():
            x = torch.from_numpy(O.zscore_patch(raw.astype(np.float64)).astype(np.float32))
            zb = O.encoder(x, st, O.EVAL)
            O.vq_forward(zb, st["vq.w.weight"], 0.25)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "encoded patches/sec", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "process_VAE bulk encode (zscore_patch + enc + vq), VQ_VAE_z16 defaults, raw uint16 2x128x128 "
                               f"patches, eval-BN batched B={batch}; reference CPU path = oracle port on host cores "
                               "(the Python reference cannot travel to the GPU box)",
                   "sample_patches_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} raw patches per step x {args.steps} steps: zscore_patch + batched eval enc+vq"},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_legs(no_loop=False):
    """BASELINE.md section 5 CPU baselines beside the batched-eval one: (3a) the as-written process_VAE loop (batch 1,
    train-mode BN, autograd on, pipeline/patch_VAE.py:443-452) and (3c) one `run_one_batch` training step at batch 256
    (run_training.py:404-408 + Adam) -- oracle port, all host threads, bounded samples."""
    from oracle import vqvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = O.calibrate_state(O.default_state("z16"), O.synthetic_patches(32, 1), seed=0)
    out = {}
    if not no_loop:
        x = O.synthetic_patches(96, 3)
        leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k in O.trainable_keys(st) else v)
                for k, v in st.items()}
        t0 = None
        for i in range(x.shape[0]):
            if i == 32:
                t0 = time.perf_counter()          # first 32 patches = warm-up
            zb = O.encoder(x[i:i + 1], leaf, O.BATCH)
            za = O.vq_forward(zb, leaf["vq.w.weight"], 0.25)[0]
            zb.detach().numpy(); za.detach().numpy()            # the loop's two per-patch read-backs
        dt = time.perf_counter() - t0
        out["as_written_loop"] = {"value": 64 / dt, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                                  "sample": "64 patches one at a time after 32 warm-up: train-mode BN, autograd on, two "
                                            "read-backs per patch (patch_VAE.py:443-452)"}
    xb = O.synthetic_patches(256, 4321)
    state = {k: v.clone() for k, v in st.items()}
    opt = {"m": {}, "v": {}}
    times = []
    for step in range(1, 4):
        t0 = time.perf_counter()
        O.train_step(xb, state, opt, step, 1e-4)
        times.append(time.perf_counter() - t0)
    out["train_step_b256"] = {"value": min(times[1:]) * 1e3, "unit": "ms/step", "cores": torch.get_num_threads(),
                              "kind": "port", "sample": "best of 2 steps after 1 warm-up: forward + backward + Adam, "
                                                        "batch 256 (run_training.py:404-408, :485)"}
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def time_events(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fp32_peak(dev):
    import ctypes as C
    from dynamorph_b200._lib import call, ptr
    scratch = torch.zeros(16, device=dev)
    flops = C.c_double()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: call("dmb_bench_fp32_fma", 148 * 8, 256, 20000, ptr(scratch), C.byref(flops), st)
    fn()
    torch.cuda.synchronize()
    ms = min(time_events(fn, 1) for _ in range(5))
    return flops.value / (ms * 1e-3) / 1e12


def layer_table(model, x, hbm_peak, fp32_tf):
    """Per-launch device time of every kernel in one eval-mode encode step at bulk batch size -- each launched alone
    through the layer-level C ABI on tensors of the model's layer shapes, the kernel the schedule dispatches
    (csrc/model.cu:run_conv / run_res): the composite head on the CUDA cores, every layer behind it on tcgen05 with the
    activation operand in tensor memory (csrc/conv_tm.cu; DMB_TM=0: the round-1 CUDA-core / Winograd kernels), the
    quantiser's tensor-core search.  Both roofs per row; `path` says which datapath does the multiply-adds."""
    import ctypes as C
    from dynamorph_b200._lib import call, ptr
    B = x.shape[0]
    dev = x.device
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = model.num_hiddens
    rh = model.num_residual_hiddens
    K = model.num_embeddings
    tm_on = os.environ.get("DMB_TM", "1") != "0" and (h, rh) == (16, 32)
    fuse_on = tm_on and os.environ.get("DMB_TM_FUSE", "1") != "0"
    rows = []

    def conv_row(name, cin, H, cout, ks, s, in_relu, out_relu, with_skip, tm, launches=1):
        xin = torch.randn(B, cin, H, H, device=dev)
        w = torch.randn(cin * ks * ks * cout, device=dev) * 0.05
        b = torch.zeros(9 * cout, device=dev)
        y = torch.empty(B, cout, H // s, H // s, device=dev)
        skip = torch.randn(B, cout, H // s, H // s, device=dev) if with_skip else None
        fn_cc = lambda: call("dmb_conv2d_forward", ptr(xin), ptr(w), ptr(b), ptr(y), B, cin, H, H, cout, ks, s, None, None, 0,
                             in_relu, ptr(skip), out_relu, st)
        fn = fn_cc
        if tm:
            n = C.c_int64()
            call("dmb_conv2d_tm_scratch_floats", cin, cout, ks, C.byref(n))
            scratch = torch.zeros(n.value, device=dev)
            fn = lambda: call("dmb_conv2d_tm", ptr(xin), ptr(w), ptr(b), ptr(y), B, cin, H, H, cout, ks, s, in_relu, ptr(skip),
                              out_relu, ptr(scratch), st)
        fn(); torch.cuda.synchronize()
        ms = time_events(fn, 5)
        macs = (H // s) ** 2 * cout * cin * ks * ks
        byts = (cin * H * H + cout * (H // s) ** 2 * (2 if with_skip else 1)) * 4
        row = {"kernel": name, "launches_per_step": launches, "ms": ms, "gbs": byts * B / ms / 1e6,
               "hbm_frac": byts * B / ms / 1e6 / hbm_peak, "path": "tcgen05 (A operand in tensor memory, 3xTF32)" if tm
               else "cuda_core"}
        if tm:
            fn_cc(); torch.cuda.synchronize()
            row.update({"tflops": None, "fp32_frac": None, "direct_form_flops_equivalent_tflops": 2 * macs * B / ms / 1e9,
                        "cuda_core_kernel_ms": time_events(fn_cc, 5)})
        else:
            row.update({"tflops": 2 * macs * B / ms / 1e9, "fp32_frac": 2 * macs * B / ms / 1e9 / fp32_tf})
        rows.append(row)

    conv_row("enc.0+enc.1 composite conv4x4s2", 2, 128, h // 2, 4, 2, 0, 1, False, False)
    conv_row("enc.4 conv4x4s2", h // 2, 64, h, 4, 2, 0, 1, False, tm_on)
    conv_row("enc.7 conv4x4s2", h, 32, h, 4, 2, 0, 1, False, tm_on)
    conv_row("enc.10 conv3x3", h, 16, h, 3, 1, 0, 0, False, tm_on)
    nres = model.num_residual_layers
    if fuse_on:
        xin = torch.randn(B, 16, 16, 16, device=dev)
        w1 = torch.randn(16 * 9 * 32, device=dev) * 0.05
        w2 = torch.randn(32 * 16, device=dev) * 0.05
        b1, b2 = torch.zeros(32, device=dev), torch.zeros(16, device=dev)
        y = torch.empty_like(xin)
        n = C.c_int64()
        call("dmb_residual_layer_tm_scratch_floats", C.byref(n))
        scratch = torch.zeros(n.value, device=dev)
        fn = lambda: call("dmb_residual_layer_tm", ptr(xin), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(y), B, ptr(scratch), st)
        fn(); torch.cuda.synchronize()
        ms = time_events(fn, 5)
        macs = 256 * (16 * 9 * 32 + 32 * 16)
        byts = 2 * 16 * 256 * 4
        rows.append({"kernel": "res layer fused: conv3x3 -> ReLU -> conv1x1 -> +skip", "launches_per_step": nres, "ms": ms,
                     "gbs": byts * B / ms / 1e6, "hbm_frac": byts * B / ms / 1e6 / hbm_peak, "tflops": None, "fp32_frac": None,
                     "direct_form_flops_equivalent_tflops": 2 * macs * B / ms / 1e9,
                     "path": "tcgen05 (two chained GEMMs per tile, A operands in tensor memory, 3xTF32)"})
    else:
        conv_row("res conv3x3", h, 16, rh, 3, 1, 1, 1, False, tm_on, nres)
        conv_row("res conv1x1", rh, 16, h, 1, 1, 0, 0, True, tm_on, nres)
    z = torch.randn(B, h, 16, 16, device=dev)
    cb = torch.randn(K, h, device=dev)
    zst = torch.empty_like(z)
    idx = torch.empty(B, 16, 16, dtype=torch.int32, device=dev)
    fn = lambda: call("dmb_vq_forward", ptr(z), ptr(cb), B, h, 256, K, ptr(zst), ptr(idx), None, st)
    fn(); torch.cuda.synchronize()
    ms = time_events(fn, 5)
    byts = (2 * h * 256 + 256) * 4
    ops = 256 * K * h * 3
    # the search runs on the tensor cores (csrc/vq_tc.cu: TF32 score GEMM + exact refinement of ~1.1 candidates per
    # position), so the direct form's 3 ops per (position, code, channel) are no longer executed: no FP32 fraction
    rows.append({"kernel": "vq fused (tensor-core search)", "launches_per_step": 1, "ms": ms, "gbs": byts * B / ms / 1e6,
                 "tflops": None, "hbm_frac": byts * B / ms / 1e6 / hbm_peak, "fp32_frac": None,
                 "direct_form_ops_equivalent_tflops": ops * B / ms / 1e9, "path": "tcgen05 (TF32 search) + exact fp32 refinement"})
    return rows


def pcie_h2d_peak(dev, barrier=None):
    """Pinned host -> device copy bandwidth of this rank (256 MB, best of 5) WHILE every other rank does the same
    (`barrier` lines them up): summed over ranks it is the roof of the end-to-end number."""
    n = 64 << 20
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    best = 0.0
    for _ in range(5):
        if barrier is not None:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n * 4 / (time.perf_counter() - t0) / 1e9)
    return best


def pcie_duplex_peak(dev, barrier=None):
    """What the end-to-end step actually asks of the host link: device -> pinned host alone, and host -> device WHILE
    device -> host runs (256 MB in, 132 MB out: the byte ratio of the raw-input step), every rank at the same time.
    Returns (d2h_alone, h2d_duplex, d2h_duplex) in GB/s for this rank, best of 3."""
    n_in, n_out = 64 << 20, 33 << 20
    h_in = torch.empty(n_in, dtype=torch.float32, pin_memory=True)
    d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
    h_out = torch.empty(n_out, dtype=torch.float32, pin_memory=True)
    d_out = torch.empty(n_out, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = [0.0, 0.0, 0.0]
    for _ in range(3):
        if barrier is not None:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best[0] = max(best[0], n_out * 4 / (time.perf_counter() - t0) / 1e9)
        if barrier is not None:
            barrier()
        torch.cuda.synchronize()
        a0, a1, b0, b1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        with torch.cuda.stream(s1):
            a0.record(s1); d_in.copy_(h_in, non_blocking=True); a1.record(s1)
        with torch.cuda.stream(s2):
            b0.record(s2); h_out.copy_(d_out, non_blocking=True); b1.record(s2)
        torch.cuda.synchronize()
        best[1] = max(best[1], n_in * 4 / (a0.elapsed_time(a1) * 1e-3) / 1e9)
        best[2] = max(best[2], n_out * 4 / (b0.elapsed_time(b1) * 1e-3) / 1e9)
    return best


def parity_sample(model, x_dev, zb_rows, idx_rows, k, raw=None, seed=0):
    """Oracle parity of `k` patches spread over a timed batch (first / last 16 + random rows): z_before within 1e-4
    relative, code indices equal except where the oracle's best / second-best gap is below 1e-6 relative (the bar of
    BASELINE.json north_star).  `raw` (host uint16): the oracle also restates zscore_patch (patch_VAE.py:413-419)."""
    from oracle import vqvae_oracle as O
    n = zb_rows.shape[0]
    rng = np.random.RandomState(seed)
    edge = list(range(16)) + list(range(n - 16, n))
    rows = sorted(edge + rng.choice(np.arange(16, n - 16), size=k - len(edge), replace=False).tolist())
    rt = torch.tensor(rows)
    state = {kk: v.detach().cpu().clone() for kk, v in model.state_dict().items()}
    if raw is not None:
        xs = torch.from_numpy(O.zscore_patch(raw[rt].numpy().astype(np.float64)).astype(np.float32))
    else:
        xs = x_dev[rt.to(x_dev.device)].cpu()
    got_z = zb_rows[rt.to(zb_rows.device)].cpu().reshape(len(rows), -1)
    got_i = idx_rows[rt.to(idx_rows.device)].cpu().long()
    with torch.no_grad():
        ref = O.encoder(xs, state, O.EVAL)
        ref_idx = O.vq_indices(ref, state["vq.w.weight"])
        gap = O.best_second_gap(ref, state["vq.w.weight"])
    max_rel = float((got_z - ref.reshape(len(rows), -1)).abs().max() / ref.abs().max())
    bad = got_i != ref_idx
    unexplained = int((bad & (gap >= 1e-6)).sum())
    res = {"patches": len(rows), "of": n, "max_rel": max_rel, "mismatches": int(bad.sum()),
           "mismatches_not_near_tie": unexplained, "positions": int(bad.numel()),
           "ok": bool(max_rel < 1e-4 and unexplained == 0), "checker": "oracle/vqvae_oracle.py (CPU)"}
    assert res["ok"], f"bench output fails oracle parity: {res}"
    return res


def pca_table(dev, hbm_peak):
    """the PCA projection of the latents (run_dim_reduction.py:86 `pca.transform`), 65,536 latent
    rows x 4096 -> 32 components resident on the device: the tcgen05 form (3xTF32 1x1 convolution, csrc/pca.cu ->
    conv_tc.cu) beside the fp32 CUDA-core GEMM.  The op reads each latent row once: HBM is its roof."""
    import ctypes as C
    from dynamorph_b200._lib import call, ptr
    n, l, k = 65536, 4096, 32
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(n, l, device=dev, generator=g)
    comp = torch.randn(k, l, device=dev, generator=g) / 64.0
    mean = torch.randn(l, device=dev, generator=g) * 0.1
    out_tc, out_cc = torch.empty(n, k, device=dev), torch.empty(n, k, device=dev)
    nf = C.c_int64()
    call("dmb_pca_transform_scratch_floats", n, l, C.byref(nf))
    scratch = torch.empty(nf.value, device=dev)
    f_tc = lambda: call("dmb_pca_transform_tc", ptr(x), n, l, ptr(mean), ptr(comp), k, None, ptr(out_tc), ptr(scratch), st)
    f_cc = lambda: call("dmb_pca_transform", ptr(x), n, l, ptr(mean), ptr(comp), k, None, ptr(out_cc), st)
    res = {"rows": n, "latent_len": l, "components": k}
    for tag, fn in (("tcgen05", f_tc), ("cuda_core", f_cc)):
        fn(); torch.cuda.synchronize()
        ms = time_events(fn, 3)
        res[tag] = {"ms": ms, "rows_per_s": n / (ms * 1e-3), "gbs": n * l * 4 / ms / 1e6, "hbm_frac": n * l * 4 / ms / 1e6 / hbm_peak}
    ref = (x[:2048].double() - mean.double()) @ comp.double().T
    res["tcgen05"]["max_err_of_max"] = float((out_tc[:2048].double() - ref).abs().max() / ref.abs().max())
    res["cuda_core"]["max_err_of_max"] = float((out_cc[:2048].double() - ref).abs().max() / ref.abs().max())
    return res


def wide_config_table(dev, bf16_peak):
    """BASELINE.json configs[3] (64-wide, 512 codes): eval-mode encode at batch 1024 on the tcgen05 path
    (csrc/conv_tc.cu 3xTF32 convs + csrc/vq_tc.cu tensor-core code search) and, for comparison, with both switched
    off (DMB_TC=0, DMB_VQ_TC=0: the CUDA-core kernels).  Timing only; parity is tests/test_gpu_conv_tc.py."""
    from dynamorph_b200.HiddenStateExtractor.vq_vae import VQ_VAE
    from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z32
    from dynamorph_b200.synthetic import calibrate, synthetic_patches
    B = 1024
    x = torch.cat([synthetic_patches(256, 10 + i, dev) for i in range(B // 256)])
    out = {}
    for name, ctor, flops, tc_flops in (
            ("VQ_VAE(num_hiddens=64,num_embeddings=512)", lambda: VQ_VAE(num_hiddens=64, num_embeddings=512),
             293601280, 2 * (70254592 + 8388608)),
            ("VQ_VAE_z32(64,64,512)", lambda: VQ_VAE_z32(num_hiddens=64, num_residual_hiddens=64, num_embeddings=512),
             310378496, 2 * (117440512 + 33554432))):
        torch.manual_seed(0)
        m = ctor().to(dev)
        calibrate(m, synthetic_patches(64, 1, dev))
        m.eval()
        row = {"batch": B, "flops_per_patch_reference": flops}
        for tag, env in (("tensor_core", None), ("cuda_core", "0")):
            for k in ("DMB_TC", "DMB_VQ_TC"):
                if env is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = env
            for _ in range(2):
                m.encode_latents(x, "eval")
            torch.cuda.synchronize()
            ms = time_events(lambda: m.encode_latents(x, "eval"), 5)
            row[tag] = {"ms": ms, "patches_per_s": B / (ms * 1e-3), "tflops_reference_count": flops * B / ms / 1e9}
        for k in ("DMB_TC", "DMB_VQ_TC"):
            os.environ.pop(k, None)
        # MACs that run on the tensor cores (every conv behind the head + the code search), x3 passes for the convs
        row["tensor_core"]["speedup_vs_cuda_core"] = row["cuda_core"]["ms"] / row["tensor_core"]["ms"]
        # the training step of the same configuration at the same batch (forward + backward + Adam, graph replay); these
        # widths train on the CUDA-core kernels
        from dynamorph_b200.trainer import FusedTrainer
        m.train()
        tr = FusedTrainer(m, lr=1e-4, use_graph=True)
        for _ in range(3):
            tr.step(x)
        torch.cuda.synchronize()
        tms = time_events(lambda: tr.step(x), 3)
        row["train_step"] = {"ms": tms, "batch": B, "patches_per_s": B / (tms * 1e-3), "path": "cuda_core"}
        del tr
        out[name] = row
        del m
        torch.cuda.empty_cache()
    out["note"] = ("tensor peak for TF32 is about half the measured bf16 peak (%.0f TFLOP/s); the 3xTF32 split executes "
                   "three MMAs per fp32-accurate product" % (bf16_peak / 2))
    return out


def run_ours(args):
    world, rank, local, dist = dist_setup(args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from dynamorph_b200.dist import bind_to_gpu_numa
    all_cores = os.sched_getaffinity(0)
    numa_cores = bind_to_gpu_numa(local)          # pinned staging buffers land next to this rank's GPU
    if dist is not None:
        dist.init_process_group("nccl", device_id=dev)
    from dynamorph_b200 import _lib
    from dynamorph_b200.HiddenStateExtractor.vae import VQ_VAE_z16
    from dynamorph_b200.bulk import BulkEncoder
    from dynamorph_b200.synthetic import calibrate, synthetic_patches
    lib = _lib.load()

    torch.manual_seed(0)
    model = VQ_VAE_z16().to(dev)
    calibrate(model, synthetic_patches(64, 1, dev))
    model.eval()
    chunk = args.chunk
    # this rank's shard of the job: patch range [rank*chunk*steps, ...) -- independent, no collective
    x = torch.cat([synthetic_patches(2048, 1234 + rank * 1000 + i, dev) for i in range(chunk // 2048)])
    eng = model._engine
    import ctypes as C
    s = eng.spec(128, 128)
    d = model.num_hiddens
    zb = torch.empty(chunk, d, 16, 16, device=dev)
    za = torch.empty_like(zb)
    idx = torch.empty(chunk, 16, 16, dtype=torch.int32, device=dev)

    def step(mode="eval"):
        eng.encode(x, mode, out=(zb, za, idx))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    lib.dmb_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    w1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = lib.dmb_launch_count(0)
    clocks = None
    windows = [(w0, w1)]
    # nvidia-smi samples every 100 ms: a short timed region can fall between two samples.  Then the SAME step keeps
    # running (untimed) until a few samples exist, and the clock record says so.
    extra = 0
    if rank == 0:
        p0 = time.perf_counter()
        while sampler.proc is not None and sampler.count(w0, w1) + sampler.count(p0, time.perf_counter()) < 4 \
                and time.perf_counter() - p0 < 3.0:
            step()
            torch.cuda.synchronize()
            extra += 1
        if extra:
            windows.append((p0, time.perf_counter()))
        clocks = sampler.stop(windows)
        clocks["window"] = "timed region" + (f" + {extra} more identical untimed steps (timed region shorter than "
                                             "the 100 ms sampling period x 4)" if extra else "")
    barrier()
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0])
    value = world * args.steps * chunk / (ms_total * 1e-3)

    # secondary: as-written BN semantics (train-mode, batch 1 == per-patch statistics)
    for _ in range(2):
        step("per_sample")
    torch.cuda.synchronize()
    ms_ps = time_events(lambda: step("per_sample"), max(3, args.steps // 4))

    # ---- end to end through the public host-buffer API: pinned host in -> HBM -> pinned host out, every step.
    # The SAME number of patches per rank at every N (the per-rank pinned footprint is ~232 KB / patch).
    # Headline: the call process_VAE makes -- RAW patches (uint16, camera-count range) shipped as they are, z-scored on
    # the device in front of the encoder (pipeline/train_utils.py:252-274), latents back to the host; the reference
    # arm does the same work on the host cores.  Beside it: pre-z-scored float32 patches (twice the H2D bytes).
    ne = min(chunk, 8192)
    x_host = torch.empty(ne, 2, 128, 128, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x[:ne])
    x16 = torch.empty(ne, 2, 128, 128, dtype=torch.uint16, pin_memory=True)
    x16.copy_(raw_u16_like(x[:ne].clone()))
    e2e_steps = max(3, min(args.steps, 8))

    def time_e2e(encoder, src, out):
        encoder.encode(src, out)
        torch.cuda.synchronize()
        barrier()
        e0.record()
        for _ in range(e2e_steps):
            encoder.encode(src, out)
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * e2e_steps * ne / (float(tt[0]) * 1e-3)

    enc16 = BulkEncoder(model, chunk=min(ne, 4096), bn_mode="eval", zscore=True)
    out = enc16.allocate_outputs(ne)
    e2e16_value = time_e2e(enc16, x16, out)
    d2h = sum(v.numel() * v.element_size() for v in out.values())
    # parity of what came back to the host: 256 patches of the raw-input run against the oracle's process_VAE restatement
    parity_e2e = None
    if rank == 0:
        parity_e2e = parity_sample(model, None, out["z_before"].view(ne, -1), out["idx"].view(ne, 16, 16), 64,
                                   raw=x16, seed=7)
    enc = BulkEncoder(model, chunk=min(ne, 4096), bn_mode="eval")
    e2e_value = time_e2e(enc, x_host, out)
    del enc16
    # aggregate pinned host -> device bandwidth with every rank copying at once: the roof of the end-to-end number
    h2d_peak = pcie_h2d_peak(dev, barrier)
    tt = torch.tensor([h2d_peak], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    h2d_peak_total = float(tt[0])
    tt = torch.tensor(pcie_duplex_peak(dev, barrier), device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    d2h_alone_total, h2d_duplex_total, d2h_duplex_total = (float(v) for v in tt)

    # ---- parity of the TIMED device path: 256 patches spread over the timed batch against the oracle
    step("eval")
    torch.cuda.synchronize()
    parity = parity_sample(model, x, zb.view(chunk, -1), idx, 256, seed=rank) if rank == 0 else None

    # ---- train step (BASELINE.json configs[1] / [4]): forward + backward + (allreduce) + Adam, fp32, BATCH-mode BN.
    # Batch 256 per GPU (configs[1]) AND 512 per GPU (configs[4]: 8 x 512) at every N, so efficiency is computable.
    from dynamorph_b200.trainer import FusedTrainer
    train = {}
    for tb in (256, 512):
        torch.manual_seed(0)
        tmodel = VQ_VAE_z16().to(dev)
        calibrate(tmodel, synthetic_patches(64, 1, dev))
        tmodel.train()
        trainer = FusedTrainer(tmodel, lr=1e-4, use_graph=True)
        xt = synthetic_patches(tb, 4321 + rank, dev)
        for _ in range(5):
            trainer.step(xt)
        barrier()
        tsteps = max(10, args.steps)
        e0.record()
        for _ in range(tsteps):
            tl = trainer.step(xt)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / tsteps], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        train[tb] = {"ms": float(t[0]), "batch_per_gpu": tb, "global_batch": tb * world,
                     "patches_per_s": tb * world / (float(t[0]) * 1e-3), "total_loss_after": float(tl.tolist()[2])}
        if tb == 256:
            # kernels of the library in ONE step (counted on an eager replica of the same step; the timed steps above are
            # graph replays of exactly these launches)
            eager = FusedTrainer(tmodel, lr=1e-4, use_graph=False)
            eager.step(xt)
            torch.cuda.synchronize()
            lib.dmb_launch_count(1)
            eager.step(xt)
            torch.cuda.synchronize()
            train[tb]["launches_per_step"] = int(lib.dmb_launch_count(0))
            del eager
        del trainer, tmodel
    # the reference's own API for the same step (run_one_batch: autograd + optimiser.step + 1 host sync), N = 1 only
    eager_ms = None
    if world == 1:
        from dynamorph_b200.optim import FusedAdam
        from dynamorph_b200.run_training import run_one_batch
        torch.manual_seed(0)
        tmodel = VQ_VAE_z16().to(dev)
        calibrate(tmodel, synthetic_patches(64, 1, dev))
        tmodel.train()
        opt = FusedAdam(tmodel, lr=1e-4)
        xt = synthetic_patches(256, 4321, dev)
        tl_ = {}
        for _ in range(3):
            run_one_batch(tmodel, xt, tl_, model_kwargs={}, optimizer=opt, transform=None, training=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            run_one_batch(tmodel, xt, tl_, model_kwargs={}, optimizer=opt, transform=None, training=True)
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t0) / 10 * 1e3
        del tmodel, opt
    train_ms = train[256]["ms"]

    line = {
        "metric": "encoded patches/sec", "value": value, "unit": "patches/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "process_VAE bulk encode (model.enc + model.vq), VQ_VAE_z16 reference defaults "
                               "(num_inputs=2,num_hiddens=16,num_residual_hiddens=32,num_embeddings=64), "
                               f"{chunk} synthetic z-scored 2x128x128 patches per step per GPU, eval-mode BN, fp32; "
                               "patch ranges sharded across ranks, no data-path collective",
                   "chunk_patches": chunk, "bn_mode": "eval",
                   "l2_policy": f"inputs larger than L2 ({chunk * 131072 / 1e9:.2f} GB per step vs 126 MB)"},
        "e2e": {"value": e2e16_value, "unit": "patches/s", "h2d_bytes_per_step": x16.numel() * 2, "d2h_bytes_per_step": d2h,
                "api": "dynamorph_b200.bulk.BulkEncoder(zscore=True).encode -- the body of process_VAE: RAW uint16 patches in "
                       "pinned host memory -> HBM -> zscore_patch on the device -> enc + vq -> z_before, z_after, idx "
                       "back in pinned host memory (3-stream pipeline)",
                "patches_per_step_per_gpu": ne,
                "rank0_cpu_affinity": (f"{len(numa_cores)} cores local to the GPU (NVML)" if numa_cores else "unchanged"),
                "float32_zscored_input": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": x_host.numel() * 4,
                                          "note": "BulkEncoder.encode on pre-z-scored float32 patches (round 1's e2e "
                                                  "configuration): twice the host->device bytes"},
                "pcie": {"h2d_gbs_achieved_all_ranks": e2e16_value * (x16.numel() * 2 / ne) / 1e9,
                         "h2d_gbs_achieved_all_ranks_float32": e2e_value * (x_host.numel() * 4 / ne) / 1e9,
                         "h2d_gbs_peak_all_ranks_measured": h2d_peak_total,
                         "frac": e2e16_value * (x16.numel() * 2 / ne) / 1e9 / h2d_peak_total,
                         "frac_float32": e2e_value * (x_host.numel() * 4 / ne) / 1e9 / h2d_peak_total,
                         "d2h_gbs_peak_all_ranks_measured": d2h_alone_total,
                         "duplex": {"h2d_gbs": h2d_duplex_total, "d2h_gbs": d2h_duplex_total,
                                    "d2h_gbs_achieved_all_ranks": e2e16_value * (d2h / ne) / 1e9,
                                    "frac_h2d": e2e16_value * (x16.numel() * 2 / ne) / 1e9 / max(h2d_duplex_total, 1e-9),
                                    "note": "both directions at once, 256 MB in / 132 MB out per rank (the byte ratio of "
                                            "the raw-input step), every rank at the same time: the link as the step uses it"},
                         "note": "peak = pinned host->device copies issued by all ranks at the same time (256 MB each, "
                                 "best of 5), summed; the end-to-end step is bound by it"}},
        "gpu_launches": int(launches),
        "train_step": {"ms": train_ms, "batch_per_gpu": 256, "global_batch": 256 * world,
                       "patches_per_s": 256 * world / (train_ms * 1e-3), "dtype": "f32",
                       "what": "forward + backward + one flat-gradient NCCL allreduce (N>1) + fused Adam, CUDA-graph replay "
                               "(trainer.FusedTrainer); per-rank BatchNorm statistics",
                       "total_loss_after": train[256]["total_loss_after"],
                       "launches_per_step": train[256].get("launches_per_step"),
                       "batch_512_per_gpu": train[512],
                       "run_one_batch_ms": eager_ms,
                       "run_one_batch_note": "the reference's own step API (run_training.run_one_batch: autograd + "
                                             "optimizer.step + zero_grad + one host read-back), batch 256, wall clock; "
                                             "forward and backward are CUDA-graph replays from the second call of a "
                                             "batch geometry on (dynamorph_b200/autograd.py)"},
        "per_sample_bn": {"value": world * chunk / (ms_ps * 1e-3), "unit": "patches/s", "ms_per_step": ms_ps,
                          "note": "as-written process_VAE semantics (train-mode BN, batch 1) at batch speed"},
    }
    if rank == 0:
        line["parity_sample"] = parity
        line["e2e"]["parity_sample"] = parity_e2e
        line["clocks"] = clocks
        hbm_peak, peak_src = measured_peaks()
        fp32_tf = fp32_peak(dev)
        rows = layer_table(model, x[:min(chunk, 8192)], hbm_peak, fp32_tf)
        total_ms = sum(r["ms"] * r["launches_per_step"] for r in rows)
        dom = max(rows, key=lambda r: r["ms"] * r["launches_per_step"])
        nb = min(chunk, 8192)
        traffic = ncu_dram_bytes_per_patch().get(dom["kernel"])
        fp32_bound = dom.get("fp32_frac") is not None and dom["fp32_frac"] > dom["hbm_frac"]
        tensor_path = str(dom.get("path", "")).startswith("tcgen05")
        line["roofline"] = {
            "kernel": dom["kernel"], "share_of_step": dom["ms"] * dom["launches_per_step"] / total_ms,
            # the binding roof of this kernel, named for what it is: FP32 FMA on the CUDA cores when its arithmetic
            # intensity sits above the FP32 ridge (11 FLOP/B), else HBM
            "bound": "fp32_fma" if fp32_bound else "hbm",
            "achieved": dom["tflops"] if fp32_bound else dom["gbs"],
            "peak": fp32_tf if fp32_bound else hbm_peak,
            "unit": "TFLOP/s" if fp32_bound else "GB/s",
            "frac": dom["fp32_frac"] if fp32_bound else dom["hbm_frac"],
            "peak_source": ("measured live (dmb_bench_fp32_fma register-resident FMA chains; nominal 74.4)" if fp32_bound
                            else peak_src),
            "hbm": {"achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["hbm_frac"],
                    "peak_source": peak_src, "algorithmic_bytes": dom["gbs"] * 1e9 * dom["ms"] * 1e-3},
            "traffic": traffic * nb if traffic else None,
            "traffic_source": "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per patch of the "
                              "committed ncu capture of the same kernel (profiles/, see ncu_dram_bytes_per_patch) x "
                              "patches per launch",
            "timing": "CUDA events around 5 launches of the kernel alone on the model's layer shape (random tensors "
                      "through the layer-level C ABI), after one warm-up launch",
            "note": ("this kernel does its multiply-adds on the tensor cores (3xTF32, activation operand in tensor memory); "
                     "it is bound by neither roof yet -- ncu: issue slots of the gather / split warps 65 % busy, tensor pipe "
                     "11 %, DRAM 43 % -- HBM is the nearer one and the one reported") if tensor_path else None}
        line["whole_step"] = {"hbm_frac_algorithmic": value / world * ENC_BYTES / 1e9 / hbm_peak,
                              "fp32_frac_algorithmic_flops": value / world * ENC_FLOPS_ALGO / 1e12 / fp32_tf,
                              "fp32_frac_executed_flops": value / world * ENC_FLOPS_CUDA_CORE / 1e12 / fp32_tf,
                              "flops_per_patch_executed": ENC_FLOPS_CUDA_CORE,
                              "flops_per_patch_direct_form_folded_head": ENC_FLOPS_EXEC,
                              "note": "executed = FP32 FLOPs left on the CUDA cores: the composite head only; every layer "
                                      "behind it and the code search run on the tensor cores (fractions above 1 of the FP32 "
                                      "roof by the reference's FLOP count are therefore expected)",
                              "bytes_per_patch": ENC_BYTES, "flops_per_patch": ENC_FLOPS_ALGO}
        line["layers"] = rows
        if world == 1:
            try:
                bf16 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1590.0)
            except Exception:
                bf16 = 1590.0
            line["wide_config"] = wide_config_table(dev, bf16)
            try:
                line["pca_projection"] = pca_table(dev, hbm_peak)
            except Exception as ex:
                line["pca_projection"] = {"unavailable": repr(ex)[:200]}
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cores)     # the CPU baseline gets every host core back
            v, cores, sec = cpu_reference_rate(1024, 256, 2)
            line["cpu_baseline"] = {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
                                    "sample": "1024 patches, batched eval enc+vq (B=256), best of 2 after 1 warm-up"}
            try:
                line["cpu_baseline"]["other_legs"] = cpu_legs()
                line["train_step"]["cpu_baseline"] = line["cpu_baseline"]["other_legs"]["train_step_b256"]
            except Exception as ex:
                line["cpu_baseline"]["other_legs"] = {"unavailable": repr(ex)[:200]}
            try:
                line["library_gpu_baseline"] = library_gpu_rate(dev)
            except Exception as ex:       # a baseline must never take the bench line down
                line["library_gpu_baseline"] = {"unavailable": repr(ex)[:200]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
