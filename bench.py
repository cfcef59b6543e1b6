#!/usr/bin/env python
"""bench.py -- MSDA forward+backward throughput on B200, with roofline, CPU baseline and end-to-end numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (N>1: launched by torch.distributed.run)
    python bench.py --impl reference [...]                        # the reference's CPU route on the host cores

Workload (BASELINE.json configs[1], the configuration the reference's published numbers are quoted on,
scripts/benchmark.py:25-31): B=4 images per GPU, Q=10 000 queries, H=8, D=32, L=4 (64^2,32^2,16^2,8^2), K=4, fp32,
padding_mode="border", align_corners=True, synthetic seeded inputs (img~N(0,1), points~U[0,1), weights=softmax(N(0,1))
over K, grad_out~U[0,1)).  One STEP = one forward + one backward (all three gradients) over that batch.
Multi-GPU: the path shards by batch with no data-path collective, every rank owns B=4 images -> weak
scaling; value = all ranks' queries / max-over-ranks device time.  At EVERY N the line also carries (extra.*):
the B=64 encoder training shape batch-sharded 64/N images per rank (strong scaling), the DETR encoder B=2
query-sharded over all N ranks with the grad_img reduce-scatter (checked against the unsharded operator in the
warm-up), a pinned-memcpy duplex probe run by all ranks at once (the host-side ceiling of `e2e`), and the CPU baseline.

Timing: every step is bracketed by CUDA events on the launching stream; a 256 MiB buffer is overwritten between
steps OUTSIDE the event pair so each step starts with a cold L2 (the reference's do_bench does the same).  The K steps
as a whole are bracketed by barrier + torch.cuda.synchronize().
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "msda-triton_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import torch  # noqa: E402

BENCH_PYRAMID = [(64, 64), (32, 32), (16, 16), (8, 8)]
DETR_PYRAMID = [(100, 167), (50, 84), (25, 42), (13, 21)]
WORKLOADS = {
    # name: (B, Q, H, D, pyramid, K, padding, align)
    "bench_q10k_border": (4, 10000, 8, 32, BENCH_PYRAMID, 4, "border", True),
    "bench_q10k_zeros": (4, 10000, 8, 32, BENCH_PYRAMID, 4, "zeros", False),
    "detr_encoder_zeros": (2, 22223, 8, 32, DETR_PYRAMID, 4, "zeros", False),
    "train_b64_encoder_zeros": (64, 22223, 8, 32, DETR_PYRAMID, 4, "zeros", False),   # BASELINE configs[4], whole batch
    "readme_q900_zeros": (2, 900, 8, 32, BENCH_PYRAMID, 4, "zeros", False),            # BASELINE configs[0] (README example)
    # encoder self-attention with realistic locality: query q sits on pixel q of the pyramid and samples
    # N(0, 2 px) around its own normalised position on every level (SURVEY.md 8d, M2 "encoder-realistic" variant)
    "detr_encoder_local_zeros": (2, 22223, 8, 32, DETR_PYRAMID, 4, "zeros", False),
    # encoder self-attention of a freshly initialised Deformable-DETR: query q sits on pixel q and every query uses the
    # SAME offsets -- point k of head h at (k+1) pixels (of each level) along direction 2 pi h / H, the model's bias
    # initialisation -- plus N(0, 0.1 px) of per-query variation: neighbouring queries share cells on the coarser levels
    "detr_encoder_init_zeros": (2, 22223, 8, 32, DETR_PYRAMID, 4, "zeros", False),
}
HEADLINE = "bench_q10k_border"
METRIC = "MSDA fwd+bwd throughput, 10k-query benchmark shape (fp32)"
UNIT = "queries/s"


def workload_string(name=None):
    """config.workload -- the SAME string in both arms (the driver compares them)."""
    name = name or HEADLINE
    B, Q, H, D, pyr, K, pm, ac = WORKLOADS[name]
    return (f"{name}: B={B} images per GPU, Q={Q} H={H} D={D} L={len(pyr)} K={K} pyramid="
            f"{'x'.join(str(h) for h, _ in pyr)} rows, {pm}/align_corners={ac}, fp32, fwd+bwd (all 3 grads)")


def make_inputs_device(name, seed, batch=None):
    """Same distributions as make_inputs(), generated ON the current CUDA device (large batches: B=64 is 5 GB)."""
    B, Q, H, D, pyr, K, pm, ac = WORKLOADS[name]
    B = B if batch is None else batch
    g = torch.Generator(device="cuda").manual_seed(seed)
    L = len(pyr)
    npix = sum(h * w for h, w in pyr)
    d = dict(device="cuda", generator=g)
    t = {
        "img": torch.randn(B, npix, H, D, **d),
        "pts": torch.rand(B, Q, H, L, K, 2, **d),
        "aw": torch.softmax(torch.randn(B, Q, H, L, K, **d), dim=-1),
        "go": torch.rand(B, Q, H, D, **d),
    }
    return t, torch.tensor(pyr, dtype=torch.int64, device="cuda")


def make_inputs(name, seed, device="cpu", pin=False):
    B, Q, H, D, pyr, K, pm, ac = WORKLOADS[name]
    g = torch.Generator().manual_seed(seed)
    L = len(pyr)
    npix = sum(h * w for h, w in pyr)
    pts = torch.rand(B, Q, H, L, K, 2, generator=g)
    if "init" in name:
        import math
        centres = []
        for (h, w) in pyr:
            ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            centres.append(torch.stack(((xs.reshape(-1) + 0.5) / w, (ys.reshape(-1) + 0.5) / h), dim=-1))
        ref = torch.cat(centres)[:Q]                                                     # [Q, 2]
        wh = torch.tensor([[w, h] for (h, w) in pyr], dtype=torch.float32)               # per level (w, h)
        theta = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
        direction = torch.stack((theta.cos(), theta.sin()), -1)
        direction = direction / direction.abs().max(-1, keepdim=True).values             # [H, 2], as the model's init
        steps = torch.arange(1, K + 1, dtype=torch.float32)                              # point k: (k + 1) pixels
        off_px = direction[:, None, None, :] * steps[None, None, :, None]                # [H, 1, K, 2]
        off_px = off_px + torch.randn(B, Q, H, L, K, 2, generator=g) * 0.1
        pts = ref[None, :, None, None, None, :] + off_px / wh[None, None, None, :, None, :]
    elif "local" in name:
        # reference point = centre of the query's own pixel (queries enumerate the pyramid in storage order)
        centres = []
        for (h, w) in pyr:
            ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            centres.append(torch.stack(((xs.reshape(-1) + 0.5) / w, (ys.reshape(-1) + 0.5) / h), dim=-1))
        ref = torch.cat(centres)[:Q]                                                     # [Q, 2]
        wh = torch.tensor([[w, h] for (h, w) in pyr], dtype=torch.float32)               # per level (w, h)
        off = torch.randn(B, Q, H, L, K, 2, generator=g) * 2.0 / wh[None, None, None, :, None, :]
        pts = ref[None, :, None, None, None, :] + off
    t = {
        "img": torch.randn(B, npix, H, D, generator=g),
        "pts": pts,
        "aw": torch.softmax(torch.randn(B, Q, H, L, K, generator=g), dim=-1),
        "go": torch.rand(B, Q, H, D, generator=g),
    }
    shapes = torch.tensor(pyr, dtype=torch.int64)
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    elif device != "cpu":
        t = {k: v.to(device) for k, v in t.items()}
        shapes = shapes.to(device)
    return t, shapes


def byte_model(name, unique_rows=None):
    """Algorithmic bytes (SURVEY.md 8d / BASELINE.md 3), fp32."""
    B, Q, H, D, pyr, K, _, _ = WORKLOADS[name]
    L, e = len(pyr), 4
    npix = sum(h * w for h, w in pyr)
    V = B * npix * H * D * e
    U = V if unique_rows is None else unique_rows * D * e
    S = B * Q * H * L * K * 2 * e
    A = B * Q * H * L * K * e
    O = B * Q * H * D * e  # noqa: E741
    G = B * Q * H * L * K * 4 * D * e
    return {"fwd": U + S + A + O, "bwd": (O + U + S + A) + (V + S + A), "gather": G, "V": V, "U": U}


def count_unique_rows(name, t, shapes):
    """Distinct (b, pixel, h) rows the bilinear corners touch (torch ops, setup only; not on any timed path)."""
    B, Q, H, D, pyr, K, pm, ac = WORKLOADS[name]
    pts = t["pts"].double()
    dev = pts.device
    npix = sum(h * w for h, w in pyr)
    hw = torch.tensor(pyr, dtype=torch.float64, device=dev)
    wh = hw.flip(-1)[None, None, None, :, None, :]
    xy = pts * (wh - 1) if ac else pts * wh - 0.5
    x0 = xy.floor()
    offs = torch.tensor([0] + [h * w for h, w in pyr[:-1]], device=dev).cumsum(0)[None, None, None, :, None]
    wi = hw[:, 1].long()[None, None, None, :, None]
    hi = hw[:, 0].long()[None, None, None, :, None]
    b = torch.arange(B, device=dev)[:, None, None, None, None]
    h = torch.arange(H, device=dev)[None, None, :, None, None]
    keys = []
    for dy in (0, 1):
        for dx in (0, 1):
            xi = (x0[..., 0].long() + dx)
            yi = (x0[..., 1].long() + dy)
            valid = (xi >= 0) & (xi < wi) & (yi >= 0) & (yi < hi)
            xi = xi.clamp(min=0).minimum(wi - 1)
            yi = yi.clamp(min=0).minimum(hi - 1)
            key = (b * npix + offs + yi * wi + xi) * H + h
            keys.append(key[valid] if pm == "zeros" else key.reshape(-1))
    return int(torch.unique(torch.cat(keys)).numel())


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:  # noqa: BLE001
            return local_index
    return local_index


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU route (torch grid_sample per level) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_route_sample(name, reps, warm=1, budget_s=60.0):
    """Times fwd+bwd of the reference's native-torch route (port: oracle/grid_sample_port.py of frontend.py:15-68) with
    all host threads on the WHOLE batch of the workload (a step of the CPU arm is the same step as ours); the number of
    repetitions is bounded by `budget_s` seconds.  One image alone (B=1, grid_sample batch N = H) is timed beside it.
    Returns the cpu_baseline dict; value = queries/s of the whole-batch median."""
    from oracle import grid_sample_port as port
    B, Q, H, D, pyr, K, pm, ac = WORKLOADS[name]
    t, shapes = make_inputs(name, seed=0)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)

    def timed(tensors, n, w):
        times = []
        for i in range(w + n):
            t0 = time.perf_counter()
            port.forward_backward(tensors["img"], shapes, tensors["pts"], tensors["aw"], tensors["go"], pm, ac)
            dt = time.perf_counter() - t0
            if i >= w:
                times.append(dt)
            if i == 0:
                n = max(1, min(n, int(budget_s / max(dt, 1e-3)) - w))   # bound the sample by the time budget
            if len(times) >= n:
                break
        times.sort()
        return times[len(times) // 2], len(times)

    med, n = timed(t, reps, warm)
    one = {k: v[:1].contiguous() for k, v in t.items()}
    med1, n1 = timed(one, min(reps, 5), 1)
    return {"value": B * Q / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"fwd+bwd of the whole batch (B={B}, Q={Q}), median of {n} after {warm} warm-up",
            "ms_per_step": med * 1e3,
            "one_image": {"value": Q / med1, "unit": UNIT, "ms": med1 * 1e3,
                          "sample": f"fwd+bwd of 1 of the {B} images (B=1), median of {n1}"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, Q, H, D, pyr, K, pm, ac = WORKLOADS[HEADLINE]
    cpu = cpu_route_sample(HEADLINE, reps=max(1, args.steps), warm=max(1, args.warmup))
    qps = cpu["value"]
    line = {
        "impl": "reference",
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(), "l2": "n/a (CPU)",
                   "note": "each step = fwd+bwd of the whole batch on the host cores (torch grid_sample route); the number "
                           "of timed steps is bounded to about a minute"},
        "cpu_baseline": cpu,
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def time_workload(name, steps, warmup, K, flush, dist_sync=None, clock_index=None, cold=True):
    """Returns dict with per-step mean ms for fwd, bwd, total (device events; cold L2 unless cold=False)."""
    B, Q, H, D, pyr, Kp, pm, ac = WORKLOADS[name]
    t, shapes = make_inputs(name, seed=0, device="cuda")
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]

    def one_step(e=None):
        if cold:
            flush.fill_(1.0)                  # cold L2: 256 MiB written between steps, outside the event pair
        if e:
            e[0].record()
        out = K.b200_multi_scale_deformable_attention_fwd(t["img"], shapes, t["pts"], t["aw"], pm, ac)
        if e:
            e[1].record()
        grads = K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], shapes, t["pts"], t["aw"], pm, ac)
        if e:
            e[2].record()
        return out, grads

    for _ in range(warmup):
        one_step()
    sampler = ClockSampler(clock_index) if clock_index is not None else None
    if dist_sync:
        dist_sync()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    if sampler:
        sampler.__enter__()
    for i in range(steps):
        one_step(ev[i])
    torch.cuda.synchronize()
    if sampler:
        sampler.__exit__()
    if dist_sync:
        dist_sync()
    wall = time.perf_counter() - wall0
    fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    res = {"fwd_ms": fwd, "bwd_ms": bwd, "step_ms": fwd + bwd, "wall_s": wall,
           "clocks": sampler.summary() if sampler else None}
    # exact count of distinct touched rows (U); for the B=64 workload U = V is assumed (every row is touched)
    res["unique_rows"] = count_unique_rows(name, t, shapes) if B < 32 else None
    return res


def time_e2e(name, steps, warmup, dist_sync=None, pipelined=True, chunks=None):
    """End to end with HOST buffers: every step copies its inputs (img, points, weights, grad_out) from pinned host
    memory to the device, runs forward + backward, and copies out + the three gradients back to pinned host memory --
    all inside the timed region.
    pipelined=True : the package's host-buffer entry point msda_triton.host.HostMsda (per-image chunks, H2D / kernels /
                     D2H overlapped on three streams).
    pipelined=False: plain user code around the autograd op (tensor.to('cuda') ... out.backward(go) ... .cpu())."""
    import msda_triton
    from msda_triton.host import HostMsda
    B, Q, H, D, pyr, Kp, pm, ac = WORKLOADS[name]
    host, shapes = make_inputs(name, seed=0, pin=True)
    shapes_dev = shapes.cuda()
    h_out = torch.empty(B, Q, H, D).pin_memory()
    h_gi = torch.empty_like(host["img"]).pin_memory()
    h_gp = torch.empty_like(host["pts"]).pin_memory()
    h_ga = torch.empty_like(host["aw"]).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = sum(v.numel() * v.element_size() for v in (h_out, h_gi, h_gp, h_ga))
    pipe = HostMsda(B, host["img"].shape[1], H, D, Q, len(pyr), Kp, chunks=chunks) if pipelined else None

    def one_step():
        if pipelined:
            pipe.run(host["img"], shapes_dev, host["pts"], host["aw"], pm, ac, h_out, host["go"], h_gi, h_gp, h_ga)
            return
        img = host["img"].to("cuda", non_blocking=True).requires_grad_(True)
        pts = host["pts"].to("cuda", non_blocking=True).requires_grad_(True)
        aw = host["aw"].to("cuda", non_blocking=True).requires_grad_(True)
        go = host["go"].to("cuda", non_blocking=True)
        out = msda_triton.multiscale_deformable_attention(img, shapes_dev, pts, aw, pm, ac)
        out.backward(go)
        h_out.copy_(out.detach(), non_blocking=True)
        h_gi.copy_(img.grad, non_blocking=True)
        h_gp.copy_(pts.grad, non_blocking=True)
        h_ga.copy_(aw.grad, non_blocking=True)

    for _ in range(warmup):
        one_step()
    if dist_sync:
        dist_sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_step()
    if pipelined:
        torch.cuda.current_stream().wait_stream(pipe.d2h)   # the timed region ends when the last result is on the host
    e1.record()
    torch.cuda.synchronize()
    if dist_sync:
        dist_sync()
    return e0.elapsed_time(e1) / steps, h2d, d2h


def time_module(flush, steps=20):
    """BASELINE configs[3]: Grounding-DINO decoder through the nn.Module -- B=8, Q=900, emb=hidden=256, H=8, L=4, K=4,
    border / align_corners=True, bf16 parameters and inputs; fwd+bwd of the whole module and of the op inside it."""
    from msda_triton import MultiscaleDeformableAttention, kernels as K
    B, Q, emb, H, L, Kp = 8, 900, 256, 8, 4, 4
    res = {}
    for pname, pyr in (("pyramid_5440", BENCH_PYRAMID), ("pyramid_22223", DETR_PYRAMID)):
        npix = sum(h * w for h, w in pyr)
        g = torch.Generator().manual_seed(0)
        dt = torch.bfloat16
        img = torch.randn(B, npix, emb, generator=g).to("cuda", dt).requires_grad_(True)
        queries = torch.randn(B, Q, emb, generator=g).to("cuda", dt).requires_grad_(True)
        ref = torch.rand(B, Q, 2, generator=g).to("cuda", dt)
        shapes = torch.tensor(pyr, device="cuda")
        mod = MultiscaleDeformableAttention(emb, emb, L, H, Kp, "border", True).to("cuda", dt)
        gout = torch.rand(B, Q, emb, generator=g).to("cuda", dt)

        def step():
            out = mod(img, shapes, queries, ref)
            out.backward(gout)

        # the op alone on module-like operands
        v = torch.randn(B, npix, H, emb // H, generator=g).to("cuda", dt)
        pts = torch.rand(B, Q, H, L, Kp, 2, generator=g).to("cuda", dt)
        aw = torch.softmax(torch.randn(B, Q, H, L * Kp, generator=g), -1).reshape(B, Q, H, L, Kp).to("cuda", dt)
        go = torch.rand(B, Q, H, emb // H, generator=g).to("cuda", dt)

        def op_step():
            K.b200_multi_scale_deformable_attention_fwd(v, shapes, pts, aw, "border", True)
            K.b200_multi_scale_deformable_attention_bwd(go, v, shapes, pts, aw, "border", True)

        out = {}
        for label, fn in (("module_fwd_bwd_ms", step), ("op_fwd_bwd_ms", op_step)):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(steps):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            out[label] = ts[len(ts) // 2]
        res[pname] = out
    return res


def l2_probe(K_lib, kinds=("gather", "scatter")):
    """Random 128-byte-row gather / red.add.v4 throughput over an L2-resident 32 MiB buffer (the access shape of the
    kernels) -- the 'L2 gather roof' SURVEY.md 8d asks to measure next to every result."""
    import ctypes
    lib = K_lib
    rows_buf = (32 << 20) // 128
    buf = torch.zeros(rows_buf * 32, device="cuda")
    sink = torch.zeros(4, device="cuda")
    n = 64 * 1024 * 1024 // 8   # 8.4M rows = 1 GiB of row traffic
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for kind in kinds:
        def launch(seed):
            if kind == "gather":
                return lib.msda_probe_gather(ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(buf.data_ptr()),
                                             rows_buf, n, seed, st)
            return lib.msda_probe_scatter(ctypes.c_void_p(buf.data_ptr()), rows_buf, n, seed, st)
        for s in range(3):
            assert launch(s) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(5):
            launch(10 + s)
        e1.record()
        torch.cuda.synchronize()
        res[kind + "_gbs"] = n * 128 / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e9
    return res


def shard_range(total, rank, world):
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def time_strong_b64(rank, world, K, flush, sync, steps=5, warmup=3):
    """BASELINE configs[4]: training fwd+bwd on the B=64 encoder shapes, batch-sharded 64/N images per rank (strong
    scaling, no data-path collective).  Inputs are generated on the device (5 GB at N=1)."""
    name = "train_b64_encoder_zeros"
    B, Q, H, D, pyr, Kp, pm, ac = WORKLOADS[name]
    lo, hi = shard_range(B, rank, world)
    t, shapes = make_inputs_device(name, seed=100 + rank, batch=hi - lo)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]

    def one(e=None):
        flush.fill_(1.0)
        if e:
            e[0].record()
        K.b200_multi_scale_deformable_attention_fwd(t["img"], shapes, t["pts"], t["aw"], pm, ac)
        if e:
            e[1].record()
        K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], shapes, t["pts"], t["aw"], pm, ac)
        if e:
            e[2].record()

    for _ in range(warmup):
        one()
    sync()
    for i in range(steps):
        one(ev[i])
    torch.cuda.synchronize()
    sync()
    fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    del t
    torch.cuda.empty_cache()
    return fwd, bwd, hi - lo


def time_query_sharded(rank, world, dist, flush, sync, steps=20, warmup=5):
    """DETR encoder B=2 with the queries sharded over ALL ranks (they share both images): forward all-gathers the pixel
    shards of `img`, backward reduce-scatters grad_img (msda_triton.distributed.query_sharded_msda).  The warm-up checks
    the sharded results against the unsharded operator run on the same rank."""
    import msda_triton
    from msda_triton import distributed as D
    name = "detr_encoder_zeros"
    B, Q, H, Dh, pyr, Kp, pm, ac = WORKLOADS[name]
    t, shapes = make_inputs(name, seed=0, device="cuda")          # the same full problem on every rank
    npix = t["img"].shape[1]
    res = {"workload": workload_string(name).replace("images per GPU", "images shared by all ranks"),
           "queries_per_rank": shard_range(Q, rank, world)[1] - shard_range(Q, rank, world)[0]}
    if world == 1:
        a, b, c = (t[k].clone().requires_grad_(True) for k in ("img", "pts", "aw"))

        def step():
            out = msda_triton.multiscale_deformable_attention(a, shapes, b, c, pm, ac)
            out.backward(t["go"])
            a.grad = b.grad = c.grad = None
        res["note"] = "N=1: the unsharded operator through the same autograd entry point, no collective"
    else:
        shard = D.shard_pixels(t["img"], rank, world).clone().requires_grad_(True)
        qlo, qhi = shard_range(Q, rank, world)
        pts = t["pts"][:, qlo:qhi].contiguous().requires_grad_(True)
        aw = t["aw"][:, qlo:qhi].contiguous().requires_grad_(True)
        go = t["go"][:, qlo:qhi].contiguous()

        def step(keep=False):
            out = D.query_sharded_msda(shard, npix, shapes, pts, aw, pm, ac)
            out.backward(go)
            g = (out.detach(), shard.grad, pts.grad, aw.grad) if keep else None
            shard.grad = pts.grad = aw.grad = None
            return g

        # ---- equivalence (warm-up): sharded == unsharded on the same inputs ----
        a, b, c = (t[k].clone().requires_grad_(True) for k in ("img", "pts", "aw"))
        ref = msda_triton.multiscale_deformable_attention(a, shapes, b, c, pm, ac)
        ref.backward(t["go"])
        got = step(keep=True)
        want = (ref.detach()[:, qlo:qhi], D.shard_pixels(a.grad, rank, world), b.grad[:, qlo:qhi], c.grad[:, qlo:qhi])
        errs = torch.tensor([float((g - w).abs().max() / w.abs().max().clamp_min(1e-30)) for g, w in zip(got, want)],
                            device="cuda", dtype=torch.float64)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        res["sharded_vs_unsharded_max_err_over_max"] = dict(zip(("out", "grad_img_shard", "grad_points", "grad_weights"),
                                                                 errs.tolist()))
        res["equivalent"] = bool(errs[0] == 0 and errs[2] <= 1e-5 and errs[3] <= 1e-5 and errs[1] <= 1e-4)
        del a, b, c, ref, got, want
    def timed(fn):
        for _ in range(warmup):
            fn()
        sync()
        ts = []
        for _ in range(steps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        sync()
        ts.sort()
        ms = torch.tensor([ts[len(ts) // 2]], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if world == 1:
        res["fwd_bwd_ms"] = timed(step)
    else:
        res["nccl"] = {"fwd_bwd_ms": timed(step),
                       "how": "all_gather_into_tensor / reduce_scatter_tensor per image, zero-copy (distributed.query_sharded_msda)"}
        # ---- the same step over NVLink peer memory: the library's own all-gather / reduce-scatter kernels ----
        try:
            ex = D.PeerPixelExchange(B, npix, H, Dh)

            def peer_step(keep=False):
                out = D.peer_query_sharded_msda(ex, shard, shapes, pts, aw, pm, ac)
                out.backward(go)
                g = (out.detach(), shard.grad, pts.grad, aw.grad) if keep else None
                shard.grad = pts.grad = aw.grad = None
                return g

            got = peer_step(keep=True)
            base = step(keep=True)
            errs = torch.tensor([float((g - w).abs().max() / w.abs().max().clamp_min(1e-30)) for g, w in zip(got, base)],
                                device="cuda", dtype=torch.float64)
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            res["peer_memory"] = {
                "fwd_bwd_ms": timed(peer_step),
                "vs_nccl_route_max_err_over_max": dict(zip(("out", "grad_img_shard", "grad_points", "grad_weights"),
                                                           errs.tolist())),
                "how": "msda_peer_all_gather / msda_peer_reduce_scatter (csrc/msda_peer.cu): P2P loads over NVLink from "
                       "symmetric memory, one launch per collective, the backward accumulates in the buffer the peers read"}
            # the whole step (all-gather, forward, backward, reduce-scatter) as ONE CUDA graph: the peer collectives keep
            # their call counts on the device, so nothing in the step depends on the host
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        peer_step()
                torch.cuda.current_stream().wait_stream(side)
                sync()
                graph = torch.cuda.CUDAGraph()
                shard.grad = pts.grad = aw.grad = None
                with torch.cuda.graph(graph):
                    out_g = D.peer_query_sharded_msda(ex, shard, shapes, pts, aw, pm, ac)
                    out_g.backward(go)
                graph.replay()
                sync()
                got_g = (out_g.detach(), shard.grad, pts.grad, aw.grad)
                errs_g = torch.tensor([float((g - w).abs().max() / w.abs().max().clamp_min(1e-30))
                                       for g, w in zip(got_g, base)], device="cuda", dtype=torch.float64)
                dist.all_reduce(errs_g, op=dist.ReduceOp.MAX)
                res["peer_memory"]["cuda_graph"] = {
                    "fwd_bwd_ms": timed(graph.replay),
                    "vs_nccl_route_max_err_over_max": dict(zip(("out", "grad_img_shard", "grad_points", "grad_weights"),
                                                               errs_g.tolist()))}
                shard.grad = pts.grad = aw.grad = None
                del graph
            except Exception as ex_g:  # noqa: BLE001
                res["peer_memory"]["cuda_graph"] = {"unavailable": f"{type(ex_g).__name__}: {ex_g}"}
            sh_det = shard.detach()
            res["peer_memory"]["all_gather_ms"] = timed(lambda: ex.all_gather(sh_det))
            res["peer_memory"]["reduce_scatter_ms"] = timed(lambda: ex.reduce_scatter())
            res["fwd_bwd_ms"] = min(res["nccl"]["fwd_bwd_ms"], res["peer_memory"]["fwd_bwd_ms"],
                                    res["peer_memory"]["cuda_graph"].get("fwd_bwd_ms", float("inf")))
        except Exception as ex_:  # noqa: BLE001
            res["peer_memory"] = {"unavailable": f"{type(ex_).__name__}: {ex_}"}
            res["fwd_bwd_ms"] = res["nccl"]["fwd_bwd_ms"]
    if world > 1:
        # the two collectives alone, same buffers / message sizes as inside the step
        chunk = D.pixel_chunk(npix, world)
        full = torch.empty(B, world * chunk, H, Dh, device="cuda")
        sh = shard.detach()
        outb = torch.empty_like(sh)
        colls = {}
        for label, fn in (("all_gather_into_tensor", lambda bb: dist.all_gather_into_tensor(full[bb], sh[bb])),
                          ("reduce_scatter_tensor", lambda bb: dist.reduce_scatter_tensor(outb[bb], full[bb]))):
            for _ in range(3):
                for bb in range(B):
                    fn(bb)
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                for bb in range(B):
                    fn(bb)
            e1.record()
            torch.cuda.synchronize()
            m = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            colls[label + "_ms"] = float(m)
        colls["message_bytes_per_collective"] = int(full.numel() * 4)
        colls["backend"] = "NCCL " + ".".join(str(v) for v in torch.cuda.nccl.version())
        res["collectives"] = colls
    return res


def pcie_probe(h2d_bytes, d2h_bytes, dist, sync, reps=10):
    """Pure pinned-memory copies of the e2e step's byte counts, host->device and device->host concurrently on two
    streams, ALL RANKS AT THE SAME TIME: the host-side ceiling of `e2e` at this N (no kernels, no staging logic)."""
    h_in = torch.empty(h2d_bytes // 4).pin_memory()
    h_out = torch.empty(d2h_bytes // 4).pin_memory()
    d_in = torch.empty(h2d_bytes // 4, device="cuda")
    d_out = torch.empty(d2h_bytes // 4, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    for _ in range(3):
        both()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        both()
    for st in (s1, s2):
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def load_reference_package():
    """The UNMODIFIED reference package staged into baseline/_ref (scripts/stage_reference.py), under the alias
    ref_msda_triton (ours owns the name msda_triton).  None when it is not there (or Triton cannot be imported)."""
    import importlib.util
    pkg = ROOT / "baseline" / "_ref" / "msda_triton"
    if not pkg.is_dir():
        return None, "baseline/_ref not staged"
    try:
        if str(pkg.parent) not in sys.path:
            sys.path.append(str(pkg.parent))   # at the END: only so that importlib.metadata finds the dist-info
        spec = importlib.util.spec_from_file_location("ref_msda_triton", pkg / "__init__.py",
                                                      submodule_search_locations=[str(pkg)])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_msda_triton"] = mod
        spec.loader.exec_module(mod)
        from ref_msda_triton import frontend
        return frontend, None
    except Exception as ex:  # noqa: BLE001
        return None, f"{type(ex).__name__}: {ex}"


def reference_triton_gpu(flush, reps=20):
    """BASELINE.md section 2's on-box bar: the reference's own Triton kernels, JIT-compiled for sm_100a, timed in this
    run with the same harness (CUDA events, cold L2, public functional API, fwd under no_grad and fwd + out.backward),
    next to ours, plus the max error of ours against them."""
    ref, why = load_reference_package()
    if ref is None:
        return {"unavailable": why}
    from msda_triton.frontend import b200_multiscale_deformable_attention as ours_fn
    ref_fn = ref.triton_multiscale_deformable_attention
    out = {}
    try:
        import triton
        out["triton"] = triton.__version__
    except Exception:  # noqa: BLE001
        pass
    for name in (HEADLINE, "detr_encoder_zeros"):
        B, Q, H, D, pyr, Kp, pm, ac = WORKLOADS[name]
        t, shapes = make_inputs(name, seed=0, device="cuda")
        row = {}

        def grads_of(fn):
            a, b, c = (t[k].clone().requires_grad_(True) for k in ("img", "pts", "aw"))
            o = fn(a, shapes, b, c, pm, ac)
            o.backward(t["go"])
            return [o.detach().double(), a.grad.double(), b.grad.double(), c.grad.double()]

        mine, theirs = grads_of(ours_fn), grads_of(ref_fn)
        row["ours_vs_reference_max_err_over_max"] = {
            what: float((m - r).abs().max() / r.abs().max().clamp_min(1e-30))
            for what, m, r in zip(("out", "grad_img", "grad_points", "grad_weights"), mine, theirs)}
        # grad_sampling_points is piecewise constant in the cell index: the JIT-compiled Triton kernel contracts
        # x*w - 0.5 into an FMA, the reference's written op order (and ours, and the CPU oracle) rounds twice, so a point
        # within 1 ulp of a cell boundary lands in the neighbouring cell there -- isolated elements, counted here
        gp_m, gp_r = mine[2], theirs[2]
        bad = (gp_m - gp_r).abs() > 1e-4 * gp_r.abs() + 1e-5 * gp_r.abs().max()
        row["grad_points_elements_outside_tolerance"] = {"count": int(bad.sum()), "of": int(bad.numel())}
        del mine, theirs, gp_m, gp_r, bad
        for who, fn in (("reference_triton", ref_fn), ("ours", ours_fn)):
            a, b, c = (t[k].clone().requires_grad_(True) for k in ("img", "pts", "aw"))

            def fwd():
                with torch.no_grad():
                    fn(a, shapes, b, c, pm, ac)

            def fwdbwd():
                o = fn(a, shapes, b, c, pm, ac)
                o.backward(t["go"])
                a.grad = b.grad = c.grad = None

            for label, f in (("fwd_ms", fwd), ("fwd_bwd_ms", fwdbwd)):
                for _ in range(5):
                    f()
                ts = []
                for _ in range(reps):
                    flush.fill_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    f()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
                row.setdefault(who, {})[label] = ts[len(ts) // 2]
        row["speedup"] = {k: row["reference_triton"][k] / row["ours"][k] for k in ("fwd_ms", "fwd_bwd_ms")}
        out[name] = row
    out["harness"] = f"CUDA events, cold L2, 5 warm-up + {reps} reps, medians; fwd+bwd through autograd on both sides"
    return out


def time_deterministic(K, flush, steps=10):
    """Backward in the deterministic (sorted-segment) mode on the headline shape, next to the atomic mode."""
    B, Q, H, D, pyr, Kp, pm, ac = WORKLOADS[HEADLINE]
    t, shapes = make_inputs(HEADLINE, seed=0, device="cuda")
    out = {}
    for label, det in (("atomic_bwd_ms", False), ("deterministic_bwd_ms", True)):
        def f():
            return K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], shapes, t["pts"], t["aw"], pm, ac,
                                                               deterministic=det)
        for _ in range(3):
            g = f()
        ts = []
        for _ in range(steps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out[label] = ts[len(ts) // 2]
        if det:
            g2 = f()
            out["bit_reproducible"] = bool(torch.equal(g[0], g2[0]))
    out["ratio"] = out["deterministic_bwd_ms"] / out["atomic_bwd_ms"]
    out["how"] = "exact (quantised) fp32 row adds, csrc/msda_bwd_detq.cu: amax + row-bound pre-passes, then the tuned backward"
    # the sorted-segment path (radix sort + ordered segment sums; full fp32 accuracy for any input) for comparison
    from msda_triton import _lib
    os.environ["MSDA_B200_DET_VARIANT"] = "0"
    _lib.reload_tuning()
    try:
        def f0():
            return K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], shapes, t["pts"], t["aw"], pm, ac,
                                                               deterministic=True)
        for _ in range(2):
            f0()
        ts = []
        for _ in range(5):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f0()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out["sorted_segment_bwd_ms"] = ts[len(ts) // 2]
    finally:
        os.environ.pop("MSDA_B200_DET_VARIANT", None)
        _lib.reload_tuning()
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MSDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep stdout to the one JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # host-side waiting (ranks 1.. idle while rank 0 times the CPU baseline): a gloo barrier blocks on a socket, an
        # NCCL barrier would spin one host core per waiting rank
        host_group = dist.new_group(backend="gloo")

    from msda_triton import _lib, kernels as K
    numa_cores = None
    if world > 1 and os.environ.get("MSDA_BENCH_NUMA_BIND", "1") != "0":
        # one process per GPU: keep each rank's pinned host buffers (e2e) in the memory next to its GPU
        from msda_triton.host import bind_host_thread_to_gpu
        numa_cores = bind_host_thread_to_gpu(torch.device("cuda", local))

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 18, device="cuda")  # 256 MiB of fp32
    B, Q, H, D, pyr, Kp, pm, ac = WORKLOADS[HEADLINE]

    head = time_workload(HEADLINE, args.steps, args.warmup, K, flush, dist_sync=sync,
                         clock_index=physical_gpu_index(local))
    # sustained back-to-back steps: one chunk per call (fewest, largest copies; consecutive calls overlap H2D / D2H)
    e2e_ms, h2d, d2h = time_e2e(HEADLINE, max(3, args.steps), 3, dist_sync=sync, pipelined=True, chunks=1)
    e2e_plain_ms, _, _ = time_e2e(HEADLINE, max(3, min(args.steps, 10)), 3, dist_sync=sync, pipelined=False)

    # max over ranks of the device time
    tmax = torch.tensor([head["step_ms"], head["fwd_ms"], head["bwd_ms"], e2e_ms, e2e_plain_ms], device="cuda",
                        dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    step_ms, fwd_ms, bwd_ms, e2e_ms, e2e_plain_ms = tmax.tolist()

    extra = {}
    if not args.quick:
        # ---- collective sections: every rank takes part, at every N ----
        copy_ms = pcie_probe(h2d, d2h, dist, sync)
        extra["pcie_probe"] = {
            "ms_per_step_bytes": copy_ms, "h2d_gbs_per_rank": h2d / (copy_ms * 1e-3) / 1e9,
            "d2h_gbs_per_rank": d2h / (copy_ms * 1e-3) / 1e9,
            "aggregate_gbs_both_directions": world * (h2d + d2h) / (copy_ms * 1e-3) / 1e9,
            "e2e_ceiling_queries_per_s": world * B * Q / (copy_ms * 1e-3),
            "e2e_fraction_of_ceiling": copy_ms / e2e_ms,
            "what": "the e2e step's H2D and D2H byte counts as plain pinned cudaMemcpyAsync on two streams, all ranks "
                    "concurrently, max over ranks: the host-memory / PCIe ceiling of the e2e figure at this N"}
        f64, b64, imgs = time_strong_b64(rank, world, K, flush, sync)
        t64 = torch.tensor([f64, b64], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t64, op=dist.ReduceOp.MAX)
        f64, b64 = t64.tolist()
        B64, Q64 = WORKLOADS["train_b64_encoder_zeros"][:2]
        extra["train_b64_encoder_strong"] = {
            "workload": "train_b64_encoder_zeros: B=64 images IN TOTAL, batch-sharded 64/N per rank, Q=22223 (DETR "
                        "encoder pyramid), zeros/False, fp32, fwd+bwd, no collective",
            "images_per_rank": imgs, "fwd_ms": f64, "bwd_ms": b64, "fwd_bwd_ms": f64 + b64, "scaling": "strong",
            "queries_per_s": B64 * Q64 / ((f64 + b64) * 1e-3)}
        extra["detr_query_sharded"] = time_query_sharded(rank, world, dist, flush, sync)

    cpu = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        bm = byte_model(HEADLINE, head["unique_rows"])
        traffic = None
        tr = ROOT / "profiles" / "traffic.json"
        if tr.exists():
            try:
                traffic = json.loads(tr.read_text()).get("msda_bwd", {}).get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                traffic = None
        ach = bm["bwd"] / (bwd_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "msda_bwd (backward pass incl. its grad_img zero-fill)",
                    "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "algorithmic_bytes": bm["bwd"],
                    "fwd": {"achieved": bm["fwd"] / (fwd_ms * 1e-3) / 1e9, "algorithmic_bytes": bm["fwd"],
                            "frac": bm["fwd"] / (fwd_ms * 1e-3) / 1e9 / peak},
                    "gather_traffic_bytes": bm["gather"],
                    "gather_gbs": {"fwd": bm["gather"] / (fwd_ms * 1e-3) / 1e9,
                                   "bwd_read_plus_atomic": 2 * bm["gather"] / (bwd_ms * 1e-3) / 1e9}}
        if not args.quick:
            try:
                extra["deterministic"] = time_deterministic(K, flush)
            except Exception as ex:  # noqa: BLE001
                extra["deterministic"] = {"error": str(ex)}
        if world == 1 and not args.quick:
            try:
                extra["l2_probe"] = l2_probe(_lib.get_lib())
                # roofline_time = max(compulsory / HBM peak, corner-row traffic / L2 peak); the L2 peaks
                # are measured live by the probes (their ncu captures: profiles/r2_probe_*.csv)
                g_peak, a_peak = extra["l2_probe"]["gather_gbs"], extra["l2_probe"]["scatter_gbs"]
                t_hbm_b, t_l2_b = bm["bwd"] / (peak * 1e9) * 1e3, bm["gather"] / (a_peak * 1e9) * 1e3
                t_hbm_f, t_l2_f = bm["fwd"] / (peak * 1e9) * 1e3, bm["gather"] / (g_peak * 1e9) * 1e3
                # every gathered row passes the SM's L1 data stage, one 128-byte wavefront per clock and SM (the unit ncu
                # shows at ~80 % for the forward, l1tex__data_pipe_lsu_wavefronts, profiles/r2_ncu_summary.md)
                props = torch.cuda.get_device_properties(local)
                sm_hz = ((head["clocks"] or {}).get("sm_mhz") or 1965) * 1e6
                l1_peak = props.multi_processor_count * 128 * sm_hz / 1e9
                t_l1_f = bm["gather"] / (l1_peak * 1e9) * 1e3
                roofline["survey_8d"] = {
                    "definition": "roofline_time = max(compulsory_bytes / HBM_peak, corner_row_bytes / on-chip peak of the "
                                  "unit the rows must cross); frac = roofline_time / measured_time",
                    "bwd": {"hbm_term_ms": t_hbm_b, "l2_atomic_term_ms": t_l2_b, "roofline_ms": max(t_hbm_b, t_l2_b),
                            "measured_ms": bwd_ms, "frac": max(t_hbm_b, t_l2_b) / bwd_ms,
                            "l2_peak_gbs": a_peak,
                            "l2_peak_source": "msda_probe_scatter: red.global.add.v4.f32 of random 128-byte rows over "
                                              "an L2-resident 32 MiB buffer, this run; ncu of the probe: the SM's "
                                              "L1TEX->XBAR request port at 84 % (profiles/r2_ncu_raw_probe_scatter.csv)"},
                    "fwd": {"hbm_term_ms": t_hbm_f, "l1_pipe_term_ms": t_l1_f, "roofline_ms": max(t_hbm_f, t_l1_f),
                            "measured_ms": fwd_ms, "frac": max(t_hbm_f, t_l1_f) / fwd_ms,
                            "l1_pipe_peak_gbs": l1_peak,
                            "l1_pipe_peak_source": f"{props.multi_processor_count} SMs x 128 B per clock x "
                                                   f"{sm_hz / 1e6:.0f} MHz (SM clock sampled under load in this run)",
                            "l2_gather_term_ms": t_l2_f, "l2_gather_peak_gbs": g_peak,
                            "l2_gather_peak_source": "msda_probe_gather: ld.global.nc.v4 of random 128-byte rows, every "
                                                     "row from L2, this run; ncu of the probe: L2->XBAR return path at "
                                                     "78-84 % (profiles/r2_ncu_raw_probe_gather.csv).  Applies to the rows "
                                                     "that miss L1 only (the forward serves about half of them from L1), "
                                                     "so it is reported, not used as the roof"},
                    "hbm_peak_gbs": peak, "hbm_peak_source": peak_src}
            except Exception as ex:  # noqa: BLE001
                extra["l2_probe"] = {"error": str(ex)}
            for name in WORKLOADS:
                if name == HEADLINE or WORKLOADS[name][0] >= 32:
                    continue
                r = time_workload(name, max(5, args.steps // 4), 3, K, flush)
                bmn = byte_model(name, r["unique_rows"])
                extra[name] = {"fwd_ms": r["fwd_ms"], "bwd_ms": r["bwd_ms"], "fwd_bwd_ms": r["step_ms"],
                               "fwd_hbm_frac": bmn["fwd"] / (r["fwd_ms"] * 1e-3) / 1e9 / peak,
                               "bwd_hbm_frac": bmn["bwd"] / (r["bwd_ms"] * 1e-3) / 1e9 / peak,
                               "fwd_gather_gbs": bmn["gather"] / (r["fwd_ms"] * 1e-3) / 1e9}
            warm = time_workload(HEADLINE, max(5, args.steps // 4), 3, K, flush, cold=False)
            extra["headline_warm_l2"] = {"fwd_ms": warm["fwd_ms"], "bwd_ms": warm["bwd_ms"],
                                         "note": "back-to-back steps without the L2 flush"}
            try:
                extra["gdino_decoder_module_bf16"] = time_module(flush)                      # fused module core
                os.environ["MSDA_B200_FUSED_MODULE"] = "0"
                extra["gdino_decoder_module_bf16_composed"] = time_module(flush)             # softmax etc. in torch
            except Exception as ex:  # noqa: BLE001
                extra["gdino_decoder_module_bf16"] = {"error": str(ex)}
            finally:
                os.environ.pop("MSDA_B200_FUSED_MODULE", None)
            try:
                extra["reference_triton_gpu"] = reference_triton_gpu(flush)
            except Exception as ex:  # noqa: BLE001
                extra["reference_triton_gpu"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
        if not args.quick:
            cpu = cpu_route_sample(HEADLINE, reps=10, warm=1)
        line = {
            "metric": METRIC, "value": world * B * Q / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(),
                       "l2": "flushed between steps (256 MiB overwrite outside the event pair)",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
            "clocks": head["clocks"],
            "e2e": {"value": world * B * Q / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_numa_bind": bool(numa_cores),
                    "api": "msda_triton.host.HostMsda(chunks=1).run -- pinned host buffers, two staging sets, H2D of "
                           "step i+1 overlapped with kernels / D2H of step i",
                    "autograd_unpipelined": {"value": world * B * Q / (e2e_plain_ms * 1e-3),
                                             "ms_per_step": e2e_plain_ms}},
            # per step: forward kernel + grad_img zero-fill (memset node) + backward kernel
            "gpu_launches": 3 * args.steps,
            "gpu_launches_note": "per step: msda_fwd_tiled_kernel, cudaMemsetAsync(grad_img), msda_bwd_tiled_kernel",
            "roofline": roofline,
            "cpu_baseline": cpu,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(group=host_group)     # ranks 1.. sleep on a socket while rank 0 runs the CPU baseline
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline only: skip extra workloads, probes, cpu baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
