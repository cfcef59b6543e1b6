"""Side by side on ONE B200: the reference's own Triton path (staged unmodified into baseline/_ref by
scripts/stage_reference.py) against this repository's CUDA path, through the same public functional API and the same
harness -- the measurement SURVEY.md section 8(d) asks for ("same harness times the reference Triton path on the same
box in the same run") and the GPU-side parity target of section 8(c).

For every workload: parity of out / grad_img / grad_points / grad_weights (max error relative to the largest reference
magnitude), then forward (no_grad) and forward+backward (out.backward(g), grads reset, as scripts/benchmark.py:90-94 of
the reference does), cold L2 (256 MiB overwrite between repetitions, outside the event pair), median / p20 / p80, plus
peak-memory deltas measured as the reference's memory benchmark does (benchmark.py:160-172).

    python scripts/stage_reference.py          # once, in the build container
    gpurun -- python scripts/bench_vs_reference_gpu.py [--reps 50] [--out gpurun_out/ref_vs_ours.json]

Not part of bench.py or the tests: baseline/_ref is measurement material only.
"""
import argparse
import importlib.util
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
REF_DIR = ROOT / "baseline" / "_ref"

WORKLOADS = {
    # name: (B, Q, H, D, pyramid, K, padding_mode, align_corners, dtype)
    "readme_q900_border": (2, 900, 8, 32, [(64, 64), (32, 32), (16, 16), (8, 8)], 4, "border", True, torch.float32),
    "bench_q10k_border": (4, 10000, 8, 32, [(64, 64), (32, 32), (16, 16), (8, 8)], 4, "border", True, torch.float32),
    "bench_q10k_zeros": (4, 10000, 8, 32, [(64, 64), (32, 32), (16, 16), (8, 8)], 4, "zeros", False, torch.float32),
    "detr_encoder_zeros": (2, 22223, 8, 32, [(100, 167), (50, 84), (25, 42), (13, 21)], 4, "zeros", False,
                           torch.float32),
    "bench_q10k_border_fp16": (4, 10000, 8, 32, [(64, 64), (32, 32), (16, 16), (8, 8)], 4, "border", True,
                               torch.float16),
}


def load_reference():
    """The reference package under the alias ref_msda_triton (ours already owns the name msda_triton)."""
    pkg = REF_DIR / "msda_triton"
    if not pkg.is_dir():
        raise SystemExit("baseline/_ref is empty: run scripts/stage_reference.py in the build container first")
    sys.path.append(str(REF_DIR))   # at the END: only so that importlib.metadata finds the dist-info
    spec = importlib.util.spec_from_file_location("ref_msda_triton", pkg / "__init__.py",
                                                  submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_msda_triton"] = mod
    spec.loader.exec_module(mod)
    from ref_msda_triton import frontend
    return frontend


def make(name):
    B, Q, H, D, pyr, K, pm, ac, dt = WORKLOADS[name]
    g = torch.Generator(device="cuda").manual_seed(0)
    L, npix = len(pyr), sum(h * w for h, w in pyr)
    img = torch.randn(B, npix, H, D, device="cuda", generator=g).to(dt)
    pts = torch.rand(B, Q, H, L, K, 2, device="cuda", generator=g).to(dt)
    aw = torch.softmax(torch.randn(B, Q, H, L, K, device="cuda", generator=g), -1).to(dt)
    go = torch.rand(B, Q, H, D, device="cuda", generator=g).to(dt)
    return img, torch.tensor(pyr, device="cuda"), pts, aw, go, pm, ac


def quantiles(fn, flush, reps, warmup):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return {"median_ms": ts[len(ts) // 2], "p20_ms": ts[len(ts) // 5], "p80_ms": ts[(4 * len(ts)) // 5]}


def peak_mb(fn):
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    before = torch.cuda.max_memory_allocated()
    fn()
    torch.cuda.synchronize()
    return (torch.cuda.max_memory_allocated() - before) / 2 ** 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "ref_vs_ours.json"))
    ns = ap.parse_args()
    ref = load_reference()
    from msda_triton.frontend import b200_multiscale_deformable_attention as ours_fn
    ref_fn = ref.triton_multiscale_deformable_attention
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    import triton
    report = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "triton": triton.__version__,
              "harness": f"CUDA events, cold L2, {ns.warmup} warm-up + {ns.reps} reps", "workloads": {}}
    for name in WORKLOADS:
        img, shapes, pts, aw, go, pm, ac = make(name)
        row = {}

        def grads_of(fn):
            a, b, c = (t.clone().requires_grad_(True) for t in (img, pts, aw))
            out = fn(a, shapes, b, c, pm, ac)
            out.backward(go)
            return [out.detach().double(), a.grad.double(), b.grad.double(), c.grad.double()]

        mine, theirs = grads_of(ours_fn), grads_of(ref_fn)
        row["parity_max_err_over_max_ref"] = {
            what: float((m - t).abs().max() / t.abs().max().clamp_min(1e-30))
            for what, m, t in zip(("out", "grad_img", "grad_points", "grad_weights"), mine, theirs)}
        del mine, theirs
        for who, fn in (("reference_triton", ref_fn), ("ours_cuda", ours_fn)):
            a, b, c = (t.clone().requires_grad_(True) for t in (img, pts, aw))

            def fwd():
                with torch.no_grad():
                    fn(a, shapes, b, c, pm, ac)

            def fwdbwd():
                out = fn(a, shapes, b, c, pm, ac)
                out.backward(go)
                a.grad = b.grad = c.grad = None

            row[who] = {"fwd": quantiles(fwd, flush, ns.reps, ns.warmup),
                        "fwd_bwd": quantiles(fwdbwd, flush, ns.reps, ns.warmup),
                        "fwd_peak_mb": peak_mb(fwd), "fwd_bwd_peak_mb": peak_mb(fwdbwd)}
        r, o = row["reference_triton"], row["ours_cuda"]
        row["speedup"] = {"fwd": r["fwd"]["median_ms"] / o["fwd"]["median_ms"],
                          "fwd_bwd": r["fwd_bwd"]["median_ms"] / o["fwd_bwd"]["median_ms"]}
        report["workloads"][name] = row
        print(f"{name}: reference fwd {r['fwd']['median_ms']:.3f} ms, fwd+bwd {r['fwd_bwd']['median_ms']:.3f} ms | "
              f"ours fwd {o['fwd']['median_ms']:.3f} ms, fwd+bwd {o['fwd_bwd']['median_ms']:.3f} ms | "
              f"x{row['speedup']['fwd']:.1f} / x{row['speedup']['fwd_bwd']:.1f} | parity "
              + ", ".join(f"{k} {v:.1e}" for k, v in row["parity_max_err_over_max_ref"].items()), flush=True)
    Path(ns.out).parent.mkdir(parents=True, exist_ok=True)
    Path(ns.out).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
