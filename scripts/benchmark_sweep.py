"""Re-implementation of the reference's benchmark sweeps (``/root/reference/scripts/benchmark.py``) without Triton or
matplotlib: same shapes (B=4, H=8, C=32, P=4, pyramid 64^2..8^2, fp32, border / align_corners=True, ``:25-31``), same
query counts (``:13``), same three quantities -- forward ms (``:23-55``), forward+backward ms through autograd with a
fresh ``rand_like`` gradient per run (``:71-107``) and peak extra memory of forward+backward (``:123-174``) -- and the
same timing discipline as ``triton.testing.do_bench``: L2 flushed before every repetition, median of the reps.

    python scripts/benchmark_sweep.py [--providers cuda torch reference_triton] [--csv profiles/r2c_benchmark_sweep.csv]
    python scripts/plot_sweep_svg.py profiles/r2c_benchmark_sweep.csv      # the three figures of benchmark.py:178-180 as SVG

providers: ``cuda`` = this repository's kernels (public API), ``torch`` = the torch grid_sample route on the GPU (the
reference's "Torch" line).  Prints a markdown table and optionally writes a CSV.
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "msda-triton_b200"))

import torch  # noqa: E402

from msda_triton.frontend import (  # noqa: E402
    native_multiscale_deformable_attention,
    triton_multiscale_deformable_attention,
)

QUERIES = [10, 100, 300, 900, 1000, 10000]
SHAPES = [(64, 64), (32, 32), (16, 16), (8, 8)]


def make(num_queries, requires_grad):
    B, H, C, P = 4, 8, 32, 4
    L = len(SHAPES)
    npix = sum(h * w for h, w in SHAPES)
    g = torch.Generator(device="cuda").manual_seed(num_queries)
    img = torch.randn(B, npix, H, C, device="cuda", generator=g).requires_grad_(requires_grad)
    shapes = torch.tensor(SHAPES, device="cuda")
    pts = torch.rand(B, num_queries, H, L, P, 2, device="cuda", generator=g).requires_grad_(requires_grad)
    aw = torch.softmax(torch.randn(B, num_queries, H, L, P, device="cuda", generator=g), dim=-1)
    aw.requires_grad_(requires_grad)
    return img, shapes, pts, aw


def do_bench(fn, flush, warmup=10, reps=50):
    for _ in range(warmup):
        fn()
    times = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--providers", nargs="+", default=["cuda", "torch"])
    ap.add_argument("--csv", default=None)
    ns = ap.parse_args()
    ops = {"cuda": triton_multiscale_deformable_attention, "torch": native_multiscale_deformable_attention}
    if "reference_triton" in ns.providers:
        # the UNMODIFIED reference package staged into baseline/_ref (scripts/stage_reference.py): its own Triton kernels,
        # JIT-compiled for sm_100a -- the "Triton" line of the reference's published plots, on this GPU
        sys.path.insert(0, str(ROOT))
        import bench
        ref, why = bench.load_reference_package()
        if ref is None:
            print(f"reference_triton unavailable: {why}", file=sys.stderr)
            ns.providers = [p for p in ns.providers if p != "reference_triton"]
        else:
            ops["reference_triton"] = ref.triton_multiscale_deformable_attention
    flush = torch.empty(256 << 20, dtype=torch.int8, device="cuda")
    rows = []
    for n in QUERIES:
        for prov in ns.providers:
            op = ops[prov]
            img, shapes, pts, aw = make(n, False)

            def fwd():
                with torch.no_grad():
                    op(img, shapes, pts, aw, "border", True)

            gimg, gshapes, gpts, gaw = make(n, True)

            def fwd_bwd():
                out = op(gimg, gshapes, gpts, gaw, "border", True)
                out.backward(torch.rand_like(out))
                gimg.grad = gpts.grad = gaw.grad = None

            t_f = do_bench(fwd, flush)
            t_fb = do_bench(fwd_bwd, flush)
            mem = 0.0
            for _ in range(10):
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats()
                start = torch.cuda.memory_allocated()
                fwd_bwd()
                torch.cuda.synchronize()
                mem += (torch.cuda.max_memory_allocated() - start) / 1e6
            rows.append((n, prov, t_f, t_fb, mem / 10))
    print("| num_queries | provider | fwd ms | fwd+bwd ms | peak extra memory MB |\n|---:|---|---:|---:|---:|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]:.4f} | {r[3]:.4f} | {r[4]:.2f} |")
    if ns.csv:
        with open(ns.csv, "w") as f:
            f.write("num_queries,provider,fwd_ms,fwd_bwd_ms,peak_extra_memory_mb\n")
            for r in rows:
                f.write(f"{r[0]},{r[1]},{r[2]:.5f},{r[3]:.5f},{r[4]:.3f}\n")


if __name__ == "__main__":
    main()
