"""Backward / forward time of shapes with more than 16 sampling points per unit (5-level pyramids, K = 8): tuned
kernels versus the generic kernels (MSDA_B200_FORCE_GENERIC=1).  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import _lib, kernels as K  # noqa: E402

SHAPES = {
    "5 levels down to 4x4, K=4": ([(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)], 4),   # 16-pixel level: hot-spot bound
    "dino_5scale 800x1333 K=4": ([(100, 167), (50, 84), (25, 42), (13, 21), (7, 11)], 4),
    "k8 L=4 K=8": ([(64, 64), (32, 32), (16, 16), (8, 8)], 8),
    "L=3 K=8": ([(64, 64), (32, 32), (16, 16)], 8),
    "rtdetr L=3 K=4": ([(64, 64), (32, 32), (16, 16)], 4),
    "L=4 K=7": ([(64, 64), (32, 32), (16, 16), (8, 8)], 7),
    "baseline L=4 K=4": ([(64, 64), (32, 32), (16, 16), (8, 8)], 4),
}


def median_ms(fn, flush, steps=20):
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
    return sorted(ts)[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B, Q, H, D = 4, 10000, 8, 32
    for dt in (torch.float32, torch.bfloat16):
        for name, (pyr, kp) in SHAPES.items():
            L = len(pyr)
            npix = sum(h * w for h, w in pyr)
            g = torch.Generator().manual_seed(0)
            img = torch.randn(B, npix, H, D, generator=g).to("cuda", dt)
            pts = torch.rand(B, Q, H, L, kp, 2, generator=g).to("cuda", dt)
            aw = torch.softmax(torch.randn(B, Q, H, L * kp, generator=g), -1).reshape(B, Q, H, L, kp).to("cuda", dt)
            go = torch.rand(B, Q, H, D, generator=g).to("cuda", dt)
            shapes = torch.tensor(pyr, device="cuda")
            row = {}
            for generic in ("0", "1"):
                os.environ["MSDA_B200_FORCE_GENERIC"] = generic
                _lib.reload_tuning()   # the library reads its knobs once
                row["fwd" + generic] = median_ms(
                    lambda: K.b200_multi_scale_deformable_attention_fwd(img, shapes, pts, aw, "border", True), flush)
                row["bwd" + generic] = median_ms(
                    lambda: K.b200_multi_scale_deformable_attention_bwd(go, img, shapes, pts, aw, "border", True), flush)
            os.environ.pop("MSDA_B200_FORCE_GENERIC")
            _lib.reload_tuning()
            print(f"{dt} {name}: fwd tuned {row['fwd0']:.3f} generic {row['fwd1']:.3f} ms | "
                  f"bwd tuned {row['bwd0']:.3f} generic {row['bwd1']:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
