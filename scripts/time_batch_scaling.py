"""Per-image time of the DETR-encoder shape as the batch grows (B = 2 ... 64): does anything besides the work scale?
Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import kernels as K  # noqa: E402

PYR = [(100, 167), (50, 84), (25, 42), (13, 21)]


def median_ms(fn, flush, steps=7):
    for _ in range(2):
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
    H, D, L, kp = 8, 32, 4, 4
    npix = sum(h * w for h, w in PYR)
    Q = npix
    shapes = torch.tensor(PYR, device="cuda")
    for B in (2, 4, 8, 16, 32, 64):
        g = torch.Generator(device="cuda").manual_seed(0)
        img = torch.randn(B, npix, H, D, device="cuda", generator=g)
        pts = torch.rand(B, Q, H, L, kp, 2, device="cuda", generator=g)
        aw = torch.rand(B, Q, H, L, kp, device="cuda", generator=g)
        go = torch.rand(B, Q, H, D, device="cuda", generator=g)
        out = torch.empty(B, Q, H, D, device="cuda")
        grads = (torch.empty_like(img), torch.empty_like(pts), torch.empty_like(aw))
        fwd = median_ms(lambda: K.b200_multi_scale_deformable_attention_fwd(img, shapes, pts, aw, "zeros", False, out=out),
                        flush)
        bwd = median_ms(lambda: K.b200_multi_scale_deformable_attention_bwd(go, img, shapes, pts, aw, "zeros", False,
                                                                            grads=grads), flush)
        noimg = median_ms(lambda: K.b200_multi_scale_deformable_attention_bwd(
            go, img, shapes, pts, aw, "zeros", False, needs=(False, True, True), grads=grads), flush)
        print(f"B={B}: per image fwd {fwd / B * 1e3:.1f} us, bwd {bwd / B * 1e3:.1f} us, bwd without grad_img "
              f"{noimg / B * 1e3:.1f} us   (totals {fwd:.3f} / {bwd:.3f} / {noimg:.3f} ms)", flush=True)
        del img, pts, aw, go, out, grads


if __name__ == "__main__":
    main()
