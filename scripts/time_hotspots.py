"""Backward time of the bench problem (B=4, Q=10k, H=8, D=32, L=4, K=4, fp32) on pyramids with the same number of
row adds but different hot-spot structure: is the backward bound by L2 atomic THROUGHPUT or by same-address
serialisation on the small coarse levels?  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import kernels as K  # noqa: E402

PYRAMIDS = {
    "bench 64,32,16,8": [(64, 64), (32, 32), (16, 16), (8, 8)],
    "flat 4 x 64x64": [(64, 64)] * 4,
    "flat 4 x 32x32": [(32, 32)] * 4,
    "flat 4 x 16x16": [(16, 16)] * 4,
    "flat 4 x 8x8": [(8, 8)] * 4,
    "flat 4 x 4x4": [(4, 4)] * 4,
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
    B, Q, H, D, kp = 4, 10000, 8, 32, 4
    for name, pyr in PYRAMIDS.items():
        L = len(pyr)
        npix = sum(h * w for h, w in pyr)
        g = torch.Generator().manual_seed(0)
        img = torch.randn(B, npix, H, D, generator=g).cuda()
        pts = torch.rand(B, Q, H, L, kp, 2, generator=g).cuda()
        aw = torch.softmax(torch.randn(B, Q, H, L * kp, generator=g), -1).reshape(B, Q, H, L, kp).cuda()
        go = torch.rand(B, Q, H, D, generator=g).cuda()
        shapes = torch.tensor(pyr, device="cuda")
        fwd = median_ms(lambda: K.b200_multi_scale_deformable_attention_fwd(img, shapes, pts, aw, "border", True), flush)
        full = median_ms(lambda: K.b200_multi_scale_deformable_attention_bwd(go, img, shapes, pts, aw, "border", True),
                         flush)
        no_img = median_ms(lambda: K.b200_multi_scale_deformable_attention_bwd(
            go, img, shapes, pts, aw, "border", True, needs=(False, True, True)), flush)
        print(f"{name}: rows/(b,h) {npix}  fwd {fwd:.3f}  bwd {full:.3f}  bwd without grad_img {no_img:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
