"""Host-side cost per call of the Python boundary (tiny problem, so the GPU is never the bottleneck): the kernel-layer
wrappers, the autograd op, and a cProfile breakdown.  Run on the GPU box."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
import msda_triton  # noqa: E402
from msda_triton import kernels as K  # noqa: E402


def per_call_us(fn, n=3000):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


def main():
    pyr = [(8, 8), (4, 4), (2, 2), (1, 1)]
    B, Q, H, D, L, Kp = 1, 8, 8, 32, 4, 4
    npix = sum(h * w for h, w in pyr)
    img = torch.randn(B, npix, H, D, device="cuda", requires_grad=True)
    pts = torch.rand(B, Q, H, L, Kp, 2, device="cuda", requires_grad=True)
    aw = torch.rand(B, Q, H, L, Kp, device="cuda", requires_grad=True)
    go = torch.rand(B, Q, H, D, device="cuda")
    shapes = torch.tensor(pyr, device="cuda")
    a, b, c = img.detach(), pts.detach(), aw.detach()

    def fwd():
        K.b200_multi_scale_deformable_attention_fwd(a, shapes, b, c, "border", True)

    def bwd():
        K.b200_multi_scale_deformable_attention_bwd(go, a, shapes, b, c, "border", True)

    def op_nograd():
        with torch.no_grad():
            msda_triton.multiscale_deformable_attention(a, shapes, b, c, "border", True)

    def autograd_step():
        out = msda_triton.multiscale_deformable_attention(img, shapes, pts, aw, "border", True)
        out.backward(go)
        img.grad = pts.grad = aw.grad = None

    def empty3():
        torch.empty_like(a), torch.empty_like(b), torch.empty_like(c)

    for name, fn in (("kernels fwd wrapper", fwd), ("kernels bwd wrapper", bwd), ("functional op (no_grad)", op_nograd),
                     ("autograd fwd+bwd", autograd_step), ("3x torch.empty_like (yardstick)", empty3)):
        print(f"{name}: {per_call_us(fn):.1f} us/call", flush=True)
    for name, fn in (("fwd", fwd), ("bwd", bwd), ("autograd", autograd_step)):
        prof = cProfile.Profile()
        prof.enable()
        for _ in range(2000):
            fn()
        prof.disable()
        print(f"---- cProfile {name} (2000 calls)")
        pstats.Stats(prof).sort_stats("tottime").print_stats(14)


if __name__ == "__main__":
    main()
