"""Times the nn.Module forward+backward with the fused core on and off for hidden 256 (head_dim 32) and hidden 512
(head_dim 64, the reference README's module example), bf16 and fp32.  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import MultiscaleDeformableAttention  # noqa: E402

PYRAMID = [(64, 64), (32, 32), (16, 16), (8, 8)]


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
    B, Q, emb, H, L, K = 8, 900, 256, 8, 4, 4
    npix = sum(h * w for h, w in PYRAMID)
    shapes = torch.tensor(PYRAMID, device="cuda")
    for dt in (torch.bfloat16, torch.float32):
        for hidden in (256, 512):
            torch.manual_seed(0)
            img = torch.randn(B, npix, emb, device="cuda", dtype=dt, requires_grad=True)
            queries = torch.randn(B, Q, emb, device="cuda", dtype=dt, requires_grad=True)
            ref = torch.rand(B, Q, 2, device="cuda", dtype=dt)
            gout = torch.rand(B, Q, emb, device="cuda", dtype=dt)
            mod = MultiscaleDeformableAttention(emb, hidden, L, H, K, "border", True).to("cuda", dt)

            def step():
                mod(img, shapes, queries, ref).backward(gout)

            row = {}
            for fused in ("1", "0"):
                os.environ["MSDA_B200_FUSED_MODULE"] = fused
                row[fused] = median_ms(step, flush)
            print(f"{dt} hidden={hidden} head_dim={hidden // H}: fused {row['1']:.3f} ms  composed {row['0']:.3f} ms",
                  flush=True)


if __name__ == "__main__":
    main()
