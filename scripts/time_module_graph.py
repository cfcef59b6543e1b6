"""Grounding-DINO-decoder-sized module step (B=8, Q=900, emb=hidden=256, bf16): eager versus one CUDA graph of the
whole forward+backward.  The step launches ~40 small kernels, so eager time is host-launch bound; the library's entry
points never allocate or synchronise, hence the whole step is capturable.  Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import MultiscaleDeformableAttention  # noqa: E402

PYRAMID = [(64, 64), (32, 32), (16, 16), (8, 8)]


def median_ms(fn, flush, steps=30):
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
        torch.manual_seed(0)
        img = torch.randn(B, npix, emb, device="cuda", dtype=dt, requires_grad=True)
        queries = torch.randn(B, Q, emb, device="cuda", dtype=dt, requires_grad=True)
        ref = torch.rand(B, Q, 2, device="cuda", dtype=dt)
        gout = torch.rand(B, Q, emb, device="cuda", dtype=dt)
        mod = MultiscaleDeformableAttention(emb, emb, L, H, K, "border", True).to("cuda", dt)
        params = [img, queries] + list(mod.parameters())

        def step():
            for p in params:
                p.grad = None
            mod(img, shapes, queries, ref).backward(gout)

        eager = median_ms(step, flush)
        step()
        want = [p.grad.clone() for p in params]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        for p in params:
            p.grad = None
        with torch.cuda.graph(graph):
            mod(img, shapes, queries, ref).backward(gout)
        graph.replay()
        torch.cuda.synchronize()
        same = all(torch.allclose(p.grad.float(), w.float(), rtol=2e-2, atol=2e-2 * float(w.float().abs().max())) for p, w in zip(params, want))
        graphed = median_ms(graph.replay, flush)
        print(f"{dt}: eager {eager:.3f} ms, CUDA graph {graphed:.3f} ms (graph grads match eager grads: {same})",
              flush=True)


if __name__ == "__main__":
    main()
