"""PCIe ceiling for the e2e path: 125 MB host->device and 125 MB device->host per step (the bench workload's
h2d/d2h bytes), pinned memory, one direction at a time and both directions concurrently on two streams."""
import torch

n = 124682240 // 4
h_in = torch.empty(n).pin_memory()
h_out = torch.empty(n).pin_memory()
d_in = torch.empty(n, device="cuda")
d_out = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 4 / 1e9
for name, fn in (("H2D only", h2d), ("D2H only", d2h), ("H2D + D2H concurrently", both)):
    ms = timed(fn)
    print(f"{name:26s} {ms:7.3f} ms per 124.7 MB (each way)  -> {gb / (ms * 1e-3):6.1f} GB/s per direction")
