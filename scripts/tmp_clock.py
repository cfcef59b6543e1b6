import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import kernels as K
PYR = [(100, 167), (50, 84), (25, 42), (13, 21)]
H, D, L, kp = 8, 32, 4, 4
npix = sum(h * w for h, w in PYR); Q = npix
shapes = torch.tensor(PYR, device="cuda")
samples = []
stop = False
def sampler():
    while not stop:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append(o)
        time.sleep(0.05)
for B in (16, 64):
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.randn(B, npix, H, D, device="cuda", generator=g)
    pts = torch.rand(B, Q, H, L, kp, 2, device="cuda", generator=g)
    aw = torch.rand(B, Q, H, L, kp, device="cuda", generator=g)
    go = torch.rand(B, Q, H, D, device="cuda", generator=g)
    out = torch.empty(B, Q, H, D, device="cuda")
    grads = (torch.empty_like(img), torch.empty_like(pts), torch.empty_like(aw))
    for what in ("fwd", "bwd"):
        fn = (lambda: K.b200_multi_scale_deformable_attention_fwd(img, shapes, pts, aw, "zeros", False, out=out)) if what == "fwd" else \
             (lambda: K.b200_multi_scale_deformable_attention_bwd(go, img, shapes, pts, aw, "zeros", False, grads=grads))
        fn(); torch.cuda.synchronize()
        samples.clear(); stop = False
        th = threading.Thread(target=sampler); th.start()
        t0 = time.time(); n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 2.5:
            for _ in range(5): fn()
            n += 5
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        stop = True; th.join()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B} {what}: {ms / B * 1e3:.1f} us/image back-to-back; samples (MHz, W, reasons): {samples[2:12]}", flush=True)
    del img, pts, aw, go, out, grads
