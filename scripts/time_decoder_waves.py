"""Decoder-shaped operator (BASELINE C4: B=8, Q=900, bf16 storage, pyramid 22 223): forward / backward time against the
number of (b,h) slices per L2 wave (MSDA_B200_SLICES_PER_WAVE; 0 = the library's own choice)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from msda_triton import _lib, kernels as K  # noqa: E402

flush = torch.empty(256 << 18, device="cuda")
B, Q, H, D, L, Kp = 8, 900, 8, 32, 4, 4
for pname, pyr in (("22223", bench.DETR_PYRAMID), ("5440", bench.BENCH_PYRAMID)):
    npix = sum(h * w for h, w in pyr)
    g = torch.Generator().manual_seed(0)
    for dt in (torch.bfloat16, torch.float32):
        v = torch.randn(B, npix, H, D, generator=g).to("cuda", dt)
        pts = torch.rand(B, Q, H, L, Kp, 2, generator=g).to("cuda", dt)
        aw = torch.softmax(torch.randn(B, Q, H, L * Kp, generator=g), -1).reshape(B, Q, H, L, Kp).to("cuda", dt)
        go = torch.rand(B, Q, H, D, generator=g).to("cuda", dt)
        shapes = torch.tensor(pyr, device="cuda")
        row = {}
        for spw in (0, 8, 16, 24, 32, 64):
            os.environ["MSDA_B200_SLICES_PER_WAVE"] = str(spw)
            _lib.reload_tuning()
            res = []
            for fn in (lambda: K.b200_multi_scale_deformable_attention_fwd(v, shapes, pts, aw, "border", True),
                       lambda: K.b200_multi_scale_deformable_attention_bwd(go, v, shapes, pts, aw, "border", True)):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(15):
                    flush.fill_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
                res.append(round(ts[len(ts) // 2] * 1e3, 1))
            row[spw] = res
        print(pname, str(dt).split(".")[-1], "slices/wave -> [fwd us, bwd us (incl. memset + rounding)]", row, flush=True)
os.environ.pop("MSDA_B200_SLICES_PER_WAVE", None)
