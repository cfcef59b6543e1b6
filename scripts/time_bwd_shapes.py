"""Backward launch shape (MSDA_B200_BWD_SHAPE): 16 warps x 128 registers (0) against 12 warps x 168 registers (1) and the
variants with the tap exchange issued one batch ahead (3, 4); cold L2, medians of 25, alternating order.
MSDA_TIME_DTYPE=bf16|f16 times the 16-bit-storage backward (incl. its fp32 scratch zero-fill and rounding pass)."""
import json
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
bench.WORKLOADS["decoder_q900_detr_pyramid"] = (8, 900, 8, 32, bench.DETR_PYRAMID, 4, "zeros", False)
bench.WORKLOADS["encoder_b16_zeros"] = (16, 22223, 8, 32, bench.DETR_PYRAMID, 4, "zeros", False)


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2], 4)


DTYPE = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[os.environ.get("MSDA_TIME_DTYPE", "f32")]
shapes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1"]
names = sys.argv[2:] or ["bench_q10k_border", "bench_q10k_zeros", "detr_encoder_zeros", "detr_encoder_local_zeros",
                         "readme_q900_zeros", "decoder_q900_detr_pyramid", "encoder_b16_zeros", "train_b64_encoder_zeros"]
for name in names:
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    gen = name if name in ("detr_encoder_local_zeros", "detr_encoder_init_zeros") else None
    if gen:
        t, s = bench.make_inputs(name, 0, device="cuda")
    else:
        g = torch.Generator(device="cuda").manual_seed(1)
        npix = sum(h * w for h, w in pyr)
        t = {"img": torch.randn(B, npix, H, D, device="cuda", generator=g),
             "pts": torch.rand(B, Q, H, len(pyr), Kp, 2, device="cuda", generator=g),
             "aw": torch.rand(B, Q, H, len(pyr), Kp, device="cuda", generator=g),
             "go": torch.rand(B, Q, H, D, device="cuda", generator=g)}
        s = torch.tensor(pyr, device="cuda")
    t = {k: v.to(DTYPE) for k, v in t.items()}
    reps = 7 if B >= 16 else 25
    row = {}
    for rnd in range(2):                      # two rounds, alternating, to see the run-to-run spread
        for sh in shapes:
            os.environ["MSDA_B200_BWD_SHAPE"] = sh
            _lib.reload_tuning()
            run = lambda: K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], pm, ac)  # noqa: E731
            row.setdefault("shape=" + sh, []).append(timeit(run, reps))
    print(name, "backward ms:", json.dumps(row), flush=True)
    del t
    torch.cuda.empty_cache()
os.environ.pop("MSDA_B200_BWD_SHAPE", None)
