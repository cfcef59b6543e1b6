"""Backward with n owner warps accumulating the coarsest level in registers (MSDA_B200_BWD_DENSE=n, MSDA_B200_DENSE_PF=p)
against the plain tuned backward: cold L2, medians, plus the largest deviation of grad_img from the plain result."""
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


def timeit(fn, reps=25, warm=4):
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


variants = [("0", "2", "0"), ("0", "2", "1")] + [(n, pf, "1") for n in ("2", "4") for pf in ("2", "3")]
for name in sys.argv[1:] or ["bench_q10k_border", "bench_q10k_zeros", "readme_q900_zeros", "detr_encoder_zeros"]:
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    t, s = bench.make_inputs(name, 0, device="cuda")
    row, base = {}, None
    for nown, pf, shape in variants:
        os.environ["MSDA_B200_BWD_SHAPE"] = shape
        os.environ["MSDA_B200_BWD_DENSE"] = nown
        os.environ["MSDA_B200_DENSE_PF"] = pf
        _lib.reload_tuning()
        run = lambda: K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], pm, ac)  # noqa: E731
        gi = run()[0]
        torch.cuda.synchronize()
        if base is None:
            base = gi.clone()
        err = float((gi - base).abs().max() / base.abs().max())
        row[f"owners={nown},pf={pf},shape={shape}"] = [timeit(run), f"{err:.1e}"]
    print(name, "backward ms, max|d grad_img|/max:", json.dumps(row), flush=True)
os.environ.pop("MSDA_B200_BWD_DENSE", None)
os.environ.pop("MSDA_B200_DENSE_PF", None)
os.environ.pop("MSDA_B200_BWD_SHAPE", None)
