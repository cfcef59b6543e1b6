"""Backward with / without the pair aggregation of neighbouring queries' row adds (MSDA_B200_BWD_AGG), cold L2, medians."""
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


for name in sys.argv[1:] or ["detr_encoder_init_zeros", "detr_encoder_local_zeros", "detr_encoder_zeros", "bench_q10k_border"]:
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    t, s = bench.make_inputs(name, 0, device="cuda")
    row = {}
    for agg in ("0", "1"):
        os.environ["MSDA_B200_BWD_AGG"] = agg
        _lib.reload_tuning()
        row["agg=" + agg] = timeit(lambda: K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], pm, ac))
    print(name, "backward ms:", json.dumps(row), flush=True)
os.environ.pop("MSDA_B200_BWD_AGG", None)
