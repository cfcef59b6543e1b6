"""Forward / backward time against MSDA_B200_L1_KEEP_KB (pyramid KB per (b,h) slice the gathers keep in L1; levels finer
than that are gathered with no-allocate loads; -1 = never stream)."""
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


names = sys.argv[1:] or ["bench_q10k_border", "bench_q10k_zeros", "detr_encoder_zeros", "detr_encoder_local_zeros"]
for name in names:
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    t, s = bench.make_inputs(name, 0, device="cuda")
    row = {}
    for kb in (-1, 40, 176, 100000):
        os.environ["MSDA_B200_L1_KEEP_KB"] = str(kb)
        _lib.reload_tuning()
        row[kb] = [timeit(lambda: K.b200_multi_scale_deformable_attention_fwd(t["img"], s, t["pts"], t["aw"], pm, ac)),
                   timeit(lambda: K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], pm, ac))]
    print(name, "keep KB -> [fwd ms, bwd ms]", json.dumps(row), flush=True)
os.environ.pop("MSDA_B200_L1_KEEP_KB", None)
