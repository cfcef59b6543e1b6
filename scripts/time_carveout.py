"""Forward / backward of the bench shapes under different shared-memory carve-outs (MSDA_B200_CARVEOUT, percent of the
unified L1 / shared array given to shared memory; -1 = driver default).  Cold L2, medians of 25."""
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


def timeit(fn, reps=25, warm=3):
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


for name in sys.argv[1:] or ["bench_q10k_border", "detr_encoder_zeros"]:
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    t, s = bench.make_inputs(name, 0, device="cuda")
    row = {}
    for co in ("-1", "0", "25", "50", "100", "-1"):
        os.environ["MSDA_B200_CARVEOUT"] = co
        _lib.reload_tuning()
        f = timeit(lambda: K.b200_multi_scale_deformable_attention_fwd(t["img"], s, t["pts"], t["aw"], pm, ac))
        b = timeit(lambda: K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], pm, ac))
        row.setdefault("carveout=" + co, []).append([f, b])
    print(name, "fwd / bwd ms:", json.dumps(row), flush=True)
os.environ.pop("MSDA_B200_CARVEOUT", None)
