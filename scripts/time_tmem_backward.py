"""A/B timing of the tensor-memory backward (csrc/msda_bwd_tmem.cu) against the plain tuned backward: cold L2, CUDA
events, medians.  Knobs: MSDA_B200_BWD_TMEM (0/1), MSDA_B200_TMEM_LEVELS (0/1/2).

    python scripts/time_tmem_backward.py [workload ...] [--json gpurun_out/tmem_timing.json]
"""
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


def timeit(fn, reps=30, warm=4):
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
    return ts[len(ts) // 2]


VARIANTS = [
    ("plain", {"MSDA_B200_BWD_TMEM": "0"}),
    ("tmem_levels0", {"MSDA_B200_BWD_TMEM": "1", "MSDA_B200_TMEM_LEVELS": "0"}),
    ("tmem_levels1", {"MSDA_B200_BWD_TMEM": "1", "MSDA_B200_TMEM_LEVELS": "1"}),
    ("tmem_levels2", {"MSDA_B200_BWD_TMEM": "1", "MSDA_B200_TMEM_LEVELS": "2"}),
    ("tmem_levels0_w12", {"MSDA_B200_BWD_TMEM": "1", "MSDA_B200_TMEM_LEVELS": "0", "MSDA_B200_TMEM_WARPS": "12"}),
    ("tmem_levels2_w12", {"MSDA_B200_BWD_TMEM": "1", "MSDA_B200_TMEM_LEVELS": "2", "MSDA_B200_TMEM_WARPS": "12"}),
    ("auto", {}),
]
KNOBS = ("MSDA_B200_BWD_TMEM", "MSDA_B200_TMEM_LEVELS", "MSDA_B200_TMEM_WARPS")

out_json = None
names = []
argv = sys.argv[1:]
i = 0
while i < len(argv):
    if argv[i] == "--json":
        out_json = argv[i + 1]
        i += 2
    else:
        names.append(argv[i])
        i += 1
names = names or ["bench_q10k_border", "bench_q10k_zeros", "detr_encoder_zeros", "detr_encoder_local_zeros",
                  "readme_q900_zeros"]
results = {}
for name in names:
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    t, s = bench.make_inputs(name, 0, device="cuda")
    row = {}
    for label, env in VARIANTS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        _lib.reload_tuning()
        row[label] = round(timeit(lambda: K.b200_multi_scale_deformable_attention_bwd(
            t["go"], t["img"], s, t["pts"], t["aw"], pm, ac, needs=(1, 1, 1), deterministic=False)), 4)
    for k in KNOBS:
        os.environ.pop(k, None)
    _lib.reload_tuning()
    row["fwd"] = round(timeit(lambda: K.b200_multi_scale_deformable_attention_fwd(t["img"], s, t["pts"], t["aw"], pm, ac)), 4)
    results[name] = row
    print(name, json.dumps(row), flush=True)
if out_json:
    Path(out_json).parent.mkdir(parents=True, exist_ok=True)
    Path(out_json).write_text(json.dumps(results, indent=1))
