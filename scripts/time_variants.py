"""Quick device-time probes of the kernels under different flag combinations (cold L2, CUDA events)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from msda_triton import kernels as K  # noqa: E402

flush = torch.empty(256 << 18, device="cuda")


def timeit(fn, reps=20, warm=3):
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


for name in [a for a in sys.argv[1:] if not a.startswith('--')] or list(bench.WORKLOADS):
    B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
    t, s = bench.make_inputs(name, 0, device="cuda")
    for dt in ("bf16", "f16"):
        if "--" + dt in sys.argv:
            t = {k: v.to(torch.bfloat16 if dt == "bf16" else torch.float16) for k, v in t.items()}
    fwd = timeit(lambda: K.b200_multi_scale_deformable_attention_fwd(t["img"], s, t["pts"], t["aw"], pm, ac))
    res = {"fwd": fwd}
    for label, needs in (("bwd_all", (1, 1, 1)), ("bwd_img_only", (1, 0, 0)), ("bwd_no_img", (0, 1, 1))):
        res[label] = timeit(lambda: K.b200_multi_scale_deformable_attention_bwd(
            t["go"], t["img"], s, t["pts"], t["aw"], pm, ac, needs=needs, deterministic=False))
    if "--nodet" not in sys.argv:
      res["bwd_deterministic"] = timeit(lambda: K.b200_multi_scale_deformable_attention_bwd(
        t["go"], t["img"], s, t["pts"], t["aw"], pm, ac, deterministic=True), reps=5)
    print(name, {k: round(v, 4) for k, v in res.items()})
