"""Runs a few forward+backward steps of one bench workload for ncu (no timing, no flush).

    ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 2 -o gpurun_out/prof \
        python scripts/profile_step.py --workload bench_q10k_border --steps 3
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
from msda_triton import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default=bench.HEADLINE)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--points", default="unit", choices=["unit", "local"])
ap.add_argument("--batch", type=int, default=0, help="override the workload's batch size")
ns = ap.parse_args()

if ns.batch:
    bench.WORKLOADS[ns.workload] = (ns.batch,) + tuple(bench.WORKLOADS[ns.workload][1:])
B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[ns.workload]
t, shapes = bench.make_inputs(ns.workload, seed=0, device="cuda")
for _ in range(ns.steps):
    out = K.b200_multi_scale_deformable_attention_fwd(t["img"], shapes, t["pts"], t["aw"], pm, ac)
    g = K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], shapes, t["pts"], t["aw"], pm, ac)
torch.cuda.synchronize()
print("done", float(out.sum()), float(g[0].sum()))
