"""Reads the clock instrumentation of a -DMSDA_TMEM_PROF build of msda_bwd_tmem.cu (debug builds only):
    MSDA_B200_NVCC_EXTRA=-DMSDA_TMEM_PROF python msda-triton_b200/build.py --force
    python scripts/tmem_prof.py [workload]
"""
import ctypes
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from msda_triton import _lib, kernels as K  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "bench_q10k_border"
B, Q, H, D, pyr, Kp, pm, ac = bench.WORKLOADS[name]
t, s = bench.make_inputs(name, 0, device="cuda")
for levels in ("1", "2"):
    os.environ["MSDA_B200_BWD_TMEM"] = "1"
    os.environ["MSDA_B200_TMEM_LEVELS"] = levels
    _lib.reload_tuning()
    for _ in range(3):
        K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], pm, ac, needs=(1, 1, 1),
                                                    deterministic=False)
    torch.cuda.synchronize()
    buf = np.zeros(148 * 16 * 8, dtype=np.int64)
    lib = _lib.get_lib()
    lib.msda_debug_tmem_prof.argtypes = [ctypes.c_void_p]
    rc = lib.msda_debug_tmem_prof(buf.ctypes.data)
    p = buf.reshape(148, 16, 8)[:, :12].astype(np.float64)
    tiles = B * H * ((Q + 3) // 4) / 148 / 12
    print(f"{name} levels={levels} rc={rc} tiles/warp={tiles:.1f}")
    for i, label in enumerate(("ring wait", "turn work", "flush", "kernel", "prologue", "check+ld+wait", "ffma+wait_st", "sttm issue")):
        print(f"  {label:10s} mean {p[:, :, i].mean():12.0f} clk  per tile {p[:, :, i].mean() / tiles:9.0f}  max {p[:, :, i].max():12.0f}")
