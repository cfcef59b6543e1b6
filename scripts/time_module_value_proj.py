"""BASELINE config C4 (Grounding-DINO decoder module, bf16, B=8, Q=900) with and without the fused value-projection node
(MSDA_B200_FUSED_VALUE_PROJ): module fwd+bwd, cold L2, medians of 20."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

flush = torch.empty(256 << 18, device="cuda")
res = {}
for label, env in (("separate_nodes", "0"), ("fused_value_proj", "1"), ("separate_nodes_again", "0")):
    os.environ["MSDA_B200_FUSED_VALUE_PROJ"] = env
    r = bench.time_module(flush)
    res[label] = {k: round(v["module_fwd_bwd_ms"], 4) for k, v in r.items()}
os.environ.pop("MSDA_B200_FUSED_VALUE_PROJ", None)
print(json.dumps(res))
