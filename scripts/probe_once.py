"""Runs the two L2 probes of libmsda_b200.so a few times (target of `ncu --set full -k regex:probe`, see profiles/)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from msda_triton import _lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
print(bench.l2_probe(_lib.get_lib(), kinds=("gather", "scatter") if which == "both" else (which,)))
torch.cuda.synchronize()
