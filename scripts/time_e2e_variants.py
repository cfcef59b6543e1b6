"""e2e (pinned host buffers in and out) on the bench shape: HostMsda with 1 / 2 / 4 chunks per call, next to the plain
duplex copy of the same bytes (the ceiling)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

for chunks in (1, 2, 4):
    ms, h2d, d2h = bench.time_e2e(bench.HEADLINE, 20, 3, pipelined=True, chunks=chunks)
    print(f"HostMsda chunks={chunks}: {ms:.3f} ms per step")
print("copy ceiling:", round(bench.pcie_probe(h2d, d2h, None, torch.cuda.synchronize), 3), "ms")
