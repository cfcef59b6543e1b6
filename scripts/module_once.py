"""One fwd+bwd step of the Grounding-DINO decoder module config (BASELINE C4: B=8, Q=900, emb 256, bf16) -- target of
`ncu --metrics gpu__time_duration.sum` launch lists.   python scripts/module_once.py [5440|22223]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from msda_triton import MultiscaleDeformableAttention  # noqa: E402

pyr = bench.DETR_PYRAMID if (len(sys.argv) < 2 or sys.argv[1] == "22223") else bench.BENCH_PYRAMID
B, Q, emb, H, L, Kp = 8, 900, 256, 8, 4, 4
npix = sum(h * w for h, w in pyr)
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16
img = torch.randn(B, npix, emb, generator=g).to("cuda", dt).requires_grad_(True)
queries = torch.randn(B, Q, emb, generator=g).to("cuda", dt).requires_grad_(True)
ref = torch.rand(B, Q, 2, generator=g).to("cuda", dt)
shapes = torch.tensor(pyr, device="cuda")
mod = MultiscaleDeformableAttention(emb, emb, L, H, Kp, "border", True).to("cuda", dt)
gout = torch.rand(B, Q, emb, generator=g).to("cuda", dt)
for _ in range(4):
    out = mod(img, shapes, queries, ref)
    out.backward(gout)
torch.cuda.synchronize()
