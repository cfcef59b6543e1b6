import sys; sys.path.insert(0, "msda-triton_b200"); sys.path.insert(0, ".")
import torch, bench
from msda_triton import kernels as K
t, s = bench.make_inputs("bench_q10k_border", 0, device="cuda")
for _ in range(3):
    K.b200_multi_scale_deformable_attention_bwd(t["go"], t["img"], s, t["pts"], t["aw"], "border", True, deterministic=True)
torch.cuda.synchronize()
