"""The library's peer-memory all-gather / reduce-scatter (csrc/msda_peer.cu) against NCCL's, DETR-encoder message size
(B=2 x 22 223 pixels x 8 heads x 32 channels fp32 = 45.5 MB), back to back without an L2 flush.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/time_peer_collectives.py"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from msda_triton import distributed as D  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, npix, H, Dh = 2, 22223, 8, 32
ex = D.PeerPixelExchange(B, npix, H, Dh)
shard = torch.randn(ex.shape_shard, device="cuda")
full = torch.empty(ex.shape_full, device="cuda")
out = torch.empty(ex.shape_shard, device="cuda")


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def nccl_ag():
    for b in range(B):
        dist.all_gather_into_tensor(full[b], shard[b])


def nccl_rs():
    for b in range(B):
        dist.reduce_scatter_tensor(out[b], full[b])


res = {"peer_all_gather_ms": timed(lambda: ex.all_gather(shard)), "peer_reduce_scatter_ms": timed(lambda: ex.reduce_scatter(out)),
       "nccl_all_gather_ms": timed(nccl_ag), "nccl_reduce_scatter_ms": timed(nccl_rs)}
remote = (world - 1) / world * full.numel() * 4
res["peer_all_gather_GBps_in"] = remote / (res["peer_all_gather_ms"] * 1e-3) / 1e9
res["peer_reduce_scatter_GBps_in"] = remote / (res["peer_reduce_scatter_ms"] * 1e-3) / 1e9
# correctness of the pulls against NCCL
ex.all_gather(shard)
nccl_ag()
torch.cuda.synchronize()
res["all_gather_equal"] = bool(torch.equal(ex.full, full))
if rank == 0:
    print(world, "ranks:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
dist.barrier()
dist.destroy_process_group()
