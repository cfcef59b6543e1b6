"""Query-sharded MSDA across the ranks of one node: every rank holds a pixel shard of `img` and a query
shard of the DETR-encoder workload; forward all-gathers the pixel shards, backward reduce-scatters grad_img over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_query_sharded.py [--steps 30]

Prints one JSON line from rank 0: device time of fwd+bwd (max over ranks), time of the two collectives alone, and the
equivalence error against the unsharded op computed on rank 0.
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "msda-triton_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import msda_triton  # noqa: E402
from msda_triton import distributed as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--workload", default="detr_encoder_zeros")
    ns = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, Q, H, Dh, pyr, K, pm, ac = bench.WORKLOADS[ns.workload]
    t, shapes = bench.make_inputs(ns.workload, seed=0, device="cuda")
    npix = t["img"].shape[1]
    shard = D.shard_pixels(t["img"], rank, world).clone().requires_grad_(True)
    pts = D.shard_queries(t["pts"], rank, world).contiguous().requires_grad_(True)
    aw = D.shard_queries(t["aw"], rank, world).contiguous().requires_grad_(True)
    go = D.shard_queries(t["go"], rank, world).contiguous()

    def step():
        out = D.query_sharded_msda(shard, npix, shapes, pts, aw, pm, ac)
        out.backward(go)
        g = (shard.grad, pts.grad, aw.grad)
        shard.grad = pts.grad = aw.grad = None
        return out, g

    for _ in range(5):
        out, grads = step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ns.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / ns.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    # the collectives alone (same message sizes)
    full = torch.empty((world,) + tuple(shard.shape), device="cuda")
    e0.record()
    for _ in range(ns.steps):
        dist.all_gather_into_tensor(full.view(world * shard.shape[0], *shard.shape[1:]), shard.detach())
        dist.reduce_scatter_tensor(shard.detach().clone(), full.view(world * shard.shape[0], *shard.shape[1:]))
    e1.record()
    torch.cuda.synchronize()
    coll = torch.tensor([e0.elapsed_time(e1) / ns.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(coll, op=dist.ReduceOp.MAX)

    # equivalence against the unsharded op
    a, b, c = (t[k].clone().requires_grad_(True) for k in ("img", "pts", "aw"))
    ref = msda_triton.multiscale_deformable_attention(a, shapes, b, c, pm, ac)
    ref.backward(t["go"])
    qlo, qhi = D.shard_range(Q, rank, world)
    err_out = float((out.detach() - ref.detach()[:, qlo:qhi]).abs().max())
    want = D.shard_pixels(a.grad, rank, world)
    err_gimg = float((grads[0] - want).abs().max() / want.abs().max())
    errs = torch.tensor([err_out, err_gimg], device="cuda", dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        msg = B * npix * H * Dh * 4
        print(json.dumps({
            "workload": ns.workload, "n_gpus": world, "partition": "query-sharded, pixel-sharded value",
            "fwd_bwd_ms": ms.item(), "collectives_alone_ms": coll.item(),
            "all_gather_and_reduce_scatter_message_bytes": msg,
            "max_abs_err_out": errs[0].item(), "max_rel_err_grad_img": errs[1].item()}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
