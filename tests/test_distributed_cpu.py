"""world_size-2 gloo test (CPU) of the sharding logic: batch-sharded and query-sharded MSDA equal the unsharded op,
including the reduce-scatter of grad_img.  On the B200 box the same code runs over NCCL (tests/test_distributed_gpu.py)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, backend, device_kind, results):
    for p in (ROOT, ROOT / "msda-triton_b200", ROOT / "tests"):
        sys.path.insert(0, str(p))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank) if device_kind == "cuda" else torch.device("cpu")
    if device_kind == "cuda":
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        import msda_triton
        from msda_triton import distributed as D
        from util import make_inputs

        B, Q, H, Dh, K = 4, 37, 2, 8, 2
        shapes = [(7, 5), (4, 3), (2, 2)]
        img, s, pts, aw, go = make_inputs(B, Q, H, Dh, shapes, K, seed=21, points="wide", weights="softmax_lk")
        img, s, pts, aw, go = (t.to(dev) for t in (img, s, pts, aw, go))
        npix = img.shape[1]
        pm, ac = "zeros", False

        # unsharded truth (same on every rank)
        a, b, c = (t.clone().requires_grad_(True) for t in (img, pts, aw))
        ref = msda_triton.multiscale_deformable_attention(a, s, b, c, pm, ac)
        ref.backward(go)

        # --- batch sharding: no collective at all ---
        ab, bb, cb = (D.shard_batch(t, rank, world).clone().requires_grad_(True) for t in (img, pts, aw))
        out_b = msda_triton.multiscale_deformable_attention(ab, s, bb, cb, pm, ac)
        out_b.backward(D.shard_batch(go, rank, world))
        lo, hi = D.shard_range(B, rank, world)
        tol = dict(rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out_b, ref[lo:hi].detach(), **tol)
        torch.testing.assert_close(ab.grad, a.grad[lo:hi], **tol)
        torch.testing.assert_close(bb.grad, b.grad[lo:hi], **tol)
        torch.testing.assert_close(cb.grad, c.grad[lo:hi], **tol)

        # --- query sharding: all-gather of pixel shards forward, reduce-scatter of grad_img backward ---
        shard = D.shard_pixels(img, rank, world).clone().requires_grad_(True)
        pq, wq = (D.shard_queries(t, rank, world).clone().requires_grad_(True) for t in (pts, aw))
        out_q = D.query_sharded_msda(shard, npix, s, pq, wq, pm, ac)
        out_q.backward(D.shard_queries(go, rank, world).contiguous())
        qlo, qhi = D.shard_range(Q, rank, world)
        torch.testing.assert_close(out_q, ref[:, qlo:qhi].detach(), **tol)
        torch.testing.assert_close(pq.grad, b.grad[:, qlo:qhi], **tol)
        torch.testing.assert_close(wq.grad, c.grad[:, qlo:qhi], **tol)
        chunk = D.pixel_chunk(npix, world)
        want = D.shard_pixels(a.grad, rank, world)
        assert shard.grad.shape == (B, chunk, H, Dh)
        torch.testing.assert_close(shard.grad, want, rtol=1e-5, atol=1e-5)
        if device_kind == "cuda":
            # --- the same over NVLink peer memory (library kernels instead of NCCL), several steps in a row ---
            ex = D.PeerPixelExchange(B, npix, H, Dh)
            for step in range(3):
                sh2 = D.shard_pixels(img, rank, world).clone().requires_grad_(True)
                pq2, wq2 = (D.shard_queries(t, rank, world).clone().requires_grad_(True) for t in (pts, aw))
                out_p = D.peer_query_sharded_msda(ex, sh2, s, pq2, wq2, pm, ac)
                out_p.backward(D.shard_queries(go, rank, world).contiguous())
                assert torch.equal(out_p, out_q), f"peer forward differs (step {step})"
                assert torch.equal(pq2.grad, pq.grad) and torch.equal(wq2.grad, wq.grad)
                torch.testing.assert_close(sh2.grad, want, rtol=1e-5, atol=1e-5)
            # forward only, twice (no reduce-scatter in between): the staging shard is re-used safely
            with torch.no_grad():
                for _ in range(2):
                    o = D.peer_query_sharded_msda(ex, D.shard_pixels(img, rank, world).contiguous(), s, pq2.detach(),
                                                  wq2.detach(), pm, ac)
                assert torch.equal(o, out_q)
            torch.cuda.synchronize()
            # the whole peer step as one CUDA graph (device-side call counts): replays reproduce the eager results
            sh3 = D.shard_pixels(img, rank, world).clone().requires_grad_(True)
            pq3, wq3 = (D.shard_queries(t, rank, world).clone().requires_grad_(True) for t in (pts, aw))
            go3 = D.shard_queries(go, rank, world).contiguous()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    D.peer_query_sharded_msda(ex, sh3, s, pq3, wq3, pm, ac).backward(go3)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            dist.barrier()
            sh3.grad = pq3.grad = wq3.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out_g = D.peer_query_sharded_msda(ex, sh3, s, pq3, wq3, pm, ac)
                out_g.backward(go3)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out_g, out_q), "graph replay: forward differs"
            assert torch.equal(pq3.grad, pq.grad) and torch.equal(wq3.grad, wq.grad)
            torch.testing.assert_close(sh3.grad, want, rtol=1e-5, atol=1e-5)
            del graph
        results[rank] = "ok"
    except Exception as ex:  # noqa: BLE001
        import traceback
        results[rank] = "FAILED: " + "".join(traceback.format_exception(ex))
    finally:
        dist.destroy_process_group()


def run_world(world, backend, device_kind):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        results = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, backend, device_kind, results)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        for p in procs:
            if p.is_alive():
                p.terminate()
        got = dict(results)
    assert len(got) == world, f"ranks reported: {sorted(got)}"
    for r in range(world):
        assert got[r] == "ok", f"rank {r}: {got[r]}"


def test_sharded_equals_unsharded_gloo_world2():
    run_world(2, "gloo", "cpu")


def test_shard_range_covers_everything():
    sys.path.insert(0, str(ROOT / "msda-triton_b200"))
    from msda_triton.distributed import pixel_chunk, shard_range
    for total in (0, 1, 7, 64, 22223):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
            assert pixel_chunk(total, world) * world >= total
