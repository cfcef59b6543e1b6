"""2-GPU NCCL run of the sharding test (skipped on a 1-GPU box; run with `gpurun --gpus 2`)."""
import pytest
import torch

from test_distributed_cpu import run_world

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_unsharded_nccl_world2():
    run_world(2, "nccl", "cuda")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tensors_on_a_device_that_is_not_current():
    """Inputs on cuda:1 while cuda:0 is the current device: the launch goes to cuda:1's current stream, the current
    device is restored, and results equal those computed with cuda:1 current (reference: Triton launches under
    torch.cuda.device_of as well)."""
    import msda_triton
    from util import BENCH_PYRAMID, make_inputs
    assert torch.cuda.current_device() == 0
    img, shapes, pts, aw, go = make_inputs(2, 300, 8, 32, BENCH_PYRAMID, 4, seed=5)
    a, b, c = (t.to("cuda:1").requires_grad_(True) for t in (img, pts, aw))
    out = msda_triton.multiscale_deformable_attention(a, shapes.to("cuda:1"), b, c, "border", True)
    out.backward(go.to("cuda:1"))
    assert torch.cuda.current_device() == 0 and out.device.index == 1
    with torch.cuda.device(1):
        x, y, z = (t.to("cuda:1").requires_grad_(True) for t in (img, pts, aw))
        want = msda_triton.multiscale_deformable_attention(x, shapes.to("cuda:1"), y, z, "border", True)
        want.backward(go.to("cuda:1"))
    torch.cuda.synchronize(1)
    assert torch.equal(out, want)
    assert torch.allclose(a.grad, x.grad, rtol=1e-4, atol=1e-5)
    assert torch.equal(b.grad, y.grad) and torch.equal(c.grad, z.grad)
