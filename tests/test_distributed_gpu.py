"""2-GPU NCCL run of the sharding test (skipped on a 1-GPU box; run with `gpurun --gpus 2`)."""
import pytest
import torch

from test_distributed_cpu import run_world

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_unsharded_nccl_world2():
    run_world(2, "nccl", "cuda")
