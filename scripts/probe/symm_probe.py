"""Does torch symmetric memory work on this box?  (2 ranks: allocate, rendezvous, read the peer's buffer with a torch op.)"""
import os
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
t = symm.empty((1024,), dtype=torch.float32, device=torch.device("cuda", local))
t.fill_(float(rank + 1))
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), "signal pads", len(hdl.signal_pad_ptrs), flush=True)
hdl.barrier(channel=0)
torch.cuda.synchronize()
dist.destroy_process_group()
