"""Probe torch symmetric memory (peer pointers, multicast, barrier) on this box."""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = 64 * 1024 * 1024
t = symm_mem.empty((n,), dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "mc", hex(hdl.multicast_ptr or 0),
      flush=True)
t.fill_(float(rank + 1))
hdl.barrier()
peer = (rank + 1) % world
pbuf = hdl.get_buffer(peer, (n,), torch.float64)
chunk = n // world
src = torch.full((chunk,), 10.0 + rank, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
hdl.barrier()
t0 = time.perf_counter()
for _ in range(10):
    pbuf[rank * chunk:(rank + 1) * chunk].copy_(src)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
hdl.barrier()
torch.cuda.synchronize()
print(rank, "peer copy GB/s", chunk * 8 / dt / 1e9, "my buffer now", t[peer * chunk].item(),
      t[rank * chunk].item(), flush=True)
dist.destroy_process_group()
