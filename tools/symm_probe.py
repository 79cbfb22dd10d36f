"""Probe: does torch symmetric memory give usable peer pointers on this box?
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
try:
    n = 1 << 20
    t = symm.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
    t.fill_(float(rank + 1))
    hdl = symm.rendezvous(t, dist.group.WORLD)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    hdl.barrier()
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (n,), torch.float32)
    val = float(pt[:16].sum().item()) / 16
    torch.cuda.synchronize()
    # bandwidth of a plain peer read (torch copy kernel over NVLink)
    big = symm.empty(1 << 28, dtype=torch.float32, device=torch.device("cuda", local))  # 1 GiB
    h2 = symm.rendezvous(big, dist.group.WORLD)
    h2.barrier()
    src = h2.get_buffer(peer, (1 << 28,), torch.float32)
    dst = torch.empty(1 << 28, dtype=torch.float32, device=torch.device("cuda", local))
    dst.copy_(src)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(src)
    torch.cuda.synchronize()
    gbs = 5 * (1 << 30) / (time.perf_counter() - t0) / 1e9
    h2.barrier()
    print(f"rank {rank}: ptrs ok ({len(ptrs)}), peer value {val} (want {peer + 1}), "
          f"peer read {gbs:.0f} GB/s", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"rank {rank}: symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
