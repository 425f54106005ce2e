#!/usr/bin/env python
"""NCCL all-reduce (sum, fp32) device time at a few message sizes under torchrun: the practical ceiling of the
NVSwitch reduction path on this box, to put the fused exchange kernels' numbers next to.  Rank 0 prints JSON."""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
out = []
for mb in (64, 256, 1024):
    n = mb * (1 << 20) // 4
    buf = torch.ones(n, device=dev)
    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    alg = mb * (1 << 20) / (ms * 1e-3) / 1e9
    out.append({"MiB": mb, "ms": ms, "algbw_GBps": alg, "busbw_GBps": alg * 2 * (world - 1) / world})
    del buf
if rank == 0:
    print(json.dumps({"world": world, "nccl_allreduce": out}))
dist.destroy_process_group()
