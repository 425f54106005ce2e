#!/usr/bin/env python
"""Config C (BASELINE configs[2]) on N GPUs under torchrun: Voxurf SDF fine stage, ray-sharded, 8192 rays per rank,
ONE all-reduce of the flat gradient buffer per step (DistributedVoxurf).  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseprobe_b200 import ops, synthetic, voxurf  # noqa: E402

G, N, STEPS = 160, 8192, int(os.environ.get("STEPS", 20))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(777)
    sdf = synthetic.make_sdf_grid(G, radius=0.5, noise=0.02)
    k0 = 0.1 * torch.randn(1, 12, G, G, G)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    ws, bs = ops.rgbnet_from_sequential(net)
    kw = dict(stepsize=0.5, s_val=0.05, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35, lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3)
    args = (sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], [-1.0, -1, -1], [1.0, 1, 1])
    tr = voxurf.DistributedVoxurf(*args, mode=os.environ.get('PP_VOXURF_DIST_MODE', 'auto'), **kw) if world > 1 else None
    eng = tr.eng if tr else voxurf.FusedVoxurf(*args, **kw)
    batches = []
    for b in range(4):
        o, d, v = synthetic.make_rays(N, seed=777 + 1000 * rank + b)
        t = torch.rand(N, 3, generator=torch.Generator().manual_seed(99 + 1000 * rank + b))
        batches.append(tuple(x.to(dev) for x in (o, d, v, t)))
    step = (lambda i: tr.train_step(*batches[i % 4])) if tr else (lambda i: eng.train_step(*batches[i % 4]))
    for i in range(4):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(STEPS):
        step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # the all-reduce alone (device time)
    ar_ms = 0.0
    if world > 1 and tr.mode == "nccl":
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.all_reduce(eng.grad_flat)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            dist.all_reduce(eng.grad_flat)
        z.record()
        torch.cuda.synchronize()
        ar_ms = a.elapsed_time(z) / 5
    if rank == 0:
        print(json.dumps({"workload": "Voxurf SDF fine stage (BASELINE configs[2]): 160^3 sdf + 12-ch k0 + rgbnet (tcgen05), "
                                      "8192 rays/step/GPU, fwd+bwd+Adam, ray-sharded dp%d, one flat gradient all-reduce" % world,
                          "mode": tr.mode if tr else "single", "n_gpus": world, "train_rays_per_s": world * N / (float(ms) * 1e-3), "ms_per_step": float(ms),
                          "allreduce_ms": ar_ms, "allreduce_bytes": eng.grad_flat.numel() * 4,
                          "samples_per_ray": eng.n_samples() / N}))
    if world > 1:
        tr.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
