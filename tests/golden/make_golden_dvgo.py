#!/usr/bin/env python
"""Golden vectors for oracle/dvgo_cpu.DenseBaseline (the CPU baseline / `--impl reference` arm of bench.py), produced
by the REFERENCE'S OWN CODE run in this container: class `DirectVoxGO` and its helpers from lib/dvgo_ori.py and the
optimiser `Adam` / `adam` from lib/utils.py, extracted with `ast` and executed unmodified (lib/dvgo_ori.py cannot be
imported: mcubes, matplotlib are missing).  One forward + three optimisation steps of the coarse stage (rgbnet_dim=0).

    python tests/golden/make_golden_dvgo.py        # needs /root/reference, CPU only
"""
import math
import os
import sys
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_voxurf import REF, extract  # noqa: E402


def main():
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "os": os, "time": __import__("time"), "math": math}
    exec(compile(extract(os.path.join(REF, "lib", "dvgo_ori.py"),
                         ["DirectVoxGO", "MaskCache", "cumprod_exclusive", "get_ray_marching_ray", "total_variation"]),
                 "lib/dvgo_ori.py", "exec"), ns)
    ns_u = {"torch": torch, "math": math, "List": List, "Optional": Optional, "Tensor": torch.Tensor}
    exec(compile(extract(os.path.join(REF, "lib", "utils.py"), ["Adam", "adam"]), "lib/utils.py", "exec"), ns_u)
    DirectVoxGO, Adam = ns["DirectVoxGO"], ns_u["Adam"]

    torch.manual_seed(99)
    G, N = 24, 128
    model = DirectVoxGO(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=G ** 3, num_voxels_base=G ** 3,
                        alpha_init=1e-2, fast_color_thres=0, rgbnet_dim=0)
    assert list(model.density.shape[2:]) == [G, G, G]
    with torch.no_grad():
        model.density.copy_(torch.randn(1, 1, G, G, G))
        model.k0.copy_(0.1 * torch.randn(1, 3, G, G, G))
    from poseprobe_b200 import synthetic
    rays_o, rays_d, viewdirs = synthetic.make_rays(N, seed=55)
    target = torch.rand(N, 3)
    rk = dict(near=0.2, far=1e9, stepsize=0.5, bg=0.3)
    out = {"density0": model.density.detach().numpy().copy(), "k00": model.k0.detach().numpy().copy(),
           "rays_o": rays_o.numpy(), "rays_d": rays_d.numpy(), "target": target.numpy(),
           "voxel_size": np.float64(float(model.voxel_size))}
    with torch.no_grad():
        ret = model(rays_o, rays_d, viewdirs, global_step=None, **rk)
    out["rgb_marched"] = ret["rgb_marched"].numpy()
    out["alphainv_last"] = ret["alphainv_cum"][..., -1].numpy()
    out["n_inbox"] = np.int64(int((~ret["mask_outbbox"]).sum()))
    # three steps of the live optimiser (lib/utils.py:53-198; betas as create_optimizer_or_freeze_model :342)
    opt = Adam([{"params": [model.density], "lr": 0.1}, {"params": [model.k0], "lr": 0.1}], betas=(0.9, 0.99))
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        ret = model(rays_o, rays_d, viewdirs, global_step=None, **rk)
        loss = F.mse_loss(ret["rgb_marched"], target)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    out["losses"] = np.array(losses)
    out["density3"], out["k03"] = model.density.detach().numpy(), model.k0.detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "dvgo_dense.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes, losses", losses)


if __name__ == "__main__":
    main()
