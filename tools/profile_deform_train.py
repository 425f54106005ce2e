#!/usr/bin/env python
"""A few live train steps (FusedVoxurf.train_step(use_deform=True), recon_scene.py:603) at the fine-stage shapes, for
an ncu launch list: python tools/profile_deform_train.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from poseprobe_b200 import synthetic, voxurf  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(777)
sdf = synthetic.make_sdf_grid(160, radius=0.5, noise=0.02)
k0 = 0.1 * torch.randn(1, 12, 160, 160, 160)
ws, bs = bench._fine_net()
eng = voxurf.FusedVoxurf(sdf.to(dev), k0.to(dev), ws, bs, bench.XYZ_MIN, bench.XYZ_MAX, stepsize=0.5, s_val=0.05, bg=0.0,
                         sdf_alpha=10.0, sdf_beta=0.35, lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3, lr_warp=1e-6,
                         w_entropy=0.01, w_mask=0.1, w_tv_k0=1e-5, w_eikonal=0.1, w_deform=0.1)
dims = [3, 128, 128, 128, 128, 4]
wws = [torch.randn(dims[i + 1], dims[i]) * (2.0 / dims[i]) ** 0.5 for i in range(5)]
wws[-1] = 0.005 * torch.randn(4, 128)
wbs = [0.1 * torch.randn(dims[i + 1]) for i in range(5)]
eng.set_warp_network(wws, wbs, 2.0)
o, d, v = (t.to(dev) for t in synthetic.make_rays(8192, seed=777))
tgt = torch.rand(8192, 3, device=dev)
mask = (torch.rand(8192, device=dev) > 0.3).float()
for _ in range(3):
    eng.train_step(o, d, v, tgt, use_deform=True, mask=mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.train_step(o, d, v, tgt, use_deform=True, mask=mask)
e1.record()
torch.cuda.synchronize()
print("deform train step: %.3f ms" % (e0.elapsed_time(e1) / 5))
