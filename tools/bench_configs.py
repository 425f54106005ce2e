#!/usr/bin/env python
"""BASELINE configs[3] and configs[4] on one GPU (parity-case configs measured for reference, not bench lines):

  D: 256^3 grid, 800x800 = 640 000 rays, forward only (full-image render in 8192-ray chunks), SDF + 12-channel
     features + rgbnet on tcgen05 (FusedVoxurf.render) and the coarse DVGO grid (FusedDVGO.render).
  E: dual branch, one rank's share: object 160^3 + "scene" 320^3 fine-stage grids, 8192 rays per step through each
     (synthetic: the reference's scene branch is an MLP NeRF), fwd + bwd + Adam.
"""
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseprobe_b200 import dvgo, ops, synthetic, voxurf  # noqa: E402

MN, MX = [-1.0, -1, -1], [1.0, 1, 1]


def make_net():
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    return ops.rgbnet_from_sequential(net)


def fine_engine(G, dev, s_val=0.05):
    sdf = synthetic.make_sdf_grid(G, radius=0.5, noise=0.02)
    k0 = 0.1 * torch.randn(1, 12, G, G, G)
    ws, bs = make_net()
    return voxurf.FusedVoxurf(sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], MN, MX,
                              stepsize=0.5, s_val=s_val, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35, lr_sdf=1e-4,
                              lr_k0=1e-2, lr_rgbnet=1e-3)


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(777)
    out = {}
    # ---- config D: full-image render, 256^3 ----
    chunks = []
    for c in range(8):
        o, d, v = synthetic.make_rays(8192, seed=900 + c)
        chunks.append(tuple(x.to(dev) for x in (o, d, v)))
    n_chunks = 640000 // 8192 + 1
    eng = fine_engine(256, dev)
    state = {"i": 0}

    def render_image_fine():
        for c in range(n_chunks):
            eng.render(*chunks[c % 8])
    ms = timed(render_image_fine, 2)
    m = eng.n_samples()
    out["D_fine_256"] = {"workload": "800x800 image = %d chunks of 8192 rays, 256^3 SDF + 12-ch k0 + rgbnet (tcgen05), fwd only"
                                     % n_chunks, "ms_per_image": ms, "render_mrays_per_s": n_chunks * 8192 / (ms * 1e-3) / 1e6,
                         "composited_samples_per_ray": m / 8192}
    del eng
    torch.cuda.empty_cache()
    density, k0c = synthetic.make_dvgo_grids(256)
    dv = dvgo.FusedDVGO(density.to(dev), k0c.to(dev), MN, MX, stepsize=0.5)

    def render_image_coarse():
        for c in range(n_chunks):
            dv.render(*chunks[c % 8][:2], want_depth=True)
    ms = timed(render_image_coarse, 3)
    out["D_coarse_256"] = {"workload": "same image, 256^3 density + k0(3) grid (DVGO colour), fwd only",
                           "ms_per_image": ms, "render_mrays_per_s": n_chunks * 8192 / (ms * 1e-3) / 1e6}
    del dv
    torch.cuda.empty_cache()
    # ---- config E: object 160^3 + scene 320^3, 8192 rays through each per step ----
    obj, scene = fine_engine(160, dev), fine_engine(320, dev)
    tgt = torch.rand(8192, 3, device=dev)

    def dual_step():
        o, d, v = chunks[state["i"] % 8]
        state["i"] += 1
        obj.train_step(o, d, v, tgt)
        scene.train_step(o, d, v, tgt)
    ms = timed(dual_step, 6)
    out["E_dual_160_320"] = {"workload": "one rank of config E: fine-stage train step on a 160^3 object grid and a 320^3 "
                                         "scene grid, 8192 rays each, fwd+bwd+Adam",
                             "ms_per_step": ms, "train_rays_per_s": 8192 / (ms * 1e-3),
                             "samples_per_ray_obj": obj.n_samples() / 8192, "samples_per_ray_scene": scene.n_samples() / 8192,
                             "resident_GB": torch.cuda.memory_allocated() / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
