#!/usr/bin/env python
"""Golden vectors for the product's `DirectVoxGO` module (poseprobe_b200/dvgo.py), produced by the REFERENCE'S OWN
`DirectVoxGO.forward` / `MaskCache` (lib/dvgo_ori.py:14-475, extracted with `ast`, executed unmodified on the CPU):
the complete nine-entry result dict and the gradients of a loss on rgb_marched / depth / weights / alphainv_cum for

  a) the coarse configuration under TRAINING (global_step given: jittered samples) with fast_color_thres = 1e-4
     (masked samples contribute sigmoid(0) colour),
  b) the fine configuration: 12-channel feature grid + rgbnet (depth 3, width 128, posbase_pe 5, viewbase_pe 4) and a
     mask cache built from a coarse checkpoint, evaluation mode.

    python tests/golden/make_golden_dvgo_module.py        # needs /root/reference, CPU only
"""
import math
import os
import sys
import tempfile

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
    DirectVoxGO = ns["DirectVoxGO"]
    from poseprobe_b200 import synthetic
    out = {}
    G, N = 20, 96
    rk = dict(near=0.2, far=6.0, stepsize=0.5, bg=0.3)

    def record(tag, model, ret, params, extra_loss_w):
        for k in ("alphainv_cum", "weights", "rgb_marched", "raw_alpha", "raw_rgb", "depth", "disp", "mask",
                  "mask_outbbox"):
            out["%s_out_%s" % (tag, k)] = ret[k].detach().numpy()
        w_rgb, w_depth, w_w, w_c = extra_loss_w
        loss = (ret["rgb_marched"] * w_rgb).sum() + (ret["depth"] * w_depth).sum() + (ret["weights"] * w_w).sum() + \
            (ret["alphainv_cum"] * w_c).sum()
        grads = torch.autograd.grad(loss, [p for _, p in params])
        for (name, _), g in zip(params, grads):
            out["%s_grad_%s" % (tag, name)] = g.numpy()
        for name, p in model.state_dict().items():
            out["%s_param_%s" % (tag, name)] = p.detach().numpy().copy()

    # ---- a) coarse, training (jitter), fast_color_thres ----
    torch.manual_seed(5)
    model = DirectVoxGO(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=G ** 3, num_voxels_base=G ** 3,
                        alpha_init=1e-2, fast_color_thres=1e-4, rgbnet_dim=0)
    with torch.no_grad():
        model.density.copy_(torch.randn(1, 1, G, G, G) * 2 + 1)
        model.k0.copy_(torch.randn(1, 3, G, G, G))
    rays_o, rays_d, viewdirs = synthetic.make_rays(N, seed=56)
    S = int(np.linalg.norm(np.array([G, G, G]) + 1) / rk["stepsize"]) + 1
    ws = [torch.randn(N, 3), torch.randn(N) * 0.1, torch.randn(N, S) * 0.1, torch.randn(N, S + 1) * 0.1]
    torch.manual_seed(1234)
    ret = model(rays_o, rays_d, viewdirs, global_step=10, **rk)
    torch.manual_seed(1234)
    out["a_jitter"] = torch.rand(N, 1).numpy()          # the draw `rng += torch.rand_like(rng[:,[0]])` made
    record("a", model, ret, [("density", model.density), ("k0", model.k0)], ws)
    for k, v in zip(("rays_o", "rays_d", "viewdirs", "w_rgb", "w_depth", "w_weights", "w_cum"),
                    [rays_o, rays_d, viewdirs] + ws):
        out["a_" + k] = v.numpy()

    # ---- b) fine: feature grid + rgbnet + mask cache, evaluation ----
    torch.manual_seed(6)
    coarse_density = torch.randn(1, 1, 12, 12, 12) * 3
    ckpt = {"MaskCache_kwargs": {"xyz_min": [-1., -1., -1.], "xyz_max": [1., 1., 1.],
                                 "act_shift": float(np.log(1 / (1 - 1e-2) - 1)), "voxel_size_ratio": 1.0},
            "model_state_dict": {"density": coarse_density}}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "coarse_last.tar")
        torch.save(ckpt, path)
        model = DirectVoxGO(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=G ** 3, num_voxels_base=G ** 3,
                            alpha_init=1e-2, fast_color_thres=1e-4, rgbnet_dim=12, rgbnet_direct=False, rgbnet_depth=3,
                            rgbnet_width=128, posbase_pe=5, viewbase_pe=4, mask_cache_path=path,
                            mask_cache_thres=1e-3)
    with torch.no_grad():
        keep = model.density.detach() <= -100                      # written by _set_nonempty_mask
        model.density.copy_(torch.where(keep, model.density.detach(), torch.randn(1, 1, G, G, G) * 2 + 1))
        model.k0.copy_(torch.randn(1, 12, G, G, G) * 0.5)
    rays_o, rays_d, viewdirs = synthetic.make_rays(N, seed=57)
    ws = [torch.randn(N, 3), torch.randn(N) * 0.1, torch.randn(N, S) * 0.1, torch.randn(N, S + 1) * 0.1]
    ret = model(rays_o, rays_d, viewdirs, global_step=None, **rk)
    params = [("density", model.density), ("k0", model.k0)] + [("rgbnet." + n, p) for n, p in
                                                                 model.rgbnet.named_parameters()]
    record("b", model, ret, params, ws)
    out["b_coarse_density"] = coarse_density.numpy()
    out["b_nonempty_mask"] = model.nonempty_mask.numpy()
    for k, v in zip(("rays_o", "rays_d", "viewdirs", "w_rgb", "w_depth", "w_weights", "w_cum"),
                    [rays_o, rays_d, viewdirs] + ws):
        out["b_" + k] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "dvgo_module.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
