#!/usr/bin/env python
"""Golden vectors for the host-side grid modules (poseprobe_b200/grid.py), produced by the REFERENCE'S OWN
`DenseGrid` and `MaskGrid` (lib/grid.py:38-89, :212-245), extracted with `ast` and executed unmodified on the CPU.
lib/grid.py JIT-compiles two CUDA extensions at import; they are replaced by the C oracle here (the kernels themselves
have their own goldens), so what this pins is the host logic: the world->index affine buffers, the checkpoint ->
occupancy recipe, scale_volume_grid's resampling, the mask broadcast in total_variation_add_grad, output shapes.

    python tests/golden/make_golden_grid.py        # needs /root/reference, CPU only
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_voxurf import REF, extract  # noqa: E402


def main():
    from oracle import oracle

    def tv_plain(param, grad, wx, wy, wz, dense_mode):
        oracle.total_variation_add_grad(param.detach(), grad, wx, wy, wz, dense_mode)

    def tv_masked(param, grad, mask, wx, wy, wz, dense_mode):
        oracle.total_variation_add_grad(param.detach(), grad, wx, wy, wz, dense_mode, mask=mask)

    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "os": os,
          "render_utils_cuda": types.SimpleNamespace(maskcache_lookup=oracle.maskcache_lookup),
          "total_variation_cuda": types.SimpleNamespace(total_variation_add_grad=tv_plain,
                                                        total_variation_add_grad_new=tv_masked)}
    exec(compile(extract(os.path.join(REF, "lib", "grid.py"), ["DenseGrid", "MaskGrid"]), "lib/grid.py", "exec"), ns)
    DenseGrid, MaskGrid = ns["DenseGrid"], ns["MaskGrid"]

    torch.manual_seed(4242)
    out = {}
    lo, hi = [-1.0, -0.7, -1.3], [0.9, 1.1, 0.8]
    # MaskGrid from an explicit mask
    occ = torch.rand(5, 7, 6) > 0.5
    mg = MaskGrid(mask=occ, xyz_min=lo, xyz_max=hi)
    pts = (torch.rand(4, 50, 3) * 2.6 - 1.4)
    out.update(mg_mask=occ.numpy(), mg_lo=np.array(lo, np.float32), mg_hi=np.array(hi, np.float32),
               mg_scale=mg.xyz2ijk_scale.numpy(), mg_shift=mg.xyz2ijk_shift.numpy(), mg_pts=pts.numpy(),
               mg_hit=mg(pts).numpy())
    # MaskGrid from a coarse-stage checkpoint
    density = torch.randn(1, 1, 6, 5, 8) * 3
    ckpt = {"model_state_dict": {"density.grid": density, "act_shift": torch.tensor(-2.0)},
            "model_kwargs": {"voxel_size_ratio": 0.75, "xyz_min": lo, "xyz_max": hi}}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "coarse_last.tar")
        torch.save(ckpt, path)
        mc = MaskGrid(path=path, mask_cache_thres=0.3)
    out.update(ck_density=density.numpy(), ck_mask=mc.mask.numpy(), ck_scale=mc.xyz2ijk_scale.numpy(),
               ck_shift=mc.xyz2ijk_shift.numpy())
    # DenseGrid: resampling and the TV entry points
    dg = DenseGrid(channels=3, world_size=[4, 6, 5], xyz_min=lo, xyz_max=hi)
    with torch.no_grad():
        dg.grid.copy_(torch.randn_like(dg.grid))
    out["dg_grid"] = dg.grid.detach().numpy().copy()
    dg.grid.grad = torch.randn_like(dg.grid)
    out["dg_grad0"] = dg.grid.grad.numpy().copy()
    dg.total_variation_add_grad(0.3, 0.2, 0.1, True)
    out["dg_grad_tv"] = dg.grid.grad.numpy().copy()
    mask1 = (torch.rand(1, 1, 4, 6, 5) > 0.4)
    out["dg_tvmask"] = mask1.numpy()
    dg.total_variation_add_grad(0.3, 0.2, 0.1, False, mask=mask1)
    out["dg_grad_tv_masked"] = dg.grid.grad.numpy().copy()
    q = torch.rand(3, 20, 3) * 2.4 - 1.2
    out["dg_q"] = q.numpy()
    out["dg_val"] = dg(q).detach().numpy()
    dg -= 0.25
    out["dg_isub"] = dg.grid.detach().numpy().copy()
    dg.scale_volume_grid([7, 5, 9])
    out["dg_scaled"] = dg.grid.detach().numpy()
    d1 = DenseGrid(channels=1, world_size=[4, 6, 5], xyz_min=lo, xyz_max=hi)
    out["d1_shape"] = np.array(d1(q).shape)
    path = os.path.join(ROOT, "tests", "golden", "grid_modules.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
