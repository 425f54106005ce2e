#!/usr/bin/env python
"""Golden vectors for the pose chain and ray generation, produced by the REFERENCE'S OWN CODE run in this container:
classes `Pose`, `Lie` of lib/camera.py (cannot be imported: easydict, termcolor, ipdb missing) and `get_rays`,
`get_rays_of_a_view` of lib/voxurf_coarse.py, extracted with `ast` and executed unmodified on the CPU.

    python tests/golden/make_golden_camera.py        # needs /root/reference, CPU only
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_voxurf import REF, extract  # noqa: E402


def main():
    ns = {"torch": torch, "np": np}
    exec(compile(extract(os.path.join(REF, "lib", "camera.py"), ["Pose", "Lie"]), "lib/camera.py", "exec"), ns)
    ns_r = {"torch": torch, "np": np}
    exec(compile(extract(os.path.join(REF, "lib", "voxurf_coarse.py"), ["get_rays", "ndc_rays", "get_rays_of_a_view"]),
                 "lib/voxurf_coarse.py", "exec"), ns_r)
    pose, lie = ns["Pose"](), ns["Lie"]()
    torch.manual_seed(7)
    n_img = 5
    wu = torch.randn(n_img, 6) * 0.3
    wu[0] = 0                                   # theta = 0: the Taylor branches
    wu[1, :3] *= 1e-4
    SE3 = lie.se3_to_SE3(wu)
    pnp = lie.se3_to_SE3(torch.randn(n_img, 6) * 0.5)
    pnp[..., 3] += torch.tensor([0.0, 0.0, 2.5])
    composed = pose.compose_pair(SE3, pnp)      # recon_scene.py:67-69: refine o pnp
    c2w = pose.invert(composed)
    out = {"wu": wu.numpy(), "SE3": SE3.numpy(), "pnp": pnp.numpy(), "composed": composed.numpy(), "c2w": c2w.numpy()}
    H, W = 9, 13
    K = torch.tensor([[40.0, 0, 6.2], [0, 42.0, 4.4], [0, 0, 1]])
    for inverse_y in (False, True):
        ro, rd, vd = ns_r["get_rays_of_a_view"](H, W, K, c2w[2], ndc=False, inverse_y=inverse_y, flip_x=False,
                                                  flip_y=inverse_y, mode="center")
        tag = "iy%d" % int(inverse_y)
        out["rays_o_" + tag], out["rays_d_" + tag], out["viewdirs_" + tag] = ro.numpy(), rd.numpy(), vd.numpy()
    out["K"], out["HW"] = K.numpy(), np.array([H, W])
    path = os.path.join(ROOT, "tests", "golden", "camera.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
