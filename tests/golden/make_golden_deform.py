#!/usr/bin/env python
"""Golden vectors for the implicit-deformation branch of the fine stage, produced by the
REFERENCE'S OWN CODE run in this container: `Voxurf.forward(use_deform=True)` (lib/voxurf_coarse.py:922-1091, deform
branch :958-985) with the reference's `DeformedImplicitField` (lib/deformation/deform_net.py) built from its own
`SingleBVPNet` / `FCBlock` / `BatchLinear` (lib/deformation/modules.py, meta_utils.py) and the regularisers of
`deform_implicit_loss` (lib/losses.py:6-24), all extracted with `ast` and executed unmodified on the CPU (same
stand-ins as make_golden_voxurf.py).  Stored: the forward outputs, the four constraint terms, and the gradients of
    L = mse(rgb_marched, target) + grad_constraint + 0.05 (grad_deform + sdf_correct + sdf_deform constraints)
with respect to the warp network, the colour network, k0 and the rays (the pose path).
The reference zero-initialises the warp network's last layer (modules.py:166-170), which would make the branch trivial;
the fixture overwrites that layer with small random values (stored).

    python tests/golden/make_golden_deform.py        # needs /root/reference, CPU only
"""
import math
import os
import re
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_voxurf import REF, extract, reference_model  # noqa: E402


def reference_warp_network(out_range):
    d = os.path.join(REF, "lib", "deformation")
    ns_meta = {"nn": nn, "re": re, "OrderedDict": OrderedDict}
    exec(compile(extract(os.path.join(d, "meta_utils.py"), ["MetaModule", "MetaSequential", "get_subdict"]),
                 "meta_utils.py", "exec"), ns_meta)
    ns_mod = {"torch": torch, "nn": nn, "np": np, "math": math, "F": F, "OrderedDict": OrderedDict,
              **{k: ns_meta[k] for k in ("MetaModule", "MetaSequential", "get_subdict")}}
    names = ["BatchLinear", "Sine", "FCBlock", "SingleBVPNet", "init_weights_normal", "init_weights_selu",
             "init_weights_elu", "init_weights_xavier", "sine_init", "first_layer_sine_init", "last_layer_sine_init"]
    exec(compile(extract(os.path.join(d, "modules.py"), names), "modules.py", "exec"), ns_mod)
    ns_net = {"torch": torch, "nn": nn, "F": F, "modules": types.SimpleNamespace(**{k: ns_mod[k] for k in names})}
    exec(compile(extract(os.path.join(d, "deform_net.py"), ["DeformedImplicitField"]), "deform_net.py", "exec"), ns_net)
    return ns_net["DeformedImplicitField"](range_shape=out_range)


def main():
    m, ns, sz = reference_model()
    N, gstep = sz["N"], sz["gstep"]
    ns_loss = {"torch": torch, "F": F}
    exec(compile(extract(os.path.join(REF, "lib", "losses.py"), ["deform_implicit_loss"]), "lib/losses.py", "exec"),
         ns_loss)
    deform_implicit_loss = ns_loss["deform_implicit_loss"]

    torch.manual_seed(2024)
    m.range_shape = (m.xyz_max - m.xyz_min)
    m.warp_network = reference_warp_network(m.range_shape)
    m.bool_mapping_deformspace = False
    linears = [q for q in m.warp_network.modules() if isinstance(q, nn.Linear)]
    assert [tuple(q.weight.shape) for q in linears] == [(128, 3), (128, 128), (128, 128), (128, 128), (4, 128)]
    assert float(linears[-1].weight.abs().sum()) == 0.0        # the reference's zero init of the last layer
    with torch.no_grad():
        linears[-1].weight.copy_(0.02 * torch.randn(4, 128))
        linears[-1].bias.copy_(0.01 * torch.randn(4))
        for q in linears[:-1]:
            q.bias.copy_(0.1 * torch.randn(128))
    for p in m.sdf.parameters():
        p.requires_grad = False                                 # voxurf_coarse.py:136-138

    from poseprobe_b200 import synthetic
    rays_o, rays_d, viewdirs = synthetic.make_rays(N, seed=78)
    target = torch.rand(N, 3)
    render_kwargs = dict(near=0.2, far=1e9, stepsize=0.5, bg=0.3)
    rays_o.requires_grad_(True)
    rays_d.requires_grad_(True)
    torch.manual_seed(4321)
    ret = m.forward(rays_o, rays_d, viewdirs, use_deform=True, global_step=gstep, **render_kwargs)
    torch.manual_seed(4321)
    jitter = torch.rand(N, 1)[:, 0]
    terms = deform_implicit_loss(ret, True)
    loss = F.mse_loss(ret["rgb_marched"], target) + terms["grad_constraint"] + 0.05 * (
        terms["grad_deform_constraint"] + terms["sdf_correct_constraint"] + terms["sdf_deform_constraint"])
    params = [p for q in linears for p in (q.weight, q.bias)]
    rgb_lin = [q for q in m.rgbnet.modules() if isinstance(q, nn.Linear)]
    rgb_params = [p for q in rgb_lin for p in (q.weight, q.bias)]
    grads = torch.autograd.grad(loss, params + rgb_params + [m.k0.grid, rays_o, rays_d])
    out = {"rays_o": rays_o.detach().numpy(), "rays_d": rays_d.detach().numpy(), "viewdirs": viewdirs.numpy(),
           "target": target.numpy(), "jitter": jitter.numpy(), "gstep": gstep, "N_iters": sz["N_iters"],
           "out_range": np.float64(float(m.range_shape.max())), "voxel_size": np.float64(float(m.voxel_size)),
           "s_val": np.float64(float(m.s_val)), "loss": np.float64(float(loss)),
           "sdf": m.sdf.grid.detach().numpy(), "k0": m.k0.grid.detach().numpy().astype(np.float16)}
    for k in ("rgb_marched", "alphainv_cum", "cum_weights", "weights", "raw_alpha", "gradient", "sdf_deform",
              "grad_deform", "sdf_correct"):
        out["ret_" + k] = ret[k].detach().numpy()
    for k, v in terms.items():
        out["term_" + k] = np.float64(float(v))
    for i, q in enumerate(linears):
        out["ww%d" % i], out["wb%d" % i] = q.weight.detach().numpy(), q.bias.detach().numpy()
        out["g_ww%d" % i], out["g_wb%d" % i] = grads[2 * i].numpy(), grads[2 * i + 1].numpy()
    base = 2 * len(linears)
    for i, q in enumerate(rgb_lin):
        out["w%d" % i], out["b%d" % i] = q.weight.detach().numpy(), q.bias.detach().numpy()
        out["g_w%d" % i], out["g_b%d" % i] = grads[base + 2 * i].numpy(), grads[base + 2 * i + 1].numpy()
    base += 2 * len(rgb_lin)
    g_k0 = grads[base].numpy()
    nz = np.flatnonzero(g_k0)
    out["g_k0_idx"], out["g_k0_val"] = nz.astype(np.int32), g_k0.reshape(-1)[nz]
    out["g_rays_o"], out["g_rays_d"] = grads[base + 1].numpy(), grads[base + 2].numpy()
    # the depth-only render through the deformation (query_sdf_point_wocuda_render, :839-920, use_deform=True: what
    # lib/recon_scene.py:622-637 calls every step), eval mode, non-unit directions; last, so nothing above changes
    q_pts, q_mask, q_depth = m.query_sdf_point_wocuda_render(rays_o, 1.3 * rays_d, global_step=None,
                                                             keep_dim=True, use_deform=True, **render_kwargs)
    out.update(query_pts=q_pts.detach().numpy(), query_mask=q_mask.numpy(), query_depth=q_depth.detach().numpy(),
               query_s_val=np.float64(float(m.s_val)), query_dir_scale=np.float64(1.3))
    path = os.path.join(ROOT, "tests", "golden", "voxurf_deform.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss), {k: float(v) for k, v in terms.items()})


if __name__ == "__main__":
    main()
