#!/usr/bin/env python
"""Golden vectors for the fine-stage oracle, produced by the REFERENCE'S OWN CODE run in this container.

`lib/voxurf_coarse.py` cannot be imported here (torch_scatter, mcubes, ... are missing and its constructor calls
.cuda()), so this script extracts the class `Voxurf`, `Alphas2Weights` and `total_variation` (and `DenseGrid` from
`lib/grid.py`) from the reference sources with `ast`, executes them UNMODIFIED in a namespace where only the
third-party / native pieces are replaced -

    torch_scatter.segment_coo   -> index_add_ (reduce='sum', sorted index: the documented semantics)
    render_utils_cuda.alpha2weight[_backward] -> oracle/pp_oracle.c (bit-exact with the reference kernel, see
                                   tests/test_oracle_golden.py)
    Tensor.cuda()               -> identity (CPU run)

- builds an instance without running `__init__` (attributes set as the constructor would for the shipped surface
config, configs/dtu_e2e/scan1.py), and calls `Voxurf.forward(..., use_deform=False)` and `total_variation`.
Outputs go to tests/golden/voxurf_forward.npz; tests/test_oracle_golden.py replays them against oracle/voxurf_cpu.py.

    python tests/golden/make_golden_voxurf.py        # needs /root/reference, CPU only
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("PP_REFERENCE", "/root/reference")


def extract(path, names):
    """source text of the top-level classes / functions `names` of a reference file, unmodified"""
    src = open(path).read()
    tree = ast.parse(src)
    lines = src.splitlines(keepends=True)
    out = []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            start = min([node.lineno] + [d.lineno for d in node.decorator_list]) - 1
            out.append("".join(lines[start:node.end_lineno]))
    assert len(out) == len(names), (path, names)
    return "\n\n".join(out)


def segment_coo(src, index, out, reduce="sum"):
    assert reduce == "sum"
    return out.index_add_(0, index, src)


def reference_model():
    """-> (reference Voxurf instance on the CPU, namespace the class was executed in, dict of the sizes used)"""
    from einops import rearrange
    from oracle import oracle
    torch.Tensor.cuda = lambda self, *a, **k: self            # CPU run
    ru = types.SimpleNamespace(
        alpha2weight=lambda alpha, ray_id, n: list(oracle.alpha2weight(alpha.detach(), ray_id, n)),
        alpha2weight_backward=oracle.alpha2weight_backward)
    ns_grid = {"torch": torch, "nn": nn, "F": F, "np": np, "os": os}
    exec(compile(extract(os.path.join(REF, "lib", "grid.py"), ["DenseGrid"]), "lib/grid.py", "exec"), ns_grid)
    cudnn = types.SimpleNamespace(enabled=True)
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "os": os, "time": __import__("time"), "cudnn": cudnn,
          "rearrange": rearrange, "segment_coo": segment_coo, "render_utils_cuda": ru,
          "grid": types.SimpleNamespace(**{k: v for k, v in ns_grid.items() if k == "DenseGrid"}),
          "camera": None, "DeformedImplicitField": None, "extract_geometry": None}
    exec(compile(extract(os.path.join(REF, "lib", "voxurf_coarse.py"), ["Voxurf", "total_variation", "Alphas2Weights"]),
                 "lib/voxurf_coarse.py", "exec"), ns)
    Voxurf, DenseGrid, total_variation = ns["Voxurf"], ns_grid["DenseGrid"], ns["total_variation"]

    torch.manual_seed(1234)
    G, N, N_iters, gstep = 16, 96, 1000, 300
    xyz_min, xyz_max = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    m = object.__new__(Voxurf)
    nn.Module.__init__(m)
    m.register_buffer("xyz_min", torch.Tensor(xyz_min))
    m.register_buffer("xyz_max", torch.Tensor(xyz_max))
    m.device = torch.device("cpu")
    m.N_iters, m.barf_c2f = N_iters, [0.1, 0.5]
    m.use_explict_deform, m.use_implict_deform, m.use_deep_sdf = False, True, False
    m.num_voxels = G ** 3
    m.voxel_size = ((m.xyz_max - m.xyz_min).prod() / m.num_voxels).pow(1 / 3)       # _set_grid_resolution
    m.world_size = ((m.xyz_max - m.xyz_min) / m.voxel_size).long()
    ws = m.world_size.tolist()
    assert ws == [G, G, G]
    m.sdf = DenseGrid(channels=1, world_size=ws, xyz_min=xyz_min, xyz_max=xyz_max)
    m.k0 = DenseGrid(channels=12, world_size=ws, xyz_min=xyz_min, xyz_max=xyz_max)
    lin = torch.linspace(-1, 1, G)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    with torch.no_grad():
        m.sdf.grid.copy_(((x * x + y * y + z * z).sqrt() - 0.5 + 0.01 * torch.randn(G, G, G))[None, None])
        m.k0.grid.copy_((0.3 * torch.randn(1, 12, G, G, G)).half().float())      # fp16-representable: small fixture
    posbase_pe, viewbase_pe, depth, width = 5, 1, 3, 128
    m.rgbnet_kwargs = {"posbase_pe": posbase_pe, "viewbase_pe": viewbase_pe}
    m.rgbnet_full_implicit, m.fast_color_thres, m.geo_rgb_dim = False, 0, 3
    m.register_buffer("posfreq", torch.FloatTensor([(2 ** i) for i in range(posbase_pe)]))
    m.register_buffer("viewfreq", torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
    dim0 = (3 + 3 * posbase_pe * 2) + (3 + 3 * viewbase_pe * 2) + 12 + 3
    m.rgbnet = nn.Sequential(nn.Linear(dim0, width), nn.ReLU(inplace=True),
                             *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
                             nn.Linear(width, 3))
    nn.init.constant_(m.rgbnet[-1].bias, 0)
    with torch.no_grad():
        m.rgbnet[0].weight.mul_(2.0)
        m.rgbnet[-1].weight.mul_(4.0)
    m.use_mapping_fun, m.activation = True, nn.Softplus(beta=10)
    m.sdf_alpha = nn.Parameter(torch.Tensor([10.0]))
    m.sdf_beta = nn.Parameter(torch.Tensor([0.35]))
    m.smooth_sdf, m.nearest, m.grad_mode = False, False, "interpolate"
    m.s_ratio, m.s_start, m.s_learn, m.step_start = 50, 0.2, False, 0
    m.s_val = nn.Parameter(torch.ones(1), requires_grad=False)
    m.progress = nn.Parameter(torch.tensor(0.))
    m.warp_network = types.SimpleNamespace(progress=nn.Parameter(torch.tensor(0.)))
    m.nonempty_mask = None
    return m, ns, dict(G=G, N=N, N_iters=N_iters, gstep=gstep)


def main():
    m, ns, sz = reference_model()
    total_variation = ns["total_variation"]
    N, N_iters, gstep = sz["N"], sz["N_iters"], sz["gstep"]
    from poseprobe_b200 import synthetic
    rays_o, rays_d, viewdirs = synthetic.make_rays(N, seed=77)
    render_kwargs = dict(near=0.2, far=1e9, stepsize=0.5, bg=0.3)
    out = {}
    for tag, gs in (("eval", None), ("train", gstep)):
        torch.manual_seed(4321)              # the only random draw of forward is the per-ray jitter (:713)
        with torch.no_grad():
            ret = m.forward(rays_o, rays_d, viewdirs, use_deform=False, global_step=gs, **render_kwargs)
        for k in ("rgb_marched", "alphainv_cum", "cum_weights", "weights", "raw_alpha", "raw_rgb", "depth", "gradient"):
            out["%s_%s" % (tag, k)] = ret[k].detach().numpy()
        out["%s_s_val" % tag] = np.float64(float(m.s_val))
        out["%s_k0_tv" % tag] = np.float64(float(ret["k0_tv"]))
    torch.manual_seed(4321)
    out["train_jitter"] = torch.rand(N, 1)[:, 0].numpy()      # rand_like(rng[:, [0]]) under the same seed
    # Voxurf.grid_sampler with the custom 8-gather sampler (`use_custom=True` -> grid_sample_3d, :545-659) and with
    # F.grid_sample (border), on points inside and slightly outside the box
    gs_grid = torch.randn(1, 4, 6, 7, 5)
    gs_xyz = (torch.rand(400, 3) * 2.2 - 1.1)
    out["gs_grid"], out["gs_xyz"] = gs_grid.numpy(), gs_xyz.numpy()
    with torch.no_grad():
        out["gs_custom"] = m.grid_sampler(gs_xyz, gs_grid, use_custom=True).numpy()
        out["gs_border"] = m.grid_sampler(gs_xyz, gs_grid).numpy()
    # total_variation with and without mask
    v = torch.randn(1, 3, 7, 6, 5)
    mask = torch.rand(1, 3, 7, 6, 5) > 0.3
    out["tv_v"], out["tv_mask"] = v.numpy(), mask.numpy()
    out["tv_plain"], out["tv_masked"] = float(total_variation(v)), float(total_variation(v, mask))
    out.update(sdf=m.sdf.grid.detach().numpy(), k0=m.k0.grid.detach().numpy().astype(np.float16),
               rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), viewdirs=viewdirs.numpy(),
               voxel_size=np.float64(float(m.voxel_size)), N_iters=N_iters, gstep=gstep)
    for i, p in enumerate([q for q in m.rgbnet.modules() if isinstance(q, nn.Linear)]):
        out["w%d" % i], out["b%d" % i] = p.weight.detach().numpy(), p.bias.detach().numpy()
    # query_sdf_point_wocuda_render (:839-920, the depth-only render of lib/recon_scene.py:622-637) on the template SDF,
    # with NON-unit directions (depth = t_min + n_step / |d|); deterministic (global_step=None), placed last so that the
    # random stream of everything above is unchanged
    with torch.no_grad():
        q_pts, q_mask, q_depth = m.query_sdf_point_wocuda_render(rays_o, 1.7 * rays_d, global_step=None, keep_dim=True,
                                                                 use_deform=False, **render_kwargs)
    out.update(query_pts=q_pts.numpy(), query_mask=q_mask.numpy(), query_depth=q_depth.numpy(),
               query_s_val=np.float64(float(m.s_val)), query_dir_scale=np.float64(1.7))
    path = os.path.join(ROOT, "tests", "golden", "voxurf_forward.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
