#!/usr/bin/env python
"""Golden vectors for the scene-branch compositing, produced by the REFERENCE'S OWN CODE run
in this container: `NeRF.composite` (lib/bg_nerf/source/models/frequency_nerf.py:283-343 - exp-cumsum transmittance over
dense [B, R, N] samples, rgb / depth / opacity / variances / all_cumulated) and `Graph.sample_depth_from_pdf`
(lib/bg_nerf/source/models/renderer.py:703-738 - inverse-CDF resampling of the coarse weights), both methods extracted
with `ast` and executed unmodified on the CPU (self is a stub carrying `device`; `opt` carries the two flags the
methods read).  Stored: inputs, every output, and the gradients of a loss on rgb / depth / opacity / weights with
respect to rgb_samples, density_samples and the ray.

    python tests/golden/make_golden_bgnerf.py        # needs /root/reference, CPU only
"""
import ast
import os
import sys
import textwrap
import types
from typing import Any, Dict, List

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("POSEPROBE_REFERENCE", "/root/reference")


def extract_method(path, cls, name):
    """source text of method `name` of top-level class `cls`, unmodified (dedented)"""
    src = open(path).read()
    lines = src.splitlines(keepends=True)
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    return textwrap.dedent("".join(lines[sub.lineno - 1:sub.end_lineno]))
    raise KeyError((path, cls, name))


def reference_functions():
    d = os.path.join(REF, "lib", "bg_nerf", "source", "models")
    ns = {"torch": torch, "Dict": Dict, "Any": Any, "List": List}
    exec(compile(extract_method(os.path.join(d, "frequency_nerf.py"), "NeRF", "composite"), "frequency_nerf.py", "exec"), ns)
    exec(compile(extract_method(os.path.join(d, "renderer.py"), "Graph", "sample_depth_from_pdf"), "renderer.py", "exec"),
         ns)
    return ns["composite"], ns["sample_depth_from_pdf"]


def main():
    composite, sample_depth_from_pdf = reference_functions()
    me = types.SimpleNamespace(device="cpu")
    out = {}
    g = torch.Generator().manual_seed(4242)
    for tag, (B, R, N), opaque in (("a", (2, 37, 64), True), ("b", (1, 50, 128), False)):
        opt = types.SimpleNamespace(nerf=types.SimpleNamespace(setbg_opaque=opaque), mask_img=False)
        ray = torch.randn(B, R, 3, generator=g) * 0.7
        ray.requires_grad_(True)
        depth_min, depth_max = 0.5, 4.0
        # stratified depths as Graph.sample_depth produces them: one sample per interval of the depth range
        edges = torch.linspace(depth_min, depth_max, N + 1)
        depth = (edges[:-1] + (edges[1:] - edges[:-1]) * torch.rand(B, R, N, generator=g))[..., None]
        rgb_s = torch.rand(B, R, N, 3, generator=g).requires_grad_(True)
        # densities after softplus: mostly thin, some opaque stretches, a few exact zeros
        dens = torch.nn.functional.softplus(torch.randn(B, R, N, generator=g) * 3 - 2)
        dens = dens * (torch.rand(B, R, N, generator=g) > 0.1)
        dens.requires_grad_(True)
        pred = composite(me, opt, ray, {"rgb_samples": rgb_s, "density_samples": dens}, depth)
        w_rgb, w_depth, w_op = torch.randn(B, R, 3, generator=g), torch.randn(B, R, 1, generator=g), \
            torch.randn(B, R, 1, generator=g)
        w_w = torch.randn(B, R, N, 1, generator=g) * 0.1
        loss = (pred["rgb"] * w_rgb).sum() + (pred["depth"] * w_depth).sum() + (pred["opacity"] * w_op).sum() + \
            (pred["weights"] * w_w).sum()
        g_rgb, g_dens, g_ray = torch.autograd.grad(loss, [rgb_s, dens, ray])
        for k, v in (("ray", ray), ("depth_samples", depth), ("rgb_samples", rgb_s), ("density_samples", dens),
                     ("w_rgb", w_rgb), ("w_depth", w_depth), ("w_opacity", w_op), ("w_weights", w_w),
                     ("g_rgb_samples", g_rgb), ("g_density_samples", g_dens), ("g_ray", g_ray)):
            out["%s_%s" % (tag, k)] = v.detach().numpy()
        for k in ("rgb", "rgb_var", "depth", "depth_var", "opacity", "weights", "all_cumulated"):
            out["%s_out_%s" % (tag, k)] = pred[k].detach().numpy()
        out["%s_opaque" % tag] = np.int64(opaque)
        # hierarchical resampling of the coarse weights, deterministic and with the reference's shared random grid
        Nf = 48 if tag == "a" else 128
        w = pred["weights"][..., 0].detach()
        out["%s_fine_det" % tag] = sample_depth_from_pdf(me, opt, w, N, Nf, [depth_min, depth_max], True).numpy()
        torch.manual_seed(99)
        out["%s_fine_rand" % tag] = sample_depth_from_pdf(me, opt, w, N, Nf, [depth_min, depth_max], False).numpy()
        torch.manual_seed(99)
        out["%s_fine_grid" % tag] = torch.rand(Nf + 1).numpy()          # the grid the call above drew
        out["%s_depth_range" % tag] = np.array([depth_min, depth_max], dtype=np.float32)
    path = os.path.join(ROOT, "tests", "golden", "bgnerf_composite.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
