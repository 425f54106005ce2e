#!/usr/bin/env python
"""Render with the implicit deformation (Voxurf.inference, lib/voxurf_coarse.py:1093-1215) at the fine-stage shapes:
G^3 SDF + 12-channel feature grid, warp network 3-128x4-4 with its Jacobian, rgbnet 57-128-128-128-3, N-ray chunks.
Per-stage CUDA-event times and one JSON line (Mrays/s, samples per ray, achieved tensor TFLOP/s of the warp network)."""
import ctypes
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseprobe_b200 import ops, synthetic, voxurf  # noqa: E402
from poseprobe_b200._lib import check, lib, ptr, stream  # noqa: E402

G = int(os.environ.get("G", 160))
N = int(os.environ.get("N", 8192))
MN, MX = [-1.0, -1, -1], [1.0, 1, 1]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    z.record()
    torch.cuda.synchronize()
    return a.elapsed_time(z) / n


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(777)
    sdf = synthetic.make_sdf_grid(G, radius=0.5, noise=0.02)
    k0 = 0.1 * torch.randn(1, 12, G, G, G)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    ws, bs = ops.rgbnet_from_sequential(net)
    eng = voxurf.FusedVoxurf(sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], MN, MX,
                             stepsize=0.5, s_val=0.05, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35)
    dims = [3, 128, 128, 128, 128, 4]
    wws = [torch.randn(dims[i + 1], dims[i]) * (2.0 / dims[i]) ** 0.5 for i in range(5)]
    wws[-1] = 0.005 * torch.randn(4, 128)
    wbs = [0.1 * torch.randn(dims[i + 1]) for i in range(5)]
    eng.set_warp_network(wws, wbs, 2.0)
    o, d, v = (t.to(dev) for t in synthetic.make_rays(N, seed=777))
    t_plain = timeit(lambda: eng.render(o, d, v, use_deform=False))
    t_all = timeit(lambda: eng.render(o, d, v, use_deform=True))
    b = eng._bufs
    samples = b["dcounter"].tolist()[0]
    composited = eng.n_samples()
    wk, bk, R = eng.warp
    cap = b["cap"]
    with torch.cuda.device(dev):
        t_rows = timeit(lambda: check(lib.pp_voxurf_deform_rows(
            ctypes.byref(eng.scene), ptr(o), ptr(d), None, N, cap, ptr(b["drows"]), ptr(b["dray_off"]),
            ptr(b["dcounter"]), stream()), "rows"))
        t_net = timeit(lambda: check(lib.pp_warpnet_forward_packed(
            ptr(b["drows"]), 4 * cap, len(wk), 128, eng._arr(wk), eng._arr(bk), ptr(b["dout"]), ptr(b["dcounter"][2:]),
            stream()), "warpnet"))
        t_march = timeit(lambda: eng._march_deform(b, o, d, v, None)) - t_rows - t_net
        t_rgb = timeit(lambda: eng._mlp_forward(b, N))
    flop = 4 * samples * 2.0 * (16 * 128 + 3 * 128 * 128 + 128 * 16)
    out = {"workload": "render with implicit deformation, %d^3, %d rays/chunk" % (G, N),
           "render_deform_mrays_per_s": N / t_all / 1e3, "render_plain_mrays_per_s": N / t_plain / 1e3,
           "ms": {"all": t_all, "plain": t_plain, "rows": t_rows, "warpnet": t_net, "march_deform": t_march,
                  "rgbnet": t_rgb},
           "samples_per_ray": samples / N, "composited_per_ray": composited / N,
           "warpnet_tflops": flop / (t_net * 1e-3) / 1e12,
           "warpnet_rows_per_s": 4 * samples / (t_net * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
