#!/usr/bin/env python
"""Config C (BASELINE configs[2], one rank's share): Voxurf SDF fine stage, 160^3 SDF + 12-channel feature grid,
rgbnet 57-128-128-128-3, 8192 rays / step, fwd + bwd + optimiser.  Prints a per-kernel breakdown (CUDA events) and
one JSON line."""
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseprobe_b200 import ops, synthetic, voxurf  # noqa: E402

G = int(os.environ.get("G", 160))
N = int(os.environ.get("N", 8192))
STEPS = int(os.environ.get("STEPS", 20))
MN, MX = [-1.0, -1, -1], [1.0, 1, 1]


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(777)
    sdf = synthetic.make_sdf_grid(G, radius=0.5, noise=0.02)
    k0 = 0.1 * torch.randn(1, 12, G, G, G)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)), nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    ws, bs = ops.rgbnet_from_sequential(net)
    eng = voxurf.FusedVoxurf(sdf.to(dev), k0.to(dev), [w.detach() for w in ws], [b.detach() for b in bs], MN, MX,
                             stepsize=0.5, s_val=0.05, bg=0.0, sdf_alpha=10.0, sdf_beta=0.35, posbase_pe=5,
                             viewbase_pe=1, lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3)
    batches = []
    for b in range(4):
        o, d, v = synthetic.make_rays(N, seed=777 + b)
        t = torch.rand(N, 3, generator=torch.Generator().manual_seed(99 + b))
        batches.append(tuple(x.to(dev) for x in (o, d, v, t)))
    for i in range(5):
        eng.train_step(*batches[i % 4])
    torch.cuda.synchronize()
    m = eng.n_samples()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(STEPS):
        eng.train_step(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    # breakdown: time each stage by re-running it
    import ctypes
    from poseprobe_b200._lib import check, lib, ptr, stream
    o, d, v, t = batches[0]
    b = eng._buffers(N)

    def timeit(fn, n=10):
        fn()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        z.record()
        torch.cuda.synchronize()
        return a.elapsed_time(z) / n

    t_march = timeit(lambda: eng._march(b, o, d, v, None))
    t_fwd = timeit(lambda: eng._mlp_forward(b, N))
    eng.loss_buf.zero_()
    check(lib.pp_voxurf_ray_loss(N, ptr(b["acc"]), ptr(b["cum_w"]), ptr(t), 0.0, 1.0, ptr(b["g_ray"]), ptr(b["rgb"]),
                                 ptr(eng.loss_buf), stream()), "loss")
    t_bwd_mlp = timeit(lambda: check(lib.pp_rgbnet_backward_fused(
        ptr(b["feat"]), b["cap"], ptr(b["counter"]), eng.d0, eng.depth, 128, eng._arr(eng.weights), eng._arr(eng.biases),
        ptr(b["w"]), ptr(b["rid"]), ptr(b["g_ray"]), ptr(b["gw"]), ptr(b["dx"]), 16, eng._arr(eng.g_w), eng._arr(eng.g_b),
        stream()), "bwd"))
    t_back = timeit(lambda: check(lib.pp_voxurf_backward(
        ctypes.byref(eng.scene), ctypes.byref(eng.cfg), ptr(o), ptr(d), ptr(v), None, N, ptr(b["rec"]), ptr(b["gw"]),
        ptr(b["dx"]), 16, ptr(b["ray_off"]), ptr(b["ray_cnt"]), ptr(b["ray_end"]), ptr(b["g_ray"]), ptr(eng.g_srec),
        ptr(eng.g_k0rec), None, None, None, None, None, None, stream()), "back"))
    t_opt = timeit(eng.optimizer_step)
    flops = 2.0 * m * (64 * 128 + 128 * 128 * 2 + 128 * 16) * 4     # fwd + recompute + 2x bwd, padded shapes
    line = {"workload": "Voxurf SDF fine stage (BASELINE configs[2], one rank): %d^3 sdf + 12-ch k0, rgbnet 57-128x3-3, "
                        "%d rays/step, stepsize 0.5, fwd+bwd+Adam" % (G, N),
            "train_rays_per_s": N / (ms * 1e-3), "ms_per_step": ms, "composited_samples_per_step": m,
            "samples_per_ray": m / N,
            "breakdown_ms": {"march_features": t_march, "rgbnet_forward": t_fwd, "rgbnet_backward_fused": t_bwd_mlp,
                             "reverse_march_scatter": t_back, "optimizer_dense_tail": t_opt},
            "mlp_tflops_effective": flops / ((t_fwd + t_bwd_mlp) * 1e-3) / 1e12}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
