#!/usr/bin/env python
"""Stress loop for the tcgen05 kernels (looking for an intermittent 'tensor memory not completely freed'):
    python tools/stress_tmem.py [seconds]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from poseprobe_b200 import synthetic, voxurf  # noqa: E402

dev = torch.device("cuda:0")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
sdf, k0 = bench._sphere_sdf(256, dev, 4242)
ws, bs = bench._fine_net()
eng = voxurf.FusedVoxurf(sdf, k0, ws, bs, bench.XYZ_MIN, bench.XYZ_MAX, stepsize=0.5, s_val=0.05, bg=0.0, sdf_alpha=10.0,
                         sdf_beta=0.35, posbase_pe=5, viewbase_pe=1)
pool = []
for c in range(8):
    o, d, v = synthetic.make_rays(8192, seed=900 + c)
    pool.append(tuple(x.to(dev) for x in (o, d, v)))
out_rgb = torch.empty(8192, 3, pin_memory=True)
t0 = time.time()
it = 0
sizes = [int(v) for v in os.environ.get('CHUNKS', '8192,8192,8192,8192,8192,8192,8192,8192,8192,1024').split(',')]
try:
    while time.time() - t0 < (0.5 if os.environ.get('MLP_ONLY') else budget):
        for c, n in enumerate(sizes):
            o, d, v = pool[c % 8]
            res = eng.render(o[:n], d[:n], v[:n], want_depth=True)
            out_rgb[:n].copy_(res[0], non_blocking=True)
        it += 1
        if it % 20 == 0:
            torch.cuda.synchronize()
            if os.environ.get('VERBOSE'):
                print('it', it, flush=True)
    torch.cuda.synchronize()
    print("render stress ok: %d images-worth of chunks in %.1f s" % (it, time.time() - t0))
except Exception as e:
    print("render stress FAILED at iteration %d: %s" % (it, str(e)[:200]))
    sys.exit(1)

# ---- tight loop on the MLP kernel alone (rows of the last march stay in the buffers) ----
b = eng._bufs
n = 8192
t0 = time.time()
it = 0
try:
    while time.time() - t0 < budget:
        for _ in range(50):
            with torch.cuda.device(dev):
                eng._mlp_forward(b, n)
        it += 50
        torch.cuda.synchronize()
    print("mlp-only stress ok: %d launches in %.1f s" % (it, time.time() - t0))
except Exception as e:
    print("mlp-only stress FAILED after %d launches: %s" % (it, str(e)[:120]))
    sys.exit(1)

# ---- train loops: rgbnet forward + fused backward, then the deformation network forward / backward ----
if os.environ.get("TRAIN"):
    del eng
    torch.cuda.empty_cache()
    sdf, k0 = bench._sphere_sdf(160, dev, 7)
    eng = voxurf.FusedVoxurf(sdf, k0, ws, bs, bench.XYZ_MIN, bench.XYZ_MAX, stepsize=0.5, s_val=0.05, bg=0.0,
                             sdf_alpha=10.0, sdf_beta=0.35, lr_sdf=1e-4, lr_k0=1e-2, lr_rgbnet=1e-3, lr_warp=1e-6,
                             w_eikonal=0.1, w_deform=0.1)
    dims = [3, 128, 128, 128, 128, 4]
    wws = [torch.randn(dims[i + 1], dims[i]) * (2.0 / dims[i]) ** 0.5 for i in range(5)]
    wws[-1] = 0.005 * torch.randn(4, 128)
    wbs = [0.1 * torch.randn(dims[i + 1]) for i in range(5)]
    eng.set_warp_network(wws, wbs, 2.0)
    tgt = torch.rand(8192, 3, device=dev)
    for use_deform in (False, True):
        t0 = time.time()
        it = 0
        try:
            while time.time() - t0 < budget / 2:
                o, d, v = pool[it % 8]
                loss = eng.train_step(o, d, v, tgt, use_deform=use_deform)
                it += 1
                if it % 10 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            assert torch.isfinite(loss).all()
            print("train stress ok (use_deform=%s): %d steps in %.1f s, loss %.5f" % (use_deform, it, time.time() - t0,
                                                                                     float(loss[0])))
        except Exception as e:
            print("train stress FAILED (use_deform=%s) at step %d: %s" % (use_deform, it, str(e)[:160]))
            sys.exit(1)
