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
