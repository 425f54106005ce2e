#!/usr/bin/env python
"""Stand-alone run of the op-level compositing kernels on config-B shapes (for ncu / quick timing):
    python tools/profile_a2w.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poseprobe_b200 as pp  # noqa: E402
from poseprobe_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
G, N = 160, 8192
o, d, _ = synthetic.make_rays(N)
mn, mx = torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1])
ru = pp.render_utils_cuda
pts, mask, ray_id, *_ = ru.sample_pts_on_rays(o.to(dev), d.to(dev), mn.to(dev), mx.to(dev), 0.2, 1e9, 0.5 * 2 / G)
keep = ~mask
pts, ray_id = pts[keep].contiguous(), ray_id[keep].contiguous()
density, _ = synthetic.make_dvgo_grids(G)
dens = ops.trilinear(density.to(dev), pts, mn, mx)[:, 0].contiguous()
_, alpha = ru.raw2alpha(dens, synthetic.act_shift(), 0.5)
M = alpha.numel()
gw, gl = torch.randn(M, device=dev), torch.randn(N, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a long kernel in front hides the host-side launch overhead of the op behind it
        flush.fill_(2.0)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


w, T, last, i_s, i_e = ru.alpha2weight(alpha, ray_id, N)
print("M = %d" % M)
print("alpha2weight fwd (bounds + march): %.1f us" % timed(lambda: ru.alpha2weight(alpha, ray_id, N)))
print("alpha2weight bwd:                  %.1f us" % timed(
    lambda: ru.alpha2weight_backward(alpha, w, T, last, i_s, i_e, N, gw, gl)))
print("raw2alpha fwd:                     %.1f us" % timed(lambda: ru.raw2alpha(dens, synthetic.act_shift(), 0.5)))
print("trilinear fwd C=1:                 %.1f us" % timed(lambda: ops.trilinear(density.to(dev), pts, mn, mx)))
