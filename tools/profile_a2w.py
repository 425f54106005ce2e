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


def timed(fn, n=20):
    """device time of fn() alone: CUDA graphs of [L2 flush; fn] and [L2 flush] replayed back to back, difference of
    the two (no host launch gaps, cold L2, clocks under load)"""
    def graph_ms(body):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def both():
        flush.fill_(1.0)
        fn()
    return (graph_ms(both) - graph_ms(lambda: flush.fill_(1.0))) * 1e3


w, T, last, i_s, i_e = ru.alpha2weight(alpha, ray_id, N)
print("M = %d" % M)
print("alpha2weight fwd (bounds + march): %.1f us" % timed(lambda: ru.alpha2weight(alpha, ray_id, N)))
print("alpha2weight bwd:                  %.1f us" % timed(
    lambda: ru.alpha2weight_backward(alpha, w, T, last, i_s, i_e, N, gw, gl)))
print("raw2alpha fwd:                     %.1f us" % timed(lambda: ru.raw2alpha(dens, synthetic.act_shift(), 0.5)))
dens_grid = density.to(dev)
print("trilinear fwd C=1:                 %.1f us" % timed(lambda: ops.trilinear(dens_grid, pts, mn, mx)))
rgb = torch.rand(M, 3, device=dev)
print("segment_sum [M,3]:                 %.1f us" % timed(lambda: ops.segment_sum(rgb, ray_id, N)))
