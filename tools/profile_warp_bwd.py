#!/usr/bin/env python
"""Stand-alone run of the deformation network's tcgen05 kernels on a config-C sized batch (for ncu):
    python tools/profile_warp_bwd.py [n_samples]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poseprobe_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda:0")
torch.manual_seed(0)
dims = [3, 128, 128, 128, 128, 4]
ws = [(torch.randn(dims[i + 1], dims[i]) * (2.0 / dims[i]) ** 0.5).to(dev) for i in range(5)]
ws[-1] = (0.01 * torch.randn(4, 128)).to(dev)
bs = [(0.1 * torch.randn(dims[i + 1])).to(dev) for i in range(5)]
x = (torch.rand(n, 3) * 2 - 1).to(dev)
g = (1e-6 * torch.randn(n, 4, 4)).to(dev)
for it in range(3):
    ops.warpnet_forward(x, ws, bs, 2.0)
    ops.warpnet_backward(x, ws, bs, g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(5):
    ops.warpnet_backward(x, ws, bs, g)
e1.record()
torch.cuda.synchronize()
print("warpnet_backward (incl. pack + absmax): %.3f ms for %d samples = %d rows" % (e0.elapsed_time(e1) / 5, n, 4 * n))
