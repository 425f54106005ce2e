#!/usr/bin/env python
"""Times the rgbnet kernels (tcgen05) at config-C size against torch (cuBLAS) fwd+bwd."""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poseprobe_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n, d0, depth = int(os.environ.get("N", 2733151)), 57, 4
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(d0, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(),
                        nn.Linear(128, 3)).cuda()
    lin = [m for m in net if isinstance(m, nn.Linear)]
    x = torch.randn(n, d0, device="cuda")
    gout = torch.randn(n, 3, device="cuda") / n
    ws = [m.weight.detach().clone().requires_grad_(True) for m in lin]
    bs = [m.bias.detach().clone().requires_grad_(True) for m in lin]
    flops_fwd = 2.0 * n * (d0 * 128 + 128 * 128 * 2 + 128 * 3)

    t_fwd = timeit(lambda: ops.rgbnet_forward(x, ws, bs))
    xg = x.clone().requires_grad_(True)

    def ours():
        for t in ws + bs + [xg]:
            t.grad = None
        ops.rgbnet(xg, ws, bs).backward(gout)
    t_ours = timeit(ours)

    xr = x.clone().requires_grad_(True)

    def ref32():
        net.zero_grad(set_to_none=True)
        xr.grad = None
        torch.sigmoid(net(xr)).backward(gout)
    t_ref32 = timeit(ref32, 3)
    net16 = nn.Sequential(*[m for m in net]).to(torch.bfloat16)
    x16 = x.to(torch.bfloat16).requires_grad_(True)
    g16 = gout.to(torch.bfloat16)

    def ref16():
        net16.zero_grad(set_to_none=True)
        x16.grad = None
        torch.sigmoid(net16(x16)).backward(g16)
    t_ref16 = timeit(ref16, 3)
    # the fused kernel alone, on its own formats (bf16 [n,64] in, fp32 [n,64] input gradient out)
    import ctypes
    from poseprobe_b200._lib import check, lib, ptr, stream
    k0p = int(lib.pp_rgbnet_padded_width(d0))
    xp = torch.empty(n, k0p, dtype=torch.bfloat16, device="cuda")
    check(lib.pp_rgbnet_pack_input(ptr(x), n, d0, ptr(xp), stream()), "pack")
    gxp = torch.empty(n, k0p, dtype=torch.float32, device="cuda")
    gws = [torch.zeros_like(w) for w in ws]
    gbs = [torch.zeros_like(b) for b in bs]
    arr = lambda ts: (ctypes.c_void_p * depth)(*[t.data_ptr() for t in ts])
    wsd, bsd = [w.detach() for w in ws], [b.detach() for b in bs]

    def packed():
        check(lib.pp_rgbnet_backward_packed(ptr(xp), ptr(gout), n, d0, depth, 128, arr(wsd), arr(bsd), 1, ptr(gxp),
                                            arr(gws), arr(gbs), None, stream()), "bwd")
    t_packed = timeit(packed)

    def packed_nodx():
        check(lib.pp_rgbnet_backward_packed(ptr(xp), ptr(gout), n, d0, depth, 128, arr(wsd), arr(bsd), 1, None,
                                            arr(gws), arr(gbs), None, stream()), "bwd")
    t_packed_nodx = timeit(packed_nodx)
    if os.environ.get("PP_MLP_PROF"):
        packed()
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 24)()
        if lib.pp_rgbnet_backward_profile(buf) == 0:
            v = list(buf)
            tiles = ((n + 127) // 128 + 147) // 148
            print("  prof (cycles per tile, CTA 0, %d tiles): mma warp: wait epi0 %d, epi1 %d, x %d, total %d | "
                  "epi warp0: wait acc %d, dw-in-tile %d, dw-end %d, cp.async %d, total %d | epi warp8: wait acc %d, "
                  "dw-in-tile %d, dw-end %d, total %d" % (tiles, v[0] // tiles, v[1] // tiles, v[2] // tiles, v[6] // tiles,
                                                          v[8] // tiles, v[9] // tiles, v[10] // tiles, v[11] // tiles,
                                                          v[14] // tiles, v[16] // tiles, v[17] // tiles, v[18] // tiles,
                                                          v[22] // tiles))
    outp = torch.empty(n, 3, device="cuda")

    def fwd_packed():
        check(lib.pp_rgbnet_forward_packed(ptr(xp), n, d0, depth, 128, arr(wsd), arr(bsd), 1, ptr(outp), None, None,
                                           None, None, stream()), "fwd")
    t_fwdp = timeit(fwd_packed)
    print("rgbnet 57-128-128-128-3, n=%d" % n)
    print("  forward kernel alone (packed bf16 in): %.3f ms  (%.0f TFLOP/s)" % (t_fwdp, flops_fwd / t_fwdp / 1e9))
    print("  fused fwd+bwd kernel alone (packed formats): %.3f ms  (%.0f TFLOP/s incl. recompute, %.2f us/tile/CTA); "
          "without dX: %.3f ms" % (t_packed, 4 * flops_fwd / t_packed / 1e9,
                                   1e3 * t_packed / ((n + 127) // 128 / 148.0), t_packed_nodx))
    print("  ours fwd only        : %.3f ms  (%.0f TFLOP/s)" % (t_fwd, flops_fwd / t_fwd / 1e9))
    print("  ours fwd+bwd (fused) : %.3f ms  (%.0f TFLOP/s algorithmic 3x fwd; %.0f incl. recompute)"
          % (t_ours, 3 * flops_fwd / t_ours / 1e9, 4 * flops_fwd / (t_ours - 0) / 1e9))
    print("  torch fp32 fwd+bwd   : %.3f ms" % t_ref32)
    print("  torch bf16 fwd+bwd   : %.3f ms" % t_ref16)


if __name__ == "__main__":
    main()
