#!/usr/bin/env python
"""Scene-branch compositing kernels (csrc/scene.cu) and the dense DVGO compositing op on the shapes the reference
trains with, timed against the reference's own formulation (oracle/bgnerf_cpu.py = NeRF.composite /
sample_depth_from_pdf restated; oracle.get_ray_marching_ray) executed by torch ON THE SAME GPU:
    python tools/profile_scene.py            # CUDA-graph timing, cold L2
Measurement infrastructure: the oracle is only the comparison arm here."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bgnerf_cpu, oracle  # noqa: E402
from poseprobe_b200 import ops, scene  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timed(fn, n=10, graph=True):
    def run_ms(body):
        body()
        torch.cuda.synchronize()
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            call = g.replay
        else:
            call = body
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            call()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def both():
        flush.fill_(1.0)
        fn()
    return (run_ms(both) - run_ms(lambda: flush.fill_(1.0))) * 1e3


def main():
    opt = types.SimpleNamespace(nerf=types.SimpleNamespace(setbg_opaque=False), mask_img=False)
    rows = []
    for B, R, N in ((1, 2048, 128), (1, 2048, 256), (1, 65536, 256)):
        gen = torch.Generator(device=dev).manual_seed(1)
        ray = torch.randn(B, R, 3, device=dev, generator=gen).requires_grad_(True)
        depth = scene.sample_depth(N, [0.5, 6.0], B, R, dev, stratified=True,
                                   rand=torch.rand(B, R, N, 1, device=dev, generator=gen))
        rgb_s = torch.rand(B, R, N, 3, device=dev, generator=gen).requires_grad_(True)
        dens = (torch.nn.functional.softplus(torch.randn(B, R, N, device=dev, generator=gen) * 2 - 3)).requires_grad_(True)
        g_rgb = torch.randn(B, R, 3, device=dev, generator=gen)

        def ours():
            out = scene.composite(opt, ray, {"rgb_samples": rgb_s, "density_samples": dens}, depth)
            return torch.autograd.grad((out["rgb"] * g_rgb).sum() + out["depth"].sum(), [rgb_s, dens, ray])

        def ref():
            out = bgnerf_cpu.composite(ray, rgb_s, dens, depth)
            return torch.autograd.grad((out["rgb"] * g_rgb).sum() + out["depth"].sum(), [rgb_s, dens, ray])

        def ours_fwd():
            with torch.no_grad():
                return scene.composite(opt, ray, {"rgb_samples": rgb_s, "density_samples": dens}, depth)

        def ref_fwd():
            with torch.no_grad():
                return bgnerf_cpu.composite(ray, rgb_s, dens, depth)
        w = ours_fwd()["weights"][..., 0].contiguous()
        t = [timed(ours_fwd), timed(ref_fwd), timed(ours, graph=False), timed(ref, graph=False),
             timed(lambda: scene.sample_depth_from_pdf(w, N, N, [0.5, 6.0], det=True)),
             timed(lambda: bgnerf_cpu.sample_depth_from_pdf(w, N, N, [0.5, 6.0], torch.linspace(0, 1, N + 1, device=dev)))]
        M = B * R * N
        rows.append("| %d x %d | %.1f | %.1f | %.1fx | %.0f | %.1f | %.1f | %.1fx | %.1f | %.1f | %.1fx |" % (
            R, N, t[0], t[1], t[1] / t[0], 24.0 * M / t[0] / 1e3, t[2], t[3], t[3] / t[2], t[4], t[5], t[5] / t[4]))
    print("# Scene-branch compositing, B200 (us; reference = the reference's torch formulation on the same GPU)\n")
    print("| rays x samples | composite fwd | reference | speed-up | fwd GB/s (24 B/sample) | fwd+bwd (autograd, eager) | "
          "reference | speed-up | sample_depth_from_pdf | reference | speed-up |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))
    # dense DVGO compositing (config-B dense shape: 8192 rays x 558 samples)
    alpha = (torch.rand(8192, 558, device=dev) ** 4).requires_grad_(True)
    gw = torch.randn(8192, 558, device=dev)

    def d_ours():
        w_, c_ = ops.get_ray_marching_ray(alpha)
        return torch.autograd.grad((w_ * gw).sum() + c_[:, -1].sum(), alpha)

    def d_ref():
        w_, c_ = oracle.get_ray_marching_ray(alpha)
        return torch.autograd.grad((w_ * gw).sum() + c_[:, -1].sum(), alpha)
    print("\nget_ray_marching_ray fwd+bwd, 8192 x 558 dense samples: ours %.1f us, torch cumprod formulation %.1f us"
          % (timed(d_ours, graph=False), timed(d_ref, graph=False)))


if __name__ == "__main__":
    main()
