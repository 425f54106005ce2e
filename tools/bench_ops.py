#!/usr/bin/env python
"""Per-operator timings on one B200: this library vs the reference's own kernels (oracle/_ref, compiled
unmodified for sm_100a) or, where the reference calls into torch (F.grid_sample, index_add_/segment_coo,
nn.Linear), vs that torch path.  Shapes are BASELINE config B (160^3, 8192 rays, stepsize 0.5).

    python tools/bench_ops.py > gpurun_out/ops.md

CUDA events, 5 warm-up + 20 timed iterations, L2 flushed (256 MB write) before every timed iteration.
This is test/measurement infrastructure: the reference modules are only ever used here and in tests/.
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import poseprobe_b200 as pp  # noqa: E402
from oracle import build_ref  # noqa: E402
from poseprobe_b200 import dvgo, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush_buf.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3   # us


rows = []


def row(name, ours_us, ref_us, ref_what, bytes_alg=None):
    gbs = "%.0f" % (bytes_alg / ours_us / 1e3) if bytes_alg else "-"
    rows.append("| %s | %.1f | %s | %s | %s | %s |" % (name, ours_us, "%.1f" % ref_us if ref_us else "-",
                                                       "%.1fx" % (ref_us / ours_us) if ref_us else "-", ref_what, gbs))


def main():
    ref_ru = build_ref.load_ref("render_utils_cuda")
    ref_tv = build_ref.load_ref("total_variation_cuda")
    ref_ad = build_ref.load_ref("adam_upd_cuda")
    G, N = 160, 8192
    o, d, _ = synthetic.make_rays(N)
    o, d = o.to(dev), d.to(dev)
    mn, mx = torch.tensor([-1.0, -1, -1], device=dev), torch.tensor([1.0, 1, 1], device=dev)
    stepdist = 0.5 * 2 / G
    ru = pp.render_utils_cuda
    # ---- sampling
    t_o = timeit(lambda: ru.sample_pts_on_rays(o, d, mn, mx, 0.2, 1e9, stepdist))
    t_r = timeit(lambda: ref_ru.sample_pts_on_rays(o, d, mn, mx, 0.2, 1e9, stepdist)) if ref_ru else None
    pts, mask, ray_id, step_id, *_ = ru.sample_pts_on_rays(o, d, mn, mx, 0.2, 1e9, stepdist)
    M0 = pts.shape[0]
    row("sample_pts_on_rays (N=8192 -> M=%d)" % M0, t_o, t_r, "reference kernels", 48 * N + 29 * M0)
    keep = ~mask
    pts, ray_id = pts[keep].contiguous(), ray_id[keep].contiguous()
    M = pts.shape[0]
    density, k0 = synthetic.make_dvgo_grids(G)
    density, k0 = density.to(dev), k0.to(dev)
    # ---- trilinear
    mn_c, mx_c = mn.cpu(), mx.cpu()

    def gs(grid):
        ind = ((pts.reshape(1, 1, 1, -1, 3) - mn) / (mx - mn)).flip((-1,)) * 2 - 1
        return F.grid_sample(grid, ind, mode="bilinear", align_corners=True).reshape(grid.shape[1], -1).T
    for C, grid in ((1, density), (3, k0)):
        t_o = timeit(lambda: ops.trilinear(grid, pts, mn_c, mx_c))
        t_r = timeit(lambda: gs(grid))
        row("DenseGrid trilinear fwd C=%d (M=%d)" % (C, M), t_o, t_r, "F.grid_sample (ATen)", (12 + 32 * C + 4 * C) * M)
        gr = grid.clone().requires_grad_(True)
        go = torch.randn(M, C, device=dev)

        def ours_bwd():
            gr.grad = None
            ops.trilinear(gr, pts, mn_c, mx_c).backward(go)

        def ref_bwd():
            gr.grad = None
            gs(gr).backward(go)
        t_o, t_r = timeit(ours_bwd), timeit(ref_bwd)
        row("DenseGrid trilinear fwd+bwd C=%d" % C, t_o, t_r, "F.grid_sample + autograd", (12 + 96 * C + 8 * C) * M)
    # ---- raw2alpha
    dens = ops.trilinear(density, pts, mn_c, mx_c)[:, 0].contiguous()
    shift = synthetic.act_shift()
    t_o = timeit(lambda: ru.raw2alpha(dens, shift, 0.5))
    t_r = timeit(lambda: ref_ru.raw2alpha(dens, shift, 0.5)) if ref_ru else None
    row("raw2alpha fwd (M=%d)" % M, t_o, t_r, "reference kernel", 12 * M)
    exp_d, alpha = ru.raw2alpha(dens, shift, 0.5)
    gb = torch.randn(M, device=dev)
    t_o = timeit(lambda: ru.raw2alpha_backward(exp_d, gb, 0.5))
    t_r = timeit(lambda: ref_ru.raw2alpha_backward(exp_d, gb, 0.5)) if ref_ru else None
    row("raw2alpha bwd", t_o, t_r, "reference kernel", 12 * M)
    # ---- alpha2weight
    t_o = timeit(lambda: ru.alpha2weight(alpha, ray_id, N))
    t_r = timeit(lambda: ref_ru.alpha2weight(alpha, ray_id, N)) if ref_ru else None
    row("alpha2weight fwd (8192 rays, M=%d)" % M, t_o, t_r, "reference kernel", 20 * M)
    w, T, last, i_s, i_e = ru.alpha2weight(alpha, ray_id, N)
    gw, gl = torch.randn(M, device=dev), torch.randn(N, device=dev)
    t_o = timeit(lambda: ru.alpha2weight_backward(alpha, w, T, last, i_s, i_e, N, gw, gl))
    t_r = timeit(lambda: ref_ru.alpha2weight_backward(alpha, w, T, last, i_s, i_e, N, gw, gl)) if ref_ru else None
    row("alpha2weight bwd", t_o, t_r, "reference kernel", 24 * M)
    # ---- segment sum
    rgb = torch.rand(M, 3, device=dev)
    t_o = timeit(lambda: ops.segment_sum(rgb, ray_id, N))
    t_r = timeit(lambda: torch.zeros(N, 3, device=dev).index_add_(0, ray_id, rgb))
    row("segment_sum [M,3]->[N,3]", t_o, t_r, "index_add_ (segment_coo stand-in)", 20 * M)
    # ---- dense ops
    param = torch.randn(1, 4, G, G, G, device=dev)
    grad = torch.randn(1, 4, G, G, G, device=dev)
    grad[torch.rand_like(grad) < 0.2] = 0
    V = param.numel()
    t_o = timeit(lambda: pp.total_variation_cuda.total_variation_add_grad(param, grad, 1e-3, 1e-3, 1e-3, True))
    t_r = timeit(lambda: ref_tv.total_variation_add_grad(param, grad, 1e-3, 1e-3, 1e-3, True)) if ref_tv else None
    row("total_variation_add_grad dense (4x160^3)", t_o, t_r, "reference kernel", 12 * V)
    m_, v_ = torch.zeros_like(param), torch.zeros_like(param)
    t_o = timeit(lambda: pp.adam_upd_cuda.masked_adam_upd(param, grad, m_, v_, 3, 0.9, 0.99, 0.1, 1e-8))
    t_r = timeit(lambda: ref_ad.masked_adam_upd(param, grad, m_, v_, 3, 0.9, 0.99, 0.1, 1e-8)) if ref_ad else None
    row("masked_adam_upd (4x160^3)", t_o, t_r, "reference kernel", 28 * V)
    # ---- MLP
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(),
                        nn.Linear(128, 3)).to(dev)
    Mc = M
    feat = torch.randn(Mc, 57, device=dev)
    ws, bs = ops.rgbnet_from_sequential(net)
    with torch.no_grad():
        t_o = timeit(lambda: ops.rgbnet_forward(feat, ws, bs))
        t_r = timeit(lambda: torch.sigmoid(net(feat)))
        flops = 2 * (57 * 128 + 128 * 128 * 2 + 128 * 3) * Mc
        row("rgbnet fwd 57-128-128-128-3 (M=%d), %.0f TFLOP/s" % (Mc, flops / t_o / 1e6), t_o, t_r,
            "nn.Sequential fp32 (cuBLAS)", (57 * 4 + 12) * Mc)
        net16 = net.to(torch.bfloat16)
        f16 = feat.to(torch.bfloat16)
        t_r2 = timeit(lambda: torch.sigmoid(net16(f16)))
        row("rgbnet fwd (same)", t_o, t_r2, "nn.Sequential bf16 (cuBLAS)", (57 * 4 + 12) * Mc)
    feat_g = feat.clone().requires_grad_(True)
    gout = torch.randn(Mc, 3, device=dev) / Mc
    wsg = [w.detach().clone().requires_grad_(True) for w in ws]
    bsg = [b.detach().clone().requires_grad_(True) for b in bs]

    def ours_fb():
        for t in wsg + bsg + [feat_g]:
            t.grad = None
        ops.rgbnet(feat_g, wsg, bsg).backward(gout)
    net32 = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(),
                          nn.Linear(128, 3)).to(dev)

    def ref_fb(nt, x_, g_):
        nt.zero_grad(set_to_none=True)
        x_.grad = None
        torch.sigmoid(nt(x_)).backward(g_)
    t_o = timeit(ours_fb)
    t_r = timeit(lambda: ref_fb(net32, feat_g, gout))
    row("rgbnet fwd+bwd (dX, dW, db), fp32 [M,57] in/out", t_o, t_r, "nn.Sequential fp32 autograd (cuBLAS)",
        (57 * 4 * 2 + 24) * Mc)
    f16g = feat.to(torch.bfloat16).requires_grad_(True)
    g16 = gout.to(torch.bfloat16)
    t_r2 = timeit(lambda: ref_fb(net16, f16g, g16))
    row("rgbnet fwd+bwd (same)", t_o, t_r2, "nn.Sequential bf16 autograd (cuBLAS)", (57 * 4 * 2 + 24) * Mc)
    # ---- fused render vs the composed operator chain
    sc = dvgo.make_scene((G, G, G), [-1, -1, -1], [1, 1, 1], 0.2, 1e9, stepdist, shift, 0.5)
    vox = dvgo.pack_voxels(density, k0)
    t_o = timeit(lambda: dvgo.render_forward(sc, vox, o, d))

    def chain(mod):
        p_, m_k, rid, sid, *_ = mod.sample_pts_on_rays(o, d, mn, mx, 0.2, 1e9, stepdist)
        kk = ~m_k
        p_, rid = p_[kk], rid[kk]
        ind = ((p_.reshape(1, 1, 1, -1, 3) - mn) / (mx - mn)).flip((-1,)) * 2 - 1
        dn = F.grid_sample(density, ind, mode="bilinear", align_corners=True).reshape(-1)
        _, al = mod.raw2alpha(dn, shift, 0.5)
        ww, _, la, _, _ = mod.alpha2weight(al, rid, N)
        col = torch.sigmoid(F.grid_sample(k0, ind, mode="bilinear", align_corners=True).reshape(3, -1).T)
        return torch.zeros(N, 3, device=dev).index_add_(0, rid, ww[:, None] * col) + la[:, None] * 0.0
    if ref_ru:
        t_r = timeit(lambda: chain(ref_ru))
        row("render forward, 8192 rays (fused march kernel)", t_o, t_r, "reference ops + ATen chain",
            48 * N + 32 * 4 * M)
    # ---- whole train step: reference operator chain + autograd + reference masked Adam vs the fused step
    if ref_ru and ref_ad:
        class RefR2A(torch.autograd.Function):
            @staticmethod
            def forward(ctx, dn_, sh_, iv_):
                e_, a_ = ref_ru.raw2alpha(dn_, sh_, iv_)
                ctx.save_for_backward(e_)
                ctx.iv = iv_
                return a_

            @staticmethod
            def backward(ctx, g_):
                return ref_ru.raw2alpha_backward(ctx.saved_tensors[0], g_.contiguous(), ctx.iv), None, None

        class RefA2W(torch.autograd.Function):
            @staticmethod
            def forward(ctx, al_, rid_, n_):
                w_, T_, la_, is_, ie_ = ref_ru.alpha2weight(al_, rid_, n_)
                ctx.save_for_backward(al_, w_, T_, la_, is_, ie_)
                ctx.n = n_
                return w_, la_

            @staticmethod
            def backward(ctx, gw_, gl_):
                al_, w_, T_, la_, is_, ie_ = ctx.saved_tensors
                return ref_ru.alpha2weight_backward(al_, w_, T_, la_, is_, ie_, ctx.n, gw_.contiguous(),
                                                    gl_.contiguous()), None, None
        dens_p = density.clone().requires_grad_(True)
        k0_p = k0.clone().requires_grad_(True)
        st = [torch.zeros_like(dens_p), torch.zeros_like(dens_p), torch.zeros_like(k0_p), torch.zeros_like(k0_p)]
        tgt = torch.rand(N, 3, device=dev)
        step_no = [0]

        def ref_step():
            dens_p.grad = None
            k0_p.grad = None
            p_, m_k, rid, sid, *_ = ref_ru.sample_pts_on_rays(o, d, mn, mx, 0.2, 1e9, stepdist)
            kk = ~m_k
            p_, rid = p_[kk], rid[kk]
            ind = ((p_.reshape(1, 1, 1, -1, 3) - mn) / (mx - mn)).flip((-1,)) * 2 - 1
            dn = F.grid_sample(dens_p, ind, mode="bilinear", align_corners=True).reshape(-1)
            al = RefR2A.apply(dn, shift, 0.5)
            ww, la = RefA2W.apply(al, rid, N)
            col = torch.sigmoid(F.grid_sample(k0_p, ind, mode="bilinear", align_corners=True).reshape(3, -1).T)
            rgbm = torch.zeros(N, 3, device=dev).index_add_(0, rid, ww[:, None] * col) + la[:, None] * 0.0
            F.mse_loss(rgbm, tgt).backward()
            step_no[0] += 1
            with torch.no_grad():
                ref_ad.masked_adam_upd(dens_p, dens_p.grad, st[0], st[1], step_no[0], 0.9, 0.99, 0.1, 1e-8)
                ref_ad.masked_adam_upd(k0_p, k0_p.grad, st[2], st[3], step_no[0], 0.9, 0.99, 0.1, 1e-8)
        eng = dvgo.FusedDVGO(density, k0, [-1, -1, -1], [1, 1, 1], stepsize=0.5)
        t_o = timeit(lambda: eng.train_step(o, d, tgt))
        t_r = timeit(ref_step, 5)
        row("TRAIN STEP 8192 rays: fwd + bwd + masked Adam (fused: 2 launches)", t_o, t_r,
            "reference kernels + ATen grid_sample/index_add_ + autograd + reference masked_adam_upd",
            60 * N + 96 * 4 * M + 32 * 4 * G ** 3)
    print("# Per-operator timings, B200, config-B shapes (us per call; GB/s = algorithmic bytes / our time)\n")
    print("| operator | ours (us) | reference (us) | speed-up | reference is | ours GB/s |")
    print("|---|---|---|---|---|---|")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
