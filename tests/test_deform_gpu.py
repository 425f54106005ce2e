"""Implicit-deformation render path (forward): the deformation network with its input Jacobian
on tcgen05 (mlp.cu, kJac) and the warped SDF lookup of Voxurf.inference (voxurf_fused.cu, kDeform) against

* plain PyTorch: the same network evaluated in fp32 with torch.autograd for the Jacobian, and an emulation of the
  kernel's arithmetic (bf16 operands, fp32 accumulation, tangents gated by the value row) for a tight check;
* oracle/voxurf_cpu.forward(warp=...), the restatement of the reference's deform branch that
  tests/test_oracle_golden.py pins against the reference's own code (tests/golden/voxurf_deform.npz).

Tolerances: the warp network runs on fp16 operands (2^-12 relative rounding per operand, accumulation in the tensor
core's fp32 datapath): 2e-3 of the output scale against fp32, 1e-3 against the fp16 emulation (measured 3e-4: the
MMA's internal accumulation is not a sequence of IEEE fp32 adds).  Its Jacobian is compared per sample because a ReLU
gate that flips under rounding changes one sample's Jacobian by a rank-1 term: >= 90 % of the samples within 5e-3
relative (measured 94-97 %, the same fraction the fp16 emulation shows against fp32), median < 2e-3 (fp32 autograd) /
1e-3 (emulation).  Per-ray render outputs downstream of the bf16 colour network: 2e-2 absolute.
"""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

MN, MX = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]


def make_warp(seed, last_scale=0.005, depth=5):
    torch.manual_seed(seed)
    dims = [3] + [128] * (depth - 1) + [4]
    ws, bs = [], []
    for i in range(depth):
        w = torch.empty(dims[i + 1], dims[i])
        nn.init.kaiming_normal_(w, a=0.0, nonlinearity="relu", mode="fan_in")       # modules.py:127-130
        ws.append(w)
        bs.append(0.1 * torch.randn(dims[i + 1]))
    ws[-1] = last_scale * torch.randn(4, 128)
    bs[-1] = 0.01 * torch.randn(4)
    return ws, bs


def net_fp32(x, ws, bs):
    h = x
    for i, (w, b) in enumerate(zip(ws, bs)):
        h = torch.nn.functional.linear(h, w, b)
        if i < len(ws) - 1:
            h = torch.relu(h)
    return h


def bf(t):
    return t.to(torch.float16).to(torch.float32)      # the warp network's operand format


def net_emulated(x, ws, bs):
    """the kernel's arithmetic: fp16 operands, fp32 accumulation, tangents gated by the value row's pattern"""
    hi = bf(x)
    lo = bf(x - hi)
    w0 = bf(ws[0])
    z = hi @ w0.T + lo @ w0.T + bs[0]
    gate = (z > 0).float()
    h = bf(torch.relu(z))
    t = bf(gate[:, None, :] * w0.T[None])                  # [n, 3, 128]: tangent k = column k of W_0
    for w, b in zip(ws[1:-1], bs[1:-1]):
        wb = bf(w)
        z = h @ wb.T + b
        gate = (z > 0).float()
        t = bf(gate[:, None, :] * (t @ wb.T))
        h = bf(torch.relu(z))
    wl = bf(ws[-1])
    return h @ wl.T + bs[-1], (t @ wl.T).transpose(1, 2)   # [n,4], [n,4,3]


@pytest.mark.parametrize("n,depth", [(1000, 5), (37, 3), (4096, 5)])
def test_warpnet_forward_with_jacobian(cuda, n, depth):
    from poseprobe_b200 import ops
    ws, bs = make_warp(3 + depth, last_scale=0.05, depth=depth)
    torch.manual_seed(n)
    x = torch.rand(n, 3) * 2 - 1
    R = 2.0
    d, c, jac = ops.warpnet_forward(x.to(cuda), [w.to(cuda) for w in ws], [b.to(cuda) for b in bs], out_range=R)
    assert d.shape == (n, 3) and c.shape == (n, 1) and jac.shape == (n, 4, 3)
    got = torch.cat([d, c], 1).cpu()
    # fp32 reference with autograd Jacobian (what the reference computes, voxurf_coarse.py:966-985)
    xg = x.clone().requires_grad_(True)
    y = net_fp32(xg, ws, bs) * R
    want_jac = torch.stack([torch.autograd.grad(y[:, k].sum(), xg, retain_graph=True)[0] for k in range(4)], 1)
    scale = float(y.detach().abs().max())
    assert float((got - y.detach()).abs().max()) < 2e-3 * scale

    def per_sample_rel(a, b):
        return (a - b).flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1).clamp_min(1e-12)

    # a gate that flips because z was rounded changes that sample's Jacobian by a rank-1 term (the Jacobian of a
    # ReLU network is discontinuous there): bounded fraction of such samples, the rest agrees to operand rounding
    rel = per_sample_rel(jac.cpu(), want_jac)
    assert float((rel < 5e-3).float().mean()) > 0.9 and float(rel.median()) < 2e-3, (
        float((rel < 5e-3).float().mean()), float(rel.median()))
    assert float((jac.cpu() - want_jac).norm() / want_jac.norm()) < 0.1
    # emulation of the kernel's arithmetic: same gates up to accumulation order
    ey, ejac = net_emulated(x, ws, bs)
    assert float((got - ey * R).abs().max()) < 1e-3 * scale, float((got - ey * R).abs().max()) / scale
    rel = per_sample_rel(jac.cpu(), ejac * R)
    assert float((rel < 5e-3).float().mean()) > 0.9 and float(rel.median()) < 1e-3, (
        float((rel < 5e-3).float().mean()), float(rel.median()))


def make_scene(G, n_rays, seed):
    from poseprobe_b200 import synthetic
    torch.manual_seed(seed)
    sdf = synthetic.make_sdf_grid(G, radius=0.5, noise=0.01)
    k0 = 0.3 * torch.randn(1, 12, G, G, G)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(inplace=True), nn.Sequential(nn.Linear(128, 128), nn.ReLU(inplace=True)),
                        nn.Linear(128, 3))
    nn.init.constant_(net[-1].bias, 0)
    with torch.no_grad():
        net[0].weight.mul_(2.0)
        net[-1].weight.mul_(4.0)
    o, d, v = synthetic.make_rays(n_rays, seed=seed + 5)
    kw = dict(stepsize=0.5, s_val=0.05, bg=0.3, posbase_pe=5, viewbase_pe=1, barf_c2f=(0.1, 0.5), sdf_alpha=10.0,
              sdf_beta=0.35)
    return sdf, k0, net, o, d, v, kw


@pytest.mark.parametrize("use_jitter", [False, True])
def test_render_deform_vs_oracle(cuda, use_jitter):
    from oracle import voxurf_cpu
    from poseprobe_b200 import ops, voxurf
    G, N, R = 32, 384, 2.0
    sdf, k0, net, o, d, v, kw = make_scene(G, N, seed=11)
    ws, bs = make_warp(21)
    rw, rb = ops.rgbnet_from_sequential(net)
    eng = voxurf.FusedVoxurf(sdf.to(cuda), k0.to(cuda), [w.detach() for w in rw], [b.detach() for b in rb], MN, MX, **kw)
    eng.set_warp_network(ws, bs, R)
    jitter = torch.rand(N) if use_jitter else None
    rgb, last, cumw, depth, normal = eng.render(o.to(cuda), d.to(cuda), v.to(cuda), want_depth=True, want_normal=True,
                                                jitter=jitter.to(cuda) if use_jitter else None)
    b = eng._bufs
    assert b["dcounter"].tolist()[1] == 0
    want = voxurf_cpu.forward(sdf, k0, net, o, d, v, MN, MX, kw["stepsize"], eng.voxel_size, kw["s_val"], bg=kw["bg"],
                              sdf_alpha=kw["sdf_alpha"], sdf_beta=kw["sdf_beta"], posbase_pe=5, viewbase_pe=1,
                              barf_c2f=kw["barf_c2f"], progress=1.0, jitter=jitter, warp=(ws, bs, R))
    want = {k: (t.detach() if torch.is_tensor(t) else t) for k, t in want.items()}
    # all in-box samples went through the network, in the reference's compacted order
    assert b["dcounter"].tolist()[0] == want["ray_id"].numel() and b["dcounter"].tolist()[2] == 4 * want["ray_id"].numel()
    # per-sample SDF / normal of the composited samples (b["rec"]) against the oracle, ray by ray
    m = eng.n_samples()
    ray_cnt, ray_off = b["ray_cnt"][:N].cpu().long(), b["ray_off"][:N].cpu().long()
    rid_ref = want["ray_id"]
    per_ray_ref = torch.zeros(N, dtype=torch.long).index_add_(0, rid_ref, torch.ones_like(rid_ref))
    starts_ref = torch.cumsum(per_ray_ref, 0) - per_ray_ref
    rec = b["rec"][:m].cpu()
    checked, grad_err = 0, []
    for r in range(0, N, 5):
        c = int(ray_cnt[r])
        if c == 0:
            continue
        sl = slice(int(ray_off[r]), int(ray_off[r]) + c)
        rs = slice(int(starts_ref[r]), int(starts_ref[r]) + c)
        torch.testing.assert_close(rec[sl, 0], want["sdf"][rs], rtol=0, atol=1e-3)
        grad_err.append((rec[sl, 1:] - want["gradient"][rs]).abs().max(-1).values)
        checked += c
    assert checked > 500
    grad_err = torch.cat(grad_err)       # gate flips (see test_warpnet_forward_with_jacobian) hit a few samples
    assert float((grad_err < 5e-3).float().mean()) > 0.95 and float(grad_err.median()) < 1e-3
    torch.testing.assert_close(cumw.cpu(), want["cum_weights"][:, 0], rtol=0, atol=2e-2)
    torch.testing.assert_close(last.cpu(), want["alphainv_cum"], rtol=0, atol=2e-2)
    torch.testing.assert_close(rgb.cpu(), want["rgb_marched"], rtol=0, atol=2e-2)
    rid, w, g = want["ray_id"], want["weights"], want["gradient"]
    dist = kw["stepsize"] * eng.voxel_size
    steps = want["step_id"].float() + (jitter[rid] if use_jitter else 0.0)
    want_depth = torch.zeros(N).index_add_(0, rid, w * steps * dist)
    want_normal = torch.zeros(N, 3).index_add_(0, rid, w[:, None] * g / (g.norm(2, -1, keepdim=True) + 1e-6))
    torch.testing.assert_close(depth.cpu(), want_depth, rtol=0, atol=3e-2)
    torch.testing.assert_close(normal.cpu(), want_normal, rtol=0, atol=3e-2)
    # the deformation is not a no-op in this scene: the plain lookup renders something else
    rgb0 = eng.render(o.to(cuda), d.to(cuda), v.to(cuda), use_deform=False,
                      jitter=jitter.to(cuda) if use_jitter else None)[0]
    assert float((rgb0 - rgb).abs().max()) > 0.05


def test_voxurf_module_inference(cuda):
    """pp.Voxurf.inference (the reference's render entry point, lib/nvs_fun.py:81-83 -> voxurf_coarse.py:1093-1215) with
    a non-trivial warp network against the oracle's deform branch."""
    import poseprobe_b200 as pp
    from oracle import voxurf_cpu
    G, N = 24, 200
    sdf, k0, net, o, d, v, kw = make_scene(G, N, seed=31)
    ws, bs = make_warp(41)
    m = pp.Voxurf(MN, MX, num_voxels=G ** 3, num_voxels_base=G ** 3, rgbnet_depth=3, barf_c2f=kw["barf_c2f"],
                  N_iters=1000).to(cuda)
    assert "warp_network.deform_net.net.net.4.0.weight" in m.state_dict()
    with torch.no_grad():
        m.sdf.grid.copy_(sdf)
        m.k0.grid.copy_(k0)
        for dst, src in zip([q for q in m.rgbnet.modules() if isinstance(q, nn.Linear)],
                            [q for q in net.modules() if isinstance(q, nn.Linear)]):
            dst.weight.copy_(src.weight)
            dst.bias.copy_(src.bias)
        for lin, w, b in zip(m.warp_network.linears(), ws, bs):
            lin.weight.copy_(w)
            lin.bias.copy_(b)
        m.sdf_alpha.fill_(kw["sdf_alpha"])
        m.sdf_beta.fill_(kw["sdf_beta"])
        m.s_val.fill_(kw["s_val"])
    out = m.inference(o.to(cuda), d.to(cuda), v.to(cuda), near=0.2, far=1e9, stepsize=0.5, bg=kw["bg"])
    with torch.no_grad():       # forward() defaults to the deformation branch like the reference (:922)
        out2 = m(o.to(cuda), d.to(cuda), v.to(cuda), near=0.2, far=1e9, stepsize=0.5, bg=kw["bg"])
    # (same kernels; the colour sums are float atomics, so equal up to summation order)
    torch.testing.assert_close(out2["rgb_marched"], out["rgb_marched"], rtol=0, atol=1e-5)
    R = m.warp_network.output_range
    assert R == 2.0
    want = voxurf_cpu.forward(sdf, k0, net, o, d, v, MN, MX, 0.5, float(m.voxel_size), kw["s_val"], bg=kw["bg"],
                              sdf_alpha=kw["sdf_alpha"], sdf_beta=kw["sdf_beta"], posbase_pe=5, viewbase_pe=1,
                              barf_c2f=kw["barf_c2f"], progress=1.0, warp=(ws, bs, R))
    torch.testing.assert_close(out["rgb_marched"].cpu(), want["rgb_marched"].detach(), rtol=0, atol=2e-2)
    torch.testing.assert_close(out["cum_weights"].cpu(), want["cum_weights"].detach(), rtol=0, atol=2e-2)
    torch.testing.assert_close(out["alphainv_cum"].cpu(), want["alphainv_cum"].detach(), rtol=0, atol=2e-2)
    assert out["depth"].shape == (N,) and out["normal_marched"].shape == (N, 3)
    # the module's own network call: deformation and correction of free-standing points
    x = (torch.rand(300, 3) * 2 - 1)
    dfm, cor = m.warp_network(x.to(cuda))
    y = net_fp32(x, ws, bs) * R
    assert float((torch.cat([dfm, cor], 1).cpu() - y).abs().max()) < 2e-3 * float(y.abs().max())


def test_render_deform_edge_cases(cuda):
    """Non-cubic grid and box, rays that miss the box mixed in (no rows reserved for them), a ray count that is not a
    multiple of the warps per block, and a sample capacity that is too small (flagged, nothing overrun)."""
    from oracle import voxurf_cpu
    from poseprobe_b200 import ops, synthetic, voxurf
    torch.manual_seed(60)
    X, Y, Z = 30, 22, 36
    mn, mx = [-0.4, -0.35, -0.5], [0.45, 0.35, 0.6]
    lx, ly, lz = [torch.linspace(a, b, n) for a, b, n in zip(mn, mx, (X, Y, Z))]
    gx, gy, gz = torch.meshgrid(lx, ly, lz, indexing="ij")
    sdf = ((gx * gx + gy * gy + gz * gz).sqrt() - 0.25 + 0.005 * torch.randn(X, Y, Z))[None, None].contiguous()
    k0 = 0.3 * torch.randn(1, 12, X, Y, Z)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 3))
    with torch.no_grad():
        net[0].weight.mul_(2.0)
        net[-1].weight.mul_(4.0)
    N = 203
    o, d, v = synthetic.make_rays(N, seed=61, radius=1.5, target_half=0.3)
    o[::9] = torch.tensor([5.0, 5.0, 5.0])           # these miss the box
    d[::9] = torch.tensor([1.0, 0.0, 0.0])
    v[::9] = torch.tensor([1.0, 0.0, 0.0])
    ws, bs = make_warp(62, last_scale=0.002)
    R = 1.1
    rw, rb = ops.rgbnet_from_sequential(net)
    kw = dict(stepsize=0.5, s_val=0.03, bg=0.2, posbase_pe=5, viewbase_pe=1, barf_c2f=None, sdf_alpha=10.0,
              sdf_beta=0.35)
    eng = voxurf.FusedVoxurf(sdf.to(cuda), k0.to(cuda), [w.detach() for w in rw], [b.detach() for b in rb], mn, mx, **kw)
    eng.set_warp_network(ws, bs, R)
    rgb, last, cumw = eng.render(o.to(cuda), d.to(cuda), v.to(cuda))
    want = voxurf_cpu.forward(sdf, k0, net, o, d, v, mn, mx, 0.5, eng.voxel_size, kw["s_val"], bg=kw["bg"],
                              sdf_alpha=10.0, sdf_beta=0.35, posbase_pe=5, viewbase_pe=1, barf_c2f=None,
                              warp=(ws, bs, R))
    assert eng._bufs["dcounter"].tolist()[0] == want["ray_id"].numel()
    assert int((eng._bufs["dray_off"][:N][::9] >= 0).sum()) == len(range(0, N, 9))    # empty runs are legal, not -1
    torch.testing.assert_close(rgb.cpu(), want["rgb_marched"].detach(), rtol=0, atol=2e-2)
    torch.testing.assert_close(cumw.cpu(), want["cum_weights"].detach()[:, 0], rtol=0, atol=2e-2)
    torch.testing.assert_close(rgb.cpu()[::9], torch.full((len(range(0, N, 9)), 3), kw["bg"]))
    assert torch.equal(last.cpu()[::9], torch.ones(len(range(0, N, 9))))
    small = voxurf.FusedVoxurf(sdf.to(cuda), k0.to(cuda), [w.detach() for w in rw], [b.detach() for b in rb], mn, mx,
                               samples_per_ray_cap=3, **kw)
    small.set_warp_network(ws, bs, R)
    small.render(o.to(cuda), d.to(cuda), v.to(cuda))
    torch.cuda.synchronize()
    assert small._bufs["dcounter"].tolist()[1] == 1
    with pytest.raises(RuntimeError, match="capacity"):
        small.n_samples()


def test_deform_backward_pieces(cuda):
    """Training through the deformation, stage by stage (the network's own backward: test_warpnet_backward):
    (a) the reverse march's per-sample {dL/d sdf_final, dL/d gradient} against autograd through the oracle's deform
        branch (the rgbnet in between runs on bf16 operands: direction / norm);
    (b) the per-sample backward of the warped lookup + regularisers (voxurf_deform_sample_backward_kernel) against the
        closed form of oracle/deform_analytic.py evaluated on the kernel's OWN inputs (network outputs, positions,
        upstream gradients): tight;
    (c) the four regulariser terms of deform_implicit_loss."""
    from oracle import deform_analytic as da
    from oracle import voxurf_cpu
    from poseprobe_b200 import ops, voxurf
    G, N, R = 32, 256, 2.0
    sdf, k0, net, o, d, v, kw = make_scene(G, N, seed=71)
    target = torch.rand(N, 3)
    ws, bs = make_warp(72)
    rw, rb = ops.rgbnet_from_sequential(net)
    eng = voxurf.FusedVoxurf(sdf.to(cuda), k0.to(cuda), [w.detach() for w in rw], [b.detach() for b in rb], MN, MX, **kw)
    eng.set_warp_network(ws, bs, R)
    wreg = (1.0, 0.05, 0.07, 0.09)
    eng.loss_buf.zero_()
    eng.backward_deform_pieces(o.to(cuda), d.to(cuda), v.to(cuda), target.to(cuda), reg_weights=wreg)
    torch.cuda.synchronize()
    b = eng._bufs
    M = b["dcounter"].tolist()[0]
    assert b["dcounter"].tolist()[1] == 0 and M > 5000
    D = b["dout"][:4 * M].reshape(M, 4, 4).cpu()
    pts = b["dpts"][:M, :3].cpu()
    g_val = b["g_val"][:M].cpu()
    g_dout = b["g_dout"][:4 * M].reshape(M, 4, 4).cpu()
    g_pts = b["g_pts"][:M].cpu()

    def rel(p, q):
        return float((p - q).norm() / (q.norm() + 1e-12))

    def cos(p, q):
        return float((p * q).sum() / (p.norm() * q.norm() + 1e-30))

    # ---- (b) + (c): closed form on the kernel's own inputs
    grid = voxurf_cpu.softplus10(kw["sdf_alpha"]) * (torch.sigmoid(voxurf_cpu.softplus10(kw["sdf_beta"]) * sdf) - 0.5)
    y, Jt = D[:, 0, :], D[:, 1:, :]              # raw network outputs (not yet multiplied by the output range)
    sdf2, grad2, A, sd2, c2, cache = da.lookup_forward(grid, pts, y, Jt, R, MN, MX)
    _, G_g, G_A, G_c, G_d = da.regulariser_grads(grad2, A, c2, sd2, *wreg)
    G_y, G_Jt, G_x = da.lookup_backward(cache, g_val[:, 0], g_val[:, 1:] + G_g, G_A, G_c, G_d)
    G_x = G_x + torch.einsum("mik,mk->mi", Jt, G_y)      # + J^T G_y: through the network's value row
    want_dout = torch.cat([G_y[:, None, :], G_Jt], 1)
    assert rel(g_dout, want_dout) < 2e-3, rel(g_dout, want_dout)
    assert rel(g_dout[:, 0], want_dout[:, 0]) < 2e-3 and rel(g_dout[:, 1:], want_dout[:, 1:]) < 2e-3
    assert rel(g_pts, G_x) < 2e-3, rel(g_pts, G_x)
    terms = torch.stack([(grad2.norm(dim=-1) - 1).abs().mean(), A.norm(dim=-1).mean(), c2.abs().mean(), sd2.abs().mean()])
    torch.testing.assert_close(b["reg"].cpu(), terms, rtol=1e-3, atol=1e-6)

    # ---- (a): upstream gradients against autograd through the oracle
    want = voxurf_cpu.forward(sdf, k0, net, o, d, v, MN, MX, kw["stepsize"], eng.voxel_size, kw["s_val"], bg=kw["bg"],
                              sdf_alpha=kw["sdf_alpha"], sdf_beta=kw["sdf_beta"], posbase_pe=5, viewbase_pe=1,
                              barf_c2f=kw["barf_c2f"], progress=1.0, warp=(ws, bs, R))
    loss = voxurf_cpu.mse_loss(want, target)
    assert abs(float(eng.loss_buf[0]) - float(loss.detach())) < 3e-2 * float(loss.detach())
    Gs_ref, Gg_ref = torch.autograd.grad(loss, [want["sdf"], want["gradient"]])
    rid = want["ray_id"]
    assert rid.numel() == M
    per_ray = torch.zeros(N, dtype=torch.long).index_add_(0, rid, torch.ones_like(rid))
    starts = torch.cumsum(per_ray, 0) - per_ray
    doff = b["dray_off"][:N].cpu().long()
    idx = doff[rid] + (torch.arange(M) - starts[rid])        # kernel slot of every oracle sample
    assert int(idx.min()) >= 0 and int(idx.max()) < M and idx.unique().numel() == M
    torch.testing.assert_close(pts[idx], want["pts"].detach(), rtol=0, atol=1e-5)
    gk = g_val[idx]
    assert cos(gk[:, 0], Gs_ref) > 0.99 and 0.9 < float(gk[:, 0].norm() / Gs_ref.norm()) < 1.1, (
        cos(gk[:, 0], Gs_ref), float(gk[:, 0].norm() / Gs_ref.norm()))
    assert cos(gk[:, 1:], Gg_ref) > 0.99 and 0.9 < float(gk[:, 1:].norm() / Gg_ref.norm()) < 1.1, (
        cos(gk[:, 1:], Gg_ref), float(gk[:, 1:].norm() / Gg_ref.norm()))
    # the feature-grid gradient of the deform step (k0 is sampled at the unwarped point, voxurf_coarse.py:1007)
    k0_r = k0.clone().requires_grad_(True)
    want2 = voxurf_cpu.forward(sdf, k0_r, net, o, d, v, MN, MX, kw["stepsize"], eng.voxel_size, kw["s_val"],
                               bg=kw["bg"], sdf_alpha=kw["sdf_alpha"], sdf_beta=kw["sdf_beta"], posbase_pe=5,
                               viewbase_pe=1, barf_c2f=kw["barf_c2f"], progress=1.0, warp=(ws, bs, R))
    (g_k0_ref,) = torch.autograd.grad(voxurf_cpu.mse_loss(want2, target), [k0_r])
    g_k0 = torch.empty((1, 12, G, G, G), device=cuda)
    from poseprobe_b200._lib import check, lib, ptr, stream
    with torch.cuda.device(cuda):
        check(lib.pp_voxurf_k0_unpack(ptr(eng.g_k0rec), eng.n_voxels, ptr(g_k0), stream()), "k0_unpack")
    assert cos(g_k0.cpu(), g_k0_ref) > 0.99


def _emulated_rows_backward(x, ws, bs, G_y, G_Jt):
    """oracle/deform_analytic.py::warp_rows_backward on the kernel's numerics: fp16 operands (weights, activations,
    tangents, x = x_hi + x_lo), fp32 accumulation, gates from the value row."""
    from oracle import deform_analytic as da
    hi = bf(x)
    xs = hi + bf(x - hi)
    wq = [bf(w) for w in ws]
    h, T, acts = xs, torch.eye(3).expand(x.shape[0], 3, 3), []
    for w, b in zip(wq[:-1], bs[:-1]):
        z = h @ w.T + b
        gate = (bf(z) > 0).float()
        acts.append((h, T, gate))
        h = bf(torch.relu(z))
        T = bf(gate[:, None, :] * (T @ w.T))
    acts.append((h, T, None))
    return da.warp_rows_backward(wq, acts, G_y, G_Jt)


@pytest.mark.parametrize("n,gscale", [(4096, 1.0), (1000, 1e-7), (33, 3e3)])
def test_warpnet_backward(cuda, n, gscale):
    """pp_warpnet_backward_packed (csrc/warp_bwd.cu) against the closed-form four-row backward of
    oracle/deform_analytic.py (which tests/test_voxurf_cpu.py pins against autograd's double backward): every dW_l and
    db_l.  Tolerances: 1e-2 relative (Frobenius) against the fp16-operand emulation (measured ~2e-3: fp16 rounding of
    the dZ tiles), 8e-2 against plain fp32 (gate flips of samples within rounding of a ReLU kink, as in the forward
    test).  gscale exercises the power-of-two loss scale (gradients far below / above fp16's range)."""
    from oracle import deform_analytic as da
    from poseprobe_b200 import ops
    from poseprobe_b200._lib import lib
    ws, bs = make_warp(91, last_scale=0.05)
    torch.manual_seed(n)
    x = torch.rand(n, 3) * 2 - 1
    G = gscale * torch.randn(n, 4, 4) * torch.rand(n, 1, 1)
    G[:, 1:, :] *= 0.3
    gws, gbs = ops.warpnet_backward(x.to(cuda), [w.to(cuda) for w in ws], [b.to(cuda) for b in bs], G.to(cuda))
    torch.cuda.synchronize()
    assert lib.pp_warpnet_backward_stalled() == 0
    G_y, G_Jt = G[:, 0, :], G[:, 1:, :]
    _, _, acts = da.warp_rows_forward(x, ws, bs)
    want_w, want_b, _ = da.warp_rows_backward(ws, acts, G_y, G_Jt)
    emu_w, emu_b, _ = _emulated_rows_backward(x, ws, bs, G_y, G_Jt)

    def rel(p, q):
        return float((p - q).norm() / (q.norm() + 1e-30))

    for l in range(5):
        assert gws[l].shape == ws[l].shape and gbs[l].shape == bs[l].shape
        assert rel(gws[l].cpu(), emu_w[l]) < 1e-2, ("dW", l, rel(gws[l].cpu(), emu_w[l]))
        assert rel(gbs[l].cpu(), emu_b[l]) < 1e-2, ("db", l, rel(gbs[l].cpu(), emu_b[l]))
        if n >= 1000:      # (with a few dozen samples one flipped gate is a visible fraction of the sum)
            assert rel(gws[l].cpu(), want_w[l]) < 8e-2, ("dW fp32", l, rel(gws[l].cpu(), want_w[l]))
            assert rel(gbs[l].cpu(), want_b[l]) < 8e-2, ("db fp32", l, rel(gbs[l].cpu(), want_b[l]))


def _golden_engine(cuda, g, **extra):
    import numpy as np
    from poseprobe_b200 import voxurf
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    sdf = t("sdf")
    k0 = torch.from_numpy(g["k0"].astype(np.float32))
    rw, rb = [t("w%d" % i) for i in range(3)], [t("b%d" % i) for i in range(3)]
    ww, wb = [t("ww%d" % i) for i in range(5)], [t("wb%d" % i) for i in range(5)]
    eng = voxurf.FusedVoxurf(sdf.to(cuda), k0.to(cuda), [w.to(cuda) for w in rw], [b.to(cuda) for b in rb], MN, MX,
                             stepsize=0.5, s_val=float(g["s_val"]), bg=0.3, posbase_pe=5, viewbase_pe=1,
                             barf_c2f=(0.1, 0.5), sdf_alpha=10.0, sdf_beta=0.35, voxel_size=float(g["voxel_size"]),
                             pose_grad=True, **extra)
    eng.set_schedule(progress=float(g["gstep"]) / float(g["N_iters"]))
    eng.set_warp_network(ww, wb, float(g["out_range"]))
    return eng, t


def test_deform_training_gradients_vs_reference_golden(cuda):
    """The reference's live train step (recon_scene.py:603: Voxurf.forward(use_deform=True) -> loss -> backward) on the
    fused kernels against tests/golden/voxurf_deform.npz = the reference's OWN forward, deform_implicit_loss and
    autograd double backward executed on the CPU (tests/golden/make_golden_deform.py): the loss and every gradient -
    deformation network (tcgen05 backward, csrc/warp_bwd.cu), rgbnet, feature grid, rays.
    Bars: cosine >= 0.999 and norm within 1 % for the feature grid, the colour network and the deformation network's
    weights; the ray gradients (sums over ~30 samples each, a few of which sit on a gate that fp16 flips) cosine >= 0.99."""
    import os
    import numpy as np
    from poseprobe_b200._lib import lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    eng, t = _golden_engine(cuda, g)
    o, d, v = t("rays_o").to(cuda), t("rays_d").to(cuda), t("viewdirs").to(cuda)
    N = o.shape[0]
    eng.loss_buf.zero_()
    eng.backward_deform(o, d, v, t("target").to(cuda), jitter=t("jitter").to(cuda), reg_weights=(1.0, 0.05, 0.05, 0.05))
    torch.cuda.synchronize()
    assert lib.pp_warpnet_backward_stalled() == 0 and lib.pp_rgbnet_backward_stalled() == 0
    b = eng._bufs
    M = b["dcounter"].tolist()[0]
    assert M == g["ret_weights"].shape[0] and b["dcounter"].tolist()[1] == 0
    reg = b["reg"].cpu()
    for i, k in enumerate(("grad_constraint", "grad_deform_constraint", "sdf_correct_constraint", "sdf_deform_constraint")):
        assert abs(float(reg[i]) - float(g["term_" + k])) < 2e-3 * max(1.0, abs(float(g["term_" + k]))), (k, float(reg[i]))
    loss = float(eng.loss_buf[0]) + float(reg[0]) + 0.05 * float(reg[1] + reg[2] + reg[3])
    assert abs(loss - float(g["loss"])) < 1e-2 * float(g["loss"]), (loss, float(g["loss"]))

    def cos(p, q):
        return float((p * q).sum() / (p.norm() * q.norm() + 1e-30))

    def check(name, got, want, c_min=0.999, n_tol=0.01):
        got, want = got.detach().cpu().reshape(-1), want.reshape(-1)
        c, r = cos(got, want), float(got.norm() / (want.norm() + 1e-30))
        assert c >= c_min and abs(r - 1) <= n_tol, (name, c, r)

    gww, gwb = eng.warp_gradients()
    for i in range(5):
        check("warp weight %d" % i, gww[i], t("g_ww%d" % i))
        check("warp bias %d" % i, gwb[i], t("g_wb%d" % i), c_min=0.998, n_tol=0.02)
    _, g_k0, g_w, g_b = eng.grid_gradients()
    for i in range(3):
        check("rgbnet weight %d" % i, g_w[i], t("g_w%d" % i))
        check("rgbnet bias %d" % i, g_b[i], t("g_b%d" % i))
    want_k0 = torch.zeros(g_k0.numel())
    want_k0[torch.from_numpy(g["g_k0_idx"]).long()] = t("g_k0_val")
    check("k0", g_k0, want_k0)
    g_rays, _ = eng.ray_gradients()
    check("rays_o", g_rays[:N, :3], t("g_rays_o"), c_min=0.99, n_tol=0.03)
    check("rays_d", g_rays[:N, 3:], t("g_rays_d"), c_min=0.99, n_tol=0.03)


def test_deform_train_step_reduces_the_live_loss(cuda):
    """FusedVoxurf.train_step(use_deform=True) with the reference's full object loss (lib/losses.py:34-74: photometric
    with mask, entropy, mask BCE, k0 TV, Eikonal over all in-box samples, the three deformation priors) and its
    optimiser (dense Python Adam): the loss goes down, the SDF grid stays frozen, the deformation network, the feature
    grid and the rgbnet move."""
    import os
    import numpy as np
    from poseprobe_b200._lib import lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    eng, t = _golden_engine(cuda, g, w_rgb=1.0, w_entropy=0.01, w_mask=0.1, w_tv_k0=0.01, w_eikonal=1.0, w_deform=0.05,
                            lr_k0=0.05, lr_rgbnet=1e-3, lr_warp=1e-3)
    o, d, v = t("rays_o").to(cuda), t("rays_d").to(cuda), t("viewdirs").to(cuda)
    target = t("target").to(cuda)
    mask = (torch.rand(o.shape[0], device=cuda) > 0.3).float()
    sdf0, k0_0 = eng.sdf.clone(), eng.k0rec.clone()
    w0 = [w.clone() for w in eng.warp_w]
    r0 = eng.weights[1].clone()
    losses = []
    for _ in range(12):
        lb = eng.train_step(o, d, v, target, use_deform=True, mask=mask)
        reg = eng._bufs["reg"]
        losses.append(float(lb[0]) + float(reg[0]) + 0.05 * float(reg[1] + reg[2] + reg[3]))
    torch.cuda.synchronize()
    assert lib.pp_warpnet_backward_stalled() == 0
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], losses
    assert torch.equal(eng.sdf, sdf0)                                     # frozen (voxurf_coarse.py:136-138)
    assert float((eng.k0rec - k0_0).abs().max()) > 0 and float((eng.weights[1] - r0).abs().max()) > 0
    assert all(float((a - b).abs().max()) > 0 for a, b in zip(eng.warp_w, w0))
    assert float(lb[4]) > 0                                               # the k0 TV term was evaluated
    # the kernel-side copy of the doubled first layer follows the master weights
    torch.testing.assert_close(eng.warp[0][0], torch.cat([eng.warp_w[0], eng.warp_w[0]], 1))


# ---- the nn.Module against the reference's own code (golden replays) ---------------------------------------------
def _golden_module(cuda, g, n_iters):
    import poseprobe_b200 as pp
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    G = g["sdf"].shape[-1]
    m = pp.Voxurf(MN, MX, num_voxels=G ** 3, num_voxels_base=G ** 3, rgbnet_depth=3, barf_c2f=[0.1, 0.5],
                  N_iters=n_iters, s_ratio=50, s_start=0.2).to(cuda)
    with torch.no_grad():
        m.sdf.grid.copy_(t("sdf"))
        m.k0.grid.copy_(torch.from_numpy(g["k0"].astype("float32")))
        lin = [q for q in m.rgbnet.modules() if isinstance(q, nn.Linear)]
        for i, q in enumerate(lin):
            q.weight.copy_(t("w%d" % i))
            q.bias.copy_(t("b%d" % i))
        m.sdf_alpha.fill_(10.0)
        m.sdf_beta.fill_(0.35)
    assert abs(float(m.voxel_size) - float(g["voxel_size"])) < 1e-6
    return m, t


@pytest.mark.parametrize("tag", ["eval", "train"])
def test_voxurf_module_replays_the_references_forward(cuda, tag):
    """pp.Voxurf.forward(use_deform=False) against tests/golden/voxurf_forward.npz = the reference's OWN
    Voxurf.forward executed on the CPU (tests/golden/make_golden_voxurf.py), eval and train mode (jitter, BARF progress,
    s_val schedule): every entry of the result dict the callers read (voxurf_coarse.py:1069-1091; lib/losses.py:39-66,
    lib/recon_scene.py:725-736).  Per-ray outputs pass the bf16 colour network: 2e-2; geometry-only entries 2e-4."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_forward.npz"))
    m, t = _golden_module(cuda, g, int(g["N_iters"]))
    o, d, v = t("rays_o").to(cuda), t("rays_d").to(cuda), t("viewdirs").to(cuda)
    kw = dict(near=0.2, far=1e9, stepsize=0.5, bg=0.3)
    if tag == "train":
        out = m(o, d, v, use_deform=False, global_step=int(g["gstep"]), jitter=t("train_jitter").to(cuda), **kw)
    else:
        with torch.no_grad():
            m.s_val.fill_(float(g["eval_s_val"]))
        out = m(o, d, v, use_deform=False, **kw)
    w = lambda k: t("%s_%s" % (tag, k))  # noqa: E731
    assert abs(out["s_val"] - float(g[tag + "_s_val"])) < 1e-6
    M = g[tag + "_weights"].shape[0]
    assert out["weights"].shape == (M,) and out["gradient"].shape == (M, 3) and out["raw_rgb"].shape == (M, 3)
    torch.testing.assert_close(out["gradient"].detach().cpu(), w("gradient"), rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(out["raw_alpha"].cpu(), w("raw_alpha"), rtol=0, atol=2e-4)
    torch.testing.assert_close(out["weights"].cpu(), w("weights"), rtol=0, atol=2e-4)
    torch.testing.assert_close(out["cum_weights"].detach().cpu(), w("cum_weights"), rtol=0, atol=2e-4)
    torch.testing.assert_close(out["alphainv_cum"].detach().cpu(), w("alphainv_cum"), rtol=0, atol=2e-4)
    torch.testing.assert_close(out["rgb_marched"].detach().cpu(), w("rgb_marched"), rtol=0, atol=2e-2)
    torch.testing.assert_close(out["raw_rgb"].cpu(), w("raw_rgb"), rtol=0, atol=3e-2)
    torch.testing.assert_close(out["depth"].cpu(), w("depth"), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["disp"].cpu(), 1 / w("depth"), rtol=1e-3, atol=1e-4)
    assert abs(float(out["k0_tv"]) - float(g[tag + "_k0_tv"])) < 1e-5 * max(1.0, float(g[tag + "_k0_tv"]))
    assert out["mask"].dtype == torch.bool and int(out["mask"].sum()) == M and not bool(out["mask_outbbox"].any())
    assert set(["alphainv_cum", "weights", "cum_weights", "rgb_marched", "normal_marched", "raw_alpha", "raw_rgb", "depth",
                "disp", "mask", "mask_outbbox", "gradient", "s_val", "k0_tv"]) <= set(out.keys())
    if tag == "eval":
        # query_sdf_point_wocuda_render (voxurf_coarse.py:839-920, the depth-only render of lib/recon_scene.py:622-637):
        # depth = t_min + sum_i w_i step_i / |d| - for the unit directions of the golden rays the forward's own depth
        pts, valid, depth = m.query_sdf_point_wocuda_render(o, d, global_step=None, keep_dim=True, use_deform=False, **kw)
        torch.testing.assert_close(depth[:, 0].cpu(), w("depth"), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(pts, o + d * depth, rtol=0, atol=1e-6)
        assert valid.dtype == torch.bool and valid.shape == (o.shape[0],) and int(valid.sum()) > 0
        # the surface point does not depend on the length of the direction vector (depth scales with 1 / |d|)
        pts2, valid2, depth2 = m.query_sdf_point_wocuda_render(o, 2.0 * d, global_step=None, keep_dim=True,
                                                               use_deform=False, **kw)
        torch.testing.assert_close(pts2[valid], pts[valid], rtol=0, atol=2e-4)
        assert torch.equal(valid, valid2)
        only = m.query_sdf_point_wocuda_render(o, d, global_step=None, use_deform=False, **kw)
        assert only.shape == (int(valid.sum()), 3)
        # ... and against the reference's OWN method on non-unit directions (golden query_*: make_golden_voxurf.py)
        with torch.no_grad():
            m.s_val.fill_(float(g["query_s_val"]))
        qp, qm, qd = m.query_sdf_point_wocuda_render(o, float(g["query_dir_scale"]) * d, global_step=None,
                                                     keep_dim=True, use_deform=False, **kw)
        assert torch.equal(qm.cpu(), t("query_mask"))
        torch.testing.assert_close(qd.cpu(), t("query_depth"), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(qp.cpu(), t("query_pts"), rtol=0, atol=3e-4)


def test_voxurf_module_deform_branch_trains_with_the_references_losses(cuda):
    """pp.Voxurf.forward(use_deform=True) - the reference's default and live train step (recon_scene.py:603) - under
    autograd with the reference's loss written in plain Python on the result dict (deform_implicit_loss +
    photometric term, exactly the combination of tests/golden/make_golden_deform.py): the forward entries and every
    parameter gradient against tests/golden/voxurf_deform.npz (the reference's own double backward)."""
    import os
    import numpy as np
    import torch.nn.functional as F
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    m, t = _golden_module(cuda, g, int(g["N_iters"]))
    with torch.no_grad():
        for lin, i in zip(m.warp_network.linears(), range(5)):
            lin.weight.copy_(t("ww%d" % i))
            lin.bias.copy_(t("wb%d" % i))
    assert abs(m.warp_network.output_range - float(g["out_range"])) < 1e-6
    o = t("rays_o").to(cuda).requires_grad_(True)
    d = t("rays_d").to(cuda).requires_grad_(True)
    v = t("viewdirs").to(cuda)
    out = m(o, d, v, global_step=int(g["gstep"]), jitter=t("jitter").to(cuda), near=0.2, far=1e9, stepsize=0.5, bg=0.3)
    M = g["ret_weights"].shape[0]
    assert out["gradient"].shape == (M, 3) and out["grad_deform"].shape == (M, 3, 3)
    assert out["sdf_correct"].shape == (M, 1) and out["sdf_deform"].shape == (M,)
    # fp16 operands in the deformation network: 2e-3 of the output scale (test_warpnet_forward_with_jacobian)
    scale = max(1.0, float(t("ret_sdf_correct").abs().max()) / 1.0)
    torch.testing.assert_close(out["sdf_correct"].detach().cpu(), t("ret_sdf_correct"), rtol=0, atol=2e-3 * scale)
    torch.testing.assert_close(out["sdf_deform"].detach().cpu(), t("ret_sdf_deform"), rtol=0, atol=2e-3 * scale)
    err = (out["grad_deform"].detach().cpu() - t("ret_grad_deform")).flatten(1).norm(dim=1)
    assert float((err < 5e-3).float().mean()) > 0.9 and float(err.median()) < 1e-3       # gate flips, see above
    torch.testing.assert_close(out["rgb_marched"].detach().cpu(), t("ret_rgb_marched"), rtol=0, atol=2e-2)
    # (NeuS alpha amplifies an SDF error by 1 / s_val: the handful of samples on a flipped gate move visibly)
    werr = (out["weights"].cpu() - t("ret_weights")).abs()
    assert float((werr < 5e-3).float().mean()) > 0.99 and float(werr.max()) < 5e-2, (float(werr.max()),)
    # lib/losses.py:6-24 on the dict, as the reference's trainer combines it
    grad_constraint = (out["gradient"].norm(dim=-1) - 1).abs().mean()
    loss = F.mse_loss(out["rgb_marched"], t("target").to(cuda)) + grad_constraint + 0.05 * (
        out["grad_deform"].norm(dim=-1).mean() + out["sdf_correct"].abs().mean() + out["sdf_deform"].abs().mean())
    assert abs(float(loss) - float(g["loss"])) < 1e-2 * float(g["loss"])
    loss.backward()

    def check(name, got, want, c_min=0.999, n_tol=0.01):
        got, want = got.detach().cpu().reshape(-1), want.reshape(-1)
        c = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
        r = float(got.norm() / (want.norm() + 1e-30))
        assert c >= c_min and abs(r - 1) <= n_tol, (name, c, r)

    for i, lin in enumerate(m.warp_network.linears()):
        check("warp weight %d" % i, lin.weight.grad, t("g_ww%d" % i))
        check("warp bias %d" % i, lin.bias.grad, t("g_wb%d" % i), c_min=0.998, n_tol=0.02)
    for i, lin in enumerate([q for q in m.rgbnet.modules() if isinstance(q, nn.Linear)]):
        check("rgbnet weight %d" % i, lin.weight.grad, t("g_w%d" % i))
        check("rgbnet bias %d" % i, lin.bias.grad, t("g_b%d" % i))
    want_k0 = torch.zeros(m.k0.grid.numel())
    want_k0[torch.from_numpy(g["g_k0_idx"]).long()] = t("g_k0_val")
    check("k0", m.k0.grid.grad, want_k0)
    check("rays_o", o.grad, t("g_rays_o"), c_min=0.99, n_tol=0.03)
    check("rays_d", d.grad, t("g_rays_d"), c_min=0.99, n_tol=0.03)
    assert m.sdf.grid.grad is None or float(m.sdf.grid.grad.abs().max()) == 0.0       # frozen in this branch


def test_fused_step_eikonal_term_on_all_inbox_samples(cuda):
    """Non-deform fused step with the live Eikonal term (lib/losses.py:7-10, weight 1.0, mean over ALL in-box samples,
    also past the early termination): SDF gradient of backward_only(w_eikonal=1) minus that of w_eikonal=0 against
    autograd through the oracle's `gradient`."""
    from oracle import voxurf_cpu
    from poseprobe_b200 import ops, voxurf
    G, N = 24, 300
    sdf, k0, net, o, d, v, kw = make_scene(G, N, seed=81)
    target = torch.rand(N, 3)
    rw, rb = ops.rgbnet_from_sequential(net)
    grads = []
    for w_eik in (0.0, 1.0):
        eng = voxurf.FusedVoxurf(sdf.to(cuda), k0.to(cuda), [w.detach() for w in rw], [b.detach() for b in rb], MN, MX,
                                 w_eikonal=w_eik, **kw)
        eng.backward_only(o.to(cuda), d.to(cuda), v.to(cuda), target.to(cuda))
        grads.append(eng.grid_gradients()[0].cpu())
        if w_eik:
            term = float(eng._bufs["reg"][0])
    sdf_r = sdf.clone().requires_grad_(True)
    want = voxurf_cpu.forward(sdf_r, k0, net, o, d, v, MN, MX, kw["stepsize"], eng.voxel_size, kw["s_val"], bg=kw["bg"],
                              sdf_alpha=kw["sdf_alpha"], sdf_beta=kw["sdf_beta"], posbase_pe=5, viewbase_pe=1,
                              barf_c2f=kw["barf_c2f"], progress=1.0)
    eik = (want["gradient"].norm(dim=-1) - 1).abs().mean()
    (g_ref,) = torch.autograd.grad(eik, [sdf_r])
    assert abs(term - float(eik)) < 1e-4 * max(1.0, float(eik))
    got = grads[1] - grads[0]
    rel = float((got - g_ref).norm() / g_ref.norm())
    assert rel < 2e-3, rel


def test_query_sdf_point_through_the_deformation_replays_the_reference(cuda):
    """Voxurf.query_sdf_point_wocuda_render(use_deform=True) - the depth-only second render of every train step
    (lib/recon_scene.py:622-637, voxurf_coarse.py:839-920) - against the reference's OWN method executed on the CPU
    (golden query_* of voxurf_deform.npz, non-unit directions).  The deformation network runs on fp16 tensor-core
    operands here: a sample whose warped SDF moves by ~1e-3 shifts the surface depth by a fraction of a step."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    m, t = _golden_module(cuda, g, int(g["N_iters"]))
    with torch.no_grad():
        for lin, i in zip(m.warp_network.linears(), range(5)):
            lin.weight.copy_(t("ww%d" % i))
            lin.bias.copy_(t("wb%d" % i))
        m.s_val.fill_(float(g["query_s_val"]))
    o, d = t("rays_o").to(cuda), t("rays_d").to(cuda)
    pts, valid, depth = m.query_sdf_point_wocuda_render(o, float(g["query_dir_scale"]) * d, global_step=None,
                                                        keep_dim=True, use_deform=True, near=0.2, far=1e9, stepsize=0.5,
                                                        bg=0.3)
    assert torch.equal(valid.cpu(), t("query_mask"))
    err = (depth.cpu() - t("query_depth")).abs()[:, 0]
    step = 0.5 * float(g["voxel_size"])
    print("query depth error / step: median %.2e max %.2e" % (float(err.median()) / step, float(err.max()) / step))
    # measured: median 5.3e-4 of a step, worst ray 0.10 of a step (a gate flip of the fp16 network next to the surface)
    assert float(err.median()) < 2e-3 * step and float(err.max()) < 0.2 * step, (float(err.median()), float(err.max()))
    torch.testing.assert_close(pts.cpu(), t("query_pts"), rtol=0, atol=0.3 * step)
