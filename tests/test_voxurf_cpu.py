"""CPU tier for the fine-stage path: the Voxurf oracle (oracle/voxurf_cpu.py) is pinned on closed forms stated by
the reference's code, and the host-side column map of the fused kernels is checked without a GPU."""
import math

import torch
import torch.nn as nn

from oracle import voxurf_cpu


def test_gradient_grid_is_the_reference_central_difference():
    """neus_sdf_gradient(mode='interpolate'), lib/voxurf_coarse.py:463-467, literally."""
    torch.manual_seed(0)
    sdf = torch.randn(1, 1, 7, 6, 5)
    vs = 0.25
    want = torch.zeros(1, 3, 7, 6, 5)
    want[:, 0, 1:-1, :, :] = (sdf[:, 0, 2:, :, :] - sdf[:, 0, :-2, :, :]) / 2 / vs
    want[:, 1, :, 1:-1, :] = (sdf[:, 0, :, 2:, :] - sdf[:, 0, :, :-2, :]) / 2 / vs
    want[:, 2, :, :, 1:-1] = (sdf[:, 0, :, :, 2:] - sdf[:, 0, :, :, :-2]) / 2 / vs
    assert torch.equal(voxurf_cpu.sdf_gradient_grid(sdf, vs), want)


def test_linear_sdf_gives_constant_gradient_and_normal():
    """An SDF that is linear in x: the sampled gradient is (slope, 0, 0) away from the border slices."""
    G = 16
    lin = torch.linspace(-1, 1, G)
    sdf = (0.7 * lin)[:, None, None].expand(G, G, G)[None, None].contiguous()
    vs = 2.0 / (G - 1)
    g = voxurf_cpu.sdf_gradient_grid(sdf, vs)
    torch.testing.assert_close(g[0, 0, 1:-1], torch.full((G - 2, G, G), 0.7), rtol=1e-5, atol=1e-6)
    assert float(g[0, 1].abs().max()) == 0 and float(g[0, 2].abs().max()) == 0


def test_barf_weights_closed_form():
    """positional_encoding_barf, :721-732: w_k = (1 - cos(pi clamp(alpha - k, 0, 1))) / 2."""
    from poseprobe_b200.voxurf import barf_weights as host_w
    for progress in (0.0, 0.12, 0.3, 0.49, 0.9):
        w = voxurf_cpu.barf_weights(progress, (0.1, 0.5), 5)
        alpha = (progress - 0.1) / 0.4 * 5
        for k in range(5):
            c = min(max(alpha - k, 0.0), 1.0)
            assert abs(float(w[k]) - (1 - math.cos(c * math.pi)) / 2) < 1e-6
        torch.testing.assert_close(w, torch.tensor(host_w(progress, (0.1, 0.5), 5)), rtol=1e-6, atol=1e-6)
    assert torch.equal(voxurf_cpu.barf_weights(0.3, None, 5), torch.ones(5))


def test_feature_permutation_covers_the_reference_row():
    from poseprobe_b200.voxurf import feature_permutation
    perm, d0 = feature_permutation(5, 1)
    assert d0 == 57                                   # 12 + 33 + 9 + 3 (voxurf_coarse.py:198-206)
    used = [p for p in perm if p >= 0]
    assert sorted(used) == list(range(57))            # every reference column feeds exactly one kernel column
    assert perm[:12] == list(range(12))               # k0 first
    assert perm[12:15] == [54, 55, 56] and perm[15] == -1   # normal next: the backward stores 16 gradient columns
    perm3, d3 = feature_permutation(3, 0)
    assert d3 == 12 + 3 + 18 + 3 + 3 and sorted(p for p in perm3 if p >= 0) == list(range(d3))


def test_oracle_forward_properties():
    """Partition of unity of the compositing, clamp range, feature row layout, finite gradients."""
    from poseprobe_b200 import synthetic
    G, N = 20, 48
    sdf = synthetic.make_sdf_grid(G, 0.5, 0.01).requires_grad_(True)
    k0 = (0.3 * torch.randn(1, 12, G, G, G)).requires_grad_(True)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 3))
    o, d, v = synthetic.make_rays(N, seed=3)
    out = voxurf_cpu.forward(sdf, k0, net, o, d, v, [-1, -1, -1], [1, 1, 1], 0.5, 2.0 / G, 0.05, bg=0.5,
                             sdf_alpha=10.0, sdf_beta=0.35, barf_c2f=(0.1, 0.5), progress=0.3)
    cum = out["cum_weights"][:, 0]
    torch.testing.assert_close(cum + out["alphainv_cum"], torch.ones(N), rtol=0, atol=1e-4)
    assert float(out["rgb_marched"].min()) >= 0 and float(out["rgb_marched"].max()) <= 1
    f = out["rgb_feat"]
    assert f.shape[1] == 57
    torch.testing.assert_close(f[:, 54:57].norm(dim=-1), torch.ones(f.shape[0]), rtol=1e-3, atol=1e-3)   # normal
    torch.testing.assert_close(f[:, 45:48], v[out["ray_id"]])                                          # viewdirs
    voxurf_cpu.mse_loss(out, torch.rand(N, 3)).backward()
    assert torch.isfinite(sdf.grad).all() and float(sdf.grad.abs().sum()) > 0
    assert torch.isfinite(k0.grad).all() and float(k0.grad.abs().sum()) > 0


def test_deformation_network_mirror_keeps_the_reference_layout():
    """poseprobe_b200.deform.DeformedImplicitField: state_dict keys / shapes of the reference module
    (lib/deformation/deform_net.py + modules.py nesting, checked against the reference class when
    tests/golden/make_golden_deform.py was written), the reference's initialisation (zero last layer), and the
    golden fixture's warp weights load into it."""
    import os
    import numpy as np
    import torch
    from poseprobe_b200.deform import DeformedImplicitField
    net = DeformedImplicitField(range_shape=torch.tensor([2.0, 1.5, 1.0]))
    want = {"progress": ()}
    shapes = [(128, 3), (128, 128), (128, 128), (128, 128), (4, 128)]
    for i, shp in enumerate(shapes):
        want["deform_net.net.net.%d.0.weight" % i] = shp
        want["deform_net.net.net.%d.0.bias" % i] = (shp[0],)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == want
    assert net.output_range == 2.0
    lin = net.linears()
    assert float(lin[-1].weight.abs().sum()) == 0.0 and float(lin[-1].bias.abs().sum()) == 0.0
    assert 0.5 < float(lin[1].weight.std()) * (128 / 2) ** 0.5 < 1.5        # Kaiming normal, fan_in, ReLU gain
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    sd = {"progress": torch.tensor(0.3)}
    for i in range(5):
        sd["deform_net.net.net.%d.0.weight" % i] = torch.from_numpy(g["ww%d" % i])
        sd["deform_net.net.net.%d.0.bias" % i] = torch.from_numpy(g["wb%d" % i])
    net.load_state_dict(sd)
    ws, bs = net.parameters_for_kernel()
    assert torch.equal(ws[4].detach(), torch.from_numpy(g["ww4"]))


def test_deform_closed_form_matches_autograd():
    """oracle/deform_analytic.py (the formulas the deformation kernels implement: Jacobian by tangent rows, warped
    lookup with its cell Hessian, the four-row backward of the warp network) against autograd through
    oracle/voxurf_cpu's deform branch, on the reference-pinned fixture's network (tests/golden/voxurf_deform.npz)."""
    import os
    import numpy as np
    import torch
    from oracle import deform_analytic as da
    from oracle import voxurf_cpu
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    torch.manual_seed(5)
    ws = [torch.from_numpy(g["ww%d" % i]).clone().requires_grad_(True) for i in range(5)]
    bs = [torch.from_numpy(g["wb%d" % i]).clone().requires_grad_(True) for i in range(5)]
    R = float(g["out_range"])
    grid = voxurf_cpu.softplus10(10.0) * (torch.sigmoid(voxurf_cpu.softplus10(0.35) * torch.from_numpy(g["sdf"])) - 0.5)
    mn, mx = torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1])
    M = 700
    pts = (torch.rand(M, 3) * 1.9 - 0.95).requires_grad_(True)
    # autograd side: the oracle's ops
    deform, corr = voxurf_cpu.warp_mlp(pts, ws, bs, R)
    new = deform + pts
    ones = torch.ones(M)
    grad_deform = torch.stack([torch.autograd.grad(new[:, k], [pts], grad_outputs=ones, create_graph=True)[0]
                               for k in range(3)], dim=2)
    sdf = voxurf_cpu.grid_sample_custom(new, grid, mn, mx)[:, 0] + corr.squeeze(-1)
    sdf_deform = sdf - voxurf_cpu.grid_sample_custom(pts, grid, mn, mx)[:, 0]
    gradient = torch.autograd.grad(sdf, [pts], grad_outputs=torch.ones_like(sdf), create_graph=True)[0]
    a, B = torch.randn(M), torch.randn(M, 3)                       # stand-ins for the render loss's dL/dsdf, dL/dgradient
    reg = voxurf_cpu.deform_implicit_loss({"gradient": gradient, "grad_deform": grad_deform, "sdf_correct": corr,
                                           "sdf_deform": sdf_deform})
    w = (1.0, 0.05, 0.07, 0.09)
    loss = (sdf * a).sum() + (gradient * B).sum() + w[0] * reg["grad_constraint"] + w[1] * reg["grad_deform_constraint"] \
        + w[2] * reg["sdf_correct_constraint"] + w[3] * reg["sdf_deform_constraint"]
    want = torch.autograd.grad(loss, ws + bs + [pts])
    # closed form
    with torch.no_grad():
        wd, bd, x = [t.detach() for t in ws], [t.detach() for t in bs], pts.detach()
        y, Jt, acts = da.warp_rows_forward(x, wd, bd)
        s2, g2, A, sd2, c2, cache = da.lookup_forward(grid, x, y, Jt, R, mn, mx)
        torch.testing.assert_close(s2, sdf.detach(), rtol=0, atol=2e-6)
        torch.testing.assert_close(g2, gradient.detach(), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(A, grad_deform.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(sd2, sdf_deform.detach(), rtol=0, atol=2e-6)
        val, G_g, G_A, G_c, G_d = da.regulariser_grads(g2, A, c2, sd2, *w)
        G_y, G_Jt, G_x = da.lookup_backward(cache, a, B + G_g, G_A, G_c, G_d)
        gw, gb, dx = da.warp_rows_backward(wd, acts, G_y, G_Jt)

    def rel(p, q):
        return float((p - q).norm() / (q.norm() + 1e-12))

    for i in range(5):
        assert rel(gw[i], want[i]) < 1e-4, ("dW", i, rel(gw[i], want[i]))
        assert rel(gb[i], want[5 + i]) < 1e-4, ("db", i, rel(gb[i], want[5 + i]))
    assert rel(G_x + dx, want[10]) < 1e-4, rel(G_x + dx, want[10])
