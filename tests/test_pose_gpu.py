"""GPU parity of ray generation and the fused per-image pose-gradient reduction against the reference's
get_rays_of_a_view / Lie.se3_to_SE3 restated on the CPU (oracle/oracle.py)."""
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


def _cams(n_img, seed=0):
    g = torch.Generator().manual_seed(seed)
    wu = torch.randn(n_img, 6, generator=g) * 0.2
    base = oracle.se3_to_SE3(wu)
    base[..., 3] += torch.tensor([0.0, 0.0, 2.5])
    H, W = 24, 32
    K = torch.tensor([[40.0, 0, 16.0], [0, 42.0, 12.0], [0, 0, 1]]).repeat(n_img, 1, 1)
    K[:, 0, 0] += torch.arange(n_img)
    return base.contiguous(), K.contiguous(), H, W


@pytest.mark.parametrize("inverse_y", [False, True])
def test_rays_from_pixels_matches_get_rays_of_a_view(cuda, inverse_y):
    from poseprobe_b200 import camera
    c2w, K, H, W = _cams(3)
    for im in range(3):
        ro, rd, vd = oracle.get_rays_of_a_view(H, W, K[im], c2w[im], inverse_y=inverse_y)
        jj, ii = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        px = (ii.flatten() + 0.5).float()
        py = (jj.flatten() + 0.5).float()
        img_id = torch.full((H * W,), im, dtype=torch.int32)
        o, d = camera.rays_from_pixels(c2w.cuda(), K.cuda(), img_id.cuda(), px.cuda(), py.cuda(), inverse_y)
        assert torch.equal(o.cpu(), ro.reshape(-1, 3))
        torch.testing.assert_close(d.cpu(), rd.reshape(-1, 3), rtol=1e-6, atol=1e-7)


def test_pose_gradient_reduction_matches_autograd(cuda):
    from poseprobe_b200 import camera
    n_img = 3
    c2w, K, H, W = _cams(n_img, seed=1)
    g = torch.Generator().manual_seed(2)
    n = 5000
    img_id = torch.randint(0, n_img, (n,), generator=g, dtype=torch.int32)
    px = torch.rand(n, generator=g) * W
    py = torch.rand(n, generator=g) * H
    g_o = torch.randn(n, 3, generator=g)
    g_d = torch.randn(n, 3, generator=g)
    # torch restatement with autograd, fp64
    cd = c2w.double().requires_grad_(True)
    Kd = K.double()
    idl = img_id.long()
    dirs = torch.stack([(px.double() - Kd[idl, 0, 2]) / Kd[idl, 0, 0], -(py.double() - Kd[idl, 1, 2]) / Kd[idl, 1, 1],
                        -torch.ones(n, dtype=torch.float64)], -1)
    u = torch.sum(dirs[:, None, :] * cd[idl, :3, :3], -1)
    rd = u / u.norm(dim=-1, keepdim=True)
    ro = cd[idl, :3, 3]
    ((ro * g_o.double()).sum() + (rd * g_d.double()).sum()).backward()
    cg = c2w.cuda().requires_grad_(True)
    o, d = camera.rays_from_pixels(cg, K.cuda(), img_id.cuda(), px.cuda(), py.cuda())
    ((o * g_o.cuda()).sum() + (d * g_d.cuda()).sum()).backward()
    torch.testing.assert_close(cg.grad.double().cpu(), cd.grad, rtol=1e-4, atol=1e-4)


def test_se3_chain_end_to_end_gradient(cuda):
    """se3_refine -> se3_to_SE3 -> compose -> invert -> rays -> fused march -> loss: gradient flows to the
    6-dof parameters and matches a finite difference."""
    from poseprobe_b200 import camera, dvgo, synthetic
    n_img, G = 3, 24
    c2w0, K, H, W = _cams(n_img, seed=3)
    _, k0 = synthetic.make_dvgo_grids(G, seed=5)
    # a smooth blob that fades to empty space well inside the box: samples that pop in / out at the box
    # boundary when the pose moves then carry no weight, so the loss is differentiable in the pose
    lin = torch.linspace(-1, 1, G)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    density = (8.0 - 40.0 * (x * x + y * y + z * z))[None, None].contiguous()
    k0 = (k0 * 10 + torch.stack([x, y, z])[None] * 3).contiguous()
    eng = dvgo.FusedDVGO(density.cuda(), k0.cuda(), [-1, -1, -1], [1, 1, 1], stepsize=0.5, bg=0.2, pose_grad=True)
    g = torch.Generator().manual_seed(4)
    n = 2048
    img_id = torch.randint(0, n_img, (n,), generator=g, dtype=torch.int32).cuda()
    px = (torch.rand(n, generator=g) * W).cuda()
    py = (torch.rand(n, generator=g) * H).cuda()
    target = torch.rand(n, 3, generator=g).cuda()
    w2c0 = camera.invert(c2w0).cuda()

    def loss_of(se3, want_grad):
        pose = camera.compose_pair(camera.se3_to_SE3(se3), w2c0)       # refine o pnp (recon_scene.py:67-69)
        c2w = camera.invert(pose)
        o, d = camera.rays_from_pixels(c2w, K.cuda(), img_id, px, py)
        if not want_grad:
            rgb, *_ = eng.render(o.detach().contiguous(), d.detach().contiguous())
            return float(torch.nn.functional.mse_loss(rgb.double(), target.double()))
        grad_rays = torch.zeros(n, 6, device="cuda")
        eng.grad.zero_()
        eng.loss_buf.zero_()
        eng.backward_only(o.detach().contiguous(), d.detach().contiguous(), target, grad_rays=grad_rays)
        torch.autograd.backward([o, d], [grad_rays[:, :3].contiguous(), grad_rays[:, 3:].contiguous()])
        return float(eng.loss_buf[1])

    se3 = (torch.randn(n_img, 6, generator=g) * 0.01).cuda().requires_grad_(True)
    loss_of(se3, True)
    grad = se3.grad.clone()
    assert float(grad.abs().max()) > 0
    # directional finite difference
    delta = torch.randn(n_img, 6, generator=g).cuda()
    eps = 1e-3
    fd = (loss_of((se3 + eps * delta).detach(), False) - loss_of((se3 - eps * delta).detach(), False)) / (2 * eps)
    lin = float((grad.double() * delta.double()).sum())
    print('pose chain directional derivative: fd %.6g  analytic %.6g' % (fd, lin))
    assert abs(fd - lin) <= 0.1 * abs(lin) + 1e-6, (fd, lin)
