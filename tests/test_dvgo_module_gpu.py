"""The product's `DirectVoxGO` module (dense compositing, the reference's nine result entries) on the GPU against
tests/golden/dvgo_module.npz = the reference's OWN `DirectVoxGO.forward` / `MaskCache` (lib/dvgo_ori.py) run on the CPU
(tests/golden/make_golden_dvgo_module.py), and the dense compositing op against its oracle restatement.
Tolerances: 1e-5 relative on weights / rgb / gradients (north star); stated at each assert."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dvgo_module.npz")
RK = dict(near=0.2, far=6.0, stepsize=0.5, bg=0.3)


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _load(tag, g, cuda, **ctor):
    import poseprobe_b200 as pp
    m = pp.DirectVoxGO(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=20 ** 3, num_voxels_base=20 ** 3,
                       alpha_init=1e-2, fast_color_thres=1e-4, **ctor).to(cuda)
    sd = {k[len(tag) + 7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "_param_")}
    sd = {k: v for k, v in sd.items() if not k.startswith("mask_cache.") and k != "nonempty_mask"}
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith("mask_cache.") for k in missing.missing_keys), missing
    return m


def _check(tag, g, m, ret, params, cuda, grad_rel):
    t = lambda k: torch.from_numpy(g["%s_%s" % (tag, k)])  # noqa: E731
    assert set(ret) == {"alphainv_cum", "weights", "rgb_marched", "raw_alpha", "raw_rgb", "depth", "disp", "mask",
                        "mask_outbbox"}
    assert torch.equal(ret["mask_outbbox"].cpu(), t("out_mask_outbbox"))
    # the colour mask is weights > 1e-4: a weight within rounding of the threshold may fall on the other side
    flips = ret["mask"].cpu() != t("out_mask")
    assert int(flips.sum()) <= 2
    for k, tol in (("raw_alpha", 1e-6), ("weights", 1e-6), ("alphainv_cum", 1e-6), ("raw_rgb", 2e-6),
                   ("rgb_marched", 2e-6)):
        got, want = ret[k].detach().cpu(), t("out_" + k)
        if k == "raw_rgb":
            got, want = got[~flips], want[~flips]
        torch.testing.assert_close(got, want, rtol=1e-5, atol=tol, msg=lambda s, k=k: k + s)
    torch.testing.assert_close(ret["depth"].detach().cpu(), t("out_depth"), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ret["disp"].detach().cpu(), t("out_disp"), rtol=1e-4, atol=1e-6)
    loss = (ret["rgb_marched"] * t("w_rgb").to(cuda)).sum() + (ret["depth"] * t("w_depth").to(cuda)).sum() + \
        (ret["weights"] * t("w_weights").to(cuda)).sum() + (ret["alphainv_cum"] * t("w_cum").to(cuda)).sum()
    grads = torch.autograd.grad(loss, [p for _, p in params])
    for (name, _), got in zip(params, grads):
        want = t("grad_" + name)
        rel = float((got.cpu() - want).norm() / (want.norm() + 1e-30))
        assert rel < grad_rel, (name, rel)


def test_coarse_training_forward_replays_the_reference(cuda):
    g = np.load(GOLD)
    m = _load("a", g, cuda)
    t = lambda k: torch.from_numpy(g["a_" + k]).to(cuda)  # noqa: E731
    ret = m(t("rays_o"), t("rays_d"), t("viewdirs"), global_step=10, jitter=t("jitter"), **RK)
    _check("a", g, m, ret, [("density", m.density), ("k0", m.k0)], cuda, grad_rel=2e-5)


def test_fine_rgbnet_mask_cache_replays_the_reference(cuda):
    import poseprobe_b200 as pp
    from poseprobe_b200 import dvgo
    g = np.load(GOLD)
    m = _load("b", g, cuda, rgbnet_dim=12, rgbnet_direct=False, rgbnet_depth=3, rgbnet_width=128, posbase_pe=5,
              viewbase_pe=4)
    state = {"MaskCache_kwargs": {"xyz_min": [-1., -1., -1.], "xyz_max": [1., 1., 1.],
                                  "act_shift": float(np.log(1 / (1 - 1e-2) - 1)), "voxel_size_ratio": 1.0},
             "model_state_dict": {"density": torch.from_numpy(g["b_coarse_density"])}}
    density_before = m.density.detach().clone()
    m.set_mask_cache(dvgo.MaskCache(state=state, mask_cache_thres=1e-3))
    assert torch.equal(m.nonempty_mask.cpu(), torch.from_numpy(g["b_nonempty_mask"]))
    assert torch.equal(m.density.detach(), density_before)      # the golden parameters already carry the -100 fill
    t = lambda k: torch.from_numpy(g["b_" + k]).to(cuda)  # noqa: E731
    ret = m(t("rays_o"), t("rays_d"), t("viewdirs"), global_step=None, **RK)
    params = [("density", m.density), ("k0", m.k0)] + [("rgbnet." + n, p) for n, p in m.rgbnet.named_parameters()]
    # (the rgbnet of this non-live module is nn.Sequential: cuBLAS fp32 / TF32-off, so 1e-4 on its gradients)
    _check("b", g, m, ret, params, cuda, grad_rel=1e-4)
    assert isinstance(pp.DirectVoxGO, type)
    # the same module with its colour head on the tcgen05 MLP kernels: bf16 operands -> 2e-2 on colours, and the
    # gradients still point the same way
    ret = m(t("rays_o"), t("rays_d"), t("viewdirs"), global_step=None, **RK)        # (fresh graph)
    m.rgbnet_tensor_cores = True
    ret_tc = m(t("rays_o"), t("rays_d"), t("viewdirs"), global_step=None, **RK)
    torch.testing.assert_close(ret_tc["rgb_marched"], ret["rgb_marched"], rtol=0, atol=2e-2)
    assert torch.equal(ret_tc["weights"], ret["weights"])
    tgt = torch.rand_like(ret["rgb_marched"])
    g_tc = torch.autograd.grad(((ret_tc["rgb_marched"] - tgt) ** 2).mean(), [m.k0, m.rgbnet[0].weight])
    g_32 = torch.autograd.grad(((ret["rgb_marched"] - tgt) ** 2).mean(), [m.k0, m.rgbnet[0].weight])
    for a, b in zip(g_tc, g_32):
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.99, cos


@pytest.mark.parametrize("shape", [(1, 1), (3, 2), (7, 31), (5, 33), (64, 226), (257, 558)])
def test_get_ray_marching_ray_vs_oracle(cuda, shape):
    """pp_dense_composite_* against oracle.get_ray_marching_ray (lib/dvgo_ori.py:478-485) in fp64, incl. alpha = 1
    (the clamp_min(1e-10) branch), alpha = 0 runs and ragged S; value and gradient."""
    from oracle import oracle
    from poseprobe_b200 import ops
    n, S = shape
    gen = torch.Generator().manual_seed(n * 1000 + S)
    alpha = torch.rand(n, S, generator=gen) ** 3
    alpha[torch.rand(n, S, generator=gen) < 0.2] = 0.0
    alpha[torch.rand(n, S, generator=gen) < 0.02] = 1.0
    gw, gc = torch.randn(n, S, generator=gen), torch.randn(n, S + 1, generator=gen)
    a64 = alpha.double().requires_grad_(True)
    w64, c64 = oracle.get_ray_marching_ray(a64)
    (ga64,) = torch.autograd.grad((w64 * gw.double()).sum() + (c64 * gc.double()).sum(), a64)
    a = alpha.to(cuda).requires_grad_(True)
    w, c = ops.get_ray_marching_ray(a)
    assert w.shape == (n, S) and c.shape == (n, S + 1)
    (ga,) = torch.autograd.grad((w * gw.to(cuda)).sum() + (c * gc.to(cuda)).sum(), a)
    torch.testing.assert_close(w.detach().cpu().double(), w64.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(c.detach().cpu().double(), c64.detach(), rtol=1e-5, atol=1e-7)
    # the gradient divides by clamp_min(1 - alpha, 1e-10): where alpha = 1 it is ~1e10 * a cancelling sum; compare
    # the rest relatively and those entries for finiteness only
    ok = alpha < 1.0
    rel = float((ga.cpu().double()[ok] - ga64[ok]).norm() / (ga64[ok].norm() + 1e-30))
    assert rel < 1e-5, rel
    assert bool(torch.isfinite(ga).all())


def test_voxel_count_views_and_scale_volume_grid(cuda):
    """`voxel_count_views` (lib/dvgo_ori.py:201-233) through the trilinear scatter kernel == the same count through
    F.grid_sample's autograd; `scale_volume_grid` resamples both grids and keeps the module renderable."""
    import poseprobe_b200 as pp
    import torch.nn.functional as F
    from poseprobe_b200 import synthetic
    m = pp.DirectVoxGO(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=16 ** 3, num_voxels_base=16 ** 3,
                       alpha_init=1e-2).to(cuda)
    H = W = 12
    o, d, _ = synthetic.make_rays(2 * H * W, seed=3)
    rays_o, rays_d = o.reshape(2 * H, W, 3).to(cuda), d.reshape(2 * H, W, 3).to(cuda)
    got = m.voxel_count_views(rays_o, rays_d, imsz=[H, H], near=0.2, far=6.0, stepsize=0.5)
    # the reference formulation with ATen's sampler
    N_samples = int(np.linalg.norm(np.array(m.density.shape[2:]) + 1) / 0.5) + 1
    rng = torch.arange(N_samples, device=cuda)[None].float()
    want = torch.zeros_like(m.density.detach())
    for ro, rd in zip(rays_o.split([H, H]), rays_d.split([H, H])):
        ones = torch.ones_like(m.density).requires_grad_()
        ro, rd = ro.flatten(0, -2), rd.flatten(0, -2)
        vec = torch.where(rd == 0, torch.full_like(rd, 1e-6), rd)
        t_min = torch.minimum((m.xyz_max - ro) / vec, (m.xyz_min - ro) / vec).amax(-1).clamp(min=0.2, max=6.0)
        pts = ro[..., None, :] + rd[..., None, :] * (t_min[..., None] + 0.5 * m.voxel_size * rng / rd.norm(dim=-1, keepdim=True))[..., None]
        ind = ((pts.reshape(1, 1, 1, -1, 3) - m.xyz_min) / (m.xyz_max - m.xyz_min)).flip((-1,)) * 2 - 1
        F.grid_sample(ones, ind, mode="bilinear", align_corners=True).sum().backward()
        want += (ones.grad > 1)
    # a voxel whose accumulated weight is within rounding of 1 may be counted differently (atomic order): allow a few
    assert float((got != want).float().mean()) < 2e-3 and float(got.sum()) > 0
    with torch.no_grad():
        m.density.normal_()
        m.k0.normal_()
    m.scale_volume_grid(24 ** 3)
    assert tuple(m.density.shape) == (1, 1, 24, 24, 24) and tuple(m.k0.shape) == (1, 3, 24, 24, 24)
    ro, rd = rays_o.flatten(0, -2), rays_d.flatten(0, -2)
    out = m(ro, rd, rd, global_step=None, near=0.2, far=6.0, stepsize=0.5, bg=0.0)
    assert out["rgb_marched"].shape == (2 * H * W, 3) and bool(torch.isfinite(out["rgb_marched"]).all())
