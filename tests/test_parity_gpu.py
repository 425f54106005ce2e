"""GPU parity: every operator of the C ABI (called through the reference-shaped shims) against
(a) the CPU oracle (oracle/pp_oracle.c, oracle/oracle.py) and, when oracle/_ref is present,
(b) the reference's own kernels compiled for sm_100a, on the same seeded inputs.

Bars: integer / bool / index outputs bit-exact; fp32 within rtol 1e-5 (stated per test)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle
from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def pp(cuda):
    import poseprobe_b200
    return poseprobe_b200


def close(a, b, rtol=RTOL, atol=1e-7):
    torch.testing.assert_close(a.cpu(), b.cpu(), rtol=rtol, atol=atol, equal_nan=True)


def exact(a, b):
    a, b = a.cpu(), b.cpu()
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.equal(a, b), "mismatch at %s" % (torch.nonzero(a != b)[:5].tolist(),)


def dev(*ts):
    return [t.cuda() for t in ts]


# --------------------------------------------------------------------------------------------
# sampling
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 33, 1000])
@pytest.mark.parametrize("box", ["unit", "aniso"])
def test_infer_helpers(pp, ref_modules, n, box):
    ru = pp.render_utils_cuda
    o, d = cases.rays_with_edge_cases(n)
    mn, mx = cases.bbox() if box == "unit" else cases.bbox_aniso()
    near, far, stepdist = 0.2, 1e9, 0.5 * 2 / 160
    o_, d_, mn_, mx_ = dev(o, d, mn, mx)
    t_min, t_max = ru.infer_t_minmax(o_, d_, mn_, mx_, near, far)
    ot_min, ot_max = oracle.infer_t_minmax(o, d, mn, mx, near, far)
    exact(t_min, ot_min)        # division/min/max only: correctly rounded on both sides
    exact(t_max, ot_max)
    ns = ru.infer_n_samples(d_, t_min, t_max, stepdist)
    exact(ns, oracle.infer_n_samples(d, ot_min, ot_max, stepdist))
    start, rdir = ru.infer_ray_start_dir(o_, d_, t_min)
    ostart, odir = oracle.infer_ray_start_dir(o, d, ot_min)
    exact(start, ostart)
    exact(rdir, odir)
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None and n > 0:
        r_min, r_max = ref.infer_t_minmax(o_, d_, mn_, mx_, near, far)
        exact(t_min, r_min)
        exact(t_max, r_max)
        exact(ns, ref.infer_n_samples(d_, r_min, r_max, stepdist))
        rs, rd = ref.infer_ray_start_dir(o_, d_, r_min)
        exact(start, rs)
        exact(rdir, rd)


@pytest.mark.parametrize("n,stepdist", [(0, 0.01), (1, 0.01), (257, 0.00625), (4096, 0.0125), (777, 0.05)])
def test_sample_pts_on_rays(pp, ref_modules, n, stepdist):
    ru = pp.render_utils_cuda
    o, d = cases.rays_with_edge_cases(n, seed=n)
    mn, mx = cases.bbox()
    near, far = 0.2, 1e9
    o_, d_, mn_, mx_ = dev(o, d, mn, mx)
    got = ru.sample_pts_on_rays(o_, d_, mn_, mx_, near, far, stepdist)
    want = oracle.sample_pts_on_rays(o, d, mn, mx, near, far, stepdist)
    assert len(got) == 9
    names = ["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max", "rays_start", "rays_dir"]
    for name, g, w in zip(names, got, want):
        exact(g, w)   # every op on this path is a correctly rounded IEEE op (or a pinned FMA)
    # structural invariants
    pts, mask, ray_id, step_id, n_steps = got[:5]
    assert int(n_steps.sum()) == ray_id.numel() == step_id.numel() == mask.numel() == pts.shape[0]
    if n:
        assert int(n_steps.min()) >= 1
        assert bool((ray_id[1:] >= ray_id[:-1]).all())
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None and n > 0:
        rgot = ref.sample_pts_on_rays(o_, d_, mn_, mx_, near, far, stepdist)
        for name, g, w in zip(names, got, rgot):
            exact(g, w)


def test_sample_pts_full_size_properties(pp):
    """BASELINE config sizes (160^3, 8192 rays, stepsize 0.5): size-independent invariants."""
    ru = pp.render_utils_cuda
    from poseprobe_b200 import synthetic
    o, d, _ = synthetic.make_rays(8192)
    mn, mx = cases.bbox()
    stepdist = 0.5 * (2.0 / 160)
    o_, d_, mn_, mx_ = dev(o, d, mn, mx)
    pts, mask, ray_id, step_id, n_steps, t_min, t_max, start, rdir = ru.sample_pts_on_rays(
        o_, d_, mn_, mx_, 0.2, 1e9, stepdist)
    assert int(n_steps.sum()) == ray_id.numel()
    # ray_id is the run-length expansion of N_steps; step_id restarts at 0 on every ray
    exact(ray_id, torch.repeat_interleave(torch.arange(8192, device="cuda"), n_steps))
    cum = torch.cumsum(n_steps, 0) - n_steps
    exact(step_id, torch.arange(ray_id.numel(), device="cuda") - cum[ray_id])
    # points recomputed the way the reference's Python wrapper does (voxurf_coarse.py:682-683)
    again = start[ray_id] + rdir[ray_id] * step_id[:, None] * stepdist
    close(pts, again, rtol=1e-5, atol=1e-5)
    inside = ((pts >= mn_) & (pts <= mx_)).all(-1)
    exact(mask, ~inside)


@pytest.mark.parametrize("n,S", [(0, 8), (5, 2), (300, 65)])
def test_sample_ndc_and_bg(pp, ref_modules, n, S):
    ru = pp.render_utils_cuda
    o, d = cases.rays_with_edge_cases(n, seed=3)
    mn, mx = cases.bbox()
    o_, d_, mn_, mx_ = dev(o, d, mn, mx)
    pts, mask = ru.sample_ndc_pts_on_rays(o_, d_, mn_, mx_, S)
    opts, omask = oracle.sample_ndc_pts_on_rays(o, d, mn, mx, S)
    exact(pts, opts)
    exact(mask, omask)
    t_max = torch.rand(n) + 0.5
    bg = ru.sample_bg_pts_on_rays(o_, d_, t_max.cuda(), 0.5, S)
    close(bg, oracle.sample_bg_pts_on_rays(o, d, t_max, 0.5, S), rtol=1e-5, atol=1e-6)
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None and n > 0:
        rp, rm = ref.sample_ndc_pts_on_rays(o_, d_, mn_, mx_, S)
        exact(pts, rp)
        exact(mask, rm)
        close(bg, ref.sample_bg_pts_on_rays(o_, d_, t_max.cuda(), 0.5, S), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n", [0, 1, 5000])
def test_maskcache_lookup(pp, ref_modules, n):
    ru = pp.render_utils_cuda
    g = torch.Generator().manual_seed(5)
    world = torch.rand(17, 23, 11, generator=g) > 0.5
    mn, mx = cases.bbox_aniso()
    xyz = (torch.rand(n, 3, generator=g) * 2.6 - 1.3)     # some points outside the box
    if n > 10:
        xyz[0] = mn            # exact corners / half-way rounding cases
        xyz[1] = mx
        xyz[2] = (mn + mx) / 2
        xyz[3] = torch.tensor([float("nan"), 0.0, 0.0])
        xyz[4] = torch.tensor([1e30, 0.0, 0.0])
    scale = (torch.tensor(world.shape, dtype=torch.float32) - 1) / (mx - mn)
    shift = -mn * scale
    got = ru.maskcache_lookup(*dev(world, xyz, scale, shift))
    exact(got, oracle.maskcache_lookup(world, xyz, scale, shift))
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None and n > 0:
        exact(got, ref.maskcache_lookup(*dev(world, xyz, scale, shift)))


# --------------------------------------------------------------------------------------------
# density -> alpha
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 3, 4, 7, 100003, 262144])
def test_raw2alpha(pp, ref_modules, n):
    ru = pp.render_utils_cuda
    g = torch.Generator().manual_seed(11)
    density = torch.randn(n, generator=g) * 4
    if n > 8:
        density[:4] = torch.tensor([100.0, -100.0, 0.0, 88.0])     # exp overflow / underflow
    shift, interval = -4.595, 0.5
    gb = torch.randn(n, generator=g)
    iv = torch.rand(n, generator=g) + 0.1
    exp_d, alpha = ru.raw2alpha(density.cuda(), shift, interval)
    oexp, oalpha = oracle.raw2alpha(density, shift, interval)
    close(exp_d, oexp, rtol=1e-5, atol=0)
    close(alpha, oalpha, rtol=1e-5, atol=3e-7)   # 1 - pow(): libm vs libdevice pow differ by <= 2 ulp at 1.0
    grad = ru.raw2alpha_backward(exp_d, gb.cuda(), interval)
    close(grad, oracle.raw2alpha_backward(exp_d.cpu(), gb, interval), rtol=1e-5, atol=1e-30)
    exp_n, alpha_n = ru.raw2alpha_nonuni(density.cuda(), shift, iv.cuda())
    oexp_n, oalpha_n = oracle.raw2alpha_nonuni(density, shift, iv)
    close(alpha_n, oalpha_n, rtol=1e-5, atol=3e-7)
    grad_n = ru.raw2alpha_nonuni_backward(exp_n, gb.cuda(), iv.cuda())
    close(grad_n, oracle.raw2alpha_nonuni_backward(exp_n.cpu(), gb, iv), rtol=1e-5, atol=1e-30)
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None and n > 0:
        rexp, ralpha = ref.raw2alpha(density.cuda(), shift, interval)
        exact(exp_d, rexp)          # same libdevice functions, same operand types
        exact(alpha, ralpha)
        exact(grad, ref.raw2alpha_backward(rexp, gb.cuda(), interval))
        rexp_n, ralpha_n = ref.raw2alpha_nonuni(density.cuda(), shift, iv.cuda())
        exact(alpha_n, ralpha_n)
        exact(grad_n, ref.raw2alpha_nonuni_backward(rexp_n, gb.cuda(), iv.cuda()))


def test_raw2alpha_unaligned_views_and_fp64(pp, ref_modules):
    """Views that start 4 B into an allocation (no 16-byte alignment), short arrays and fp64 give the same bits as
    the aligned fp32 call, element by element, and match the reference kernels (a four-elements-per-thread variant
    was measured and rejected in round 2 - profiles/r2_results.md section 8 - this test pinned it and stays as the
    guard for any later vector form)."""
    ru = pp.render_utils_cuda
    g = torch.Generator().manual_seed(12)
    n = 4099
    density = (torch.randn(n + 1, generator=g) * 4).cuda()
    gb = torch.randn(n + 1, generator=g).cuda()
    iv = (torch.rand(n + 1, generator=g) + 0.1).cuda()
    shift, interval = -4.595, 0.5
    base = ru.raw2alpha(density[1:].clone(), shift, interval)
    for d in (density[1:],):                      # contiguous, data_ptr % 16 == 4
        assert d.data_ptr() % 16 == 4 and d.is_contiguous()
        got = ru.raw2alpha(d, shift, interval)
        exact(got[0], base[0])
        exact(got[1], base[1])
    grad_a = ru.raw2alpha_backward(base[0], gb[1:].clone(), interval)
    exact(ru.raw2alpha_backward(base[0], gb[1:], interval), grad_a)
    en, an = ru.raw2alpha_nonuni(density[1:].clone(), shift, iv[1:].clone())
    en2, an2 = ru.raw2alpha_nonuni(density[1:].clone(), shift, iv[1:])
    exact(an2, an)
    exact(ru.raw2alpha_nonuni_backward(en, gb[1:], iv[1:]), ru.raw2alpha_nonuni_backward(en, gb[1:].clone(), iv[1:].clone()))
    # element-by-element equality of short arrays with single-element calls
    for k in range(1, 5):
        one = torch.cat([ru.raw2alpha(density[j:j + 1].clone(), shift, interval)[1] for j in range(1, 1 + k)])
        exact(ru.raw2alpha(density[1:1 + k].clone(), shift, interval)[1], one)
    d64 = density[1:].double()
    e64, a64 = ru.raw2alpha(d64, shift, interval)
    e_t = torch.exp(d64.cpu() + float(np.float32(shift)))
    close(e64, e_t, rtol=1e-13, atol=0)
    close(a64, 1 - torch.pow(1 + e_t, -float(np.float32(interval))), rtol=1e-12, atol=1e-15)
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None:
        rexp, ralpha = ref.raw2alpha(density[1:].clone(), shift, interval)
        exact(base[0], rexp)
        exact(base[1], ralpha)
        exact(grad_a, ref.raw2alpha_backward(rexp, gb[1:].clone(), interval))
        r64 = ref.raw2alpha(d64, shift, interval)
        exact(e64, r64[0])
        exact(a64, r64[1])


# --------------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_rays,max_len,p_big", [(1, 5, 0.0), (7, 40, 0.3), (100, 600, 0.02), (2049, 90, 0.1),
                                                  (64, 0, 0.0)])
def test_alpha2weight_fwd_bwd(pp, ref_modules, n_rays, max_len, p_big):
    ru = pp.render_utils_cuda
    alpha, ray_id, lens = cases.alpha_and_ray_id(n_rays, max_len, seed=n_rays + max_len, p_big=p_big)
    a_, r_ = dev(alpha, ray_id)
    got = ru.alpha2weight(a_, r_, n_rays)
    want = oracle.alpha2weight(alpha, ray_id, n_rays)
    assert len(got) == 5
    # the recurrence is evaluated in the reference's order and precision: everything is bit-exact
    for g, w in zip(got, want):
        exact(g, w)
    weight, T, last, i_start, i_end = got
    g = torch.Generator().manual_seed(99)
    gw = torch.randn(alpha.numel(), generator=g)
    gl = torch.randn(n_rays, generator=g)
    grad = ru.alpha2weight_backward(a_, weight, T, last, i_start, i_end, n_rays, gw.cuda(), gl.cuda())
    ograd = oracle.alpha2weight_backward(alpha, *want, n_rays, gw, gl)
    exact(grad, ograd)
    ref = ref_modules.get("render_utils_cuda")
    if ref is not None and alpha.numel() > 0:
        rgot = ref.alpha2weight(a_, r_, n_rays)
        for g_, w_ in zip(got, rgot):
            exact(g_, w_)
        exact(grad, ref.alpha2weight_backward(a_, *rgot, n_rays, gw.cuda(), gl.cuda()))


def test_alpha2weight_matches_cumprod_and_autograd(pp):
    """Closed form: without early termination w_i = a_i * prod_{j<i}(1-a_j); backward == autograd."""
    ru = pp.render_utils_cuda
    n_rays, L = 50, 37
    g = torch.Generator().manual_seed(4)
    alpha = (torch.rand(n_rays, L, generator=g) * 0.02).double()
    ray_id = torch.arange(n_rays).repeat_interleave(L)
    a = alpha.reshape(-1).float().cuda().requires_grad_(True)
    w, last = pp.Alphas2Weights.apply(a, ray_id.cuda(), n_rays)
    ad = alpha.clone().requires_grad_(True)
    T = torch.cat([torch.ones(n_rays, 1, dtype=torch.float64), (1 - ad).cumprod(-1)], -1)
    wd = ad * T[:, :-1]
    close(w.double(), wd.reshape(-1), rtol=1e-5, atol=1e-9)
    close(last.double(), T[:, -1], rtol=1e-5, atol=1e-9)
    gw = torch.randn(n_rays, L, generator=g, dtype=torch.float64)
    gl = torch.randn(n_rays, generator=g, dtype=torch.float64)
    ((wd * gw).sum() + (T[:, -1] * gl).sum()).backward()
    ((w * gw.reshape(-1).float().cuda()).sum() + (last * gl.float().cuda()).sum()).backward()
    close(a.grad.double(), ad.grad.reshape(-1), rtol=1e-4, atol=1e-6)


def test_alpha2weight_full_size(pp):
    """8192 rays x ~330 samples (config B): early-termination + fill-value invariants."""
    ru = pp.render_utils_cuda
    g = torch.Generator().manual_seed(8)
    n_rays = 8192
    lens = torch.randint(200, 460, (n_rays,), generator=g)
    ray_id = torch.repeat_interleave(torch.arange(n_rays), lens).cuda()
    alpha = (torch.rand(int(lens.sum()), generator=g) * 0.06).cuda()
    weight, T, last, i_start, i_end = ru.alpha2weight(alpha, ray_id, n_rays)
    cum = torch.cumsum(lens, 0).cuda()
    exact(i_start, cum - lens.cuda())
    assert bool((i_end <= cum).all()) and bool((i_end > i_start).all())
    idx = torch.arange(alpha.numel(), device="cuda")
    past = idx >= i_end[ray_id]
    assert bool((weight[past] == 0).all()) and bool((T[past] == 1).all())
    terminated = i_end < cum
    assert bool((last[terminated] < 1e-3).all())
    live = last >= 1e-3                      # a ray may also terminate exactly on its last sample
    assert bool((i_end[live] == cum[live]).all())
    # sum of weights + final transmittance == 1 on every ray (telescoping identity)
    tot = torch.zeros(n_rays, device="cuda").index_add_(0, ray_id, weight) + last
    close(tot, torch.ones(n_rays), rtol=0, atol=1e-4)    # ~460 float roundings per ray


# --------------------------------------------------------------------------------------------
# dense-grid ops
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 5, 6, 7), (1, 3, 16, 9, 33), (1, 2, 40, 40, 40)])
@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("masked", [False, True])
def test_total_variation(pp, ref_modules, shape, dense, masked):
    tv = pp.total_variation_cuda
    g = torch.Generator().manual_seed(21)
    param = torch.randn(shape, generator=g) * 2          # exercises the +-1 clamp
    grad = torch.randn(shape, generator=g)
    grad[torch.rand(shape, generator=g) < 0.6] = 0        # sparse-mode skips
    mask = (torch.rand(shape, generator=g) > 0.3).float()
    wx, wy, wz = 0.3, 0.5, 0.7
    want = grad.clone()
    oracle.total_variation_add_grad(param, want, wx, wy, wz, dense, mask if masked else None)
    got = grad.clone().cuda()
    if masked:
        tv.total_variation_add_grad_new(param.cuda(), got, mask.cuda(), wx, wy, wz, dense)
    else:
        tv.total_variation_add_grad(param.cuda(), got, wx, wy, wz, dense)
    exact(got, want)     # products and sums are individually rounded on both sides
    ref = ref_modules.get("total_variation_cuda")
    if ref is not None:
        rgot = grad.clone().cuda()
        if masked:
            ref.total_variation_add_grad_new(param.cuda(), rgot, mask.cuda(), wx, wy, wz, dense)
        else:
            ref.total_variation_add_grad(param.cuda(), rgot, wx, wy, wz, dense)
        exact(got, rgot)


@pytest.mark.parametrize("numel", [1, 1023, 4096, 160 * 160 * 4 + 2])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_adam(pp, ref_modules, numel, mode):
    ad = pp.adam_upd_cuda
    g = torch.Generator().manual_seed(31 + mode)
    p = torch.randn(numel, generator=g)
    grad = torch.randn(numel, generator=g)
    grad[torch.rand(numel, generator=g) < 0.5] = 0
    m = torch.randn(numel, generator=g) * 0.1
    v = torch.rand(numel, generator=g) * 0.01
    perlr = torch.rand(numel, generator=g)
    step, b1, b2, lr, eps = 7, 0.9, 0.99, 0.1, 1e-8
    wp, wm, wv = p.clone(), m.clone(), v.clone()
    oracle.adam_upd(mode, wp, grad, wm, wv, perlr if mode == 2 else None, step, b1, b2, lr, eps)
    gp, gm, gv = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
    fn = [ad.adam_upd, ad.masked_adam_upd, None][mode]
    if mode == 2:
        ad.adam_upd_with_perlr(gp, grad.cuda(), gm, gv, perlr.cuda(), step, b1, b2, lr, eps)
    else:
        fn(gp, grad.cuda(), gm, gv, step, b1, b2, lr, eps)
    exact(gm, wm)
    exact(gv, wv)
    exact(gp, wp)
    ref = ref_modules.get("adam_upd_cuda")
    if ref is not None:
        rp, rm, rv = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
        if mode == 2:
            ref.adam_upd_with_perlr(rp, grad.cuda(), rm, rv, perlr.cuda(), step, b1, b2, lr, eps)
        else:
            [ref.adam_upd, ref.masked_adam_upd][mode](rp, grad.cuda(), rm, rv, step, b1, b2, lr, eps)
        exact(gm, rm)
        exact(gv, rv)
        exact(gp, rp)


@pytest.mark.parametrize("n_rays,n_pts", [(1, 1), (33, 65), (500, 128), (3, 1000)])
def test_cumdist_thres(pp, ref_modules, n_rays, n_pts):
    g = torch.Generator().manual_seed(41)
    dist = torch.rand(n_rays, n_pts, generator=g) * 0.1
    thres = 0.33
    got = pp.ub360_utils_cuda.cumdist_thres(dist.cuda(), thres)
    exact(got, oracle.cumdist_thres(dist, thres))
    ref = ref_modules.get("ub360_utils_cuda")
    if ref is not None:
        exact(got, ref.cumdist_thres(dist.cuda(), thres))


# --------------------------------------------------------------------------------------------
# python-side operators of the reference
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,shape", [(1, (20, 24, 28)), (3, (16, 16, 16)), (12, (9, 10, 11))])
@pytest.mark.parametrize("padding", ["zeros", "border"])
def test_trilinear_vs_grid_sample(pp, C, shape, padding):
    g = torch.Generator().manual_seed(51)
    grid = torch.randn(1, C, *shape, generator=g)
    mn, mx = cases.bbox_aniso()
    n = 4000
    xyz = mn + (mx - mn) * (torch.rand(n, 3, generator=g) * 1.2 - 0.1)     # 10% margin outside the box
    xyz[0] = mn
    xyz[1] = mx
    # (a) CPU oracle = the reference's own call into F.grid_sample
    want = oracle.dense_grid_forward(grid, xyz, mn, mx, padding_mode=padding).reshape(n, C)
    grid_ = grid.cuda().requires_grad_(True)
    xyz_ = xyz.cuda().requires_grad_(True)
    got = pp.ops.trilinear(grid_, xyz_, mn, mx, padding_mode=padding)
    close(got, want, rtol=1e-5, atol=1e-6)
    # (b) torch's CUDA grid_sampler on the same device, forward and backward
    grid_t = grid.cuda().requires_grad_(True)
    xyz_t = xyz.cuda().requires_grad_(True)
    ind = ((xyz_t.reshape(1, 1, 1, -1, 3) - mn.cuda()) / (mx.cuda() - mn.cuda())).flip((-1,)) * 2 - 1
    ref = F.grid_sample(grid_t, ind, mode="bilinear", align_corners=True, padding_mode=padding)
    ref = ref.reshape(C, -1).T
    close(got, ref, rtol=1e-5, atol=1e-6)
    go = torch.randn(n, C, generator=g).cuda()
    got.backward(go)
    ref.backward(go)
    close(grid_.grad, grid_t.grad, rtol=1e-4, atol=1e-5)     # atomics: summation order differs
    # points exactly on a cell boundary have a one-sided derivative; compare strictly inside cells
    close(xyz_.grad[2:], xyz_t.grad[2:], rtol=1e-3, atol=1e-3)


def test_trilinear_spatial_gradient(pp):
    """Analytic d value / d xyz equals autograd through the gather."""
    g = torch.Generator().manual_seed(52)
    grid = torch.randn(1, 2, 12, 13, 14, generator=g).cuda()
    mn, mx = cases.bbox_aniso()
    xyz = (mn + (mx - mn) * torch.rand(500, 3, generator=g)).cuda().requires_grad_(True)
    val, dxyz = pp.ops.trilinear_with_grad(grid, xyz, mn, mx)
    out = pp.ops.trilinear(grid, xyz, mn, mx)
    close(val, out, rtol=0, atol=0)
    for c in range(2):
        (gx,) = torch.autograd.grad(out[:, c].sum(), xyz, retain_graph=True)
        close(dxyz[:, c], gx, rtol=1e-5, atol=1e-6)


def test_dense_grid_module(pp):
    """DenseGrid drop-in: same forward/backward as the reference module's F.grid_sample path."""
    g = torch.Generator().manual_seed(53)
    mn, mx = cases.bbox()
    for C in (1, 3):
        dg = pp.DenseGrid(C, [8, 9, 10], mn.tolist(), mx.tolist()).cuda()
        with torch.no_grad():
            dg.grid.copy_(torch.randn(dg.grid.shape, generator=g))
        xyz = (torch.rand(7, 11, 3, generator=g) * 2 - 1).cuda()
        out = dg(xyz)
        want = oracle.dense_grid_forward(dg.grid.detach().cpu(), xyz.cpu(), mn, mx)
        assert out.shape == want.shape
        close(out, want, rtol=1e-5, atol=1e-6)
        out.sum().backward()
        assert dg.grid.grad is not None and float(dg.grid.grad.abs().sum()) > 0
        dg.total_variation_add_grad(1.0, 1.0, 1.0, True)


def test_grid_modules_replay_reference_golden(pp):
    """DenseGrid / MaskGrid on the GPU against tests/golden/grid_modules.npz (the reference's own lib/grid.py classes,
    see tests/golden/make_golden_grid.py): forward values and shapes, TV add-grad with and without the broadcast
    [1,1,X,Y,Z] mask, the mask lookup (bit-exact)."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grid_modules.npz"))
    lo, hi = g["mg_lo"].tolist(), g["mg_hi"].tolist()
    mg = pp.MaskGrid(mask=torch.from_numpy(g["mg_mask"]), xyz_min=lo, xyz_max=hi).cuda()
    hit = mg(torch.from_numpy(g["mg_pts"]).cuda())
    assert hit.dtype == torch.bool and hit.shape == g["mg_hit"].shape
    assert np.array_equal(hit.cpu().numpy(), g["mg_hit"])
    dg = pp.DenseGrid(channels=3, world_size=[4, 6, 5], xyz_min=lo, xyz_max=hi).cuda()
    with torch.no_grad():
        dg.grid.copy_(torch.from_numpy(g["dg_grid"]))
    dg.grid.grad = torch.from_numpy(g["dg_grad0"]).cuda()
    dg.total_variation_add_grad(0.3, 0.2, 0.1, True)
    close(dg.grid.grad, torch.from_numpy(g["dg_grad_tv"]), rtol=1e-6, atol=1e-6)
    dg.total_variation_add_grad(0.3, 0.2, 0.1, False, mask=torch.from_numpy(g["dg_tvmask"]).cuda())
    close(dg.grid.grad, torch.from_numpy(g["dg_grad_tv_masked"]), rtol=1e-6, atol=1e-6)
    val = dg(torch.from_numpy(g["dg_q"]).cuda())
    assert val.shape == g["dg_val"].shape
    close(val, torch.from_numpy(g["dg_val"]), rtol=1e-5, atol=1e-6)
    d1 = pp.DenseGrid(channels=1, world_size=[4, 6, 5], xyz_min=lo, xyz_max=hi).cuda()
    assert list(d1(torch.from_numpy(g["dg_q"]).cuda()).shape) == g["d1_shape"].tolist()


@pytest.mark.parametrize("C", [1, 3])
def test_segment_sum_and_cumsum(pp, C):
    alpha, ray_id, lens = cases.alpha_and_ray_id(300, 700, seed=61)
    g = torch.Generator().manual_seed(62)
    src = torch.randn(ray_id.numel(), C, generator=g)
    got = pp.ops.segment_sum(src.cuda(), ray_id.cuda(), 300)
    close(got, oracle.segment_sum(src, ray_id, 300), rtol=1e-4, atol=1e-4)
    if C == 1:
        s = src[:, 0].contiguous()
        cs = pp.ops.segment_cumsum(s.cuda(), ray_id.cuda())
        close(cs, oracle.segment_cumsum(s, ray_id), rtol=1e-4, atol=1e-4)


def test_neus_alpha(pp):
    g = torch.Generator().manual_seed(71)
    n_rays, m = 50, 5000
    ray_id = torch.sort(torch.randint(0, n_rays, (m,), generator=g)).values
    viewdirs = F.normalize(torch.randn(n_rays, 3, generator=g), dim=-1)
    sdf = torch.randn(m, generator=g) * 0.05
    grad = F.normalize(torch.randn(m, 3, generator=g), dim=-1) * (1 + 0.1 * torch.randn(m, 1, generator=g))
    dist, inv_s = 0.00625, 64.0
    sdf_t = sdf.clone().requires_grad_(True)
    grad_t = grad.clone().requires_grad_(True)
    want = oracle.neus_alpha(viewdirs, ray_id, dist, sdf_t, grad_t, inv_s)
    sdf_g = sdf.cuda().requires_grad_(True)
    grad_g = grad.cuda().requires_grad_(True)
    got = pp.ops.neus_alpha(viewdirs.cuda(), ray_id.cuda(), dist, sdf_g, grad_g, inv_s)
    close(got, want, rtol=1e-5, atol=1e-6)
    go = torch.randn(m, generator=g)
    got.backward(go.cuda())
    # the two sigmoid branches nearly cancel in d alpha / d sdf, so fp32 autograd is itself only good to
    # ~1e-5 of the gradient scale: judge against the same formula evaluated in fp64
    sdf_d = sdf.double().requires_grad_(True)
    grad_d = grad.double().requires_grad_(True)
    oracle.neus_alpha(viewdirs.double(), ray_id, dist, sdf_d, grad_d, inv_s).backward(go.double())
    for mine, ref64 in ((sdf_g.grad, sdf_d.grad), (grad_g.grad, grad_d.grad)):
        scale = float(ref64.abs().max())
        close(mine.double(), ref64, rtol=1e-4, atol=2e-6 * scale)


# --------------------------------------------------------------------------------------------
# fp64 dispatch (the reference dispatches float and double) and error behaviour
# --------------------------------------------------------------------------------------------
def test_fp64_against_reference(pp, ref_modules):
    ref = ref_modules.get("render_utils_cuda")
    if ref is None:
        pytest.skip("oracle/_ref not built")
    ru = pp.render_utils_cuda
    o, d = cases.rays_with_edge_cases(300, seed=9)
    mn, mx = cases.bbox()
    o_, d_, mn_, mx_ = [t.double().cuda() for t in (o, d, mn, mx)]
    got = ru.sample_pts_on_rays(o_, d_, mn_, mx_, 0.2, 1e9, 0.0125)
    want = ref.sample_pts_on_rays(o_, d_, mn_, mx_, 0.2, 1e9, 0.0125)
    for g_, w_ in zip(got, want):
        exact(g_, w_)
    alpha, ray_id, _ = cases.alpha_and_ray_id(200, 80, seed=5, p_big=0.2)
    a_, r_ = alpha.double().cuda(), ray_id.cuda()
    g2 = ru.alpha2weight(a_, r_, 200)
    w2 = ref.alpha2weight(a_, r_, 200)
    for g_, w_ in zip(g2, w2):
        exact(g_, w_)
    gen = torch.Generator().manual_seed(1)
    gw = torch.randn(alpha.numel(), generator=gen, dtype=torch.float64).cuda()
    gl = torch.randn(200, generator=gen, dtype=torch.float64).cuda()
    close(ru.alpha2weight_backward(a_, *g2, 200, gw, gl), ref.alpha2weight_backward(a_, *w2, 200, gw, gl),
          rtol=1e-12, atol=1e-14)
    dens = torch.randn(1000, generator=gen, dtype=torch.float64).cuda()
    e1, a1 = ru.raw2alpha(dens, -4.0, 0.5)
    e2, a2 = ref.raw2alpha(dens, -4.0, 0.5)
    close(a1, a2, rtol=1e-13, atol=1e-15)
    close(ru.raw2alpha_backward(e1, gw[:1000].contiguous(), 0.5), ref.raw2alpha_backward(e2, gw[:1000].contiguous(), 0.5),
          rtol=1e-13, atol=1e-15)


def test_error_behaviour(pp):
    ru = pp.render_utils_cuda
    a = torch.rand(10)
    with pytest.raises(RuntimeError, match="alpha must be a CUDA tensor"):
        ru.alpha2weight(a, torch.zeros(10, dtype=torch.int64), 1)
    nc = torch.rand(10, 2).cuda()[:, 0]
    with pytest.raises(RuntimeError, match="density must be contiguous"):
        ru.raw2alpha(nc, 0.0, 0.5)
    with pytest.raises(RuntimeError, match="not implemented for 'Half'"):
        ru.raw2alpha(torch.rand(4).cuda().half(), 0.0, 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.ops.trilinear(torch.zeros(1, 1, 2, 2, 2), torch.zeros(1, 3), torch.zeros(3), torch.ones(3))


@pytest.mark.parametrize("masked", [False, True])
def test_total_variation_loss_python_side(cuda, masked):
    """ops.total_variation == the reference's Python `total_variation` (voxurf_coarse.py:1298-1313) and its autograd."""
    from oracle import oracle
    from poseprobe_b200 import ops
    torch.manual_seed(3)
    v = torch.randn(1, 5, 19, 33, 21)
    v[0, 0, 3:6, 4:9] = 0.25                       # flat region: |x|' = 0 at 0, like torch.abs
    mask = None
    if masked:
        mask = (torch.rand(1, 1, 19, 33, 21) > 0.3).repeat(1, 5, 1, 1, 1)
    vr = v.clone().requires_grad_(True)
    want = oracle.total_variation(vr, mask)
    want.backward()
    vc = v.cuda().requires_grad_(True)
    got = ops.total_variation(vc, mask.cuda() if masked else None)
    (2.5 * got).backward()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(vc.grad.cpu(), 2.5 * vr.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("padding", ["custom", "border", "zeros"])
def test_trilinear_is_twice_differentiable(pp, padding):
    """Op-level double backward of the trilinear gather (a6 / SURVEY 8f rank 2): the reference's own sampler
    grid_sample_3d (lib/voxurf_coarse.py:545-659) exists because F.grid_sample had no second derivative; its use is
    torch.autograd.grad(value, xyz, create_graph=True) followed by a loss on that gradient (:966-985, lib/losses.py:
    7-10).  Same computation through pp.ops.trilinear against CPU autograd (fp64) through the oracle's restatement of
    the sampler: first-order gradient, and the gradients of a loss on value AND spatial gradient w.r.t. grid and
    points (cell-derivative scatter, off-diagonal cell Hessian).  fp32 kernels: 1e-4 relative."""
    from oracle import voxurf_cpu
    from poseprobe_b200 import ops
    g = torch.Generator().manual_seed(91)
    mn, mx = torch.tensor([-1.0, -0.8, -1.2]), torch.tensor([1.1, 0.9, 1.0])
    grid = torch.randn(1, 3, 9, 7, 11, generator=g)
    n = 3000
    xyz = mn + (mx - mn) * (torch.rand(n, 3, generator=g) * (1.3 if padding == "custom" else 0.98) + (
        -0.15 if padding == "custom" else 0.01))
    # the spatial gradient of a trilinear interpolant jumps at cell faces: drop the few points that fp32 and fp64 could
    # place on different sides of one
    vox = (xyz.double() - mn.double()) / (mx.double() - mn.double()) * (torch.tensor([9.0, 7.0, 11.0]).double() - 1)
    frac = vox - vox.floor()
    keep = ((frac > 1e-3) & (frac < 1 - 1e-3)).all(-1)
    xyz = xyz[keep]
    n = xyz.shape[0]
    wv, wg = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)

    def run(sampler, gr, pts, dt):
        val = sampler(gr, pts)
        (dx,) = torch.autograd.grad(val.sum(), pts, create_graph=True)
        loss = (val * wv.to(pts.device, dt)).sum() + (dx * wg.to(pts.device, dt)).sum() + (dx.norm(dim=-1) - 1).abs().mean()
        gg, gp = torch.autograd.grad(loss, [gr, pts])
        return val.detach(), dx.detach(), gg, gp

    def ref_sampler(gr, pts):
        # (F.grid_sample has no second derivative - the reason the reference carries its own sampler; for the zeros /
        # border modes the points are kept strictly inside the box, where all three modes are the same function)
        return voxurf_cpu.grid_sample_custom(pts, gr, mn.double(), mx.double())

    want = run(ref_sampler, grid.double().requires_grad_(True), xyz.double().requires_grad_(True), torch.float64)
    got = run(lambda gr, pts: ops.trilinear(gr, pts, mn, mx, padding), grid.cuda().requires_grad_(True),
              xyz.cuda().requires_grad_(True), torch.float32)
    for name, a, b in zip(("value", "d value / d xyz", "d loss / d grid", "d loss / d xyz"), got, want):
        rel = float((a.cpu().double() - b).norm() / b.norm())
        assert rel < 1e-4, (name, rel)
    # the module-level mirrors of the reference's wrappers
    import poseprobe_b200 as ppkg
    if padding == "custom":
        opt = (((xyz - mn) / (mx - mn)).flip(-1) * 2 - 1).reshape(1, 1, 1, -1, 3)
        img = ops.grid_sample_3d(grid.cuda(), opt.cuda())
        assert img.shape == (1, 3, 1, 1, n)
        torch.testing.assert_close(img.reshape(3, -1).T.cpu(), want[0].float(), rtol=1e-4, atol=1e-5)


def test_alpha2weight_termination_at_tile_boundaries(pp):
    """Early termination exactly before / at / after the 16-sample tile boundaries of the chain kernel, rays shorter
    than, equal to and one longer than a tile, empty rays, a ray count that is not a multiple of 32, and a ray of 2500
    samples: forward and backward bit-exact against the C oracle (render_utils_kernel.cu:577-707)."""
    ru = pp.render_utils_cuda
    # (alpha, length): 0.65^n crosses 1e-3 at n = 17, 0.649 at 16, 0.8031 at 32, 0.8085 at 33, 0.35 at 15/16 ...
    spec = [(0.35, 40), (0.351, 16), (0.351, 17), (0.1969, 64), (0.1915, 33), (0.1915, 32), (0.5, 1), (0.0, 0),
            (0.9, 15), (1.0, 20), (1e-4, 2500), (0.3508, 31), (0.0, 16), (0.02, 47)] * 3 + [(0.2, 5)]
    n_rays = len(spec) + 2                      # two trailing rays without samples
    alpha = torch.cat([torch.full((n,), a) for a, n in spec])
    g = torch.Generator().manual_seed(0)
    alpha = (alpha * (1 + 1e-3 * torch.randn(alpha.numel(), generator=g))).clamp(0, 1)
    ray_id = torch.cat([torch.full((n,), r, dtype=torch.int64) for r, (_, n) in enumerate(spec)])
    got = ru.alpha2weight(alpha.cuda(), ray_id.cuda(), n_rays)
    want = oracle.alpha2weight(alpha, ray_id, n_rays)
    for a_, b_ in zip(got, want):
        exact(a_, b_)
    cut = (want[4] - want[3])                   # composited samples per ray
    assert int(cut[0]) in (16, 17) and int(cut[10]) == 2500 and int(cut[7]) == 0 and int(cut[-1]) == 0
    gw = torch.randn(alpha.numel(), generator=g)
    gl = torch.randn(n_rays, generator=g)
    grad = ru.alpha2weight_backward(alpha.cuda(), *got, n_rays, gw.cuda(), gl.cuda())
    exact(grad, oracle.alpha2weight_backward(alpha, *want, n_rays, gw, gl))
