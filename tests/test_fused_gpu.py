"""GPU parity of the fused DVGO render / train kernels against the composed-operator CPU oracle
(oracle/dvgo_cpu.py) and size-independent properties at the BASELINE config-B shape."""
import pytest
import torch

from oracle import dvgo_cpu, oracle
from poseprobe_b200 import synthetic
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp(cuda):
    import poseprobe_b200
    from poseprobe_b200 import dvgo
    return dvgo


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.cpu(), b.cpu(), rtol=rtol, atol=atol)


def small_scene(G=20, n=300, seed=0, density_scale=3.0):
    density, k0 = synthetic.make_dvgo_grids(G, seed=seed)
    density = density * density_scale + 2.0          # dense enough that some rays terminate early
    o, d = cases.rays_with_edge_cases(n, seed=seed)
    d = d / d.norm(dim=-1, keepdim=True)
    mn, mx = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    kw = dict(near=0.2, far=1e9, stepdist=0.5 * 2.0 / G, act_shift=synthetic.act_shift(), interval=0.5)
    return density, k0, o.contiguous(), d.contiguous(), mn, mx, kw


@pytest.mark.parametrize("bg,thres", [(0.0, 0.0), (1.0, 0.0), (0.5, 1e-4)])
def test_render_forward_matches_composed_ops(pp, bg, thres):
    density, k0, o, d, mn, mx, kw = small_scene()
    want = dvgo_cpu.render_composed(density, k0, o, d, mn, mx, bg=bg, fast_color_thres=thres, **kw)
    sc = pp.make_scene(density.shape[2:], mn, mx, kw["near"], kw["far"], kw["stepdist"], kw["act_shift"],
                       kw["interval"], bg, thres)
    vox = pp.pack_voxels(density.cuda(), k0.cuda())
    rgb, last, depth, cnt = pp.render_forward(sc, vox, o.cuda(), d.cuda(), want_count=True)
    assert int((want["n_marched"] < torch.bincount(want["ray_id"], minlength=o.shape[0])).sum()) > 0, \
        "scene should exercise early termination"
    # same samples and the same early-termination index.  The cut is decided by `T < 1e-3` on fp32 values
    # that differ in the last bits between implementations (libm vs libdevice pow already do), so a ray
    # whose transmittance lands within rounding of the threshold may stop one sample apart: allow < 1 %.
    same = cnt.cpu() == want["n_marched"]
    assert float(same.float().mean()) >= 0.99, float(same.float().mean())
    assert int((cnt.cpu() - want["n_marched"]).abs().max()) <= 1
    close(last.cpu()[same], want["alphainv_last"][same], rtol=1e-5, atol=1e-7)
    close(rgb.cpu()[same], want["rgb_marched"][same], rtol=1e-5, atol=1e-6)
    close(depth.cpu()[same], want["depth"][same], rtol=1e-5, atol=1e-6)
    # ... and the rays that stop one sample apart are bounded, not excluded: the extra / missing sample carries a weight
    # below T = 1e-3 (colour <= 1, distance <= the box diagonal), and alphainv_last moves by at most that T
    close(rgb.cpu()[~same], want["rgb_marched"][~same], rtol=0, atol=2e-3)
    close(last.cpu()[~same], want["alphainv_last"][~same], rtol=0, atol=1e-3)
    close(depth.cpu()[~same], want["depth"][~same], rtol=0, atol=4e-3)
    # exact mode (PP_MARCH_FLAGS=2: raw2alpha's powf expression and alpha2weight's serial double-precision recurrence
    # in the reference's order): the early-termination index is the operator chain's on EVERY ray
    import os
    os.environ["PP_MARCH_FLAGS"] = "2"
    try:
        rgb_x, last_x, depth_x, cnt_x = pp.render_forward(sc, vox, o.cuda(), d.cuda(), want_count=True)
    finally:
        del os.environ["PP_MARCH_FLAGS"]
    assert torch.equal(cnt_x.cpu(), want["n_marched"]), int((cnt_x.cpu() != want["n_marched"]).sum())
    close(last_x.cpu(), want["alphainv_last"], rtol=1e-5, atol=1e-7)
    close(rgb_x.cpu(), want["rgb_marched"], rtol=1e-5, atol=1e-6)
    close(depth_x.cpu(), want["depth"], rtol=1e-5, atol=1e-6)
    # pack / unpack round trip is exact
    d2, k2 = pp.unpack_voxels(vox, density.shape[2:])
    assert torch.equal(d2.cpu(), density) and torch.equal(k2.cpu(), k0)


@pytest.mark.parametrize("thres", [0.0, 1e-4])
def test_render_backward_matches_autograd_of_composed_ops(pp, thres):
    density, k0, o, d, mn, mx, kw = small_scene(seed=1)
    bg = 0.7
    dens_t = density.clone().requires_grad_(True)
    k0_t = k0.clone().requires_grad_(True)
    out = dvgo_cpu.render_composed(dens_t, k0_t, o, d, mn, mx, bg=bg, fast_color_thres=thres, **kw)
    g = torch.Generator().manual_seed(3)
    g_rgb = torch.randn(o.shape[0], 3, generator=g)
    g_last = torch.randn(o.shape[0], generator=g)
    ((out["rgb_marched"] * g_rgb).sum() + (out["alphainv_last"] * g_last).sum()).backward()
    sc = pp.make_scene(density.shape[2:], mn, mx, kw["near"], kw["far"], kw["stepdist"], kw["act_shift"],
                       kw["interval"], bg, thres)
    dens_g = density.cuda().requires_grad_(True)
    k0_g = k0.cuda().requires_grad_(True)
    rgb, last, depth = pp._RenderFn.apply(dens_g, k0_g, o.cuda(), d.cuda(), sc)
    ((rgb * g_rgb.cuda()).sum() + (last * g_last.cuda()).sum()).backward()
    for mine, ref in ((dens_g.grad, dens_t.grad), (k0_g.grad, k0_t.grad)):
        scale = float(ref.abs().max())
        close(mine, ref, rtol=1e-4, atol=1e-5 * scale)      # float atomics: summation order differs


@pytest.mark.parametrize("loss_w", [(1.0, 0.0, 0.0), (1.0, 0.01, 0.1)])
def test_train_step_matches_composed_ops_and_masked_adam(pp, loss_w):
    density, k0, o, d, mn, mx, kw = small_scene(seed=2, n=256)
    target = torch.rand(o.shape[0], 3, generator=torch.Generator().manual_seed(5))
    w_rgb, w_ent, w_per = loss_w
    dens_t = density.clone().requires_grad_(True)
    k0_t = k0.clone().requires_grad_(True)
    out = dvgo_cpu.render_composed(dens_t, k0_t, o, d, mn, mx, bg=1.0, **kw)
    loss, mse = dvgo_cpu.losses(out, target, w_rgb, w_ent, w_per)
    loss.backward()
    eng = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, bg=1.0, lr_density=0.1, lr_k0=0.05,
                       w_rgb=w_rgb, w_entropy_last=w_ent, w_rgbper=w_per)
    rgb = torch.empty(o.shape[0], 3, device="cuda")
    eng.backward_only(o.cuda(), d.cuda(), target.cuda(), rgb_out=rgb)
    close(rgb, out["rgb_marched"].detach(), rtol=1e-5, atol=1e-6)
    close(eng.loss_buf[0], loss.detach(), rtol=1e-4, atol=1e-7)
    close(eng.loss_buf[1], mse.detach(), rtol=1e-4, atol=1e-7)
    gd, gk = eng.export_grad()
    for mine, ref in ((gd, dens_t.grad), (gk, k0_t.grad)):
        close(mine, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))
    # optimiser: masked Adam of the reference kernel on the oracle's gradients
    pd, pk = density.clone(), k0.clone()
    for p, gr, lr in ((pd, dens_t.grad, 0.1), (pk, k0_t.grad, 0.05)):
        dvgo_cpu.masked_adam(p, gr, torch.zeros_like(p), torch.zeros_like(p), 1, 0.9, 0.99, lr, 1e-8)
    eng.optimizer_step()
    nd, nk = eng.export_grids()
    # the first Adam step moves every touched element by ~lr * sign(grad): compare where the gradient is
    # clearly non-zero (tiny gradients may differ in sign between the two summation orders)
    for mine, ref, gr in ((nd, pd, dens_t.grad), (nk, pk, k0_t.grad)):
        solid = gr.abs() > 1e-4 * float(gr.abs().max())
        close(mine.cpu()[solid], ref[solid], rtol=1e-4, atol=1e-5)
        untouched = gr == 0
        assert torch.equal(mine.cpu()[untouched], ref[untouched])
    assert float(eng.grad.abs().max()) == 0.0            # gradients were reset in the same pass


def test_pose_gradient_matches_autograd(pp):
    """d loss / d rays_o, rays_d through the sample positions == autograd of the dense python sampler
    (lib/voxurf_coarse.py:697-719 formulas) on a scene without early termination."""
    G = 16
    density, k0 = synthetic.make_dvgo_grids(G, seed=4)
    o, d, _ = synthetic.make_rays(64, seed=9)
    mn, mx = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    stepdist, shift, interval = 0.5 * 2.0 / G, synthetic.act_shift(), 0.5
    target = torch.rand(64, 3, generator=torch.Generator().manual_seed(6))
    eng = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, bg=0.3, pose_grad=True)
    grad_rays = torch.zeros(64, 6, device="cuda")
    eng.backward_only(o.cuda(), d.cuda(), target.cuda(), grad_rays=grad_rays)
    # fp64 torch restatement with differentiable sample positions
    od = o.double().requires_grad_(True)
    dd = d.double().requires_grad_(True)
    mn_t, mx_t = torch.tensor(mn).double(), torch.tensor(mx).double()
    pts_f, mask, ray_id, step_id, *_ = oracle.sample_pts_on_rays(o, d, torch.tensor(mn), torch.tensor(mx), 0.2, 1e9,
                                                                 stepdist)
    keep = ~mask
    ray_id, step_id = ray_id[keep], step_id[keep]
    vec = torch.where(dd == 0, torch.full_like(dd, 1e-6), dd)
    t_min = torch.minimum((mx_t - od) / vec, (mn_t - od) / vec).amax(-1).clamp(min=0.2, max=1e9)
    interpx = t_min[ray_id] + (step_id * stepdist) / dd.norm(dim=-1)[ray_id]
    pts = od[ray_id] + dd[ray_id] * interpx[:, None]
    dens = oracle.dense_grid_forward(density.double(), pts, mn_t, mx_t)
    alpha = 1 - torch.exp(-torch.nn.functional.softplus(dens + shift) * interval)
    # per-ray exclusive cumprod (no early termination in this scene)
    logT = torch.log1p(-alpha)
    excl = torch.cumsum(logT, 0) - logT
    first = torch.cat([torch.tensor([True]), ray_id[1:] != ray_id[:-1]])
    idx = torch.arange(len(logT))
    seg_start = torch.cummax(torch.where(first, idx, torch.zeros_like(idx)), 0).values
    T = torch.exp(excl - excl[seg_start])
    w = T * alpha
    rgb = torch.sigmoid(oracle.dense_grid_forward(k0.double(), pts, mn_t, mx_t))
    last = torch.ones(64, dtype=torch.float64).index_reduce_(0, ray_id, 1 - alpha, "prod")
    marched = torch.zeros(64, 3, dtype=torch.float64).index_add_(0, ray_id, w[:, None] * rgb) + last[:, None] * 0.3
    assert float(last.min()) > 1e-3
    loss = torch.nn.functional.mse_loss(marched, target.double())
    loss.backward()
    want = torch.cat([od.grad, dd.grad], -1)
    scale = float(want.abs().max())
    close(grad_rays.double(), want, rtol=2e-3, atol=2e-4 * scale)


def test_tv_adam_pingpong_matches_reference_sequence(pp):
    """fused TV + masked Adam == total_variation_add_grad (old kernel) followed by masked_adam_upd."""
    G = 12
    density, k0 = synthetic.make_dvgo_grids(G, seed=7)
    mn, mx = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    o, d, _ = synthetic.make_rays(40, seed=3)
    target = torch.rand(40, 3, generator=torch.Generator().manual_seed(8))
    for dense in (False, True):
        eng = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, tv_density=0.3, tv_k0=0.2, tv_dense=dense,
                           lr_density=0.1, lr_k0=0.02)
        eng.backward_only(o.cuda(), d.cuda(), target.cuda())
        gd, gk = [t.cpu() for t in eng.export_grad()]
        pd, pk = density.clone(), k0.clone()
        oracle.total_variation_add_grad(pd, gd, 0.3, 0.3, 0.3, dense)
        oracle.total_variation_add_grad(pk, gk, 0.2, 0.2, 0.2, dense)
        for p, gr, lr in ((pd, gd, 0.1), (pk, gk, 0.02)):
            dvgo_cpu.masked_adam(p, gr, torch.zeros_like(p), torch.zeros_like(p), 1, 0.9, 0.99, lr, 1e-8)
        eng.optimizer_step()
        nd, nk = eng.export_grids()
        assert torch.equal(nd.cpu(), pd) and torch.equal(nk.cpu(), pk)     # same roundings as the reference kernels


def test_module_drop_in(pp):
    """DirectVoxGO keeps the reference's parameter names and result-dict keys."""
    m = pp.DirectVoxGO([-1, -1, -1], [1, 1, 1], num_voxels=20 ** 3, num_voxels_base=20 ** 3, alpha_init=1e-2).cuda()
    assert set(m.state_dict().keys()) >= {"density", "k0", "xyz_min", "xyz_max"}
    assert tuple(m.density.shape) == (1, 1, 20, 20, 20) and tuple(m.k0.shape) == (1, 3, 20, 20, 20)
    with torch.no_grad():
        m.density.normal_()
        m.k0.normal_()
    o, d, v = synthetic.make_rays(100)
    for fwd in (m.forward, m.forward_fused):
        m.zero_grad()
        out = fwd(o.cuda(), d.cuda(), v.cuda(), global_step=1, near=0.2, far=1e9, stepsize=0.5, bg=1.0)
        assert set(out) >= {"rgb_marched", "alphainv_cum", "depth", "disp"}
        torch.nn.functional.mse_loss(out["rgb_marched"], torch.rand(100, 3, device="cuda")).backward()
        assert float(m.density.grad.abs().sum()) > 0 and float(m.k0.grad.abs().sum()) > 0


def test_config_b_full_size_properties(pp):
    """160^3 grid, 8192 rays, stepsize 0.5 (BASELINE config B): size-independent invariants."""
    G, N = 160, 8192
    density, k0 = synthetic.make_dvgo_grids(G)
    o, d, _ = synthetic.make_rays(N)
    mn, mx = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    o_, d_ = o.cuda(), d.cuda()
    # (1) partition of unity: with saturated colour and bg = 1, sum(w) + T_last = 1 on every ray
    eng = pp.FusedDVGO(density.cuda() + 3.0, torch.full_like(k0, 40.0).cuda(), mn, mx, stepsize=0.5, bg=1.0)
    rgb, last, depth, cnt = eng.render(o_, d_, want_count=True)
    close(rgb, torch.ones(N, 3), rtol=0, atol=1e-4)
    assert bool((last >= 0).all()) and bool((last <= 1).all())
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= synthetic.dense_n_samples(G, 0.5)
    assert bool((last[cnt < cnt.max()] < 1.0).all())
    # (2) the train kernel's forward half is the render kernel
    eng = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, bg=0.0)
    target = torch.rand(N, 3, generator=torch.Generator().manual_seed(1)).cuda()
    rgb_r, *_ = eng.render(o_, d_)
    rgb_t = torch.empty(N, 3, device="cuda")
    eng.loss_buf.zero_()
    eng.backward_only(o_, d_, target, rgb_out=rgb_t)
    assert torch.equal(rgb_r, rgb_t)
    close(eng.loss_buf[1], torch.nn.functional.mse_loss(rgb_r, target), rtol=1e-4, atol=1e-7)
    # (3) directional derivative: <grad, delta> matches a central finite difference of the loss
    g = torch.Generator().manual_seed(2)
    delta = torch.randn(eng.vox.shape, generator=g).cuda()
    lin = float((eng.grad.double() * delta.double()).sum())
    eps = 1e-2
    vox0 = eng.vox.clone()

    def loss_at(v):
        eng.vox = v
        r, *_ = eng.render(o_, d_)
        return float(torch.nn.functional.mse_loss(r.double(), target.double()))
    fd = (loss_at(vox0 + eps * delta) - loss_at(vox0 - eps * delta)) / (2 * eps)
    eng.vox = vox0
    assert abs(fd - lin) <= 2e-2 * abs(lin) + 1e-7, (fd, lin)
    # (4) masked Adam: untouched voxels are bit-identical after the step, gradients are reset
    touched = (eng.grad != 0)
    before = eng.vox.clone()
    eng.optimizer_step()
    assert torch.equal(eng.vox[~touched], before[~touched])
    assert bool((eng.vox[touched] != before[touched]).float().mean() > 0.99)
    assert float(eng.grad.abs().max()) == 0.0


def test_config_b_scene_against_the_oracle_on_a_ray_subset(pp):
    """The REAL config-B scene (160^3 density + k0 grids of bench.py, stepsize 0.5) on a 256-ray subset of the bench
    batch against the composed-operator oracle: rendered outputs, early-termination indices (exact mode), the loss and
    the grid gradients of a train step - an oracle comparison at the benchmark's own size, not only invariants."""
    import os
    G, N = 160, 256
    density, k0 = synthetic.make_dvgo_grids(G)
    o, d, _ = synthetic.make_rays(8192)
    sel = torch.arange(0, 8192, 8192 // N)[:N]
    o, d = o[sel].contiguous(), d[sel].contiguous()
    mn, mx = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    kw = dict(near=0.2, far=1e9, stepdist=0.5 * 2.0 / G, act_shift=synthetic.act_shift(), interval=0.5)
    target = torch.rand(N, 3, generator=torch.Generator().manual_seed(11))
    dens_t, k0_t = density.clone().requires_grad_(True), k0.clone().requires_grad_(True)
    want = dvgo_cpu.render_composed(dens_t, k0_t, o, d, mn, mx, bg=0.0, **kw)
    loss, mse = dvgo_cpu.losses(want, target, 1.0, 0.0, 0.0)
    loss.backward()
    eng = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, bg=0.0)
    rgb, last, depth, cnt = eng.render(o.cuda(), d.cuda(), want_count=True)
    assert int((cnt.cpu() - want["n_marched"]).abs().max()) <= 1
    close(rgb, want["rgb_marched"].detach(), rtol=0, atol=2e-3)
    close(last, want["alphainv_last"].detach(), rtol=0, atol=1e-3)
    os.environ["PP_MARCH_FLAGS"] = "2"
    try:
        rgb_x, last_x, depth_x, cnt_x = eng.render(o.cuda(), d.cuda(), want_count=True)
    finally:
        del os.environ["PP_MARCH_FLAGS"]
    assert torch.equal(cnt_x.cpu(), want["n_marched"])
    close(rgb_x, want["rgb_marched"].detach(), rtol=1e-5, atol=2e-6)
    close(last_x, want["alphainv_last"].detach(), rtol=1e-5, atol=1e-7)
    close(depth_x, want["depth"].detach(), rtol=1e-5, atol=2e-6)
    eng.loss_buf.zero_()
    eng.backward_only(o.cuda(), d.cuda(), target.cuda())
    close(eng.loss_buf[0], loss.detach(), rtol=1e-4, atol=1e-7)
    gd, gk = eng.export_grad()
    for mine, ref in ((gd, dens_t.grad), (gk, k0_t.grad)):
        rel = float((mine.cpu() - ref).norm() / ref.norm())
        assert rel < 1e-4, rel
        close(mine, ref, rtol=1e-3, atol=1e-5 * float(ref.abs().max()))


def _torchrun(script, nproc, port, timeout=600):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", script)]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=root)


@pytest.mark.parametrize("script,port", [("dist_check.py", 29533), ("dist_check_voxurf.py", 29534)])
def test_two_ranks_equal_one_rank_on_the_concatenated_batch(pp, script, port):
    """N ranks with their own ray batches == ONE rank stepping on the concatenated batch (up to fp32 summation order),
    for the fused peer-memory exchange kernels (p2p, and NVLS where a multicast address exists), the NCCL formulation
    and the fine stage's flat all-reduce: tests/dist_check*.py under torchrun with 2 ranks.  Needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = _torchrun(script, 2, port)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " ok: world=2" in r.stdout, r.stdout[-3000:]


def test_fused_dvgo_reference_adam_mode(pp):
    """FusedDVGO(adam="reference"): the optimiser the reference runs (dense Python Adam, lib/utils.py:176-198) - every
    voxel moves each step (zero-gradient ones through their momentum), eps after the bias correction - against
    oracle.python_adam_step on the engine's own gradients, two steps."""
    G, N = 20, 256
    density, k0 = synthetic.make_dvgo_grids(G)
    o, d, _ = synthetic.make_rays(N, seed=5)
    tgt = torch.rand(N, 3, generator=torch.Generator().manual_seed(6))
    mn, mx = [-1.0, -1, -1], [1.0, 1, 1]
    eng = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, adam="reference", lr_density=0.1, lr_k0=0.05)
    pd, pk = density.clone(), k0.clone()
    st = [torch.zeros_like(pd), torch.zeros_like(pd), torch.zeros_like(pk), torch.zeros_like(pk)]
    for step in (1, 2):
        if step == 2:                     # other rays: some voxels of step 1 get no gradient now
            o, d, _ = synthetic.make_rays(N // 4, seed=7)
            tgt = tgt[:N // 4]
        eng.loss_buf.zero_()
        eng.backward_only(o.cuda(), d.cuda(), tgt.cuda())
        gd, gk = (t.cpu() for t in eng.export_grad())
        eng.optimizer_step()
        oracle.python_adam_step(pd, gd, st[0], st[1], step, 0.9, 0.99, 0.1, 1e-8)
        oracle.python_adam_step(pk, gk, st[2], st[3], step, 0.9, 0.99, 0.05, 1e-8)
        nd, nk = eng.export_grids()
        close(nd.cpu(), pd, rtol=1e-5, atol=1e-6)
        close(nk.cpu(), pk, rtol=1e-5, atol=1e-6)
    # a voxel no ray touched in step 2 still moved (momentum): the masked kernel would have left it alone
    untouched = (gd == 0) & (st[0] != 0)
    assert bool(untouched.any())
    with pytest.raises(NotImplementedError):
        pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, adam="reference", tv_density=0.1)


@pytest.mark.parametrize("adam", ["masked", "reference"])
def test_train_step_host_graph_equals_eager_steps(pp, adam):
    """FusedDVGO.train_step_host (pinned rays in, step, pinned loss out as ONE CUDA-graph launch, pp_dvgo_step_graph_*)
    against the same steps issued eagerly on a second engine: same losses, same parameters after four steps fed from
    DIFFERENT pinned buffers (the ray-copy node is re-pointed per launch, Adam's bias-corrected step sizes are patched
    per step).  Gradient sums are unordered fp32 atomics in both engines, so equality is to rounding, not to the bit."""
    G, N = 24, 512
    density, k0 = synthetic.make_dvgo_grids(G)
    mn, mx = [-1.0, -1, -1], [1.0, 1, 1]
    kw = dict(stepsize=0.5, adam=adam, lr_density=0.1, lr_k0=0.05)
    a = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, **kw)
    b = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, **kw)
    loss_host = torch.zeros(2).pin_memory()
    before = pp.lib.pp_launch_count()
    for step in range(4):
        o, d, _ = synthetic.make_rays(N, seed=20 + step)
        tgt = torch.rand(N, 3, generator=torch.Generator().manual_seed(30 + step))
        packed = torch.stack([o, d, tgt]).pin_memory()          # a fresh pinned buffer every step
        a.train_step_host(packed, loss_host)
        torch.cuda.current_stream().synchronize()
        want = b.train_step(o.cuda(), d.cuda(), tgt.cuda()).cpu()
        close(loss_host, want, rtol=1e-5, atol=1e-7)
        assert float(loss_host[0]) > 0
    assert a.step == b.step == 4 and len(a._step_graphs) == 1
    # replays are counted like eager launches (2 kernels per step on either engine)
    assert pp.lib.pp_launch_count() - before == 4 * 2 + 4 * 2
    for got, ref in zip(a.export_grids(), b.export_grids()):
        diff = (got - ref).abs().flatten()
        assert float(diff.quantile(0.999)) <= 1e-6, float(diff.quantile(0.999))
        assert float(diff.max()) <= 2 * 0.1 * 4       # a gradient within rounding of zero may flip one Adam step
    # the second batch size gets its own graph; unpinned input is rejected, nothing falls back to pageable copies
    o, d, _ = synthetic.make_rays(64, seed=40)
    small = torch.stack([o, d, torch.rand(64, 3)]).pin_memory()
    a.train_step_host(small, loss_host)
    torch.cuda.current_stream().synchronize()
    assert len(a._step_graphs) == 2 and torch.isfinite(loss_host).all()
    with pytest.raises(RuntimeError, match="pinned"):
        a.train_step_host(torch.stack([o, d, torch.rand(64, 3)]), loss_host)
    with pytest.raises(RuntimeError, match=r"\[3, N, 3\]"):
        a.train_step_host(torch.zeros(2, 64, 3).pin_memory(), loss_host)


def test_train_step_host_with_tv_runs_the_eager_sequence(pp):
    """TV on: the step is not graph-captured (the TV + Adam pass ping-pongs the parameter buffers); train_step_host
    issues copy-in, train_step, copy-out on the stream and gives train_step's numbers."""
    G, N = 20, 256
    density, k0 = synthetic.make_dvgo_grids(G)
    mn, mx = [-1.0, -1, -1], [1.0, 1, 1]
    a = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, tv_density=0.1, tv_k0=0.1)
    b = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, stepsize=0.5, tv_density=0.1, tv_k0=0.1)
    o, d, _ = synthetic.make_rays(N, seed=3)
    tgt = torch.rand(N, 3, generator=torch.Generator().manual_seed(4))
    loss_host = torch.zeros(2).pin_memory()
    a.train_step_host(torch.stack([o, d, tgt]).pin_memory(), loss_host)
    torch.cuda.current_stream().synchronize()
    close(loss_host, b.train_step(o.cuda(), d.cuda(), tgt.cuda()).cpu(), rtol=1e-5, atol=1e-7)
    assert len(a._step_graphs) == 0
    for got, ref in zip(a.export_grids(), b.export_grids()):
        assert float((got - ref).abs().flatten().quantile(0.999)) <= 1e-6


def test_host_pipeline_overlapped_steps_equal_eager_steps(pp):
    """FusedDVGO.host_pipeline: the copy of batch k+1 is queued on the copy stream while step k runs, the host reads
    the loss of step k while step k+1 runs.  Six steps with one batch of look-ahead must give the eager engine's losses
    in order and its parameters; the slot discipline (depth steps in flight at most, result() before reuse) is
    enforced."""
    G, N = 24, 512
    density, k0 = synthetic.make_dvgo_grids(G)
    mn, mx = [-1.0, -1, -1], [1.0, 1, 1]
    kw = dict(stepsize=0.5, lr_density=0.1, lr_k0=0.05)
    a = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, **kw)
    b = pp.FusedDVGO(density.cuda(), k0.cuda(), mn, mx, **kw)
    batches = []
    for step in range(6):
        o, d, _ = synthetic.make_rays(N, seed=50 + step)
        batches.append(torch.stack([o, d, torch.rand(N, 3, generator=torch.Generator().manual_seed(60 + step))]))
    pinned = [t.pin_memory() for t in batches]
    pipe = a.host_pipeline(N, depth=2)
    with pytest.raises(RuntimeError, match="nothing in flight"):
        pipe.result()
    got = []
    pipe.submit(pinned[0])
    for k in range(6):
        if k + 1 < 6:
            pipe.submit(pinned[k + 1])
        got.append(pipe.result())
    assert not pipe.pending and a.step == 6
    for k, blk in enumerate(batches):
        want = b.train_step(blk[0].cuda(), blk[1].cuda(), blk[2].cuda()).cpu()
        assert got[k][0] == pytest.approx(float(want[0]), rel=1e-5) and got[k][1] == pytest.approx(float(want[1]), rel=1e-5)
    for x, y in zip(a.export_grids(), b.export_grids()):
        assert float((x - y).abs().flatten().quantile(0.999)) <= 1e-6
    pipe.submit(pinned[0])
    pipe.submit(pinned[1])
    with pytest.raises(RuntimeError, match="in flight"):
        pipe.submit(pinned[2])
    assert len(pipe.drain()) == 2
    with pytest.raises(RuntimeError, match="pinned"):
        pipe.submit(batches[0])


def test_render_host_chunk_pipeline_equals_device_render(pp):
    """FusedDVGO.render_host (pinned rays in, pinned rgb / depth / alphainv_last out, chunk copies on copy streams under
    the neighbouring chunks' kernels) against one device-resident render of the same rays: bit-equal (the forward march
    has no atomics), for a ray count that is not a multiple of the chunk, repeated calls on the same streams, a
    larger chunk afterwards, and no rays at all."""
    G, N = 24, 1000
    density, k0 = synthetic.make_dvgo_grids(G)
    eng = pp.FusedDVGO(density.cuda() * 3 + 2, k0.cuda(), [-1.0, -1, -1], [1.0, 1, 1], stepsize=0.5)
    o, d, _ = synthetic.make_rays(N, seed=9)
    rgb, last, depth, _ = eng.render(o.cuda(), d.cuda(), want_depth=True)
    rays_host = torch.stack([o, d]).pin_memory()
    for chunk in (256, 256, 4096):
        rgb_h = torch.full((N, 3), -1.0).pin_memory()
        depth_h = torch.full((N,), -1.0).pin_memory()
        last_h = torch.full((N,), -1.0).pin_memory()
        eng.render_host(rays_host, rgb_h, depth_h, last_h, chunk=chunk)
        assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(depth_h, depth.cpu()) and torch.equal(last_h, last.cpu())
    rgb_only = torch.full((N, 3), -1.0).pin_memory()
    eng.render_host(rays_host, rgb_only, chunk=300)
    assert torch.equal(rgb_only, rgb.cpu())
    eng.render_host(torch.zeros(2, 0, 3).pin_memory(), torch.zeros(0, 3).pin_memory())
    with pytest.raises(RuntimeError, match="pinned"):
        eng.render_host(torch.stack([o, d]), rgb_only)
    with pytest.raises(RuntimeError, match="shape"):
        eng.render_host(rays_host, torch.zeros(N - 1, 3).pin_memory())
