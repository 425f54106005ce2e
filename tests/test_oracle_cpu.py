"""CPU tier: known-answer tests that pin the oracle on tiny hand-computable inputs, closed forms the
reference's Python code states (cumprod compositing, softplus form of raw2alpha, Python Adam), and
structural properties."""
import math

import pytest
import torch

from oracle import oracle
from tests import cases


def test_sampling_known_answers():
    # one axis-aligned ray through the unit box: enters at z=-1 (t=2), leaves at z=1 (t=4)
    o = torch.tensor([[0.0, 0.0, -3.0]])
    d = torch.tensor([[0.0, 0.0, 1.0]])
    mn, mx = cases.bbox()
    t_min, t_max = oracle.infer_t_minmax(o, d, mn, mx, 0.2, 1e9)
    assert t_min.item() == 2.0 and t_max.item() == 4.0
    # zero components are replaced by 1e-6: the x/y slabs become (-1e6, 1e6)
    n = oracle.infer_n_samples(d, t_min, t_max, 0.5)
    assert n.item() == 4
    pts, mask, ray_id, step_id, n_steps, *_ = oracle.sample_pts_on_rays(o, d, mn, mx, 0.2, 1e9, 0.5)
    assert pts[:, 2].tolist() == [-1.0, -0.5, 0.0, 0.5]
    assert mask.tolist() == [False] * 4          # strict comparisons: the face itself is inside
    assert ray_id.tolist() == [0] * 4 and step_id.tolist() == [0, 1, 2, 3]
    # a ray that misses still gets one (out-of-box) sample
    o2 = torch.tensor([[5.0, 5.0, 5.0]])
    d2 = torch.tensor([[1.0, 0.0, 0.0]])
    pts, mask, ray_id, step_id, n_steps, t_min, t_max, *_ = oracle.sample_pts_on_rays(o2, d2, mn, mx, 0.2, 1e9, 0.5)
    assert n_steps.tolist() == [1] and mask.tolist() == [True] and t_max.item() <= t_min.item()


def test_sampling_matches_dense_python_sampler():
    """Points of the enumerating sampler are the in-box prefix of the reference's dense sampler
    (lib/dvgo_ori.py:263-287, no jitter)."""
    o, d = cases.rays_with_edge_cases(64, seed=2)
    d = d / d.norm(dim=-1, keepdim=True)
    mn, mx = cases.bbox()
    stepdist = 0.03
    pts, mask, ray_id, step_id, n_steps, t_min, t_max, start, rdir = oracle.sample_pts_on_rays(
        o, d, mn, mx, 0.2, 1e9, stepdist)
    vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
    ra, rb = (mx - o) / vec, (mn - o) / vec
    tmin_py = torch.minimum(ra, rb).amax(-1).clamp(min=0.2, max=1e9)
    tmax_py = torch.maximum(ra, rb).amin(-1).clamp(min=0.2, max=1e9)
    assert torch.equal(t_min, tmin_py) and torch.equal(t_max, tmax_py)
    interp = t_min[ray_id] + stepdist * step_id / d.norm(dim=-1)[ray_id]
    pts_py = o[ray_id] + d[ray_id] * interp[:, None]
    torch.testing.assert_close(pts, pts_py, rtol=1e-5, atol=1e-5)


def test_raw2alpha_equals_softplus_form():
    """raw2alpha == 1 - exp(-softplus(d + shift) * interval)  (lib/dvgo_ori.py:242-244)."""
    d = torch.linspace(-12, 12, 997)
    shift, interval = -4.595, 0.5
    exp_d, alpha = oracle.raw2alpha(d, shift, interval)
    want = 1 - torch.exp(-torch.nn.functional.softplus(d.double() + shift) * interval)
    torch.testing.assert_close(alpha.double(), want, rtol=1e-4, atol=1e-6)
    # backward == autograd of the closed form
    dd = d.double().requires_grad_(True)
    a = 1 - torch.exp(-torch.nn.functional.softplus(dd + shift) * interval)
    gb = torch.cos(d)
    a.backward(gb.double())
    grad = oracle.raw2alpha_backward(exp_d, gb, interval)
    torch.testing.assert_close(grad.double(), dd.grad, rtol=1e-4, atol=1e-7)


def test_alpha2weight_known_answers():
    alpha = torch.tensor([0.5, 0.5, 0.5, 0.25])
    ray_id = torch.tensor([0, 0, 0, 2])
    w, T, last, i_s, i_e = oracle.alpha2weight(alpha, ray_id, 3)
    assert w.tolist() == [0.5, 0.25, 0.125, 0.25]
    assert T.tolist() == [1.0, 0.5, 0.25, 1.0]
    assert last.tolist() == [0.125, 1.0, 0.75]
    assert i_s.tolist() == [0, 0, 3] and i_e.tolist() == [3, 0, 4]     # ray 1 has no samples
    # early termination: T < 1e-3 after the 2nd sample -> i_end rewritten, tail keeps fill values
    alpha = torch.tensor([0.99, 0.99, 0.5, 0.5])
    w, T, last, i_s, i_e = oracle.alpha2weight(alpha, torch.zeros(4, dtype=torch.int64), 1)
    assert i_e.tolist() == [2]
    assert w[2:].tolist() == [0.0, 0.0] and T[2:].tolist() == [1.0, 1.0]
    assert last.item() == pytest.approx(1e-4, rel=1e-5)


def test_alpha2weight_equals_cumprod_compositing():
    """Without early termination: weights == alpha * cumprod_exclusive(1-alpha) (dvgo_ori.py:478-485)."""
    g = torch.Generator().manual_seed(0)
    alpha = torch.rand(20, 30, generator=g) * 0.03
    ray_id = torch.arange(20).repeat_interleave(30)
    w, T, last, _, i_e = oracle.alpha2weight(alpha.reshape(-1), ray_id, 20)
    wd, cum = oracle.get_ray_marching_ray(alpha)
    torch.testing.assert_close(w.reshape(20, 30), wd, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(last, cum[:, -1], rtol=1e-5, atol=1e-8)
    assert i_e.tolist() == list(range(30, 601, 30))


def test_alpha2weight_backward_equals_autograd():
    g = torch.Generator().manual_seed(1)
    alpha = torch.rand(11, 17, generator=g) * 0.05
    ray_id = torch.arange(11).repeat_interleave(17)
    fwd = oracle.alpha2weight(alpha.reshape(-1), ray_id, 11)
    gw = torch.randn(11 * 17, generator=g)
    gl = torch.randn(11, generator=g)
    grad = oracle.alpha2weight_backward(alpha.reshape(-1), *fwd, 11, gw, gl)
    a = alpha.double().requires_grad_(True)
    wd, cum = oracle.get_ray_marching_ray(a)
    ((wd.reshape(-1) * gw.double()).sum() + (cum[:, -1] * gl.double()).sum()).backward()
    torch.testing.assert_close(grad.double(), a.grad.reshape(-1), rtol=1e-4, atol=1e-6)


def test_total_variation_axis_quirks():
    """Old kernel: (k, j, i) weights = (wz, wy, wz); new kernel: (wx, wy, wz) -- SURVEY.md A14."""
    p = torch.zeros(1, 1, 3, 3, 3)
    p[0, 0, 1, 1, 1] = 0.6
    for masked in (False, True):
        grad = torch.zeros_like(p)
        mask = torch.ones_like(p) if masked else None
        oracle.total_variation_add_grad(p, grad, 6 * 1.0, 6 * 10.0, 6 * 100.0, True, mask)
        wk = 1.0 if masked else 100.0
        centre = 2 * 0.6 * (wk + 10.0 + 100.0)
        assert grad[0, 0, 1, 1, 1].item() == pytest.approx(centre, rel=1e-6)
        assert grad[0, 0, 1, 1, 0].item() == pytest.approx(-0.6 * wk, rel=1e-6)      # k neighbour
        assert grad[0, 0, 1, 0, 1].item() == pytest.approx(-0.6 * 10.0, rel=1e-6)    # j neighbour
        assert grad[0, 0, 0, 1, 1].item() == pytest.approx(-0.6 * 100.0, rel=1e-6)   # i neighbour
    # clamp to [-1, 1] and sparse mode
    p = torch.zeros(1, 1, 1, 1, 2)
    p[..., 1] = 5.0
    grad = torch.tensor([0.0, 1.0]).reshape(1, 1, 1, 1, 2)
    oracle.total_variation_add_grad(p, grad, 6.0, 6.0, 6.0, False)
    assert grad.flatten().tolist() == [0.0, 2.0]     # first voxel skipped (grad == 0), second += clamp(5)=1


def test_adam_kernel_vs_python_adam():
    """The kernel's update differs from lib/utils.py:176-198 only in where eps enters; with eps = 0 they
    agree, and masked mode leaves zero-grad elements untouched."""
    g = torch.Generator().manual_seed(2)
    p = torch.randn(100, generator=g)
    grad = torch.randn(100, generator=g)
    grad[::3] = 0
    m = torch.randn(100, generator=g) * 0.1
    v = torch.rand(100, generator=g) * 0.01 + 1e-4
    kp, km, kv = p.clone(), m.clone(), v.clone()
    oracle.adam_upd(0, kp, grad, km, kv, None, 5, 0.9, 0.99, 0.01, 0.0)
    pp_, pm, pv = p.clone().double(), m.clone().double(), v.clone().double()
    oracle.python_adam_step(pp_, grad.double(), pm, pv, 5, 0.9, 0.99, 0.01, 0.0)
    torch.testing.assert_close(kp.double(), pp_, rtol=1e-5, atol=1e-7)
    kp2, km2, kv2 = p.clone(), m.clone(), v.clone()
    oracle.adam_upd(1, kp2, grad, km2, kv2, None, 5, 0.9, 0.99, 0.01, 1e-8)
    assert torch.equal(kp2[::3], p[::3]) and torch.equal(km2[::3], m[::3]) and torch.equal(kv2[::3], v[::3])
    assert not torch.equal(kp2[1::3], p[1::3])


def test_cumdist_known_answer():
    dist = torch.tensor([[0.2, 0.2, 0.2, 0.2, 0.2, 0.2]])
    assert oracle.cumdist_thres(dist, 0.5).tolist() == [[False, False, True, False, False, True]]


def test_maskcache_known_answer():
    world = torch.zeros(2, 2, 2, dtype=torch.bool)
    world[1, 0, 1] = True
    scale = torch.ones(3)
    shift = torch.zeros(3)
    xyz = torch.tensor([[1.0, 0.0, 1.0], [0.6, 0.4, 0.5], [1.6, 0.0, 1.0], [-0.6, 0.0, 1.0]])
    # round-half-away: 0.5 -> 1; 1.6 -> 2 is out of range; -0.6 -> -1 out of range
    assert oracle.maskcache_lookup(world, xyz, scale, shift).tolist() == [True, True, False, False]


def test_segment_ops_and_neus_alpha_reference_forms():
    src = torch.arange(6, dtype=torch.float32)
    rid = torch.tensor([0, 0, 2, 2, 2, 3])
    assert oracle.segment_sum(src[:, None], rid, 5).flatten().tolist() == [1.0, 0.0, 9.0, 5.0, 0.0]
    assert oracle.segment_cumsum(src, rid).tolist() == [0.0, 1.0, 2.0, 5.0, 9.0, 5.0]
    # a surface crossed head-on: alpha rises from ~0 to a peak around sdf = 0
    sdf = torch.linspace(0.1, -0.1, 9)
    grads = torch.tensor([[0.0, 0.0, 1.0]]).repeat(9, 1)
    view = torch.tensor([[0.0, 0.0, -1.0]])
    a = oracle.neus_alpha(view, torch.zeros(9, dtype=torch.int64), 0.025, sdf, grads, 64.0)
    assert float(a[0]) < 0.05 and float(a[-1]) > 0.5 and bool((a[1:] >= a[:-1]).all())   # opacity rises through the surface
    assert bool((a >= 0).all()) and bool((a <= 1).all())


# ---- randomised structural properties (hypothesis): ragged, empty and degenerate inputs -----------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=40), min_size=1, max_size=12), st.integers(0, 2 ** 31 - 1))
def test_alpha2weight_properties_on_ragged_segments(lens, seed):
    """Any segmentation (including empty rays and an empty input): weights + final transmittance partition unity while
    the ray is not terminated, weights are zero past i_end, i_start / i_end delimit the ray's segment, and the
    backward of a constant upstream gradient equals autograd of the cumprod form."""
    g = torch.Generator().manual_seed(seed)
    n_rays = len(lens)
    ray_id = torch.repeat_interleave(torch.arange(n_rays), torch.tensor(lens))
    alpha = torch.rand(ray_id.numel(), generator=g) * 0.9
    w, T, last, i_start, i_end = oracle.alpha2weight(alpha, ray_id, n_rays)
    assert w.shape == alpha.shape and last.shape == (n_rays,)
    off = 0
    for r, n in enumerate(lens):
        if n == 0:
            assert last[r].item() == 1.0
            continue
        assert i_start[r].item() == off and off < i_end[r].item() <= off + n
        seg = slice(off, off + n)
        stop = i_end[r].item() - off
        assert float(w[seg][stop:].abs().sum()) == 0.0                     # nothing is composited past the stop
        total = float(w[seg].double().sum() + last[r].double())
        assert abs(total - 1.0) < 1e-5
        ref_T = torch.cumprod(torch.cat([torch.ones(1), 1 - alpha[seg][:stop]]), 0)
        torch.testing.assert_close(w[seg][:stop], (ref_T[:-1] * alpha[seg][:stop]), rtol=1e-5, atol=1e-7)
        off += n


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 40), st.floats(0.01, 0.2), st.integers(0, 2 ** 31 - 1))
def test_sampling_enumeration_is_a_run_length_code(n_rays, stepdist, seed):
    """ray_id is sorted, step_id restarts at 0 on every ray and counts up by one, the per-ray counts are N_steps, every
    ray (even one that misses the box) owns at least one sample, and in-box flags agree with the point coordinates."""
    o, d = cases.rays_with_edge_cases(n_rays, seed=seed % 1000)
    mn, mx = cases.bbox()
    pts, mask, ray_id, step_id, n_steps, t_min, t_max, start, rdir = oracle.sample_pts_on_rays(
        o, d, mn, mx, 0.2, 1e9, float(stepdist))
    assert int(n_steps.min()) >= 1 and int(n_steps.sum()) == ray_id.numel()
    assert bool((ray_id[1:] >= ray_id[:-1]).all())
    counts = torch.zeros(n_rays, dtype=torch.long).index_add_(0, ray_id, torch.ones_like(ray_id))
    assert torch.equal(counts, n_steps)
    starts = torch.cumsum(n_steps, 0) - n_steps
    assert torch.equal(step_id, torch.arange(ray_id.numel()) - starts[ray_id])
    outside = ((mn > pts) | (pts > mx)).any(-1)
    assert torch.equal(outside, mask)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 6), st.integers(1, 50), st.floats(0.05, 2.0), st.integers(0, 2 ** 31 - 1))
def test_cumdist_thres_is_a_running_sum_with_reset(n, s, thres, seed):
    """ub360_utils_kernel.cu:12-47: mask[i] is set where the fp32 running sum of dist since the last set point
    exceeds thres, and the sum restarts there."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    dist = torch.rand(n, s, generator=g)
    mask = oracle.cumdist_thres(dist, float(thres))
    th = np.float32(thres)
    for r in range(n):
        acc, want = np.float32(0.0), []
        for i in range(s):
            acc = np.float32(acc + np.float32(dist[r, i].item()))
            over = bool(acc > th)
            want.append(over)
            if over:
                acc = np.float32(0.0)
        assert mask[r].tolist() == want


def test_scene_oracle_structural_properties():
    """oracle/bgnerf_cpu.py (NeRF.composite / sample_depth_from_pdf restated): weights are a sub-probability whose sum
    is the opacity; with positive densities the 1e10 last interval makes every ray opaque; the exclusive transmittance
    is non-increasing; resampled depths stay inside the range, are sorted for a sorted grid, and concentrate where the
    weights are."""
    import torch
    from oracle import bgnerf_cpu
    g = torch.Generator().manual_seed(3)
    B, R, N = 2, 17, 40
    ray = torch.randn(B, R, 3, generator=g)
    edges = torch.linspace(0.5, 4.0, N + 1)
    depth = (edges[:-1] + (edges[1:] - edges[:-1]) * torch.rand(B, R, N, generator=g))[..., None]
    rgb = torch.rand(B, R, N, 3, generator=g)
    dens = torch.rand(B, R, N, generator=g) * 2 + 1e-3
    out = bgnerf_cpu.composite(ray, rgb, dens, depth)
    w = out["weights"][..., 0]
    assert float(w.min()) >= 0
    torch.testing.assert_close(w.sum(-1, keepdim=True), out["opacity"])
    torch.testing.assert_close(out["opacity"], torch.ones_like(out["opacity"]), rtol=0, atol=1e-5)
    assert float(out["rgb"].min()) >= 0 and float(out["rgb"].max()) <= 1 + 1e-5
    assert float(out["depth"].min()) >= 0.5 - 1e-5
    # zero density everywhere except the last sample: the whole weight sits on the last sample
    d0 = torch.zeros(B, R, N)
    d0[..., -1] = 1.0
    o0 = bgnerf_cpu.composite(ray, rgb, d0, depth)
    torch.testing.assert_close(o0["weights"][..., -1, 0], torch.ones(B, R))
    torch.testing.assert_close(o0["rgb"], rgb[:, :, -1])
    # resampling: a single peaked bin attracts all fine samples into that bin's depth interval
    peak = torch.zeros(1, 1, N)
    peak[..., 7] = 1.0
    fine, idx = bgnerf_cpu.sample_depth_from_pdf(peak, N, 16, [0.5, 4.0], torch.linspace(0, 1, 17))
    lo, hi = 0.5 + 3.5 * 7 / N, 0.5 + 3.5 * 8 / N
    assert float(fine.min()) >= lo - 1e-5 and float(fine.max()) <= hi + 1e-5
    assert bool((fine[0, 0, 1:, 0] >= fine[0, 0, :-1, 0]).all())
    fine2, _ = bgnerf_cpu.sample_depth_from_pdf(w, N, 32, [0.5, 4.0], torch.linspace(0, 1, 33))
    assert float(fine2.min()) >= 0.5 - 1e-5 and float(fine2.max()) <= 4.0 + 1e-5


def test_dense_compositing_oracle_properties():
    """oracle.get_ray_marching_ray (lib/dvgo_ori.py:478-485): weights + final transmittance partition unity, the
    clamp keeps a fully opaque sample from zeroing the tail exactly (1e-10), alpha = 0 samples get zero weight."""
    import torch
    from oracle import oracle
    g = torch.Generator().manual_seed(4)
    a = torch.rand(9, 33, generator=g) ** 2
    a[:, 5] = 0.0
    w, c = oracle.get_ray_marching_ray(a)
    assert c.shape == (9, 34) and bool((c[:, 0] == 1).all())
    torch.testing.assert_close(w.sum(-1) + c[:, -1], torch.ones(9), rtol=0, atol=1e-5)
    assert bool((w[:, 5] == 0).all()) and bool((c[:, 1:] <= c[:, :-1] + 1e-7).all())
    a[:, 10] = 1.0
    w, c = oracle.get_ray_marching_ray(a)
    assert float(c[:, 11:].max()) <= 1.0001e-10 and float(c[:, 11].min()) > 0
