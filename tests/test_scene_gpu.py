"""Scene-branch compositing on the GPU (csrc/scene.cu through poseprobe_b200.scene) against
  * tests/golden/bgnerf_composite.npz = the reference's OWN NeRF.composite / Graph.sample_depth_from_pdf run on the CPU
    (frequency_nerf.py:283-343, renderer.py:703-738; tests/golden/make_golden_bgnerf.py), and
  * oracle/bgnerf_cpu.py (itself pinned to that golden) on fresh seeded inputs, incl. ragged N (not a multiple of 32),
    N = 2 (the reference indexes T[:, :, -2]: N >= 2), zero densities, empty batches.
Tolerance: 1e-5 relative on fp32 weights / rgb / gradients (the north star's bar), written at each assert."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "bgnerf_composite.npz")


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


# rgb_var = sum_n w_n sum_c (rgb_nc - rgb_c) is a cancellation: it is (1 - opacity) * sum_c rgb_c up to rounding, i.e. ~0 for
# an opaque ray, and what is left is the fp32 summation noise of N terms of size w_n (a diagnostic image in the
# reference, training/base.py:648) - compared absolutely at the size of that noise
_ATOL = {"rgb_var": 2e-5}


def _opt(opaque):
    return types.SimpleNamespace(nerf=types.SimpleNamespace(setbg_opaque=opaque), mask_img=False)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_composite_replays_the_references_golden(cuda, tag):
    from poseprobe_b200 import scene
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g["%s_%s" % (tag, k)])  # noqa: E731
    ray = t("ray").to(cuda).requires_grad_(True)
    rgb_s = t("rgb_samples").to(cuda).requires_grad_(True)
    dens = t("density_samples").to(cuda).requires_grad_(True)
    depth = t("depth_samples").to(cuda)
    pred = scene.composite(_opt(bool(g[tag + "_opaque"])), ray, {"rgb_samples": rgb_s, "density_samples": dens}, depth)
    for k in ("rgb", "rgb_var", "depth", "depth_var", "opacity", "weights", "all_cumulated"):
        assert pred[k].shape == t("out_" + k).shape, k
        torch.testing.assert_close(pred[k].detach().cpu(), t("out_" + k), rtol=1e-5, atol=_ATOL.get(k, 2e-6),
                                   msg=lambda m, k=k: k + m)
    loss = (pred["rgb"] * t("w_rgb").to(cuda)).sum() + (pred["depth"] * t("w_depth").to(cuda)).sum() + \
        (pred["opacity"] * t("w_opacity").to(cuda)).sum() + (pred["weights"] * t("w_weights").to(cuda)).sum()
    grads = torch.autograd.grad(loss, [rgb_s, dens, ray])
    for got, k in zip(grads, ("g_rgb_samples", "g_density_samples", "g_ray")):
        want = t(k)
        rel = float((got.cpu() - want).norm() / want.norm())
        assert rel < 1e-5, (k, rel)
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sample_depth_from_pdf_replays_the_references_golden(cuda, tag):
    from poseprobe_b200 import scene
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g["%s_%s" % (tag, k)])  # noqa: E731
    w = t("out_weights")[..., 0].to(cuda)
    N = w.shape[-1]
    rng = [float(v) for v in g[tag + "_depth_range"]]
    Nf = g[tag + "_fine_det"].shape[2]
    det = scene.sample_depth_from_pdf(w, N, Nf, rng, det=True).cpu()
    rnd = scene.sample_depth_from_pdf(w, N, Nf, rng, det=False, grid=t("fine_grid")).cpu()
    for got, want in ((det, t("fine_det")), (rnd, t("fine_rand"))):
        assert got.shape == want.shape
        # the cdf is a float cumsum: a mid point within one ulp of a cdf entry may land in the neighbouring bin on a
        # different summation order, which moves that one sample by < one bin; everything else agrees to 1e-5
        err = (got - want).abs()
        assert float((err < 1e-5 * 4.0).float().mean()) > 0.999, float(err.max())
        assert float(err.max()) < (rng[1] - rng[0]) / N


@pytest.mark.parametrize("shape", [(1, 1, 2), (2, 5, 3), (1, 33, 31), (3, 70, 97), (1, 1000, 128), (2, 300, 256)])
@pytest.mark.parametrize("opaque", [False, True])
def test_composite_vs_oracle_ragged(cuda, shape, opaque):
    from oracle import bgnerf_cpu
    from poseprobe_b200 import scene
    B, R, N = shape
    gen = torch.Generator().manual_seed(B * 1000 + R * 10 + N)
    ray = torch.randn(B, R, 3, generator=gen)
    edges = torch.linspace(0.3, 5.0, N + 1)
    depth = (edges[:-1] + (edges[1:] - edges[:-1]) * torch.rand(B, R, N, generator=gen))[..., None]
    rgb_s = torch.rand(B, R, N, 3, generator=gen)
    dens = torch.nn.functional.softplus(torch.randn(B, R, N, generator=gen) * 3 - 2) * (
        torch.rand(B, R, N, generator=gen) > 0.2)
    wts = [torch.randn(B, R, 3, generator=gen), torch.randn(B, R, 1, generator=gen), torch.randn(B, R, 1, generator=gen),
           torch.randn(B, R, N, 1, generator=gen) * 0.1]

    def run(fn, dev, dt):
        a, b, c = (x.to(dev, dt).requires_grad_(True) for x in (ray, rgb_s, dens))
        out = fn(a, b, c, depth.to(dev, dt))
        loss = sum((out[k] * w.to(dev, dt)).sum() for k, w in zip(("rgb", "depth", "opacity", "weights"), wts))
        return out, torch.autograd.grad(loss, [b, c, a])

    want, gw = run(lambda a, b, c, d: bgnerf_cpu.composite(a, b, c, d, setbg_opaque=opaque), "cpu", torch.float64)
    got, gg = run(lambda a, b, c, d: scene.composite(_opt(opaque), a, {"rgb_samples": b, "density_samples": c}, d), cuda,
                  torch.float32)
    for k in ("rgb", "rgb_var", "depth", "depth_var", "opacity", "weights", "all_cumulated"):
        torch.testing.assert_close(got[k].detach().cpu().double(), want[k].detach(), rtol=1e-5, atol=_ATOL.get(k, 2e-6),
                                   msg=lambda m, k=k: k + m)
    for a, b, k in zip(gg, gw, ("rgb_samples", "density_samples", "ray")):
        rel = float((a.cpu().double() - b).norm() / (b.norm() + 1e-30))
        assert rel < 1e-5, (k, rel)


def test_composite_empty_and_errors(cuda):
    from poseprobe_b200 import scene
    z = lambda *s: torch.zeros(*s, device=cuda)  # noqa: E731
    out = scene.composite(_opt(False), z(1, 0, 3), {"rgb_samples": z(1, 0, 8, 3), "density_samples": z(1, 0, 8)},
                          z(1, 0, 8, 1))
    assert out["rgb"].shape == (1, 0, 3) and out["weights"].shape == (1, 0, 8, 1)
    with pytest.raises(RuntimeError):
        scene.composite(_opt(False), torch.zeros(1, 2, 3), {"rgb_samples": torch.zeros(1, 2, 8, 3),
                                                            "density_samples": torch.zeros(1, 2, 8)},
                        torch.zeros(1, 2, 8, 1))


def test_composite_full_size_properties(cuda):
    """The reference's training shape (2048 rays x 128 coarse + 128 fine samples... here 65536 rays of BASELINE
    configs[4] x 256 samples): weights are a sub-probability (sum = opacity <= 1), opacity + T_end = 1 through the
    1e10 last interval whenever the last density is positive, rgb is a convex combination, and the resampled depths
    are sorted and inside the range."""
    from poseprobe_b200 import scene
    B, R, N = 1, 65536, 256
    gen = torch.Generator(device=cuda).manual_seed(5)
    ray = torch.randn(B, R, 3, generator=gen, device=cuda)
    depth = scene.sample_depth(N, [0.5, 6.0], B, R, cuda, stratified=True,
                               rand=torch.rand(B, R, N, 1, generator=gen, device=cuda))
    rgb_s = torch.rand(B, R, N, 3, generator=gen, device=cuda)
    dens = torch.nn.functional.softplus(torch.randn(B, R, N, generator=gen, device=cuda) * 2 - 3) + 1e-3
    out = scene.composite(_opt(False), ray, {"rgb_samples": rgb_s, "density_samples": dens}, depth)
    w = out["weights"][..., 0]
    assert float(w.min()) >= 0
    torch.testing.assert_close(w.sum(-1, keepdim=True), out["opacity"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["opacity"], torch.ones_like(out["opacity"]), rtol=0, atol=2e-5)
    assert float(out["rgb"].min()) >= 0 and float(out["rgb"].max()) <= 1 + 1e-5
    fine = scene.sample_depth_from_pdf(w, N, 128, [0.5, 6.0], det=True)
    assert float(fine.min()) >= 0.5 - 1e-5 and float(fine.max()) <= 6.0 + 1e-5
    assert bool((fine[:, :, 1:, 0] >= fine[:, :, :-1, 0] - 1e-6).all())


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (2, 7, 64, 48), (1, 100, 128, 128), (1, 33, 500, 524), (3, 5, 17, 0)])
def test_merge_depth_samples_is_cat_sort(cuda, shape):
    """pp_nerf_sort_depth == torch.cat(...).sort(dim=2).values (renderer.py:612-614): exact (a sort moves values)."""
    from poseprobe_b200 import scene
    B, R, N, Nf = shape
    gen = torch.Generator().manual_seed(N * 7 + Nf)
    a = torch.rand(B, R, N, 1, generator=gen) * 5
    b = torch.rand(B, R, Nf, 1, generator=gen) * 5
    if Nf > 3:
        b[..., :3, :] = a[..., :3, :]                      # ties
    want = torch.cat([a, b], dim=2).sort(dim=2).values
    got = scene.merge_depth_samples(a.to(cuda), b.to(cuda)).cpu()
    assert got.shape == want.shape and torch.equal(got, want)
