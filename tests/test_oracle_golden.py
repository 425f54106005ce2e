"""CPU tier: replays the golden vectors in tests/golden/ (outputs of the REFERENCE's own kernels,
generated on a B200 by tests/golden/make_golden.py from oracle/_ref) against the C oracle.
Bars: integers / bools / indices bit-exact; fp32 bit-exact wherever the path has no transcendental
(libm vs libdevice exp/pow differ by <= 2 ulp), rtol 1e-5 otherwise."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden vectors %s not generated yet" % name)
    z = np.load(path)
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def exact(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    assert torch.equal(a, b), "first mismatches: %s" % (torch.nonzero(a != b)[:5].tolist(),)


def test_golden_sampling():
    z = load("sampling")
    near, far, stepdist = float(z["near"]), float(z["far"]), float(z["stepdist"])
    got = oracle.sample_pts_on_rays(z["rays_o"], z["rays_d"], z["xyz_min"], z["xyz_max"], near, far, stepdist)
    names = ["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max", "rays_start", "rays_dir"]
    for name, g in zip(names, got):
        exact(g, z[name])
    tm = oracle.infer_t_minmax(z["rays_o"], z["rays_d"], z["xyz_min"], z["xyz_max"], near, far)
    exact(oracle.infer_n_samples(z["rays_d"], tm[0], tm[1], stepdist), z["infer_n_samples"])
    p, m = oracle.sample_ndc_pts_on_rays(z["rays_o"], z["rays_d"], z["xyz_min"], z["xyz_max"], int(z["ndc_S"]))
    exact(p, z["ndc_pts"])
    exact(m, z["ndc_mask"])
    bg = oracle.sample_bg_pts_on_rays(z["rays_o"], z["rays_d"], z["bg_t_max"], float(z["bg_preserve"]), int(z["ndc_S"]))
    exact(bg, z["bg_pts"])


def test_golden_maskcache():
    z = load("maskcache")
    exact(oracle.maskcache_lookup(z["world"], z["xyz"], z["scale"], z["shift"]), z["out"])


def test_golden_raw2alpha():
    z = load("raw2alpha")
    shift, interval = float(z["shift"]), float(z["interval"])
    exp_d, alpha = oracle.raw2alpha(z["density"], shift, interval)
    torch.testing.assert_close(exp_d, z["exp_d"], rtol=1e-5, atol=0)
    torch.testing.assert_close(alpha, z["alpha"], rtol=1e-5, atol=3e-7)
    # backward from the reference's own exp_d: only pow differs
    grad = oracle.raw2alpha_backward(z["exp_d"], z["grad_back"], interval)
    torch.testing.assert_close(grad, z["grad"], rtol=1e-5, atol=1e-30)
    _, alpha_n = oracle.raw2alpha_nonuni(z["density"], shift, z["interval_pts"])
    torch.testing.assert_close(alpha_n, z["alpha_nonuni"], rtol=1e-5, atol=3e-7)
    grad_n = oracle.raw2alpha_nonuni_backward(z["exp_d"], z["grad_back"], z["interval_pts"])
    torch.testing.assert_close(grad_n, z["grad_nonuni"], rtol=1e-5, atol=1e-30)


def test_golden_alpha2weight():
    z = load("alpha2weight")
    n_rays = int(z["n_rays"])
    got = oracle.alpha2weight(z["alpha"], z["ray_id"], n_rays)
    for name, g in zip(["weight", "T", "alphainv_last", "i_start", "i_end"], got):
        exact(g, z[name])
    grad = oracle.alpha2weight_backward(z["alpha"], z["weight"], z["T"], z["alphainv_last"], z["i_start"], z["i_end"],
                                        n_rays, z["grad_weights"], z["grad_last"])
    exact(grad, z["grad"])


def test_golden_total_variation():
    z = load("total_variation")
    wx, wy, wz = [float(v) for v in z["w"]]
    for dense in (0, 1):
        g = z["grad"].clone()
        oracle.total_variation_add_grad(z["param"], g, wx, wy, wz, bool(dense), None)
        exact(g, z["old_dense%d" % dense])
        g = z["grad"].clone()
        oracle.total_variation_add_grad(z["param"], g, wx, wy, wz, bool(dense), z["mask"])
        exact(g, z["new_dense%d" % dense])


def test_golden_adam():
    z = load("adam")
    step, b1, b2, lr, eps = z["hp"].tolist()
    for mode in range(3):
        p, m, v = z["param"].clone(), z["exp_avg"].clone(), z["exp_avg_sq"].clone()
        oracle.adam_upd(mode, p, z["grad"], m, v, z["perlr"] if mode == 2 else None, int(step), b1, b2, lr, eps)
        exact(m, z["exp_avg_%d" % mode])
        exact(v, z["exp_avg_sq_%d" % mode])
        exact(p, z["param_%d" % mode])


def test_golden_cumdist():
    z = load("cumdist")
    exact(oracle.cumdist_thres(z["dist"], float(z["thres"])), z["mask"])
