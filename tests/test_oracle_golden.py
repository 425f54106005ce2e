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


def test_voxurf_oracle_matches_the_references_own_forward():
    """oracle/voxurf_cpu.py against tests/golden/voxurf_forward.npz = outputs of the reference's OWN `Voxurf.forward`
    (use_deform=False), `neus_sdf_gradient`, `neus_alpha_from_sdf_scatter`, `positional_encoding_barf`,
    `sample_ray_ori` and `total_variation`, executed unmodified on the CPU by tests/golden/make_golden_voxurf.py
    (only torch_scatter.segment_coo and the CUDA alpha2weight are replaced).  The reference samples with the dense
    torch sampler (`o + d * (t_min + step / |d|)`), the oracle with the enumerating sampler's arithmetic
    (`start + dir * dist`): positions differ by fp32 rounding, hence 2e-5 on per-ray outputs."""
    import os
    import numpy as np
    import torch
    import torch.nn as nn
    from oracle import oracle, voxurf_cpu
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_forward.npz"))
    sdf, k0 = torch.from_numpy(g["sdf"]), torch.from_numpy(g["k0"].astype(np.float32))
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 3))
    with torch.no_grad():
        for i, lin in enumerate([net[0], net[2], net[4]]):
            lin.weight.copy_(torch.from_numpy(g["w%d" % i]))
            lin.bias.copy_(torch.from_numpy(g["b%d" % i]))
    o, d, v = (torch.from_numpy(g[k]) for k in ("rays_o", "rays_d", "viewdirs"))
    vs = float(g["voxel_size"])
    cases = (("eval", None, 1.0), ("train", torch.from_numpy(g["train_jitter"]), float(g["gstep"]) / float(g["N_iters"])))
    for tag, jitter, progress in cases:
        with torch.no_grad():
            out = voxurf_cpu.forward(sdf, k0, net, o, d, v, [-1, -1, -1], [1, 1, 1], 0.5, vs, float(g[tag + "_s_val"]),
                                     bg=0.3, sdf_alpha=10.0, sdf_beta=0.35, posbase_pe=5, viewbase_pe=1,
                                     barf_c2f=(0.1, 0.5), progress=progress, jitter=jitter)
        for key in ("rgb_marched", "alphainv_cum", "cum_weights"):
            torch.testing.assert_close(out[key], torch.from_numpy(g["%s_%s" % (tag, key)]), rtol=0, atol=2e-5,
                                       msg=lambda m, key=key, tag=tag: "%s %s: %s" % (tag, key, m))
        # per-sample outputs: same in-box sample set (a point within rounding of a face may differ: not in this scene)
        assert out["weights"].shape[0] == g[tag + "_weights"].shape[0]
        torch.testing.assert_close(out["raw_alpha"], torch.from_numpy(g[tag + "_raw_alpha"]), rtol=0, atol=2e-5)
        torch.testing.assert_close(out["weights"], torch.from_numpy(g[tag + "_weights"]), rtol=0, atol=2e-5)
        torch.testing.assert_close(out["raw_rgb"], torch.from_numpy(g[tag + "_raw_rgb"]), rtol=0, atol=2e-5)
        torch.testing.assert_close(out["gradient"], torch.from_numpy(g[tag + "_gradient"]), rtol=1e-4, atol=2e-4)
    # Voxurf.grid_sampler: F.grid_sample(border) and the custom 8-gather `grid_sample_3d` (:545-659) vs the oracle's
    # trilinear with border padding; the custom sampler extrapolates its weights outside the box, so it is compared
    # on the in-box points only (the reference only feeds it in-box samples)
    gs_grid, gs_xyz = torch.from_numpy(g["gs_grid"]), torch.from_numpy(g["gs_xyz"])
    mn, mx = torch.tensor([-1.0, -1, -1]), torch.tensor([1.0, 1, 1])
    got = oracle.dense_grid_forward(gs_grid, gs_xyz, mn, mx, padding_mode="border")
    torch.testing.assert_close(got, torch.from_numpy(g["gs_border"]), rtol=1e-6, atol=1e-6)
    inside = ((gs_xyz >= -1) & (gs_xyz <= 1)).all(-1)
    assert int(inside.sum()) > 200 and int((~inside).sum()) > 20
    torch.testing.assert_close(got[inside], torch.from_numpy(g["gs_custom"])[inside], rtol=1e-5, atol=1e-5)
    # the s_val schedule of neus_alpha_from_sdf_scatter (:487)
    assert abs(float(g["train_s_val"]) - 1.0 / (float(g["gstep"]) + 50 / 0.2) * 50) < 1e-7
    # total_variation (:1298-1313)
    tv_v, tv_mask = torch.from_numpy(g["tv_v"]), torch.from_numpy(g["tv_mask"])
    assert abs(float(oracle.total_variation(tv_v)) - float(g["tv_plain"])) < 1e-6
    assert abs(float(oracle.total_variation(tv_v, tv_mask)) - float(g["tv_masked"])) < 1e-6
    # k0_total_variation() of the forward (:1067) is total_variation(k0.grid)
    assert abs(float(oracle.total_variation(k0)) - float(g["eval_k0_tv"])) < 1e-6


def test_dense_baseline_matches_the_references_own_dvgo_and_adam():
    """oracle/dvgo_cpu.DenseBaseline (bench.py's CPU baseline / reference arm) against tests/golden/dvgo_dense.npz =
    the reference's OWN `DirectVoxGO.forward` (lib/dvgo_ori.py:289-379) and optimiser (`Adam`, lib/utils.py:53-198),
    executed unmodified on the CPU by tests/golden/make_golden_dvgo.py: render, then three optimisation steps."""
    import os
    import numpy as np
    import torch
    from oracle import dvgo_cpu
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dvgo_dense.npz"))
    base = dvgo_cpu.DenseBaseline(torch.from_numpy(g["density0"]), torch.from_numpy(g["k00"]), [-1, -1, -1], [1, 1, 1],
                                  0.5, alpha_init=1e-2, near=0.2, far=1e9, bg=0.3, lr=0.1, betas=(0.9, 0.99))
    assert abs(base.voxel_size - float(g["voxel_size"])) < 1e-7
    o, d, t = (torch.from_numpy(g[k]) for k in ("rays_o", "rays_d", "target"))
    with torch.no_grad():
        rgb, last, n_inbox = base.forward(o, d)
    assert n_inbox == int(g["n_inbox"])
    torch.testing.assert_close(rgb, torch.from_numpy(g["rgb_marched"]), rtol=0, atol=1e-6)
    torch.testing.assert_close(last, torch.from_numpy(g["alphainv_last"]), rtol=0, atol=1e-6)
    losses = [base.train_step(o, d, t)[0] for _ in range(3)]
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
    torch.testing.assert_close(base.density.detach(), torch.from_numpy(g["density3"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(base.k0.detach(), torch.from_numpy(g["k03"]), rtol=1e-4, atol=1e-5)


def test_pose_chain_and_rays_match_the_references_own_code():
    """oracle.se3_to_SE3 / get_rays_of_a_view AND the product's host-side pose algebra (poseprobe_b200/camera.py, pure
    torch, runs on the CPU) against tests/golden/camera.npz = the reference's own `Lie.se3_to_SE3`,
    `Pose.compose_pair`, `Pose.invert` (lib/camera.py) and `get_rays_of_a_view` (lib/voxurf_coarse.py:1339-1407)
    executed unmodified by tests/golden/make_golden_camera.py (incl. theta = 0 and theta ~ 1e-4 twists)."""
    import os
    import numpy as np
    import torch
    from oracle import oracle
    from poseprobe_b200 import camera
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera.npz"))
    wu = torch.from_numpy(g["wu"])
    for impl in (oracle.se3_to_SE3, camera.se3_to_SE3):
        torch.testing.assert_close(impl(wu), torch.from_numpy(g["SE3"]), rtol=1e-6, atol=1e-7)
    composed = camera.compose_pair(torch.from_numpy(g["SE3"]), torch.from_numpy(g["pnp"]))
    torch.testing.assert_close(composed, torch.from_numpy(g["composed"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(camera.invert(composed), torch.from_numpy(g["c2w"]), rtol=1e-6, atol=1e-6)
    H, W = [int(x) for x in g["HW"]]
    K, c2w = torch.from_numpy(g["K"]), torch.from_numpy(g["c2w"])[2]
    for inverse_y in (False, True):
        tag = "iy%d" % int(inverse_y)
        ro, rd, vd = oracle.get_rays_of_a_view(H, W, K, c2w, inverse_y=inverse_y, flip_y=inverse_y)
        assert torch.equal(ro, torch.from_numpy(g["rays_o_" + tag]))
        torch.testing.assert_close(rd, torch.from_numpy(g["rays_d_" + tag]), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(vd, torch.from_numpy(g["viewdirs_" + tag]), rtol=1e-6, atol=1e-7)


def test_grid_modules_host_logic_matches_the_references_own_classes(tmp_path):
    """poseprobe_b200/grid.py (the drop-in DenseGrid / MaskGrid) against tests/golden/grid_modules.npz = the reference's
    own classes (lib/grid.py:38-89, :212-245) executed unmodified by tests/golden/make_golden_grid.py.  CPU part: the
    world->index affine buffers (bit-exact), the checkpoint -> occupancy recipe, `-=` and scale_volume_grid; the
    kernel-backed methods replay the same file in tests/test_parity_gpu.py."""
    import os
    import numpy as np
    import torch
    from poseprobe_b200.grid import DenseGrid, MaskGrid
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "grid_modules.npz"))
    lo, hi = g["mg_lo"].tolist(), g["mg_hi"].tolist()
    mg = MaskGrid(mask=torch.from_numpy(g["mg_mask"]), xyz_min=lo, xyz_max=hi)
    assert mg.mask.dtype == torch.bool and set(mg.state_dict()) == {"mask", "xyz2ijk_scale", "xyz2ijk_shift"}
    assert np.array_equal(mg.xyz2ijk_scale.numpy(), g["mg_scale"])
    assert np.array_equal(mg.xyz2ijk_shift.numpy(), g["mg_shift"])
    ckpt = {"model_state_dict": {"density.grid": torch.from_numpy(g["ck_density"]), "act_shift": torch.tensor(-2.0)},
            "model_kwargs": {"voxel_size_ratio": 0.75, "xyz_min": lo, "xyz_max": hi}}
    path = str(tmp_path / "coarse_last.tar")
    torch.save(ckpt, path)
    mc = MaskGrid(path=path, mask_cache_thres=0.3)
    assert mc.mask_cache_thres == 0.3 and np.array_equal(mc.mask.numpy(), g["ck_mask"])
    assert np.array_equal(mc.xyz2ijk_scale.numpy(), g["ck_scale"])
    assert np.array_equal(mc.xyz2ijk_shift.numpy(), g["ck_shift"])
    dg = DenseGrid(channels=3, world_size=[4, 6, 5], xyz_min=lo, xyz_max=hi)
    assert set(dg.state_dict()) == {"grid", "xyz_min", "xyz_max"} and tuple(dg.grid.shape) == (1, 3, 4, 6, 5)
    with torch.no_grad():
        dg.grid.copy_(torch.from_numpy(g["dg_grid"]))
    dg -= 0.25
    assert np.array_equal(dg.grid.detach().numpy(), g["dg_isub"])
    dg.scale_volume_grid([7, 5, 9])
    assert isinstance(dg.grid, torch.nn.Parameter) and dg.grid.requires_grad
    np.testing.assert_allclose(dg.grid.detach().numpy(), g["dg_scaled"], rtol=0, atol=1e-6)
    assert dg.get_dense_grid() is dg.grid


def test_voxurf_deform_oracle_matches_the_references_own_forward_and_gradients():
    """oracle/voxurf_cpu.py, implicit-deformation branch (``warp=``), against tests/golden/voxurf_deform.npz = the
    reference's OWN `Voxurf.forward(use_deform=True)` with its `DeformedImplicitField` and `deform_implicit_loss`,
    executed unmodified on the CPU by tests/golden/make_golden_deform.py: forward outputs (incl. the warp Jacobian
    `grad_deform`, the correction and `sdf_deform`), the four constraint terms, and the gradients of the combined loss
    with respect to the warp network, the colour network, k0 and the rays (second-order terms through the normal)."""
    import os
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import voxurf_cpu
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxurf_deform.npz"))
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    sdf = t("sdf")
    k0 = torch.from_numpy(g["k0"].astype(np.float32)).requires_grad_(True)
    net = nn.Sequential(nn.Linear(57, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 3))
    rgb_lin = [net[0], net[2], net[4]]
    with torch.no_grad():
        for i, lin in enumerate(rgb_lin):
            lin.weight.copy_(t("w%d" % i))
            lin.bias.copy_(t("b%d" % i))
    ww = [t("ww%d" % i).clone().requires_grad_(True) for i in range(5)]
    wb = [t("wb%d" % i).clone().requires_grad_(True) for i in range(5)]
    o, d = t("rays_o").clone().requires_grad_(True), t("rays_d").clone().requires_grad_(True)
    progress = float(g["gstep"]) / float(g["N_iters"])
    out = voxurf_cpu.forward(sdf, k0, net, o, d, t("viewdirs"), [-1, -1, -1], [1, 1, 1], 0.5, float(g["voxel_size"]),
                             float(g["s_val"]), bg=0.3, sdf_alpha=10.0, sdf_beta=0.35, posbase_pe=5, viewbase_pe=1,
                             barf_c2f=(0.1, 0.5), progress=progress, jitter=t("jitter"), differentiable_points=True,
                             warp=(ww, wb, float(g["out_range"])))
    assert out["weights"].shape[0] == g["ret_weights"].shape[0]
    for key in ("rgb_marched", "alphainv_cum", "cum_weights", "weights", "raw_alpha", "sdf_deform", "sdf_correct"):
        torch.testing.assert_close(out[key].detach().reshape(g["ret_" + key].shape), t("ret_" + key), rtol=0, atol=3e-5,
                                   msg=lambda m, key=key: "%s: %s" % (key, m))
    torch.testing.assert_close(out["gradient"].detach(), t("ret_gradient"), rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(out["grad_deform"].detach(), t("ret_grad_deform"), rtol=1e-4, atol=1e-5)
    terms = voxurf_cpu.deform_implicit_loss(out)
    for k, v in terms.items():
        assert abs(float(v.detach()) - float(g["term_" + k])) < 2e-5, k
    loss = F.mse_loss(out["rgb_marched"], t("target")) + terms["grad_constraint"] + 0.05 * (
        terms["grad_deform_constraint"] + terms["sdf_correct_constraint"] + terms["sdf_deform_constraint"])
    assert abs(float(loss.detach()) - float(g["loss"])) < 2e-5
    rgb_params = [p for lin in rgb_lin for p in (lin.weight, lin.bias)]
    grads = torch.autograd.grad(loss, ww + wb + rgb_params + [k0, o, d])

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-12))

    for i in range(5):
        assert rel(grads[i], t("g_ww%d" % i)) < 2e-3, ("warp weight", i, rel(grads[i], t("g_ww%d" % i)))
        assert rel(grads[5 + i], t("g_wb%d" % i)) < 2e-3, ("warp bias", i)
    for i in range(3):
        assert rel(grads[10 + 2 * i], t("g_w%d" % i)) < 2e-3, ("rgbnet weight", i)
        assert rel(grads[11 + 2 * i], t("g_b%d" % i)) < 2e-3, ("rgbnet bias", i)
    g_k0 = torch.zeros(k0.numel())
    g_k0[torch.from_numpy(g["g_k0_idx"]).long()] = t("g_k0_val")
    assert rel(grads[16].reshape(-1), g_k0) < 2e-3
    assert rel(grads[17], t("g_rays_o")) < 5e-3 and rel(grads[18], t("g_rays_d")) < 5e-3


def test_scene_branch_compositing_oracle_matches_the_references_own_code():
    """oracle/bgnerf_cpu.py against tests/golden/bgnerf_composite.npz = the reference's OWN `NeRF.composite`
    (frequency_nerf.py:283-343) and `Graph.sample_depth_from_pdf` (renderer.py:703-738) executed on the CPU
    (tests/golden/make_golden_bgnerf.py): every output, the gradients of a loss on rgb / depth / opacity / weights
    w.r.t. rgb_samples, density_samples and the ray, and the resampled depths (deterministic and shared-random grid)."""
    import os
    import numpy as np
    import torch
    from oracle import bgnerf_cpu
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bgnerf_composite.npz"))
    for tag in ("a", "b"):
        t = lambda k: torch.from_numpy(g["%s_%s" % (tag, k)])  # noqa: E731
        ray = t("ray").clone().requires_grad_(True)
        rgb_s = t("rgb_samples").clone().requires_grad_(True)
        dens = t("density_samples").clone().requires_grad_(True)
        out = bgnerf_cpu.composite(ray, rgb_s, dens, t("depth_samples"), setbg_opaque=bool(g[tag + "_opaque"]))
        for k in ("rgb", "rgb_var", "depth", "depth_var", "opacity", "weights", "all_cumulated"):
            torch.testing.assert_close(out[k].detach(), t("out_" + k), rtol=1e-5, atol=1e-6, msg=lambda m, k=k: k + m)
        loss = (out["rgb"] * t("w_rgb")).sum() + (out["depth"] * t("w_depth")).sum() + \
            (out["opacity"] * t("w_opacity")).sum() + (out["weights"] * t("w_weights")).sum()
        grads = torch.autograd.grad(loss, [rgb_s, dens, ray])
        for got, k in zip(grads, ("g_rgb_samples", "g_density_samples", "g_ray")):
            torch.testing.assert_close(got, t(k), rtol=1e-4, atol=1e-6)
        w = t("out_weights")[..., 0]
        N = w.shape[-1]
        rng = [float(v) for v in g[tag + "_depth_range"]]
        Nf = g[tag + "_fine_det"].shape[2]
        det, _ = bgnerf_cpu.sample_depth_from_pdf(w, N, Nf, rng, torch.linspace(0, 1, Nf + 1))
        assert torch.equal(det, t("fine_det"))
        rnd, _ = bgnerf_cpu.sample_depth_from_pdf(w, N, Nf, rng, t("fine_grid"))
        assert torch.equal(rnd, t("fine_rand"))


def test_dense_compositing_oracle_matches_the_references_own_dvgo_forward():
    """oracle.get_ray_marching_ray (lib/dvgo_ori.py:478-485) against tests/golden/dvgo_module.npz = the reference's own
    `DirectVoxGO.forward` (make_golden_dvgo_module.py): weights and alphainv_cum recomputed from the stored raw_alpha,
    for both stored configurations (the GPU tier replays the whole module against the same file)."""
    import os
    import numpy as np
    import torch
    from oracle import oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dvgo_module.npz"))
    for tag in ("a", "b"):
        alpha = torch.from_numpy(g[tag + "_out_raw_alpha"])
        w, c = oracle.get_ray_marching_ray(alpha)
        torch.testing.assert_close(w, torch.from_numpy(g[tag + "_out_weights"]), rtol=1e-6, atol=1e-8)
        torch.testing.assert_close(c, torch.from_numpy(g[tag + "_out_alphainv_cum"]), rtol=1e-6, atol=1e-8)
        assert torch.equal(w > 1e-4, torch.from_numpy(g[tag + "_out_mask"]))
