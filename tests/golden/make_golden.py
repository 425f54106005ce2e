"""Generates tests/golden/*.npz from the REFERENCE's own kernels (oracle/_ref, built unmodified from
/root/reference/lib/cuda by oracle/build_ref.py) on a CUDA device.

Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy the .npz files into tests/golden/ and commit them.  Inputs are stored next to the outputs so
the CPU-side replay (tests/test_oracle_golden.py) does not depend on RNG reproducibility.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from tests import cases  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ru = build_ref.load_ref("render_utils_cuda")
    tv = build_ref.load_ref("total_variation_cuda")
    ad = build_ref.load_ref("adam_upd_cuda")
    ub = build_ref.load_ref("ub360_utils_cuda")
    assert ru and tv and ad and ub, "oracle/_ref is not built"
    dev = "cuda"
    g = torch.Generator().manual_seed(2024)

    # ---- sampling -------------------------------------------------------------------------------
    o, d = cases.rays_with_edge_cases(48, seed=1)
    mn, mx = cases.bbox_aniso()
    near, far, stepdist = 0.2, 1e9, 0.05
    outs = ru.sample_pts_on_rays(o.to(dev), d.to(dev), mn.to(dev), mx.to(dev), near, far, stepdist)
    names = ["rays_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max", "rays_start", "rays_dir"]
    rec = dict(rays_o=npy(o), rays_d=npy(d), xyz_min=npy(mn), xyz_max=npy(mx),
               near=np.float32(near), far=np.float32(far), stepdist=np.float32(stepdist))
    rec.update({k: npy(v) for k, v in zip(names, outs)})
    tm = ru.infer_t_minmax(o.to(dev), d.to(dev), mn.to(dev), mx.to(dev), near, far)
    rec["infer_n_samples"] = npy(ru.infer_n_samples(d.to(dev), tm[0], tm[1], stepdist))
    S = 9
    p, m = ru.sample_ndc_pts_on_rays(o.to(dev), d.to(dev), mn.to(dev), mx.to(dev), S)
    rec.update(ndc_S=np.int32(S), ndc_pts=npy(p), ndc_mask=npy(m))
    t_max = torch.rand(48, generator=g) + 0.5
    rec.update(bg_t_max=npy(t_max), bg_preserve=np.float32(0.5),
               bg_pts=npy(ru.sample_bg_pts_on_rays(o.to(dev), d.to(dev), t_max.to(dev), 0.5, S)))
    np.savez_compressed(os.path.join(out_dir, "sampling.npz"), **rec)

    # ---- maskcache ------------------------------------------------------------------------------
    world = torch.rand(7, 9, 5, generator=g) > 0.5
    xyz = torch.rand(200, 3, generator=g) * 2.6 - 1.3
    xyz[0], xyz[1], xyz[2] = mn, mx, (mn + mx) / 2
    scale = (torch.tensor(world.shape, dtype=torch.float32) - 1) / (mx - mn)
    shift = -mn * scale
    out = ru.maskcache_lookup(world.to(dev), xyz.to(dev), scale.to(dev), shift.to(dev))
    np.savez_compressed(os.path.join(out_dir, "maskcache.npz"), world=npy(world), xyz=npy(xyz), scale=npy(scale),
                        shift=npy(shift), out=npy(out))

    # ---- raw2alpha ------------------------------------------------------------------------------
    dens = torch.randn(256, generator=g) * 4
    dens[:4] = torch.tensor([100.0, -100.0, 0.0, 88.0])
    gb = torch.randn(256, generator=g)
    iv = torch.rand(256, generator=g) + 0.1
    shift_, interval = -4.595, 0.5
    e, a = ru.raw2alpha(dens.to(dev), shift_, interval)
    gr = ru.raw2alpha_backward(e, gb.to(dev), interval)
    en, an = ru.raw2alpha_nonuni(dens.to(dev), shift_, iv.to(dev))
    grn = ru.raw2alpha_nonuni_backward(en, gb.to(dev), iv.to(dev))
    np.savez_compressed(os.path.join(out_dir, "raw2alpha.npz"), density=npy(dens), grad_back=npy(gb), interval_pts=npy(iv),
                        shift=np.float32(shift_), interval=np.float32(interval), exp_d=npy(e), alpha=npy(a),
                        grad=npy(gr), alpha_nonuni=npy(an), grad_nonuni=npy(grn))

    # ---- alpha2weight ---------------------------------------------------------------------------
    alpha, ray_id, _ = cases.alpha_and_ray_id(40, 50, seed=77, p_big=0.15)
    n_rays = 40
    w, T, last, i_s, i_e = ru.alpha2weight(alpha.to(dev), ray_id.to(dev), n_rays)
    gw = torch.randn(alpha.numel(), generator=g)
    gl = torch.randn(n_rays, generator=g)
    gr = ru.alpha2weight_backward(alpha.to(dev), w, T, last, i_s, i_e, n_rays, gw.to(dev), gl.to(dev))
    np.savez_compressed(os.path.join(out_dir, "alpha2weight.npz"), alpha=npy(alpha), ray_id=npy(ray_id),
                        n_rays=np.int64(n_rays), weight=npy(w), T=npy(T), alphainv_last=npy(last), i_start=npy(i_s),
                        i_end=npy(i_e), grad_weights=npy(gw), grad_last=npy(gl), grad=npy(gr))

    # ---- total variation ------------------------------------------------------------------------
    shape = (1, 2, 6, 7, 8)
    param = torch.randn(shape, generator=g) * 2
    grad = torch.randn(shape, generator=g)
    grad[torch.rand(shape, generator=g) < 0.6] = 0
    mask = (torch.rand(shape, generator=g) > 0.3).float()
    rec = dict(param=npy(param), grad=npy(grad), mask=npy(mask), w=np.array([0.3, 0.5, 0.7], np.float32))
    for dense in (0, 1):
        a_ = grad.clone().to(dev)
        tv.total_variation_add_grad(param.to(dev), a_, 0.3, 0.5, 0.7, bool(dense))
        rec["old_dense%d" % dense] = npy(a_)
        b_ = grad.clone().to(dev)
        tv.total_variation_add_grad_new(param.to(dev), b_, mask.to(dev), 0.3, 0.5, 0.7, bool(dense))
        rec["new_dense%d" % dense] = npy(b_)
    np.savez_compressed(os.path.join(out_dir, "total_variation.npz"), **rec)

    # ---- adam -----------------------------------------------------------------------------------
    n = 257
    p0 = torch.randn(n, generator=g)
    gr0 = torch.randn(n, generator=g)
    gr0[torch.rand(n, generator=g) < 0.5] = 0
    m0 = torch.randn(n, generator=g) * 0.1
    v0 = torch.rand(n, generator=g) * 0.01
    plr = torch.rand(n, generator=g)
    hp = dict(step=7, beta1=0.9, beta2=0.99, lr=0.1, eps=1e-8)
    rec = dict(param=npy(p0), grad=npy(gr0), exp_avg=npy(m0), exp_avg_sq=npy(v0), perlr=npy(plr),
               hp=np.array([hp["step"], hp["beta1"], hp["beta2"], hp["lr"], hp["eps"]], np.float64))
    for mode, fn in enumerate([ad.adam_upd, ad.masked_adam_upd, ad.adam_upd_with_perlr]):
        p_, m_, v_ = p0.clone().to(dev), m0.clone().to(dev), v0.clone().to(dev)
        args = (p_, gr0.to(dev), m_, v_) + ((plr.to(dev),) if mode == 2 else ())
        fn(*args, hp["step"], hp["beta1"], hp["beta2"], hp["lr"], hp["eps"])
        rec["param_%d" % mode], rec["exp_avg_%d" % mode], rec["exp_avg_sq_%d" % mode] = npy(p_), npy(m_), npy(v_)
    np.savez_compressed(os.path.join(out_dir, "adam.npz"), **rec)

    # ---- cumdist_thres --------------------------------------------------------------------------
    dist = torch.rand(9, 40, generator=g) * 0.1
    mask = ub.cumdist_thres(dist.to(dev), 0.33)
    np.savez_compressed(os.path.join(out_dir, "cumdist.npz"), dist=npy(dist), thres=np.float32(0.33), mask=npy(mask))
    torch.cuda.synchronize()
    print("golden vectors written to", out_dir, "from", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__)))
